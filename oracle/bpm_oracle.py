"""bpm_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of the reference's beam-propagation pre-processor `bpm.py` (the script that
writes `bessel-normal.dat`, the 512x512 fp64 intensity the `image` source reads).  Pinned against
the reference itself: tests/golden/bpm_v1.npz is a fingerprint of the file the unmodified script
writes in this container (tests/golden/make_bpm_golden.py).

What the script computes once its commented-out blocks are set aside (bpm.py:84-151, :203-205):
  * a ring-shaped field exp(-((r - 1612)/300)^2) on a 512x512 grid of 5000 um       (:100-104,:116-121)
  * nz/10 = 100 split-step free-space propagations, each ifft2(fft2(e) * exp(i * arg)) with
    arg = -dz (k1^2 + k2^2) / (2k) on the folded frequency grid                     (:57-80,:106-115,:126-127)
  * a thin-lens phase exp(-i k r^2 / 2R) (it does not change the intensity)          (:136)
  * out = |e^T|^2                                                                    (:203-204)
"""
import numpy as np

DEFAULTS = dict(w0=582.0 * 4, wavelength=0.785, axicon_angle=5.0, n=1.45, xymax=5000.0, nxy=512, nz=1000,
                ring_radius=1612.0, ring_width=300.0, steps=None)


def bessel_intensity(**kw):
    p = dict(DEFAULTS)
    p.update(kw)
    k = 2.0 * np.pi / p["wavelength"]                                   # bpm.py:86
    k_r = k * (p["n"] - 1.0) * p["axicon_angle"] * np.pi / 360.0         # :90
    nxy, nz = int(p["nxy"]), int(p["nz"])
    zmax = p["w0"] * (k / k_r)                                           # :96
    L = 3.0 * zmax                                                       # :97
    R = L
    dz = L / nz                                                          # :100
    dx = p["xymax"] / nxy                                                # :102
    dk = (2.0 * np.pi / dx) / nxy                                        # :103-104
    nmid = nxy // 2
    v = np.arange(0, nxy)
    x, y = np.meshgrid(v, v)
    x = x * dx - p["xymax"] / 2                                          # :109-110
    y = y * dx - p["xymax"] / 2
    fold = v > nmid                                                      # :112-113
    v = np.where(fold, nxy - v, v) * dk                                  # :114
    k2, k1 = np.meshgrid(v, v)
    arg = -dz * (k1 ** 2 + k2 ** 2) / (2.0 * k)                          # :117
    r = np.sqrt(x ** 2 + y ** 2)
    e = np.exp(-(r - p["ring_radius"]) ** 2 / p["ring_width"] ** 2).astype(np.complex128)   # :120-121
    steps = int(nz / 10) if p["steps"] is None else int(p["steps"])     # :126
    freq = np.exp(1j * arg)
    for _ in range(steps):
        e = np.fft.ifft2(np.fft.fft2(e) * freq)                          # :78-80
    e = e * np.exp(-1j * k * r ** 2 / (2.0 * R))                         # :136
    return np.abs(e.T) ** 2                                              # :203
