"""Known answers the oracle did not author: numbers held by the REFERENCE's own
files and by closed-form optics.

  * every shipped plano-convex file states its focal length f and the back focal length fb (lines 4-5,
    e.g. reference res/planoConvex-f39.9mm.params:4-5), every doublet file f and fb (lines 7-8,
    res/achromaticDoublet-f50.0mm.params:7-8) -- catalogue values at the design wavelength.  A paraxial
    ray traced through the lens must focus where those lines say, and exactly where the paraxial (ABCD)
    formulas put it for the radii, thicknesses and Sellmeier indices of the same file;
  * the hemispherical (cosine-weighted) average of the unpolarised Fresnel reflectance has a closed
    form (Walsh 1926); the Brewster angle has R = ((n^2-1)/(n^2+1))^2 / 2.

The same rays go through the CUDA path in the -m gpu twin of each test.
"""
import glob
import math
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import oracle_lib

RES = oracle_lib.RES
PLANOS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(RES, "planoConvex*.params")))
DOUBLETS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(RES, "achromaticDoublet*.params")))
LAMBDA_D = 587.6e-9      # N-BK7 plano-convex singlets are specified at the d line
LAMBDA_B = 855e-9        # the "-B" (NIR) achromats at 855 nm
H = 1e-5                 # ray height: spherical aberration ~ (h / R)^2 ~ 1e-7 relative
# reference res/achromaticDoublet-f75.0mm.params states f = 75.0 mm, fb = 69.9 mm (lines 7-8) but its element-1
# Sellmeier lines (10-15: B1 = 1.5851495 ...) with R1 = 36.90, R2 = 42.17, R3 = 417.8 mm give a 59.4 mm lens:
# the data of the file, which is what the reference traces, not its catalogue line
DATA_QUIRK = {"achromaticDoublet-f75.0mm.params": 59.4258e-3}


def tokens(name):
    return [float(l.split()[0].replace("d", "e").replace("D", "e")) for l in open(os.path.join(RES, name)) if l.strip()]


def sellmeier(lam_m, B, C):
    l2 = (lam_m * 1e6) ** 2
    return math.sqrt(1.0 + sum(b * l2 / (l2 - c) for b, c in zip(B, C)))


def parallel_rays(scene, stop, tracer):
    """two rays parallel to the axis at heights +-H through L2 (stop 3) or L2 + L3 (stop 4), every
    interface transmitting (constant draw 0.999 > any reflectance here) -> (pos, dir) at the last surface"""
    job = abi.default_job(2, use_bottle=False, uniform_override=0.999, stop_after=stop)
    p = np.array([[H, -H], [0.0, 0.0], [0.0, 0.0]])
    d = np.array([[0.0, 0.0], [0.0, 0.0], [1.0, 1.0]])
    out = tracer(job, scene, 2, p, d)
    assert list(out["status"]) == [25, 25], out["status"]        # ORT_ST_STOPPED: alive at the stop
    return out["pos"], out["dir"]


def focus(pos, d):
    """paraxial focus of the exit ray: (z where it crosses the axis, effective focal length h / -slope)"""
    slope = d[0, 0] / d[2, 0]
    assert slope < 0                                              # converging
    assert abs(d[0, 1] + d[0, 0]) < 1e-15 and abs(pos[0, 1] + pos[0, 0]) < 1e-15   # mirror ray: mirror image
    return pos[2, 0] - pos[0, 0] / slope, H / -slope


def check_plano(name, tracer):
    t = tokens(name)
    th, R, f_file, fb_file = t[0], t[1], t[3], t[4]
    n = sellmeier(LAMBDA_D, t[6:9], t[9:12])
    scene = oracle_lib.make_scene(l2=name, lens_wavelength=LAMBDA_D)
    assert abs(scene.L2.n2 - n) < 1e-14
    pos, d = parallel_rays(scene, abi.STOP_L2 if hasattr(abi, "STOP_L2") else 3, tracer)
    zc, efl = focus(pos, d)
    vertex = scene.L2.centre[2] + R                               # the curved vertex, = fb + thickness
    assert abs(vertex - (scene.L2.fb + th)) < 1e-15
    bfd = zc - vertex
    # collimated light enters the flat face undeviated: one refracting surface, f = R / (n - 1), measured from it
    f_paraxial = R / (n - 1.0)
    assert abs(efl / f_paraxial - 1.0) < 1e-6 and abs(bfd / f_paraxial - 1.0) < 1e-6, (efl, bfd, f_paraxial)
    # ... which is the catalogue focal length of the file, to its rounding (0.1 mm) and the 1 % catalogue tolerance
    assert abs(efl - f_file) < 0.2e-3 and abs(efl / f_file - 1.0) < 0.01, (name, efl, f_file)
    # the file's fb is the focal distance on the FLAT side, f - t / n (where the reference puts the point source)
    assert abs((f_paraxial - th / n) - fb_file) < 0.15e-3, (name, f_paraxial - th / n, fb_file)
    return efl, bfd


def doublet_paraxial(t, lam):
    """ABCD matrices of the three surfaces and two gaps: -> (effective focal length, back focal distance)"""
    t1, t2, R1, R2, R3 = t[0], t[1], t[2], t[3], t[4]
    n1 = t[8]
    n2 = sellmeier(lam, t[9:12], t[12:15])
    n3 = sellmeier(lam, t[15:18], t[18:21])
    # reduced-angle convention: state (y, n u); surface: nu' = nu - y (n' - n) / R; gap: y' = y + (nu) t / n.
    # The reference's doublet: convex R1 first (centre behind it: +R1), then R2 and R3 with centres in FRONT
    # (src/lens.f90:122-124: c2 = ... + t1 - R2, c3 = ... + t - R3): both concave towards the source: -R2, -R3.
    y, nu = 1.0, 0.0
    nu -= y * (n2 - n1) / R1
    y += nu * t1 / n2
    nu -= y * (n3 - n2) / (-R2)
    y += nu * t2 / n3
    nu -= y * (n1 - n3) / (-R3)
    return -1.0 / (nu / n1), -y / (nu / n1), n2, n3


def check_doublet(name, tracer):
    t = tokens(name)
    f_file, fb_file = t[6], t[7]
    efl_m, bfd_m, n2, n3 = doublet_paraxial(t, LAMBDA_B)
    scene = oracle_lib.make_scene(l3=name, lens_wavelength=LAMBDA_B)
    assert abs(scene.L3.n2 - n2) < 1e-14 and abs(scene.L3.n3 - n3) < 1e-14
    scene.L2.n2 = scene.L2.n1                                     # L2 becomes a window of air: L3 sees collimated light
    pos, d = parallel_rays(scene, 4, tracer)
    zc, efl = focus(pos, d)
    last_vertex = scene.L3.centre3[2] + t[4]                      # c3 = vertex - R3
    bfd = zc - last_vertex
    assert abs(efl / efl_m - 1.0) < 1e-6 and abs(bfd / bfd_m - 1.0) < 1e-6, (efl, efl_m, bfd, bfd_m)
    if name in DATA_QUIRK:
        # the reference's own file is inconsistent: its radii, thicknesses and Sellmeier lines give this focal
        # length, not the one on its line 7 (see DATA_QUIRK); ray trace and paraxial matrix agree on it
        assert abs(efl - DATA_QUIRK[name]) < 1e-5, (name, efl)
        return efl, bfd
    # the catalogue lines of the file (f +-1 %, fb to its rounding)
    assert abs(efl / f_file - 1.0) < 0.01, (name, efl, f_file)
    assert abs(bfd - fb_file) < 0.01 * f_file, (name, bfd, fb_file)
    return efl, bfd


@pytest.mark.parametrize("name", PLANOS)
def test_plano_convex_focal_lengths_oracle(orc, name):
    check_plano(name, orc.trace_rays)


@pytest.mark.parametrize("name", DOUBLETS)
def test_doublet_focal_lengths_oracle(orc, name):
    check_doublet(name, orc.trace_rays)


@pytest.mark.gpu
def test_focal_lengths_cuda(ort, orc):
    for name in PLANOS:
        a, b = check_plano(name, ort.trace_rays), check_plano(name, orc.trace_rays)
        assert np.allclose(a, b, rtol=1e-9, atol=0)
    for name in DOUBLETS:
        a, b = check_doublet(name, ort.trace_rays), check_doublet(name, orc.trace_rays)
        assert np.allclose(a, b, rtol=1e-9, atol=0)


def walsh(n):
    """hemispherical average of the unpolarised Fresnel reflectance, light entering index n from 1 (Walsh 1926)"""
    return (0.5 + (n - 1) * (3 * n + 1) / (6 * (n + 1) ** 2) + (n ** 2 * (n ** 2 - 1) ** 2 / (n ** 2 + 1) ** 3) * math.log((n - 1) / (n + 1))
            - 2 * n ** 3 * (n ** 2 + 2 * n - 1) / ((n ** 2 + 1) * (n ** 4 - 1)) + (8 * n ** 4 * (n ** 4 + 1) / ((n ** 2 + 1) * (n ** 4 - 1) ** 2)) * math.log(n))


@pytest.mark.parametrize("n", [1.33, 1.5, 1.511079564908228, 1.785335731036205])
def test_fresnel_hemispherical_integral_and_brewster(orc, n):
    """2 int_0^1 R(c) c dc over the oracle's fresnel() (src/surfaces.f90:336-372) against Walsh's closed form"""
    x, w = np.polynomial.legendre.leggauss(400)
    c = 0.5 * (x + 1.0)
    L = orc.lib()
    tot = 0.0
    for ci, wi in zip(c, w):
        I = oracle_lib.v3((math.sqrt(1 - ci * ci), 0.0, ci))
        N = oracle_lib.v3((0.0, 0.0, -1.0))
        tot += 0.5 * wi * 2.0 * ci * L.orc_fresnel(I, N, 1.0, n)
    assert abs(tot - walsh(n)) < 1e-11, (tot, walsh(n))
    tb = math.atan(n)                                             # Brewster: r_p = 0
    I = oracle_lib.v3((math.sin(tb), 0.0, math.cos(tb)))
    assert abs(L.orc_fresnel(I, oracle_lib.v3((0.0, 0.0, -1.0)), 1.0, n) - 0.5 * ((n * n - 1) / (n * n + 1)) ** 2) < 1e-15
