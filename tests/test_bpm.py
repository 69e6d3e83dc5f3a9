"""Beam-propagation pre-processor: ort_bpm_bessel against the numpy
restatement of the reference's bpm.py, which is itself pinned by a fingerprint of the file the
UNMODIFIED reference script writes (tests/golden/bpm_v1.npz, made by make_bpm_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bpm_oracle  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "bpm_v1.npz"))
TOL = 1e-9   # of the peak intensity; FFT libraries differ in rounding (observed 4e-14 after 100 steps)


def _check_fingerprint(img, tol, argmax=False):
    peak = float(GOLD["peak"])
    assert img.shape == (512, 512)
    assert np.abs(img[256, :] - GOLD["row"]).max() <= tol * peak
    assert np.abs(img[:, 256] - GOLD["col"]).max() <= tol * peak
    assert np.abs(img[::8, ::8] - GOLD["sub"]).max() <= tol * peak
    assert abs(img.sum() - float(GOLD["total"])) <= tol * float(GOLD["total"])
    # the ring has mirror-image maxima that differ in the last bits: only a bit-faithful
    # restatement can be asked for the same argmax
    if argmax:
        assert tuple(np.unravel_index(np.argmax(img), img.shape)) == tuple(GOLD["argmax"])
    assert abs(img.max() - peak) <= tol * peak


def test_oracle_reproduces_the_reference_script():
    _check_fingerprint(bpm_oracle.bessel_intensity(), 1e-14, argmax=True)


def test_bpm_defaults_are_the_scripts_constants(ortlib):
    p = ortlib.bpm_defaults()
    d = bpm_oracle.DEFAULTS
    assert (p.w0, p.wavelength, p.axicon_deg, p.n_axicon, p.xymax, p.ring_radius, p.ring_width) == \
        (d["w0"], d["wavelength"], d["axicon_angle"], d["n"], d["xymax"], d["ring_radius"], d["ring_width"])
    assert (p.nxy, p.nz) == (d["nxy"], d["nz"]) and p.steps < 0
    import ctypes as C
    from opticalraytrace_b200 import abi
    assert ortlib.load().ort_bpm_struct_size() == C.sizeof(abi.Bpm)


def test_bpm_needs_an_initialised_library(ortlib):
    from opticalraytrace_b200.lib import OrtError
    with pytest.raises(OrtError):
        ortlib.bpm_bessel()


@pytest.mark.gpu
def test_cuda_bpm_matches_reference_output(ort):
    img = ort.bpm_bessel()
    _check_fingerprint(img, TOL)
    ref = bpm_oracle.bessel_intensity()
    assert np.abs(img - ref).max() <= TOL * ref.max()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(nxy=256, steps=10), dict(nxy=128, steps=0), dict(steps=37, ring_radius=900.0),
                                dict(nxy=384, nz=500, wavelength=0.633, axicon_angle=2.0, n=1.5, ring_width=120.0)])
def test_cuda_bpm_parameter_cases(ort, kw):
    p = ort.bpm_defaults()
    names = dict(axicon_angle="axicon_deg", n="n_axicon")
    for k, v in kw.items():
        setattr(p, names.get(k, k), v)
    img = ort.bpm_bessel(p)
    ref = bpm_oracle.bessel_intensity(**kw)
    assert img.shape == ref.shape
    assert np.abs(img - ref).max() <= TOL * ref.max()


@pytest.mark.gpu
def test_bpm_feeds_the_image_source(ort, orc, tmp_path):
    """bpm -> bessel-normal.dat -> init_emit_image -> point loop with source_type = image"""
    from opticalraytrace_b200 import abi
    from tests import cases
    path = str(tmp_path / "bessel-normal.dat")
    ort.bpm_write_file(path)
    assert os.path.getsize(path) == 512 * 512 * 8
    n = 60_000
    budget = ort.load_image_source(path, n)
    assert np.array_equal(budget, orc.load_image_source(path, n))
    scene = cases.scene_for(orc, cases.C2, 2)
    try:
        ort.set_image_source(budget)
        orc.set_image_source(budget)
        job = abi.default_job(2, n, source="image")
        img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
        oimg, olost, ohist = orc.trace(job, scene)
        assert np.array_equal(img, oimg) and np.array_equal(hist, ohist)
    finally:
        ort.set_image_source(None)
        orc.set_image_source(None)
