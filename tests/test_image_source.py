"""source_type = image: init_emit_image + emit_image + emit
(reference src/sourceMod.f90:303-408).  The input file is what bpm.py writes: 512 x 512 float64,
C order.  None is shipped, so the tests synthesise a Bessel-like ring pattern."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import rel_err


def _bessel_like(path):
    y, x = np.mgrid[0:512, 0:512]
    r = np.hypot(x - 255.5, y - 255.5)
    img = np.exp(-((r - 120.0) / 9.0) ** 2) + 0.3 * np.exp(-(r / 25.0) ** 2) * (x > 255)  # not symmetric
    img.astype(np.float64).tofile(path)
    return img


@pytest.fixture()
def src_file(tmp_path):
    p = str(tmp_path / "bessel-test.dat")
    return p, _bessel_like(p)


def test_budget_matches_oracle_and_intensity(ortlib, orc, src_file):
    path, img = src_file
    n = 3_000_000
    b = ortlib.load_image_source(path, n)
    bo = orc.load_image_source(path, n)
    assert np.array_equal(b, bo)
    assert abs(int(b.sum()) - n) < 600                         # random rounding of 262144 fractions
    # budget[(j-1)*512 + (i-1)] = imgin(i,j) with imgout(i,j) = file row i, column j
    B = b.reshape(512, 512).T                                  # B[i, j]
    want = n * img / img.sum()
    assert np.abs(B - want).max() <= 1.0
    with pytest.raises(Exception):
        ortlib.load_image_source(path + ".missing", n)


def test_host_math_image_source(orc, harness, src_file):
    path, _ = src_file
    n = 200_000
    budget = orc.load_image_source(path, n)
    orc.set_image_source(budget)
    harness.set_image_source(budget)
    try:
        scene = cases.scene_for(orc, cases.C1, 2)
        for stop in (1, 0):
            job = abi.default_job(2, source="image", stop_after=stop)
            a = orc.trace_rays(job, scene, n)
            b = harness(job, scene, n)
            assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])
            e = np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"]))
            assert np.nanmax(e) < 1e-9
        # source positions follow the intensity pattern: pixel pitch 5 mm / 512, centred
        job = abi.default_job(2, source="image", stop_after=1)
        a = orc.trace_rays(job, scene, n)
        ok = a["status"] == abi.ST_STOPPED
        assert ok.sum() == min(n, int(budget.sum()))
        px = np.floor((a["pos"][0, ok] + 2500e-6) / (5000e-6 / 512)).astype(int)   # first index j
        py = np.floor((a["pos"][1, ok] + 2500e-6) / (5000e-6 / 512)).astype(int)   # second index i
        counts = np.zeros((512, 512), dtype=np.int64)
        np.add.at(counts, (py, px), 1)
        # rays are dealt out in scan order: the first sum(budget[:k]) rays fill the first k pixels
        flat = counts.ravel()                                   # [i*512 + j] = scan order
        done = np.cumsum(budget.astype(np.int64)) <= ok.sum()
        assert np.array_equal(flat[done], budget[done])
    finally:
        orc.set_image_source(None)
        harness.set_image_source(None)


@pytest.mark.gpu
def test_cuda_image_source(ort, orc, src_file):
    path, _ = src_file
    n = 500_003
    budget = ort.load_image_source(path, n)
    assert np.array_equal(budget, orc.load_image_source(path, n))
    from opticalraytrace_b200.lib import OrtError
    scene = cases.scene_for(orc, cases.C1, 2)
    ort.set_image_source(None)
    with pytest.raises(OrtError, match="ort_set_image_source"):
        ort.trace(abi.default_job(2, 100, source="image"), scene)
    ort.set_image_source(budget)
    orc.set_image_source(budget)
    try:
        job = abi.default_job(2, source="image")
        a = orc.trace_rays(job, scene, 200_000)
        b = ort.trace_rays(job, scene, 200_000)
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])
        e = np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"]))
        assert np.nanmax(e) < 1e-9
        job = abi.default_job(2, n, source="image")
        img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
        oimg, olost, ohist = orc.trace(job, scene)
        assert np.array_equal(hist, ohist) and np.array_equal(img, oimg) and np.array_equal(lost, olost)
        assert hist[0, 26] == max(0, n - int(budget.sum()))     # rays beyond the budget
        ring = ort.trace(abi.default_job(1, n, source="image"), cases.scene_for(orc, cases.C1, 1))
        plain = ort.trace(abi.default_job(1, n), cases.scene_for(orc, cases.C1, 1))
        assert np.array_equal(ring[0], plain[0])                # the ring loop still uses ring()
    finally:
        ort.set_image_source(None)
        orc.set_image_source(None)
