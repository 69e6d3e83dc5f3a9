"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the
pinned oracle): the oracle still reproduces them, the host-compiled product math matches them,
and -- on the GPU box, where /root/reference does not exist -- so does the CUDA path."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import rel_err

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = cases.RAY_CASES + cases.SCATTER_CASES + cases.SOURCE_CASES
IDS = [c[0] for c in ALL]
NRAYS, NIMG = 96, 200_000


@pytest.fixture(scope="module")
def gold():
    return (np.load(os.path.join(HERE, "rays_v2.npz")), np.load(os.path.join(HERE, "images_v2.npz")),
            np.load(os.path.join(HERE, "uniforms_v2.npz")))


def _dense(g, cid):
    img = np.zeros(abi.ORT_IMG_BINS, dtype=np.uint64)
    img[g["%s/idx" % cid]] = g["%s/cnt" % cid].astype(np.uint64)
    return img.reshape(abi.ORT_IMG_N, abi.ORT_IMG_N)


def _check_rays(r, g, cid, tol):
    assert np.array_equal(r["status"], g["%s/status" % cid])
    assert np.array_equal(r["bin"], g["%s/bin" % cid])
    e = np.maximum(rel_err(g["%s/pos" % cid], r["pos"]), rel_err(g["%s/dir" % cid], r["dir"]))
    assert np.nanmax(e) < tol, np.nanmax(e)


@pytest.mark.parametrize("cid,files,phase,kw", ALL, ids=IDS)
def test_oracle_reproduces_golden(orc, gold, cid, files, phase, kw):
    scene = cases.scene_for(orc, files, phase, kw)
    r = orc.trace_rays(abi.default_job(phase, **kw), scene, NRAYS)
    _check_rays(r, gold[0], cid, 1e-15)
    img, lost, hist = orc.trace(abi.default_job(phase, NIMG, **kw), scene)
    assert np.array_equal(img[0], _dense(gold[1], cid))
    assert np.array_equal(hist[0], gold[1]["%s/hist" % cid])
    assert np.array_equal(lost, gold[1]["%s/lost" % cid])


def test_uniforms_golden(orc, gold):
    for key in gold[2].files:
        p, r = key.split("/")
        assert np.array_equal(orc.uniforms(123456789, int(p[1:]), int(r[1:]), 0, 24), gold[2][key])


@pytest.mark.parametrize("cid,files,phase,kw", ALL, ids=IDS)
def test_host_math_matches_golden(orc, harness, gold, cid, files, phase, kw):
    scene = cases.scene_for(orc, files, phase, kw)
    r = harness(abi.default_job(phase, **kw), scene, NRAYS)
    _check_rays(r, gold[0], cid, 1e-6 if "scatter" in cid else 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("cid,files,phase,kw", ALL, ids=IDS)
def test_cuda_matches_golden(ort, gold, cid, files, phase, kw):
    """Does not touch the oracle: scenes come from the product's own readers."""
    res = os.path.join(os.path.dirname(HERE), "..", "res")
    st = ort.make_settings(*files, source_type=kw.get("source", "point"))
    scene, _ = ort.build_scene(st, res, 843e-9 if phase == 2 else None)
    r = ort.trace_rays(abi.default_job(phase, **kw), scene, NRAYS)
    _check_rays(r, gold[0], cid, 1e-6 if "scatter" in cid else 1e-9)
    img, lost, hist, _ = ort.trace(abi.default_job(phase, NIMG, **kw), scene, allow_trap=True)
    assert np.array_equal(img[0], _dense(gold[1], cid))
    assert np.array_equal(hist[0], gold[1]["%s/hist" % cid])
    assert np.array_equal(lost, gold[1]["%s/lost" % cid])


@pytest.mark.gpu
def test_cuda_uniforms_golden(ort, gold):
    for key in gold[2].files:
        p, r = key.split("/")
        assert np.array_equal(ort.uniforms(123456789, int(p[1:]), int(r[1:]), 0, 24), gold[2][key])
