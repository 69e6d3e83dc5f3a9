"""The product's reformulated per-ray arithmetic (ort_optics.cuh: one-division quadratics, fused
Fresnel + Snell, hoisted normals, squared-radius tests) compiled for the HOST by the test-only
harness and compared with the oracle -- the no-GPU half of the parity argument; the -m gpu
tests repeat it with the same header compiled for sm_100a through the C-ABI."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import both_err, rel_err, scatter_conditioning

TOL = 1e-9          # north_star: 1e-9 relative, checked on the vector AND per component (conftest.comp_err)


_CASES = cases.RAY_CASES + cases.SOURCE_CASES


@pytest.mark.parametrize("cid,files,phase,kw", _CASES, ids=[c[0] for c in _CASES])
def test_reformulated_math_matches_oracle(orc, harness, cid, files, phase, kw):
    n = 100_000
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a = orc.trace_rays(job, scene, n)
    b = harness(job, scene, n)
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["bin"], b["bin"])
    e, ec = both_err(a, b)
    assert np.nanmax(e) < TOL and np.nanmax(ec) < TOL, (np.nanmax(e), np.nanmax(ec))


@pytest.mark.parametrize("cid,files,phase,kw", cases.SCATTER_CASES, ids=[c[0] for c in cases.SCATTER_CASES])
def test_scatter_math_matches_oracle(orc, harness, cid, files, phase, kw):
    n = 100_000
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a, resp, stable = scatter_conditioning(orc, job, scene, n)
    b = harness(job, scene, n)
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["bin"], b["bin"])
    e = np.nan_to_num(np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"])))
    # no blanket waiver: every ray within 1e-9, or within 4x what +-2 ulp in the libm calls does to THIS ray
    assert np.all(e[stable] <= np.maximum(TOL, 4.0 * resp[stable])), float((e / np.maximum(TOL, 4.0 * resp))[stable].max())
    assert np.all(e[resp < TOL / 4] < TOL)
    assert np.mean(e < TOL) > 0.999
    assert (a["status"] == 2).any() or (a["status"] == 6).any() or "faithful" in cid  # absorption exercised


@pytest.mark.parametrize("stop", [1, 2, 3, 4])
def test_stage_outputs(orc, harness, stop):
    for phase in (1, 2):
        scene = cases.scene_for(orc, cases.C1, phase)
        job = abi.default_job(phase, stop_after=stop, first_ray=3 * 10 ** 10)
        a = orc.trace_rays(job, scene, 20_000)
        b = harness(job, scene, 20_000)
        assert np.array_equal(a["status"], b["status"])
        e = np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"]))
        assert np.nanmax(e) < TOL


def test_on_axis_ray_sees_zero_reflectance(orc, harness):
    """at EXACTLY normal incidence the reference's fresnel() returns 0, so an
    on-axis ray is never reflected -- even with a draw (0.01) far below the true 4 % reflectance."""
    for phase, bottle in ((2, True), (2, False), (1, False)):
        scene = cases.scene_for(orc, cases.C2, phase)
        job = abi.default_job(2, use_bottle=bottle, uniform_override=0.01)
        p = np.zeros((3, 2))
        d = np.array([[0.0, 1e-3], [0.0, 0.0], [1.0, np.sqrt(1 - 1e-6)]])
        a = orc.trace_rays(job, scene, 2, p, d)
        b = harness(job, scene, 2, p, d)
        assert a["status"][0] == 0 and tuple(a["bin"][:, 0]) == (0, 0)      # axial: sails through
        assert a["status"][1] != 0                                           # 1 mrad off axis: reflected
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])
        assert np.array_equal(a["dir"][:, 0], [0.0, 0.0, 1.0]) and np.array_equal(b["dir"][:, 0], [0.0, 0.0, 1.0])
