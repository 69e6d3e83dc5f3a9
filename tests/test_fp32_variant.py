"""The fp32 variant (ort_job.precision = 32): same draws, same decisions except within float
rounding of a threshold.  north_star's tolerance for it is 1e-5 relative on exit positions and
directions; histograms must agree within Poisson noise."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import rel_err

CORE = [c for c in cases.RAY_CASES if c[0] in (
    "c1-ring", "c1-point", "c2-ring", "c2-point", "c2-point-nobottle", "c1-point-iris-before",
    "ellipse-ring", "ellipse-point-fixed", "other-ring", "other-point")]
TOL32 = 1e-5


def _check(a, b, n):
    same = a["status"] == b["status"]
    assert np.mean(~same) < 5e-4                       # decisions flip only next to a threshold
    # a ray rejected by the NA test or flagged FAR was carried to the image plane along a direction
    # that may be almost parallel to it: its position there is pos + dir * (z / dir.z), arbitrarily
    # ill-conditioned, so only its direction is compared
    skew = np.isin(a["status"][same], (21, 22))
    e_pos = np.where(skew, 0.0, rel_err(a["pos"][:, same], b["pos"][:, same]))
    e = np.maximum(e_pos, rel_err(a["dir"][:, same], b["dir"][:, same]))
    # a flipped decision can still end in the same status (reflected at L2's flat face and then at the
    # back of its sphere, against transmitted and reflected at the curved face: both status 11); such
    # rays are somewhere else entirely and count as flips, not as numerical error
    elsewhere = e > 1e-2
    assert (np.sum(~same) + np.sum(elsewhere)) / n < 5e-4
    e = e[~elsewhere]
    assert np.nanquantile(e, 0.999) < TOL32, np.nanquantile(e, 0.999)
    assert np.nanmedian(e) < 1e-6
    assert np.nanmax(e) < 1e-3                         # isolated grazing / near-TIR rays
    binsame = (a["bin"] == b["bin"]).all(axis=0)
    assert np.mean(~binsame & same) < 1e-3             # rays within float rounding of a bin edge


@pytest.mark.parametrize("cid,files,phase,kw", CORE, ids=[c[0] for c in CORE])
def test_fp32_host_math(orc, harness, cid, files, phase, kw):
    n = 200_000
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a = orc.trace_rays(job, scene, n)
    job.precision = 32
    _check(a, harness(job, scene, n), n)


@pytest.mark.gpu
@pytest.mark.parametrize("cid,files,phase,kw", CORE, ids=[c[0] for c in CORE])
def test_fp32_cuda_rays(ort, orc, cid, files, phase, kw):
    n = 200_000
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a = orc.trace_rays(job, scene, n)
    job.precision = 32
    _check(a, ort.trace_rays(job, scene, n), n)


@pytest.mark.gpu
@pytest.mark.parametrize("phase", [1, 2])
def test_fp32_image_agrees_with_fp64(ort, orc, phase):
    """Same ray set through both variants: the images differ only by the handful of rays next to a
    decision threshold or a bin edge; against an independent ray set both are Poisson-consistent."""
    n = 30_000_000 if phase == 1 else 4_000_000   # the ring loop bins only 2.5 % of its rays
    scene = cases.scene_for(orc, cases.C1, phase)
    j64 = abi.default_job(phase, n)
    j32 = abi.default_job(phase, n)
    j32.precision = 32
    i64, l64, h64, _ = ort.trace(j64, scene)
    i32, l32, h32, _ = ort.trace(j32, scene)
    assert int(h32.sum()) == n and int(i32.sum()) == int(h32[0, 0])
    assert np.abs(h64.astype(np.int64) - h32.astype(np.int64)).sum() < 2e-3 * n
    assert np.abs(i64.astype(np.int64) - i32.astype(np.int64)).sum() < 4e-3 * i64.sum()
    # independent rays (another index range): chi-square over well-filled bins
    other, _, _, _ = ort.trace(abi.default_job(phase, n, first_ray=10 * n), scene)
    m = (i32 + other) >= 30
    a, b = i32[m].astype(float), other[m].astype(float)
    chi2 = ((a - b) ** 2 / (a + b)).sum()
    dof = m.sum()
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)
    assert abs(int(l32[0]) - int(l64[0])) < 2e-3 * n


@pytest.mark.gpu
def test_fp32_megakernel_equals_explicit_kernel(ort, orc):
    """The two fp32 kernels run the same arithmetic: binning the explicit kernel's output gives the
    megakernel's image bit for bit."""
    n = 300_000
    scene = cases.scene_for(orc, cases.C2, 2)
    job = abi.default_job(2, n)
    job.precision = 32
    img, lost, hist, _ = ort.trace(job, scene)
    r = ort.trace_rays(job, scene, n)
    ok = r["status"] == 0
    ref = np.zeros((401, 401), dtype=np.uint64)
    np.add.at(ref, (r["bin"][1, ok] + 200, r["bin"][0, ok] + 200), 1)
    assert np.array_equal(ref, img[0])
    assert np.array_equal(np.bincount(r["status"], minlength=32), hist[0])
