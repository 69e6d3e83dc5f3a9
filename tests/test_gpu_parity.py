"""GPU parity: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Tolerances: 1e-9 relative on exit position / direction (BASELINE.json
north_star), on the vector and per component (conftest.comp_err); status codes and detector bins
identical; integer images and counters bit-exact.  The scatter path has no blanket waiver: a ray
may exceed 1e-9 only as far as +-2 ulp in the libm calls of tauint / stokes move THAT ray in the
oracle itself (conftest.scatter_conditioning)."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import both_err, rel_err, scatter_conditioning

pytestmark = pytest.mark.gpu

TOL = 1e-9          # fp64 tolerance stated by BASELINE.json north_star


def test_uniforms_match_oracle(ort, orc):
    for phase in (1, 2):
        for ray in (0, 1, 12345, 2 ** 32 + 7, 10 ** 11 - 1):
            a = ort.uniforms(123456789, phase, ray, 0, 40)
            b = orc.uniforms(123456789, phase, ray, 0, 40)
            assert np.array_equal(a, b)
    assert not np.array_equal(ort.uniforms(1, 1, 0, 0, 8), ort.uniforms(2, 1, 0, 0, 8))


def _compare(ort, orc, files, phase, kw, n, tol, stop=0, first_ray=0):
    scene = cases.scene_for(orc, files, phase, kw)
    kw = dict(kw)
    job = abi.default_job(phase, stop_after=stop, first_ray=first_ray, **kw)
    a = orc.trace_rays(job, scene, n)
    b = ort.trace_rays(job, scene, n)
    assert np.array_equal(a["status"], b["status"]), np.flatnonzero(a["status"] != b["status"])[:10]
    assert np.array_equal(a["bin"], b["bin"])
    e, ec = both_err(a, b)
    return np.maximum(e, ec), a       # vector-relative and per-component, whichever is worse


@pytest.mark.parametrize("cid,files,phase,kw", cases.RAY_CASES, ids=[c[0] for c in cases.RAY_CASES])
def test_rays_match_oracle(ort, orc, cid, files, phase, kw):
    e, a = _compare(ort, orc, files, phase, kw, 200_000, TOL)
    assert np.nanmax(e) < TOL, np.nanmax(e)
    print("%s: worst of (vector-relative, per-component) error %.3g" % (cid, np.nanmax(e)))
    # unit directions wherever the ray is still a ray
    ok = a["status"] == 0
    if ok.any():
        assert np.abs(np.linalg.norm(a["dir"][:, ok], axis=0) - 1).max() < 1e-12


@pytest.mark.parametrize("cid,files,phase,kw", cases.SOURCE_CASES, ids=[c[0] for c in cases.SOURCE_CASES])
def test_other_sources_match_oracle(ort, orc, cid, files, phase, kw):
    """crs / isors / spot emitters: per-ray state after the source and at the detector, and the
    megakernel's image + histogram (isors trips the reference's `error stop` on the ~2.8 % of rays
    its axicon face reflects: counted as status 26, return code ORT_ETRACE)."""
    e, a = _compare(ort, orc, files, phase, kw, 100_000, TOL, stop=1)
    assert np.nanmax(e) < TOL
    e, a = _compare(ort, orc, files, phase, kw, 100_000, TOL)
    assert np.nanmax(e) < TOL
    n = 400_003
    scene = cases.scene_for(orc, files, phase, kw)
    kw2 = dict(kw)
    kw2.setdefault("total_rays", n)
    job = abi.default_job(phase, n, **kw2)
    img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
    oimg, olost, ohist = orc.trace(job, scene)
    assert np.array_equal(hist, ohist), list(zip(abi.STATUS_NAMES, hist[0], ohist[0]))
    assert np.array_equal(img, oimg) and np.array_equal(lost, olost)
    if kw.get("source") == "isors" and phase == 1:
        assert 0.02 < hist[0, 26] / n < 0.04


@pytest.mark.parametrize("stop", [1, 2, 3, 4])
@pytest.mark.parametrize("phase", [1, 2])
def test_rays_stage_by_stage(ort, orc, phase, stop):
    e, _ = _compare(ort, orc, cases.C2, phase, {}, 100_000, TOL, stop=stop, first_ray=10 ** 10)
    assert np.nanmax(e) < TOL


@pytest.mark.parametrize("cid,files,phase,kw", cases.SCATTER_CASES, ids=[c[0] for c in cases.SCATTER_CASES])
def test_scatter_rays_match_oracle(ort, orc, cid, files, phase, kw):
    n = 200_000
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a, resp, stable = scatter_conditioning(orc, job, scene, n)
    b = ort.trace_rays(job, scene, n)
    assert np.array_equal(a["status"], b["status"]), np.flatnonzero(a["status"] != b["status"])[:10]
    assert np.array_equal(a["bin"], b["bin"])
    e = np.nan_to_num(np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"])))
    # every ray within 1e-9, or within 4x what +-2 ulp in the libm calls does to THIS ray in the oracle
    allowed = np.maximum(TOL, 4.0 * resp)
    assert np.all(e[stable] <= allowed[stable]), float((e / allowed)[stable].max())
    assert np.all(e[resp < TOL / 4] < TOL)
    print("%s: %d of %d rays above 1e-9 (largest %.3g, its own conditioning bound %.3g); %d rays whose status "
          "the jitter can flip" % (cid, int((e >= TOL).sum()), n, e.max(), allowed[np.argmax(e)], int((~stable).sum())))
    assert np.mean(e < TOL) > 0.999


def test_explicit_input_rays(ort, orc):
    """Caller-supplied ray list (grid + random) through the point-phase path."""
    rng = np.random.default_rng(7)
    n = 100_000
    th = rng.uniform(0, 0.35, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    d = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
    p = np.stack([rng.uniform(-2e-3, 2e-3, n), rng.uniform(-2e-3, 2e-3, n), rng.uniform(-5e-3, 5e-3, n)])
    scene = cases.scene_for(orc, cases.C2, 2)
    job = abi.default_job(2)
    a = orc.trace_rays(job, scene, n, p, d)
    b = ort.trace_rays(job, scene, n, p, d)
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["bin"], b["bin"])
    assert max(np.nanmax(rel_err(a["pos"], b["pos"])), np.nanmax(rel_err(a["dir"], b["dir"]))) < TOL


def test_known_answer_rays(ort, orc):
    """SURVEY 8(c) KAT-A (bin (60,0)) and KAT-C (collimated axis ray -> bin (0,0))."""
    s843 = cases.scene_for(orc, cases.C2, 2)
    j = abi.default_job(2, uniform_override=0.5)
    p = np.zeros((3, 1))
    d = np.array([[np.sin(0.1)], [0.0], [np.cos(0.1)]])
    r = ort.trace_rays(j, s843, 1, p, d)
    assert r["status"][0] == 0 and tuple(r["bin"][:, 0]) == (60, 0)
    assert abs(r["pos"][0, 0] - 1.49695788349947e-3) < 1e-12
    s785 = cases.scene_for(orc, cases.C2, 1)
    j = abi.default_job(2, use_bottle=False, uniform_override=0.5)
    d = np.array([[np.sin(0.05) * np.cos(1.0)], [np.sin(0.05) * np.sin(1.0)], [np.cos(0.05)]])
    r = ort.trace_rays(j, s785, 1, p, d)
    assert r["status"][0] == 0 and tuple(r["bin"][:, 0]) == (0, 0)
    assert abs(r["pos"][0, 0] - 7.0433e-6) < 1e-9 and abs(r["pos"][1, 0] - 1.09693e-5) < 1e-9
    # u = 0 always reflects where R > 0: first lossy interface ends the ray
    j0 = abi.default_job(2, uniform_override=0.0)
    d = np.array([[np.sin(0.1)], [0.0], [np.cos(0.1)]])
    r = ort.trace_rays(j0, s843, 1, p, d)
    assert r["status"][0] == abi.STATUS_NAMES.index("bottle_inner_reflect")


IMG_CASES = [c for c in cases.RAY_CASES if c[0] in (
    "c1-ring", "c1-point", "c2-ring", "c2-point", "c2-point-nobottle", "c1-point-iris-before",
    "ellipse-ring", "ellipse-point-fixed", "other-point")] + cases.SCATTER_CASES[:3]


@pytest.mark.parametrize("cid,files,phase,kw", IMG_CASES, ids=[c[0] for c in IMG_CASES])
def test_trace_image_bit_exact(ort, orc, cid, files, phase, kw):
    """The production megakernel (device-side sources, warp compaction, aggregated atomics):
    integer image, loss counter and per-status histogram identical to the oracle's."""
    n = 1_000_003  # not a multiple of 32 on purpose
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, n, **kw)
    img, lost, hist, tm = ort.trace(job, scene)
    oimg, olost, ohist = orc.trace(job, scene)
    assert int(hist[..., :27].sum()) == n
    assert np.array_equal(hist, ohist), list(zip(abi.STATUS_NAMES, hist[0], ohist[0]))
    assert np.array_equal(lost, olost)
    assert np.array_equal(img, oimg)
    assert int(img.sum()) == int(hist[0, 0])
    assert tm.kernel_launches >= 1 and tm.trace_seconds > 0


@pytest.mark.parametrize("phase", [1, 2])
def test_flat_kernel_equals_megakernel(ort, orc, phase):
    scene = cases.scene_for(orc, cases.C2, phase)
    n = 777_777
    a = ort.trace(abi.default_job(phase, n), scene)
    b = ort.trace(abi.default_job(phase, n, flags=abi.FLAG_NO_COMPACTION), scene)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_batched_scenes_equal_individual_runs(ort, orc):
    """BASELINE config 3: the 15 clearBottle-large offset files in one launch (SURVEY quirk 5:
    +2..+14 mm collapse to one geometry through the offset guard)."""
    names = ["clearBottle-large_%dmm.params" % mm for mm in range(-14, 15, 2)]
    n = 100_000
    for phase in (1, 2):
        scenes = [cases.scene_for(orc, (nm,) + cases.C2[1:], phase) for nm in names]
        job = abi.default_job(phase, n)
        img, lost, hist, _ = ort.trace(job, scenes)
        for i in (0, 3, 7, 8, 14):
            one = ort.trace(job, scenes[i])
            assert np.array_equal(img[i], one[0][0]) and np.array_equal(hist[i], one[2][0])
        oimg, olost, ohist = orc.trace(job, scenes)
        assert np.array_equal(img, oimg) and np.array_equal(hist, ohist)
        for i in range(9, 15):  # guard-collapsed geometries are the same run
            assert np.array_equal(img[8], img[i])


def test_ray_range_partition_sums_to_whole(ort, orc):
    """G-invariance: disjoint ray-index ranges traced separately add up to the
    whole job bit-for-bit -- what the multi-GPU reduce relies on."""
    from opticalraytrace_b200 import partition
    scene = cases.scene_for(orc, cases.C2, 2)
    n = 600_001
    whole = ort.trace(abi.default_job(2, n, first_ray=5), scene)
    acc_img = np.zeros_like(whole[0])
    acc_hist = np.zeros_like(whole[2])
    for r in range(3):
        first, cnt = partition(n, r, 3, first_ray=5)
        part = ort.trace(abi.default_job(2, cnt, first_ray=first), scene)
        acc_img += part[0]
        acc_hist += part[2]
    assert np.array_equal(acc_img, whole[0]) and np.array_equal(acc_hist, whole[2])


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 4097])
def test_ragged_sizes(ort, orc, n):
    scene = cases.scene_for(orc, cases.C1, 2)
    job = abi.default_job(2, n)
    img, lost, hist, _ = ort.trace(job, scene)
    oimg, olost, ohist = orc.trace(job, scene)
    assert int(hist[..., :27].sum()) == n
    assert np.array_equal(img, oimg) and np.array_equal(hist, ohist)


def test_full_size_properties(ort, orc):
    """BASELINE.json-size run (config 2, 2^28 rays per phase here): size-independent properties."""
    n = 1 << 28
    for phase in (1, 2):
        scene = cases.scene_for(orc, cases.C2, phase)
        img, lost, hist, tm = ort.trace(abi.default_job(phase, n), scene)
        assert int(hist[..., :27].sum()) == n                      # every ray accounted for exactly once
        assert int(img.sum()) == int(hist[0, 0])         # image mass == binned count
        assert int(lost[0]) == sum(int(hist[0, s]) for s in range(32) if abi.status_is_lost(s))
        assert hist[0, 18] == 0 and hist[0, 24] == 0     # no `error stop` invariants hit
        # linearity: two half-size runs over disjoint ray ranges add to the same image
        h1 = ort.trace(abi.default_job(phase, n // 2), scene)
        h2 = ort.trace(abi.default_job(phase, n // 2, first_ray=n // 2), scene)
        assert np.array_equal(h1[0] + h2[0], img)
        # agreement with the oracle's fractions within 5 sigma (binomial)
        m = 2_000_000
        _, _, oh = orc.trace(abi.default_job(phase, m, first_ray=n), scene)
        for s in range(26):
            p = hist[0, s] / n
            sigma = np.sqrt(max(p * (1 - p), 1e-12) / m)
            assert abs(oh[0, s] / m - p) < 5 * sigma + 1e-6, (s, p, oh[0, s] / m)


def test_bad_arguments(ort, orc):
    from opticalraytrace_b200.lib import OrtError
    scene = cases.scene_for(orc, cases.C1, 1)
    with pytest.raises(OrtError):
        ort.trace(abi.default_job(3, 10), scene)
    j = abi.default_job(1, 10)
    j.precision = 16
    with pytest.raises(OrtError):
        ort.trace(j, scene)
    with pytest.raises(OrtError):
        ort.trace(abi.default_job(1, 10), [])


def test_fast_math_accuracy(ort):
    """The library's slow-path-free reciprocal / division / sqrt / rsqrt: <= 2 ulp on the device."""
    worst = ort.math_selftest(1 << 24)
    print("max ulp error:", worst)
    assert worst["rcp"] <= 2 and worst["div"] <= 2 and worst["sqrt"] <= 2 and worst["rsqrt"] <= 3, worst


def test_ring_without_aim_plane_shortcut(ort, orc):
    """A hand-built scene whose L2 is shifted 1 mm along z: the flat face no longer lies in the
    ring source's aim plane, so stage 0 must run the general source + aperture arithmetic."""
    scene = cases.scene_for(orc, cases.C1, 1)
    scene.L2.centre[2] += 1e-3
    n = 500_003
    job = abi.default_job(1, n)
    img, lost, hist, _ = ort.trace(job, scene)
    oimg, olost, ohist = orc.trace(job, scene)
    assert np.array_equal(hist, ohist) and np.array_equal(img, oimg) and np.array_equal(lost, olost)
    a = orc.trace_rays(abi.default_job(1), scene, 100_000)
    b = ort.trace_rays(abi.default_job(1), scene, 100_000)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])
    assert max(np.nanmax(rel_err(a["pos"], b["pos"])), np.nanmax(rel_err(a["dir"], b["dir"]))) < TOL


def test_on_axis_ray_sees_zero_reflectance(ort, orc):
    """SURVEY quirk 3 on the device: exactly normal incidence -> R = 0 -> never reflected."""
    for bottle in (True, False):
        scene = cases.scene_for(orc, cases.C2, 2)
        job = abi.default_job(2, use_bottle=bottle, uniform_override=0.01)
        p = np.zeros((3, 2))
        d = np.array([[0.0, 1e-3], [0.0, 0.0], [1.0, np.sqrt(1 - 1e-6)]])
        a = orc.trace_rays(job, scene, 2, p, d)
        b = ort.trace_rays(job, scene, 2, p, d)
        assert b["status"][0] == 0 and tuple(b["bin"][:, 0]) == (0, 0) and b["status"][1] != 0
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])
        assert np.array_equal(b["dir"][:, 0], [0.0, 0.0, 1.0])


def test_more_rays_than_one_launch_holds(ort, orc):
    """Ray ids inside a launch are 32-bit offsets, so ort_trace splits jobs into launches of at
    most 2^31 rays (2^29-ray slices, two kernels each, on the ring loop's filter path): a job of
    2^32 + 5 rays must account for every ray once and equal the sum of its parts, also far out in
    the ray-index space (1e11-ray jobs, config 5)."""
    scene = cases.scene_for(orc, cases.C2, 1)
    n = (1 << 32) + 5
    first = 10 ** 11
    img, lost, hist, tm = ort.trace(abi.default_job(1, n, first_ray=first), scene)
    assert tm.kernel_launches == 2 * 9 and int(hist[..., :27].sum()) == n and int(img.sum()) == int(hist[0, 0])
    img0, lost0, hist0, tm0 = ort.trace(abi.default_job(1, n, first_ray=first, flags=abi.FLAG_NO_FILTER), scene)
    assert tm0.kernel_launches == 3 and np.array_equal(img0, img) and np.array_equal(hist0, hist)
    pimg, _, phist, ptm = ort.trace(abi.default_job(2, n, first_ray=first), cases.scene_for(orc, cases.C2, 2))
    assert ptm.kernel_launches == 3 and int(phist[..., :27].sum()) == n and int(pimg.sum()) == int(phist[0, 0])
    acc_img, acc_hist = np.zeros_like(img), np.zeros_like(hist)
    for lo, cnt in ((0, 1 << 31), (1 << 31, 1 << 31), (1 << 32, 5)):
        part = ort.trace(abi.default_job(1, cnt, first_ray=first + lo), scene)
        acc_img += part[0]
        acc_hist += part[2]
    assert np.array_equal(acc_img, img) and np.array_equal(acc_hist, hist)
    # the tail of the index space matches the oracle ray by ray
    a = orc.trace_rays(abi.default_job(1, first_ray=first + (1 << 32) - 1000), scene, 1005)
    b = ort.trace_rays(abi.default_job(1, first_ray=first + (1 << 32) - 1000), scene, 1005)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["bin"], b["bin"])


def test_small_batched_call_is_the_same_on_one_stream_and_on_several(ort, orc):
    """A batched call with few rays per scene spreads its scenes over several streams;
    ORT_FLAG_ONE_LANE runs them back to back.  Same images, histograms and loss counters either way,
    equal to one call per scene, for both loops and for a mix of clear, elliptic and scattering bottles
    (the kernel is chosen per scene)."""
    mix = [cases.C1, cases.C2, cases.ELL, cases.OTHER, cases.SCATTER_CASES[0][1], cases.C2, cases.OTHER2]
    n = 300_007
    for phase in (1, 2):
        scenes = [cases.scene_for(orc, f, phase) for f in mix]
        img, lost, hist, tm = ort.trace(abi.default_job(phase, n), scenes, allow_trap=True)
        img1, lost1, hist1, _ = ort.trace(abi.default_job(phase, n, flags=abi.FLAG_ONE_LANE), scenes, allow_trap=True)
        assert np.array_equal(img, img1) and np.array_equal(hist, hist1) and np.array_equal(lost, lost1)
        for k, sc in enumerate(scenes):
            i0, l0, h0, _ = ort.trace(abi.default_job(phase, n), sc, allow_trap=True)
            assert np.array_equal(i0[0], img[k]) and np.array_equal(h0[0], hist[k]) and l0[0] == lost[k], (phase, k)
        oimg, olost, ohist = orc.trace(abi.default_job(phase, n), scenes)
        assert np.array_equal(oimg, img) and np.array_equal(ohist, hist)
