"""The ring loop's single-precision culling filter (ort_ring_filter, ort_filter.cuh).

A verdict s > 0 claims "this ray ends with status s, no need to trace it in fp64": it has to agree
with the oracle on EVERY ray; a verdict 0 hands the ray to the fp64 stage and is always safe.  On
CPU the filter is the host-compiled header; on the GPU box ORT_FLAG_VERIFY_FILTER makes the kernel
run both paths on every ray and count disagreements."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.test_fuzz_scenes import random_case

RING_SETUPS = [(cases.C1, {}), (cases.C2, {}), (cases.ELL, {}), (cases.ELLS, {}), (cases.OTHER, {}),
               (cases.OTHER2, {}), (cases.C2, dict(iris="before", iris_radius=0.6)),
               (cases.C1, dict(iris="before", iris_radius=0.3))]


def _check_filter(orc, harness, scene, job, n):
    verdict, shortcut = harness.ring_filter(job, scene, n)
    ref = orc.trace_rays(job, scene, n)["status"]
    if not shortcut:            # L2 moved out of the aim plane: the launcher does not use the filter
        return 0, 0, 0
    stage_a = verdict == -1
    assert np.all(ref[stage_a] == 9)
    certain = verdict > 0
    bad = certain & (verdict != ref)
    assert not bad.any(), (np.flatnonzero(bad)[:5], verdict[bad][:5], ref[bad][:5])
    # rays the filter hands back although they do end before L3's aperture test: wasted, not wrong
    in_b = ~stage_a
    ended_in_b = in_b & np.isin(ref, [9, 10, 11, 12, 13, 14])
    return int(in_b.sum()), int(certain.sum()), int(ended_in_b.sum())


@pytest.mark.parametrize("k", range(len(RING_SETUPS)))
def test_filter_verdicts_are_exact(orc, harness, k):
    files, kw = RING_SETUPS[k]
    scene = cases.scene_for(orc, files, 1)
    n = 1_500_000
    job = abi.default_job(1, first_ray=7 * 10 ** 9 * k, **kw)
    in_b, certain, ended = _check_filter(orc, harness, scene, job, n)
    # the filter must actually cull: nearly every ray that ends in stage B is called in fp32
    assert certain > 0.85 * ended, (in_b, certain, ended)


@pytest.mark.parametrize("k", range(0, 40, 1))
def test_filter_on_random_scenes(orc, harness, k):
    scene, phase, kw = random_case(orc, k)
    kw.pop("use_bottle")
    job = abi.default_job(1, first_ray=k * 10 ** 9, **kw)
    _check_filter(orc, harness, scene, job, 200_000)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(RING_SETUPS)))
def test_cuda_filter_kernel_equals_oracle_and_unfiltered(ort, orc, k):
    files, kw = RING_SETUPS[k]
    scene = cases.scene_for(orc, files, 1)
    n = 400_003
    job = abi.default_job(1, n, first_ray=3 * 10 ** 9 * k, **kw)
    img, lost, hist, tm = ort.trace(job, scene)
    oimg, olost, ohist = orc.trace(job, scene)
    assert np.array_equal(hist, ohist) and np.array_equal(img, oimg) and np.array_equal(lost, olost)
    # at scale: filter on / off / verify agree to the last count
    big = 3_000_000_000
    job = abi.default_job(1, big, first_ray=10 ** 12 + k, **kw)
    img, lost, hist, tm = ort.trace(job, scene)
    job.flags = abi.FLAG_NO_FILTER
    img0, lost0, hist0, tm0 = ort.trace(job, scene)
    assert np.array_equal(hist, hist0) and np.array_equal(img, img0) and np.array_equal(lost, lost0)
    job.flags = abi.FLAG_VERIFY_FILTER
    img1, lost1, hist1, _ = ort.trace(job, scene)
    called, wrong = int(hist1[0, abi.FILTER_SLOT_CALLED]), int(hist1[0, abi.FILTER_SLOT_WRONG])
    ended_in_b = int(hist0[0, 10:15].sum())
    assert wrong == 0
    assert called > 0.85 * ended_in_b, (called, ended_in_b)
    hist1[0, 30:] = 0
    assert np.array_equal(hist1, hist0) and np.array_equal(img1, img0)
    print("setup %d: %.3g rays/s filtered, %.3g rays/s unfiltered, filter called %.2f%% of stage-B deaths"
          % (k, big / tm.trace_seconds, big / tm0.trace_seconds, 100.0 * called / max(ended_in_b, 1)))


@pytest.mark.gpu
def test_cuda_filter_on_random_scenes(ort, orc):
    jobs, scenes = [], []
    for k in range(40):
        scene, phase, kw = random_case(orc, k)
        kw.pop("use_bottle")
        n = 200_000_000
        job = abi.default_job(1, n, first_ray=k * 10 ** 10, **kw)
        job.flags |= abi.FLAG_VERIFY_FILTER
        img1, lost1, hist1, _ = ort.trace(job, scene, allow_trap=True)
        assert int(hist1[0, abi.FILTER_SLOT_WRONG]) == 0, k
        job.flags &= ~abi.FLAG_VERIFY_FILTER
        img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
        job.flags |= abi.FLAG_NO_FILTER
        img0, lost0, hist0, _ = ort.trace(job, scene, allow_trap=True)
        assert np.array_equal(hist, hist0) and np.array_equal(img, img0), k


def test_integer_aperture_cut_is_the_fp64_test(orc, harness):
    """Stage A compares the raw 64-bit draw with a cut found by bisection: the draws just below and
    at the cut must fall on the two sides of the fp64 expression u2 * lens_r2 > radius^2, for every
    lens of the shipped set (ort_ring_aim_cut, ort_flatten.h)."""
    for files in (cases.C1, cases.C2, cases.OTHER, cases.OTHER2):
        scene = cases.scene_for(orc, files, 1)
        job = abi.default_job(1)
        cut, have = harness.ring_aim_cut(job, scene)
        assert have and cut % 2048 == 0
        lens_r2 = (scene.L2.radius + 10e-3) ** 2
        rad2 = scene.L2.radius ** 2
        for w, want in ((cut - 1, False), (cut, True), (cut + 2047, True), (cut - 2048, False)):
            u2 = float(w >> 11) * 2.0 ** -53
            assert (u2 * lens_r2 > rad2) is want
        assert abs((cut >> 11) * 2.0 ** -53 - rad2 / lens_r2) < 1e-15


@pytest.mark.gpu
def test_survivor_list_overflow_is_reported():
    """The list of rays handed to fp64 is sized 16 sigma above its expectation; if it ever were
    too small the call must fail loudly instead of dropping rays.  The capacity override is a hook
    of the assert-instrumented build only (libort_debug.so); the release library ignores it."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dbg = os.path.join(root, "opticalraytrace_b200", "libort_debug.so")
    if not os.path.exists(dbg):
        pytest.skip("libort_debug.so not built (make -C opticalraytrace_b200/csrc DEBUG=1)")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from opticalraytrace_b200 import abi, lib\n"
        "from tests import cases, oracle_lib as orc\n"
        "lib.init(1)\n"
        "scene = cases.scene_for(orc, cases.C1, 1)\n"
        "try:\n"
        "    lib.trace(abi.default_job(1, 2_000_000), scene)\n"
        "    print('NO ERROR')\n"
        "except lib.OrtError as e:\n"
        "    print('ERROR', e)\n" % root)
    env = dict(os.environ, ORT_TEST_RING_LIST_CAP="100")
    out = subprocess.run([sys.executable, "-c", code], env=dict(env, ORT_LIB=dbg), capture_output=True, text=True)
    assert "survivor list overflowed" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)   # release build
    assert "NO ERROR" in out.stdout, out.stdout + out.stderr


EXTREMES = [
    ("nearly flat L2", lambda sc: setattr(sc.L2, "n2", 1.0005)),
    ("dense L2", lambda sc: setattr(sc.L2, "n2", 2.6)),
    ("dense L3", lambda sc: (setattr(sc.L3, "n2", 2.2), setattr(sc.L3, "n3", 2.9))),
    ("L3 far away", lambda sc: sc.L3.centre1.__setitem__(2, sc.L3.centre1[2] + 0.4)),
    ("L3 nearly touching", lambda sc: sc.L3.centre1.__setitem__(2, sc.L3.centre1[2] - 0.02)),
    ("huge L3 radius of curvature", lambda sc: (setattr(sc.L3, "R1", sc.L3.R1 * 50),
                                                sc.L3.centre1.__setitem__(2, sc.L3.centre1[2] + sc.L3.R1 * 49 / 50))),
    ("tiny L3 aperture", lambda sc: setattr(sc.L3, "radius", sc.L3.radius * 0.05)),
    ("wide ring", lambda sc: (setattr(sc, "r1", sc.r1 * 0.1), setattr(sc, "r2", sc.r2 * 3.0))),
]


@pytest.mark.parametrize("k", range(len(EXTREMES)))
def test_filter_on_extreme_geometries(orc, harness, k):
    """Far outside the shipped parameter range the filter may hand everything to fp64, but a
    verdict it does give must still be the oracle's."""
    name, tweak = EXTREMES[k]
    scene = cases.scene_for(orc, cases.C2, 1)
    tweak(scene)
    job = abi.default_job(1, first_ray=5 * 10 ** 8 * k)
    in_b, certain, ended = _check_filter(orc, harness, scene, job, 400_000)
    print(name, in_b, certain, ended)


@pytest.mark.gpu
def test_cuda_filter_on_extreme_geometries(ort, orc):
    for k, (name, tweak) in enumerate(EXTREMES):
        scene = cases.scene_for(orc, cases.C2, 1)
        tweak(scene)
        job = abi.default_job(1, 300_000_000, first_ray=5 * 10 ** 8 * k, flags=abi.FLAG_VERIFY_FILTER)
        hist = ort.trace(job, scene, allow_trap=True)[2]
        assert int(hist[0, abi.FILTER_SLOT_WRONG]) == 0, name
        job.nrays = 300_000
        job.flags = 0
        img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
        oimg, olost, ohist = orc.trace(job, scene)
        assert np.array_equal(hist, ohist) and np.array_equal(img, oimg), name


def test_usable_flag(orc, harness):
    """The launcher runs the filter when its bound constants are finite and small enough to decide
    anything (ort_make_filter): true for every shipped set-up; a system moved 50 m off the origin
    rounds its fp32 coordinates by micrometres and runs the all-fp64 kernel instead.  A pin-hole
    aperture or iris is no longer a reason to switch off -- the bounds are per ray and relative to
    the aperture -- but whatever the filter says there must still be the oracle's status."""
    job = abi.default_job(1)
    for files in (cases.C1, cases.C2, cases.ELL, cases.ELLS, cases.OTHER, cases.OTHER2):
        assert harness.ring_filter_in_range(job, cases.scene_for(orc, files, 1))
    far = cases.scene_for(orc, cases.C2, 1)
    for c in (far.bottle.centre, far.L2.centre, far.L3.centre1, far.L3.centre2, far.L3.centre3):
        c[2] += 50.0
    assert not harness.ring_filter_in_range(job, far)
    pin = cases.scene_for(orc, cases.C2, 1)
    pin.L3.radius *= 0.02
    _check_filter(orc, harness, pin, job, 400_000)
    assert harness.filter_bounds(job, pin, 400_000, fuzz=True)[2]["violations"] == 0
    iris = abi.default_job(1, iris="before", iris_radius=0.01)
    _check_filter(orc, harness, cases.scene_for(orc, cases.C2, 1), iris, 400_000)
    assert harness.filter_bounds(iris, cases.scene_for(orc, cases.C2, 1), 400_000, fuzz=True)[2]["violations"] == 0


def test_high_word_rule_of_stage_a(orc, harness):
    """The cull kernel decides stage A on the high 32 bits of the draw: below the cut's high word
    the fp64 aperture expression must say inside, above it outside; only an equal high word
    (2^-32 of the rays) is left to fp64."""
    rng = np.random.default_rng(5)
    scene = cases.scene_for(orc, cases.C2, 1)
    cut, have = harness.ring_aim_cut(abi.default_job(1), scene)
    assert have
    cut_hi = cut >> 32
    lens_r2 = (scene.L2.radius + 10e-3) ** 2
    rad2 = scene.L2.radius ** 2
    his = np.concatenate([rng.integers(0, 2 ** 32, 20000, dtype=np.uint64),
                          np.array([cut_hi - 1, cut_hi + 1, 0, 2 ** 32 - 1], dtype=np.uint64)])
    los = rng.integers(0, 2 ** 32, his.size, dtype=np.uint64)
    for hi, lo in zip(his.tolist(), los.tolist()):
        if hi == cut_hi:
            continue
        u2 = float(((hi << 32) | lo) >> 11) * 2.0 ** -53
        assert (u2 * lens_r2 > rad2) == (hi > cut_hi)


def _edge_cases():
    import json
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edge_rays_v2.json")))
    return [(name, c["cut_hi"], ray) for name, c in sorted(d["cases"].items()) for ray in c["rays"]]


def test_edge_ray_fixture_matches_the_generator(orc, harness):
    """tests/golden/edge_rays_v2.json (make_edge_rays.py): rays whose aim word EQUALS the cut's high word, the
    one case stage A cannot decide on the word.  Checked against the oracle's own draw: slot 2 of each
    listed ray, as 53 bits, has exactly that high word."""
    for name, cut_hi, ray in _edge_cases():
        files = cases.C1 if name == "c1" else cases.C2
        scene = cases.scene_for(orc, files, 1)
        cut, have = harness.ring_aim_cut(abi.default_job(1), scene)
        assert have and cut >> 32 == cut_hi
        u2 = orc.uniforms(abi.default_job(1).seed, 1, ray, 2, 1)[0]
        assert (int(u2 * 2.0 ** 53) << 11) >> 32 == cut_hi, (name, ray)


@pytest.mark.gpu
def test_cuda_rays_on_the_aperture_cut(ort, orc):
    """A few rays around every listed edge ray through the culling kernel (the ray goes on to fp64), the
    all-fp64 kernel (it evaluates the whole aperture expression) and the oracle: identical counts and images;
    the fp32 kernel takes the same branch and must at least account for every ray.  The windows start off a
    multiple of four, so the first and last quad of the launch are partial."""
    for name, cut_hi, ray in _edge_cases():
        files = cases.C1 if name == "c1" else cases.C2
        scene = cases.scene_for(orc, files, 1)
        for first, n in ((ray - 3, 9), (ray, 1), (ray - 130, 263)):
            job = abi.default_job(1, n, first_ray=first)
            oimg, olost, ohist = orc.trace(job, scene)
            img, lost, hist, _ = ort.trace(job, scene)
            assert np.array_equal(hist, ohist) and np.array_equal(img, oimg) and np.array_equal(lost, olost), (name, ray, first, n)
            job.flags = abi.FLAG_NO_FILTER
            img0, lost0, hist0, _ = ort.trace(job, scene)
            assert np.array_equal(hist0, ohist) and np.array_equal(img0, oimg), (name, ray, first, n)
            job32 = abi.default_job(1, n, first_ray=first)
            job32.precision = 32
            _, _, hist32, _ = ort.trace(job32, scene)
            assert int(hist32[..., :27].sum()) == n


def test_every_filter_parameter_is_set_and_used():
    """OrtfParamsT (ort_filter.cuh) is the flat list of constants both lane policies read; a field that
    ortf_make_params forgets would be an uninitialised kernel parameter."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "opticalraytrace_b200", "csrc", "ort_filter.cuh")).read()
    a = src.index("struct OrtfParamsT {")
    body = src[a:src.index("};", a)]
    fields = []
    for line in body.split("\n"):
        m = re.match(r"\s*C\s+(.*);", line.split("/*")[0])
        if m:
            fields += [f.strip() for f in m.group(1).split(",")]
    assert len(fields) > 60
    assigned = set(re.findall(r"ORTF_SET\((\w+),", src)) - {"field"}
    used = set(re.findall(r"\bk\.(\w+)", src[src.index("template <typename P, bool FROM_FLAT>"):]))
    assert set(fields) == assigned
    assert set(fields) <= used


@pytest.mark.parametrize("k", range(len(RING_SETUPS)))
def test_two_rays_per_lane_equal_one_ray_per_lane(orc, harness, k):
    """The culling kernel runs the filter on two rays per lane (ortf_filter<OrtfTwo>).  A ray's verdict must
    not depend on its lane partner -- a partner that ends earlier, later, or trips a guard -- nor on which
    half it sits in: for several pairings, both halves give the one-ray instantiation's verdict on every
    ray that passes stage A.  (Host build: the packed operations are done per half; the masks, the exits
    and the selects under test are the device's.)"""
    files, kw = RING_SETUPS[k]
    scene = cases.scene_for(orc, files, 1)
    n = 200_000
    job = abi.default_job(1, n, first_ray=7 * 10 ** 9 + k, **kw)
    one, usable = harness.ring_filter(job, scene, n)
    assert usable
    passed = one >= 0
    assert passed.sum() > 0.25 * n
    for shift in (0, 1, 17, 99_991):
        lo, hi = harness.ring_filter_pairs(job, scene, n, shift)
        assert np.array_equal(lo[passed], one[passed]), shift
        assert np.array_equal(hi[passed], one[passed]), shift


@pytest.mark.parametrize("k", range(1, 40, 3))
def test_two_rays_per_lane_on_random_scenes(orc, harness, k):
    scene, _, kw = random_case(orc, k)
    kw.pop("use_bottle")
    n = 60_000
    job = abi.default_job(1, n, first_ray=11 * 10 ** 9 + k, **kw)
    one, usable = harness.ring_filter(job, scene, n)
    if not usable:
        pytest.skip("the launcher does not run the filter on this scene")
    passed = one >= 0
    for shift in (1, 29_989):
        lo, hi = harness.ring_filter_pairs(job, scene, n, shift)
        assert np.array_equal(lo[passed], one[passed]) and np.array_equal(hi[passed], one[passed]), shift
