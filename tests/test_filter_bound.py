"""The proof obligation of the ring loop's single-precision filter (ort_filter.cuh):
every quantity the filter holds an error bound for must lie within that bound of the value exact
arithmetic gives.  The host harness runs the filter with its trace next to a double-precision twin
(tests/host_harness.cpp: twin_filter) and compares them record by record -- positions, directions,
normals, N.I, sin^2, cos^2 theta_t, cos theta_t, the Fresnel decision value, h, c, the discriminant,
the path length, rho^2 -- on every ray that passes stage A, for the shipped set-ups, 40 randomised
and 8 extreme scenes.  The `fuzz` build pushes every MUFU stand-in to +- the error the bounds assume
(rule R4), so the bounds are exercised at their limit; the device's own MUFU errors are measured
exhaustively by ort_mufu_selftest (GPU test below) and must stay below half of what is assumed."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.test_fuzz_scenes import random_case
from tests.test_ring_filter import EXTREMES, RING_SETUPS


def _check(harness, job, scene, n, fuzz):
    usable, ratios, c = harness.filter_bounds(job, scene, n, fuzz=fuzz)
    if usable is None:      # L2 moved out of the aim plane: the launcher never runs the filter here
        return usable, ratios, c
    assert c["violations"] == 0, (ratios, c)
    assert c["wrong"] == 0, c
    assert all(r <= 1.0 for r in ratios.values()), ratios
    return usable, ratios, c


@pytest.mark.parametrize("fuzz", [False, True], ids=["exact-mufu", "worst-mufu"])
@pytest.mark.parametrize("k", range(len(RING_SETUPS)))
def test_bounds_hold_on_shipped_setups(orc, harness, k, fuzz):
    files, kw = RING_SETUPS[k]
    scene = cases.scene_for(orc, files, 1)
    job = abi.default_job(1, first_ray=11 * 10 ** 9 * k, **kw)
    usable, ratios, c = _check(harness, job, scene, 1_000_000, fuzz)
    assert usable
    assert c["records"] > 5 * c["passed"]          # the comparison is not vacuous
    assert c["called"] > 0.15 * c["passed"]        # ... and the filter does decide
    # the bounds are bounds, not estimates: the worst observed error stays well inside them
    assert max(ratios.values()) < 0.6, ratios


@pytest.mark.parametrize("k", range(40))
def test_bounds_hold_on_random_scenes(orc, harness, k):
    scene, phase, kw = random_case(orc, k)
    kw.pop("use_bottle")
    job = abi.default_job(1, first_ray=k * 10 ** 9, **kw)
    _check(harness, job, scene, 150_000, fuzz=bool(k & 1))


@pytest.mark.parametrize("k", range(len(EXTREMES)))
def test_bounds_hold_on_extreme_geometries(orc, harness, k):
    name, tweak = EXTREMES[k]
    scene = cases.scene_for(orc, cases.C2, 1)
    tweak(scene)
    job = abi.default_job(1, first_ray=5 * 10 ** 8 * k)
    usable, ratios, c = _check(harness, job, scene, 300_000, fuzz=True)
    print(name, usable, c, max(ratios.values()))


def test_far_from_the_origin_the_bounds_grow_and_still_hold(orc, harness):
    """fp32 coordinates are absolute: a system 50 m from the origin rounds every coordinate by ~4e-6 m.
    The bounds say so (every constant carries the u |centre| terms): the filter then proves little
    or is switched off by the launcher, but what it does prove is still right."""
    for shift in (0.5, 5.0, 50.0):
        scene = cases.scene_for(orc, cases.C2, 1)
        for c in (scene.bottle.centre, scene.L2.centre, scene.L3.centre1, scene.L3.centre2, scene.L3.centre3):
            c[2] += shift
        scene.L2.fb += shift
        scene.img_plane += shift
        job = abi.default_job(1)
        usable, ratios, c = _check(harness, job, scene, 300_000, fuzz=True)
        print(shift, usable, c)
    assert not usable


@pytest.mark.gpu
def test_mufu_errors_are_below_half_of_what_the_bounds_assume(ort):
    """Rule R4: all 2^32 fp32 arguments through rcp / rsqrt / sqrt / sin / cos .approx.ftz.f32 on the
    device, against fp64."""
    worst, assumed = ort.mufu_selftest()
    print("measured", worst)
    print("assumed ", assumed)
    for name in worst:
        assert 0.0 < worst[name] <= 0.5 * assumed[name], (name, worst[name], assumed[name])
