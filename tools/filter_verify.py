#!/usr/bin/env python3
"""Evidence for the ring loop's single-precision filter (ort_filter.cuh).

CPU part (default): for the shipped, 40 randomised and 8 extreme scenes, runs the filter with its
trace next to the double-precision twin (tests/host_harness.cpp) with the MUFU stand-ins pushed to
the assumed error, and prints per scene the largest |fp32 - exact| / bound per kind of quantity, the
number of bound violations and wrong verdicts (both must be 0), and how many of the rays that pass
stage A the filter calls.

GPU part (--gpu RAYS): ORT_FLAG_VERIFY_FILTER over RAYS rays per scene on the device -- the culling
kernel lists every ray that passes stage A together with the verdict of its own filter (two rays per
lane, packed arithmetic), the fp64 kernel traces them all and counts verdicts that differ -- plus the
exhaustive MUFU error measurement (ort_mufu_selftest).

    python tools/filter_verify.py [--rays 1000000]
    python tools/filter_verify.py --gpu 300000000000 [--first-ray N] [--only shipped]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from opticalraytrace_b200 import abi  # noqa: E402
from tests import cases, oracle_lib as orc  # noqa: E402
from tests.test_fuzz_scenes import random_case  # noqa: E402
from tests.test_ring_filter import EXTREMES, RING_SETUPS  # noqa: E402


def setups():
    out = []
    for i, (f, kw) in enumerate(RING_SETUPS):
        out.append(("shipped %d" % i, cases.scene_for(orc, f, 1), dict(kw)))
    for k in range(40):
        sc, _, kw = random_case(orc, k)
        kw.pop("use_bottle")
        out.append(("random %d" % k, sc, kw))
    for k, (name, tweak) in enumerate(EXTREMES):
        sc = cases.scene_for(orc, cases.C2, 1)
        tweak(sc)
        out.append(("extreme %d (%s)" % (k, name), sc, {}))
    return out


def cpu(nrays):
    import tests.conftest as cf
    harness = cf.harness.__wrapped__()
    tot = dict(records=0, violations=0, called=0, wrong=0, passed=0)
    for name, sc, kw in setups():
        job = abi.default_job(1, **kw)
        usable, ratios, c = harness.filter_bounds(job, sc, nrays, fuzz=True)
        if usable is None:
            print("%-40s L2 is not in the aim plane: the launcher never runs the filter" % name)
            continue
        for k in tot:
            tot[k] += c[k]
        worst = max(ratios, key=ratios.get)
        print("%-40s usable %d  records %9d  violations %d  called %7d of %7d  wrong %d  worst err/bound %.3f (%s)"
              % (name, usable, c["records"], c["violations"], c["called"], c["passed"], c["wrong"], ratios[worst], worst))
    print("TOTAL", tot)


def gpu(nrays, first_ray=0, only=None, shipped_factor=1):
    from opticalraytrace_b200 import lib
    lib.init(1)
    try:
        worst, assumed = lib.mufu_selftest()
        for k in worst:
            print("MUFU %-6s measured %.4g = 2^%.2f   assumed %.4g   ratio %.3f"
                  % (k, worst[k], __import__("math").log2(worst[k]), assumed[k], worst[k] / assumed[k]))
        called = wrong = 0
        for name, sc, kw in setups():
            if only and not name.startswith(only):
                continue
            n = nrays * (shipped_factor if name.startswith("shipped") else 1)
            job = abi.default_job(1, n, first_ray=first_ray, flags=kw.pop("flags", 0) | abi.FLAG_VERIFY_FILTER, **kw)
            _, _, hist, tm = lib.trace(job, sc, want_image=False, allow_trap=True)
            c, w = int(hist[0, abi.FILTER_SLOT_CALLED]), int(hist[0, abi.FILTER_SLOT_WRONG])
            called += c
            wrong += w
            print("%-40s %.3g rays: filter called %d, wrong %d  (%.1f s)" % (name, n, c, w, tm.trace_seconds), flush=True)
        print("TOTAL verdicts %d wrong %d" % (called, wrong))
    finally:
        lib.finalize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--rays", type=int, default=1_000_000)
    ap.add_argument("--gpu", type=int, default=0, metavar="RAYS")
    ap.add_argument("--first-ray", type=int, default=0, help="start of the ray-index range (GPU part)")
    ap.add_argument("--only", default=None, help="only the set-ups whose name starts with this (GPU part)")
    ap.add_argument("--shipped-factor", type=int, default=1, help="the shipped set-ups get this many times RAYS (GPU part)")
    a = ap.parse_args()
    gpu(a.gpu, a.first_ray, a.only, a.shipped_factor) if a.gpu else cpu(a.rays)
