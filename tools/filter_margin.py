#!/usr/bin/env python3
"""How much margin does the ring loop's fp32 culling filter have?  

CPU part (default): builds the test-only host harness with ORT_FILTER_TOL = 0, 1e-7, 1e-5 and the
shipped value, and counts, over shipped and randomised geometries, how many filter verdicts
disagree with the oracle and how many rays are handed to fp64.  With NO margin fp32 misjudges a few
rays in 1e7; with the shipped margin none, at the price of ~1 % unnecessary fp64 passes.

GPU part (--gpu RAYS): ORT_FLAG_VERIFY_FILTER over RAYS rays per set-up on the device: the kernel
runs filter and fp64 on every ray and counts disagreements.  Build variants of libort.so with
-DORT_FILTER_TOL=... / -DORT_FILTER_COND_I=... and point ORT_LIB at them to map the wrong-verdict
rate against the margin (profiles/r01_filter_margin.txt).

    python tools/filter_margin.py [--rays 4000000]
    python tools/filter_margin.py --gpu 100000000000
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from opticalraytrace_b200 import abi  # noqa: E402
from tests import cases, oracle_lib as orc  # noqa: E402
from tests.test_fuzz_scenes import random_case  # noqa: E402


def setups():
    out = [("shipped %d" % i, cases.scene_for(orc, f, 1), {}) for i, f in
           enumerate((cases.C1, cases.C2, cases.ELL, cases.OTHER))]
    for k in (1, 2, 5, 7, 10, 11, 13, 14):
        sc, _, kw = random_case(orc, k)
        kw.pop("use_bottle")
        out.append(("random %d" % k, sc, kw))
    return out


def cpu(nrays):
    cases_ = setups()
    refs = [orc.trace_rays(abi.default_job(1, **kw), sc, nrays)["status"] for _, sc, kw in cases_]
    with tempfile.TemporaryDirectory() as tmp:
        for tol in ("0.0f", "1e-7f", "1e-5f", None):
            so = os.path.join(tmp, "hh_%s.so" % (tol or "shipped"))
            cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-march=native", "-ffp-contract=off",
                   "-shared", "-o", so, os.path.join(ROOT, "tests", "host_harness.cpp")]
            if tol:
                cmd.insert(1, "-DORT_FILTER_TOL=" + tol)
            subprocess.check_call(cmd)
            H = C.CDLL(so)
            H.hh_ring_filter.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64, C.c_void_p]
            wrong = back = inb = 0
            for (name, sc, kw), ref in zip(cases_, refs):
                job = abi.default_job(1, **kw)
                v = np.zeros(nrays, np.int32)
                if not H.hh_ring_filter(C.byref(job), C.byref(sc), nrays, v.ctypes.data):
                    continue            # L2 outside the aim plane: the filter is not used
                wrong += int(((v > 0) & (v != ref)).sum())
                back += int((v == 0).sum())
                inb += int((v != -1).sum())
            print("margin %-8s wrong verdicts %6d   handed to fp64 %9d of %d rays past L2's aperture"
                  % (tol or "shipped", wrong, back, inb))


def gpu(nrays, first_ray=0, only=None):
    from opticalraytrace_b200 import lib
    lib.init(1)
    try:
        for name, sc, kw in setups():
            if only and not name.startswith(only):
                continue
            job = abi.default_job(1, nrays, first_ray=first_ray, flags=kw.pop("flags", 0) | abi.FLAG_VERIFY_FILTER, **kw)
            _, _, hist, tm = lib.trace(job, sc, want_image=False, allow_trap=True)
            print("%-10s %.3g rays: filter called %d, wrong %d  (%.1f s)"
                  % (name, nrays, hist[0, abi.FILTER_SLOT_CALLED], hist[0, abi.FILTER_SLOT_WRONG], tm.trace_seconds))
    finally:
        lib.finalize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--rays", type=int, default=4_000_000)
    ap.add_argument("--gpu", type=int, default=0, metavar="RAYS")
    ap.add_argument("--first-ray", type=int, default=0, help="start of the ray-index range (GPU part)")
    ap.add_argument("--only", default=None, help="only the set-ups whose name starts with this (GPU part)")
    a = ap.parse_args()
    gpu(a.gpu, a.first_ray, a.only) if a.gpu else cpu(a.rays)
