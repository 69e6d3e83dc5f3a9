#!/usr/bin/env python3
"""Finds the ring-loop rays that sit ON the integer aperture cut of the CUDA kernels' stage A.

Stage A (ort_ring_quads_pass, ort_kernels.cuh) ends a ray when the high word of its aim-disc r^2 draw is
above the high word of the cut (ort_ring_aim_cut) and lets it pass when it is below; a ray whose word
EQUALS the cut's -- 2^-32 of all rays -- takes a path of its own (the all-fp64 kernel evaluates the whole
expression, the culling kernel hands the ray to fp64).  A parity run of 1e6 rays never meets one, so
their indices are found here by brute force (oracle/orc_find_aim_word, the first 2^35 rays of each
shipped set-up, ~1 min on 8 cores) and committed as tests/golden/edge_rays_v2.json;
tests/test_gpu_parity.py::test_rays_on_the_aperture_cut traces a few rays around each.

    python tests/golden/make_edge_rays.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from opticalraytrace_b200 import abi  # noqa: E402
from tests import cases, oracle_lib as orc  # noqa: E402

SPAN = 1 << 35


def main():
    L = orc.lib()
    L.orc_find_aim_word.restype = C.c_int64
    L.orc_find_aim_word.argtypes = [C.c_uint64, C.c_int32, C.c_uint32, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests"), "-s"])
    H = C.CDLL(os.path.join(ROOT, "tests", "libhost_harness.so"))
    H.hh_ring_aim_cut.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.POINTER(C.c_int)]
    H.hh_ring_aim_cut.restype = C.c_uint64
    out = {"span": SPAN, "cases": {}}
    for name, files in (("c1", cases.C1), ("c2", cases.C2)):
        scene = cases.scene_for(orc, files, 1)
        job = abi.default_job(1, 1)
        have = C.c_int(0)
        cut = H.hh_ring_aim_cut(C.byref(job), C.byref(scene), C.byref(have))
        assert have.value
        ids = np.zeros(64, np.int64)
        n = L.orc_find_aim_word(job.seed, 1, cut >> 32, 0, SPAN, ids.ctypes.data, 64)
        out["cases"][name] = {"seed": int(job.seed), "cut_hi": int(cut >> 32), "rays": sorted(int(i) for i in ids[:min(n, 64)])}
        print(name, "cut_hi", hex(cut >> 32), "found", n, out["cases"][name]["rays"])
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "edge_rays_v2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
