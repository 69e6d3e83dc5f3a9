#!/usr/bin/env python3
"""CPU baseline figures of BASELINE.md section 2: the C++/OpenMP oracle (restatement of the Fortran
path) on 1 thread (config 1, mirrors `./install.sh -n 1`) and on all host cores (config 2)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opticalraytrace_b200 import abi  # noqa: E402
from tests import cases, oracle_lib as O  # noqa: E402


def rate(files, phase, n, threads):
    scene = cases.scene_for(O, files, phase)
    O.trace(abi.default_job(phase, n // 10), scene, nthreads=threads)
    t0 = time.perf_counter()
    O.trace(abi.default_job(phase, n, first_ray=1 << 36), scene, nthreads=threads)
    return n / (time.perf_counter() - t0)


if __name__ == "__main__":
    cores = os.cpu_count()
    out = {"cores": cores, "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")}
    out["config1_ring_1thread"] = rate(cases.C1, 1, 10_000_000, 1)
    out["config1_point_1thread"] = rate(cases.C1, 2, 10_000_000, 1)
    out["config2_ring_allcores"] = rate(cases.C2, 1, 100_000_000, cores)
    out["config2_point_allcores"] = rate(cases.C2, 2, 100_000_000, cores)
    print(json.dumps(out))
