#!/usr/bin/env python3
"""The 75-case L2 x L3 focal-length experiment of runner.py (:231-261: 5 doublets x 5 plano-convex lenses x
3 bottles) as ONE batched ort_trace call per ray loop, at a small ray count per case (a quick-look sweep):
rays/s with the scenes overlapped on several streams (default) and back to back on one
(ORT_FLAG_ONE_LANE), and the same cases as 75 separate calls.

    python tools/sweep_bench.py [--rays 1000000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from opticalraytrace_b200 import abi, lib, sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    lib.init(1)
    cases = sweep.experiment_cases("lens")
    out = {"cases": len(cases), "rays_per_case": a.rays}
    for phase, name in ((1, "ring"), (2, "point")):
        scenes = []
        for kw in cases:
            kw = {k: v for k, v in kw.items() if k in ("bottle", "l2", "l3", "use_bottle")}
            st = lib.make_settings(nphotons=a.rays, **kw)
            scenes.append(lib.build_scene(st, os.path.join(ROOT, "res"), None if phase == 1 else 843e-9)[0])
        ref = None
        for label, flags, batched in (("lanes", 0, True), ("one_lane", abi.FLAG_ONE_LANE, True), ("separate_calls", 0, False)):
            best_dev, best_wall, img0 = 1e30, 1e30, None
            for rep in range(a.reps + 1):
                job = lib.job_from_settings(st, phase)
                job.nrays = a.rays
                job.flags |= flags
                t0 = time.perf_counter()
                if batched:
                    img, lost, hist, tm = lib.trace(job, scenes)
                    dev = tm.trace_seconds
                else:
                    dev, imgs = 0.0, []
                    for sc in scenes:
                        i1, _, _, tm = lib.trace(job, sc)
                        dev += tm.trace_seconds
                        imgs.append(i1[0])
                    img = np.stack(imgs)
                wall = time.perf_counter() - t0
                if rep:     # the first repetition warms up (allocations, lazy module load)
                    best_dev, best_wall = min(best_dev, dev), min(best_wall, wall)
                img0 = img
            if ref is None:
                ref = img0
            assert np.array_equal(ref, img0), "the three ways must give identical images"
            total = float(a.rays) * len(scenes)
            out["%s_%s" % (name, label)] = {"rays_per_s_device": total / best_dev, "rays_per_s_wall": total / best_wall,
                                            "ms_device": best_dev * 1e3, "ms_wall": best_wall * 1e3}
    lib.finalize()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
