#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both loops, every kernel
variant (ring shortcut, clear bottle, scatter bottle, crs / isors / spot emitters, flat kernel,
explicit-ray kernel), 1e5 rays each.  usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opticalraytrace_b200 import abi, lib  # noqa: E402

RES = os.path.join(ROOT, "res")
lib.init(1)
n = 100_000
total = 0
for bottle, src in (("clearBottle-large.params", "point"), ("scatterBottle-small.params", "point"),
                    ("clearBottle-ellipse-long.params", "crs"), ("clearBottle-small.params", "isors"),
                    ("clearBottle-small.params", "spot")):
    st = lib.make_settings(bottle, nphotons=n, source_type=src)
    for phase, lam in ((1, None), (2, 843e-9)):
        scene, _ = lib.build_scene(st, RES, lam)
        job = lib.job_from_settings(st, phase)
        for flags in (0, abi.FLAG_NO_COMPACTION):
            job.flags = flags
            img, lost, hist, _ = lib.trace(job, [scene, scene], allow_trap=True)
            assert int(hist[0, :27].sum()) == n and int(img[0].sum()) == int(hist[0, 0])
            total += int(hist[0, 0])
        r = lib.trace_rays(job, scene, 4096)
        assert r["status"].min() >= 0
lib.math_selftest(1 << 16)
lib.measure_fp64_peak()
lib.finalize()
print("sanitize_case ok, binned", total)
