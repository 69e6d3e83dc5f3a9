#!/usr/bin/env python3
"""Writes the committed golden fixtures tests/golden/*.npz.

Provenance: the reference (Fortran) cannot be built in this image and ships no golden data, so
these vectors come from the CPU oracle (oracle/ort_oracle.cpp, g++ -O2 -ffp-contract=off) after
it was pinned by tests/test_oracle_kat.py.  They freeze today's oracle so that a later edit of
the oracle, the generator's slot map or the params library shows up as a diff, and they give the
GPU tests a fixture that does not depend on building the oracle on the GPU box.

  python tests/golden/make_golden.py        (from the repo root)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from opticalraytrace_b200 import abi  # noqa: E402
from tests import cases, oracle_lib as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NRAYS = 96
NIMG = 200_000


def main():
    rays, images = {}, {}
    for cid, files, phase, kw in cases.RAY_CASES + cases.SCATTER_CASES + cases.SOURCE_CASES:
        scene = cases.scene_for(O, files, phase, kw)
        job = abi.default_job(phase, **kw)
        r = O.trace_rays(job, scene, NRAYS)
        for k in ("pos", "dir", "status", "bin"):
            rays["%s/%s" % (cid, k)] = r[k]
        job = abi.default_job(phase, NIMG, **kw)
        img, lost, hist = O.trace(job, scene)
        nz = np.flatnonzero(img[0])
        images["%s/idx" % cid] = nz.astype(np.int32)
        images["%s/cnt" % cid] = img[0].ravel()[nz].astype(np.int32)
        images["%s/hist" % cid] = hist[0]
        images["%s/lost" % cid] = lost
    np.savez_compressed(os.path.join(HERE, "rays_v2.npz"), **rays)
    np.savez_compressed(os.path.join(HERE, "images_v2.npz"), **images)
    # first uniforms of three rays of each phase: freezes the generator + slot map
    u = {"p%d/r%d" % (p, r): O.uniforms(123456789, p, r, 0, 24)
         for p in (1, 2) for r in (0, 1, 2 ** 33 + 5)}
    np.savez_compressed(os.path.join(HERE, "uniforms_v2.npz"), **u)
    for f in ("rays_v2.npz", "images_v2.npz", "uniforms_v2.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
