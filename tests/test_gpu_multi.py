"""Multi-GPU paths (skipped on a one-GPU box): single-process mode (`ort_init(G)`, ncclCommInitAll,
one ncclReduce inside ort_trace) gives the same image as one device, bit for bit."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases

pytestmark = pytest.mark.gpu


def test_single_process_multi_device_equals_one_device(ort, orc):
    ortlib = ort
    if ortlib.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2_000_001
    results = {}
    try:
        for g in (1, 2, min(ortlib.device_count(), 4)):
            ortlib.init(g)   # re-initialises (finalizes the previous mode first)
            for phase in (1, 2):
                scene = cases.scene_for(orc, cases.C2, phase)
                img, lost, hist, tm = ortlib.trace(abi.default_job(phase, n), scene)
                results[(g, phase)] = (img, lost, hist)
                if g > 1:
                    assert tm.reduce_seconds > 0
    finally:
        ortlib.init(1)       # what the session fixture promised the other tests
    for (g, phase), (img, lost, hist) in results.items():
        ref = results[(1, phase)]
        assert np.array_equal(img, ref[0]) and np.array_equal(hist, ref[2]) and np.array_equal(lost, ref[1])
    oimg, olost, ohist = orc.trace(abi.default_job(2, n), cases.scene_for(orc, cases.C2, 2))
    assert np.array_equal(results[(1, 2)][0], oimg)
