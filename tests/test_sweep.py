"""Sweep front-end: runner.py's experiments as batched ort_trace calls."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi, sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "res")


def test_experiment_case_lists():
    """Same case counts as the loops of the reference's runner.py (:113-261)."""
    assert len(sweep.experiment_cases("point")) == 4
    assert len(sweep.experiment_cases("spot")) == 4
    assert len(sweep.experiment_cases("iris")) == 4 * (5 + 5 + 1)
    assert len(sweep.experiment_cases("offset")) == 7
    assert len(sweep.experiment_cases("lens")) == 5 * 5 * 4
    lens = sweep.experiment_cases("lens")
    assert {c["l2"] for c in lens} == {"planoConvex-f%smm.params" % f for f in ("59.8", "49.8", "39.9", "34.9", "29.9")}
    assert all(os.path.exists(os.path.join(RES, c["l3"])) and os.path.exists(os.path.join(RES, c["l2"])) for c in lens)
    with pytest.raises(ValueError):
        sweep.experiment_cases("bessel")


@pytest.mark.gpu
def test_batched_sweep_equals_single_runs(ort, orc, tmp_path):
    n = 60_000
    cases = sweep.experiment_cases("iris")[:14] + sweep.experiment_cases("lens")[36:44] + \
        sweep.experiment_cases("offset")
    res = sweep.run_sweep(cases, RES, str(tmp_path), nphotons=n, verbose=False)
    assert res[-1] is None                                   # the -16 mm file does not exist
    done = [r for r in res if r is not None]
    assert len(done) == len(cases) - 1
    for r in done[::3]:
        st = r["st"]
        for phase, key, cnt in ((1, "ring", "rcount"), (2, "point", "pcount")):
            job = ort.job_from_settings(st, phase)
            scene = r["ring_scene"] if phase == 1 else r["point_scene"]
            img, lost, hist, _ = ort.trace(job, scene)
            assert np.array_equal(img[0], r[key]) and int(lost[0]) == r[cnt]
            oimg, olost, _ = orc.trace(job, scene)
            assert np.array_equal(oimg[0], r[key])
    # per-case outputs: one trans-stats line per case and folder, images unless deselected
    for folder, count in (("iris", 14), ("images-lens", 8), ("images-offset", 6)):
        lines = open(tmp_path / folder / "trans-stats.dat").read().splitlines()
        assert len(lines) == count + 1
    assert not any(f.endswith(".dat") and "_image" in f for f in os.listdir(tmp_path / "images-lens"))
    assert sum(f.endswith("_image-total.dat") for f in os.listdir(tmp_path / "iris")) == 14


@pytest.mark.gpu
def test_spot_sweep_writes_tracks(ort, tmp_path):
    res = sweep.run_sweep(sweep.experiment_cases("spot")[:2], RES, str(tmp_path), verbose=False)
    files = os.listdir(tmp_path / "spot-diag")
    assert sum(f.endswith("-pointtrace.dat") for f in files) == 2
    assert sum(f.endswith("-ringtrace.dat") for f in files) == 2
    assert not any("_image" in f for f in files)
