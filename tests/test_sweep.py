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
    assert len(sweep.experiment_cases("lens")) == 5 * 5 * 3      # runner.py:394-397: three bottles
    assert len(sweep.experiment_cases("bessel")) == 4
    isb = sweep.experiment_cases("isb")
    assert len(isb) == 14 and [c["source_type"] for c in isb] == ["isors"] * 7 + ["point"] * 7
    assert isb[6]["isors_offset"] == 1.5e-3 and isb[13]["bottle_z"] == 1.5e-3
    lens = sweep.experiment_cases("lens")
    assert {c["l2"] for c in lens} == {"planoConvex-f%smm.params" % f for f in ("59.8", "49.8", "39.9", "34.9", "29.9")}
    assert all(os.path.exists(os.path.join(RES, c["l3"])) and os.path.exists(os.path.join(RES, c["l2"])) for c in lens)
    with pytest.raises(ValueError):
        sweep.experiment_cases("bessel_params")


def test_bessel_bottle_position():
    """runner.py:280-312 with the shipped files: fb = 35.7 mm, radius a = 17.5 mm."""
    import math
    z = sweep.bessel_bottle_z(RES, 0.75e-3)
    want = 35.7e-3 * (0.75e-3 + 0.5e-3) / (97.3e-3 * math.tan(math.radians(5.0) * 0.45)) - 17.5e-3
    assert z == pytest.approx(want, rel=1e-14)


@pytest.mark.gpu
def test_batched_sweep_equals_single_runs(ort, orc, tmp_path):
    n = 60_000
    cases = sweep.experiment_cases("iris")[:14] + sweep.experiment_cases("lens")[36:44] + \
        sweep.experiment_cases("offset")
    assert cases[-1]["bottle"].endswith("_-16mm.params")
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


@pytest.mark.gpu
def test_isors_vs_bessel_sweep(ort, orc, tmp_path):
    n = 50_000
    res = sweep.run_sweep(sweep.experiment_cases("isb"), RES, str(tmp_path), nphotons=n, verbose=False)
    assert len(res) == 14 and all(r is not None for r in res)
    zs = [r["point_scene"].bottle.centre[2] for r in res[7:]]
    assert zs == sorted(zs) and len(set(zs)) == 7            # the generated bottle files were read
    assert all(r["point_scene"].bottle.centre[2] == 0.0 for r in res[:7])
    for r in res[::2]:
        st = r["st"]
        for phase, key, cnt in ((1, "ring", "rcount"), (2, "point", "pcount")):
            job = ort.job_from_settings(st, phase)
            scene = r["ring_scene"] if phase == 1 else r["point_scene"]
            oimg, olost, _ = orc.trace(job, scene)
            assert np.array_equal(oimg[0], r[key]) and int(olost[0]) == r[cnt]
    lines = open(tmp_path / "iSORS_vs_Bessel" / "trans-stats.dat").read().splitlines()
    assert len(lines) == 15
    assert sum(",isors," in ln for ln in lines) == 7


@pytest.mark.gpu
def test_bessel_sweep(ort, orc, tmp_path):
    """-b: the image source; the reference ships no bessel-smear.dat, so a synthetic one goes into a
    scratch res directory next to links to the shipped files."""
    resdir = tmp_path / "res"
    resdir.mkdir()
    for f in os.listdir(RES):
        os.symlink(os.path.join(RES, f), resdir / f)
    yy, xx = np.mgrid[0:512, 0:512]
    r = np.hypot(xx - 255.5, yy - 255.5)
    np.exp(-((r - 90.0) / 12.0) ** 2).astype(np.float64).tofile(resdir / "bessel-smear.dat")
    n = 80_000
    res = sweep.run_sweep(sweep.experiment_cases("bessel"), str(resdir), str(tmp_path / "data"),
                          nphotons=n, verbose=False)
    try:
        orc.set_image_source(orc.load_image_source(str(resdir / "bessel-smear.dat"), n))
        for rr in res:
            job = ort.job_from_settings(rr["st"], 2)
            oimg, olost, _ = orc.trace(job, rr["point_scene"])
            assert np.array_equal(oimg[0], rr["point"]) and int(olost[0]) == rr["pcount"]
    finally:
        orc.set_image_source(None)
        ort.set_image_source(None)
    assert sum(f.endswith("_image-point.dat") for f in os.listdir(tmp_path / "data" / "images")) == 4
