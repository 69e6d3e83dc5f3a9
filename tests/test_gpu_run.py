"""End to end on the GPU box: the Python host (opticalraytrace_b200.run) and the C++ `raytrace`
program, driven like the reference binary (`cd bin && ./raytrace <settings>`), produce the
reference's output files; their content equals the oracle's."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "res")


def _settings_file(tmp_path, n, folder="run1", bottle="clearBottle-small.params"):
    lines = open(os.path.join(RES, "settings-config1.params")).read().splitlines()
    lines[2] = "%d   # rays" % n
    lines[13] = bottle
    lines[17] = folder
    res = tmp_path / "res"
    shutil.copytree(RES, res)
    (res / "job.params").write_text("\n".join(lines) + "\n")
    return res


def _oracle_images(orc, n, files):
    ring = orc.trace(abi.default_job(1, n), orc.make_scene(*files))
    point = orc.trace(abi.default_job(2, n), orc.make_scene(*files, lens_wavelength=843e-9))
    return ring, point


def test_python_host_run(ort, orc, tmp_path):
    import opticalraytrace_b200 as pkg
    n = 300_000
    res = _settings_file(tmp_path, n)
    out = pkg.run(str(res / "job.params"), str(res), str(tmp_path / "data"), verbose=False)
    ring, point = _oracle_images(orc, n, cases.C1)
    assert np.array_equal(out["ring"], ring[0][0]) and np.array_equal(out["point"], point[0][0])
    assert out["rcount"] == ring[1][0] and out["pcount"] == point[1][0]
    base = os.path.join(out["folder"], out["name"] + "_image")
    for suffix, want in (("-ring.dat", ring[0][0]), ("-point.dat", point[0][0]),
                         ("-total.dat", ring[0][0] + point[0][0])):
        raw = np.fromfile(base + suffix, dtype=np.float64)
        assert raw.size == 401 * 401
        assert np.array_equal(raw.reshape(401, 401), want.astype(np.float64))
    assert os.path.exists(os.path.join(out["folder"], "trans-stats.dat"))


def test_cpp_raytrace_binary(orc, tmp_path):
    n = 200_000
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "src"), "-s"])
    res = _settings_file(tmp_path, n, folder="cpp")
    bindir = tmp_path / "bin"
    bindir.mkdir()
    shutil.copy(os.path.join(ROOT, "src", "raytrace"), bindir / "raytrace")
    env = dict(os.environ, ORT_NUM_GPUS="1",
               LD_LIBRARY_PATH=os.path.join(ROOT, "opticalraytrace_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run(["./raytrace", "job.params"], cwd=bindir, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    ring, point = _oracle_images(orc, n, cases.C1)
    assert " Using job.params settings." in p.stdout
    assert "Ring  transmitted:  %8.2f%%" % (100 * (1 - ring[1][0] / n)) in p.stdout
    assert "Point transmitted:  %8.2f%%" % (100 * (1 - point[1][0] / n)) in p.stdout
    folder = tmp_path / "data" / "cpp"
    files = sorted(os.listdir(folder))
    assert "trans-stats.dat" in files and len(files) == 4
    ringf = [f for f in files if f.endswith("_image-ring.dat")][0]
    assert ringf.startswith("point_bottle_T_Ra_0.01750_Rb_0.01750_offset_0.00000__F_F_1.00000_L2f_0.0399_L3f_0.0500")
    raw = np.fromfile(folder / ringf, dtype=np.float64).reshape(401, 401)
    assert np.array_equal(raw, ring[0][0].astype(np.float64))
    total = np.fromfile(folder / ringf.replace("-ring", "-total"), dtype=np.float64)
    assert total.sum() == ring[0].sum() + point[0].sum()
    # unsupported source types stop with a message instead of silently running something else
    lines = (res / "job.params").read_text().splitlines()
    lines[10] = "image"
    (res / "img.params").write_text("\n".join(lines) + "\n")
    p = subprocess.run(["./raytrace", "img.params"], cwd=bindir, env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "cannot open image source" in p.stderr   # no bessel-*.dat is shipped
    # the shipped settings.params (crs source, 14-line bottle file, tracker on) runs as is
    p = subprocess.run(["./raytrace", "settings.params"], cwd=bindir, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Ring  transmitted:" in p.stdout and "Deselecting makeImages" in p.stdout
    out = tmp_path / "data" / "settings-testysors"
    names = sorted(os.listdir(out))
    assert [n.split("-")[-1] for n in names if "trace" in n] == ["pointtrace.dat", "ringtrace.dat"]
    assert not any(n.endswith("_image-ring.dat") for n in names)   # tracker => no images


def test_all_kernel_variants_bounds_checked():
    """Every kernel variant once, through the assert-instrumented build (make DEBUG=1): the
    project's own substitute for compute-sanitizer, which is closed on this GPU pool."""
    dbg = os.path.join(ROOT, "opticalraytrace_b200", "libort_debug.so")
    if not os.path.exists(dbg):
        pytest.skip("libort_debug.so not built (make -C opticalraytrace_b200/csrc DEBUG=1)")
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")],
                       env=dict(os.environ, ORT_LIB=dbg), capture_output=True, text=True)
    assert p.returncode == 0 and "sanitize_case ok" in p.stdout, p.stdout + p.stderr


def _read_tracks(path):
    """debug-plot.py's reader, condensed: blocks of position lines separated by blank lines."""
    rays, cur, blanks = [], [], 0
    for line in open(path):
        if len(line) > 3:
            cur.append([float(v) for v in line.split()])
            blanks = 0
        else:
            if blanks == 0 and cur:
                rays.append(cur)
            cur = []
            blanks += 1
    return rays


def test_tracker_files(ort, orc, tmp_path):
    """reference src/stackMod.f90 + src/main.f90:103-107,144-160: popped stack (image plane first),
    `3(F10.7,1x)` lines, three blank lines per ray; lens-lost rays leave only blanks."""
    n = 500
    for phase in (1, 2):
        scene = cases.scene_for(orc, cases.C1, phase)
        job = abi.default_job(phase, n)
        path = str(tmp_path / ("p%d.dat" % phase))
        ort.write_tracks(job, scene, path)
        text = open(path).read().splitlines()
        assert all(len(l) == 33 or l == "  " for l in text)
        full = orc.trace_rays(job, scene, n)
        st = full["status"]
        survivors = np.flatnonzero((st == 0) | (st >= 21) & (st <= 23))
        bottle_lost = np.flatnonzero((st >= 1) & (st <= 8))
        rays = _read_tracks(path)
        assert len(rays) == len(survivors) + len(bottle_lost)
        k = 0
        for i in range(n):
            if i in survivors:
                r = rays[k]; k += 1
                assert len(r) == (5 if phase == 2 else 4)
                assert np.allclose(r[0], full["pos"][:, i], atol=6e-8)            # image plane first
                src = orc.trace_rays(abi.default_job(phase, stop_after=1, first_ray=i), scene, 1)
                assert np.allclose(r[-1], src["pos"][:, 0], atol=6e-8)            # source last
                zs = [p[2] for p in r]
                assert zs == sorted(zs, reverse=True)
            elif i in bottle_lost:
                r = rays[k]; k += 1
                assert len(r) == 2 and np.allclose(r[0], full["pos"][:, i], atol=6e-8)
        blank = sum(1 for l in text if l == "  ")
        assert blank == 3 * n + 3 * len(bottle_lost)
    from opticalraytrace_b200.lib import OrtError
    with pytest.raises(OrtError, match="Too many photons"):
        ort.write_tracks(abi.default_job(1, 10001), scene, str(tmp_path / "x.dat"))
