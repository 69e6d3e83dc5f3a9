"""Host side of the drop-in surface: params readers, prologue scalars, output names and files
(no GPU needed).  The product's readers (libort.so) are checked against the oracle's."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "res")


def test_res_library_regenerates():
    """res/ is exactly what tools/make_res.py writes (the committed fixtures are not hand-edited)."""
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location("make_res", os.path.join(ROOT, "tools", "make_res.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with tempfile.TemporaryDirectory() as d:
        m.main(d)
        names = sorted(os.listdir(d))
        assert names == sorted(os.listdir(RES))
        for nm in names:
            assert open(os.path.join(d, nm)).read() == open(os.path.join(RES, nm)).read(), nm
    assert len(names) == 56


def _bottles():
    return sorted(f for f in os.listdir(RES) if "Bottle" in f)


@pytest.mark.parametrize("bottle", _bottles())
def test_scene_identical_to_oracle_for_every_bottle(ortlib, orc, bottle):
    for lam in (None, 843e-9):
        st = ortlib.make_settings(bottle, *cases.C2[1:])
        sc, pre = ortlib.build_scene(st, RES, lam)
        so = orc.make_scene(bottle, *cases.C2[1:], lens_wavelength=lam)
        assert bytes(sc) == bytes(so)


def test_scene_identical_for_every_lens_pair(ortlib, orc):
    planos = sorted(f for f in os.listdir(RES) if f.startswith("planoConvex"))
    doublets = sorted(f for f in os.listdir(RES) if f.startswith("achromaticDoublet"))
    for p in planos:
        for d in doublets:
            st = ortlib.make_settings("clearBottle-small.params", p, d)
            sc, _ = ortlib.build_scene(st, RES, 843e-9)
            assert bytes(sc) == bytes(orc.make_scene("clearBottle-small.params", p, d, lens_wavelength=843e-9))


def test_bottle_optional_mu_lines(ortlib, tmp_path):
    """12 lines -> mu = 0; 16 lines -> all four; the shipped 14-line file is
    tolerated instead of aborting like the reference."""
    import ctypes as C
    b = abi.Bottle()
    L = ortlib.load()
    assert L.ort_load_bottle(os.path.join(RES, "clearBottle-small.params").encode(), 785e-9, C.byref(b)) == 0
    assert (b.mua_b, b.mus_b, b.mua_c, b.mus_c, b.scatter_b, b.scatter_c, b.ellipse) == (0, 0, 0, 0, 0, 0, 0)
    assert L.ort_load_bottle(os.path.join(RES, "scatterBottle-large.params").encode(), 785e-9, C.byref(b)) == 0
    assert (b.mua_b, b.mus_b, b.mua_c, b.mus_c, b.scatter_b, b.scatter_c) == (0.5, 20.0, 1.0, 30.0, 1, 1)
    assert L.ort_load_bottle(os.path.join(RES, "clearBottle-small_0.0mm.params").encode(), 785e-9, C.byref(b)) == 0
    assert (b.mua_b, b.mus_b, b.mua_c, b.mus_c, b.scatter_b) == (0, 0, 0, 0, 0) and b.thickness == 2e-3
    assert L.ort_load_bottle(os.path.join(RES, "clearBottle-ellipse-short.params").encode(), 785e-9, C.byref(b)) == 0
    assert b.ellipse == 1 and b.radiusa == 17.5e-3 and b.radiusb == 35e-3
    short = tmp_path / "short.params"
    short.write_text("1.0\n2.0\n")
    assert L.ort_load_bottle(str(short).encode(), 785e-9, C.byref(b)) == abi.ORT_EPARSE
    assert L.ort_load_bottle(b"/nonexistent/x.params", 785e-9, C.byref(b)) == abi.ORT_EIO


def test_list_directed_tokens(ortlib, tmp_path):
    """d exponents, integer-as-real, logical spellings, quoted names, trailing junk, no final
    newline, blank lines."""
    text = "\n".join([
        "0.5D-3   ! ring", "785d-9 #", "", "2500000000  rays (needs int64)", "5", "1.45 axicon",
        ".TRUE.", "F", "t", "1.d-2,extra", "0.0", "'point'", "before", "0.5", "clearBottle-large.params junk",
        "planoConvex-f39.9mm.params", "\"achromaticDoublet-f50.0mm.params\"", "bessel-normal.dat",
        "outdir", "1.5d-3", "1.d-3 no newline at the end"])
    f = tmp_path / "s.params"
    f.write_text(text)
    s = ortlib.read_settings(str(f))
    assert s.ring_width == 0.5e-3 and s.wavelength == 785e-9 and s.nphotons == 2_500_000_000
    assert (s.use_bottle, s.use_tracker, s.make_images) == (1, 0, 1)
    assert s.image_diameter == 1e-2 and s.source_type == b"point"
    assert (s.iris_before, s.iris_after, s.iris_radius) == (1, 0, 0.5)
    assert s.l3_file == b"achromaticDoublet-f50.0mm.params" and s.folder == b"outdir"
    assert s.isors_offset == 1.5e-3 and s.spot_size == 1e-3


def test_settings_errors(ortlib, tmp_path):
    from opticalraytrace_b200.lib import OrtError
    base = open(os.path.join(RES, "settings.params")).read().splitlines()
    bad = list(base)
    bad[10] = "laser   # unknown source"
    f = tmp_path / "a.params"
    f.write_text("\n".join(bad))
    with pytest.raises(OrtError, match="No such source type"):
        ortlib.read_settings(str(f))
    bad = list(base)
    bad[11] = "middle"
    f.write_text("\n".join(bad))
    with pytest.raises(OrtError, match="No such iris position"):
        ortlib.read_settings(str(f))
    f.write_text("\n".join(base[:15]))
    with pytest.raises(OrtError, match="end of file"):
        ortlib.read_settings(str(f))
    s = ortlib.read_settings(os.path.join(RES, "settings.params"))
    assert s.nphotons == 100 and s.source_type == b"crs" and s.bottle_file == b"clearBottle-small_0.0mm.params"


def test_output_basename_recipe(ortlib):
    """src/main.f90:45-48 with str() = first len chars of f100.16 (src/utils.f90:351-369)."""
    st = ortlib.make_settings(*cases.C2)
    sc, pre = ortlib.build_scene(st, RES)
    assert ortlib.output_basename(st, sc, pre) == (
        "point_bottle_T_Ra_0.03500_Rb_0.03500_offset_-0.0020__F_F_1.00000_L2f_0.0399_L3f_0.0500"
        "_fo_0.00000_alp_5.00000_bwidth_0.00050_sep_0.00150")
    # the name keeps the pre-guard offset (quirk 5) while the scene carries the guarded one
    st = ortlib.make_settings("clearBottle-large_14mm.params", *cases.C2[1:], iris="after", iris_radius=0.25,
                              use_bottle=False, fibre_offset=-1e-3)
    sc, pre = ortlib.build_scene(st, RES)
    assert pre == 14e-3 and sc.bottle.centre[2] == pytest.approx(-1.3e-3, abs=1e-15)
    name = ortlib.output_basename(st, sc, pre)
    assert "_bottle_F_" in name and "_offset_0.01400__F_T_0.25000_" in name and "_fo_-0.0010_" in name


def test_image_files_and_trans_stats(ortlib, tmp_path):
    ring = np.zeros((401, 401), dtype=np.uint64)
    point = np.zeros((401, 401), dtype=np.uint64)
    ring[200 + 0, 200 + 60] = 7          # bin (xp=60, yp=0)
    point[200 - 3, 200 + 1] = 2 ** 40    # beyond int32: the reference would have overflowed
    base = str(tmp_path / "name_image")
    ortlib.write_images(base, ring, point)
    for suffix, want in (("-ring.dat", ring), ("-point.dat", point), ("-total.dat", ring + point)):
        raw = np.fromfile(base + suffix, dtype=np.float64)
        assert raw.size == 401 * 401 and os.path.getsize(base + suffix) == 1286408
        assert np.array_equal(raw.reshape(401, 401), want.astype(np.float64))
    assert np.fromfile(base + "-ring.dat", dtype=np.float64)[(0 + 200) * 401 + (60 + 200)] == 7.0
    st = ortlib.make_settings(*cases.C2, nphotons=1000)
    sc, _ = ortlib.build_scene(st, RES, 843e-9)
    ortlib.append_trans_stats(str(tmp_path), st, sc, 999, 507)
    ortlib.append_trans_stats(str(tmp_path), st, sc, 1000, 0)
    lines = open(tmp_path / "trans-stats.dat").read().splitlines()
    assert len(lines) == 3 and lines[0].strip().startswith("r/%, p/%, l2%f")
    cols = [c.strip() for c in lines[1].split(",")]
    assert float(cols[0]) == pytest.approx(0.1) and float(cols[1]) == pytest.approx(49.3)
    assert float(cols[2]) == 39.9e-3 and cols[4] == "T" and cols[7] == "F F" and cols[10] == "point"
