"""Pins the CPU oracle.  The reference ships no tests or golden vectors and cannot be built here
(no Fortran compiler), so the oracle is pinned by: published Philox vectors, the closed-form /
physics known answers of SURVEY.md 8(c), invariants, and an independent second restatement
(tests/pyref.py)."""
import ctypes as C
import math

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases, pyref


RANDOM123_KAT = [  # Random123 kat_vectors, philox4x32-10: (counter, key, output)
    ([0] * 4, [0] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def _philox(L, c, k, first, n):
    c = (C.c_uint32 * 4)(*c)
    k = (C.c_uint32 * 2)(*k)
    o = (C.c_uint32 * 4)()
    L.orc_philox(c, k, first, n, o)
    return list(o)


def test_philox_random123_vectors(orc):
    """The round function and key schedule against the published philox4x32-10 vectors, and the
    7-round generator used here pinned to them: rounds 7..9 applied to its output must give the
    published 10-round output (a Philox round is a bijection, so this determines the 7-round
    output uniquely)."""
    L = orc.lib()
    assert L.orc_philox_rounds() == 7
    for c, k, out in RANDOM123_KAT:
        assert _philox(L, c, k, 0, 10) == out
        seven = _philox(L, c, k, 0, 7)
        assert seven != out
        assert _philox(L, seven, k, 7, 3) == out


def test_slot_map_rev2(orc):
    """the slot -> (block, words) table of the generator, from raw blocks"""
    L = orc.lib()
    seed, phase, ray = 0x123456789abcdef0, 2, 0x1_0000_0007
    key = [seed & 0xffffffff, seed >> 32]
    blk = [_philox(L, [ray & 0xffffffff, ray >> 32, phase, b], key, 0, 7) for b in range(12)]
    wide = lambda w, h=0: float((((w[2 * h + 1] << 32) | w[2 * h]) >> 11)) * 2.0 ** -53
    narrow = lambda x: x * 2.0 ** -32
    want = [wide(blk[0]), narrow(blk[0][2]), wide(blk[1]), narrow(blk[1][2]), narrow(blk[0][3]), narrow(blk[1][3]),
            narrow(blk[2][0]), narrow(blk[2][1]), narrow(blk[2][2]), narrow(blk[2][3]),
            wide(blk[3]), narrow(blk[3][2]), narrow(blk[3][3]), wide(blk[4]), narrow(blk[4][2]), narrow(blk[4][3]),
            wide(blk[8]), wide(blk[8], 1), wide(blk[9]), wide(blk[9], 1), wide(blk[10]), wide(blk[10], 1)]
    got = orc.uniforms(seed, phase, ray, 0, len(want))
    assert list(got) == want
    # ring loop (phase 1): the high word of slot 2 is word (ray & 3) of the block rays 4q .. 4q+3 share,
    # counter (ray >> 2, 16 + phase, 0); everything else as in the table
    blk1 = [_philox(L, [ray & 0xffffffff, ray >> 32, 1, b], key, 0, 7) for b in range(5)]
    q = ray >> 2
    shared = _philox(L, [q & 0xffffffff, q >> 32, 17, 0], key, 0, 7)
    got1 = orc.uniforms(seed, 1, ray, 0, 6)
    assert got1[2] == float((((shared[ray & 3] << 32) | blk1[1][0]) >> 11)) * 2.0 ** -53
    assert list(got1[[0, 1, 3, 4, 5]]) == [wide(blk1[0]), narrow(blk1[0][2]), narrow(blk1[1][2]), narrow(blk1[0][3]), narrow(blk1[1][3])]
    # the four rays of a quad take the four different words
    his = [int(orc.uniforms(seed, 1, 4 * q + k, 2, 1)[0] * 2.0 ** 32) for k in range(4)]
    assert his == shared


def test_uniform_stream_properties(orc):
    u = np.stack([orc.uniforms(123456789, 1, r, 0, 20) for r in range(20000)])
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    # every slot is uniform on its own (chi-square over 64 cells, 20000 samples: mean 63, sd 11.2) ...
    for k in range(20):
        h = np.bincount((u[:, k] * 64).astype(int), minlength=64)
        assert ((h - 312.5) ** 2 / 312.5).sum() < 63 + 5 * 11.2, k
    # ... uncorrelated with every other slot of the same ray and with the same slot of the next ray
    cc = np.corrcoef(u.T)
    assert np.abs(cc - np.eye(20)).max() < 0.04
    for k in range(20):
        assert abs(np.corrcoef(u[:-1, k], u[1:, k])[0, 1]) < 0.04
    # pure function of (seed, phase, ray, slot): slots can be fetched in any grouping
    a = orc.uniforms(5, 2, 77, 0, 24)
    b = np.concatenate([orc.uniforms(5, 2, 77, 0, 5), orc.uniforms(5, 2, 77, 5, 19)])
    assert np.array_equal(a, b)
    assert not np.array_equal(orc.uniforms(5, 1, 77, 0, 4), orc.uniforms(5, 2, 77, 0, 4))
    # wide draws have 53-bit resolution, narrow ones 32-bit
    assert np.all(np.mod(a * 2.0 ** 53, 1.0) == 0.0)
    narrow = [1, 3, 4, 5, 6, 7, 8, 9, 11, 12, 14, 15]
    assert np.all(np.mod(a[narrow] * 2.0 ** 32, 1.0) == 0.0)
    assert np.any(np.mod(a[[0, 2, 10, 13, 16, 17]] * 2.0 ** 32, 1.0) != 0.0)


def test_dispersion_constants(orc):
    """SURVEY 8(c)(1)"""
    s785 = orc.make_scene(*cases.C2)
    s843 = orc.make_scene(*cases.C2, lens_wavelength=843e-9)
    assert s785.L2.n2 == pytest.approx(1.511079564908228, abs=2e-15)
    assert s843.L2.n2 == pytest.approx(1.509964985996694, abs=2e-15)
    assert s785.L3.n2 == pytest.approx(1.643111336013526, abs=2e-15)
    assert s785.L3.n3 == pytest.approx(1.785335731036205, abs=2e-15)
    assert s843.L3.n2 == pytest.approx(1.641622124360634, abs=2e-15)
    assert s843.L3.n3 == pytest.approx(1.782026735556232, abs=2e-15)
    assert s785.bottle.nbottle == pytest.approx(1.517476652730365, abs=2e-15)
    assert s785.bottle.ncontents == pytest.approx(1.357668387241211, abs=2e-15)
    # the bottle keeps the settings wavelength in the point phase (src/main.f90:113-117)
    assert s843.bottle.nbottle == s785.bottle.nbottle


def test_prologue_constants(orc):
    s = orc.make_scene(*cases.C2)
    assert s.cos_theta_max == pytest.approx(0.942159141664439, abs=1e-15)
    assert math.asin(0.22) == pytest.approx(0.221814470496794, abs=1e-15)
    assert s.L2.centre[2] == pytest.approx(0.0215, abs=1e-15)
    assert s.L3.centre1[2] == pytest.approx(0.15635, abs=1e-15)
    assert s.L3.centre2[2] == pytest.approx(0.10325, abs=1e-15)
    assert s.L3.centre3[2] == pytest.approx(0.0065, abs=1e-15)
    assert s.img_plane == pytest.approx(0.1771, abs=1e-15)
    bessel = 2 * math.sqrt(s.r2)
    assert bessel == pytest.approx(3.5338e-3, rel=1e-4)
    assert s.r1 == pytest.approx(9.20393e-6, rel=1e-5) and s.r2 == pytest.approx(3.12193e-6, rel=1e-5)
    assert s.r1 > s.r2


def test_offset_guard_collapses_positive_offsets(orc):
    """with f39.9 (fb 35.7 mm) every +2..+14 mm file ends at z0 = -1.3 mm."""
    zs = []
    for mm in range(-14, 15, 2):
        s = orc.make_scene("clearBottle-large_%dmm.params" % mm, *cases.C2[1:])
        zs.append(s.bottle.centre[2])
    assert zs[:8] == pytest.approx([m * 1e-3 for m in range(-14, 1, 2)], abs=1e-18)
    assert all(z == pytest.approx(0.0357 - 0.035 - 2e-3, abs=1e-15) for z in zs[8:])


def test_kat_a_single_ray(orc):
    """clearBottle-large, f39.9/f50 at 843 nm, dir=(sin .1,0,cos .1), u=0.5."""
    s = orc.make_scene(*cases.C2, lens_wavelength=843e-9)
    p = np.zeros((3, 1))
    d = np.array([[math.sin(0.1)], [0.0], [math.cos(0.1)]])
    want = {
        abi.STOP_BOTTLE: ((3.2886656276517e-3, 0, 0.033), (0.13554067377167903, 0, 0.9907717828811636)),
        abi.STOP_L2: ((4.19593098785981e-3, 0, 0.04166814724125938), None),
        abi.STOP_L3: (None, (-0.10455640335111067, 0, 0.9945189583503574)),
    }
    for stop, (wp, wd) in want.items():
        r = orc.trace_rays(abi.default_job(2, uniform_override=0.5, stop_after=stop), s, 1, p, d)
        assert r["status"][0] == abi.ST_STOPPED
        if wp:
            assert r["pos"][:, 0] == pytest.approx(wp, abs=1e-14)
        if wd:
            assert r["dir"][:, 0] == pytest.approx(wd, abs=1e-14)
    r = orc.trace_rays(abi.default_job(2, uniform_override=0.5), s, 1, p, d)
    assert r["status"][0] == 0 and tuple(r["bin"][:, 0]) == (60, 0)
    assert r["pos"][0, 0] == pytest.approx(1.49695788349947e-3, abs=1e-14)


def test_kat_c_collimation(orc):
    """a ray from L2's back focal point leaves collimated -> lands on the axis."""
    s = orc.make_scene(*cases.C2)
    p = np.zeros((3, 1))
    d = np.array([[math.sin(0.05) * math.cos(1.0)], [math.sin(0.05) * math.sin(1.0)], [math.cos(0.05)]])
    j = abi.default_job(2, use_bottle=False, uniform_override=0.5)
    r = orc.trace_rays(j, s, 1, p, d)
    assert r["status"][0] == 0 and tuple(r["bin"][:, 0]) == (0, 0)
    assert r["pos"][:, 0] == pytest.approx((7.0433e-6, 1.09693e-5, 0.1771), abs=2e-10)
    j.stop_after = abi.STOP_L2
    r2 = orc.trace_rays(j, s, 1, p, d)
    assert abs(r2["dir"][2, 0] - 1) < 1e-5  # collimated after L2


def test_fresnel_known_values(orc):
    L = orc.lib()
    n = orc.v3((0, 0, -1))
    assert L.orc_fresnel(orc.v3((0, 0, 1)), n, 1.0, 1.5) == 0.0         # quirk 3: exactly normal
    th = 1e-4
    r = L.orc_fresnel(orc.v3((math.sin(th), 0, math.cos(th))), n, 1.0, 1.5)
    assert r == pytest.approx(((1 - 1.5) / (1 + 1.5)) ** 2, rel=1e-6)   # near-normal: 4 %
    tb = math.atan(1.5)                                                  # Brewster: Rp = 0
    r = L.orc_fresnel(orc.v3((math.sin(tb), 0, math.cos(tb))), n, 1.0, 1.5)
    ct, c2 = math.cos(tb), math.sqrt(1 - (math.sin(tb) / 1.5) ** 2)
    rs = ((ct - 1.5 * c2) / (ct + 1.5 * c2)) ** 2
    assert r == pytest.approx(0.5 * rs, rel=1e-12)
    assert L.orc_fresnel(orc.v3((math.sin(0.8), 0, math.cos(0.8))), n, 1.5, 1.0) == 1.0  # TIR


def test_snell_and_unit_directions(orc):
    L = orc.lib()
    rng = np.random.default_rng(3)
    for _ in range(200):
        th = rng.uniform(0, 1.4)
        ph = rng.uniform(0, 2 * math.pi)
        I = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
        for N in ((0, 0, -1.0), (0, 0, 1.0)):
            buf = orc.v3(I)
            L.orc_refract(buf, orc.v3(N), 1.0 / 1.5)
            T = np.array(list(buf))
            assert abs(np.linalg.norm(T) - 1) < 1e-14
            assert 1.0 * math.sin(th) == pytest.approx(1.5 * math.hypot(T[0], T[1]), abs=1e-14)
            buf = orc.v3(I)
            L.orc_reflect(buf, orc.v3(N))
            assert list(buf)[2] == pytest.approx(-I[2], abs=1e-15)


def test_intersections_closed_form(orc):
    L = orc.lib()
    t = C.c_double()
    o, d, c = orc.v3((0, 0, -5)), orc.v3((0, 0, 1)), orc.v3((0, 0, 0))
    assert L.orc_intersect_sphere(o, d, c, 2.0, C.byref(t)) == 1 and t.value == 3.0
    assert L.orc_intersect_sphere(orc.v3((0, 0, 0)), d, c, 2.0, C.byref(t)) == 1 and t.value == 2.0
    assert L.orc_intersect_sphere(orc.v3((0, 0, 5)), d, c, 2.0, C.byref(t)) == 0
    assert L.orc_intersect_sphere(orc.v3((3, 0, -5)), d, c, 2.0, C.byref(t)) == 0
    # cylinder along x: x motion is free
    dd = orc.v3((0.6, 0, 0.8))
    assert L.orc_intersect_cylinder(orc.v3((7, 0, 0)), dd, c, 2.0, C.byref(t)) == 1
    assert t.value == pytest.approx(2.5, abs=1e-15)
    # ellipse semia <-> z, semib <-> y
    assert L.orc_intersect_ellipse(c, orc.v3((0, 0, 1)), c, 3.0, 1.0, C.byref(t)) == 1 and t.value == pytest.approx(3.0)
    assert L.orc_intersect_ellipse(c, orc.v3((0, 1, 0)), c, 3.0, 1.0, C.byref(t)) == 1 and t.value == pytest.approx(1.0)


def test_stokes_direction_stays_unit_and_forward_peaked(orc):
    L = orc.lib()
    cosines = []
    for ray in range(4000):
        buf = orc.v3((0.0, 0.6, 0.8))
        L.orc_stokes(buf, 0.9, 42, ray)
        v = np.array(list(buf))
        assert abs(np.linalg.norm(v) - 1) < 1e-12
        cosines.append(float(v @ np.array([0.0, 0.6, 0.8])))
    assert np.mean(cosines) == pytest.approx(0.9, abs=0.02)   # <cos> = g for Henyey-Greenstein
    cosines = []
    for ray in range(4000):
        buf = orc.v3((0.0, 0.6, 0.8))
        L.orc_stokes(buf, 0.0, 42, ray)
        cosines.append(list(buf)[2])
    assert abs(np.mean(cosines)) < 0.03                        # isotropic


def test_transmission_fractions(orc):
    """SURVEY 8(c)(4) probe fractions (different generator -> statistical agreement only)."""
    n = 400_000
    s = orc.make_scene(*cases.C1)
    s8 = orc.make_scene(*cases.C1, lens_wavelength=843e-9)
    _, lost, hist = orc.trace(abi.default_job(1, n), s)
    assert hist[0, 0] / n == pytest.approx(0.0245, abs=0.002)
    _, lost, hist = orc.trace(abi.default_job(2, n), s8)
    assert 1 - lost[0] / n == pytest.approx(0.600, abs=0.005)
    assert hist[0, 0] / n == pytest.approx(0.476, abs=0.005)
    s = orc.make_scene(*cases.C2)
    s8 = orc.make_scene(*cases.C2, lens_wavelength=843e-9)
    _, lost, hist = orc.trace(abi.default_job(1, n), s)
    assert 1e-4 < hist[0, 0] / n < 4e-4
    assert hist[0, 9] / n == pytest.approx(1 - 0.313, abs=0.005)     # 69 % die at L2's aperture
    _, lost, hist = orc.trace(abi.default_job(2, n), s8)
    assert 1 - lost[0] / n == pytest.approx(0.492, abs=0.005)
    assert hist[0, 0] / n == pytest.approx(0.418, abs=0.005)
    # ellipse bottles: nothing gets through in the reference (quirk 2) ...
    se = orc.make_scene(*cases.ELL, lens_wavelength=843e-9)
    _, lost, hist = orc.trace(abi.default_job(2, 100_000), se)
    assert hist[0, 0] == 0 and hist[0, 5] > 0.9 * 100_000
    # ... and does with the opt-in fix
    _, lost, hist = orc.trace(abi.default_job(2, 100_000, flags=abi.FLAG_FIX_OUTER_ELLIPSE), se)
    assert hist[0, 0] > 10_000


def test_hot_bin_without_bottle(orc):
    """use_bottle = false puts the point source at L2's focus: the image peaks sharply on the
    axis bins (the hot-bin case for the detector atomics)."""
    s8 = orc.make_scene(*cases.C2, lens_wavelength=843e-9)
    img, _, hist = orc.trace(abi.default_job(2, 100_000, use_bottle=False), s8)
    assert img.sum() == hist[0, 0]
    y, x = np.unravel_index(np.argmax(img[0]), img[0].shape)
    assert abs(int(y) - 200) <= 1 and abs(int(x) - 200) <= 1
    assert img[0].max() > 0.015 * img.sum()       # ~20x the hottest bin of the bottle runs


def test_image_layout_x_fastest(orc):
    """bin (xp,yp) lands at element (yp+200)*401 + (xp+200)  (src/imageMod.f90:102-112)."""
    s8 = orc.make_scene(*cases.C2, lens_wavelength=843e-9)
    p = np.zeros((3, 1))
    d = np.array([[math.sin(0.1)], [0.0], [math.cos(0.1)]])
    r = orc.trace_rays(abi.default_job(2, uniform_override=0.5), s8, 1, p, d)
    assert tuple(r["bin"][:, 0]) == (60, 0)
    img, _, _ = orc.trace(abi.default_job(2, 50_000), s8)
    rr = orc.trace_rays(abi.default_job(2), s8, 50_000)
    ok = rr["status"] == 0
    ref = np.zeros((401, 401), dtype=np.uint64)
    np.add.at(ref, (rr["bin"][1, ok] + 200, rr["bin"][0, ok] + 200), 1)
    assert np.array_equal(ref, img[0])


@pytest.mark.parametrize("cid,files,phase,kw", cases.RAY_CASES, ids=[c[0] for c in cases.RAY_CASES])
def test_pyref_agrees_with_oracle(orc, cid, files, phase, kw):
    """Independent pure-Python restatement vs the C++ oracle, ray by ray."""
    n = 1500
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, **kw)
    a = orc.trace_rays(job, scene, n)
    for i in range(n):
        u = orc.uniforms(job.seed, phase, i, 0, 10)
        pos, d, st, xp, yp = pyref.trace_one(job, scene, lambda k: u[k])
        assert st == a["status"][i], (i, st, a["status"][i])
        if st == 0:
            assert (xp, yp) == tuple(a["bin"][:, i])
        if st in (0, 21, 23):
            assert np.allclose(pos, a["pos"][:, i], rtol=0, atol=1e-15 + 1e-12 * np.abs(pos).max())
            assert np.allclose(d, a["dir"][:, i], rtol=0, atol=1e-13)
