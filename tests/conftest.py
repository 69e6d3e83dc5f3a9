import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RES = os.path.join(ROOT, "res")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from tests import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def harness():
    """Product optics header compiled for the host (test-only)."""
    from opticalraytrace_b200 import abi
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests"), "-s"])
    H = C.CDLL(os.path.join(ROOT, "tests", "libhost_harness.so"))
    H.hh_trace_rays.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64] + [C.c_void_p] * 6

    H.hh_set_image_source.argtypes = [C.c_void_p]

    def set_image_source(budget):
        if budget is None:
            H.hh_set_image_source(None)
        else:
            b = np.ascontiguousarray(budget, dtype=np.int32)
            H.hh_set_image_source(b.ctypes.data)

    def run(job, scene, n, pos_in=None, dir_in=None):
        po, do = np.zeros((3, n)), np.zeros((3, n))
        st, b = np.zeros(n, np.int32), np.zeros((2, n), np.int32)
        pi = di = None
        if pos_in is not None:
            pin = np.ascontiguousarray(pos_in, dtype=np.float64)
            din = np.ascontiguousarray(dir_in, dtype=np.float64)
            pi, di = pin.ctypes.data, din.ctypes.data
        H.hh_trace_rays(C.byref(job), C.byref(scene), n, pi, di, po.ctypes.data, do.ctypes.data,
                        st.ctypes.data, b.ctypes.data)
        return dict(pos=po, dir=do, status=st, bin=b)
    H.hh_ring_filter.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64, C.c_void_p]

    def ring_filter(job, scene, n):
        v = np.zeros(n, np.int32)
        shortcut = H.hh_ring_filter(C.byref(job), C.byref(scene), n, v.ctypes.data)
        return v, bool(shortcut)
    H.hh_ring_filter_pairs.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]

    def ring_filter_pairs(job, scene, n, shift):
        lo = np.zeros(n, dtype=np.int32)
        hi = np.zeros(n, dtype=np.int32)
        H.hh_ring_filter_pairs(C.byref(job), C.byref(scene), n, shift, lo.ctypes.data, hi.ctypes.data)
        return lo, hi
    run.ring_filter_pairs = ring_filter_pairs
    H.hh_ring_aim_cut.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.POINTER(C.c_int)]
    H.hh_ring_aim_cut.restype = C.c_uint64

    def ring_aim_cut(job, scene):
        have = C.c_int(0)
        cut = H.hh_ring_aim_cut(C.byref(job), C.byref(scene), C.byref(have))
        return int(cut), bool(have.value)
    run.ring_aim_cut = ring_aim_cut
    HF = C.CDLL(os.path.join(ROOT, "tests", "libhost_harness_fuzz.so"))
    for lib in (H, HF):
        lib.hh_filter_bounds.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_void_p]

    def filter_bounds(job, scene, n, fuzz=False):
        """-> (usable, {kind: largest |fp32 - exact| / bound}, counts) -- see hh_filter_bounds;
        usable is None when the scene lacks the filter's premise (no ring_shortcut)"""
        mr, cnt = np.zeros(16), np.zeros(8, np.int64)
        ok = (HF if fuzz else H).hh_filter_bounds(C.byref(job), C.byref(scene), n, mr.ctypes.data, cnt.ctypes.data, None)
        kinds = ["", "pos", "dir", "normal", "n.i", "s2", "ct2", "cos_t", "F", "h", "c", "disc", "t", "rho2"]
        return (bool(ok & 2) if ok & 1 else None), {kinds[k]: mr[k] for k in range(1, 14)}, dict(
            records=int(cnt[0]), violations=int(cnt[1]), called=int(cnt[2]), wrong=int(cnt[3]), passed=int(cnt[4]))
    run.filter_bounds = filter_bounds
    H.hh_ring_filter_in_range.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene)]
    run.ring_filter_in_range = lambda job, scene: bool(H.hh_ring_filter_in_range(C.byref(job), C.byref(scene)))
    run.set_image_source = set_image_source
    run.ring_filter = ring_filter
    return run


@pytest.fixture(scope="session")
def ortlib():
    """libort.so built (no device needed)."""
    from opticalraytrace_b200 import lib
    so = lib.LIB_PATH
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "opticalraytrace_b200", "csrc"), "-s"])
    lib.load()
    return lib


@pytest.fixture(scope="session")
def ort(ortlib):
    """Initialised on cuda:0 -- fails loudly when there is no device (no CPU fallback)."""
    ortlib.init(1)
    yield ortlib
    ortlib.finalize()


def rel_err(a, b):
    """per-ray max-abs difference relative to the vector's largest component"""
    d = np.abs(a - b).max(axis=0)
    s = np.maximum(np.abs(a).max(axis=0), 1e-300)
    return d / s


def comp_err(a, b, length=0.0):
    """per-ray, per-COMPONENT error: max_i |a_i - b_i| / (|a_i| + 1e-2 |a|_max).  comp_err < 1e-9 is
    |a_i - b_i| < 1e-9 |a_i| + 1e-11 |a|_max: every component agrees to 1e-9 of ITSELF, down to
    components a hundredth of the vector; below that the floor is 1e-11 of the vector -- a hundred
    times tighter than rel_err's -- because fp64 cannot hold a small component to 1e-9 of itself:
    after seven surfaces the rounding noise of a unit direction is ~2e-12 absolute (the largest seen
    in 1e7 rays), which IS 1e-6 of a component of size 2e-6.  `length`: for positions, |a|_max is not
    allowed to fall below the length scale of the system the point was computed in (a point 0.9 mm
    from the origin reached by going 36 mm forth and 35 mm back carries the rounding of 36 mm)."""
    scale = np.maximum(np.maximum(np.abs(a).max(axis=0), length), 1e-300)
    return (np.abs(a - b) / (np.abs(a) + 1e-2 * scale)).max(axis=0)


def both_err(a, b, length=0.05):
    """max over position and direction of (vector-relative error, per-component error); `length`: the
    scale of the optical train (L2 sits ~36-100 mm from the origin in every shipped set-up)"""
    return (np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"])),
            np.maximum(comp_err(a["pos"], b["pos"], length), comp_err(a["dir"], b["dir"])))


def scatter_conditioning(orc, job, scene, n, seeds=(11, 22, 33, 44, 55, 66)):
    """How far apart may two correct implementations of the scatter path be?  The oracle against itself
    with every libm result of tauint / stokes and every intersection distance moved by a pseudo-random
    -2..+2 ulp (orc_set_jitter): -> (response[n] = largest vector-relative change of the final position /
    direction over the jitter seeds, stable[n] = the final status never changed)."""
    base = orc.trace_rays(job, scene, n)
    resp = np.zeros(n)
    stable = np.ones(n, bool)
    try:
        for s in seeds:
            orc.set_jitter(s)
            j = orc.trace_rays(job, scene, n)
            stable &= j["status"] == base["status"]
            with np.errstate(invalid="ignore"):
                e = np.maximum(rel_err(base["pos"], j["pos"]), rel_err(base["dir"], j["dir"]))
            resp = np.maximum(resp, np.nan_to_num(e, nan=np.inf))
    finally:
        orc.set_jitter(0)
    return base, resp, stable
