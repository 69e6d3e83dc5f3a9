"""ctypes binding of opticalraytrace_b200/libort.so (the C-ABI of include/ort.h).

Loading fails loudly when the CUDA library has not been built: there is no Python or CPU
fallback for the trace loop.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORT_LIB lets a tuning run pick an experimental build of the same library; default = the product
LIB_PATH = os.environ.get("ORT_LIB") or os.path.join(_HERE, "libort.so")
_lib = None

# every symbol include/ort.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "ort_init", "ort_init_rank", "ort_nccl_unique_id", "ort_finalize", "ort_synchronize", "ort_last_error",
    "ort_device_count", "ort_struct_sizes", "ort_trace", "ort_trace_rays", "ort_uniforms", "ort_measure_fp64_peak", "ort_math_selftest", "ort_mufu_selftest", "ort_write_tracks",
    "ort_load_image_source", "ort_set_image_source",
    "ort_load_plano", "ort_load_doublet", "ort_load_bottle", "ort_read_settings",
    "ort_build_scene", "ort_job_from_settings", "ort_output_basename", "ort_write_images",
    "ort_append_trans_stats",
    "ort_bpm_defaults", "ort_bpm_struct_size", "ort_bpm_bessel", "ort_bpm_write_file",
    "ort_trace_volume", "ort_write_volume",
]


class OrtError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("ort error %d: %s" % (code, text))
        self.code = code


def load():
    """Load libort.so.  Raises if it is missing -- build it with `python -c 'import
    __graft_entry__ as g; g.build()'` or `make -C opticalraytrace_b200/csrc`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrtError(abi.ORT_ENODEVICE,
                       "%s not built; the trace loop has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ort_last_error.restype = C.c_char_p
    L.ort_init.argtypes = [C.c_int]
    L.ort_init_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ort_nccl_unique_id.argtypes = [C.c_void_p]
    L.ort_trace.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.POINTER(abi.Timing)]
    L.ort_trace_rays.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
    L.ort_uniforms.argtypes = [C.c_uint64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    L.ort_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ort_math_selftest.argtypes = [C.c_int64, C.POINTER(C.c_uint64)]
    L.ort_mufu_selftest.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ort_write_tracks.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_char_p]
    L.ort_load_image_source.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, C.c_void_p]
    L.ort_set_image_source.argtypes = [C.c_void_p]
    L.ort_load_plano.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(abi.Plano)]
    L.ort_load_doublet.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(abi.Doublet)]
    L.ort_load_bottle.argtypes = [C.c_char_p, C.c_double, C.POINTER(abi.Bottle)]
    L.ort_read_settings.argtypes = [C.c_char_p, C.POINTER(abi.Settings)]
    L.ort_build_scene.argtypes = [C.POINTER(abi.Settings), C.c_char_p, C.c_double,
                                  C.POINTER(abi.Scene), C.POINTER(C.c_double)]
    L.ort_job_from_settings.argtypes = [C.POINTER(abi.Settings), C.c_int32, C.POINTER(abi.Job)]
    L.ort_output_basename.argtypes = [C.POINTER(abi.Settings), C.POINTER(abi.Scene), C.c_double,
                                      C.c_char_p, C.c_size_t]
    L.ort_write_images.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    L.ort_append_trans_stats.argtypes = [C.c_char_p, C.POINTER(abi.Settings),
                                         C.POINTER(abi.Scene), C.c_int64, C.c_int64]
    L.ort_trace_volume.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_void_p, C.c_void_p, C.c_void_p]
    L.ort_write_volume.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    L.ort_bpm_defaults.argtypes = [C.POINTER(abi.Bpm)]
    L.ort_bpm_bessel.argtypes = [C.POINTER(abi.Bpm), C.c_void_p]
    L.ort_bpm_write_file.argtypes = [C.POINTER(abi.Bpm), C.c_char_p]
    _lib = L
    return L


def check(rc, allow=()):
    if rc < 0 and rc not in allow:
        raise OrtError(rc, load().ort_last_error().decode(errors="replace"))
    return rc


# ---- lifetime ---------------------------------------------------------------------------
def init(ngpus=1):
    return check(load().ort_init(ngpus))


def init_rank(device, rank, nranks, nccl_id=None):
    buf = None
    if nccl_id is not None:
        buf = C.create_string_buffer(bytes(nccl_id), 128)
    return check(load().ort_init_rank(device, rank, nranks, buf))


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    check(load().ort_nccl_unique_id(buf))
    return buf.raw


def finalize():
    return check(load().ort_finalize())


def synchronize():
    return check(load().ort_synchronize())


def device_count():
    return load().ort_device_count()


def struct_sizes():
    out = (C.c_int32 * 8)()
    check(load().ort_struct_sizes(out))
    return list(out)


# ---- host side of the drop-in surface -------------------------------------------------------
def read_settings(path):
    s = abi.Settings()
    check(load().ort_read_settings(os.fsencode(path), C.byref(s)))
    return s


def build_scene(settings, resdir, lens_wavelength=None):
    """-> (Scene, pre_guard_offset); reference src/setupMod.f90:113-119 + src/main.f90:51-70,81"""
    sc = abi.Scene()
    pre = C.c_double(0.0)
    lw = settings.wavelength if lens_wavelength is None else lens_wavelength
    check(load().ort_build_scene(C.byref(settings), os.fsencode(resdir), lw, C.byref(sc),
                                 C.byref(pre)))
    return sc, pre.value


def make_settings(bottle="clearBottle-large.params", l2="planoConvex-f39.9mm.params",
                  l3="achromaticDoublet-f50.0mm.params", *, nphotons=1000, wavelength=785e-9,
                  alpha_deg=5.0, n_axicon=1.45, ring_width=0.5e-3, use_bottle=True,
                  image_diameter=1e-2, fibre_offset=0.0, source_type="point", iris="none",
                  iris_radius=1.0, folder="run", isors_offset=1.5e-3, spot_size=1e-3):
    """An in-memory settings.params (same fields as reference src/setupMod.f90:57-133)."""
    s = abi.Settings()
    s.ring_width, s.wavelength, s.alpha_deg, s.n_axicon = ring_width, wavelength, alpha_deg, n_axicon
    s.image_diameter, s.fibre_offset, s.iris_radius = image_diameter, fibre_offset, iris_radius
    s.isors_offset, s.spot_size, s.nphotons = isors_offset, spot_size, nphotons
    s.use_bottle, s.use_tracker, s.make_images = int(use_bottle), 0, 1
    s.iris_before, s.iris_after = int(iris == "before"), int(iris == "after")
    s.source_type, s.iris_name = source_type.encode(), iris.encode()
    s.bottle_file, s.l2_file, s.l3_file = bottle.encode(), l2.encode(), l3.encode()
    s.image_file, s.folder = b"bessel-normal.dat", folder.encode()
    return s


def job_from_settings(settings, phase):
    j = abi.Job()
    check(load().ort_job_from_settings(C.byref(settings), phase, C.byref(j)))
    return j


def output_basename(settings, scene, pre_guard_offset):
    buf = C.create_string_buffer(1024)
    check(load().ort_output_basename(C.byref(settings), C.byref(scene), pre_guard_offset, buf, 1024))
    return buf.value.decode()


def write_images(base, ring, point):
    ring = np.ascontiguousarray(ring, dtype=np.uint64)
    point = np.ascontiguousarray(point, dtype=np.uint64)
    check(load().ort_write_images(os.fsencode(base), ring.ctypes.data, point.ctypes.data))


def append_trans_stats(folder, settings, scene, rcount, pcount):
    check(load().ort_append_trans_stats(os.fsencode(folder), C.byref(settings), C.byref(scene),
                                        int(rcount), int(pcount)))


# ---- the hot path -----------------------------------------------------------------------------
def trace(job, scenes, *, want_image=True, allow_trap=False):
    """ort_trace -> (image[nscenes,401,401] uint64 | None, lost[nscenes], hist[nscenes,32], Timing)"""
    if isinstance(scenes, abi.Scene):
        scenes = [scenes]
    ns = len(scenes)
    arr = (abi.Scene * ns)(*scenes)
    image = np.empty((ns, abi.ORT_IMG_N, abi.ORT_IMG_N), dtype=np.uint64) if want_image else None   # overwritten
    lost = np.zeros(ns, dtype=np.int64)
    hist = np.zeros((ns, abi.ORT_NSTATUS), dtype=np.int64)
    tm = abi.Timing()
    rc = load().ort_trace(C.byref(job), arr, ns, image.ctypes.data if want_image else None,
                          lost.ctypes.data, hist.ctypes.data, C.byref(tm))
    check(rc, allow=(abi.ORT_ETRACE,) if allow_trap else ())
    return image, lost, hist, tm


def trace_rays(job, scene, n, pos_in=None, dir_in=None):
    """ort_trace_rays -> dict(pos[3,n], dir[3,n], status[n], bin[2,n])"""
    pos_out = np.zeros((3, n))
    dir_out = np.zeros((3, n))
    status = np.zeros(n, dtype=np.int32)
    bins = np.zeros((2, n), dtype=np.int32)
    pi = di = None
    if pos_in is not None:
        pin = np.ascontiguousarray(pos_in, dtype=np.float64)
        din = np.ascontiguousarray(dir_in, dtype=np.float64)
        assert pin.shape == (3, n) and din.shape == (3, n)
        pi, di = pin.ctypes.data, din.ctypes.data
    check(load().ort_trace_rays(C.byref(job), C.byref(scene), n, pi, di, pos_out.ctypes.data,
                                dir_out.ctypes.data, status.ctypes.data, bins.ctypes.data),
          allow=(abi.ORT_ETRACE,))
    return dict(pos=pos_out, dir=dir_out, status=status, bin=bins)


def load_image_source(path, nphotons, seed=123456789):
    """ort_load_image_source (init_emit_image) -> int32[512*512] ray budget, Fortran memory order"""
    budget = np.zeros(abi.SRCIMG_N * abi.SRCIMG_N, dtype=np.int32)
    check(load().ort_load_image_source(os.fsencode(path), int(nphotons), seed, budget.ctypes.data))
    return budget


def set_image_source(budget):
    if budget is None:
        return check(load().ort_set_image_source(None))
    b = np.ascontiguousarray(budget, dtype=np.int32)
    assert b.size == abi.SRCIMG_N * abi.SRCIMG_N
    return check(load().ort_set_image_source(b.ctypes.data))


def trace_volume(job, scene, allow_trap=False):
    """ort_trace_volume (makeImage3D) -> (volume[200, 401, 401] uint32, lost, hist[32])"""
    vol = np.zeros((abi.VOL_DEPTH, abi.ORT_IMG_N, abi.ORT_IMG_N), dtype=np.uint32)
    lost = np.zeros(1, dtype=np.int64)
    hist = np.zeros(abi.ORT_NSTATUS, dtype=np.int64)
    check(load().ort_trace_volume(C.byref(job), C.byref(scene), vol.ctypes.data, lost.ctypes.data,
                                  hist.ctypes.data), allow=(abi.ORT_ETRACE,) if allow_trap else ())
    return vol, int(lost[0]), hist


def write_volume(base, vol_ring=None, vol_point=None):
    """ort_write_volume (writeImage3D): <base>-vol-ring.dat / <base>-vol-point.dat, raw fp64"""
    vols = [None if v is None else np.ascontiguousarray(v, dtype=np.uint32) for v in (vol_ring, vol_point)]
    check(load().ort_write_volume(os.fsencode(base), *(None if v is None else v.ctypes.data for v in vols)))


def bpm_defaults():
    p = abi.Bpm()
    check(load().ort_bpm_defaults(C.byref(p)))
    return p


def bpm_bessel(params=None):
    """ort_bpm_bessel -> float64[nxy, nxy] in the element order of the reference's bessel-normal.dat"""
    p = params if params is not None else bpm_defaults()
    out = np.zeros((p.nxy, p.nxy), dtype=np.float64)
    check(load().ort_bpm_bessel(C.byref(p), out.ctypes.data))
    return out


def bpm_write_file(path, params=None):
    p = params if params is not None else bpm_defaults()
    check(load().ort_bpm_write_file(C.byref(p), os.fsencode(path)))


def write_tracks(job, scene, path):
    """ort_write_tracks: the reference's *-ringtrace.dat / *-pointtrace.dat"""
    check(load().ort_write_tracks(C.byref(job), C.byref(scene), os.fsencode(path)))


def uniforms(seed, phase, ray, first_slot, n):
    out = np.zeros(n, dtype=np.float64)
    check(load().ort_uniforms(seed, phase, ray, first_slot, n, out.ctypes.data))
    return out


def math_selftest(n=1 << 24):
    out = (C.c_uint64 * 4)()
    check(load().ort_math_selftest(n, out))
    return dict(zip(("rcp", "div", "sqrt", "rsqrt"), (int(v) for v in out)))


def mufu_selftest():
    """ort_mufu_selftest: (measured, assumed) largest errors of rcp / rsqrt / sqrt (relative) and
    sin / cos (absolute) .approx.ftz.f32 over every fp32 argument"""
    w, a = (C.c_double * 5)(), (C.c_double * 5)()
    check(load().ort_mufu_selftest(w, a))
    names = ("rcp", "rsqrt", "sqrt", "sin", "cos")
    return dict(zip(names, w)), dict(zip(names, a))


def measure_fp64_peak():
    tf, mhz = C.c_double(0), C.c_double(0)
    check(load().ort_measure_fp64_peak(C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value
