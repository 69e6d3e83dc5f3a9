"""ctypes binding of oracle/libort_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from opticalraytrace_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
RES = os.path.join(ROOT, "res")
_lib = None
_fast = None

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int32)


def build():
    so = os.path.join(ORACLE_DIR, "libort_oracle.so")
    src = os.path.join(ORACLE_DIR, "ort_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "ort.h")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                               os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def fast_lib():
    """The timing twin (oracle/Makefile: the reference's -O2 -march=native -flto -mavx -fopenmp): only
    orc_trace is bound; bench.py's CPU legs use it, the parity tests never do.  It is built on the
    machine that runs it (-march=native)."""
    global _fast
    if _fast is None:
        so = os.path.join(ORACLE_DIR, "libort_oracle_fast.so")
        stamp = so + ".host"
        host = open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0] if os.path.exists("/proc/cpuinfo") else ""
        if (not os.path.exists(so)) or (not os.path.exists(stamp)) or open(stamp).read() != host:
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B", "libort_oracle_fast.so"])
            open(stamp, "w").write(host)
        L = C.CDLL(so)
        L.orc_trace.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int]
        _fast = L
    return _fast


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_load_plano.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(abi.Plano)]
        L.orc_load_doublet.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(abi.Doublet)]
        L.orc_load_bottle.argtypes = [C.c_char_p, C.c_double, C.POINTER(abi.Bottle)]
        L.orc_derive_scene.argtypes = [C.POINTER(abi.Scene), C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_double, C.c_double]
        L.orc_uniforms.argtypes = [C.c_uint64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, DP]
        L.orc_trace_rays.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.orc_trace.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int]
        L.orc_trace_volume.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.Scene), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
        L.orc_fresnel.restype = C.c_double
        L.orc_fresnel.argtypes = [DP, DP, C.c_double, C.c_double]
        L.orc_sellmeier.restype = C.c_double
        L.orc_sellmeier.argtypes = [C.c_double] * 7
        L.orc_refract.argtypes = [DP, DP, C.c_double]
        L.orc_reflect.argtypes = [DP, DP]
        L.orc_stokes.argtypes = [DP, C.c_double, C.c_uint64, C.c_int64]
        for name in ("orc_intersect_sphere", "orc_intersect_cylinder"):
            getattr(L, name).argtypes = [DP, DP, DP, C.c_double, DP]
        L.orc_intersect_ellipse.argtypes = [DP, DP, DP, C.c_double, C.c_double, DP]
        L.orc_load_image_source.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, C.c_void_p]
        L.orc_set_image_source.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("oracle %s failed: %d" % (what, rc))


def make_scene(bottle="clearBottle-large.params", l2="planoConvex-f39.9mm.params",
               l3="achromaticDoublet-f50.0mm.params", *, wavelength=785e-9, lens_wavelength=None,
               alpha_deg=5.0, n_axicon=1.45, ring_width=0.5e-3, resdir=RES, isors=False,
               isors_offset=1.5e-3, spot_size=1e-3):
    """Oracle's restatement of the scene set-up of reference src/setupMod.f90:113-119 +
    src/main.f90:51-70,81 (lens_wavelength=843e-9 gives the point-phase lenses, main.f90:113-117)."""
    L = lib()
    lw = wavelength if lens_wavelength is None else lens_wavelength
    S = abi.Scene()
    _chk(L.orc_load_bottle(os.path.join(resdir, bottle).encode(), wavelength, C.byref(S.bottle)),
         "load_bottle")
    _chk(L.orc_load_plano(os.path.join(resdir, l2).encode(), lw, 0.0, C.byref(S.L2)), "load_plano")
    off = 2. * S.L2.fb + S.L2.thickness
    _chk(L.orc_load_doublet(os.path.join(resdir, l3).encode(), lw, off, C.byref(S.L3)),
         "load_doublet")
    _chk(L.orc_derive_scene(C.byref(S), alpha_deg, n_axicon, ring_width, 1 if isors else 0,
                            isors_offset, spot_size), "derive_scene")
    return S


def set_jitter(seed):
    """conditioning probe of the scatter path: every libm result / intersection distance of tauint and
    stokes moved by a pseudo-random -2..+2 ulp (0 = off)"""
    lib().orc_set_jitter(C.c_uint64(int(seed)))


def uniforms(seed, phase, ray, first_slot, n):
    out = np.zeros(n, dtype=np.float64)
    lib().orc_uniforms(seed, phase, ray, first_slot, n, out.ctypes.data_as(DP))
    return out


def trace_rays(job, scene, n, pos_in=None, dir_in=None):
    """-> dict(pos[3,n], dir[3,n], status[n], bin[2,n])"""
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
    _chk(lib().orc_trace_rays(C.byref(job), C.byref(scene), n, pi, di, pos_out.ctypes.data,
                              dir_out.ctypes.data, status.ctypes.data, bins.ctypes.data),
         "trace_rays")
    return dict(pos=pos_out, dir=dir_out, status=status, bin=bins)


def trace(job, scenes, nthreads=0, fast=False):
    """-> image[nscenes,401,401] (uint64, [yp+200, xp+200]), lost[nscenes], hist[nscenes,32];
    fast=True: the timing twin built with the reference's optimisation flags"""
    if isinstance(scenes, abi.Scene):
        scenes = [scenes]
    ns = len(scenes)
    arr = (abi.Scene * ns)(*scenes)
    image = np.zeros((ns, abi.ORT_IMG_N, abi.ORT_IMG_N), dtype=np.uint64)
    lost = np.zeros(ns, dtype=np.int64)
    hist = np.zeros((ns, abi.ORT_NSTATUS), dtype=np.int64)
    _chk((fast_lib() if fast else lib()).orc_trace(C.byref(job), arr, ns, image.ctypes.data, lost.ctypes.data,
                                                   hist.ctypes.data, nthreads), "trace")
    return image, lost, hist


def trace_volume(job, scene, nthreads=0):
    """makeImage3D -> volume[200,401,401] uint32, lost, hist[32]"""
    vol = np.zeros((200, abi.ORT_IMG_N, abi.ORT_IMG_N), dtype=np.uint32)
    lost = np.zeros(1, dtype=np.int64)
    hist = np.zeros(abi.ORT_NSTATUS, dtype=np.int64)
    _chk(lib().orc_trace_volume(C.byref(job), C.byref(scene), vol.ctypes.data, lost.ctypes.data,
                                hist.ctypes.data, nthreads), "trace_volume")
    return vol, int(lost[0]), hist


def load_image_source(path, nphotons, seed=123456789):
    budget = np.zeros(512 * 512, dtype=np.int32)
    _chk(lib().orc_load_image_source(os.fsencode(path), int(nphotons), seed, budget.ctypes.data),
         "load_image_source")
    return budget


def set_image_source(budget):
    if budget is None:
        lib().orc_set_image_source(None)
    else:
        b = np.ascontiguousarray(budget, dtype=np.int32)
        lib().orc_set_image_source(b.ctypes.data)


def v3(a):
    return (C.c_double * 3)(*a)
