"""ctypes mirror of include/ort.h -- field for field, same order.

Kept free of any library loading so the product binding (opticalraytrace_b200.lib) and the
test-only checker binding under tests/ can share the struct definitions.
"""
import ctypes as C

ORT_IMG_N = 401
ORT_IMG_HALF = 200
ORT_IMG_BINS = ORT_IMG_N * ORT_IMG_N
ORT_NSTATUS = 32
VOL_DEPTH = 200

ORT_OK, ORT_EINVAL, ORT_ENODEVICE, ORT_ECUDA, ORT_ENCCL, ORT_EIO, ORT_EPARSE, ORT_ETRACE = (
    0, -1, -2, -3, -4, -5, -6, -7)

PHASE_RING, PHASE_POINT = 1, 2
FLAG_FIX_OUTER_ELLIPSE, FLAG_NO_REDUCE, FLAG_NO_COMPACTION = 1, 2, 4
FLAG_NO_FILTER, FLAG_VERIFY_FILTER = 8, 16
FLAG_ONE_LANE = 32
FILTER_SLOT_CALLED, FILTER_SLOT_WRONG = 30, 31
SCATTER_EVENTS_SLOT = 27
STOP_NONE, STOP_SOURCE, STOP_BOTTLE, STOP_L2, STOP_L3 = 0, 1, 2, 3, 4
SRC_POINT, SRC_CRS, SRC_ISORS, SRC_SPOT, SRC_IMAGE = 0, 1, 2, 3, 4
SOURCE_KINDS = {"point": SRC_POINT, "crs": SRC_CRS, "isors": SRC_ISORS, "spot": SRC_SPOT,
                "image": SRC_IMAGE}
SRCIMG_N = 512

STATUS_NAMES = [
    "binned", "bottle_inner_miss", "contents_absorbed", "contents_backward",
    "bottle_inner_reflect", "bottle_outer_miss", "wall_absorbed", "wall_backward",
    "bottle_outer_reflect", "l2_aperture", "l2_sphere_miss", "l2_curved_reflect",
    "l3_iris_before", "l3_s1_miss", "l3_aperture", "l3_s1_reflect", "l3_s2_miss",
    "l3_s2_reflect", "l3_s3_miss", "l3_s3_reflect", "l3_iris_after", "na_reject", "far",
    "off_detector", "tauint_miss", "stopped", "source_miss",
]
ST_BINNED = 0
ST_STOPPED = 25
NO_BIN = -2 ** 31


def status_is_lost(s):
    return (1 <= s <= 20) or s == 24 or s == 26


class Plano(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("thickness", "diameter", "radius", "fb", "f", "n1", "n2", "curve_radius")] + [
        ("centre", C.c_double * 3), ("flat_normal", C.c_double * 3)]


class Doublet(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("thickness", "diameter", "radius", "fb", "f", "n1", "n2", "n3",
                 "thickness1", "thickness2", "R1", "R2", "R3")] + [
        ("centre1", C.c_double * 3), ("centre2", C.c_double * 3), ("centre3", C.c_double * 3)]


class Bottle(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("nbottle", "ncontents", "thickness", "radiusa", "radiusb",
                 "mua_b", "mus_b", "mua_c", "mus_c")] + [
        ("centre", C.c_double * 3),
        ("ellipse", C.c_int32), ("scatter_b", C.c_int32), ("scatter_c", C.c_int32),
        ("_pad", C.c_int32)]


class Scene(C.Structure):
    _fields_ = [("bottle", Bottle), ("L2", Plano), ("L3", Doublet),
                ("cos_theta_max", C.c_double), ("r1", C.c_double), ("r2", C.c_double),
                ("img_plane", C.c_double), ("point_offset", C.c_double),
                ("spot_size", C.c_double), ("isors_offset", C.c_double), ("ring_width", C.c_double)]


class Job(C.Structure):
    _fields_ = [("phase", C.c_int32), ("use_bottle", C.c_int32), ("iris_before", C.c_int32),
                ("iris_after", C.c_int32), ("precision", C.c_int32), ("flags", C.c_int32),
                ("stop_after", C.c_int32), ("source_kind", C.c_int32),
                ("iris_radius", C.c_double), ("fibre_offset", C.c_double),
                ("image_diameter", C.c_double), ("uniform_override", C.c_double),
                ("seed", C.c_uint64), ("first_ray", C.c_int64), ("nrays", C.c_int64),
                ("total_rays", C.c_int64)]


class Timing(C.Structure):
    _fields_ = [("trace_seconds", C.c_double), ("reduce_seconds", C.c_double),
                ("wall_seconds", C.c_double), ("d2h_seconds", C.c_double),
                ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class Bpm(C.Structure):
    """ort_bpm: the parameters of the reference's bpm.py (lengths in micrometres)"""
    _fields_ = [(n, C.c_double) for n in
                ("w0", "wavelength", "axicon_deg", "n_axicon", "xymax", "ring_radius", "ring_width")] + [
        ("nxy", C.c_int32), ("nz", C.c_int32), ("steps", C.c_int32), ("reserved", C.c_int32)]


class Settings(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("ring_width", "wavelength", "alpha_deg", "n_axicon", "image_diameter",
                 "fibre_offset", "iris_radius", "isors_offset", "spot_size")] + [
        ("nphotons", C.c_int64),
        ("use_bottle", C.c_int32), ("use_tracker", C.c_int32), ("make_images", C.c_int32),
        ("iris_before", C.c_int32), ("iris_after", C.c_int32), ("_pad", C.c_int32),
        ("source_type", C.c_char * 64), ("iris_name", C.c_char * 64),
        ("bottle_file", C.c_char * 256), ("l2_file", C.c_char * 256),
        ("l3_file", C.c_char * 256), ("image_file", C.c_char * 256), ("folder", C.c_char * 256)]


def default_job(phase, nrays=0, *, use_bottle=True, iris="none", iris_radius=1.0,
                fibre_offset=0.0, image_diameter=1e-2, seed=123456789, first_ray=0,
                flags=0, stop_after=0, uniform_override=-1.0, source="point", total_rays=0):
    """A Job with the values the reference's settings.params ships (src/setupMod.f90:57-133)."""
    j = Job()
    j.phase = phase
    j.use_bottle = 1 if use_bottle else 0
    j.iris_before = 1 if iris == "before" else 0
    j.iris_after = 1 if iris == "after" else 0
    j.precision = 64
    j.flags = flags
    j.stop_after = stop_after
    j.source_kind = SOURCE_KINDS[source]
    j.total_rays = total_rays
    j.iris_radius = iris_radius
    j.fibre_offset = fibre_offset
    j.image_diameter = image_diameter
    j.uniform_override = uniform_override
    j.seed = seed
    j.first_ray = first_ray
    j.nrays = nrays
    return j
