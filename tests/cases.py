"""The optical configurations the parity tests sweep (names follow BASELINE.json configs)."""
from opticalraytrace_b200 import abi

C1 = ("clearBottle-small.params", "planoConvex.params", "achromaticDoublet.params")
C2 = ("clearBottle-large.params", "planoConvex-f39.9mm.params", "achromaticDoublet-f50.0mm.params")
ELL = ("clearBottle-ellipse-long.params", "planoConvex-f39.9mm.params", "achromaticDoublet-f50.0mm.params")
ELLS = ("clearBottle-ellipse-short.params", "planoConvex-f49.8mm.params", "achromaticDoublet-f75.0mm.params")
SC_L = ("scatterBottle-large.params", "planoConvex-f39.9mm.params", "achromaticDoublet-f50.0mm.params")
SC_S = ("scatterBottle-small.params", "planoConvex.params", "achromaticDoublet.params")
SC_E = ("scatterBottle-ellipse-long.params", "planoConvex-f39.9mm.params", "achromaticDoublet-f50.0mm.params")
OTHER = ("clearBottle-small_-5.0mm.params", "planoConvex-f29.9mm.params", "achromaticDoublet-f40.0mm.params")
OTHER2 = ("clearBottle-large_-10mm.params", "planoConvex-f59.8mm.params", "achromaticDoublet-f60.0mm.params")

# (id, files, phase, job kwargs)
RAY_CASES = [
    ("c1-ring", C1, abi.PHASE_RING, {}),
    ("c1-point", C1, abi.PHASE_POINT, {}),
    ("c2-ring", C2, abi.PHASE_RING, {}),
    ("c2-point", C2, abi.PHASE_POINT, {}),
    ("c2-point-nobottle", C2, abi.PHASE_POINT, dict(use_bottle=False)),
    ("c1-point-iris-before", C1, abi.PHASE_POINT, dict(iris="before", iris_radius=0.6)),
    ("c1-ring-iris-after", C1, abi.PHASE_RING, dict(iris="after", iris_radius=0.5)),
    ("c2-point-iris-after", C2, abi.PHASE_POINT, dict(iris="after", iris_radius=0.7)),
    ("c2-point-fibre-offset", C2, abi.PHASE_POINT, dict(fibre_offset=2e-3, image_diameter=5e-3)),
    ("ellipse-ring", ELL, abi.PHASE_RING, {}),
    ("ellipse-point", ELL, abi.PHASE_POINT, {}),
    ("ellipse-point-fixed", ELL, abi.PHASE_POINT, dict(flags=abi.FLAG_FIX_OUTER_ELLIPSE)),
    ("ellipse-short-point-fixed", ELLS, abi.PHASE_POINT, dict(flags=abi.FLAG_FIX_OUTER_ELLIPSE)),
    ("other-ring", OTHER, abi.PHASE_RING, {}),
    ("other-point", OTHER, abi.PHASE_POINT, {}),
    ("other2-point", OTHER2, abi.PHASE_POINT, {}),
]
SCATTER_CASES = [
    ("scatter-large", SC_L, abi.PHASE_POINT, {}),
    ("scatter-small", SC_S, abi.PHASE_POINT, {}),
    ("scatter-ellipse-fixed", SC_E, abi.PHASE_POINT, dict(flags=abi.FLAG_FIX_OUTER_ELLIPSE)),
    ("scatter-ellipse-faithful", SC_E, abi.PHASE_POINT, {}),
]


# the other source types of settings.params: crs and isors change the ring
# loop's emitter, spot the point loop's; isors also moves the point source to the bottle centre
SOURCE_CASES = [
    ("c1-crs-ring", C1, abi.PHASE_RING, dict(source="crs")),
    ("c2-crs-ring", C2, abi.PHASE_RING, dict(source="crs")),
    ("ellipse-crs-ring", ELL, abi.PHASE_RING, dict(source="crs")),
    ("c1-isors-ring", C1, abi.PHASE_RING, dict(source="isors")),
    ("c2-isors-ring", C2, abi.PHASE_RING, dict(source="isors")),
    ("ellipse-isors-ring", ELL, abi.PHASE_RING, dict(source="isors")),
    ("other-isors-point", OTHER, abi.PHASE_POINT, dict(source="isors")),
    ("c1-spot-point", C1, abi.PHASE_POINT, dict(source="spot", total_rays=250_000)),
    ("c2-spot-point-nobottle", C2, abi.PHASE_POINT, dict(source="spot", total_rays=1_000_000, use_bottle=False)),
    ("c2-spot-ring", C2, abi.PHASE_RING, dict(source="spot")),
]


def scene_for(orc, files, phase, kw=None):
    lam = 843e-9 if phase == abi.PHASE_POINT else None
    isors = bool(kw) and kw.get("source") == "isors"
    return orc.make_scene(*files, lens_wavelength=lam, isors=isors)
