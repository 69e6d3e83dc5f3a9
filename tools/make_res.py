#!/usr/bin/env python3
"""Regenerate res/*.params -- the optical parameter library of the drop-in surface.

The *formats* are the reference's:
  plano-convex  12 positional lines   (reference src/lens.f90:146-159)
  doublet       21 positional lines   (reference src/lens.f90:92-114)
  bottle        12 lines + optional 4 (reference src/lens.f90:182-210)
  settings      20 positional lines   (reference src/setupMod.f90:57-133)
Only the first whitespace-separated token of a line is data; the rest is free text.
The numeric tokens below are catalogue facts (lens prescriptions, Sellmeier / Cauchy
coefficients, bottle sizes) spelled exactly as a list-directed Fortran read expects
(`d` exponents).  The trailing text on each line is ours.

Run:  python tools/make_res.py            (writes into ./res)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BK7 = ["1.03961212", "0.231792344", "1.01046945", "0.00600069867", "0.0200179144", "103.560653"]
NLAK22 = ["1.14229781", "0.535138441", "1.04088385", "0.00585778594", "0.0198546147", "100.834017"]
NBAF10 = ["1.5851495", "0.143559385", "1.08521269", "0.00926681282", "0.0424489805", "105.613573"]
NSF6 = ["1.72448482", "0.390104889", "1.04572858", "0.0134871947", "0.0569318095", "118.557185"]
SODALIME = ["1.5130", "0.003169", "0.003962"]
ETHANOL = ["1.35265", "0.00306", "0.00002"]

PLANO_LABELS = ["thickness [m]", "curve radius [m]", "diameter [m]", "focal length f [m]",
                "back focal length fb [m]", "n1 (surrounding medium)",
                "Sellmeier B1", "Sellmeier B2", "Sellmeier B3",
                "Sellmeier C1 [um^2]", "Sellmeier C2 [um^2]", "Sellmeier C3 [um^2]"]
DOUBLET_LABELS = ["thickness of element 1 [m]", "thickness of element 2 [m]",
                  "R1 [m]", "R2 [m]", "R3 [m]", "diameter [m]", "focal length f [m]",
                  "back focal length fb [m]", "n1 (surrounding medium)",
                  "element 1 Sellmeier B1", "element 1 Sellmeier B2", "element 1 Sellmeier B3",
                  "element 1 Sellmeier C1 [um^2]", "element 1 Sellmeier C2 [um^2]",
                  "element 1 Sellmeier C3 [um^2]",
                  "element 2 Sellmeier B1", "element 2 Sellmeier B2", "element 2 Sellmeier B3",
                  "element 2 Sellmeier C1 [um^2]", "element 2 Sellmeier C2 [um^2]",
                  "element 2 Sellmeier C3 [um^2]"]
BOTTLE_LABELS = ["wall thickness [m]", "radius a, along z [m]",
                 "radius b, along y [m] (a != b -> elliptical bottle)",
                 "centre x [m]", "centre y [m]", "centre z [m]",
                 "glass dispersion a", "glass dispersion b", "glass dispersion c",
                 "contents Cauchy a", "contents Cauchy b", "contents Cauchy c",
                 "mua wall [1/m]", "mus wall [1/m]", "mua contents [1/m]", "mus contents [1/m]"]
SETTINGS_LABELS = ["ring width [m]", "wavelength [m]", "number of rays per phase",
                   "axicon angle alpha [deg]", "axicon refractive index", "use bottle",
                   "use tracker", "make images", "image diameter [m]", "fibre offset [m]",
                   "source type: image | spot | point | isors | crs",
                   "iris position: before | after | none", "iris radius (fraction of lens radius)",
                   "bottle file", "L2 (plano-convex) file", "L3 (doublet) file",
                   "image-source file", "output folder under data/", "isors offset [m]",
                   "crs spot radius [m]"]


def emit(path, tokens, labels, final_newline=True):
    lines = ["%-16s # %s" % (t, l) for t, l in zip(tokens, labels)]
    text = "\n".join(lines) + ("\n" if final_newline else "")
    with open(path, "w") as fh:
        fh.write(text)


def main(out=None):
    out = out or os.path.join(ROOT, "res")
    os.makedirs(out, exist_ok=True)
    P = lambda name: os.path.join(out, name)

    # plano-convex lenses: thickness, R, diameter, f, fb
    planos = {
        "planoConvex.params": ["6.40d-3", "20.6d-3", "25.4d-3", "39.9d-3", "35.7d-3"],
        "planoConvex-f39.9mm.params": ["6.40d-3", "20.6d-3", "25.4d-3", "39.9d-3", "35.7d-3"],
        "planoConvex-f29.9mm.params": ["8.60d-3", "15.5d-3", "25.4d-3", "29.9d-3", "24.2d-3"],
        "planoConvex-f34.9mm.params": ["7.20d-3", "18.0d-3", "25.4d-3", "34.9d-3", "30.1d-3"],
        "planoConvex-f49.8mm.params": ["5.30d-3", "25.8d-3", "25.4d-3", "49.8d-3", "46.3d-3"],
        "planoConvex-f59.8mm.params": ["4.70d-3", "30.90d-3", "25.40d-3", "59.80d-3", "56.70d-3"],
        "planoConvex-smallf.params": ["3.50d-3", "12.90d-3", "25.40d-3", "24.90d-3", "22.60d-3"],
        "L1.params": ["3.60d-3", "51.50d-3", "25.4d-3", "99.70d-3", "97.30d-3"],
    }
    for name, geo in planos.items():
        emit(P(name), geo + ["1.0"] + BK7, PLANO_LABELS)

    # achromatic doublets: t1, t2, R1, R2, R3, diameter, f, fb ; glass pair
    doublets = {
        "achromaticDoublet.params": (["7.5d-3", "1.8d-3", "33.55d-3", "27.05d-3", "125.60d-3", "25.4d-3", "50d-3", "45d-3"], NLAK22),
        "achromaticDoublet-f50.0mm.params": (["7.5d-3", "1.8d-3", "33.55d-3", "27.05d-3", "125.60d-3", "25.4d-3", "50d-3", "45d-3"], NLAK22),
        "achromaticDoublet-f40.0mm.params": (["10.0d-3", "2.5d-3", "26.12d-3", "21.28d-3", "137.09d-3", "25.4d-3", "40.0d-3", "32.8d-3"], NBAF10),
        "achromaticDoublet-f45.0mm.params": (["7.8d-3", "1.6d-3", "29.38d-3", "25.05d-3", "127.06d-3", "25.4d-3", "45d-3", "39.6d-3"], NLAK22),
        "achromaticDoublet-f60.0mm.params": (["6.0d-3", "1.7d-3", "39.48d-3", "33.00d-3", "165.20d-3", "25.4d-3", "60d-3", "55.8d-3"], NLAK22),
        "achromaticDoublet-f75.0mm.params": (["5.0d-3", "1.6d-3", "36.90d-3", "42.17d-3", "417.8d-3", "25.4d-3", "75.0d-3", "69.9d-3"], NBAF10),
    }
    for name, (geo, glass1) in doublets.items():
        # like the reference's doublet files, the last line carries no newline
        emit(P(name), geo + ["1.0d0"] + glass1 + NSF6, DOUBLET_LABELS, final_newline=False)

    # bottles: thickness, Ra, Rb, x, y, z
    def bottle(name, th, ra, rb, z, extra=()):
        toks = [th, ra, rb, "0.0", "0.0", z] + SODALIME + ETHANOL + list(extra)
        emit(P(name), toks, BOTTLE_LABELS)

    bottle("clearBottle-small.params", "2.10d-3", "17.5d-3", "17.5d-3", "0.00")
    bottle("clearBottle-large.params", "2.10d-3", "35.0d-3", "35.0d-3", "-2.00d-3")
    bottle("clearBottle-ellipse.params", "2.10d-3", "35.0d-3", "17.5d-3", "0.00")
    bottle("clearBottle-ellipse-long.params", "2.10d-3", "35.0d-3", "17.5d-3", "0.00")
    bottle("clearBottle-ellipse-short.params", "2.10d-3", "17.5d-3", "35.0d-3", "0.00")
    for mm in range(-14, 15, 2):
        z = "0.00" if mm == 0 else "%d.00d-3" % mm
        th = "4.0d-3" if mm == 0 else "4.d-3"
        bottle("clearBottle-large_%dmm.params" % mm, th, "35.0d-3", "35.0d-3", z)
    for tenth in range(-175, 176, 25):
        mm = tenth / 10.0
        if tenth == 0:
            # the shipped 14-line file: two of the four optional mu lines
            bottle("clearBottle-small_0.0mm.params", "2.d-3", "17.5d-3", "17.5d-3", "0.0",
                   extra=["0.", "0.0"])
        else:
            z = ("%.4f" % (mm / 1000.0)).rstrip("0")
            bottle("clearBottle-small_%.1fmm.params" % mm, "2.10d-3", "17.5d-3", "17.5d-3", z)

    # synthetic scattering variants (SURVEY 8(c) fixtures / config 4): mua_b, mus_b, mua_c, mus_c
    bottle("scatterBottle-ellipse-long.params", "2.10d-3", "35.0d-3", "17.5d-3", "0.00",
           extra=["0.0", "0.0", "1.0", "30.0"])
    bottle("scatterBottle-large.params", "2.10d-3", "35.0d-3", "35.0d-3", "-2.00d-3",
           extra=["0.5", "20.0", "1.0", "30.0"])
    bottle("scatterBottle-small.params", "2.10d-3", "17.5d-3", "17.5d-3", "0.00",
           extra=["2.0", "40.0", "5.0", "60.0"])

    emit(P("settings.params"),
         ["0.5d-3", "785d-9", "100", "5", "1.45", "true", "true", "true", "1.d-2", "0.0", "crs",
          "none", "1.0", "clearBottle-small_0.0mm.params", "planoConvex-f39.9mm.params",
          "achromaticDoublet-f40.0mm.params", "bessel-normal.dat", "settings-testysors",
          "1.5d-3", "1.d-3"], SETTINGS_LABELS)
    # runnable job files for the BASELINE.json configs
    emit(P("settings-config1.params"),
         ["0.5d-3", "785d-9", "10000000", "5", "1.45", "true", "false", "true", "1.d-2", "0.0",
          "point", "none", "1.0", "clearBottle-small.params", "planoConvex.params",
          "achromaticDoublet.params", "bessel-normal.dat", "config1", "1.5d-3", "1.d-3"],
         SETTINGS_LABELS)
    emit(P("settings-config2.params"),
         ["0.5d-3", "785d-9", "1000000000", "5", "1.45", "true", "false", "true", "1.d-2", "0.0",
          "point", "none", "1.0", "clearBottle-large.params", "planoConvex-f39.9mm.params",
          "achromaticDoublet-f50.0mm.params", "bessel-normal.dat", "config2", "1.5d-3", "1.d-3"],
         SETTINGS_LABELS)
    emit(P("settings-config4.params"),
         ["0.5d-3", "785d-9", "100000000", "5", "1.45", "true", "false", "true", "1.d-2", "0.0",
          "point", "none", "1.0", "scatterBottle-ellipse-long.params", "planoConvex-f39.9mm.params",
          "achromaticDoublet-f50.0mm.params", "bessel-normal.dat", "config4", "1.5d-3", "1.d-3"],
         SETTINGS_LABELS)
    return out


if __name__ == "__main__":
    print(main(sys.argv[1] if len(sys.argv) > 1 else None))
