"""Beam-propagation pre-processor: the reference's `bpm.py`.

The reference script models the Bessel beam behind the axicon with an FFT split-step propagation
and writes `bessel-normal.dat`, the 512x512 fp64 intensity map that `source_type = image` samples
(`src/sourceMod.f90:363-408`; `res/` ships none, so the `image` source and `runner.py -b` cannot run
until it has been produced).  Here the propagation runs on the GPU behind the C-ABI
(`ort_bpm_bessel`, cuFFT + three small kernels); this module is the command line around it:

    python -m opticalraytrace_b200.bpm                      # writes res/bessel-normal.dat
    python -m opticalraytrace_b200.bpm -o res/bessel-smear.dat --steps 150

The script's matplotlib figure is not reproduced (plotting is out of scope).
"""
import argparse
import os
import sys

from . import lib


def main(argv=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ap = argparse.ArgumentParser(prog="python -m opticalraytrace_b200.bpm", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-o", "--output", default=os.path.join(root, "res", "bessel-normal.dat"),
                    help="file to write (raw fp64, nxy*nxy values)")
    ap.add_argument("--w0", type=float, help="beam waist [um] (bpm.py: 582*4)")
    ap.add_argument("--wavelength", type=float, help="[um] (0.785)")
    ap.add_argument("--axicon-deg", type=float, help="axicon angle [deg] (5)")
    ap.add_argument("--n-axicon", type=float, help="refractive index (1.45)")
    ap.add_argument("--xymax", type=float, help="lateral grid extent [um] (5000)")
    ap.add_argument("--ring-radius", type=float, help="[um] (1612)")
    ap.add_argument("--ring-width", type=float, help="[um] (300)")
    ap.add_argument("--nxy", type=int, help="grid points per side (512; the image source needs 512)")
    ap.add_argument("--nz", type=int, help="axial voxels (1000)")
    ap.add_argument("--steps", type=int, help="free-space steps (default nz/10)")
    args = ap.parse_args(argv)
    lib.init(1)
    try:
        p = lib.bpm_defaults()
        for name in ("w0", "wavelength", "axicon_deg", "n_axicon", "xymax", "ring_radius", "ring_width",
                     "nxy", "nz", "steps"):
            v = getattr(args, name)
            if v is not None:
                setattr(p, name, v)
        lib.bpm_write_file(args.output, p)
        print("wrote %s (%d x %d fp64)" % (args.output, p.nxy, p.nxy))
    finally:
        lib.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
