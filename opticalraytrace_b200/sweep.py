"""Experiment sweeps as batched launches.

The reference's `runner.py` runs its experiments one case at a time: write a settings file,
rebuild the program, start a process (`runner.py:26-47`).  Every case of an experiment that
shares the job-level switches (use_bottle, iris, source type ...) differs only in its optical
configuration, and `ort_trace` takes many configurations ("scenes") in one call -- so a sweep
becomes one batched call per ray loop instead of one build + process per case.  The outputs per
case are exactly what a single run writes (images, trans-stats.dat line, tracker files).

Experiments (same option letters as `runner.py:351-389`):
  -p  point/ring images for the four standard bottle set-ups        (runner.py:136-155)
  -s  spot diagrams with the tracker                                 (runner.py:113-133)
  -i  iris position x size experiment                                (runner.py:158-186)
  -o  bottle offset experiment on the large bottle                   (runner.py:189-206)
  -l  L2 x L3 focal-length experiment                                (runner.py:231-261, :394-397)
  -b  bessel images: the `image` source for the four set-ups         (runner.py:209-228);
      needs res/bessel-smear.dat -- `python -m opticalraytrace_b200.bpm -o res/bessel-smear.dat`
  --isb  iSORS against Bessel illumination over seven offsets        (runner.py:266-320; the
      reference defines this one without wiring it to an option)
(-bp is a stub in the reference, runner.py:264-265, and stays one here.)
"""
import argparse
import collections
import math
import os
import shutil
import sys
import tempfile

from . import _abi as abi
from . import lib

BOTTLES = [("clearBottle-large.params", True), ("clearBottle-small.params", True),
           ("clearBottle-ellipse.params", True), ("clearBottle-small.params", False)]


def experiment_cases(name):
    """-> list of keyword dicts for lib.make_settings, one per case (runner.py's loops)."""
    out = []
    if name == "point":
        for b, ub in BOTTLES:
            out.append(dict(bottle=b, use_bottle=ub, folder="images"))
    elif name == "spot":
        for b, ub in BOTTLES:
            out.append(dict(bottle=b, use_bottle=ub, folder="spot-diag", source_type="spot",
                            nphotons=100, use_tracker=True))
    elif name == "iris":
        for b, ub in BOTTLES:
            for iris in ("before", "after", "none"):
                for size in (1.0, 0.8, 0.6, 0.4, 0.2):
                    out.append(dict(bottle=b, use_bottle=ub, folder="iris", iris=iris, iris_radius=size))
                    if iris == "none":
                        break
    elif name == "offset":
        for mm in range(4, 17, 2):
            out.append(dict(bottle="clearBottle-large_-%dmm.params" % mm, folder="images-offset"))
    elif name == "lens":
        for l3 in ("40.0", "45.0", "50.0", "60.0", "75.0"):
            for l2 in ("59.8", "49.8", "39.9", "34.9", "29.9"):
                for b, ub in BOTTLES[:3]:          # runner.py:394-396 drops the no-bottle set-up
                    out.append(dict(bottle=b, use_bottle=ub, folder="images-lens", make_images=False,
                                    l2="planoConvex-f%smm.params" % l2,
                                    l3="achromaticDoublet-f%smm.params" % l3))
    elif name == "bessel":
        for b, ub in BOTTLES:
            out.append(dict(bottle=b, use_bottle=ub, folder="images", source_type="image",
                            image_file="bessel-smear.dat"))
    elif name == "isb":
        for source in ("isors", "point"):
            for k in range(7):
                offset = 1.5e-3 if k == 6 else k * (1.5e-3 / 6)      # np.linspace(0, 1.5e-3, 7)
                kw = dict(folder="iSORS_vs_Bessel", source_type=source, isors_offset=offset,
                          bottle="clearBottle-small_0.0mm.params")
                if source == "point":   # move the bottle so that the Bessel ring lands at that offset
                    kw["bottle_z"] = offset
                out.append(kw)
    else:
        raise ValueError("unknown experiment %r" % name)
    return out


def bessel_bottle_z(resdir, offset, l2="planoConvex-f39.9mm.params", ring_width=0.5e-3, alpha_deg=5.0,
                    n_axicon=1.45, bottle="clearBottle-small_0.0mm.params"):
    """Bottle z position that puts the Bessel ring at the iSORS spatial offset (runner.py:280-312):
    fb (offset + ring width) / (d tan(alpha (n_axicon - 1))) - radius a, d = 97.3 mm axicon to L1."""
    def first_number(path, line_no):
        with open(os.path.join(resdir, path)) as f:
            return float(f.readlines()[line_no].split()[0].lower().replace("d", "e"))
    fb = first_number(l2, 4)         # planoConvex line 5: back focal length
    radius_a = first_number(bottle, 1)
    alpha = math.radians(alpha_deg)
    return fb * (offset + ring_width) / (97.3e-3 * math.tan(alpha * (n_axicon - 1.0))) - radius_a


def _moved_bottle(resdir, overlay, src, z, tag):
    """Write a copy of bottle file `src` with its centre z replaced (what runner.py's
    create_bottle_file does for the iSORS experiment) into the overlay directory."""
    with open(os.path.join(resdir, src)) as f:
        lines = f.readlines()
    rest = lines[5].split(None, 1)
    lines[5] = "%r %s" % (z, rest[1] if len(rest) > 1 else "\n")
    name = "%s_iSORS-%s.params" % (os.path.splitext(src)[0].split("_")[0], tag)
    with open(os.path.join(overlay, name), "w") as f:
        f.writelines(lines)
    return name


def _job_key(st):
    return (st.use_bottle, st.iris_before, st.iris_after, st.iris_radius, st.source_type,
            st.fibre_offset, st.image_diameter, st.nphotons, st.image_file)


def run_sweep(cases, resdir, datadir=None, nphotons=None, verbose=True):
    """Trace every case; cases that share the job-level switches go through ort_trace together.
    Returns a list of result dicts (ring, point images, rcount, pcount, name, folder) in case order.
    The library must be initialised."""
    prepared, skipped = [], []
    overlay = None
    if any("bottle_z" in kw for kw in cases):
        # generated bottle files live in a scratch copy of res/ (the reference writes them into res/)
        overlay = tempfile.mkdtemp(prefix="ort-res-")
        for f in os.listdir(resdir):
            os.symlink(os.path.join(os.path.abspath(resdir), f), os.path.join(overlay, f))
    try:
        return _run_sweep(cases, resdir, overlay, datadir, nphotons, verbose, prepared, skipped)
    finally:
        if overlay is not None:
            shutil.rmtree(overlay, ignore_errors=True)


def _run_sweep(cases, resdir, overlay, datadir, nphotons, verbose, prepared, skipped):
    if overlay is not None:
        base_res, resdir = resdir, overlay
    for i, kw in enumerate(cases):
        kw = dict(kw)
        tracker = kw.pop("use_tracker", False)
        make_images = kw.pop("make_images", True)
        image_file = kw.pop("image_file", None)
        if "bottle_z" in kw:
            off = kw.pop("bottle_z")
            z = bessel_bottle_z(base_res, off, l2=kw.get("l2", "planoConvex-f39.9mm.params"), bottle=kw["bottle"])
            kw["bottle"] = _moved_bottle(base_res, overlay, kw["bottle"], z, "%d" % i)
        kw.setdefault("nphotons", nphotons if nphotons is not None else 1_000_000_000)
        st = lib.make_settings(**kw)
        if image_file is not None:
            st.image_file = image_file.encode()
        st.use_tracker = int(tracker)
        st.make_images = int(make_images and not tracker)
        if not os.path.exists(os.path.join(resdir, st.bottle_file.decode())):
            skipped.append(st.bottle_file.decode())   # runner.py asks for a -16 mm file nobody ships
            prepared.append(None)
            continue
        ring_scene, pre = lib.build_scene(st, resdir, st.wavelength)
        point_scene, _ = lib.build_scene(st, resdir, 843e-9)
        prepared.append(dict(i=i, st=st, ring_scene=ring_scene, point_scene=point_scene,
                             name=lib.output_basename(st, ring_scene, pre)))
    groups = collections.OrderedDict()
    for c in prepared:
        if c is not None:
            groups.setdefault(_job_key(c["st"]), []).append(c)
    launches = 0
    for key, members in groups.items():
        st0 = members[0]["st"]
        if st0.source_type.decode() == "image":     # init_emit_image, src/setupMod.f90:120-121
            lib.set_image_source(lib.load_image_source(os.path.join(resdir, st0.image_file.decode()),
                                                       st0.nphotons))
        for phase, scene_key, img_key, cnt_key in ((abi.PHASE_RING, "ring_scene", "ring", "rcount"),
                                                   (abi.PHASE_POINT, "point_scene", "point", "pcount")):
            job = lib.job_from_settings(st0, phase)
            img, lost, hist, tm = lib.trace(job, [m[scene_key] for m in members], allow_trap=True)
            launches += 1
            for k, m in enumerate(members):
                m[img_key] = img[k]
                m[cnt_key] = int(lost[k])
                m[img_key + "_seconds"] = tm.trace_seconds / len(members)
    if datadir is not None:
        for c in prepared:
            if c is None:
                continue
            st = c["st"]
            folder = os.path.join(datadir, st.folder.decode())
            os.makedirs(folder, exist_ok=True)
            c["folder"] = folder
            if st.use_tracker:
                lib.write_tracks(lib.job_from_settings(st, abi.PHASE_RING), c["ring_scene"],
                                 os.path.join(folder, c["name"] + "-ringtrace.dat"))
                lib.write_tracks(lib.job_from_settings(st, abi.PHASE_POINT), c["point_scene"],
                                 os.path.join(folder, c["name"] + "-pointtrace.dat"))
            lib.append_trans_stats(folder, st, c["point_scene"], c["rcount"], c["pcount"])
            if st.make_images:
                lib.write_images(os.path.join(folder, c["name"] + "_image"), c["ring"], c["point"])
    if verbose:
        n_ok = sum(1 for c in prepared if c is not None)
        print("sweep: %d cases in %d batched ort_trace calls (%d job groups)%s"
              % (n_ok, launches, len(groups),
                 "; skipped missing files: %s" % ", ".join(skipped) if skipped else ""))
    return prepared


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m opticalraytrace_b200.sweep", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-s", "--spot", action="store_true", help="spot diagrams")
    ap.add_argument("-p", "--point", action="store_true", help="point/ring images")
    ap.add_argument("-o", "--offset", action="store_true", help="offset experiment on the large bottle")
    ap.add_argument("-i", "--iris", action="store_true", help="iris experiment")
    ap.add_argument("-l", "--lens", action="store_true", help="lens experiment")
    ap.add_argument("-b", "--bessel", action="store_true", help="bessel/ring images (needs res/bessel-smear.dat)")
    ap.add_argument("--isb", action="store_true", help="iSORS vs Bessel offsets experiment")
    ap.add_argument("-bp", "--bessel_params", action="store_true", help="(a stub in the reference too)")
    ap.add_argument("-a", "--all", action="store_true", help="all of the above")
    ap.add_argument("-n", "--nphotons", type=int, default=1_000_000_000, help="rays per loop per case")
    ap.add_argument("--gpus", type=int, default=0, help="devices to use (0 = all visible)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--res", default=os.path.join(root, "res"))
    ap.add_argument("--data", default=os.path.join(root, "data"))
    args = ap.parse_args(argv)
    # -a covers what runner.py's -a covers (runner.py:384-397); --isb has to be asked for
    chosen = [n for n in ("bessel", "point", "spot", "offset", "iris", "lens") if args.all or getattr(args, n)]
    if args.isb:
        chosen.append("isb")
    if not chosen:
        ap.print_help()
        return 0
    lib.init(args.gpus)
    try:
        for name in chosen:
            cases = experiment_cases(name)
            if name == "bessel" and args.all and not os.path.exists(os.path.join(args.res, cases[0]["image_file"])):
                print("sweep: no %s in %s, bessel images skipped" % (cases[0]["image_file"], args.res))
                continue
            run_sweep(cases, args.res, args.data, nphotons=args.nphotons)
    finally:
        lib.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
