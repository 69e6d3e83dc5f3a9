"""Host program around the trace loop -- what reference src/main.f90 does besides its two
`!$OMP do` loops: read settings.params, build the optics at 785 nm (ring phase) and at 843 nm
(point phase, src/main.f90:113-117), call the ray loop twice, write trans-stats.dat, print the
two transmission lines and write the three raw images.
"""
import os

import numpy as np

from . import _abi as abi
from . import lib


def partition(nrays, rank, nranks, first_ray=0):
    """Contiguous ray-index range of `rank`: [first + N*r/G, first + N*(r+1)/G).
    Uniforms depend only on the ray index, so the summed image is the same for any nranks."""
    lo = nrays * rank // nranks
    hi = nrays * (rank + 1) // nranks
    return first_ray + lo, hi - lo


def run(settings_path, resdir, datadir=None, *, nphotons=None, write=True, verbose=True):
    """One reference run.  Returns dict(ring=image, point=image, rcount, pcount, names...).
    The library must already be initialised (lib.init / lib.init_rank)."""
    st = lib.read_settings(settings_path)
    if nphotons is not None:
        st.nphotons = nphotons
    if verbose:
        print(" Using %s settings." % os.path.basename(settings_path))
    if st.use_tracker:  # src/setupMod.f90:75-82
        if st.nphotons > 10000:
            raise lib.OrtError(abi.ORT_EINVAL, "Too many photons for tracker use!")
        if st.make_images:
            if verbose:
                print(" ***************\n Cannot track packets and make images!\n"
                      " Deselecting makeImages\n ***************")
            st.make_images = 0
    src = st.source_type.decode()
    if src == "image":  # init_emit_image, src/setupMod.f90:120-121
        lib.set_image_source(lib.load_image_source(os.path.join(resdir, st.image_file.decode()), st.nphotons))
    scene_ring, pre_guard = lib.build_scene(st, resdir, st.wavelength)
    scene_point, _ = lib.build_scene(st, resdir, 843e-9)
    if verbose and scene_ring.bottle.centre[2] != pre_guard:
        print(" Bottle offset too large! Adjusting so that there is a minimum of 2mm offset from lens.")
        print(" Now bottle set at z position: %r" % scene_ring.bottle.centre[2])
    name = lib.output_basename(st, scene_ring, pre_guard)

    jr = lib.job_from_settings(st, abi.PHASE_RING)
    ring, rlost, rhist, rt = lib.trace(jr, scene_ring, allow_trap=True)
    jp = lib.job_from_settings(st, abi.PHASE_POINT)
    point, plost, phist, pt = lib.trace(jp, scene_point, allow_trap=True)
    rcount, pcount = int(rlost[0]), int(plost[0])
    n = float(st.nphotons) if st.nphotons else float("nan")
    out = dict(ring=ring[0], point=point[0], rcount=rcount, pcount=pcount, name=name,
               ring_hist=rhist[0], point_hist=phist[0], ring_timing=rt, point_timing=pt,
               settings=st, scene_ring=scene_ring, scene_point=scene_point)
    if write and datadir is not None:
        folder = os.path.join(datadir, st.folder.decode())
        os.makedirs(folder, exist_ok=True)
        if st.use_tracker:  # src/main.f90:72-74,121-124
            lib.write_tracks(jr, scene_ring, os.path.join(folder, name + "-ringtrace.dat"))
            lib.write_tracks(jp, scene_point, os.path.join(folder, name + "-pointtrace.dat"))
        lib.append_trans_stats(folder, st, scene_point, rcount, pcount)
        if st.make_images:
            lib.write_images(os.path.join(folder, name + "_image"), ring[0], point[0])
        out["folder"] = folder
    if verbose:
        print("Ring  transmitted:  %8.2f%%" % (100. * (1. - rcount / n)))
        print("Point transmitted:  %8.2f%%" % (100. * (1. - pcount / n)))
    return out
