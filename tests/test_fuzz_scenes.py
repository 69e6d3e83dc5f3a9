"""Randomised optical configurations: perturbed bottles / lenses / job switches, compared ray by ray
with the oracle -- the host-compiled product math here, the CUDA path on the GPU box.  Catches
anything that only holds for the shipped geometries (e.g. the ring aim-plane shortcut must switch
itself off when L2 is moved; lens centres off the axis; elliptical bottles of any aspect)."""
import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases
from tests.conftest import rel_err

NSCENES = 40


def random_case(orc, k):
    rng = np.random.default_rng(1000 + k)
    files = [cases.C1, cases.C2, cases.ELL, cases.ELLS, cases.OTHER, cases.OTHER2, cases.SC_L, cases.SC_S][k % 8]
    phase = 1 + (k // 8) % 2
    scene = cases.scene_for(orc, files, phase)
    b, l2, l3 = scene.bottle, scene.L2, scene.L3
    # bottle: size, wall, position, indices
    s = rng.uniform(0.8, 1.2)
    b.radiusa *= s
    b.radiusb *= s * rng.uniform(0.9, 1.1) if b.ellipse else s
    b.thickness *= rng.uniform(0.5, 1.5)
    b.centre[1] = rng.uniform(-5e-4, 5e-4)
    b.centre[2] += rng.uniform(-1e-3, 1e-3)
    b.nbottle += rng.uniform(-0.03, 0.03)
    b.ncontents += rng.uniform(-0.03, 0.03)
    if b.scatter_c:
        b.mus_c *= rng.uniform(0.3, 2.0)
    # lenses: indices, small decentre / shift (shift disables the ring shortcut)
    l2.n2 += rng.uniform(-0.02, 0.02)
    l3.n2 += rng.uniform(-0.02, 0.02)
    l3.n3 += rng.uniform(-0.02, 0.02)
    if k % 3 == 0:
        l2.centre[2] += rng.uniform(-5e-4, 5e-4)
    if k % 4 == 0:
        l2.centre[0] = rng.uniform(-2e-4, 2e-4)
        l3.centre2[1] = rng.uniform(-2e-4, 2e-4)
    l3.R3 *= rng.uniform(0.9, 1.1)
    scene.img_plane += rng.uniform(-2e-3, 2e-3)
    kw = dict(iris=["none", "before", "after"][k % 3], iris_radius=float(rng.uniform(0.3, 1.0)),
              fibre_offset=float(rng.uniform(-1e-3, 1e-3)), image_diameter=float(rng.uniform(4e-3, 2e-2)),
              use_bottle=bool(k % 5), seed=int(rng.integers(1, 2 ** 62)),
              flags=abi.FLAG_FIX_OUTER_ELLIPSE if (b.ellipse and k % 2) else 0)
    return scene, phase, kw


def _check(a, b, scatter):
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["bin"], b["bin"])
    e = np.maximum(rel_err(a["pos"], b["pos"]), rel_err(a["dir"], b["dir"]))
    if scatter:
        assert np.nanmax(e) < 1e-6 and np.nanquantile(e, 0.999) < 1e-9
    else:
        assert np.nanmax(e) < 1e-9


@pytest.mark.parametrize("k", range(NSCENES))
def test_fuzz_host_math(orc, harness, k):
    scene, phase, kw = random_case(orc, k)
    job = abi.default_job(phase, first_ray=k * 10 ** 9, **kw)
    n = 30_000
    _check(orc.trace_rays(job, scene, n), harness(job, scene, n), scene.bottle.scatter_c or scene.bottle.scatter_b)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(NSCENES))
def test_fuzz_cuda(ort, orc, k):
    scene, phase, kw = random_case(orc, k)
    n = 60_000
    job = abi.default_job(phase, first_ray=k * 10 ** 9, **kw)
    _check(orc.trace_rays(job, scene, n), ort.trace_rays(job, scene, n),
           scene.bottle.scatter_c or scene.bottle.scatter_b)
    job.nrays = 150_001
    img, lost, hist, _ = ort.trace(job, scene, allow_trap=True)
    oimg, olost, ohist = orc.trace(job, scene)
    assert np.array_equal(hist, ohist) and np.array_equal(img, oimg) and np.array_equal(lost, olost)
