"""Volume image: makeImage3D / writeImage3D, src/imageMod.f90:61-90,117-133.
The reference never reaches these routines (its main passes a rank-3 image), so there is no
reference output to pin; the oracle restates the routine and the tests tie it to the 2-D path."""
import os

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases


def test_oracle_volume_is_consistent_with_the_2d_image(orc):
    """Depth 0 is the image plane itself, sampled WITHOUT the NA test: it must dominate the 2-D
    image bin by bin, differ from it by exactly the NA-rejected rays, and every deeper layer can
    only lose rays (a ray stops at its first sample outside the window)."""
    scene = cases.scene_for(orc, cases.C2, 2)
    job = abi.default_job(2, 150_000)
    vol, lost, hist = orc.trace_volume(job, scene)
    img, lost2, hist2 = orc.trace(job, scene)
    assert lost == int(lost2[0])
    assert np.all(vol[0] >= img[0])
    assert int(vol[0].sum()) == int(hist[0]) == int(hist2[0, 0] + hist2[0, 21])       # binned + NA-rejected
    assert int(hist[23]) == int(hist2[0, 22] + hist2[0, 23])                          # nothing in the window
    layer = vol.reshape(200, -1).sum(axis=1)
    assert np.all(np.diff(layer.astype(np.int64)) <= 0) and layer[-1] > 0
    assert np.array_equal(hist[1:21], hist2[0, 1:21])


@pytest.mark.gpu
@pytest.mark.parametrize("phase,files,kw", [(2, cases.C2, {}), (1, cases.C1, {}), (2, cases.C1, dict(use_bottle=False)),
                                            (2, cases.OTHER, dict(iris="after", iris_radius=0.5))])
def test_cuda_volume_equals_oracle(ort, orc, phase, files, kw):
    scene = cases.scene_for(orc, files, phase, kw)
    job = abi.default_job(phase, 300_000 if phase == 2 else 3_000_000, first_ray=77, **kw)
    vol, lost, hist = ort.trace_volume(job, scene)
    ovol, olost, ohist = orc.trace_volume(job, scene)
    assert lost == olost and np.array_equal(hist, ohist)
    assert int(vol.sum()) > 0 and np.array_equal(vol, ovol)


@pytest.mark.gpu
def test_volume_files(ort, orc, tmp_path):
    scene = cases.scene_for(orc, cases.C2, 2)
    vol, _, _ = ort.trace_volume(abi.default_job(2, 100_000), scene)
    base = str(tmp_path / "run")
    ort.write_volume(base, None, vol)
    assert not os.path.exists(base + "-vol-ring.dat")
    back = np.fromfile(base + "-vol-point.dat", dtype=np.float64)
    assert back.size == 200 * 401 * 401 and np.array_equal(back.reshape(vol.shape), vol.astype(np.float64))
    from opticalraytrace_b200.lib import OrtError
    j32 = abi.default_job(2, 10)
    j32.precision = 32
    with pytest.raises(OrtError, match="precision 64"):
        ort.trace_volume(j32, scene)
