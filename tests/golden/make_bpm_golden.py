"""Golden vectors for the beam-propagation pre-processor.

Runs the reference's own `bpm.py` in THIS container (it is a plain numpy script; matplotlib is not
installed here and only draws the final figure, so a do-nothing stand-in is put in its place) and
keeps a fingerprint of the 512x512 intensity it writes to `bessel-normal.dat`: the central row and
column, an 8-strided subsample, sum and maximum.  /root/reference does not exist on the GPU box --
only the committed tests/golden/bpm_v1.npz travels.

    python tests/golden/make_bpm_golden.py
"""
import os
import runpy
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/bpm.py"


def main():
    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter((_Anything(), [_Anything(), _Anything()]))

        def __getitem__(self, i):
            return _Anything()

    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = lambda *a, **k: (_Anything(), [_Anything(), _Anything()])
    plt.show = lambda *a, **k: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            runpy.run_path(REF, run_name="__main__")
            img = np.fromfile("bessel-normal.dat", dtype=np.float64).reshape(512, 512)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "bpm_v1.npz"), row=img[256, :], col=img[:, 256],
                        sub=img[::8, ::8].copy(), total=img.sum(), peak=img.max(),
                        argmax=np.array(np.unravel_index(np.argmax(img), img.shape)))
    print("bpm golden: sum %.12e peak %.12e at %s" % (img.sum(), img.max(), np.unravel_index(np.argmax(img), img.shape)))


if __name__ == "__main__":
    main()
