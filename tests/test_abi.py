"""The C-ABI library loads on a machine without a GPU, exports every symbol include/ort.h
declares, its structs have the sizes the Python (and Fortran) mirrors assume, and every compute
entry fails loudly -- there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from opticalraytrace_b200 import abi
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "ort.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ort_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(ortlib):
    declared = _header_functions()
    assert len(declared) >= 19
    L = ortlib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(ortlib.EXPORTS) == declared


def test_struct_sizes_match_mirrors(ortlib):
    sizes = ortlib.struct_sizes()
    mirrors = [abi.Plano, abi.Doublet, abi.Bottle, abi.Scene, abi.Job, abi.Timing, abi.Settings]
    assert sizes[:7] == [C.sizeof(m) for m in mirrors]
    assert sizes[7] == 100


def test_fortran_interface_mirrors_header():
    """fortran/ort_interface.f90 binds the same names (it cannot be compiled here: no gfortran)."""
    path = os.path.join(ROOT, "fortran", "ort_interface.f90")
    if not os.path.exists(path):
        pytest.skip("fortran interface not written yet")
    text = open(path).read().lower()
    for name in ("ort_init", "ort_trace", "ort_trace_rays", "ort_finalize", "ort_last_error",
                 "ort_init_rank", "ort_struct_sizes"):
        assert 'name="%s"' % name in text.replace("'", '"'), name


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="needs a machine without a GPU")
def test_no_cpu_fallback(ortlib, orc):
    from opticalraytrace_b200.lib import OrtError
    assert ortlib.device_count() == 0
    with pytest.raises(OrtError) as e:
        ortlib.init(1)
    assert e.value.code == abi.ORT_ENODEVICE
    scene = cases.scene_for(orc, cases.C1, 1)
    with pytest.raises(OrtError) as e:
        ortlib.trace(abi.default_job(1, 10), scene)
    assert e.value.code == abi.ORT_ENODEVICE
    with pytest.raises(OrtError):
        ortlib.trace_rays(abi.default_job(1), scene, 4)
    with pytest.raises(OrtError):
        ortlib.uniforms(1, 1, 0, 0, 4)
    with pytest.raises(OrtError):
        ortlib.measure_fp64_peak()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "opticalraytrace_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "ort_oracle" not in text and "oracle_lib" not in text and "orc_" not in text, f
                assert "libhost_harness" not in text, f
