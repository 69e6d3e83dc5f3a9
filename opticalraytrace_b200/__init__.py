"""opticalraytrace_b200 -- B200-native per-ray trace loop of OpticalRayTrace.

The product is the CUDA library `libort.so` behind the C-ABI of include/ort.h
(opticalraytrace_b200/csrc).  This package is the thin Python host: a ctypes binding (`lib`),
the struct mirror (`_abi`) and `raytrace.run`, which does what the reference's
`program raytrace` does around its two ray loops (src/main.f90).
"""
from . import _abi as abi  # noqa: F401
from . import lib  # noqa: F401
from .raytrace import run, partition  # noqa: F401

__all__ = ["abi", "lib", "run", "partition"]
