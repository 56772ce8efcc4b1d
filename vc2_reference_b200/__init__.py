"""B200-native VC-2 HQ/LD hot path.

The product is the C-ABI shared library ``libvc2b200.so`` (hand-written sm_100a CUDA kernels
behind ``include/vc2_cabi.h``) plus the C++ host layer that mirrors the bbc/vc2-reference
Library headers and command lines.  This Python package is only the thin ctypes binding the
tests and ``bench.py`` drive it through; it contains no codec arithmetic and no CPU fallback.
"""
from ._cabi import Vc2Error, lib, lib_path  # noqa: F401
from .api import (  # noqa: F401
    KERNELS, Context, Codec, Geom, quant_matrix, slice_bytes, padded_size, make_geom, hq_index_slices,
)
