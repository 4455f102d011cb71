"""vkhel_b200 -- B200-native implementation of vkhel's NTT hot path.

The product is the C-ABI shared library vkhel_b200/lib/libvkhel.so (sources in
vkhel_b200/csrc, public headers in include/vkhel).  This package is only the
ctypes binding used by the Python tests and bench.py; it mirrors the C API
one to one (same names, argument order and meaning as include/vkhel/vkhel.h
and vkhel_ext.h) and adds no compute of its own.  There is no CPU fallback:
importing works without a GPU (so that symbol checks can run), but creating a
context aborts loudly when no CUDA device is present.
"""
from .api import (  # noqa: F401
    LIB_PATH, lib, build, Context, Vector, NttTables, host_alloc, device_count,
)
from . import params  # noqa: F401
