"""ctypes binding of libvkhel.so -- the reference-side stub a maintainer would
write for a Python caller.  Every method is a direct call
of the C entry point of the same name (include/vkhel/vkhel.h,
include/vkhel/vkhel_ext.h)."""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# $VKHEL_LIB_PATH lets kernel experiments load an alternative build
LIB_PATH = os.environ.get("VKHEL_LIB_PATH",
                          os.path.join(_PKG, "lib", "libvkhel.so"))

_u64 = ctypes.c_uint64
_p64 = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p


def build():
    """Compile the library in-tree (nvcc, sm_100a)."""
    subprocess.check_call(["make", "-C", _ROOT, "-s", "-j8", "libs"])


# name -> (restype, argtypes); the 18 reference entry points first
SIGNATURES = {
    "vkhel_ctx_create": (_vp, []),
    "vkhel_ctx_destroy": (None, [_vp]),
    "vkhel_ntt_tables_create": (_vp, [_u64, _u64, _u64]),
    "vkhel_ntt_tables_destroy": (None, [_vp]),
    "vkhel_vector_create": (_vp, [_vp, _u64]),
    "vkhel_vector_create2": (_vp, [_vp, _u64, ctypes.c_bool]),
    "vkhel_vector_destroy": (None, [_vp]),
    "vkhel_vector_dup": (_vp, [_vp]),
    "vkhel_vector_copy_from_host": (None, [_vp, _p64]),
    "vkhel_vector_map": (None, [_vp, ctypes.POINTER(_vp), ctypes.c_size_t]),
    "vkhel_vector_unmap": (None, [_vp]),
    "vkhel_vector_elemfma": (None, [_vp, _vp, _vp, _u64, _u64]),
    "vkhel_vector_elemmod": (None, [_vp, _vp, _u64, _u64]),
    "vkhel_vector_elemmul": (None, [_vp, _vp, _vp, _u64]),
    "vkhel_vector_elemgtadd": (None, [_vp, _vp, _u64, _u64]),
    "vkhel_vector_elemgtsub": (None, [_vp, _vp, _u64, _u64, _u64]),
    "vkhel_vector_forward_transform": (None, [_vp, _vp, _vp]),
    "vkhel_vector_inverse_transform": (None, [_vp, _vp, _vp]),
    # private-header symbols that match the export glob
    "vkhel_vector_dbgprint": (None, [_vp]),
    "vkhel_ntt_tables_dbgprint": (None, [_vp]),
    # extensions (vkhel_ext.h)
    "vkhel_device_count": (ctypes.c_int, []),
    "vkhel_ctx_create_device": (_vp, [ctypes.c_int]),
    "vkhel_ctx_device": (ctypes.c_int, [_vp]),
    "vkhel_ctx_sync": (None, [_vp]),
    "vkhel_ctx_stream": (_vp, [_vp]),
    "vkhel_host_alloc": (_vp, [ctypes.c_size_t]),
    "vkhel_host_free": (None, [_vp]),
    "vkhel_vector_length": (_u64, [_vp]),
    "vkhel_vector_device_ptr": (_vp, [_vp]),
    "vkhel_vector_map_range": (None, [_vp, ctypes.POINTER(_vp), _u64, _u64]),
    "vkhel_ctx_readahead_hits": (_u64, [_vp]),
    "vkhel_ctx_lazy_forwards": (_u64, [_vp]),
    "vkhel_vector_upload": (None, [_vp, _vp, _u64, _u64]),
    "vkhel_vector_download": (None, [_vp, _vp, _u64, _u64]),
    "vkhel_vector_copy_peer": (None, [_vp, _u64, _vp, _u64, _u64]),
    "vkhel_vector_forward_transform_batch": (None, [_vp, _vp, _vp, _u64]),
    "vkhel_vector_inverse_transform_batch": (None, [_vp, _vp, _vp, _u64]),
    "vkhel_vector_forward_transform_rns": (None, [_vp, _vp,
                                                  ctypes.POINTER(_vp), _u64,
                                                  _u64]),
    "vkhel_vector_inverse_transform_rns": (None, [_vp, _vp,
                                                  ctypes.POINTER(_vp), _u64,
                                                  _u64]),
    "vkhel_vector_elemmul_rns": (None, [_vp, _vp, _vp, _p64, _u64, _u64,
                                        _u64]),
    "vkhel_vector_polymul_rns": (None, [_vp, _vp, _vp, ctypes.POINTER(_vp),
                                        _u64, _u64]),
    "vkhel_timer_create": (_vp, [_vp]),
    "vkhel_timer_start": (None, [_vp]),
    "vkhel_timer_stop": (None, [_vp]),
    "vkhel_timer_elapsed_ms": (ctypes.c_double, [_vp]),
    "vkhel_timer_destroy": (None, [_vp]),
    "vkhel_ctx_launch_count": (_u64, [_vp]),
    "vkhel_ctx_launch_count_noflush": (_u64, [_vp]),
    "vkhel_ctx_flush_l2": (None, [_vp]),
    "vkhel_ctx_flush": (None, [_vp]),
    "vkhel_ntt_tables_create_on": (_vp, [_vp, _u64, _u64, _u64]),
    "vkhel_ctx_deferred_stats": (None, [_vp, _p64, _p64]),
    "vkhel_ctx_fused_products": (ctypes.c_uint64, [_vp]),
    "vkhel_ctx_probe_int_peaks": (ctypes.c_int, [_vp,
                                                 ctypes.POINTER(ctypes.c_double),
                                                 ctypes.c_int]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvkhel.so is not built (%s): run `make` or "
                "__graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def device_count():
    return lib().vkhel_device_count()


def _as_u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


class PinnedArray:
    """numpy view of page-locked host memory from vkhel_host_alloc."""

    def __init__(self, count):
        self.count = count
        self.ptr = lib().vkhel_host_alloc(count * 8)
        buf = (ctypes.c_uint64 * count).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint64)

    def free(self):
        if self.ptr:
            self.array = None
            lib().vkhel_host_free(self.ptr)
            self.ptr = None


def host_alloc(count):
    return PinnedArray(count)


class NttTables:
    """struct vkhel_ntt_tables (vkhel_ntt_tables_create / _destroy)"""

    def __init__(self, n, q, w, ctx=None):
        """ctx given: vkhel_ntt_tables_create_on (generated on that GPU)"""
        self.n, self.q, self.w = n, q, w
        if ctx is None:
            self.handle = lib().vkhel_ntt_tables_create(n, q, w)
        else:
            self.handle = lib().vkhel_ntt_tables_create_on(ctx.handle, n, q, w)

    def _field(self, index):
        # struct layout of include/priv/ntt_tables.h: n, q, w, then 4 pointers
        raw = ctypes.cast(self.handle, ctypes.POINTER(ctypes.c_uint64))
        addr = raw[3 + index]
        arr = (ctypes.c_uint64 * self.n).from_address(addr)
        return np.frombuffer(arr, dtype=np.uint64).copy()

    roots_of_unity = property(lambda self: self._field(0))
    inv_roots_of_unity = property(lambda self: self._field(1))
    roots_barrett_factors = property(lambda self: self._field(2))
    inv_roots_barrett_factors = property(lambda self: self._field(3))

    def destroy(self):
        if self.handle:
            lib().vkhel_ntt_tables_destroy(self.handle)
            self.handle = None


def _table_array(tables):
    return (_vp * len(tables))(*[t.handle for t in tables])


class Vector:
    """struct vkhel_vector"""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.handle = handle

    @property
    def length(self):
        return int(lib().vkhel_vector_length(self.handle))

    def destroy(self):
        if self.handle:
            lib().vkhel_vector_destroy(self.handle)
            self.handle = None

    def dup(self):
        return Vector(self.ctx, lib().vkhel_vector_dup(self.handle))

    @property
    def device_ptr(self):
        """raw device address (vkhel_vector_device_ptr): work recorded so far
        is launched first; the caller orders its own use against the
        context's stream (Context.sync)"""
        return int(lib().vkhel_vector_device_ptr(self.handle) or 0)

    @property
    def __cuda_array_interface__(self):
        """zero-copy view for libraries that speak the CUDA array interface
        (e.g. torch.as_tensor(vec, device="cuda") for an NCCL gather); signed
        64-bit, since torch has no unsigned 64-bit arithmetic -- the bits are
        what is moved"""
        return {"shape": (self.length,), "typestr": "<i8",
                "data": (self.device_ptr, False), "version": 2}

    def copy_from_host(self, data):
        data = _as_u64(data)
        assert data.size >= self.length
        lib().vkhel_vector_copy_from_host(
            self.handle, data.ctypes.data_as(_p64))

    def to_host(self):
        """map + copy + unmap, the reference's way to read a vector"""
        mem = _vp()
        n = self.length
        lib().vkhel_vector_map(self.handle, ctypes.byref(mem), n * 8)
        if n:
            buf = (ctypes.c_uint64 * n).from_address(mem.value)
            out = np.frombuffer(buf, dtype=np.uint64).copy()
        else:
            out = np.empty(0, np.uint64)
        lib().vkhel_vector_unmap(self.handle)
        return out

    def map_range(self, offset, count):
        """vkhel_vector_map_range: numpy view of the staged elements
        [offset, offset + count); write to it, then unmap()"""
        mem = _vp()
        lib().vkhel_vector_map_range(self.handle, ctypes.byref(mem), offset,
                                     count)
        if not count:
            return np.empty(0, np.uint64)
        buf = (ctypes.c_uint64 * count).from_address(mem.value)
        return np.frombuffer(buf, dtype=np.uint64)

    def unmap(self):
        lib().vkhel_vector_unmap(self.handle)

    def read_into(self, out):
        """map, copy the vector into `out` (a numpy array), unmap: the
        reference's way to read a result, with one host copy"""
        np.copyto(out, self.map_range(0, self.length))
        self.unmap()

    def copy_peer(self, src, dst_offset=0, src_offset=0, count=None):
        """device -> device copy, possibly across GPUs (NVLink P2P)"""
        count = src.length - src_offset if count is None else count
        lib().vkhel_vector_copy_peer(self.handle, dst_offset, src.handle,
                                     src_offset, count)

    def upload(self, pinned, offset=0, count=None, host_offset=0):
        """enqueue host -> device: pinned[host_offset:+count] -> self[offset:]"""
        count = pinned.count - host_offset if count is None else count
        lib().vkhel_vector_upload(self.handle, pinned.ptr + 8 * host_offset,
                                  offset, count)

    def download(self, pinned, offset=0, count=None, host_offset=0):
        """enqueue device -> host: self[offset:+count] -> pinned[host_offset:]"""
        count = pinned.count - host_offset if count is None else count
        lib().vkhel_vector_download(self.handle, pinned.ptr + 8 * host_offset,
                                    offset, count)


class Context:
    """struct vkhel_ctx; methods are the vkhel_vector_* operators"""

    def __init__(self, device=None):
        if device is None:
            self.handle = lib().vkhel_ctx_create()
        else:
            self.handle = lib().vkhel_ctx_create_device(device)

    def destroy(self):
        if self.handle:
            lib().vkhel_ctx_destroy(self.handle)
            self.handle = None

    def sync(self):
        lib().vkhel_ctx_sync(self.handle)

    @property
    def stream(self):
        """the context's cudaStream_t (vkhel_ctx_stream).  Handing it out
        launches what is recorded and turns recording off for this context:
        the caller now orders its own work by stream position."""
        return int(lib().vkhel_ctx_stream(self.handle) or 0)

    @property
    def device(self):
        return int(lib().vkhel_ctx_device(self.handle))

    def flush_l2(self):
        lib().vkhel_ctx_flush_l2(self.handle)

    def probe_int_peaks(self):
        """integer-pipe rates of this context's device, measured now
        (vkhel_ctx_probe_int_peaks): thread-instructions, respectively
        butterflies, per clock per SM"""
        buf = (ctypes.c_double * 12)()
        n = lib().vkhel_ctx_probe_int_peaks(self.handle, buf, 12)
        names = ["sm_count", "sm_clock_mhz", "imad", "imad_wide", "imad_hi",
                 "lop3", "bfly_forward", "bfly_inverse", "shf", "iadd3",
                 "add64_3in", "csub64"]
        return {k: float(buf[i]) for i, k in enumerate(names[:n])}

    def flush(self):
        """launch the recorded single-vector transforms (no wait)"""
        lib().vkhel_ctx_flush(self.handle)

    @property
    def deferred_stats(self):
        """(batched launches, transforms carried by them) so far"""
        b, t = ctypes.c_uint64(0), ctypes.c_uint64(0)
        lib().vkhel_ctx_deferred_stats(self.handle, ctypes.byref(b),
                                       ctypes.byref(t))
        return int(b.value), int(t.value)

    @property
    def fused_products(self):
        """elemmul calls that were fused into the inverse transform after them"""
        return int(lib().vkhel_ctx_fused_products(self.handle))

    @property
    def lazy_forwards(self):
        """batched forward transforms that stored lazy residues for the
        in-place inverse transform that followed them at once"""
        return int(lib().vkhel_ctx_lazy_forwards(self.handle))

    @property
    def readahead_hits(self):
        """maps served from a device -> host copy started before the call"""
        return int(lib().vkhel_ctx_readahead_hits(self.handle))

    @property
    def launch_count(self):
        return int(lib().vkhel_ctx_launch_count(self.handle))

    @property
    def launch_count_noflush(self):
        """kernels launched so far; what is recorded stays recorded"""
        return int(lib().vkhel_ctx_launch_count_noflush(self.handle))

    def vector(self, length, zero=True):
        return Vector(self, lib().vkhel_vector_create2(self.handle, length,
                                                       zero))

    def from_host(self, data):
        data = _as_u64(data)
        v = self.vector(data.size, zero=False)
        v.copy_from_host(data)
        return v

    # element-wise
    def elemfma(self, a, b, result, multiplier, mod):
        lib().vkhel_vector_elemfma(a.handle, b.handle, result.handle,
                                   multiplier, mod)

    def elemmod(self, a, result, mod, q):
        lib().vkhel_vector_elemmod(a.handle, result.handle, mod, q)

    def elemmul(self, a, b, result, mod):
        lib().vkhel_vector_elemmul(a.handle, b.handle, result.handle, mod)

    def elemgtadd(self, a, result, bound, diff):
        lib().vkhel_vector_elemgtadd(a.handle, result.handle, bound, diff)

    def elemgtsub(self, a, result, bound, diff, mod):
        lib().vkhel_vector_elemgtsub(a.handle, result.handle, bound, diff, mod)

    def elemmul_rns(self, a, b, result, mods, n, batch):
        mods = _as_u64(mods)
        lib().vkhel_vector_elemmul_rns(a.handle, b.handle, result.handle,
                                       mods.ctypes.data_as(_p64), mods.size,
                                       n, batch)

    # transforms
    def forward_transform(self, operand, result, ntt):
        lib().vkhel_vector_forward_transform(operand.handle, result.handle,
                                             ntt.handle)

    def inverse_transform(self, operand, result, ntt):
        lib().vkhel_vector_inverse_transform(operand.handle, result.handle,
                                             ntt.handle)

    def forward_transform_batch(self, operand, result, ntt, batch):
        lib().vkhel_vector_forward_transform_batch(
            operand.handle, result.handle, ntt.handle, batch)

    def inverse_transform_batch(self, operand, result, ntt, batch):
        lib().vkhel_vector_inverse_transform_batch(
            operand.handle, result.handle, ntt.handle, batch)

    def forward_transform_rns(self, operand, result, tables, batch):
        lib().vkhel_vector_forward_transform_rns(
            operand.handle, result.handle, _table_array(tables), len(tables),
            batch)

    def inverse_transform_rns(self, operand, result, tables, batch):
        lib().vkhel_vector_inverse_transform_rns(
            operand.handle, result.handle, _table_array(tables), len(tables),
            batch)

    def polymul_rns(self, a, b, result, tables, batch):
        lib().vkhel_vector_polymul_rns(
            a.handle, b.handle, result.handle, _table_array(tables),
            len(tables), batch)

    # timing
    def timer(self):
        return Timer(self)


class Timer:
    def __init__(self, ctx):
        self.handle = lib().vkhel_timer_create(ctx.handle)

    def start(self):
        lib().vkhel_timer_start(self.handle)

    def stop(self):
        lib().vkhel_timer_stop(self.handle)

    def elapsed_ms(self):
        return float(lib().vkhel_timer_elapsed_ms(self.handle))

    def destroy(self):
        if self.handle:
            lib().vkhel_timer_destroy(self.handle)
            self.handle = None
