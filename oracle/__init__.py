"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

ctypes view of oracle/liboracle.so (oracle.c, a plain-C restatement of the
reference's algorithm) and, when present, of oracle/_ref/libvkhel_refhost.so
(the reference's own src/numbers.c + src/ntt_tables.c compiled from
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this module; vkhel_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libvkhel_refhost.so")

_u64 = ctypes.c_uint64
_p64 = ctypes.POINTER(ctypes.c_uint64)


def build():
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _load():
    if not os.path.exists(_LIB_PATH):
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    sig = {
        "oracle_shoup_factor": (_u64, [_u64, _u64]),
        "oracle_multiply_mod": (_u64, [_u64, _u64, _u64]),
        "oracle_power_mod": (_u64, [_u64, _u64, _u64]),
        "oracle_inverse_mod": (_u64, [_u64, _u64]),
        "oracle_tables": (None, [_u64, _u64, _u64, _p64, _p64, _p64, _p64]),
        "oracle_forward": (None, [_p64, _p64, _u64, _u64, _p64, _p64]),
        "oracle_inverse": (None, [_p64, _p64, _u64, _u64, _u64, _p64, _p64]),
        "oracle_barrett_defined": (ctypes.c_int, [_u64]),
        "oracle_elemmul_literal": (None, [_p64, _p64, _p64, _u64, _u64]),
        "oracle_elemmul": (None, [_p64, _p64, _p64, _u64, _u64]),
        "oracle_elemfma_literal": (None, [_p64, _p64, _p64, _u64, _u64, _u64]),
        "oracle_elemfma": (None, [_p64, _p64, _p64, _u64, _u64, _u64]),
        "oracle_elemmulconst": (None, [_p64, _p64, _u64, _u64, _u64]),
        "oracle_elemgtadd": (None, [_p64, _p64, _u64, _u64, _u64]),
        "oracle_elemgtsub": (None, [_p64, _p64, _u64, _u64, _u64, _u64,
                                    ctypes.c_int]),
        "oracle_elemmodbytwo": (None, [_p64, _p64, _u64, _u64]),
        "oracle_elemmod": (None, [_p64, _p64, _u64, _u64, _u64]),
        "oracle_negacyclic_schoolbook": (None, [_p64, _p64, _p64, _u64, _u64]),
        "oracle_forward_batch": (None, [_p64, _p64, _u64, _u64, _u64, _p64,
                                        ctypes.POINTER(_p64),
                                        ctypes.POINTER(_p64), ctypes.c_int]),
        "oracle_inverse_batch": (None, [_p64, _p64, _u64, _u64, _u64, _p64,
                                        ctypes.POINTER(_p64),
                                        ctypes.POINTER(_p64), ctypes.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = _load()


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(_p64)


# ---- scalars ------------------------------------------------------------------
def shoup_factor(f, q):
    return int(_lib.oracle_shoup_factor(f, q))


def multiply_mod(a, b, q):
    return int(_lib.oracle_multiply_mod(a, b, q))


def power_mod(base, exp, q):
    return int(_lib.oracle_power_mod(base, exp, q))


def inverse_mod(a, q):
    return int(_lib.oracle_inverse_mod(a, q))


def barrett_defined(q):
    return bool(_lib.oracle_barrett_defined(q))


# ---- tables --------------------------------------------------------------------
class Tables:
    """roots / inv_roots / Shoup companions, reference src/ntt_tables.c:17-44"""

    def __init__(self, n, q, w):
        self.n, self.q, self.w = n, q, w
        self.roots = np.empty(n, np.uint64)
        self.inv_roots = np.empty(n, np.uint64)
        self.roots_shoup = np.empty(n, np.uint64)
        self.inv_roots_shoup = np.empty(n, np.uint64)
        if q >> 62:
            # The reference's host Barrett (numbers.c:5-28) shifts the high
            # product word by 66 - bits(q) and overflows for 63-bit moduli:
            # its tables are not well defined there.  For such moduli the
            # oracle states the table CONTRACT with Python
            # integers instead of restating the broken arithmetic.
            self._fill_exact()
        else:
            _lib.oracle_tables(n, q, w, _ptr(self.roots),
                               _ptr(self.inv_roots), _ptr(self.roots_shoup),
                               _ptr(self.inv_roots_shoup))

    def _fill_exact(self):
        n, q, w = self.n, self.q, self.w
        bits = n.bit_length() - 1
        w_inv = pow(w, -1, q) if n > 1 else 1
        p = pi = 1
        for i in range(n):
            idx = int(format(i, "0%db" % bits)[::-1], 2) if bits else 0
            self.roots[idx] = p
            self.inv_roots[idx] = pi
            self.roots_shoup[idx] = (p << 64) // q
            self.inv_roots_shoup[idx] = (pi << 64) // q
            p = p * w % q
            pi = pi * w_inv % q


# ---- transforms -----------------------------------------------------------------
def forward(x, t):
    x = _arr(x)
    out = np.empty(t.n, np.uint64)
    _lib.oracle_forward(_ptr(x), _ptr(out), t.n, t.q, _ptr(t.roots),
                        _ptr(t.roots_shoup))
    return out


def inverse(x, t, out_len=None, out_init=None):
    """out_len > n reproduces the reference scaling every element of the
    result vector (src/vector.c:635-638); out_init seeds the tail."""
    x = _arr(x)
    out_len = t.n if out_len is None else out_len
    out = np.zeros(max(out_len, t.n), np.uint64)
    if out_init is not None:
        out[:len(out_init)] = out_init
    _lib.oracle_inverse(_ptr(x), _ptr(out), t.n, out_len, t.q,
                        _ptr(t.inv_roots), _ptr(t.inv_roots_shoup))
    return out


def _table_ptrs(tables, names):
    out = []
    for name in names:
        arr = (_p64 * len(tables))(*[_ptr(getattr(t, name)) for t in tables])
        out.append(arr)
    return out


def forward_batch(x, tables, threads=1):
    """x: [polys][n]; polynomial p uses tables[p % len(tables)]"""
    x = _arr(x)
    n = tables[0].n
    polys = x.size // n
    out = np.empty_like(x)
    mods = _arr([t.q for t in tables])
    r, rs = _table_ptrs(tables, ["roots", "roots_shoup"])
    _lib.oracle_forward_batch(_ptr(x), _ptr(out), n, polys, len(tables),
                              _ptr(mods), r, rs, threads)
    return out


def inverse_batch(x, tables, threads=1):
    x = _arr(x)
    n = tables[0].n
    polys = x.size // n
    out = np.empty_like(x)
    mods = _arr([t.q for t in tables])
    r, rs = _table_ptrs(tables, ["inv_roots", "inv_roots_shoup"])
    _lib.oracle_inverse_batch(_ptr(x), _ptr(out), n, polys, len(tables),
                              _ptr(mods), r, rs, threads)
    return out


# ---- element-wise ------------------------------------------------------------------
def _binary(fn, a, b, *scalars):
    a, b = _arr(a), _arr(b)
    out = np.empty_like(a)
    fn(_ptr(a), _ptr(b), _ptr(out), a.size, *scalars)
    return out


def _unary(fn, a, *scalars):
    a = _arr(a)
    out = np.empty_like(a)
    fn(_ptr(a), _ptr(out), a.size, *scalars)
    return out


def elemmul(a, b, q, literal=False):
    fn = _lib.oracle_elemmul_literal if literal else _lib.oracle_elemmul
    return _binary(fn, a, b, q)


def elemfma(a, b, mult, q, literal=False):
    fn = _lib.oracle_elemfma_literal if literal else _lib.oracle_elemfma
    return _binary(fn, a, b, mult, q)


def elemmulconst(a, b, q):
    return _unary(_lib.oracle_elemmulconst, a, b, q)


def elemgtadd(a, bound, diff):
    return _unary(_lib.oracle_elemgtadd, a, bound, diff)


def elemgtsub(a, bound, diff, q, literal=False):
    return _unary(_lib.oracle_elemgtsub, a, bound, diff, q, int(literal))


def elemmodbytwo(a, signed_bound):
    return _unary(_lib.oracle_elemmodbytwo, a, signed_bound)


def elemmod(a, mod, q):
    return _unary(_lib.oracle_elemmod, a, mod, q)


def negacyclic_schoolbook(a, b, q):
    a, b = _arr(a), _arr(b)
    out = np.empty_like(a)
    _lib.oracle_negacyclic_schoolbook(_ptr(a), _ptr(b), _ptr(out), a.size, q)
    return out


# ---- the reference's own host code (oracle/_ref), when built ------------------------
class _RefTablesStruct(ctypes.Structure):
    # reference include/priv/ntt_tables.h:6-15
    _fields_ = [("n", _u64), ("q", _u64), ("w", _u64),
                ("roots_of_unity", _p64), ("inv_roots_of_unity", _p64),
                ("roots_barrett_factors", _p64),
                ("inv_roots_barrett_factors", _p64)]


def reference_host():
    """ctypes handle on the reference's compiled numbers.c/ntt_tables.c, or
    None when oracle/_ref has not been built (no /root/reference)."""
    if not os.path.exists(_REF_PATH):
        return None
    ref = ctypes.CDLL(_REF_PATH, mode=ctypes.RTLD_LOCAL)
    ref.vkhel_ntt_tables_create.restype = ctypes.POINTER(_RefTablesStruct)
    ref.vkhel_ntt_tables_create.argtypes = [_u64, _u64, _u64]
    ref.vkhel_ntt_tables_destroy.argtypes = [ctypes.POINTER(_RefTablesStruct)]
    ref.nt_multiply_mod.restype = _u64
    ref.nt_multiply_mod.argtypes = [_u64, _u64, _u64, _u64]
    ref.nt_power_mod.restype = _u64
    ref.nt_power_mod.argtypes = [_u64, _u64, _u64]
    ref.nt_inverse_mod.restype = _u64
    ref.nt_inverse_mod.argtypes = [_u64, _u64]
    ref.nt_compute_barrett_factor.restype = _u64
    ref.nt_compute_barrett_factor.argtypes = [_u64, _u64, _u64]
    ref.nt_is_primitive_root.restype = ctypes.c_bool
    ref.nt_is_primitive_root.argtypes = [_u64, _u64, _u64]
    return ref


def reference_tables(ref, n, q, w):
    """Tables produced by the reference's vkhel_ntt_tables_create, as numpy
    arrays (roots, inv_roots, roots_shoup, inv_roots_shoup)."""
    t = ref.vkhel_ntt_tables_create(n, q, w)
    c = t.contents
    out = tuple(np.ctypeslib.as_array(getattr(c, f), shape=(n,)).copy()
                for f in ("roots_of_unity", "inv_roots_of_unity",
                          "roots_barrett_factors",
                          "inv_roots_barrett_factors"))
    ref.vkhel_ntt_tables_destroy(t)
    return out
