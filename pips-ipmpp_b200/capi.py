"""ctypes binding of libpipsgpu.so (include/pipsgpu.h). Thin: one Python method per C entry point, numpy arrays in,
numpy arrays out, errors raised as PipsGpuError with the library's message. No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpipsgpu.so")
_lib = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


class PipsGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pipsgpu error {code}: {msg}")
        self.code = code


def lib():
    """Load the shared library; fails loudly when it has not been built (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PipsGpuError(-1, f"{LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.pipsgpu_last_error.restype = C.c_char_p
        for name, rt in (("pipsgpu_ctx_launch_count", C.c_int64), ("pipsgpu_kkt_sc", C.c_void_p)):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = rt
    return _lib


def _check(rc):
    if rc != 0:
        raise PipsGpuError(rc, lib().pipsgpu_last_error().decode())


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _csr_args(m):
    """(krow, jcol, val) int32/int32/float64 arrays of a scipy CSR matrix or (None, None, None)"""
    if m is None or m.shape[0] == 0 or m.shape[1] == 0:
        return None, None, None
    m = m.tocsr()
    return _i32(m.indptr), _i32(m.indices), _f64(m.data)


def partition(n_children, n_ranks):
    lo = np.zeros(n_ranks, dtype=np.int32)
    hi = np.zeros(n_ranks, dtype=np.int32)
    _check(lib().pipsgpu_partition(C.c_int(n_children), C.c_int(n_ranks), _ip(lo), _ip(hi)))
    return [(int(a), int(b)) for a, b in zip(lo, hi)]


class Symbolic:
    """host-only symbolic analysis (pipsgpu_symbolic_*)"""

    def __init__(self, n, nx, ny, nz, krow, jcol, val, nB, brow, bcol, max_width=0):
        self.h = C.c_void_p()
        krow, jcol, val, brow, bcol = _i32(krow), _i32(jcol), _f64(val), _i32(brow), _i32(bcol)
        _check(lib().pipsgpu_symbolic_create(C.c_int(n), C.c_int(nx), C.c_int(ny), C.c_int(nz), _ip(krow), _ip(jcol),
                                             _ip(val), C.c_int(nB), C.c_int64(len(brow)), _ip(brow), _ip(bcol),
                                             C.c_int(max_width), C.byref(self.h)))
        sz = np.zeros(13, dtype=np.int64)
        fl = C.c_double(0)
        _check(lib().pipsgpu_symbolic_sizes(self.h, _ip(sz), C.byref(fl)))
        keys = ("n", "nB", "ns", "nlevels", "nrows", "Lsize", "Csize", "nnzK", "nnzB", "nnzL", "nnzL_border", "n_pairs",
                "nchild")
        self.sizes = {k: int(v) for k, v in zip(keys, sz)}
        self.flops = fl.value

    def export(self):
        s = self.sizes
        n, ns = s["n"], s["ns"]
        a = dict(perm=np.zeros(n, np.int32), ptype=np.zeros(n, np.int8), sgn=np.zeros(n, np.int8),
                 sn_col=np.zeros(ns + 1, np.int32), sn_rowptr=np.zeros(ns + 1, np.int64),
                 rowidx=np.zeros(s["nrows"], np.int32), relidx=np.zeros(s["nrows"], np.int32),
                 sn_na=np.zeros(ns, np.int32), sn_nb=np.zeros(ns, np.int32), sn_parent=np.zeros(ns, np.int32),
                 sn_level=np.zeros(ns, np.int32), Loff=np.zeros(ns + 1, np.int64), Coff=np.zeros(ns + 1, np.int64),
                 kmap=np.zeros(s["nnzK"], np.int64), bmap=np.zeros(s["nnzB"], np.int64),
                 child_ptr=np.zeros(ns + 1, np.int32), child_idx=np.zeros(s["nchild"], np.int32))
        order = ("perm", "ptype", "sgn", "sn_col", "sn_rowptr", "rowidx", "relidx", "sn_na", "sn_nb", "sn_parent",
                 "sn_level", "Loff", "Coff", "kmap", "bmap", "child_ptr", "child_idx")
        _check(lib().pipsgpu_symbolic_export(self.h, *[_ip(a[k]) for k in order]))
        a.update(n=n, nB=s["nB"], ns=ns)
        return a

    def __del__(self):
        try:
            if self.h:
                lib().pipsgpu_symbolic_destroy(self.h)
        except Exception:
            pass
