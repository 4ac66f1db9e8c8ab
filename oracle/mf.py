"""CPU ORACLE glue (test infrastructure only): ctypes driver of oracle/mf_ldlt.c on a symbolic structure exported by
the product's host-side analysis (pipsgpu_symbolic_export). Used to validate that analysis on CPU and as the
multi-frontal CPU baseline in bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """gcc-compile oracle/mf_ldlt.c into oracle/_build/libmf_oracle.so"""
    out = os.path.join(_HERE, "_build", "libmf_oracle.so")
    src = os.path.join(_HERE, "mf_ldlt.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


class _MfSym(C.Structure):
    _fields_ = [("n", C.c_int), ("nB", C.c_int), ("ns", C.c_int),
                ("sn_col", C.c_void_p), ("sn_rowptr", C.c_void_p), ("rowidx", C.c_void_p), ("relidx", C.c_void_p),
                ("sn_na", C.c_void_p), ("sn_nb", C.c_void_p), ("sn_parent", C.c_void_p), ("Loff", C.c_void_p),
                ("Coff", C.c_void_p), ("ptype", C.c_void_p), ("sgn", C.c_void_p)]


_BLAS = None      # keeps the OpenBLAS handle alive


def _attach_blas(L):
    """hand scipy's bundled OpenBLAS dgemm (LP64 build, scipy.libs/libscipy_openblas-*.so) to the C code, pinned to one
    thread per call (the blocks are spread over host threads by MfBackend). MF_ORACLE_BLAS=0 keeps the scalar loops."""
    global _BLAS
    if os.environ.get("MF_ORACLE_BLAS", "1") == "0":
        return False
    try:
        import glob
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas-*.so"))
        if not cands:
            return False
        B = C.CDLL(cands[0])
        fn = getattr(B, "scipy_cblas_dgemm")
        L.mf_set_dgemm(C.cast(fn, C.c_void_p))
        _BLAS = B
        return True
    except Exception:      # noqa: BLE001
        return False


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.mf_factorize.restype = C.c_int
        _LIB.mf_solve.restype = C.c_int
        _LIB.mf_has_dgemm.restype = C.c_int
        _attach_blas(_LIB)
    return _LIB


def uses_blas():
    return bool(lib().mf_has_dgemm())


def blas_threads(n):
    """thread count of the OpenBLAS instance behind the leaf updates AND scipy's LAPACK: 1 while the leaves are spread over
    the host threads (the dgemm calls must not oversubscribe the cores), all cores again for the root's dsytrf"""
    if _BLAS is not None and hasattr(_BLAS, "scipy_openblas_set_num_threads"):
        _BLAS.scipy_openblas_set_num_threads(C.c_int(max(1, int(n))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_SYMLIB = None


def symlib():
    """host-only build of the product's symbolic analysis (oracle/_build/libpipssym.so, g++ of csrc/symbolic.cpp +
    csrc/capi_host.cpp by __graft_entry__.build_oracle): the CPU arm gets its elimination plan without loading the CUDA
    library into its process"""
    global _SYMLIB
    if _SYMLIB is None:
        path = os.path.join(_HERE, "_build", "libpipssym.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built - run __graft_entry__.build()")
        _SYMLIB = C.CDLL(path)
        _SYMLIB.pipsgpu_last_error.restype = C.c_char_p
    return _SYMLIB


class HostSymbolic:
    """pipsgpu_symbolic_* through libpipssym.so; same interface as pips-ipmpp_b200/capi.py:Symbolic (sizes / flops / export)"""

    def __init__(self, n, nx, ny, nz, krow, jcol, val, nB, brow, bcol, max_width=0):
        L = symlib()
        self.h = C.c_void_p()
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        krow, jcol, brow, bcol = i32(krow), i32(jcol), i32(brow), i32(bcol)
        val = np.ascontiguousarray(val, dtype=np.float64)
        rc = L.pipsgpu_symbolic_create(C.c_int(n), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(krow), _p(jcol), _p(val),
                                       C.c_int(nB), C.c_int64(len(brow)), _p(brow), _p(bcol), C.c_int(max_width),
                                       C.byref(self.h))
        if rc != 0:
            raise RuntimeError(L.pipsgpu_last_error().decode())
        sz = np.zeros(13, dtype=np.int64)
        fl = C.c_double(0)
        L.pipsgpu_symbolic_sizes(self.h, _p(sz), C.byref(fl))
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
        rc = symlib().pipsgpu_symbolic_export(self.h, *[_p(a[k]) for k in order])
        if rc != 0:
            raise RuntimeError(symlib().pipsgpu_last_error().decode())
        a.update(n=n, nB=s["nB"], ns=ns)
        return a

    def __del__(self):
        try:
            if self.h:
                symlib().pipsgpu_symbolic_destroy(self.h)
        except Exception:
            pass


def leaf_kkt_and_border(lp, i):
    """lower CSR of the leaf KKT (create_kkt) and the border triplets in the compressed border index space
    [0,n0) | myl | mzl  (SC index = b for b < n0, b + my0 otherwise)."""
    from .linsys import build_leaf_kkt_lower
    ch = lp.children[i]
    krow, jcol, val, diag_pos = build_leaf_kkt_lower(ch)
    n0 = lp.root.n
    A, Cm, F, G = ch.A.tocoo(), ch.C.tocoo(), ch.F.tocoo(), ch.G.tocoo()
    brow = np.concatenate([ch.n + A.row, ch.n + ch.my + Cm.row, F.col, G.col]).astype(np.int32)
    bcol = np.concatenate([A.col, Cm.col, n0 + F.row, n0 + lp.myl + G.row]).astype(np.int32)
    bval = np.concatenate([A.data, Cm.data, F.data, G.data]).astype(np.float64)
    nB = n0 + lp.myl + lp.mzl
    return krow, jcol, val, diag_pos, nB, brow, bcol, bval


class MfLeaf:
    """numeric multifrontal LDL^T of one leaf on a given symbolic structure `sym` (dict of numpy arrays)."""

    def __init__(self, sym, pivot_eps=1e-8):
        self.sym = sym
        self.S = _MfSym(sym["n"], sym["nB"], sym["ns"], *[_p(sym[k]) for k in (
            "sn_col", "sn_rowptr", "rowidx", "relidx", "sn_na", "sn_nb", "sn_parent", "Loff", "Coff", "ptype", "sgn")])
        n = sym["n"]
        self.L = np.zeros(int(sym["Loff"][-1]) + 1)
        self.Cb = np.zeros(int(sym["Coff"][-1]) + 1)
        self.d_d, self.d_e, self.dinv_d, self.dinv_e = (np.zeros(n + 1) for _ in range(4))
        self.pivot_eps = pivot_eps
        self.inertia = np.zeros(3, dtype=np.int32)
        self.npert = C.c_int(0)

    def factorize(self, kval, bval, SCb=None, knorm=None):
        sym = self.sym
        kval = np.ascontiguousarray(kval, dtype=np.float64)
        bval = np.ascontiguousarray(bval, dtype=np.float64)
        if knorm is None:
            knorm = max(1.0, float(np.abs(kval).max(initial=0.0)))
        rc = lib().mf_factorize(C.byref(self.S), C.c_int64(len(kval)), _p(sym["kmap"]), _p(kval), C.c_int64(len(bval)),
                                _p(sym["bmap"]), _p(bval), _p(self.L), _p(self.Cb), _p(self.d_d), _p(self.d_e),
                                _p(self.dinv_d), _p(self.dinv_e), _p(SCb) if SCb is not None else None,
                                C.c_double(self.pivot_eps * knorm), _p(self.inertia), C.byref(self.npert))
        assert rc == 0
        return tuple(int(v) for v in self.inertia), self.npert.value

    def solve(self, x):
        x = np.array(x, dtype=np.float64, copy=True)
        rc = lib().mf_solve(C.byref(self.S), _p(self.sym["perm"]), _p(self.L), _p(self.dinv_d), _p(self.dinv_e), _p(x))
        assert rc == 0
        return x


class MfBackend:
    """CPU KKT backend (same interface as oracle.ipm.OracleBackend) built from the C multifrontal restatement for the
    leaves (threads over blocks; ctypes releases the GIL) and the oracle's LAPACK dsytrf/dsytrs root. This is the
    CPU path timed as `cpu_baseline` / `--impl reference` in bench.py: it does the same fused factor + Schur work as
    the reference's default PARDISO-Schur configuration (sLinsysLeafSchurSlv.C:20-35)."""

    def __init__(self, lp, symbolic_factory, threads=1, max_width=0):
        from concurrent.futures import ThreadPoolExecutor
        from .linsys import OracleRoot
        self.lp = lp
        self.threads = max(1, threads)
        self.pool = ThreadPoolExecutor(self.threads)
        self.root = OracleRoot(lp, owned=[])
        self.leaf_data = []

        def prep(i):
            ch = lp.children[i]
            krow, jcol, val, diag_pos, nB, brow, bcol, bval = leaf_kkt_and_border(lp, i)
            sym = symbolic_factory(ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val, nB, brow, bcol, max_width)
            arrs = sym.export()
            knorm = max(1.0, float(np.abs(val).max(initial=0.0)))     # off-diagonal norm (diagonal is still zero here)
            return dict(m=MfLeaf(arrs), val=val.copy(), diag_pos=diag_pos, bval=bval, brow=brow, bcol=bcol, nB=nB,
                        knorm=knorm, stats=dict(sym.sizes, flops=sym.flops))

        self.leaf_data = list(self.pool.map(prep, range(len(lp.children))))
        n0, my0, myl, mzl = lp.root.n, lp.root.my, lp.myl, lp.mzl
        self.nB = n0 + myl + mzl
        self.scidx = np.concatenate([np.arange(n0), n0 + my0 + np.arange(myl + mzl)])

    def factorize(self, dd, nomegainv, blocks=None):
        lp = self.lp
        blocks = range(len(lp.children)) if blocks is None else blocks
        nth = self.threads
        acc = [np.zeros((self.nB, self.nB)) for _ in range(nth)]
        chunks = [list(blocks)[t::nth] for t in range(nth)]

        def work(t):
            for i in chunks[t]:
                d = self.leaf_data[i]
                ch = lp.children[i]
                v = d["val"]
                v[d["diag_pos"][:ch.n]] = dd[1 + i]
                v[d["diag_pos"][ch.n:ch.n + ch.my]] = 0.0
                v[d["diag_pos"][ch.n + ch.my:]] = nomegainv[1 + i]
                d["inertia"], d["npert"] = d["m"].factorize(v, d["bval"], acc[t], knorm=d["knorm"])

        import time
        t0 = time.perf_counter()
        blas_threads(1)
        list(self.pool.map(work, range(nth)))
        blas_threads(os.cpu_count() or 1)
        SCb = np.sum(acc, axis=0)
        self.t_leaf_factor = time.perf_counter() - t0      # leaf LDL^T + Schur terms (scales with the number of blocks)
        t0 = time.perf_counter()
        r = self.root
        r.put_diagonals(dd, nomegainv)
        r.initialize_kkt()
        r.SC[np.ix_(self.scidx, self.scidx)] += np.tril(SCb)
        r.finalize_kkt_dense()
        r.factorize_kkt()
        self.t_root_factor = time.perf_counter() - t0      # root assembly + dense LDL^T (does not scale with the blocks)

    def solve_compressed(self, b, blocks=None):
        lp = self.lp
        r = self.root
        blocks = list(range(len(lp.children)) if blocks is None else blocks)
        n0 = lp.root.n
        L0 = len(b[0])
        tail = L0 - lp.myl - lp.mzl

        def b0idx(bc):
            return np.where(bc < n0, bc, bc - n0 + tail)

        def fwd(i):
            d = self.leaf_data[i]
            z = d["m"].solve(b[1 + i])
            b[1 + i][:] = z
            contrib = np.zeros(L0)
            np.add.at(contrib, b0idx(d["bcol"]), d["bval"] * z[d["brow"]])
            return contrib

        import time
        t0 = time.perf_counter()
        for c in self.pool.map(fwd, blocks):
            b[0] -= c
        t1 = time.perf_counter()
        r.solve_reduced_link_cons(b[0])
        t2 = time.perf_counter()

        def bwd(i):
            d = self.leaf_data[i]
            ch = lp.children[i]
            t = np.zeros(ch.n + ch.my + ch.mz)
            np.add.at(t, d["brow"], d["bval"] * b[0][b0idx(d["bcol"])])
            b[1 + i] -= d["m"].solve(t)

        list(self.pool.map(bwd, blocks))
        self.t_leaf_solve = (t1 - t0) + (time.perf_counter() - t2)
        self.t_root_solve = t2 - t1
        return b
