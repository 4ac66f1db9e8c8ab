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

    def border_groups(self, big_u, frac):
        """grouping of the big fronts for the deferred border x border updates (pipsgpu_symbolic_border_groups)"""
        cnt = np.zeros(3, dtype=np.int64)
        null = C.c_void_p()
        _check(lib().pipsgpu_symbolic_border_groups(self.h, C.c_int(big_u), C.c_double(frac), _ip(cnt), null, null, null,
                                                    null, null, null))
        ng, nm, npos = (int(v) for v in cnt)
        g = dict(root=np.zeros(ng, np.int32), N=np.zeros(ng, np.int32), mem_ptr=np.zeros(ng + 1, np.int32),
                 mem_front=np.zeros(nm, np.int32), pos_ptr=np.zeros(nm + 1, np.int64), pos=np.zeros(npos, np.int32))
        _check(lib().pipsgpu_symbolic_border_groups(self.h, C.c_int(big_u), C.c_double(frac), _ip(cnt), _ip(g["root"]),
                                                    _ip(g["N"]), _ip(g["mem_ptr"]), _ip(g["mem_front"]), _ip(g["pos_ptr"]),
                                                    _ip(g["pos"])))
        return g

    def __del__(self):
        try:
            if self.h:
                lib().pipsgpu_symbolic_destroy(self.h)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------
# device objects
# ---------------------------------------------------------------------------------------------------------------
class Context:
    """pipsgpu_ctx: one per process / GPU."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().pipsgpu_ctx_create(C.c_int(device), C.byref(self.h)))
        self.device = device
        self._keep = []

    def init_comm(self, uid: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(uid, 128)
        _check(lib().pipsgpu_ctx_init_comm(self.h, buf, C.c_int(rank), C.c_int(nranks)))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().pipsgpu_nccl_unique_id(buf))
        return buf.raw

    def set_stream(self, cuda_stream: int):
        """borrow a CUDA stream (raw cudaStream_t as int, e.g. torch.cuda.Stream().cuda_stream)"""
        _check(lib().pipsgpu_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def commit(self, threads=0):
        _check(lib().pipsgpu_ctx_commit(self.h, C.c_int(threads)))

    def synchronize(self):
        _check(lib().pipsgpu_ctx_synchronize(self.h))

    def factorize_leaves(self, sc=None):
        _check(lib().pipsgpu_ctx_factorize_leaves(self.h, sc.h if sc is not None else None))

    def dmma_peak_tflops(self):
        """measured FP64 tensor-pipe peak of this GPU (register-resident DMMA chains)"""
        v = C.c_double(0)
        _check(lib().pipsgpu_bench_dmma_peak(self.h, C.byref(v)))
        return v.value

    def dgemm_tflops(self, n=4096):
        """measured rate of the library's own FP64 tensor-pipe GEMM kernel (TMA-staged operands) on an n^3 lower update"""
        v = C.c_double(0)
        _check(lib().pipsgpu_bench_dgemm(self.h, C.c_int(n), C.byref(v)))
        return v.value

    def launch_count(self):
        return int(lib().pipsgpu_ctx_launch_count(self.h))

    def pool_sizes(self):
        out = np.zeros(8, dtype=np.int64)
        _check(lib().pipsgpu_ctx_pool_sizes(self.h, _ip(out)))
        keys = ("L", "C", "solve_pool", "fronts", "rows", "levels", "nnzK", "nnzB")
        return {k: int(v) for k, v in zip(keys, out)}

    def n_batches(self):
        return int(lib().pipsgpu_ctx_num_batches(self.h))

    def close(self):
        if self.h:
            _check(lib().pipsgpu_ctx_destroy(self.h))
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Leaf:
    """pipsgpu_leaf: one scenario block (DistributedLeafLinearSystem + its DoubleLinearSolver)."""

    def __init__(self, ctx: Context, n, nx, ny, nz, krow, jcol, val):
        self.ctx = ctx
        self.n, self.nx, self.ny, self.nz = n, nx, ny, nz
        self.h = C.c_void_p()
        krow, jcol, val = _i32(krow), _i32(jcol), _f64(val)
        _check(lib().pipsgpu_leaf_create(ctx.h, C.c_int(n), C.c_int(nx), C.c_int(ny), C.c_int(nz), _ip(krow), _ip(jcol),
                                         _ip(val), C.byref(self.h)))

    def set_border(self, n0, my0, myl, mzl, A, Cm, F, G):
        args = []
        for m in (A, Cm, F, G):
            args.extend(_csr_args(m))
        _check(lib().pipsgpu_leaf_set_border(self.h, C.c_int(n0), C.c_int(my0), C.c_int(myl), C.c_int(mzl),
                                             *[_ip(a) for a in args]))

    def analyze(self):
        _check(lib().pipsgpu_leaf_analyze(self.h))

    def set_diagonal(self, dd, nomegainv, regP=0.0, regDy=0.0, regDz=0.0):
        dd, nomegainv = _f64(dd), _f64(nomegainv)
        _check(lib().pipsgpu_leaf_set_diagonal(self.h, _ip(dd), _ip(nomegainv), C.c_double(regP), C.c_double(regDy),
                                               C.c_double(regDz)))

    def set_regularization(self, regP=0.0, regDy=0.0, regDz=0.0):
        _check(lib().pipsgpu_leaf_set_regularization(self.h, C.c_double(regP), C.c_double(regDy), C.c_double(regDz)))

    def set_values(self, val):
        val = _f64(val)
        _check(lib().pipsgpu_leaf_set_values(self.h, _ip(val)))

    def factorize(self):
        inertia = np.zeros(3, dtype=np.int32)
        _check(lib().pipsgpu_leaf_factorize(self.h, _ip(inertia)))
        return tuple(int(v) for v in inertia)

    def solve(self, rhs):
        """rhs: (nrhs, n) or (n,) array; returns the solution (same shape)."""
        x = np.array(rhs, dtype=np.float64, order="C", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[0]
        _check(lib().pipsgpu_leaf_solve(self.h, C.c_int(nrhs), _ip(x)))
        return x

    def inertia(self):
        inertia = np.zeros(3, dtype=np.int32)
        npert = C.c_int(0)
        _check(lib().pipsgpu_leaf_get_inertia(self.h, _ip(inertia), C.byref(npert)))
        return tuple(int(v) for v in inertia), npert.value

    def stats(self):
        out = np.zeros(8, dtype=np.int64)
        fl = C.c_double(0)
        _check(lib().pipsgpu_leaf_get_stats(self.h, _ip(out), C.byref(fl)))
        keys = ("nnzL", "nnzL_border", "stored", "nfronts", "nlevels", "n_pairs", "max_m", "contrib")
        d = {k: int(v) for k, v in zip(keys, out)}
        d["flops"] = fl.value
        f4 = np.zeros(4)
        if lib().pipsgpu_leaf_get_flops(self.h, _ip(f4)) == 0:
            d["schur_flops"], d["update_flops"], d["panel_flops"] = float(f4[1]), float(f4[2]), float(f4[3])
        return d

    def download_factor(self):
        """(L panel pool, dinv_d, dinv_e) of the last factorisation - same layout as oracle/mf_ldlt.c"""
        L = np.zeros(max(self.stats()["stored"], 1))
        dd = np.zeros(self.n)
        de = np.zeros(self.n)
        _check(lib().pipsgpu_leaf_download_factor(self.h, _ip(L), _ip(dd), _ip(de)))
        return L, dd, de

    def add_schur(self, sc):
        _check(lib().pipsgpu_leaf_add_schur(self.h, sc.h))


class SchurComplement:
    """pipsgpu_sc: dense root Schur complement + dense LDL^T solver."""

    def __init__(self, ctx: Context, n0, my0, myl, mzl, handle=None):
        self.ctx = ctx
        self.owned = handle is None
        self.h = C.c_void_p(handle) if handle is not None else C.c_void_p()
        if handle is None:
            _check(lib().pipsgpu_sc_create(ctx.h, C.c_int(n0), C.c_int(my0), C.c_int(myl), C.c_int(mzl), C.byref(self.h)))
        self.n = int(lib().pipsgpu_sc_dim(self.h))

    def zero(self):
        _check(lib().pipsgpu_sc_zero(self.h))

    def allreduce(self):
        _check(lib().pipsgpu_sc_allreduce(self.h))

    def add_sparse(self, row0, col0, M, alpha=1.0):
        kr, jc, v = _csr_args(M)
        if kr is None:
            return
        _check(lib().pipsgpu_sc_add_sparse(self.h, C.c_int(row0), C.c_int(col0), C.c_int(M.shape[0]), _ip(kr), _ip(jc),
                                           _ip(v), C.c_double(alpha)))

    def add_diag(self, offset, d, alpha=1.0):
        d = _f64(d)
        _check(lib().pipsgpu_sc_add_diag(self.h, C.c_int(offset), C.c_int(len(d)), _ip(d), C.c_double(alpha)))

    def factorize(self):
        inertia = np.zeros(3, dtype=np.int32)
        _check(lib().pipsgpu_sc_factorize(self.h, _ip(inertia)))
        return tuple(int(v) for v in inertia)

    def solve(self, rhs):
        x = np.array(rhs, dtype=np.float64, order="C", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[0]
        _check(lib().pipsgpu_sc_solve(self.h, C.c_int(nrhs), _ip(x)))
        return x

    def set_sparse_pattern(self, krow, jcol, threads=0):
        """lower CSR pattern of the Schur complement: sparse all-reduce + sparse root LDL^T (pipsgpu_sc_set_sparse_pattern)"""
        krow, jcol = _i32(krow), _i32(jcol)
        assert len(krow) == self.n + 1
        _check(lib().pipsgpu_sc_set_sparse_pattern(self.h, _ip(krow), _ip(jcol), C.c_int(threads)))

    def uses_sparse_solver(self):
        return bool(lib().pipsgpu_sc_uses_sparse_solver(self.h))

    def download(self):
        out = np.zeros((self.n, self.n))
        _check(lib().pipsgpu_sc_download(self.h, _ip(out)))
        return out

    def upload(self, lower):
        a = _f64(lower)
        assert a.shape == (self.n, self.n)
        _check(lib().pipsgpu_sc_upload(self.h, _ip(a)))

    def close(self):
        if self.owned and self.h:
            _check(lib().pipsgpu_sc_destroy(self.h))
            self.h = C.c_void_p()


class KKT:
    """pipsgpu_kkt: sLinsysRootAug for the two-level dense-Schur-complement case."""

    def __init__(self, ctx: Context, n0, my0, mz0, myl, mzl, A0, C0, F0, G0):
        self.ctx = ctx
        self.dims = (n0, my0, mz0, myl, mzl)
        self.h = C.c_void_p()
        args = []
        for m in (A0, C0, F0, G0):
            args.extend(_csr_args(m))
        _check(lib().pipsgpu_kkt_create(ctx.h, C.c_int(n0), C.c_int(my0), C.c_int(mz0), C.c_int(myl), C.c_int(mzl),
                                        *[_ip(a) for a in args], C.byref(self.h)))
        self.leaves = []

    def add_leaf(self, leaf: Leaf):
        _check(lib().pipsgpu_kkt_add_leaf(self.h, leaf.h))
        self.leaves.append(leaf)

    def setup(self, threads=0):
        _check(lib().pipsgpu_kkt_setup(self.h, C.c_int(threads)))

    def factorize(self, dd0, nomegainv0, nomegainv_link, dd_leaves, nomegainv_leaves, reg=None):
        """all arguments host arrays (or raw integer addresses of pinned buffers); returns the root inertia"""
        inertia = np.zeros(3, dtype=np.int32)
        regp = _ip(_f64(reg)) if reg is not None else None
        keep = [_f64(a) if not isinstance(a, int) else a for a in (dd0, nomegainv0, nomegainv_link, dd_leaves, nomegainv_leaves)]
        ptrs = [C.c_void_p(a) if isinstance(a, int) else _ip(a) for a in keep]
        _check(lib().pipsgpu_kkt_factorize(self.h, *ptrs, regp, _ip(inertia)))
        return tuple(int(v) for v in inertia)

    def solve_compressed(self, b0, b_leaves):
        """in place on float64 C-contiguous arrays (or raw addresses)"""
        p0 = C.c_void_p(b0) if isinstance(b0, int) else _ip(b0)
        pl = C.c_void_p(b_leaves) if isinstance(b_leaves, int) else _ip(b_leaves)
        _check(lib().pipsgpu_kkt_solve_compressed(self.h, p0, pl))

    def solve_bicgstab(self, b0, b_leaves, tol=1e-10, max_iter=75):
        """device BiCGStab preconditioned by solveCompressed, in place; returns (status, iterations, relative residual)"""
        p0 = C.c_void_p(b0) if isinstance(b0, int) else _ip(b0)
        pl = C.c_void_p(b_leaves) if isinstance(b_leaves, int) else _ip(b_leaves)
        st, it, rr = C.c_int(0), C.c_int(0), C.c_double(0)
        _check(lib().pipsgpu_kkt_solve(self.h, p0, pl, C.c_double(tol), C.c_int(max_iter), C.byref(st), C.byref(it), C.byref(rr)))
        return st.value, it.value, rr.value

    def factorize_resident(self):
        inertia = np.zeros(3, dtype=np.int32)
        _check(lib().pipsgpu_kkt_factorize_resident(self.h, _ip(inertia)))
        return tuple(int(v) for v in inertia)

    def solve_compressed_resident(self):
        _check(lib().pipsgpu_kkt_solve_compressed_resident(self.h))

    def sc(self):
        n0, my0, mz0, myl, mzl = self.dims
        return SchurComplement(self.ctx, n0, my0, myl, mzl, handle=lib().pipsgpu_kkt_sc(self.h))

    def stage_times(self):
        out = np.zeros(8)
        _check(lib().pipsgpu_kkt_stage_times(self.h, _ip(out)))
        keys = ("h2d_assembly", "leaf_ldlt_schur", "allreduce", "finalize", "root_ldlt", "lsolve", "dsolve", "ltsolve")
        return dict(zip(keys, out.tolist()))

    def kernel_times(self):
        out = np.zeros(4)
        _check(lib().pipsgpu_kkt_kernel_times(self.h, _ip(out)))
        # pack_solve_span: event span around the pack kernel of the last batch on the main stream - the group SYRKs run on their
        # own stream at the same time, so this is NOT the kernel's duration (1.3 ms at C2, 3.3 ms at C3 in the launch lists)
        return dict(zip(("psolve", "pack_solve_span", "dense_solve", "schur_kernels"), out.tolist()))

    def close(self):
        if self.h:
            _check(lib().pipsgpu_kkt_destroy(self.h))
            self.h = C.c_void_p()
