"""Host-side mirror of the reference's linear-system classes for the hot path, over the C ABI (capi.py).

  CudaLeafLinearSystem   <- LinearSystems/DistributedLeafLinearSystem.{h,C}  (create_kkt :52-72, factor2 :74-86, ...)
  CudaRootLinearSystem   <- LinearSystems/sLinsysRootAug.{h,C} + DistributedRootLinearSystem.{h,C}
                            (createChildren :711-758 with DistributedDummyLinearSystem for non-owned blocks,
                             factor2 :206-243, solveCompressed DistributedLinearSystem.C:409-420)

The C++ rendition a PIPS-IPM++ maintainer would compile in is include/pipsgpu_shim.hpp; this
Python rendition is what the tests and bench.py drive. No arithmetic happens here: the methods assemble index/value
arrays once (create_kkt) and afterwards only move diagonals and right-hand sides through the ABI."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import capi


def create_kkt_lower(ch):
    """create_kkt (DistributedLeafLinearSystem.C:52-72): lower CSR of [Q+diag . .; W 0 .; D . 0] with the full diagonal
    stored explicitly (diagonal entry last in each row), Q = 0 for LPs."""
    n, my, mz = ch.n, ch.my, ch.mz
    nK = n + my + mz
    blocks = []
    if my > 0:
        W = ch.W.tocoo()
        blocks.append((W.row + n, W.col, W.data))
    if mz > 0:
        D = ch.D.tocoo()
        blocks.append((D.row + n + my, D.col, D.data))
    rows = np.concatenate([b[0] for b in blocks] + [np.arange(nK)])
    cols = np.concatenate([b[1] for b in blocks] + [np.arange(nK)])
    vals = np.concatenate([b[2] for b in blocks] + [np.zeros(nK)])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    krow = np.zeros(nK + 1, dtype=np.int64)
    np.add.at(krow, rows + 1, 1)
    krow = np.cumsum(krow).astype(np.int32)
    return krow, cols.astype(np.int32), vals.astype(np.float64)


def compressed_kkt_residual(lp, dd, nomegainv, rhs, sol, reg=None, reduce_root=None):
    """Residual  b - K x  of the system solveCompressed solves, evaluated on the host from the
    problem data alone (plain sparse matvecs, no factorisation): per owned block  b_i - K_i x_i - Bord_i x_0  and for the
    root  b_0 - sum_i Bord_i^T x_i - K_root x_0  with  K_root = [dd0 A0^T C0^T F0^T G0^T; A0; C0 . nOm0; F0; G0 . . . nOmL]
    in the b0 layout [x0; y0; z0; yl; zl]. `reduce_root` (a function summing an array over the ranks) makes the root
    part exact when the blocks are distributed. Returns (||r||_2^2 of the owned leaf parts, root residual vector).
    Size-independent correctness property used by bench.py and the full-size parity tests."""
    regP, regDy, regDz = reg if reg is not None else (0.0, 0.0, 0.0)
    root = lp.root
    n0, my0, mz0, myl, mzl = root.n, root.my, root.mz, lp.myl, lp.mzl
    x0 = sol[0]
    u_x, u_y, u_z = x0[:n0], x0[n0:n0 + my0], x0[n0 + my0:n0 + my0 + mz0]
    u_yl, u_zl = x0[n0 + my0 + mz0:n0 + my0 + mz0 + myl], x0[n0 + my0 + mz0 + myl:]
    acc = np.zeros(n0 + my0 + mz0 + myl + mzl)          # sum_i Bord_i^T x_i in the b0 layout
    leaf_sq = 0.0
    for k, ch in enumerate(lp.children):
        xi = sol[1 + k]
        n, my, mz = ch.n, ch.my, ch.mz
        xx, xy, xz = xi[:n], xi[n:n + my], xi[n + my:]
        r = np.array(rhs[1 + k], dtype=np.float64, copy=True)
        r[:n] -= (dd[1 + k] + regP) * xx + ch.W.T @ xy + ch.D.T @ xz + ch.F.T @ u_yl + ch.G.T @ u_zl
        r[n:n + my] -= ch.W @ xx - regDy * xy + ch.A @ u_x
        r[n + my:] -= ch.D @ xx + (nomegainv[1 + k] - regDz) * xz + ch.C @ u_x
        leaf_sq += float(r @ r)
        acc[:n0] += ch.A.T @ xy + ch.C.T @ xz
        acc[n0 + my0 + mz0:n0 + my0 + mz0 + myl] += ch.F @ xx
        acc[n0 + my0 + mz0 + myl:] += ch.G @ xx
    if reduce_root is not None:
        acc = reduce_root(acc)
    r0 = np.array(rhs[0], dtype=np.float64, copy=True) - acc
    r0[:n0] -= (dd[0] + regP) * u_x + root.A.T @ u_y + root.C.T @ u_z + root.F.T @ u_yl + root.G.T @ u_zl
    r0[n0:n0 + my0] -= root.A @ u_x - regDy * u_y
    r0[n0 + my0:n0 + my0 + mz0] -= root.C @ u_x + (nomegainv[0] - regDz) * u_z
    r0[n0 + my0 + mz0:n0 + my0 + mz0 + myl] -= root.F @ u_x - regDy * u_yl
    r0[n0 + my0 + mz0 + myl:] -= root.G @ u_x + (nomegainv[-1] - regDz) * u_zl
    return leaf_sq, r0


def sc_lower_pattern(lp):
    """Structural lower-triangular CSR pattern (krow, jcol) of the root Schur complement in its index order
    [x0; y0; yl; zl] - what the reference precomputes once for its sparse root KKT (LS/sLinsysRootAug.C:79-208, the
    SparseSymmetricMatrix handed to finalizeKKTsparse, :1629-1767). Every block contributes the clique of the linking
    variables / linking rows it touches (its term Bord_i^T K_i^-1 Bord_i), the root its own blocks. `lp` must hold ALL
    blocks (every rank needs the same global pattern)."""
    import scipy.sparse as sp
    root = lp.root
    n0, my0, myl, mzl = root.n, root.my, lp.myl, lp.mzl
    n = n0 + my0 + myl + mzl
    rows, cols = [np.arange(n)], [np.arange(n)]

    def clique(idx):
        idx = np.unique(idx)
        if len(idx):
            r, c = np.meshgrid(idx, idx, indexing="ij")
            m = r >= c
            rows.append(r[m]); cols.append(c[m])

    for ch in lp.children:
        touched = [ch.A.tocoo().col, ch.C.tocoo().col,
                   n0 + my0 + np.unique(ch.F.tocoo().row), n0 + my0 + myl + np.unique(ch.G.tocoo().row)]
        clique(np.concatenate(touched))
    C0 = root.C.tocsr()
    for r in range(root.mz):                      # - C0^T diag C0: columns sharing a row
        clique(C0.indices[C0.indptr[r]:C0.indptr[r + 1]])
    for M, off in ((root.A, n0), (root.F, n0 + my0), (root.G, n0 + my0 + myl)):
        M = M.tocoo()
        rows.append(off + M.row); cols.append(M.col)
    r, c = np.concatenate(rows), np.concatenate(cols)
    P = sp.coo_matrix((np.ones(len(r)), (r, c)), shape=(n, n)).tocsr()
    P.sum_duplicates(); P.sort_indices()
    return P.indptr.astype(np.int32), P.indices.astype(np.int32)


class CudaRootLinearSystem:
    """Root linear system of a two-level block-angular LP on one GPU (one rank of the reference)."""

    def __init__(self, lp, device=0, rank=0, n_ranks=1, nccl_uid=None, threads=0, stream=None, kkts=None, sc_pattern=None):
        self.lp = lp
        root = lp.root
        self.rank, self.n_ranks = rank, n_ranks
        self.ctx = capi.Context(device)
        if stream is not None:
            self.ctx.set_stream(stream)      # borrow the caller's CUDA stream (raw cudaStream_t)
        if n_ranks > 1:
            assert nccl_uid is not None
            self.ctx.init_comm(nccl_uid, rank, n_ranks)
        # `lp.children` holds only the blocks this rank owns (problem.synthetic(block_range=...)); the reference gets
        # the same effect with DistributedDummyLinearSystem children (LS/DistributedRootLinearSystem.C:725-757)
        self.kkt = capi.KKT(self.ctx, root.n, root.my, root.mz, lp.myl, lp.mzl, root.A, root.C, root.F, root.G)
        self.leaves = []
        for k, ch in enumerate(lp.children):
            krow, jcol, val = kkts[k] if kkts is not None else create_kkt_lower(ch)      # create_kkt, possibly precomputed
            leaf = capi.Leaf(self.ctx, ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val)
            leaf.set_border(root.n, root.my, lp.myl, lp.mzl, ch.A, ch.C, ch.F, ch.G)
            self.kkt.add_leaf(leaf)
            self.leaves.append(leaf)
        self.kkt.setup(threads)
        if sc_pattern is not None:          # sparse root KKT (sLinsysRootAug with hasSparseKkt): (krow, jcol) of sc_lower_pattern
            self.kkt.sc().set_sparse_pattern(sc_pattern[0], sc_pattern[1], threads)
        self.nb0 = root.n + root.my + root.mz + lp.myl + lp.mzl
        self.sum_n = sum(ch.n + ch.my + ch.mz for ch in lp.children)
        self._b_leaves = np.zeros(max(self.sum_n, 1))
        self._b0 = np.zeros(self.nb0)

    # ---- backend interface used by the host IPM driver ----------------------------------------------------
    def factorize(self, dd, nomegainv, reg=None):
        """dd: [root, child..]; nomegainv: [root, child.., link] (owned children only)."""
        cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)
        self.inertia_root = self.kkt.factorize(np.asarray(dd[0], float), np.asarray(nomegainv[0], float),
                                               np.asarray(nomegainv[-1], float), cat(dd[1:]), cat(nomegainv[1:-1]), reg)
        return self.inertia_root

    def factorize_with_correct_inertia(self, dd, nomegainv, leaf_strategies, barrier_parameter=1.0):
        """DistributedRootLinearSystem::factor2 with REGULARIZATION = true, batched: every owned leaf runs the reference's
        per-node loop LinearSystem::factorize_with_correct_inertia (LS/LinearSystem.C:296-325; regularization.py) with its own
        strategy object. All leaves are factorised together (one batched device factorisation per round); a leaf whose
        inertia is already correct keeps its regularisation. The Schur complement is always the one of the final round
        (it is re-accumulated from zero by every pipsgpu_kkt_factorize). Returns (root inertia, per-leaf regularisation,
        number of batched factorisations)."""
        n = len(self.leaves)
        assert len(leaf_strategies) == n
        reg = []
        for st, leaf in zip(leaf_strategies, self.leaves):
            st.notify_new_step()
            reg.append(tuple(st.get_default_regularization()))
            leaf.set_regularization(*reg[-1])
        rounds = 0
        while True:
            root_inertia = self.factorize(dd, nomegainv)
            rounds += 1
            wrong = [i for i, (st, leaf) in enumerate(zip(leaf_strategies, self.leaves))
                     if not st.is_inertia_correct(leaf.inertia()[0])]
            if not wrong:
                return root_inertia, reg, rounds
            if rounds > 60:
                raise RuntimeError("factorize_with_correct_inertia: inertia still wrong after 60 rounds")
            for i in wrong:
                cur = tuple(leaf_strategies[i].get_regularization_parameters(self.leaves[i].inertia()[0], barrier_parameter))
                assert all(c >= l for c, l in zip(cur, reg[i]))
                reg[i] = cur
                self.leaves[i].set_regularization(*cur)

    def solve_bicgstab(self, rhs, tol=1e-10, max_iter=75):
        """LinearSystem::solveCompressedBiCGStab on the device (pipsgpu_kkt_solve): rhs = [b0, b_1, ..., b_N], in place.
        Returns (status, iterations, relative residual)."""
        self._b0[:] = rhs[0]
        off = 0
        for b in rhs[1:]:
            self._b_leaves[off:off + len(b)] = b
            off += len(b)
        out = self.kkt.solve_bicgstab(self._b0, self._b_leaves, tol, max_iter)
        rhs[0][:] = self._b0
        off = 0
        for b in rhs[1:]:
            b[:] = self._b_leaves[off:off + len(b)]
            off += len(b)
        return out

    def solve_compressed(self, rhs):
        """rhs = [b0, b_1, ..., b_N] (owned children), solved in place."""
        self._b0[:] = rhs[0]
        off = 0
        for b in rhs[1:]:
            self._b_leaves[off:off + len(b)] = b
            off += len(b)
        self.kkt.solve_compressed(self._b0, self._b_leaves)
        rhs[0][:] = self._b0
        off = 0
        for b in rhs[1:]:
            b[:] = self._b_leaves[off:off + len(b)]
            off += len(b)
        return rhs

    def schur_complement(self):
        """assembled (un-factorised) Schur complement, lower triangle row-major"""
        return self.kkt.sc().download()

    def close(self):
        self.kkt.close()
        self.ctx.close()
