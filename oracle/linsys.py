"""CPU ORACLE (test infrastructure only - never imported by the product path).

numpy/scipy restatement of the reference's distributed KKT linear-system layer for the two-level
(non-hierarchical) dense-Schur-complement path, function by function:

  OracleLeaf            <- LinearSystems/DistributedLeafLinearSystem.C:52-252 (create_kkt, diagonals, factor2,
                           addLniziLinkCons, addTermToSchurComplBlocked) and
                           LinearSystems/DistributedLinearSystem.C:430-483 (LniTransMult), :766-1047 (blocked
                           Bord^T K^-1 Bord), :1116-1175 (addLeftBorderTimesDenseColsToResTranspDense)
  OracleRoot            <- LinearSystems/DistributedRootLinearSystem.C:206-243 (factor2), :840-881 (initializeKKT,
                           reduceKKTdense), :1436-1464 (factorizeKKT), :1614-1707 (submatrixAllReduce*);
                           LinearSystems/sLinsysRootAug.C:60-76, 210-466, 1276-1338, 1769-1813
  solve_compressed      <- LinearSystems/DistributedLinearSystem.C:409-420 (Lsolve; Dsolve; Ltsolve)

Third-party arithmetic that is NOT in /root/reference (ThirdPartyLibs/* only holds download scripts):
  * sparse symmetric-indefinite solve of the leaf KKT (reference: PARDISO >= 7.2 / MKL PARDISO / HSL MA27, MA57 /
    MUMPS behind LinearSolvers/DoubleLinearSolver.h:24-72, no version pins exist). Here: SuperLU (scipy 1.18
    `scipy.sparse.linalg.splu`, partial pivoting) - the one sparse direct solver available offline. It plays the
    role of the solver plugin: the oracle only relies on the operator K_i^-1, exactly like the reference does.
  * dense symmetric-indefinite LDL^T of the Schur complement (reference: LAPACK dsytrf/dsytrs,
    LinearSolvers/DenseSymmetricIndefinitSolver/DeSymIndefSolver.C:56-126). Here: the same LAPACK routines from
    scipy's bundled OpenBLAS (`scipy.linalg.lapack.dsytrf/dsytrs`).

PARITY PINNING: the reference pins nothing at the solver boundary; the only golden values for this
path are end-to-end (objective, iteration count) pairs - tests/test_oracle_golden.py checks this oracle against
them (hier_approach_4blocks_2by2 -> 14.0 in <= 5 iterations; callbackExample -> 14) and against dense numpy
evaluations of the Schur complement / step direction.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.linalg import lapack


def build_leaf_kkt_lower(ch):
    """create_kkt (DistributedLeafLinearSystem.C:52-72): lower CSR of
         [ Q+diag  .    .  ]
         [ W       0    .  ]      with a full (explicitly stored) diagonal, Q = 0 for LPs.
         [ D       .    0  ]"""
    n, my, mz = ch.n, ch.my, ch.mz
    nK = n + my + mz
    K = sp.bmat([[sp.csr_matrix((n, n)), None, None],
                 [ch.W, sp.csr_matrix((my, my)), None],
                 [ch.D, None, sp.csr_matrix((mz, mz))]], format="coo")
    rows = np.concatenate([K.row, np.arange(nK)])
    cols = np.concatenate([K.col, np.arange(nK)])
    vals = np.concatenate([K.data, np.zeros(nK)])
    # keep explicit zeros of the pattern (setToDiagonal + symAtPutSubmatrix keep every structural entry)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(nK + 1, dtype=np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    diag_pos = indptr[1:] - 1          # diagonal is the last entry of each lower row
    return indptr, cols.astype(np.int32), vals.astype(np.float64), diag_pos


class OracleLeaf:
    """DistributedLeafLinearSystem restated. `blocksize` = SC_BLOCKWISE_BLOCKSIZE_MAX * threads
    (Options/PIPSIPMppOptions.C:236, default 20)."""

    def __init__(self, lp, i, blocksize=20):
        self.lp, self.ch = lp, lp.children[i]
        ch = self.ch
        self.n, self.my, self.mz = ch.n, ch.my, ch.mz
        self.nK = ch.n + ch.my + ch.mz
        self.krow, self.jcol, self.M, self.diag_pos = build_leaf_kkt_lower(ch)
        self.blocksize = blocksize
        self.lu = None
        self.pivot_eps = 1e-8
        self.perturbed = False
        # transposes used as "border_right" column sources (RACFG_BLOCK.h:23-33)
        self.A_csc = ch.A.tocsc()
        self.C_csc = ch.C.tocsc()
        self.Ft_csc = ch.F.T.tocsc()     # n_i x myl
        self.Gt_csc = ch.G.T.tocsc()

    # -- put_primal_diagonal / clear_dual_equality_diagonal / put_dual_inequalites_diagonal /
    #    add_regularization_local_kkt (DistributedLeafLinearSystem.C:88-143)
    def put_diagonals(self, dd, nomegainv, regP=0.0, regDy=0.0, regDz=0.0):
        n, my = self.n, self.my
        self.M[self.diag_pos[:n]] = dd + regP
        self.M[self.diag_pos[n:n + my]] = 0.0 - regDy
        self.M[self.diag_pos[n + my:]] = nomegainv - regDz

    def full_matrix(self):
        L = sp.csr_matrix((self.M, self.jcol, self.krow), shape=(self.nK, self.nK))
        return (L + sp.tril(L, -1).T).tocsc()

    # -- factor2 -> solver->matrixChanged()
    def factor2(self):
        """solver->matrixChanged(). The reference's default plugin (PARDISO, iparm[9]=8: pivots smaller than
        1e-8*||K|| are replaced by +-1e-8*||K||, PardisoProjectSolver.C:68-77) survives singular leaf systems such as
        scenario 2 of callbackExample.cpp (an all-zero equality row). SuperLU has no pivot perturbation, so the oracle
        emulates it: if the plain factorisation is singular or has a tiny pivot, the zero (2,2) diagonal is replaced
        by -1e-8*||K|| and the matrix is refactorised. `self.perturbed` records that this happened."""
        K = self.full_matrix()
        knorm = max(np.abs(K.data).max(initial=0.0), 1.0)
        self.perturbed = False
        opts = dict(SymmetricMode=False, DiagPivotThresh=1.0)
        try:
            lu = spla.splu(K, permc_spec="COLAMD", options=opts)
            if np.abs(lu.U.diagonal()).min() < self.pivot_eps * knorm:
                raise RuntimeError("tiny pivot")
        except RuntimeError:
            self.perturbed = True
            n, my = self.n, self.my
            M = self.M.copy()
            yd = self.diag_pos[n:n + my]
            M[yd] = np.where(np.abs(M[yd]) < self.pivot_eps * knorm, -self.pivot_eps * knorm, M[yd])
            xd = self.diag_pos[:n]
            M[xd] = np.where(np.abs(M[xd]) < self.pivot_eps * knorm, self.pivot_eps * knorm, M[xd])
            L = sp.csr_matrix((M, self.jcol, self.krow), shape=(self.nK, self.nK))
            lu = spla.splu((L + sp.tril(L, -1).T).tocsc(), permc_spec="COLAMD", options=opts)
        self.lu = lu

    def solve(self, rhs):
        """DoubleLinearSolver::solve(nrhs, rhss, colSparsity): rhs is (nrhs, nK) with RHS j contiguous in row j."""
        rhs = np.asarray(rhs, dtype=np.float64)
        if rhs.ndim == 1:
            return self.lu.solve(rhs)
        return np.ascontiguousarray(self.lu.solve(np.ascontiguousarray(rhs.T)).T)

    # -- addTermToSchurComplBlocked -> addBiTLeftKiBiRightToResBlockedParallelSolvers
    def add_term_to_schur_compl_blocked(self, SC):
        lp, ch = self.lp, self.ch
        n0, my0, myl, mzl = lp.root.n, lp.root.my, lp.myl, lp.mzl
        nS = SC.shape[0]
        n, my, mz, nK = self.n, self.my, self.mz, self.nK

        def process(col_ids, gather, sc_offset):
            # chunks of <= blocksize border columns; only structurally nonzero ones become right-hand sides
            for c0 in range(0, len(col_ids), self.blocksize):
                chunk = col_ids[c0:c0 + self.blocksize]
                cols = [c for c in chunk if gather(c, None)]
                if not cols:
                    continue
                dense = np.zeros((len(cols), nK))
                for j, c in enumerate(cols):
                    gather(c, dense[j])          # fromGetColsBlock: RHS j contiguous
                X = self.solve(dense)
                # addLeftBorderTimesDenseColsToResTranspDense: SC[row, :] -= [R^T A^T C^T; F; G] x
                for j, c in enumerate(cols):
                    x = X[j]
                    row = sc_offset + c
                    SC[row, 0:n0] -= ch.A.T @ x[n:n + my] + ch.C.T @ x[n + my:]
                    if myl > 0:
                        SC[row, nS - myl - mzl:nS - mzl] -= ch.F @ x[:n]
                    if mzl > 0:
                        SC[row, nS - mzl:] -= ch.G @ x[:n]

        def gather_RAC(c, out):
            a0, a1 = self.A_csc.indptr[c], self.A_csc.indptr[c + 1]
            c0_, c1_ = self.C_csc.indptr[c], self.C_csc.indptr[c + 1]
            if out is None:
                return (a1 - a0) + (c1_ - c0_) > 0
            out[n + self.A_csc.indices[a0:a1]] = self.A_csc.data[a0:a1]
            out[n + my + self.C_csc.indices[c0_:c1_]] = self.C_csc.data[c0_:c1_]
            return True

        def gather_T(mat):
            def g(c, out):
                p0, p1 = mat.indptr[c], mat.indptr[c + 1]
                if out is None:
                    return p1 - p0 > 0
                out[mat.indices[p0:p1]] = mat.data[p0:p1]
                return True
            return g

        process(list(range(n0)), gather_RAC, 0)
        if myl > 0:
            process(list(range(myl)), gather_T(self.Ft_csc), n0 + my0)
        if mzl > 0:
            process(list(range(mzl)), gather_T(self.Gt_csc), n0 + my0 + myl)

    # -- addLniziLinkCons (DistributedLeafLinearSystem.C:171-212)
    def add_lnizi_link_cons(self, b0, bi):
        lp, ch = self.lp, self.ch
        n, my = self.n, self.my
        zi = self.solve(bi)
        bi[:] = zi
        n0 = lp.root.n
        b0[:n0] -= ch.A.T @ zi[n:n + my] + ch.C.T @ zi[n + my:]
        L = len(b0)
        if lp.myl > 0:
            b0[L - lp.myl - lp.mzl:L - lp.mzl] -= ch.F @ zi[:n]
        if lp.mzl > 0:
            b0[L - lp.mzl:] -= ch.G @ zi[:n]

    # -- Ltsolve2 -> LniTransMult (DistributedLinearSystem.C:430-483): b_i -= K_i^-1 (Bord_i u)
    def ltsolve2(self, bi, u):
        lp, ch = self.lp, self.ch
        n, my = self.n, self.my
        n0 = lp.root.n
        t = np.zeros(self.nK)
        t[n:n + my] = ch.A @ u[:n0]
        t[n + my:] = ch.C @ u[:n0]
        L = len(u)
        if lp.myl > 0:
            t[:n] += ch.F.T @ u[L - lp.myl - lp.mzl:L - lp.mzl]
        if lp.mzl > 0:
            t[:n] += ch.G.T @ u[L - lp.mzl:]
        bi -= self.solve(t)


def partition_children(N, P):
    """mapChildrenToNSubTrees (Readers/Distributed/DistributedTree.C:437-475): contiguous ranges, floor(N/P) each,
    the N mod P leftovers spread by stride over the ranks. Returns list of (lo, hi) per rank."""
    assert 1 <= P <= max(N, 1)
    everyone_gets, n_leftovers = divmod(N, P)
    sizes = np.full(P, everyone_gets, dtype=int)
    if n_leftovers > 0:
        free_in_row = P - n_leftovers
        free_after, additional = divmod(free_in_row, n_leftovers)
        assigned = 0
        i = 0
        while i < P:
            sizes[i] += 1
            i += 1
            if additional != assigned:
                i += 1
                assigned += 1
            i += free_after
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(offs[r]), int(offs[r + 1])) for r in range(P)]


class OracleRoot:
    """sLinsysRootAug + DistributedRootLinearSystem (dense SC path). `owned` = indices of the children this rank owns
    (others are DistributedDummyLinearSystem no-ops)."""

    def __init__(self, lp, owned=None, blocksize=20, rank=0):
        self.lp = lp
        self.rank = rank
        self.owned = list(range(len(lp.children))) if owned is None else list(owned)
        self.leaves = {i: OracleLeaf(lp, i, blocksize) for i in self.owned}
        self.nS = lp.n_S
        self.SC = np.zeros((self.nS, self.nS))
        self.regDy = 0.0
        self.pivot_eps = 1e-12
        self.n_perturbed = 0

    # -- LinearSystem::factorize part that moves data: diagonals for every node
    def put_diagonals(self, dd, nomegainv, regP=0.0, regDy=0.0, regDz=0.0):
        """dd / nomegainv: lists over [root, child_0 .. child_{N-1}], nomegainv has one more trailing entry for the
        linking inequality rows (zDiagLinkCons)."""
        self.dd0 = np.asarray(dd[0], float) + regP
        self.zdiag0 = np.asarray(nomegainv[0], float) - regDz
        self.zdiag_link = np.asarray(nomegainv[-1], float) - regDz
        self.regDy = regDy
        for i, leaf in self.leaves.items():
            leaf.put_diagonals(dd[1 + i], nomegainv[1 + i], regP, regDy, regDz)

    def initialize_kkt(self):
        self.SC[:] = 0.0

    def assemble_local_kkt(self):
        for i in self.owned:
            self.leaves[i].factor2()
        for i in self.owned:
            self.leaves[i].add_term_to_schur_compl_blocked(self.SC)

    def reduce_pieces(self):
        """the three reduced pieces of reduceKKTdense (DistributedRootLinearSystem.C:860-881) as index masks."""
        lp = self.lp
        n0, my0 = lp.root.n, lp.root.my
        nS = self.nS
        m = np.zeros((nS, nS), dtype=bool)
        il = np.tril_indices(n0)
        m[il] = True                                   # submatrixAllReduceDiagLower(0, n0)
        m[n0 + my0:, 0:n0] = True                      # submatrixAllReduceFull(n0+my0, 0, myl+mzl, n0)
        k = nS - (n0 + my0)
        il = np.tril_indices(k)
        m[(il[0] + n0 + my0, il[1] + n0 + my0)] = True  # submatrixAllReduceDiagLower(n0+my0, myl+mzl)
        return m

    @staticmethod
    def reduce_kkt(roots):
        """MPI_Allreduce(SUM) of the three pieces over the 'ranks' in `roots` (in place)."""
        mask = roots[0].reduce_pieces()
        tot = np.zeros_like(roots[0].SC)
        for r in roots:
            tot[mask] += r.SC[mask]
        for r in roots:
            r.SC[mask] = tot[mask]

    def finalize_kkt_dense(self):
        lp = self.lp
        root = lp.root
        n0, my0, myl, mzl = root.n, root.my, lp.myl, lp.mzl
        SC = self.SC
        idx = np.arange(n0)
        SC[idx, idx] += self.dd0                                    # schur_complement_put_primal_block
        if root.mz > 0:                                             # schur_complement_add_CTDC_block
            CtDC = (root.C.T @ sp.diags(1.0 / self.zdiag0) @ root.C).toarray()
            SC[:n0, :n0] -= np.tril(CtDC)
        if my0 > 0:                                                 # schur_complement_put_AT_block ("yes, B")
            SC[n0:n0 + my0, :n0] += root.A.toarray()
            d = np.arange(n0, n0 + my0)
            SC[d, d] += -self.regDy
        if myl > 0:
            SC[n0 + my0:n0 + my0 + myl, :n0] += root.F.toarray()
            d = np.arange(n0 + my0, n0 + my0 + myl)
            SC[d, d] += -self.regDy
        if mzl > 0:
            SC[n0 + my0 + myl:, :n0] += root.G.toarray()
            d = np.arange(n0 + my0 + myl, self.nS)
            SC[d, d] += self.zdiag_link

    def sc_symmetric(self):
        L = np.tril(self.SC)
        return L + np.tril(L, -1).T

    def factorize_kkt(self):
        """DeSymIndefSolver::matrixChanged: copy + dsytrf (Bunch-Kaufman)."""
        S = self.sc_symmetric()
        self.ldu, self.ipiv, info = lapack.dsytrf(S, lower=1)
        assert info >= 0
        # A singular Schur complement (redundant equality rows, e.g. hier_approach_4blocks_2by2 has 15 equalities for
        # 14 variables) leaves dsytrf with a pivot of rounding-noise size. The reference's default root solver perturbs
        # such pivots to 1e-8*||SC|| (PARDISO iparm[9]=8, PardisoProjectIndefSolver.C:34-61); do the same to the 1x1
        # pivots of the LAPACK factor so that the outcome does not depend on rounding luck.
        # The threshold is relative to the largest entry of the pivot's own (original) row, so badly scaled but
        # well-posed systems are left untouched.
        perm = np.arange(self.nS)
        k = 0
        while k < self.nS:
            if self.ipiv[k] > 0:
                j = self.ipiv[k] - 1
                perm[[k, j]] = perm[[j, k]]
                k += 1
            else:
                j = -self.ipiv[k] - 1
                perm[[k + 1, j]] = perm[[j, k + 1]]
                k += 2
        rowmax = np.abs(S).max(axis=1, initial=0.0)
        self.n_perturbed = 0
        k = 0
        while k < self.nS:
            if self.ipiv[k] > 0:
                d = self.ldu[k, k]
                eps = self.pivot_eps * max(rowmax[perm[k]], 1e-300)
                if abs(d) < eps:
                    self.ldu[k, k] = -eps if d < 0 else eps
                    self.n_perturbed += 1
                k += 1
            else:
                k += 2
        return info

    def inertia(self):
        """DeSymIndefSolver.C:135-160 style: signs of the 1x1 / 2x2 diagonal blocks of D."""
        n = self.nS
        pos = neg = zero = 0
        k = 0
        d = self.ldu
        while k < n:
            if self.ipiv[k] > 0:
                v = d[k, k]
                pos += v > 0
                neg += v < 0
                zero += v == 0
                k += 1
            else:
                a, b, c = d[k, k], d[k + 1, k], d[k + 1, k + 1]
                ev = np.linalg.eigvalsh(np.array([[a, b], [b, c]]))
                pos += int((ev > 0).sum())
                neg += int((ev < 0).sum())
                k += 2
        return int(pos), int(neg), int(zero)

    def sc_solve(self, r):
        x, info = lapack.dsytrs(self.ldu, self.ipiv, r, lower=1)
        assert info == 0
        return x

    def factor2(self, roots=None):
        """DistributedRootLinearSystem::factor2 (:206-243). `roots` = all ranks' OracleRoot objects when emulating
        several ranks in one process."""
        self.initialize_kkt()
        self.assemble_local_kkt()
        if roots is None:
            OracleRoot.reduce_kkt([self])
            self.finalize_kkt_dense()
            self.factorize_kkt()

    # ---- solveCompressed ---------------------------------------------------------------------
    def lsolve_local(self, b):
        """Lsolve without the allreduce: b = [b0, b_child...] (b_child only for owned children)."""
        if self.rank > 0:
            b[0][:] = 0.0
        for i in self.owned:
            self.leaves[i].add_lnizi_link_cons(b[0], b[1 + i])

    def solve_reduced_link_cons(self, b0):
        lp = self.lp
        root = lp.root
        n0, my0, mz0, myl, mzl = root.n, root.my, root.mz, lp.myl, lp.mzl
        b1 = b0[:n0].copy()
        b2 = b0[n0:n0 + my0]
        b3 = b0[n0 + my0:n0 + my0 + mz0].copy()
        b45 = b0[n0 + my0 + mz0:]
        if mz0 > 0:
            b3 = b3 / self.zdiag0
            b1 -= root.C.T @ b3
        r = np.concatenate([b1, b2, b45])
        u = self.sc_solve(r)
        b0[:n0 + my0] = u[:n0 + my0]
        if mz0 > 0:
            t = b0[n0 + my0:n0 + my0 + mz0]
            t -= root.C @ u[:n0]
            t /= self.zdiag0
        b0[n0 + my0 + mz0:] = u[n0 + my0:]

    def ltsolve(self, b):
        for i in self.owned:
            self.leaves[i].ltsolve2(b[1 + i], b[0])

    def solve_compressed(self, b):
        """single-rank solveCompressed, in place on b = [b0(n0+my0+mz0+myl+mzl), b_1, ..., b_N]."""
        self.lsolve_local(b)
        self.solve_reduced_link_cons(b[0])
        self.ltsolve(b)
        return b


def solve_compressed_multi(roots, b):
    """solveCompressed over several emulated ranks: Lsolve locally, allreduce b0, Dsolve redundantly, Ltsolve."""
    b0s = []
    for r in roots:
        b0 = b[0].copy()
        bl = [b0] + b[1:]
        r.lsolve_local(bl)
        b0s.append(b0)
    tot = np.sum(b0s, axis=0)
    b[0][:] = tot
    roots[0].solve_reduced_link_cons(b[0])
    for r in roots:
        r.ltsolve(b)
    return b


def dense_reference_sc(lp, dd, nomegainv, regDy=0.0):
    """Independent dense evaluation of the Schur complement (Appendix B of SURVEY.md) used to pin the oracle:
    SC = sum_i -Bord_i^T K_i^-1 Bord_i + root block, full symmetric matrix via numpy.linalg.solve."""
    root = lp.root
    n0, my0, myl, mzl = root.n, root.my, lp.myl, lp.mzl
    nS = lp.n_S
    SC = np.zeros((nS, nS))
    for i, ch in enumerate(lp.children):
        n, my, mz = ch.n, ch.my, ch.mz
        nK = n + my + mz
        K = np.zeros((nK, nK))
        K[:n, :n] = np.diag(dd[1 + i])
        K[n:n + my, :n] = ch.W.toarray()
        K[n + my:, :n] = ch.D.toarray()
        K = np.tril(K) + np.tril(K, -1).T
        K[n:n + my, n:n + my] -= regDy * np.eye(my)
        K[n + my:, n + my:] = np.diag(nomegainv[1 + i])
        B = np.zeros((nK, nS))
        B[n:n + my, :n0] = ch.A.toarray()
        B[n + my:, :n0] = ch.C.toarray()
        B[:n, n0 + my0:n0 + my0 + myl] = ch.F.toarray().T
        B[:n, n0 + my0 + myl:] = ch.G.toarray().T
        SC -= B.T @ np.linalg.solve(K, B)
    SC[:n0, :n0] += np.diag(dd[0])
    if root.mz > 0:
        C0 = root.C.toarray()
        SC[:n0, :n0] -= C0.T @ np.diag(1.0 / nomegainv[0]) @ C0
    A0 = root.A.toarray()
    SC[n0:n0 + my0, :n0] += A0
    SC[:n0, n0:n0 + my0] += A0.T
    F0 = root.F.toarray()
    SC[n0 + my0:n0 + my0 + myl, :n0] += F0
    SC[:n0, n0 + my0:n0 + my0 + myl] += F0.T
    G0 = root.G.toarray()
    SC[n0 + my0 + myl:, :n0] += G0
    SC[:n0, n0 + my0 + myl:] += G0.T
    d = np.arange(n0 + my0 + myl, nS)
    SC[d, d] += nomegainv[-1]
    d = np.arange(n0, n0 + my0 + myl)
    SC[d, d] -= regDy
    return SC


# ---------------------------------------------------------------------------------------------------------------
# inertia-correcting regularisation, restated for ONE leaf on the C multifrontal oracle (oracle/mf_ldlt.c):
# LinearSystem::factorize_with_correct_inertia (LinearSystems/LinearSystem.C:296-325) with the IPOPT strategy
# (LinearSystems/IpoptRegularization.cpp:30-92, constants IpoptRegularization.hpp:22-27).
# ---------------------------------------------------------------------------------------------------------------
def oracle_leaf_inertia_loop(lp, i, dd_i, nom_i, symbolic_factory, barrier_parameter=1.0, primal_last=0.0):
    """returns (final (regP, regDy, regDz), number of factorisations, list of inertias seen)"""
    from . import mf
    ch = lp.children[i]
    krow, jcol, val, diag_pos, nB, brow, bcol, bval = mf.leaf_kkt_and_border(lp, i)
    sym = symbolic_factory(ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val, nB, brow, bcol).export()
    leaf = mf.MfLeaf(sym)
    knorm = max(1.0, float(np.abs(val).max(initial=0.0)))
    expected = (ch.n, ch.my + ch.mz, 0)

    def factor(reg):
        v = val.copy()
        v[diag_pos[:ch.n]] = dd_i + reg[0]
        v[diag_pos[ch.n:ch.n + ch.my]] = -reg[1]
        v[diag_pos[ch.n + ch.my:]] = nom_i - reg[2]
        inertia, _ = leaf.factorize(v, bval, None, knorm=knorm)
        return inertia

    reg = (0.0, 0.0, 0.0)                      # IpoptRegularization::get_default_regularization
    seen = [factor(reg)]
    primal_cur = -1.0
    dual_cur = 0.0
    first = True
    while seen[-1] != expected:
        if first:                              # get_regularization_new_matrix
            first = False
            dual_cur = barrier_parameter ** 0.25 if seen[-1][2] else 1e-4
            primal_cur = 1e-4 if primal_last == 0.0 else max(1e-20, primal_last / 3.0)
        else:                                  # get_regularization_nth_try
            primal_cur *= 100.0 if primal_last == 0.0 else 8.0
            if primal_cur > 1e40:
                raise RuntimeError("cannot factorize matrix after excessive error correction")
        reg = (primal_cur, dual_cur, dual_cur)
        seen.append(factor(reg))
        if len(seen) > 60:
            raise RuntimeError("inertia loop did not terminate")
    return reg, len(seen), seen
