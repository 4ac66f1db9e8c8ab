"""CPU tests (-m "not gpu") of the host side of the product: the C-ABI library loads and exports every symbol
include/pipsgpu.h declares, fails loudly without a device (no CPU fallback), the rank partition equals the
reference's mapChildrenToNSubTrees, and the symbolic analysis (matching, nested dissection, supernodes, front tree)
produces a structurally valid elimination plan."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pipsgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pipsgpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.capi.lib()
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pipsgpu_version() >= 100


def test_no_device_means_error_not_fallback(pkg):
    """without a GPU every compute entry point must fail with PIPSGPU_ERR_NO_DEVICE (2) - never compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.capi.PipsGpuError) as e:
        pkg.capi.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/"""
    pkgdir = os.path.join(ROOT, "pips-ipmpp_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".py") is False or "import" not in src.split("oracle/")[0][-40:], f


def _reference_partition(N, P):
    """mapChildrenToNSubTrees restated independently (Readers/Distributed/DistributedTree.C:437-475)"""
    sizes = [N // P] * P
    left = N % P
    if left:
        free = P - left
        after, extra = divmod(free, left)
        i = assigned = 0
        while i < P:
            sizes[i] += 1
            i += 1
            if extra != assigned:
                i += 1
                assigned += 1
            i += after
    out, off = [], 0
    for s in sizes:
        out.append((off, off + s))
        off += s
    return out


@pytest.mark.parametrize("N,P", [(4, 2), (64, 1), (512, 8), (128, 8), (8192, 8), (10, 4), (7, 3), (13, 5), (9, 8), (5, 5)])
def test_partition_matches_reference(pkg, N, P):
    from oracle.linsys import partition_children
    got = pkg.capi.partition(N, P)
    assert got == _reference_partition(N, P) == partition_children(N, P)
    assert got[0][0] == 0 and got[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
    sizes = [hi - lo for lo, hi in got]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == N


def test_partition_rejects_more_ranks_than_children(pkg):
    with pytest.raises(pkg.capi.PipsGpuError):
        pkg.capi.partition(3, 4)      # DistributedTree.C:56-60 enforces P <= N


def _symbolic(pkg, lp, i, max_width=0):
    from oracle import mf
    ch = lp.children[i]
    krow, jcol, val, diag_pos, nB, brow, bcol, bval = mf.leaf_kkt_and_border(lp, i)
    sym = pkg.capi.Symbolic(ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val, nB, brow, bcol, max_width)
    return sym, (krow, jcol, val, nB, brow, bcol)


@pytest.mark.parametrize("cfg,max_width", [("T1", 0), ("T2", 0), ("T3", 8), ("T3", 32)])
def test_symbolic_analysis_invariants(pkg, cfg, max_width):
    lp = pkg.problem.synthetic(cfg)
    ch = lp.children[0]
    sym, (krow, jcol, val, nB, brow, bcol) = _symbolic(pkg, lp, 0, max_width)
    a = sym.export()
    n, ns = a["n"], a["ns"]
    perm = a["perm"]
    assert sorted(perm.tolist()) == list(range(n))
    iperm = np.empty(n, int); iperm[perm] = np.arange(n)
    # 2x2 pivots: (primal, equality-row) pairs adjacent, first is a primal column, second an equality row
    pt = a["ptype"]
    firsts = np.nonzero(pt == 1)[0]
    assert np.all(pt[firsts + 1] == 2) and sym.sizes["n_pairs"] == len(firsts)
    assert np.all(perm[firsts] < ch.n) and np.all((perm[firsts + 1] >= ch.n) & (perm[firsts + 1] < ch.n + ch.my))
    # every equality row (zero diagonal) is covered by a 2x2 pivot
    assert len(firsts) == ch.my
    assert np.all(a["sgn"] == np.where(perm < ch.n, 1, -1))
    # supernodes: contiguous cover, bounded width, never split a pair
    col = a["sn_col"]
    assert col[0] == 0 and col[-1] == n and np.all(np.diff(col) >= 1)
    assert np.diff(col).max() <= (max_width or 32)
    assert not np.any(pt[col[1:-1]] == 2)
    # rows of each front ascending, strictly after its columns; parent = front owning the first non-border row
    rp, rows = a["sn_rowptr"], a["rowidx"]
    sn_of = np.repeat(np.arange(ns), np.diff(col))
    for s in range(ns):
        r = rows[rp[s]:rp[s + 1]]
        assert np.all(np.diff(r) > 0)
        assert len(r) == a["sn_na"][s] + a["sn_nb"][s]
        if len(r):
            assert r[0] >= col[s + 1]
        na = a["sn_na"][s]
        assert np.all(r[:na] < n) and np.all(r[na:] >= n) and np.all(r[na:] < n + nB)
        p = a["sn_parent"][s]
        if na > 0:
            assert p == sn_of[r[0]] and a["sn_level"][s] == a["sn_level"][p] + 1
            # relative indices address the same global row in the parent front
            prow = np.concatenate([np.arange(col[p], col[p + 1]), rows[rp[p]:rp[p + 1]]])
            assert np.array_equal(prow[a["relidx"][rp[s]:rp[s + 1]]], r)
        else:
            assert p == -1 and a["sn_level"][s] == 0
    # every entry of K and of the border lands inside its front panel, no two entries on the same slot
    assert len(np.unique(a["kmap"])) == len(a["kmap"])
    assert a["kmap"].min() >= 0 and a["kmap"].max() < a["Loff"][-1]
    if len(a["bmap"]):
        both = np.concatenate([a["kmap"], a["bmap"]])
        assert len(np.unique(both)) == len(both) and a["bmap"].max() < a["Loff"][-1]
    # the symbolic structure contains the exact Cholesky structure of P K P^T (checked against a dense symbolic run)
    if n <= 400:
        S = np.zeros((n, n), bool)
        for i in range(n):
            for q in range(krow[i], krow[i + 1]):
                pi, pj = iperm[i], iperm[jcol[q]]
                S[max(pi, pj), min(pi, pj)] = True
        for j in range(n):
            r = np.nonzero(S[j + 1:, j])[0] + j + 1
            if len(r):
                S[np.ix_(r, r)] |= np.tri(len(r), dtype=bool)
        for s in range(ns):
            front_rows = set(range(col[s], col[s + 1])) | set(int(v) for v in rows[rp[s]:rp[s + 1]] if v < n)
            for j in range(col[s], col[s + 1]):
                need = set(np.nonzero(S[j:, j])[0] + j)
                assert need <= front_rows
        assert sym.sizes["nnzL"] == int(np.tril(S).sum())


@pytest.mark.parametrize("cfg,big_u,frac", [("T2", 4, 0.6), ("T3", 8, 0.5), ("T3", 16, 0.9), ("T3", 2, 0.05)])
def test_border_groups_are_nested_and_equivalent(pkg, cfg, big_u, frac):
    """the grouping behind border_group_kernel: every big front with border rows sits in exactly one group, a group is a
    connected piece of the tree below its root, positions address the same border index in the root's row list, and
    accumulating sum_f L_B,f D_f L_B,f^T through the group positions (what the kernel does, 64-row tiles included)
    equals accumulating it front by front through the border indices"""
    lp = pkg.problem.synthetic(cfg)
    sym, (krow, jcol, val, nB, brow, bcol) = _symbolic(pkg, lp, 0)
    a = sym.export()
    g = sym.border_groups(big_u, frac)
    n, ns = a["n"], a["ns"]
    rp, rows, na, nb, par = a["sn_rowptr"], a["rowidx"], a["sn_na"], a["sn_nb"], a["sn_parent"]
    w = np.diff(a["sn_col"])
    want = {f for f in range(ns) if na[f] + nb[f] >= big_u and nb[f] > 0}
    assert sorted(g["mem_front"].tolist()) == sorted(want) and len(want) > 0
    rng = np.random.default_rng(5)
    direct = np.zeros((nB, nB))
    grouped = np.zeros((nB, nB))
    for gi in range(len(g["root"])):
        root, N = int(g["root"][gi]), int(g["N"][gi])
        mem = g["mem_front"][g["mem_ptr"][gi]:g["mem_ptr"][gi + 1]]
        assert mem[0] == root and N == nb[root]
        rootb = rows[rp[root] + na[root]:rp[root + 1]] - n           # border index of every root row
        inside = set(int(f) for f in mem)
        nblk = (N + 63) // 64
        acc = np.zeros((nblk * 64, nblk * 64))
        for q, f in zip(range(g["mem_ptr"][gi], g["mem_ptr"][gi + 1]), mem):
            pos = g["pos"][g["pos_ptr"][q]:g["pos_ptr"][q + 1]]
            fb = rows[rp[f] + na[f]:rp[f + 1]] - n
            assert len(pos) == nb[f] and np.all(np.diff(pos) > 0)
            assert np.array_equal(rootb[pos], fb)
            if f != root:
                assert int(par[f]) in inside and nb[f] >= frac * N
            LB = rng.standard_normal((nb[f], w[f]))
            D = rng.standard_normal((w[f], w[f])); D = D + D.T
            U = LB @ D @ LB.T
            direct[np.ix_(fb, fb)] += U
            # the kernel's staging: per 64-row block an inverse map position -> member row, absent rows are zero
            inv = -np.ones(nblk * 64, int); inv[pos] = np.arange(nb[f])
            A = np.where(inv[:, None] >= 0, LB[np.maximum(inv, 0)], 0.0)
            acc += A @ D @ A.T
        assert np.all(acc[N:, :] == 0) and np.all(acc[:, N:] == 0)
        grouped[np.ix_(rootb, rootb)] += acc[:N, :N]
    assert np.allclose(grouped, direct, rtol=1e-12, atol=1e-12)


def test_symbolic_is_deterministic_and_counts_flops(pkg):
    lp = pkg.problem.synthetic("T2")
    s1, _ = _symbolic(pkg, lp, 1)
    s2, _ = _symbolic(pkg, lp, 1)
    a, b = s1.export(), s2.export()
    assert all(np.array_equal(a[k], b[k]) for k in a if isinstance(a[k], np.ndarray))
    assert s1.flops == s2.flops > 0
    # sum_j eta_j (eta_j + 2) >= 3 * eta(L) for eta_j >= 1
    eta = s1.sizes["nnzL"] - s1.sizes["n"]
    assert s1.flops >= eta


def test_create_kkt_lower_layout(pkg):
    """create_kkt (DistributedLeafLinearSystem.C:52-72): lower CSR, full diagonal, W rows at nx.., D rows at nx+my.."""
    lp = pkg.problem.synthetic("T1")
    ch = lp.children[0]
    krow, jcol, val = pkg.linsys.create_kkt_lower(ch)
    nK = ch.n + ch.my + ch.mz
    assert len(krow) == nK + 1 and krow[-1] == len(jcol) == nK + ch.W.nnz + ch.D.nnz
    for i in range(nK):
        cols = jcol[krow[i]:krow[i + 1]]
        assert np.all(np.diff(cols) > 0) and cols[-1] == i
    import scipy.sparse as sp
    K = sp.csr_matrix((val, jcol, krow), shape=(nK, nK)).toarray()
    assert np.array_equal(K[ch.n:ch.n + ch.my, :ch.n], ch.W.toarray())
    assert np.array_equal(K[ch.n + ch.my:, :ch.n], ch.D.toarray())
    assert np.all(np.diag(K) == 0)


def test_synthetic_generator_is_deterministic_and_shaped(pkg):
    P = pkg.problem
    a, b = P.synthetic("T2"), P.synthetic("T2")
    assert (a.children[3].W != b.children[3].W).nnz == 0
    # a rank's slice of the blocks is identical to the same blocks of the whole problem (contiguous partition)
    part = P.synthetic("T2", block_range=(2, 4))
    assert len(part.children) == 2 and (part.children[0].W != a.children[2].W).nnz == 0
    nb, my, mz, n, n0, my0, mz0, myl, mzl = P.CONFIGS["C3"]
    assert (nb, n + my + mz, n0 + my0 + myl + mzl) == (512, 110000, 2700)
    nb, my, mz, n, n0, my0, mz0, myl, mzl = P.CONFIGS["C2"]
    assert (nb, n + my + mz, n0 + my0 + myl + mzl) == (64, 44000, 800)


@pytest.mark.parametrize("cfg", ["T2", "T3"])
def test_xform_front_solve_equals_substitution(pkg, cfg):
    """the algebra of the streaming solve kernels (psolve.cu with PIPSGPU_XFORM, psolve_direct.cu), restated in numpy on
    the oracle's factor: per front  y = X y  (X = L11^-1),  z = D^-1 y  fused into the forward sweep,  rows above -= L21 y;
    backward  t = x - L21^T x_anc,  x = X^T t.  Must equal the oracle's plain substitution solve (mf_solve)."""
    from oracle import mf
    lp = pkg.problem.synthetic(cfg)
    dd, nom = pkg.problem.synthetic_iterate(lp, pkg.problem.CONFIG_ID[cfg])
    ch = lp.children[0]
    krow, jcol, val, diag_pos, nB, brow, bcol, bval = mf.leaf_kkt_and_border(lp, 0)
    val = val.copy()
    val[diag_pos[:ch.n]] = dd[1]
    val[diag_pos[ch.n + ch.my:]] = nom[1]
    n = ch.n + ch.my + ch.mz
    none = np.zeros(0, np.int32)
    a = pkg.capi.Symbolic(n, ch.n, ch.my, ch.mz, krow, jcol, val, 0, none, none).export()
    leaf = mf.MfLeaf(a)
    leaf.factorize(val, np.zeros(0))
    rng = np.random.default_rng(7)
    b = rng.standard_normal(n)
    want = leaf.solve(b)
    col, rp, rows, ns = a["sn_col"], a["sn_rowptr"], a["rowidx"], a["ns"]
    pt, did, die = a["ptype"], leaf.dinv_d, leaf.dinv_e
    y = b[a["perm"]].copy()
    fronts = []
    for s in range(ns):
        c0, w = col[s], col[s + 1] - col[s]
        na = a["sn_na"][s]
        m = w + na + a["sn_nb"][s]
        F = leaf.L[a["Loff"][s]:a["Loff"][s] + m * w].reshape(w, m).T          # column-major panel -> (m, w)
        L11 = np.tril(F[:w, :w], -1) + np.eye(w)
        X = np.linalg.inv(L11)
        L21 = F[w:w + na, :]
        r = rows[rp[s]:rp[s] + na]
        fronts.append((c0, w, X, L21, r))
        yk = X @ y[c0:c0 + w]                                                  # y = X y
        z = did[c0:c0 + w] * yk                                                # z = D^-1 y (2x2 pivots couple neighbours)
        for k in range(w):
            if pt[c0 + k] == 1:
                z[k] += die[c0 + k] * yk[k + 1]
            elif pt[c0 + k] == 2:
                z[k] += die[c0 + k - 1] * yk[k - 1]
        y[r] -= L21 @ yk                                                       # rows above -= L21 y
        y[c0:c0 + w] = z
    for c0, w, X, L21, r in reversed(fronts):
        t = y[c0:c0 + w] - L21.T @ y[r]
        y[c0:c0 + w] = X.T @ t
    got = np.empty(n)
    got[a["perm"]] = y
    assert np.linalg.norm(got - want) <= 1e-11 * np.linalg.norm(want)


def test_host_kkt_residual_helper_against_oracle_solution(pkg):
    """linsys.compressed_kkt_residual (the size-independent check bench.py and the full-size GPU tests rely on) must
    vanish on the oracle's solveCompressed solution, with and without regularisation"""
    from oracle.linsys import OracleRoot
    P = pkg.problem
    for cfg, reg in (("T2", None), ("T3", (1e-6, 1e-7, 1e-7))):
        lp = P.synthetic(cfg)
        dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[cfg])
        ref = OracleRoot(lp)
        ref.put_diagonals(dd, nom, *(reg or (0.0, 0.0, 0.0)))
        ref.factor2()
        rng = np.random.default_rng(0)
        nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
        rhs = [rng.standard_normal(nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
        sol = ref.solve_compressed([b.copy() for b in rhs])
        lsq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol, reg)
        assert np.sqrt(lsq + r0 @ r0) / np.sqrt(sum(b @ b for b in rhs)) <= 1e-12
        # and it must NOT vanish for a wrong solution
        sol[1][0] += 1.0
        lsq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol, reg)
        assert np.sqrt(lsq + r0 @ r0) > 1e-6


def test_regularisation_loop_host_logic_against_oracle(pkg):
    """regularization.py (mirror of LinearSystem::factorize_with_correct_inertia + IpoptRegularization) driven by the C
    oracle's factorisation must take exactly the steps of the independent restatement in oracle/linsys.py"""
    from oracle import mf
    from oracle.linsys import oracle_leaf_inertia_loop
    P = pkg.problem
    R = pkg.regularization
    lp = P.synthetic("T2")
    ch = lp.children[2]
    W = ch.W.tolil()
    W[1, :] = W[0, :]
    ch.W = P._csr(W.tocsr(), W.shape)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["T2"])
    want_reg, want_n, seen = oracle_leaf_inertia_loop(lp, 2, dd[3], nom[3], mf.HostSymbolic, barrier_parameter=1e-2)
    assert seen[0] != (ch.n, ch.my + ch.mz, 0) and seen[-1] == (ch.n, ch.my + ch.mz, 0)
    krow, jcol, val, diag_pos, nB, brow, bcol, bval = mf.leaf_kkt_and_border(lp, 2)
    leaf = mf.MfLeaf(mf.HostSymbolic(ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val, nB, brow, bcol).export())
    knorm = max(1.0, float(np.abs(val).max(initial=0.0)))
    reg = [0.0, 0.0, 0.0]
    state = {}

    def add(dp, dy, dz):
        reg[0] += dp; reg[1] += dy; reg[2] += dz

    def changed():
        v = val.copy()
        v[diag_pos[:ch.n]] = dd[3] + reg[0]
        v[diag_pos[ch.n:ch.n + ch.my]] = -reg[1]
        v[diag_pos[ch.n + ch.my:]] = nom[3] - reg[2]
        state["inertia"], _ = leaf.factorize(v, bval, None, knorm=knorm)

    final, nfact = R.factorize_with_correct_inertia(R.IpoptRegularization(ch.n, ch.my + ch.mz), add, changed,
                                                    lambda: state["inertia"], 1e-2)
    assert nfact == want_n and tuple(final) == tuple(want_reg) and tuple(reg) == tuple(want_reg)
    # Friedlander-Orban: decreasing default regularisation per new step, x100 per retry
    fo = R.FriedlanderOrbanRegularization(3, 2)
    fo.notify_new_step()
    assert fo.get_default_regularization() == (1.0, 1.0, 1.0)
    assert fo.get_regularization_parameters() == (100.0, 100.0, 100.0)
    fo.notify_new_step()
    assert fo.get_default_regularization() == (10.0, 10.0, 10.0)


def test_sparse_root_pattern_covers_the_oracle_schur_complement(pkg):
    """sc_lower_pattern (what the reference precomputes for its sparse root KKT, LS/sLinsysRootAug.C:79-208) must contain
    every structural nonzero of the Schur complement the oracle assembles, and be sparse on an LP with local links."""
    import numpy as np
    import scipy.sparse as sp
    from oracle.linsys import OracleRoot
    P = pkg.problem
    for lp in (P.localize_links(P.synthetic("S1"), halo=0.5), P.synthetic("T2"), P.c1_hier_4blocks()):
        krow, jcol = pkg.linsys.sc_lower_pattern(lp)
        n = len(krow) - 1
        assert n == lp.root.n + lp.root.my + lp.myl + lp.mzl
        pat = sp.csr_matrix((np.ones(krow[-1]), jcol, krow), shape=(n, n)).toarray() > 0
        assert np.all(np.diag(pat)) and not np.any(np.triu(pat, 1))
        rng = np.random.default_rng(0)
        nodes = [lp.root] + lp.children
        dd = [rng.uniform(0.5, 2.0, nd.n) for nd in nodes]
        nom = [-rng.uniform(0.5, 2.0, nd.mz) for nd in nodes] + [-rng.uniform(0.5, 2.0, lp.mzl)]
        ref = OracleRoot(lp)
        ref.put_diagonals(dd, nom)
        ref.factor2()
        assert not np.any((np.tril(ref.SC) != 0) & ~pat)
    krow, _ = pkg.linsys.sc_lower_pattern(P.localize_links(P.synthetic("S1"), halo=0.5))
    n = len(krow) - 1
    assert krow[-1] < 0.35 * n * (n + 1) / 2
