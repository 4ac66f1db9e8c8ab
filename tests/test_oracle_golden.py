"""CPU tests (-m "not gpu"): the oracle against the reference's golden values, the committed fixtures and independent
dense numpy evaluations. The oracle is the checker of the GPU parity tests, so it is pinned here first."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _split(lp, flat, kind):
    """flat concatenation -> per-node list. kind 'dd': [root, children]; 'nom': [root, children, link];
    'rhs': [b0, b_children]"""
    out, off = [], 0
    if kind == "dd":
        sizes = [lp.root.n] + [ch.n for ch in lp.children]
    elif kind == "nom":
        sizes = [lp.root.mz] + [ch.mz for ch in lp.children] + [lp.mzl]
    else:
        sizes = [lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl] + [ch.n + ch.my + ch.mz for ch in lp.children]
    for s in sizes:
        out.append(np.array(flat[off:off + s]))
        off += s
    assert off == len(flat)
    return out


def _instances(pkg):
    P = pkg.problem
    return {"c1_hier_4blocks": P.c1_hier_4blocks, "callback_example": P.callback_example,
            "synthetic_T1": lambda: P.synthetic("T1")}


@pytest.mark.parametrize("name", ["c1_hier_4blocks", "callback_example"])
def test_reference_goldens_end_to_end(pkg, name):
    """objective 14.0 / <= 1.1 x 4 iterations (TST/IntegrationTests/gamssmall_instance_data.txt:8 via t_pips.cpp:115-119)
    and objective 14 (TST/pipsipmLinkConsTest.sh:25)"""
    from oracle.ipm import IPM, OracleBackend
    lp = _instances(pkg)[name]()
    status, obj, iters = IPM(lp, OracleBackend(lp)).solve()
    assert status == "SUCCESSFUL_TERMINATION"
    assert abs(obj - 14.0) <= 1e-4
    if "expected_iterations" in lp.meta:
        assert iters <= int(np.ceil(1.1 * lp.meta["expected_iterations"]))
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    assert iters == int(fx["iterations"]) and abs(obj - float(fx["objective"])) <= 1e-9


def test_two_emulated_ranks_match_one(pkg):
    """2 MPI ranks of BASELINE configs[0]: partition + 3-piece reduce + b0 allreduce give the 1-rank result"""
    from oracle.ipm import IPM, OracleBackend
    lp = pkg.problem.c1_hier_4blocks()
    one = IPM(lp, OracleBackend(lp, n_ranks=1)).solve()
    two = IPM(lp, OracleBackend(lp, n_ranks=2)).solve()
    assert one[0] == two[0] == "SUCCESSFUL_TERMINATION"
    assert one[2] == two[2] and abs(one[1] - two[1]) <= 1e-9


@pytest.mark.parametrize("name", ["c1_hier_4blocks", "synthetic_T1"])
def test_oracle_reproduces_fixtures(pkg, name):
    from oracle.linsys import OracleRoot, dense_reference_sc
    lp = _instances(pkg)[name]()
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    assert len(fx["picks"]) >= 3
    for k in fx["picks"]:
        dd, nom = _split(lp, fx[f"k{k}_dd"], "dd"), _split(lp, fx[f"k{k}_nom"], "nom")
        ref = OracleRoot(lp)
        ref.put_diagonals(dd, nom)
        ref.factor2()
        assert _rel(np.tril(ref.SC), fx[f"k{k}_sc"]) <= 1e-12
        assert _rel(fx[f"k{k}_sc"], np.tril(dense_reference_sc(lp, dd, nom))) <= 1e-10
        assert tuple(fx[f"k{k}_inertia"]) == ref.inertia()
        if name.startswith("synthetic"):     # quasi-definite by construction (DistributedRootLinearSystem.C:55);
            # c1 has free variables / redundant rows, its SC inertia is whatever dsytrf finds
            assert ref.inertia() == (lp.root.n, lp.root.my + lp.myl + lp.mzl, 0)
        rhs = _split(lp, fx[f"k{k}_rhs"], "rhs")
        sol = ref.solve_compressed([r.copy() for r in rhs])
        assert _rel(np.concatenate(sol), fx[f"k{k}_sol"]) <= 1e-9


@pytest.mark.parametrize("cfg", ["T1", "T2"])
def test_solve_compressed_solves_the_bordered_system(pkg, cfg):
    """solveCompressed (Lsolve; Dsolve; Ltsolve) against a dense solve of the full doubly bordered KKT matrix in the
    compressed ordering [x0 y0 z0 yl zl | x_i y_i z_i ...]"""
    from oracle.linsys import OracleRoot
    P = pkg.problem
    lp = P.synthetic(cfg)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[cfg])
    ref = OracleRoot(lp)
    ref.put_diagonals(dd, nom)
    ref.factor2()
    r = lp.root
    n0, my0, mz0, myl, mzl = r.n, r.my, r.mz, lp.myl, lp.mzl
    nb0 = n0 + my0 + mz0 + myl + mzl
    sizes = [nb0] + [ch.n + ch.my + ch.mz for ch in lp.children]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    M = np.zeros((offs[-1], offs[-1]))
    # root block [dd0 A0^T C0^T F0^T G0^T; A0; C0 nom0; F0; G0 nomL]
    M[:n0, :n0] = np.diag(dd[0])
    o = n0
    for blk, m in ((r.A, my0), (r.C, mz0), (r.F, myl), (r.G, mzl)):
        if m:
            M[o:o + m, :n0] = blk.toarray()
        o += m
    M[n0 + my0:n0 + my0 + mz0, n0 + my0:n0 + my0 + mz0] = np.diag(nom[0])
    M[nb0 - mzl:nb0, nb0 - mzl:nb0] = np.diag(nom[-1])
    for i, ch in enumerate(lp.children):
        a = offs[1 + i]
        n, my, mz = ch.n, ch.my, ch.mz
        M[a:a + n, a:a + n] = np.diag(dd[1 + i])
        M[a + n:a + n + my, a:a + n] = ch.W.toarray()
        M[a + n + my:a + n + my + mz, a:a + n] = ch.D.toarray()
        M[a + n + my:a + n + my + mz, a + n + my:a + n + my + mz] = np.diag(nom[1 + i])
        M[a + n:a + n + my, :n0] = ch.A.toarray()                                  # T_i
        M[a + n + my:a + n + my + mz, :n0] = ch.C.toarray()
        M[n0 + my0 + mz0:n0 + my0 + mz0 + myl, a:a + n] = ch.F.toarray()
        M[n0 + my0 + mz0 + myl:nb0, a:a + n] = ch.G.toarray()
    # M was filled with all coupling blocks on one side only; symmetrise
    M = np.tril(M) + np.tril(M, -1).T + (np.triu(M, 1) + np.triu(M, 1).T)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(offs[-1])
    want = np.linalg.solve(M, b)
    got = ref.solve_compressed([b[offs[k]:offs[k + 1]].copy() for k in range(len(sizes))])
    assert _rel(np.concatenate(got), want) <= 1e-8


def test_c_multifrontal_oracle_matches_superlu_oracle(pkg):
    """oracle/mf_ldlt.c (the serial C restatement of the multifrontal algorithm, also the CPU baseline of bench.py)
    against the SuperLU-based oracle: Schur complement and solveCompressed on T2"""
    from oracle import mf
    from oracle.linsys import OracleRoot
    P = pkg.problem
    lp = P.synthetic("T2")
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["T2"])
    ref = OracleRoot(lp)
    ref.put_diagonals(dd, nom)
    ref.factor2()
    be = mf.MfBackend(lp, pkg.capi.Symbolic, threads=2)
    be.factorize(dd, nom)
    assert _rel(np.tril(be.root.SC), np.tril(ref.SC)) <= 1e-10
    rng = np.random.default_rng(3)
    nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
    rhs = [rng.standard_normal(nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    want = ref.solve_compressed([r.copy() for r in rhs])
    got = be.solve_compressed([r.copy() for r in rhs])
    assert _rel(np.concatenate(got), np.concatenate(want)) <= 1e-8
