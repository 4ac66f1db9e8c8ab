"""Outer-solve rows of SURVEY.md 8 on the GPU (`-m gpu`): device BiCGStab + full-KKT matvec (a17 / f1), iterative
refinement of the leaf solves (a16), truthful inertia + the regularisation loop (a4 / f4)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _hard_iterate(lp, seed, lo=-6.0, hi=6.0):
    """late-IPM-like diagonals: 12 orders of magnitude of spread"""
    rng = np.random.default_rng(seed)
    nodes = [lp.root] + lp.children
    dd = [10.0 ** rng.uniform(lo, hi, nd.n) for nd in nodes]
    nom = [-(10.0 ** rng.uniform(lo, hi, nd.mz)) for nd in nodes] + [-(10.0 ** rng.uniform(lo, hi, lp.mzl))]
    return dd, nom


@pytest.mark.parametrize("cfg", ["T2", "T3"])
def test_device_bicgstab_against_oracle_and_host_matvec(pkg, cfg):
    """pipsgpu_kkt_solve = LinearSystem::solveCompressedBiCGStab on the device: the reported relative residual must
    equal the residual evaluated on the host from the problem data (checks the device system_mult), and the solution
    must agree with the oracle's restatement of the same Krylov loop (oracle/ipm.py:LinearSystem._bicgstab) preconditioned
    by the oracle's solveCompressed."""
    from oracle.ipm import LinearSystem, OracleBackend
    P = pkg.problem
    lp = P.synthetic(cfg)
    dd, nom = _hard_iterate(lp, 7)
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    gpu.factorize(dd, nom)
    rng = np.random.default_rng(3)
    rhs = [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    sol = [r.copy() for r in rhs]
    status, iters, relres = gpu.solve_bicgstab(sol, tol=1e-10)
    assert status in (0, 1), (status, iters, relres)          # CONVERGED or SKIPPED
    lsq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol)
    host_rel = np.sqrt(lsq + r0 @ r0) / np.sqrt(sum(b @ b for b in rhs))
    assert host_rel <= 2e-10
    assert abs(host_rel - relres) <= 1e-11 + 0.05 * host_rel
    # a plain solveCompressed on the same factors must not be more accurate than the Krylov-refined solution
    plain = gpu.solve_compressed([r.copy() for r in rhs])
    lsq1, r1 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, plain)
    assert host_rel <= 1.0001 * np.sqrt(lsq1 + r1 @ r1) / np.sqrt(sum(b @ b for b in rhs)) + 1e-16
    # oracle: same loop, oracle preconditioner
    ls = LinearSystem(lp, OracleBackend(lp))
    ls.dd = np.concatenate(dd)
    ls.nomegainv = np.concatenate(nom)
    ls.backend.factorize(dd, nom)
    bx, by, bz = ls.separate_vars(rhs)
    ox, oy, oz = ls._bicgstab(bx, by, bz)
    want = ls.join_rhs(ox, oy, oz)
    assert ls.bicg_converged
    assert _rel(np.concatenate(sol), np.concatenate(want)) <= 1e-7
    assert abs(iters - (0 if ls.bicg_skipped else ls.bicg_niter)) <= 1
    gpu.close()


def test_bicgstab_iterates_when_the_preconditioner_is_stale(pkg):
    """the Krylov loop must actually iterate (not only be skipped): precondition with the factorisation of a slightly
    different iterate by re-uploading new diagonals without refactorising is not possible through the ABI, so the
    inexactness comes from forced pivot perturbations (tiny primal diagonals), exactly the situation of the reference"""
    P = pkg.problem
    lp = P.synthetic("T2")
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["T2"])
    dd = [d.copy() for d in dd]
    for d in dd[1:]:
        d[::7] = 1e-11                      # below the static-pivot threshold 1e-8 * ||K||: perturbed pivots
    os.environ["PIPSGPU_REFINE"] = "0"      # the leaf solves stay inexact: the outer loop has to do the work
    try:
        gpu = pkg.linsys.CudaRootLinearSystem(lp)
    finally:
        del os.environ["PIPSGPU_REFINE"]
    gpu.factorize(dd, nom)
    assert sum(lf.inertia()[1] for lf in gpu.leaves) > 0          # perturbed pivots were reported
    rng = np.random.default_rng(5)
    rhs = [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    plain = gpu.solve_compressed([r.copy() for r in rhs])
    lsq0, r00 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, plain)
    res_plain = np.sqrt(lsq0 + r00 @ r00) / np.sqrt(sum(b @ b for b in rhs))
    sol = [r.copy() for r in rhs]
    status, iters, relres = gpu.solve_bicgstab(sol, tol=1e-10, max_iter=75)
    lsq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol)
    res = np.sqrt(lsq + r0 @ r0) / np.sqrt(sum(b @ b for b in rhs))
    assert res_plain > 1e-9                 # the perturbed factorisation alone is not accurate enough ...
    assert iters >= 1 and status == 0       # ... so BiCGStab iterated and converged
    assert res <= 2e-10 and res < 1e-2 * res_plain
    gpu.close()


def test_iterative_refinement_of_leaf_solves(pkg):
    """perturbed pivots: with refinement (default, PARDISO iparm[7]=2 analogue) solveCompressed and the plug-in multi-RHS
    leaf solve recover the accuracy that the unrefined solves lose"""
    P = pkg.problem
    lp = P.synthetic("T2")
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["T2"])
    dd = [d.copy() for d in dd]
    for d in dd[1:]:
        d[::7] = 1e-11
    rng = np.random.default_rng(9)
    rhs = None
    out = {}
    for steps in ("0", "2"):
        os.environ["PIPSGPU_REFINE"] = steps
        try:
            gpu = pkg.linsys.CudaRootLinearSystem(lp)
        finally:
            del os.environ["PIPSGPU_REFINE"]
        gpu.factorize(dd, nom)
        if rhs is None:
            rhs = [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
        sol = gpu.solve_compressed([r.copy() for r in rhs])
        lsq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol)
        # plug-in level: K_0 X = B with 3 right-hand sides, residual through the host matrix
        ch = lp.children[0]
        n = ch.n + ch.my + ch.mz
        B = rng.standard_normal((3, n)) if "B" not in out else out["B"]
        out["B"] = B
        X = gpu.leaves[0].solve(B)
        import scipy.sparse as sp
        K = sp.bmat([[sp.diags(dd[1]), ch.W.T, ch.D.T], [ch.W, None, None], [ch.D, None, sp.diags(nom[1])]], format="csr")
        out[steps] = (np.sqrt(lsq + r0 @ r0) / np.sqrt(sum(b @ b for b in rhs)),
                      np.linalg.norm(K @ X.T - B.T) / np.linalg.norm(B))
        gpu.close()
    assert out["0"][0] > 1e-9 and out["0"][1] > 1e-9          # static pivoting + perturbation alone is inexact
    assert out["2"][0] < 1e-3 * out["0"][0]
    assert out["2"][1] < 1e-3 * out["0"][1]


def test_wrong_inertia_is_reported_and_regularised_away(pkg):
    """a leaf whose equality block is rank deficient (duplicated row) has the wrong inertia without regularisation: the
    device must REPORT it (zero / wrongly signed computed pivot), and the reference's per-node loop
    LinearSystem::factorize_with_correct_inertia with the IPOPT strategy must end with the same regularisation and the
    same number of factorisations as its restatement on the C oracle."""
    from oracle.linsys import oracle_leaf_inertia_loop
    P = pkg.problem
    lp = P.synthetic("T2")
    bad = 2
    ch = lp.children[bad]
    W = ch.W.tolil()
    W[1, :] = W[0, :]                       # rows 0 and 1 of the equality block coincide: K is singular
    ch.W = P._csr(W.tocsr(), W.shape)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["T2"])
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    gpu.factorize(dd, nom)
    expected = [(c.n, c.my + c.mz, 0) for c in lp.children]
    got = [lf.inertia()[0] for lf in gpu.leaves]
    assert [g == e for g, e in zip(got, expected)] == [k != bad for k in range(len(expected))]
    R = pkg.regularization
    strategies = [R.IpoptRegularization(c.n, c.my + c.mz) for c in lp.children]
    root_inertia, reg, rounds = gpu.factorize_with_correct_inertia(dd, nom, strategies, barrier_parameter=1e-2)
    assert [lf.inertia()[0] for lf in gpu.leaves] == expected
    want_reg, want_nfact, seen = oracle_leaf_inertia_loop(lp, bad, dd[1 + bad], nom[1 + bad], pkg.capi.Symbolic,
                                                         barrier_parameter=1e-2)
    assert seen[0] == got[bad]              # the oracle sees the same wrong inertia without regularisation
    assert rounds == want_nfact
    assert np.allclose(reg[bad], want_reg, rtol=0, atol=0)
    assert all(r == (0.0, 0.0, 0.0) for k, r in enumerate(reg) if k != bad)
    # the regularised system is solvable: residual of solveCompressed against the regularised matrix of that leaf
    rng = np.random.default_rng(1)
    rhs = [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(c.n + c.my + c.mz) for c in lp.children]
    X = gpu.leaves[bad].solve(rhs[1 + bad])
    import scipy.sparse as sp
    rp, ry, rz = reg[bad]
    K = sp.bmat([[sp.diags(dd[1 + bad] + rp), ch.W.T, ch.D.T], [ch.W, -ry * sp.eye(ch.my), None],
                 [ch.D, None, sp.diags(nom[1 + bad] - rz)]], format="csr")
    assert np.linalg.norm(K @ X - rhs[1 + bad]) / np.linalg.norm(rhs[1 + bad]) <= 1e-8
    gpu.close()
