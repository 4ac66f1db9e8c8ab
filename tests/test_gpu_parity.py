"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs. `-m gpu`.

Tolerances (BASELINE.json north_star): assembled Schur complement relative Frobenius error <= 1e-10, IPM iteration
count within +-1, final objective within 1e-8 relative. Solves: 1e-9 relative (the KKT systems of the synthetic
iterate have condition numbers around 1e6-1e8, two different pivot orders)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _iterate(pkg, lp, cfg):
    return pkg.problem.synthetic_iterate(lp, pkg.problem.CONFIG_ID[cfg])


@pytest.mark.parametrize("cfg", ["T1", "T2", "T3"])
def test_leaf_factor_and_solve(pkg, cfg):
    from oracle.linsys import OracleLeaf
    lp = pkg.problem.synthetic(cfg)
    dd, nom = _iterate(pkg, lp, cfg)
    ctx = pkg.capi.Context(0)
    leaves = []
    for ch in lp.children[:2]:
        krow, jcol, val = pkg.linsys.create_kkt_lower(ch)
        leaves.append(pkg.capi.Leaf(ctx, ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val))
    ctx.commit()
    rng = np.random.default_rng(0)
    for i, leaf in enumerate(leaves):
        ch = lp.children[i]
        leaf.set_diagonal(dd[1 + i], nom[1 + i])
        assert leaf.factorize() == (ch.n, ch.my + ch.mz, 0)
        assert leaf.inertia()[1] == 0          # no perturbed pivots on a well-posed leaf
        ref = OracleLeaf(lp, i)
        ref.put_diagonals(dd[1 + i], nom[1 + i])
        ref.factor2()
        B = rng.standard_normal((5, leaf.n))
        X = leaf.solve(B)
        assert _rel(X, ref.solve(B)) <= 1e-9
        x1 = leaf.solve(B[0])
        assert _rel(x1, X[0]) <= 1e-14
        K = ref.full_matrix()
        assert np.linalg.norm(K @ x1 - B[0]) / np.linalg.norm(B[0]) <= 1e-9
    ctx.close()


@pytest.mark.parametrize("cfg", ["T1", "T2", "T3", "T4"])
def test_schur_complement_and_solve_compressed(pkg, cfg):
    from oracle.linsys import OracleRoot
    lp = pkg.problem.synthetic(cfg)
    dd, nom = _iterate(pkg, lp, cfg)
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    inertia = gpu.factorize(dd, nom)
    ref = OracleRoot(lp)
    ref.put_diagonals(dd, nom)
    ref.factor2()
    assert _rel(gpu.schur_complement(), np.tril(ref.SC)) <= 1e-10
    assert inertia == ref.inertia() == (lp.root.n, lp.root.my + lp.myl + lp.mzl, 0)
    rng = np.random.default_rng(2)
    rhs = [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    want = ref.solve_compressed([r.copy() for r in rhs])
    got = gpu.solve_compressed([r.copy() for r in rhs])
    assert _rel(np.concatenate(got), np.concatenate(want)) <= 1e-8
    t = gpu.kkt.stage_times()
    assert t["leaf_ldlt_schur"] > 0 and t["root_ldlt"] > 0
    assert gpu.ctx.launch_count() > 0
    gpu.close()


def test_regularised_and_repeated_factorisation(pkg):
    """symbolic analysis is reused: factorising twice with different diagonals gives each time the oracle's result"""
    from oracle.linsys import OracleRoot
    lp = pkg.problem.synthetic("T2")
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    ref = OracleRoot(lp)
    for seed_shift, reg in ((0, None), (5, (1e-6, 1e-7, 1e-7))):
        dd, nom = pkg.problem.synthetic_iterate(lp, pkg.problem.CONFIG_ID["T2"] + seed_shift)
        gpu.factorize(dd, nom, reg)
        r = reg or (0.0, 0.0, 0.0)
        ref.put_diagonals(dd, nom, *r)
        ref.factor2()
        assert _rel(gpu.schur_complement(), np.tril(ref.SC)) <= 1e-10
    gpu.close()


@pytest.mark.parametrize("n0,nneg", [(40, 25), (200, 137), (333, 90)])
def test_dense_root_solver(pkg, n0, nneg):
    """dense LDL^T + solve of a saddle-point Schur complement vs numpy; inertia reported exactly"""
    rng = np.random.default_rng(n0)
    n = n0 + nneg
    H = rng.standard_normal((n0, n0)); H = H @ H.T + n0 * np.eye(n0)
    A = rng.standard_normal((nneg, n0))
    G = np.diag(rng.uniform(0.0, 1.0, nneg))
    S = np.block([[H, A.T], [A, -G]])
    ctx = pkg.capi.Context(0)
    sc = pkg.capi.SchurComplement(ctx, n0, 0, nneg, 0)
    sc.upload(np.tril(S))
    assert sc.factorize() == (n0, nneg, 0)
    B = rng.standard_normal((3, n))
    X = sc.solve(B)
    assert _rel(X, np.linalg.solve(S, B.T).T) <= 1e-10
    assert _rel(sc.download(), np.tril(S)) == 0.0
    sc.close()
    ctx.close()


@pytest.mark.parametrize("n", [57, 300])
def test_dense_root_solver_positive_definite(pkg, n):
    """the SOLVER_DENSE_SYM_PSD role (SOL/DensePSDSolver/DeSymPSDSolver.C:24-45: dpotrf / dpotrs): a root
    without dual rows is the same static LDL^T with n0 = n; inertia (n, 0, 0), solution against numpy"""
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    ctx = pkg.capi.Context(0)
    sc = pkg.capi.SchurComplement(ctx, n, 0, 0, 0)
    sc.upload(np.tril(S))
    assert sc.factorize() == (n, 0, 0)
    B = rng.standard_normal((2, n))
    assert _rel(sc.solve(B), np.linalg.solve(S, B.T).T) <= 1e-11
    sc.close()
    ctx.close()


def test_dense_root_solver_saddle_point(pkg):
    """the SOLVER_DENSE_SYM_INDEF_SADDLE_POINT role (SOL/DenseSymmetricIndefinitSolver/DeSymIndefSolver2.C:17-71: Cholesky of Q,
    dtrsm, dsyrk, Cholesky of A Q^-1 A^T): [Q A^T; A 0] with a zero (2,2) block is exactly what the static x0-first LDL^T
    factorises without pivoting (no Bunch-Kaufman fallback: zero perturbed pivots, inertia (nx, ny, 0))"""
    rng = np.random.default_rng(5)
    nx, ny = 180, 70
    M = rng.standard_normal((nx, nx))
    Q = M @ M.T + nx * np.eye(nx)
    A = rng.standard_normal((ny, nx))
    S = np.block([[Q, A.T], [A, np.zeros((ny, ny))]])
    ctx = pkg.capi.Context(0)
    sc = pkg.capi.SchurComplement(ctx, nx, 0, ny, 0)
    sc.upload(np.tril(S))
    assert sc.factorize() == (nx, ny, 0)
    B = rng.standard_normal((2, nx + ny))
    assert _rel(sc.solve(B), np.linalg.solve(S, B.T).T) <= 1e-10
    sc.close()
    ctx.close()


def test_sc_assembly_helpers(pkg):
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    ctx = pkg.capi.Context(0)
    sc = pkg.capi.SchurComplement(ctx, 6, 2, 3, 1)
    M = sp.random(3, 6, density=0.5, random_state=1, format="csr")
    sc.add_sparse(8, 0, M, alpha=2.0)
    d = rng.standard_normal(6)
    sc.add_diag(0, d, alpha=-1.0)
    want = np.zeros((12, 12))
    want[8:11, 0:6] += 2.0 * M.toarray()
    want[np.arange(6), np.arange(6)] -= d
    assert np.allclose(sc.download(), want, atol=1e-15)
    sc.zero()
    assert np.all(sc.download() == 0)
    sc.close()
    ctx.close()


@pytest.mark.parametrize("name", ["c1", "callback", "T1"])
def test_ipm_iteration_and_objective_parity(pkg, name):
    """full IPM runs driven by the host driver: CUDA backend vs oracle backend (golden: objective 14)"""
    from oracle.ipm import IPM, OracleBackend
    P = pkg.problem
    lp = {"c1": P.c1_hier_4blocks, "callback": P.callback_example, "T1": lambda: P.synthetic("T1")}[name]()
    ref = IPM(lp, OracleBackend(lp))
    st_r, obj_r, it_r = ref.solve()
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    run = IPM(lp, gpu)
    st_g, obj_g, it_g = run.solve()
    gpu.close()
    assert st_r == st_g == "SUCCESSFUL_TERMINATION"
    assert abs(it_g - it_r) <= 1
    assert abs(obj_g - obj_r) <= 1e-8 * max(1.0, abs(obj_r))
    if "expected_objective" in lp.meta:
        assert abs(obj_g - lp.meta["expected_objective"]) <= 1e-4      # t_pips.cpp:27 tolerance
        if "expected_iterations" in lp.meta:
            assert it_g <= int(np.ceil(1.1 * lp.meta["expected_iterations"]))


def test_full_size_block_properties(pkg):
    """C2-sized blocks (44 000 KKT rows each): size-independent checks - residual of the leaf solve through the host
    matvec, symmetry/linearity of solveCompressed, Schur complement against the C multifrontal oracle."""
    from oracle import mf
    P = pkg.problem
    lp = P.synthetic("C2", n_blocks=2)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID["C2"])
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    gpu.factorize(dd, nom)
    sc = gpu.schur_complement()
    # Schur term of block 0 and 1 from the serial C multifrontal oracle on the same symbolic structure
    n0, my0, myl, mzl = lp.root.n, lp.root.my, lp.myl, lp.mzl
    nB = n0 + myl + mzl
    SCb = np.zeros((nB, nB))
    for i, ch in enumerate(lp.children):
        krow, jcol, val, diag_pos, nB_, brow, bcol, bval = mf.leaf_kkt_and_border(lp, i)
        val = val.copy()
        val[diag_pos[:ch.n]] = dd[1 + i]
        val[diag_pos[ch.n + ch.my:]] = nom[1 + i]
        sym = pkg.capi.Symbolic(ch.n + ch.my + ch.mz, ch.n, ch.my, ch.mz, krow, jcol, val, nB, brow, bcol).export()
        m = mf.MfLeaf(sym)
        inertia, npert = m.factorize(val, bval, SCb)
        assert inertia == (ch.n, ch.my + ch.mz, 0) and npert == 0
    idx = np.concatenate([np.arange(n0), n0 + my0 + np.arange(myl + mzl)])
    # remove the root's own block from the GPU matrix to compare the children's contribution only
    from oracle.linsys import OracleRoot
    root_only = OracleRoot(lp, owned=[])
    root_only.put_diagonals(dd, nom)
    root_only.initialize_kkt()
    root_only.finalize_kkt_dense()
    contrib = (sc - np.tril(root_only.SC))[np.ix_(idx, idx)]
    assert _rel(contrib, np.tril(SCb)) <= 1e-10
    # solveCompressed: linearity
    rng = np.random.default_rng(4)
    mk = lambda: [rng.standard_normal(gpu.nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    a, b = mk(), mk()
    sa = gpu.solve_compressed([v.copy() for v in a])
    sb = gpu.solve_compressed([v.copy() for v in b])
    sab = gpu.solve_compressed([2.0 * u - 3.0 * v for u, v in zip(a, b)])
    lin = np.concatenate([2.0 * u - 3.0 * v for u, v in zip(sa, sb)])
    assert _rel(np.concatenate(sab), lin) <= 1e-9
    gpu.close()


@pytest.mark.parametrize("kind", ["saddle_free", "random_indefinite", "c1_like_singular"])
def test_dense_root_solver_bunch_kaufman(pkg, kind):
    """matrices the static (x0-first) order cannot factor: the solver must fall back to Bunch-Kaufman pivoting (what
    the reference's dsytrf does, DeSymIndefSolver.C:76-90) and still match numpy / the oracle's LAPACK dsytrf inertia"""
    from scipy.linalg import lapack
    rng = np.random.default_rng(7)
    if kind == "saddle_free":
        n0 = m = 24
        A = rng.standard_normal((m, n0)) + 3 * np.eye(m)
        S = np.block([[np.zeros((n0, n0)), A.T], [A, np.zeros((m, m))]])
    elif kind == "random_indefinite":
        n0, m = 70, 80
        S = rng.standard_normal((n0 + m, n0 + m)); S = S + S.T
    else:
        n0, m = 6, 12
        A = rng.standard_normal((m, n0))
        S = np.block([[np.zeros((n0, n0)), A.T], [A, -np.diag(rng.uniform(0.5, 1.5, m))]])
    n = n0 + m
    ctx = pkg.capi.Context(0)
    sc = pkg.capi.SchurComplement(ctx, n0, 0, m, 0)
    sc.upload(np.tril(S))
    inertia = sc.factorize()
    ev = np.linalg.eigvalsh(S)
    assert inertia == (int((ev > 0).sum()), int((ev < 0).sum()), 0)
    B = rng.standard_normal((2, n))
    X = sc.solve(B)
    assert _rel(X, np.linalg.solve(S, B.T).T) <= 1e-9
    sc.close()
    ctx.close()


@pytest.mark.parametrize("name", ["c1_hier_4blocks", "synthetic_T1"])
def test_against_committed_golden_fixtures(pkg, name):
    """CUDA path vs tests/golden/*.npz (oracle output pinned against the reference's end-to-end goldens and a dense
    numpy evaluation by tests/golden/make_golden.py): Schur complement <= 1e-10 rel. Frobenius on the recorded IPM
    iterates, solveCompressed <= 1e-8."""
    import os
    from test_oracle_golden import GOLD, _split
    P = pkg.problem
    lp = {"c1_hier_4blocks": P.c1_hier_4blocks, "synthetic_T1": lambda: P.synthetic("T1")}[name]()
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    gpu = pkg.linsys.CudaRootLinearSystem(lp)
    for k in fx["picks"]:
        dd, nom = _split(lp, fx[f"k{k}_dd"], "dd"), _split(lp, fx[f"k{k}_nom"], "nom")
        inertia = gpu.factorize(dd, nom)
        assert _rel(gpu.schur_complement(), fx[f"k{k}_sc"]) <= 1e-10
        if name.startswith("synthetic"):
            assert inertia == tuple(fx[f"k{k}_inertia"])
        else:   # c1: free variables / redundant rows give the SC an (almost) zero eigenvalue whose sign is rounding noise
            assert sum(inertia) == lp.n_S
        if not name.startswith("synthetic"):
            continue   # c1's SC is numerically singular: a random right-hand side has no well-defined solution
        rhs = _split(lp, fx[f"k{k}_rhs"], "rhs")
        got = gpu.solve_compressed([r.copy() for r in rhs])
        assert _rel(np.concatenate(got), fx[f"k{k}_sol"]) <= 1e-8
    gpu.close()


def _two_gpu_lp(P, cfg, block_range=None):
    """cfg 'S1-local' = S1 with local linking structure (sparse root Schur complement)"""
    if cfg == "S1-local":
        return P.localize_links(P.synthetic("S1", block_range=block_range), halo=0.5), "S1"
    return P.synthetic(cfg, block_range=block_range), cfg


def _two_gpu_worker(rank, world, port, cfg, q):
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)    # only ships the NCCL unique id
    try:
        pkg = g.load_package()
        P = pkg.problem
        obj = [pkg.capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nblocks = P.CONFIGS["S1" if cfg == "S1-local" else cfg][0]
        lo, hi = pkg.capi.partition(nblocks, world)[rank]
        lp, base = _two_gpu_lp(P, cfg, (lo, hi))
        dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[base])
        pattern = None
        if cfg == "S1-local":       # the pattern is global: every rank derives it from the whole problem structure
            pattern = pkg.linsys.sc_lower_pattern(_two_gpu_lp(P, cfg)[0])
        sys_ = pkg.linsys.CudaRootLinearSystem(lp, device=rank, rank=rank, n_ranks=world, nccl_uid=obj[0], sc_pattern=pattern)
        inertia = sys_.factorize(dd, nom)
        sparse_used = sys_.kkt.sc().uses_sparse_solver()
        sc = sys_.schur_complement()
        rng = np.random.default_rng(11)
        b0 = rng.standard_normal(sys_.nb0)
        nK = lp.children[0].n + lp.children[0].my + lp.children[0].mz
        ball = rng.standard_normal((nblocks, nK))
        # ranks > 0 pass garbage in b0: the reference ignores it (sLinsysRootAug.C:332-333)
        rhs = [b0.copy() if rank == 0 else np.full(sys_.nb0, 1e30)] + [ball[i].copy() for i in range(lo, hi)]
        sol = sys_.solve_compressed(rhs)
        # device BiCGStab over both ranks (dot products all-reduced): the root part of the vectors is replicated
        ksol = [b0.copy()] + [ball[i].copy() for i in range(lo, hi)]
        kstat = sys_.solve_bicgstab(ksol, tol=1e-10)
        q.put((rank, lo, hi, inertia, sc, sol[0], np.stack(sol[1:]), sparse_used, kstat, ksol[0], np.stack(ksol[1:])))
        dist.barrier()
        sys_.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", ["T2", "T3", "S1-local"])
def test_two_gpus_nccl_allreduce_equals_single_rank_oracle(pkg, cfg):
    """blocks partitioned over 2 GPUs as mapChildrenToNSubTrees does over 2 MPI ranks; reduceKKT (dense pack, or the
    pattern values of the sparse root for 'S1-local'), the Lsolve b0 reduction and the dot products of the device BiCGStab
    run as ncclAllReduce. Needs 2 visible GPUs (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle.linsys import OracleRoot
    from test_multirank_gloo import _free_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    P = pkg.problem
    lp, base = _two_gpu_lp(P, cfg)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[base])
    ref = OracleRoot(lp)
    ref.put_diagonals(dd, nom)
    ref.factor2()
    rng = np.random.default_rng(11)
    b0 = rng.standard_normal(res[0][5].shape[0])
    nK = lp.children[0].n + lp.children[0].my + lp.children[0].mz
    ball = rng.standard_normal((len(lp.children), nK))
    want = ref.solve_compressed([b0.copy()] + [ball[i].copy() for i in range(len(lp.children))])
    for rank, lo, hi, inertia, sc, sol0, soll, sparse_used, kstat, ksol0, ksoll in res:
        assert _rel(sc, np.tril(ref.SC)) <= 1e-10
        assert inertia == ref.inertia()
        assert sparse_used == (cfg == "S1-local")
        assert _rel(sol0, want[0]) <= 1e-8
        assert _rel(soll, np.stack(want[1 + lo:1 + hi])) <= 1e-8
        assert kstat[0] in (0, 1) and kstat[2] <= 1e-10
        assert _rel(ksol0, want[0]) <= 1e-7
        assert _rel(ksoll, np.stack(want[1 + lo:1 + hi])) <= 1e-7
    # both ranks took the same Krylov path (the residual norms agree to rounding: the replicated root parts are updated
    # with rank-local atomics)
    assert res[0][8][:2] == res[1][8][:2] and abs(res[0][8][2] - res[1][8][2]) <= 1e-6 * max(res[0][8][2], 1e-300)


@pytest.mark.parametrize("cfg", ["T1", "T3"])
def test_factor_entries_match_c_oracle(pkg, cfg):
    """factor-level parity: the L panels and D^-1 downloaded from the device against oracle/mf_ldlt.c on the SAME
    symbolic structure (identical layout), front by front. 1e-9 relative to the largest factor entry (the device sums
    in DMMA-tile order, the oracle left to right)."""
    from oracle import mf
    P = pkg.problem
    lp = P.synthetic(cfg)
    dd, nom = _iterate(pkg, lp, cfg)
    ctx = pkg.capi.Context(0)
    ch = lp.children[0]
    n = ch.n + ch.my + ch.mz
    krow, jcol, val = pkg.linsys.create_kkt_lower(ch)
    leaf = pkg.capi.Leaf(ctx, n, ch.n, ch.my, ch.mz, krow, jcol, val)
    ctx.commit()
    leaf.set_diagonal(dd[1], nom[1])
    assert leaf.factorize() == (ch.n, ch.my + ch.mz, 0)
    Lg, dg, eg = leaf.download_factor()
    krow2, jcol2, val2, diag_pos, nB, brow, bcol, bval = mf.leaf_kkt_and_border(lp, 0)
    val2 = val2.copy()
    val2[diag_pos[:ch.n]] = dd[1]
    val2[diag_pos[ch.n + ch.my:]] = nom[1]
    none = np.zeros(0, np.int32)
    sym = pkg.capi.Symbolic(n, ch.n, ch.my, ch.mz, krow2, jcol2, val2, 0, none, none).export()
    ref = mf.MfLeaf(sym)
    assert ref.factorize(val2, np.zeros(0)) == ((ch.n, ch.my + ch.mz, 0), 0)
    assert len(Lg) == int(sym["Loff"][-1])
    scale = np.abs(ref.L[:len(Lg)]).max()
    assert np.abs(Lg - ref.L[:len(Lg)]).max() <= 1e-9 * scale
    assert np.abs(dg - ref.dinv_d[:n]).max() <= 1e-9 * np.abs(ref.dinv_d[:n]).max()
    assert np.abs(eg - ref.dinv_e[:n]).max() <= 1e-9 * max(np.abs(ref.dinv_e[:n]).max(), 1e-300)
    ctx.close()
