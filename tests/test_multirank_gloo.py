"""CPU test of the N>1 path (world_size 2, gloo): every process owns the contiguous block range the reference's
mapChildrenToNSubTrees gives its rank (through the C-ABI pipsgpu_partition), builds only those blocks, accumulates
its Schur terms, and the packed three-piece reduce of reduceKKTdense (LS/DistributedRootLinearSystem.C:849-881) and the
b0 all-reduce of Lsolve (LS/sLinsysRootAug.C:332-341) go through torch.distributed - the same protocol the CUDA path
runs over NCCL. The result must equal the single-rank oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(SC, n0, my0, nl):
    """lower(n0) | rect nl x n0 | lower(nl) - the layout of sc_pack_kernel (csrc/kernels.cu)"""
    il = np.tril_indices(n0)
    a = SC[:n0, :n0][il]
    o = n0 + my0
    b = SC[o:o + nl, :n0].ravel()
    jl = np.tril_indices(nl)
    c = SC[o:o + nl, o:o + nl][jl]
    return np.concatenate([a, b, c])


def _unpack(buf, SC, n0, my0, nl):
    il = np.tril_indices(n0)
    k = len(il[0])
    SC[:n0, :n0][il] = buf[:k]
    o = n0 + my0
    SC[o:o + nl, :n0] = buf[k:k + nl * n0].reshape(nl, n0)
    jl = np.tril_indices(nl)
    SC[o:o + nl, o:o + nl][jl] = buf[k + nl * n0:]


def _worker(rank, world, port, cfg, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from oracle.linsys import OracleRoot
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = g.load_package()
        P = pkg.problem
        nblocks = P.CONFIGS[cfg][0]
        lo, hi = pkg.capi.partition(nblocks, world)[rank]
        lp = P.synthetic(cfg, block_range=(lo, hi))          # only this rank's blocks are built
        dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[cfg])
        root = OracleRoot(lp, rank=rank)
        root.put_diagonals(dd, nom)
        root.initialize_kkt()
        root.assemble_local_kkt()
        n0, my0, nl = lp.root.n, lp.root.my, lp.myl + lp.mzl
        buf = torch.from_numpy(_pack(np.tril(root.SC), n0, my0, nl))
        dist.all_reduce(buf)                                  # reduceKKT
        root.SC[:] = 0.0
        _unpack(buf.numpy(), root.SC, n0, my0, nl)
        root.SC[:] = np.tril(root.SC) + np.tril(root.SC, -1).T
        root.finalize_kkt_dense()
        root.factorize_kkt()
        # solveCompressed: same global right-hand side on every rank, each rank holds its own children's parts
        rng = np.random.default_rng(11)
        nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
        b0 = rng.standard_normal(nb0)
        nK = lp.children[0].n + lp.children[0].my + lp.children[0].mz
        ball = rng.standard_normal((nblocks, nK))
        b = [b0.copy() if rank == 0 else np.zeros(nb0)] + [ball[i].copy() for i in range(lo, hi)]
        root.lsolve_local(b)
        t = torch.from_numpy(b[0])
        dist.all_reduce(t)                                    # Lsolve's b0 all-reduce
        root.solve_reduced_link_cons(b[0])
        root.ltsolve(b)
        q.put((rank, lo, hi, np.tril(root.SC), b[0], np.stack(b[1:])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", ["T1", "T2"])
def test_two_ranks_gloo_equal_single_rank(pkg, cfg):
    import torch.multiprocessing as mp
    from oracle.linsys import OracleRoot
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = pkg.problem
    lp = P.synthetic(cfg)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[cfg])
    ref = OracleRoot(lp)
    ref.put_diagonals(dd, nom)
    ref.factor2()
    rng = np.random.default_rng(11)
    nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
    b0 = rng.standard_normal(nb0)
    nK = lp.children[0].n + lp.children[0].my + lp.children[0].mz
    ball = rng.standard_normal((len(lp.children), nK))
    want = ref.solve_compressed([b0.copy()] + [ball[i].copy() for i in range(len(lp.children))])
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == len(lp.children)
    for rank, lo, hi, sc, sol0, soll in res:
        assert rel(sc, np.tril(ref.SC)) <= 1e-12
        assert rel(sol0, want[0]) <= 1e-9
        assert rel(soll, np.stack(want[1 + lo:1 + hi])) <= 1e-9
