#!/usr/bin/env python
"""bench.py - seconds per IPM iteration spent in the KKT factor + solve (BASELINE.json metric) on N B200s.

One "step" = the KKT work of one interior-point iteration on a fixed synthetic iterate:
    1 x factorize        (diagonals down, leaf LDL^T + Schur terms, SC all-reduce, finalize, root LDL^T)
  + S x solveCompressed  (Lsolve, Dsolve, Ltsolve; S = --solves, default 8 = predictor + corrector + 2 Gondzio
                          correctors with on average one BiCGStab iteration, LinearSystem.C:550-798)

Workload (default): BASELINE.json's headline configuration "C3" = energy-system-shaped LP, 512 blocks x 110 000 KKT
rows, n_S = 2 700, STRONG scaling: the 512 blocks are split over the N GPUs exactly like the reference's
mapChildrenToNSubTrees splits them over MPI ranks; the Schur complement and the b0 vector are all-reduced over NCCL.
`--config C2|C4|C5` selects the other BASELINE configurations (C2 defaults to weak scaling, 64 blocks per GPU, the
round-1 line); `--blocks-per-gpu` forces weak scaling.

  value : device-resident step (diagonals / right-hand sides already in HBM), CUDA events on the launching stream
  e2e   : the same step through the C-ABI calls with pinned HOST buffers (H2D of diagonals and right-hand sides and D2H
          of the solutions inside the timed region)
  check : before timing, one solveCompressed result is verified through the host KKT matvec (residual of the full
          bordered system, reduced over the ranks) and the all-reduced Schur complement is compared across ranks; the
          run FAILS if either is off
  --impl reference : the CPU path (oracle/: C multifrontal LDL^T + Schur terms + LAPACK dsytrf root, all host threads)
                     on a bounded sample of the same workload's blocks
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's own log (version banner, NCCL_DEBUG=INFO) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "sec per IPM iteration (KKT factor+solve)"
UNIT = "s/iteration"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--blocks-per-gpu", type=int, default=0,
                    help="weak scaling with this many blocks per GPU; 0 = the scaling mode's default")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="strong: the config's block count is split over the GPUs (pipsgpu_partition = "
                         "mapChildrenToNSubTrees); weak: blocks-per-gpu blocks on every GPU; auto: C2 weak, others strong")
    ap.add_argument("--total-blocks", type=int, default=0, help="strong scaling: override the config's block count")
    ap.add_argument("--solves", type=int, default=8, help="solveCompressed calls per IPM iteration")
    ap.add_argument("--cpu-sample-blocks", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--gen-procs", type=int, default=0, help="host processes that build the blocks (0 = cores / ranks)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_cores(local_rank, world):
    """NUMA placement of one rank (what an MPI launcher's --bind-to does for the reference): restrict the process to the
    cores `nvidia-smi topo -m` lists as CPU affinity of its GPU, and inside that set to the rank's share among the GPUs with
    the same affinity. Pinned staging buffers are then first-touched on the GPU's own NUMA node. Best effort: any parsing
    problem leaves the affinity untouched. Returns a short description for the bench line."""
    import re
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        aff = {}
        for line in out.splitlines():
            m = re.match(r"^GPU(\d+)\s", line)
            if not m:
                continue
            toks = [t for t in line.split() if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t)]
            if not toks:
                continue
            cores = set()
            for part in toks[0].split(","):      # first purely numeric range list after the link columns = CPU affinity
                a, _, b = part.partition("-")
                cores.update(range(int(a), int(b or a) + 1))
            aff[int(m.group(1))] = sorted(cores)
        mine = aff.get(local_rank)
        if not mine:
            return "unbound (no CPU affinity reported)"
        avail = sorted(set(mine) & set(os.sched_getaffinity(0)))
        peers = sorted(g for g, c in aff.items() if c == mine and g < world)
        if len(avail) >= 2 * len(peers) and local_rank in peers:
            k = peers.index(local_rank)
            share = len(avail) // len(peers)
            avail = avail[k * share:(k + 1) * share]
        if not avail:
            return "unbound (empty core set)"
        os.sched_setaffinity(0, avail)
        global _BOUND
        _BOUND = True
        return f"cores {avail[0]}-{avail[-1]} ({len(avail)}) = CPU affinity of GPU {local_rank}"
    except Exception as e:      # noqa: BLE001
        return f"unbound ({type(e).__name__})"


_BOUND = False


def rank_cores(world):
    """host cores this rank may use: its bound set, or an even share of the box"""
    if _BOUND:
        return max(1, len(os.sched_getaffinity(0)))
    return max(1, (os.cpu_count() or 1) // max(world, 1))


def scaling_mode(args):
    if args.scaling != "auto":
        return args.scaling
    if args.blocks_per_gpu:
        return "weak"
    return "weak" if args.config == "C2" else "strong"


def total_blocks_of(args, world, pkg):
    base = pkg.problem.CONFIGS[args.config][0]
    if scaling_mode(args) == "weak":
        return (args.blocks_per_gpu or base) * world
    return args.total_blocks or base


def workload_string(args, pkg, total, world):
    N, my, mz, n, n0, my0, mz0, myl, mzl = pkg.problem.CONFIGS[args.config]
    return (f"{args.config}: {total} blocks x {n + my + mz} KKT rows, n_S={n0 + my0 + myl + mzl}, {scaling_mode(args)} scaling "
            f"over {world} GPU(s), fixed synthetic iterate (SURVEY 8d)")


def _gen_chunk(job):
    """worker: build the blocks [a, b) of the synthetic LP and their lower-CSR KKT matrices (host preparation only)"""
    cfg, total, a, b = job
    pkg = graft.load_package()
    lp = pkg.problem.synthetic(cfg, n_blocks=total, block_range=(a, b))
    kkts = [pkg.linsys.create_kkt_lower(ch) for ch in lp.children]
    return a, lp, kkts


def build_problem(pkg, args, rank, world, total):
    """the blocks this rank owns (pipsgpu_partition = mapChildrenToNSubTrees), built by a pool of host processes"""
    import multiprocessing as mp
    P = pkg.problem
    parts = pkg.capi.partition(total, world)
    lo, hi = parts[rank]
    nproc = args.gen_procs or rank_cores(world)
    nproc = max(1, min(nproc, hi - lo))
    if nproc == 1:
        lp = P.synthetic(args.config, n_blocks=total, block_range=(lo, hi))
        kkts = [pkg.linsys.create_kkt_lower(ch) for ch in lp.children]
    else:
        step = max(1, (hi - lo + 4 * nproc - 1) // (4 * nproc))
        jobs = [(args.config, total, a, min(a + step, hi)) for a in range(lo, hi, step)]
        with mp.get_context("fork").Pool(nproc) as pool:
            res = sorted(pool.map(_gen_chunk, jobs), key=lambda t: t[0])
        lp = res[0][1]
        kkts = list(res[0][2])
        for _, part, kk in res[1:]:
            lp.children.extend(part.children)
            kkts.extend(kk)
        lp.meta["block_range"] = (lo, hi)
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[args.config])
    return lp, kkts, dd, nom, (lo, hi)


def canonical_counts(lp, leaf_stats, solves):
    """SURVEY.md 8(d) algorithmic flops / bytes of this rank's share (per step)."""
    F1 = sum(s["flops"] for s in leaf_stats)
    nnzL = sum(s["nnzL"] for s in leaf_stats)
    eta = sum(s["nnzL"] - (ch.n + ch.my + ch.mz) for s, ch in zip(leaf_stats, lp.children))
    nnzK = sum(ch.W.nnz + ch.D.nnz + ch.n + ch.my + ch.mz for ch in lp.children)
    B1 = 16.0 * nnzL + 12.0 * nnzK
    F2 = B2 = 0.0
    for s, ch in zip(leaf_stats, lp.children):
        nK = ch.n + ch.my + ch.mz
        cols = set(ch.A.indices.tolist()) | set(ch.C.indices.tolist())
        r_i = len(cols) + int((np.diff(ch.F.indptr) > 0).sum()) + int((np.diff(ch.G.indptr) > 0).sum())
        nnzB = ch.A.nnz + ch.C.nnz + ch.F.nnz + ch.G.nnz
        F2 += r_i * (4.0 * (s["nnzL"] - nK) + nK + 2.0 * nnzB)
        B2 += 24.0 * s["nnzL"] + 16.0 * nK * r_i          # r_tile = r_i: one sweep handles every border column
    nS = lp.n_S
    F3 = nS ** 3 / 3.0 + 2.0 * nS ** 2
    B3 = 16.0 * nS ** 2
    nnzBord = sum(ch.A.nnz + ch.C.nnz + ch.F.nnz + ch.G.nnz for ch in lp.children)
    nK_tot = sum(ch.n + ch.my + ch.mz for ch in lp.children)
    F4 = 2.0 * (4.0 * eta + nK_tot) + 4.0 * nnzBord + 2.0 * nS ** 2
    B4 = 48.0 * nnzL + 8.0 * nS ** 2
    return dict(F1=F1, B1=B1, F2=F2, B2=B2, F3=F3, B3=B3, F4=F4 * solves, B4=B4 * solves)


def cpu_reference(pkg, args, total_blocks, sample_blocks, steps, warmup):
    """The CPU path (oracle/mf.py MfBackend: C multifrontal LDL^T + Schur terms per block on all host threads, LAPACK
    dsytrf / dsytrs root) timed on a bounded sample of the workload's blocks. The leaf part scales with the number of
    blocks and is extrapolated to the full block count; the root part (dense n_S x n_S) does not and is added unscaled.
    The symbolic structure comes from the host-only build of the analysis (oracle/_build/libpipssym.so): the CUDA
    library is not loaded into this process."""
    from oracle import mf
    P = pkg.problem
    cores = os.cpu_count() or 1
    sample = sample_blocks or min(total_blocks, max(cores, 4))
    sample = max(1, min(sample, total_blocks))
    lp = P.synthetic(args.config, n_blocks=total_blocks, block_range=(0, sample))
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[args.config])
    threads = min(cores, sample)
    be = mf.MfBackend(lp, mf.HostSymbolic, threads=threads)
    rng = np.random.default_rng(0)
    nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
    rhs = [rng.standard_normal(nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in lp.children]
    acc = []
    for it in range(warmup + steps):
        be.factorize(dd, nom)
        b = [r.copy() for r in rhs]
        be.solve_compressed(b)
        if it >= warmup:
            acc.append((be.t_leaf_factor, be.t_root_factor, be.t_leaf_solve, be.t_root_solve))
    tlf, trf, tls, trs = (float(np.mean([a[k] for a in acc])) for k in range(4))
    scale = total_blocks / sample
    value = (tlf + args.solves * tls) * scale + (trf + args.solves * trs)
    return dict(value=value, unit=UNIT, cores=threads, kind="port",
                sample=f"{sample} of {total_blocks} blocks measured per step (1 factorize + 1 solveCompressed); leaf part x{scale:.1f}, "
                       f"dense root part unscaled, {args.solves} solves/iteration; C multifrontal LDL^T + Schur terms "
                       f"(oracle/mf_ldlt.c, front updates {'through OpenBLAS dgemm (BLAS-3)' if mf.uses_blas() else 'scalar'}) "
                       f"+ LAPACK dsytrf root, {threads} host threads",
                t_leaf_factor_sample=tlf, t_root_factor=trf, t_leaf_solve_sample=tls, t_root_solve=trs,
                measured_seconds_per_step=tlf + trf + tls + trs, sample_blocks=sample, sample_scale=scale)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        args.gpus = world
    if args.gpus > 1 and world == 1:
        print(json.dumps({"error": "launch with torch.distributed.run for --gpus > 1"}))
        return 2
    graft.build()
    pkg = graft.load_package()
    binding = bind_to_gpu_cores(local_rank, world) if (world > 1 and args.impl == "ours") else "unbound (single rank)"
    total_blocks = total_blocks_of(args, world, pkg)
    workload = workload_string(args, pkg, total_blocks, world)

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = cpu_reference(pkg, args, total_blocks, args.cpu_sample_blocks, args.steps, args.warmup)
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["value"] * 1e3, "higher_is_better": False,
                "scaling": scaling_mode(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "solves_per_iteration": args.solves, "cpu_sample": res["sample"]},
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------
    # host preparation first (forks worker processes: must happen before CUDA is initialised in this process)
    t_gen = time.perf_counter()
    lp, kkts, dd, nom, (lo, hi) = build_problem(pkg, args, rank, world, total_blocks)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        obj = [pkg.capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]

    stream = torch.cuda.Stream()
    t_setup = time.perf_counter()
    # host threads of the (one-off) symbolic phase: the ranks of one box share its cores
    sym_threads = rank_cores(world) if world > 1 else 0
    sys_ = pkg.linsys.CudaRootLinearSystem(lp, device=local_rank, rank=rank, n_ranks=world, nccl_uid=uid,
                                           threads=sym_threads, stream=stream.cuda_stream, kkts=kkts)
    del kkts
    t_setup = time.perf_counter() - t_setup
    kkt = sys_.kkt
    leaf_stats = [lf.stats() for lf in sys_.leaves]

    # pinned host buffers of one step's inputs / outputs
    def pinned(a):
        t = torch.empty(max(len(a), 1), dtype=torch.float64).pin_memory()
        t[:len(a)] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        return t
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)
    h_dd0, h_nz0, h_nzl = pinned(dd[0]), pinned(nom[0]), pinned(nom[-1])
    h_ddl, h_nzlv = pinned(cat(dd[1:])), pinned(cat(nom[1:-1]))
    rng = np.random.default_rng(1234)
    rhs0 = rng.standard_normal(sys_.nb0)                     # the root right-hand side is the same on every rank
    rhsl = np.random.default_rng(1234 + 7919 * (rank + 1)).standard_normal(sys_.sum_n)
    h_b0, h_bl = pinned(rhs0), pinned(rhsl)
    src_b0, src_bl = pinned(rhs0), pinned(rhsl)
    # solveCompressed works in place: every solve of a step gets its own pinned right-hand-side buffer, refreshed between
    # the timed steps (the caller of the real code produces a new right-hand side for every solve; copying one inside the
    # timed region would charge a host memcpy of the whole vector to the e2e number)
    e2e_bufs = [(pinned(rhs0), pinned(rhsl)) for _ in range(args.solves)]

    def refresh_e2e():
        for b0_, bl_ in e2e_bufs:
            b0_.copy_(src_b0); bl_.copy_(src_bl)

    def step_e2e():
        kkt.factorize(h_dd0.data_ptr(), h_nz0.data_ptr(), h_nzl.data_ptr(), h_ddl.data_ptr(), h_nzlv.data_ptr())
        for b0_, bl_ in e2e_bufs:
            kkt.solve_compressed(b0_.data_ptr(), bl_.data_ptr())

    def step_resident():
        kkt.factorize_resident()
        for _ in range(args.solves):
            kkt.solve_compressed_resident()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # measured FP64 tensor-pipe (DMMA) peak of this GPU: the denominator of the FP64-tensor fractions, and the
    # rate of the library's own TMA-fed GEMM kernel in isolation
    dmma_peak = sys_.ctx.dmma_peak_tflops()
    dgemm_rate = sys_.ctx.dgemm_tflops(4096)

    # ---- correctness gate: one factorize + solveCompressed through the C ABI, checked on the host ----
    check = None
    inertia_root = kkt.factorize(h_dd0.data_ptr(), h_nz0.data_ptr(), h_nzl.data_ptr(), h_ddl.data_ptr(), h_nzlv.data_ptr())
    if not args.no_check:
        h_b0.copy_(src_b0); h_bl.copy_(src_bl)
        kkt.solve_compressed(h_b0.data_ptr(), h_bl.data_ptr())
        sol0 = h_b0.numpy()[:sys_.nb0].copy()
        soll = h_bl.numpy()[:sys_.sum_n]
        offs = np.concatenate([[0], np.cumsum([ch.n + ch.my + ch.mz for ch in lp.children])]).astype(np.int64)
        sol = [sol0] + [soll[offs[i]:offs[i + 1]] for i in range(len(lp.children))]
        rhs = [rhs0] + [rhsl[offs[i]:offs[i + 1]] for i in range(len(lp.children))]

        def allsum(a):
            if world == 1:
                return a
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()
        leaf_sq, r0 = pkg.linsys.compressed_kkt_residual(lp, dd, nom, rhs, sol, None, reduce_root=allsum)
        sq = allsum(np.array([leaf_sq, float(rhsl @ rhsl)]))
        res_rel = float(np.sqrt(sq[0] + r0 @ r0) / np.sqrt(sq[1] + rhs0 @ rhs0))
        # every rank must hold the same all-reduced Schur complement and the same root solution
        sc = sys_.schur_complement()
        sig = np.array([float(sc.sum()), float(np.abs(sc).sum()), float(sol0.sum()), float(np.abs(sol0).sum())])
        same = True
        if world > 1:
            g = [torch.zeros(4, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(g, torch.from_numpy(sig).cuda())
            same = all(bool(torch.equal(g[0], x)) for x in g)
        del sc
        n0, my0, myl, mzl = lp.root.n, lp.root.my, lp.myl, lp.mzl
        leaf_inertia_ok = all(lf.inertia() == ((ch.n, ch.my + ch.mz, 0), 0) for lf, ch in zip(sys_.leaves, lp.children))
        ok = np.array([1.0 if leaf_inertia_ok else 0.0])
        if world > 1:
            t = torch.from_numpy(ok).cuda(); dist.all_reduce(t, op=dist.ReduceOp.MIN); ok = t.cpu().numpy()
        check = {"solve_residual_rel": res_rel, "tolerance": 1e-8, "root_inertia": list(inertia_root),
                 "root_inertia_expected": [n0, my0 + myl + mzl, 0], "leaf_inertia_ok": bool(ok[0] > 0),
                 "sc_and_root_solution_identical_across_ranks": bool(same),
                 "how": "residual of the full bordered KKT system through the host matvec (linsys.compressed_kkt_residual), "
                        "reduced over the ranks; SC / root-solution signatures all-gathered"}
        if not (res_rel <= 1e-8 and same and tuple(inertia_root) == (n0, my0 + myl + mzl, 0) and ok[0] > 0):
            if rank == 0:
                print(json.dumps({"error": "correctness gate failed", "check": check}))
            return 3

    for _ in range(max(args.warmup, 3)):
        refresh_e2e()
        step_e2e()
    step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sys_.ctx.launch_count()
    stage_acc = {}
    ktime_acc = {}
    ms_res = 0.0
    # resident timing, one step at a time so that the per-stage event timers can be averaged
    for _ in range(args.steps):
        ms_res += timed(step_resident, 1)
        for k_, v_ in kkt.stage_times().items():
            stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
        for k_, v_ in kkt.kernel_times().items():
            ktime_acc[k_] = ktime_acc.get(k_, 0.0) + v_
    launches = sys_.ctx.launch_count() - l0
    ms_e2e = 0.0
    for _ in range(args.steps):          # one step at a time: the right-hand-side buffers are refreshed in between (untimed)
        refresh_e2e()
        ms_e2e += timed(step_e2e, 1)
    clocks = sampler.stop()
    stage_ms = {k_: v_ / args.steps for k_, v_ in stage_acc.items()}
    kern_ms = {k_: v_ / args.steps for k_, v_ in ktime_acc.items()}
    ms_per_step = ms_res / args.steps
    ms_per_step_e2e = ms_e2e / args.steps

    # ---- algorithmic counts and the rooflines of the kernels that make up the step ----
    cnt = canonical_counts(lp, leaf_stats, args.solves)
    peak_gbs, peak_src = measured_peaks()
    pools = sys_.ctx.pool_sizes()
    nbatches = sys_.ctx.n_batches()
    nnzL = sum(s["nnzL"] for s in leaf_stats)
    tensor_src = "in-repo DMMA microbenchmark measured in this run (pipsgpu_bench_dmma_peak; MEASURED_PEAKS.json has no FP64 entry)"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(f"{args.config}:{hi - lo}", {})
    except Exception:
        pass
    roofs = []
    # (1) streaming solve: one launch = one K^-1 application over all blocks of the rank; algorithmic bytes = 24 nnz(L)
    #     (factor entry + index read once forward, once backward: B4 / 2 applications per solveCompressed)
    t_ps = kern_ms["psolve"] * 1e-3
    if t_ps > 0:
        ps_bytes = 24.0 * nnzL
        roofs.append({"kernel": "psolve_stream_kernel (persistent supernodal forward+backward solve of all blocks of the GPU, "
                                "one launch per K^-1)",
                      "bound": "hbm", "achieved": ps_bytes / t_ps / 1e9, "peak": peak_gbs, "unit": "GB/s",
                      "frac": ps_bytes / t_ps / 1e9 / peak_gbs, "traffic": traffic.get("psolve_stream_kernel"),
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": ps_bytes, "avg_launch_ms": t_ps * 1e3,
                      "launches_per_step": 2 * args.solves, "share_of_step": 2 * args.solves * t_ps * 1e3 / ms_per_step,
                      "stored_bytes_per_launch": 16.0 * pools["solve_pool"]})
    # (2) Schur-term kernels (border x border updates of the grouped fronts): algorithmic flops = the lower triangle of
    #     L_B D L_B^T over all fronts, sum_f w_f nb_f (nb_f + 1) (executed-useful flops; the canonical F2 of SURVEY 8(d) is
    #     reported separately as schur_fp64_tflops_credited)
    t_sk = kern_ms.get("schur_kernels", 0.0) * 1e-3
    schur_fl = float(sum(s["schur_flops"] for s in leaf_stats))
    if t_sk > 0 and schur_fl > 0:
        roofs.append({"kernel": "Schur-term kernels (SC -= L_B D L_B^T of the grouped fronts on the FP64 tensor pipe)",
                      "bound": "tensor", "achieved": schur_fl / t_sk / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                      "frac": schur_fl / t_sk / 1e12 / dmma_peak, "traffic": traffic.get("schur_kernels"),
                      "peak_source": tensor_src, "algorithmic_flops_per_step": schur_fl, "avg_launch_ms": t_sk * 1e3 / max(nbatches, 1),
                      "launches_per_step": nbatches, "share_of_step": t_sk * 1e3 / ms_per_step,
                      "dmma_pipe_active_pct_ncu": traffic.get("schur_dmma_pipe_active_pct")})
    # (3) dense root LDL^T (all panels of one factorisation): F3 = n_S^3 / 3 + 2 n_S^2
    t_root = stage_ms["root_ldlt"] * 1e-3
    if t_root > 0:
        roofs.append({"kernel": "root LDL^T (ldlt_diag_kernel + dgemm_nt_kernel panel / trailing updates, TMA-staged, look-ahead)",
                      "bound": "tensor", "achieved": cnt["F3"] / t_root / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                      "frac": cnt["F3"] / t_root / 1e12 / dmma_peak, "traffic": traffic.get("root_ldlt"),
                      "peak_source": tensor_src, "algorithmic_flops_per_launch_set": cnt["F3"], "avg_launch_ms": t_root * 1e3,
                      "launches_per_step": 1, "share_of_step": stage_ms["root_ldlt"] / ms_per_step,
                      "dmma_pipe_active_pct_ncu_trailing_gemm": traffic.get("root_dmma_pipe_active_pct_trailing_gemm")})
    # (4) leaf factorisation as a whole (pivot / panel / tile kernels of every level + the Schur kernels): B1 + B2
    t_leaf = stage_ms["leaf_ldlt_schur"] * 1e-3
    nlev = max(s["nlevels"] for s in leaf_stats)
    roof_leaf = {"kernel": "leaf LDL^T + Schur terms, whole stage (pivot / small_fronts / big_panel / front_update_tiles per "
                           "tree level, Schur-term kernels, pack_solve)",
                 "bound": "hbm", "achieved": (cnt["B1"] + cnt["B2"]) / t_leaf / 1e9, "peak": peak_gbs, "unit": "GB/s",
                 "frac": (cnt["B1"] + cnt["B2"]) / t_leaf / 1e9 / peak_gbs, "traffic": None, "peak_source": peak_src,
                 "algorithmic_bytes_per_step": cnt["B1"] + cnt["B2"], "tree_levels": nlev, "factor_batches": nbatches,
                 "share_of_step": stage_ms["leaf_ldlt_schur"] / ms_per_step,
                 "stage1_only_GBps": cnt["B1"] / t_leaf / 1e9,
                 "executed_update_tflops": float(sum(s["update_flops"] for s in leaf_stats)) / t_leaf / 1e12,
                 "credited_F1_F2_tflops": (cnt["F1"] + cnt["F2"]) / t_leaf / 1e12}
    roofs.sort(key=lambda r: -r["share_of_step"])
    roofline = roofs[0]
    schur_tflops = (cnt["F2"] + cnt["F3"]) / (t_leaf + t_root) / 1e12
    nbytes = 8
    h2d = nbytes * (len(dd[0]) + len(nom[0]) + len(nom[-1]) + sum(len(d) for d in dd[1:]) + sum(len(z) for z in nom[1:-1])
                    + args.solves * (sys_.nb0 + sys_.sum_n))
    d2h = nbytes * args.solves * (sys_.nb0 + sys_.sum_n)
    if world > 1:      # whole-job bytes and launches
        t = torch.tensor([float(h2d), float(d2h), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        h2d, d2h, launches = int(t[0].item()), int(t[1].item()), int(t[2].item())

    # plug-in level (DoubleLinearSolver::solve(int nrhss, double* rhss, int*), the entry the reference's blocked Schur path
    # drives, LS/DistributedLinearSystem.C:766-1047): 32 right-hand sides on one leaf through the C ABI with host buffers
    plugin = None
    if rank == 0 and sys_.leaves:
        lf0, ch0 = sys_.leaves[0], lp.children[0]
        nK0 = ch0.n + ch0.my + ch0.mz
        B = np.random.default_rng(5).standard_normal((32, nK0))
        lf0.solve(B)                                   # warm-up (allocates the leaf's persistent work buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            lf0.solve(B)
        dt = (time.perf_counter() - t0) / 3
        plugin = {"call": "pipsgpu_leaf_solve(nrhs=32) on one leaf, host buffers", "ms": dt * 1e3, "rhs_per_s": 32 / dt,
                  "leaf_rows": nK0}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys_.close()
        sys_ = None
        cpu = cpu_reference(pkg, args, total_blocks, args.cpu_sample_blocks, 1, 0)

    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step * 1e-3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False,
                "scaling": scaling_mode(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "solves_per_iteration": args.solves, "blocks_this_gpu": hi - lo,
                           "total_blocks": total_blocks,
                           "l2_policy": "working set (factor panels + solve pool, GBs) exceeds the 126 MB L2; no flush needed",
                           "e2e_policy": "per step: diagonals H2D + factorize + solves x (rhs H2D, solveCompressed, solution D2H) through "
                                         "the C ABI with pinned host buffers, one buffer per solve, refreshed between timed steps",
                           "host_generation_s": t_gen, "symbolic_setup_s": t_setup, "rank0_cpu_binding": binding},
                "e2e": {"value": ms_per_step_e2e * 1e-3, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": roofline,
                "roofline_all": roofs,
                "roofline_leaf_stage": roof_leaf,
                "check": check,
                "pools": pools,
                "stages_ms": stage_ms,
                "kernels_ms": kern_ms,
                "schur_fp64_tflops_credited": schur_tflops,
                "canonical_counts": cnt,
                "fp64_tensor_peak_measured_tflops": dmma_peak,
                "dgemm_kernel_tflops_n4096": dgemm_rate,
                "plugin_multi_rhs": plugin,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if sys_ is not None:
        sys_.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
