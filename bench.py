#!/usr/bin/env python
"""bench.py - seconds per IPM iteration spent in the KKT factor + solve (BASELINE.json metric) on N B200s.

One "step" = the KKT work of one interior-point iteration on a fixed synthetic iterate:
    1 x factorize        (diagonals down, leaf LDL^T + Schur terms, SC all-reduce, finalize, root LDL^T)
  + S x solveCompressed  (Lsolve, Dsolve, Ltsolve; S = --solves, default 8 = predictor + corrector + 2 Gondzio
                          correctors with on average one BiCGStab iteration, LinearSystem.C:550-798)
Workload at N=1: BASELINE.json configs[1] = "C2" (64 blocks x 44 000 KKT rows, n_S = 800). For N > 1 every rank owns
64 C2-shaped blocks of a 64*N-block problem (weak scaling; blocks assigned contiguously exactly like the reference's
mapChildrenToNSubTrees), the Schur complement is all-reduced over NCCL.

  value : device-resident step (diagonals / right-hand sides already in HBM), CUDA events on the launching stream
  e2e   : the same step through the C-ABI calls with pinned HOST buffers (H2D of diagonals and right-hand sides and D2H
          of the solutions inside the timed region)
  --impl reference : the CPU path (oracle/: C multifrontal LDL^T + Schur terms + LAPACK dsytrf root, all host threads)
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
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--blocks-per-gpu", type=int, default=0,
                    help="weak scaling with this many blocks per GPU; 0 = the scaling mode's default")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="strong: the config's block count is split over the GPUs (pipsgpu_partition = "
                         "mapChildrenToNSubTrees); weak: blocks-per-gpu blocks on every GPU; auto: C2 weak, others strong")
    ap.add_argument("--total-blocks", type=int, default=0, help="strong scaling: override the config's block count")
    ap.add_argument("--solves", type=int, default=8, help="solveCompressed calls per IPM iteration")
    ap.add_argument("--cpu-sample-blocks", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def scaling_mode(args):
    if args.scaling != "auto":
        return args.scaling
    if args.blocks_per_gpu:
        return "weak"
    return "weak" if args.config == "C2" else "strong"


def build_problem(pkg, args, rank, world):
    P = pkg.problem
    cfg = args.config
    base_blocks = P.CONFIGS[cfg][0]
    if scaling_mode(args) == "weak":
        total = (args.blocks_per_gpu or base_blocks) * world
    else:
        total = args.total_blocks or base_blocks
    parts = pkg.capi.partition(total, world)
    lo, hi = parts[rank]
    lp = P.synthetic(cfg, n_blocks=total, block_range=(lo, hi))
    dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[cfg])
    return lp, dd, nom, total, (lo, hi)


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


def cpu_reference(pkg, lp, dd, nom, solves, total_blocks, sample_blocks, steps, warmup):
    """time the CPU path (oracle/mf.py MfBackend, all host threads) on a bounded sample of this workload's blocks and
    scale the leaf-dominated part to the full block count."""
    from oracle import mf
    cores = os.cpu_count() or 1
    nloc = len(lp.children)
    sample = sample_blocks or min(nloc, max(cores, 4))
    sample = max(1, min(sample, nloc))
    import copy
    sub = copy.copy(lp)
    sub.children = lp.children[:sample]
    be = mf.MfBackend(sub, pkg.capi.Symbolic, threads=min(cores, sample))
    dds = [dd[0]] + dd[1:1 + sample]
    noms = [nom[0]] + nom[1:1 + sample] + [nom[-1]]
    rng = np.random.default_rng(0)
    nb0 = lp.root.n + lp.root.my + lp.root.mz + lp.myl + lp.mzl
    rhs = [rng.standard_normal(nb0)] + [rng.standard_normal(ch.n + ch.my + ch.mz) for ch in sub.children]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        be.factorize(dds, noms)
        t1 = time.perf_counter()
        b = [r.copy() for r in rhs]
        be.solve_compressed(b)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
    tf = float(np.mean([t[0] for t in times]))
    ts = float(np.mean([t[1] for t in times]))
    scale = total_blocks / sample
    value = (tf + solves * ts) * scale
    return dict(value=value, unit=UNIT, cores=min(cores, sample), kind="port",
                sample=f"{sample} of {total_blocks} blocks (1 factorize + 1 solveCompressed measured, leaf work scaled x{scale:.1f}, "
                       f"{solves} solves/iteration); C multifrontal LDL^T+Schur (oracle/mf_ldlt.c) + LAPACK dsytrf root",
                t_factorize_sample=tf, t_solve_sample=ts)


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

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        total_job = (args.blocks_per_gpu or pkg.problem.CONFIGS[args.config][0]) * world
        cores = os.cpu_count() or 1
        nsample = min(total_job, args.cpu_sample_blocks or max(cores, 4))
        P = pkg.problem
        lp = P.synthetic(args.config, n_blocks=total_job, block_range=(0, nsample))   # only the sampled blocks are built
        dd, nom = P.synthetic_iterate(lp, P.CONFIG_ID[args.config])
        res = cpu_reference(pkg, lp, dd, nom, args.solves, total_job, args.cpu_sample_blocks, args.steps, args.warmup)
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["value"] * 1e3, "higher_is_better": False, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": f"{args.config}: {total_job} blocks x {lp.n_K(0)} KKT rows, n_S={lp.n_S}",
                           "solves_per_iteration": args.solves},
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------
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

    lp, dd, nom, total_blocks, (lo, hi) = build_problem(pkg, args, rank, world)
    stream = torch.cuda.Stream()
    t_setup = time.perf_counter()
    # host threads of the (one-off) symbolic phase: the ranks of one box share its cores
    sym_threads = max(1, (os.cpu_count() or 1) // world) if world > 1 else 0
    sys_ = pkg.linsys.CudaRootLinearSystem(lp, device=local_rank, rank=rank, n_ranks=world, nccl_uid=uid,
                                           threads=sym_threads, stream=stream.cuda_stream)
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
    rng = np.random.default_rng(1234 + rank)
    rhs0 = rng.standard_normal(sys_.nb0)
    rhsl = rng.standard_normal(sys_.sum_n)
    h_b0, h_bl = pinned(rhs0), pinned(rhsl)
    src_b0, src_bl = pinned(rhs0), pinned(rhsl)

    def step_e2e():
        kkt.factorize(h_dd0.data_ptr(), h_nz0.data_ptr(), h_nzl.data_ptr(), h_ddl.data_ptr(), h_nzlv.data_ptr())
        for _ in range(args.solves):
            h_b0.copy_(src_b0); h_bl.copy_(src_bl)          # fresh right-hand side (host-side memcpy)
            kkt.solve_compressed(h_b0.data_ptr(), h_bl.data_ptr())

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

    # measured FP64 tensor-pipe (DMMA) peak of this GPU: the denominator of the FP64-tensor fractions
    dmma_peak = sys_.ctx.dmma_peak_tflops()
    for _ in range(max(args.warmup, 3)):
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
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()
    stage_ms = {k_: v_ / args.steps for k_, v_ in stage_acc.items()}
    ms_per_step = ms_res / args.steps
    ms_per_step_e2e = ms_e2e / args.steps

    # ---- algorithmic counts and the rooflines of the two kernels groups that make up the step ----
    cnt = canonical_counts(lp, leaf_stats, args.solves)
    peak_gbs, peak_src = measured_peaks()
    pools = sys_.ctx.pool_sizes()
    nnzL = sum(s["nnzL"] for s in leaf_stats)
    # (1) streaming solve: one launch = one K^-1 application over all blocks of the rank; algorithmic bytes = 24 nnz(L)
    #     (factor entry + index read once forward, once backward: B4 / 2 applications per solveCompressed)
    t_ps = ktime_acc["psolve"] / args.steps * 1e-3
    ps_bytes = 24.0 * nnzL
    ps_ach = ps_bytes / t_ps / 1e9
    roof_solve = {"kernel": "psolve_stream_kernel (persistent supernodal forward+backward solve of all blocks, one launch per K^-1)",
                  "bound": "hbm", "achieved": ps_ach, "peak": peak_gbs, "unit": "GB/s", "frac": ps_ach / peak_gbs,
                  "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": ps_bytes,
                  "avg_launch_ms": t_ps * 1e3, "launches_per_step": 2 * args.solves,
                  "share_of_step": 2 * args.solves * t_ps * 1e3 / ms_per_step,
                  "stored_bytes_per_launch": 16.0 * pools["solve_pool"],
                  "stored_bytes_GBps": 16.0 * pools["solve_pool"] / t_ps / 1e9}
    # (2) leaf factorisation + fused Schur terms: all level launches of one factorize; algorithmic bytes = B1 + B2
    t_leaf = stage_ms["leaf_ldlt_schur"] * 1e-3
    nlev = max(s["nlevels"] for s in leaf_stats)
    achieved = (cnt["B1"] + cnt["B2"]) / t_leaf / 1e9
    roof_leaf = {"kernel": "leaf LDL^T + Schur terms (pivot / small_fronts / big_panel / front_update_tiles kernels, "
                           "one launch set per tree level, + pack_solve)",
                 "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                 "traffic": None, "peak_source": peak_src,
                 "algorithmic_bytes_per_step": cnt["B1"] + cnt["B2"], "tree_levels": nlev,
                 "share_of_step": stage_ms["leaf_ldlt_schur"] / ms_per_step,
                 "stage1_only_GBps": cnt["B1"] / t_leaf / 1e9,
                 "stored_factor_GBps": 8.0 * (pools["L"] + 2 * pools["C"]) / t_leaf / 1e9,
                 "credited_F1_F2_tflops": (cnt["F1"] + cnt["F2"]) / t_leaf / 1e12}
    # the single dominant KERNEL of the step is the streaming solve (16 launches); the leaf stage is a set of ~100 launches
    # of five kernels and is reported next to it
    roofline, roofline_other = roof_solve, roof_leaf
    try:
        with open(os.path.join(ROOT, "profiles", "r1", "ncu_traffic.json")) as f:
            tr = json.load(f)["psolve_stream_kernel"]
        if args.config == "C2" and (hi - lo) == 64:      # the capture was taken on this workload
            roof_solve["traffic"] = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
            roof_solve["traffic_source"] = "profiles/r1/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
    except Exception:
        pass
    t_root = stage_ms["root_ldlt"] * 1e-3
    schur_tflops = (cnt["F2"] + cnt["F3"]) / (t_leaf + t_root) / 1e12
    nbytes = 8
    h2d = nbytes * (len(dd[0]) + len(nom[0]) + len(nom[-1]) + sum(len(d) for d in dd[1:]) + sum(len(z) for z in nom[1:-1])
                    + args.solves * (sys_.nb0 + sys_.sum_n))
    d2h = nbytes * args.solves * (sys_.nb0 + sys_.sum_n)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(pkg, lp, dd, nom, args.solves, total_blocks, args.cpu_sample_blocks, 1, 1)

    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step * 1e-3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False,
                "scaling": scaling_mode(args),
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.config}: {total_blocks} blocks x {lp.n_K(0)} KKT rows "
                                       f"({hi - lo} per GPU), n_S={lp.n_S}, fixed synthetic iterate (SURVEY 8d)",
                           "solves_per_iteration": args.solves, "blocks_per_gpu": hi - lo,
                           "l2_policy": "working set (factor + contribution pools, several GB) exceeds the 126 MB L2; no flush needed",
                           "symbolic_setup_s": t_setup},
                "e2e": {"value": ms_per_step_e2e * 1e-3, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": roofline,
                "roofline_other": roofline_other,
                "pools": pools,
                "stages_ms": stage_ms,
                "kernels_ms": {k_: v_ / args.steps for k_, v_ in ktime_acc.items()},
                "schur_fp64_tflops_credited": schur_tflops,
                "canonical_counts": cnt,
                "fp64_tensor_peak_measured_tflops": dmma_peak,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    sys_.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
