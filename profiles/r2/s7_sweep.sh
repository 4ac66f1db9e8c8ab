#!/bin/bash
# round 2, session 7: full gpu suite on the restored 64x64 group kernel + new diag kernel + slab-parallel panels; symbolic-parameter sweep on C3 (64 blocks) and C4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s7_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s7_pytest_gpu.log
tail -12 gpurun_out/s7_pytest_gpu.log
run() { name=$1; shift; timeout 1200 python bench.py "$@" > gpurun_out/s7_bench_$name.json 2> gpurun_out/s7_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s7_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"), d["pools"]["L"], d["pools"]["solve_pool"], d["pools"]["fronts"])
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s7_bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3 --no-cpu-baseline
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
C3="--config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
PIPSGPU_RELAX_SMALL=0.6 run c3_rs06 $C3
PIPSGPU_RELAX_SMALL=0.4 run c3_rs04 $C3
PIPSGPU_RELAX_MID=0.25 run c3_rm025 $C3
PIPSGPU_SUBTREE_WIDTH=16 run c3_sw16 $C3
PIPSGPU_SUBTREE_WIDTH=64 run c3_sw64 $C3
PIPSGPU_BIG_U=128 run c3_bu128 $C3
PIPSGPU_BIG_U=48 run c3_bu48 $C3
PIPSGPU_GROUP_FRAC=0.5 run c3_gf05 $C3
PIPSGPU_GROUP_FRAC=0.75 run c3_gf075 $C3
C4="--config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
PIPSGPU_GROUP_FRAC=0.5 run c4_gf05 $C4
PIPSGPU_GROUP_FRAC=0.75 run c4_gf075 $C4
PIPSGPU_BIG_U=128 run c4_bu128 $C4
du -sh gpurun_out
