#!/bin/bash
# round 2, session 16: self-contained group tiles (no dependent descriptor loads), GROUP_PIPE on C3/C4, C5 variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s16_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s16_pytest_gpu.log
tail -6 gpurun_out/s16_pytest_gpu.log
run() { name=$1; shift; timeout 1800 python bench.py "$@" > gpurun_out/s16_bench_$name.json 2> gpurun_out/s16_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s16_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s16_bench_$name.err; }
C2="--config C2 --steps 5 --warmup 3 --no-cpu-baseline"
C3="--config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline"
C4="--config C4 --steps 3 --warmup 3 --no-cpu-baseline"
C5="--config C5 --steps 3 --warmup 3 --no-cpu-baseline"
run c2 $C2
run c3_64 $C3
run c4 $C4
PIPSGPU_GROUP_PIPE=1 run c3_pipe $C3 --no-check
PIPSGPU_GROUP_PIPE=1 run c4_pipe $C4 --no-check
run c5 $C5
PIPSGPU_GROUP_BY_LEVEL=0 run c5_nolvl $C5 --no-check
PIPSGPU_BATCH_GB=40 run c5_b40 $C5 --no-check
du -sh gpurun_out
