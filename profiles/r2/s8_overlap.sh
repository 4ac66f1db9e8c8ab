#!/bin/bash
# round 2, session 8: group SYRKs overlapped with the tree (per-level launches on their own stream, deferred join across batches)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s8_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s8_pytest_gpu.log
tail -12 gpurun_out/s8_pytest_gpu.log
PIPSGPU_BATCH_GB=0.02 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "schur_complement_and_solve or golden or regularised" > gpurun_out/s8_pytest_batched.log 2>&1; echo "pytest batched rc=$?"
tail -3 gpurun_out/s8_pytest_batched.log
run() { name=$1; shift; timeout 1200 python bench.py "$@" > gpurun_out/s8_bench_$name.json 2> gpurun_out/s8_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s8_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s8_bench_$name.err; }
C2="--config C2 --steps 5 --warmup 3 --no-cpu-baseline"
C3="--config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline"
C4="--config C4 --steps 3 --warmup 3 --no-cpu-baseline"
run c2 $C2
run c3_64 $C3
run c4 $C4
PIPSGPU_GROUP_BY_LEVEL=0 run c2_nolvl $C2 --no-check
PIPSGPU_GROUP_BY_LEVEL=0 run c3_nolvl $C3 --no-check
PIPSGPU_GROUP_BY_LEVEL=0 run c4_nolvl $C4 --no-check
PIPSGPU_MIN_BATCHES=1 run c2_mb1 $C2 --no-check
PIPSGPU_MIN_BATCHES=1 run c3_mb1 $C3 --no-check
PIPSGPU_MIN_BATCHES=1 run c4_mb1 $C4 --no-check
PIPSGPU_MIN_BATCHES=4 run c2_mb4 $C2 --no-check
PIPSGPU_MIN_BATCHES=4 run c3_mb4 $C3 --no-check
PIPSGPU_MIN_BATCHES=4 run c4_mb4 $C4 --no-check
PIPSGPU_MIN_BATCHES=8 run c3_mb8 $C3 --no-check
PIPSGPU_MIN_BATCHES=8 run c4_mb8 $C4 --no-check
du -sh gpurun_out
