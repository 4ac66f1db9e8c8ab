#!/bin/bash
# round 2, session 12: final build - full gpu suite, the four configurations, ncu evidence (group launches of one C3 factorisation, diag block, trailing GEMM at C4), default bench line (C3, 512 blocks, 1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s12_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s12_pytest_gpu.log
tail -12 gpurun_out/s12_pytest_gpu.log
run() { name=$1; shift; timeout 1800 python bench.py "$@" > gpurun_out/s12_bench_$name.json 2> gpurun_out/s12_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s12_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s12_bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
timeout 900 ncu --set full --clock-control none --import-source on -k regex:border_group -s 44 -c 24 -o gpurun_out/s12_group_c3 python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s12_ncu_group.log 2>&1; echo "ncu group rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ldlt_diag|dgemm_nt" -s 60 -c 8 -o gpurun_out/s12_root_c4 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s12_ncu_root.log 2>&1; echo "ncu root rc=$?"
run default --steps 3 --warmup 3
ls -la gpurun_out | grep s12_ | tail -12
