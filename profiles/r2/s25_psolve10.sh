#!/bin/bash
# round 2, session 24: psolve_stream_kernel with 10 warps per SM (PS_SMAX 256, PS_NOUT 96)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s25_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s25_pytest_gpu.log
tail -6 gpurun_out/s25_pytest_gpu.log
run() { name=$1; shift; timeout 1800 python bench.py "$@" > gpurun_out/s25_bench_$name.json 2> gpurun_out/s25_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s25_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s25_bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3 --no-cpu-baseline
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
