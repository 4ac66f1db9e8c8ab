#!/bin/bash
# First GPU session of round 2: validate and time the candidates that were written after round 1's GPU budget ran out.
#   PIPSGPU_SOLVE=direct   psolve_direct.cu (no-staging streaming solve, 24 warps per SM)
#   PIPSGPU_EXTENDED=1     tests/test_gpu_extended.py (C3 / C4 / C5 block shapes)
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "t=$(( $(date +%s) - T0 ))"; }
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?"; tail -3 gpurun_out/pytest_gpu.log; el
PIPSGPU_SOLVE=direct timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_direct.log 2>&1; echo "pytest(direct) rc=$?"; tail -15 gpurun_out/pytest_gpu_direct.log; el
S="--steps 3 --warmup 3 --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stages_ms"]["leaf_ldlt_schur"], d["kernels_ms"])'
timeout 200 python bench.py $S 2>/dev/null | tee gpurun_out/bench_default.json | python -c "$P"; el
PIPSGPU_SOLVE=direct timeout 200 python bench.py $S 2>gpurun_out/bench_direct.err | tee gpurun_out/bench_direct.json | python -c "$P"; el
PIPSGPU_EXTENDED=1 timeout 900 python -m pytest tests/test_gpu_extended.py -x -q > gpurun_out/pytest_gpu_extended.log 2>&1; echo "pytest(extended) rc=$?"; tail -15 gpurun_out/pytest_gpu_extended.log; el
PIPSGPU_SOLVE=direct timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psolve_direct_kernel" -s 40 -c 1 -f -o gpurun_out/prof_solve_direct python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solve_direct.log 2>&1; echo "ncu direct rc=$?"; el
