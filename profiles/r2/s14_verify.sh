#!/bin/bash
# round 2, session 14: verification of the final tree (compat Schur solver in the C++ shim test, pattern re-declaration, bench line C2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s14_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s14_pytest_gpu.log
tail -25 gpurun_out/s14_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s14_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s14_smoke.log
timeout 900 python bench.py --config C2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s14_bench_c2.json 2> gpurun_out/s14_bench_c2.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/s14_bench_c2.json; tail -3 gpurun_out/s14_bench_c2.err
