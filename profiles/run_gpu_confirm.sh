#!/bin/bash
# Confirmation of the committed default configuration (1 GPU): parity suite, official bench line, launch list.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log; echo "t=$(( $(date +%s) - T0 ))"
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; echo "t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$? t=$(( $(date +%s) - T0 ))"
