#!/bin/bash
# round 2, session 1: round-1 kernels on every BASELINE config (baseline for the round), measured DMMA peak
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 1700 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; cat gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3 --no-cpu-baseline
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
run c5 --config C5 --steps 3 --warmup 3 --no-cpu-baseline
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_c3_64.csv python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config C4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
du -sh gpurun_out
