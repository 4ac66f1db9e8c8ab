#!/bin/bash
# One GPU-box session: parity tests, bench line, reference arm, ncu launch list. Outputs under gpurun_out/ (kept small).
# usage: run_gpu_round.sh [noref] [full:<tag>:<kernel-regex>:<skip>:<count> ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
if [ "$RC" != "0" ]; then
  # tell a defect of the streaming solve from a defect of the factors
  PIPSGPU_SOLVE=level timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_level.log 2>&1; echo "pytest(level solve) rc=$?" >> gpurun_out/pytest_gpu_level.log
  tail -8 gpurun_out/pytest_gpu_level.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "noref" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
for a in "$@"; do
  case "$a" in full:*) IFS=: read -r _ TAG RX SKIP CNT <<< "$a"; bash profiles/run_ncu_full.sh "$TAG" "$RX" "$SKIP" "$CNT";; esac
done
