#!/bin/bash
# 2-GPU session (gpurun --gpus 2): the NCCL tests (reduceKKT + Lsolve b0 all-reduce against the single-rank oracle) and
# one weak-scaling bench line (64 C2 blocks per GPU), launched the way the driver launches it.
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/smi_2gpu.txt 2>&1
timeout 240 python -m pytest tests -m gpu -x -q -k "two_gpus" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest(two_gpus) rc=$?" | tee -a gpurun_out/pytest_2gpu.log
tail -4 gpurun_out/pytest_2gpu.log; echo "t=$(( $(date +%s) - T0 ))"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu rc=$?"
cat gpurun_out/bench_2gpu.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu.err; echo "t=$(( $(date +%s) - T0 ))"
