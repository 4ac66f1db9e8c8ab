#!/bin/bash
# round 2, session 10 (gpurun --gpus 2): NCCL parity tests (dense + sparse root + device BiCGStab over 2 ranks), C3 strong scaling at N=2 (128 blocks)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/s10_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/s10_pytest_2gpu.log 2>&1; echo "pytest(two_gpus) rc=$?" | tee -a gpurun_out/s10_pytest_2gpu.log
tail -15 gpurun_out/s10_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config C3 --total-blocks 128 --steps 3 --warmup 3 > gpurun_out/s10_bench_c3_128_2gpu.json 2> gpurun_out/s10_bench_c3_128_2gpu.err; echo "bench 2gpu rc=$?"
cat gpurun_out/s10_bench_c3_128_2gpu.json | cut -c1-1200; tail -5 gpurun_out/s10_bench_c3_128_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --config C3 --total-blocks 128 --steps 1 --warmup 0 > gpurun_out/s10_bench_ref_2gpu.json 2> gpurun_out/s10_bench_ref_2gpu.err; echo "bench ref rc=$?"
cat gpurun_out/s10_bench_ref_2gpu.json | cut -c1-800; tail -3 gpurun_out/s10_bench_ref_2gpu.err
