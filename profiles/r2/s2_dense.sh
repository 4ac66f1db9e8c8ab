#!/bin/bash
# round 2, session 2: new dense root stage (TMA GEMM, blocked LDL^T, multi-CTA solve), batched factorisation, C3-512 at N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "dense or root_dim or sc_assembly" > gpurun_out/pytest_dense.log 2>&1; echo "pytest dense rc=$?"
tail -15 gpurun_out/pytest_dense.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
PIPSGPU_BATCH_GB=0.02 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "schur_complement_and_solve or golden or regularised" > gpurun_out/pytest_batched.log 2>&1; echo "pytest batched rc=$?"
tail -5 gpurun_out/pytest_batched.log
run() { name=$1; shift; timeout 1200 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; cat gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c3_512 --config C3 --steps 3 --warmup 3 --no-cpu-baseline
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config C4 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
du -sh gpurun_out
