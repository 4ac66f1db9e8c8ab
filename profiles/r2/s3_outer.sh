#!/bin/bash
# round 2, session 3: validate device BiCGStab / refinement / inertia loop; full gpu suite; per-kernel launch list of C2 and C3-64
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_outer_solve.py -m gpu -q > gpurun_out/s3_pytest_outer.log 2>&1; echo "pytest outer rc=$?"
tail -40 gpurun_out/s3_pytest_outer.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/s3_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_gpu.log
tail -15 gpurun_out/s3_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/s3_launches_c2.csv python bench.py --config C2 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/s3_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/s3_launches_c3_64.csv python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/s3_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
du -sh gpurun_out
