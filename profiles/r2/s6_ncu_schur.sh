#!/bin/bash
# round 2, session 6: parity with the slab-parallel big_panel_kernel; ncu --set full of schur_syrk_kernel (C3, 64 blocks); launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/s6_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s6_pytest_gpu.log
tail -12 gpurun_out/s6_pytest_gpu.log
PIPSGPU_GROUP_FRAC=0.6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_syrk -s 2 -c 1 -o gpurun_out/s6_schur_f06 python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s6_ncu1.log 2>&1; echo "ncu1 rc=$?"
PIPSGPU_GROUP_FRAC=0.6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/s6_launches_c3_64.csv python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/s6_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | tail -8
