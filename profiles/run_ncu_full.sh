#!/bin/bash
# ncu --set full captures of the top kernels of one bench step (C2, 1 GPU). Outputs under gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:psolve -s 6 -c 2 -f -o gpurun_out/prof_psolve $B > gpurun_out/ncu_psolve.log 2>&1; echo "psolve rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"front_update_tiles|small_fronts|big_panel|pivot_kernel" -s 82 -c 82 -f -o gpurun_out/prof_factor $B > gpurun_out/ncu_factor.log 2>&1; echo "factor rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_solve" -s 4 -c 1 -f -o gpurun_out/prof_dsolve $B > gpurun_out/ncu_dsolve.log 2>&1; echo "dsolve rc=$?"
ls -la gpurun_out
