#!/bin/bash
# Budget-tight 1-GPU session: parity tests, bench line, grouping A/B, ncu launch list, ncu --set full of the solve kernel
# and of one factorisation's kernels. Every stage has its own timeout; outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
T0=$(date +%s)
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log; echo "t=$(( $(date +%s) - T0 ))"
if [ "$RC" != "0" ]; then
  PIPSGPU_GROUP=0 timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_nogroup.log 2>&1; echo "pytest(no groups) rc=$?" >> gpurun_out/pytest_gpu_nogroup.log
  tail -6 gpurun_out/pytest_gpu_nogroup.log
fi
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; echo "t=$(( $(date +%s) - T0 ))"
PIPSGPU_GROUP=0 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nogroup.json 2> gpurun_out/bench_nogroup.err; echo "nogroup rc=$?"
cat gpurun_out/bench_nogroup.json; echo "t=$(( $(date +%s) - T0 ))"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
echo "t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psolve_stream_kernel|dense_solve_kernel" -s 40 -c 2 -f -o gpurun_out/prof_solve $B > gpurun_out/ncu_solve.log 2>&1; echo "full solve rc=$?"
ncu -i gpurun_out/prof_solve.ncu-rep --page raw --csv > gpurun_out/prof_solve_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_solve.ncu-rep --page details > gpurun_out/prof_solve_details.txt 2>/dev/null
echo "t=$(( $(date +%s) - T0 ))"
NF=$(grep -c "pivot_kernel\|small_fronts_kernel\|big_panel_kernel\|front_update_tiles_kernel\|border_group_kernel" gpurun_out/launches.csv)
PER=$((NF / 6))
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pivot_kernel|small_fronts_kernel|big_panel_kernel|front_update_tiles_kernel|border_group_kernel" -s $((3 * PER)) -c $PER -f -o gpurun_out/prof_factor $B > gpurun_out/ncu_factor.log 2>&1; echo "full factor rc=$? (per factorisation: $PER launches)"
ncu -i gpurun_out/prof_factor.ncu-rep --page raw --csv > gpurun_out/prof_factor_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_factor.ncu-rep --page details > gpurun_out/prof_factor_details.txt 2>/dev/null
for f in gpurun_out/prof_factor.ncu-rep gpurun_out/prof_solve.ncu-rep; do
  SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 20000000 ]; then rm -f $f; fi
done
echo "t=$(( $(date +%s) - T0 ))"
du -sh gpurun_out
