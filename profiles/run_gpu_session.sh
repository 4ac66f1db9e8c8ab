#!/bin/bash
# One GPU-box session (1 GPU): parity tests, bench line (+ A/B switches), ncu launch list, ncu --set full of one whole
# factorisation (all factor kernels) and of the solve kernels. Outputs under gpurun_out/ (kept small).
# usage: run_gpu_session.sh [ref] [ab] [full]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
if [ "$RC" != "0" ]; then
  # isolate: border groups off / level-scheduled solve
  PIPSGPU_GROUP=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_nogroup.log 2>&1; echo "pytest(no groups) rc=$?" >> gpurun_out/pytest_gpu_nogroup.log
  tail -8 gpurun_out/pytest_gpu_nogroup.log
  PIPSGPU_SOLVE=level timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_level.log 2>&1; echo "pytest(level solve) rc=$?" >> gpurun_out/pytest_gpu_level.log
  tail -8 gpurun_out/pytest_gpu_level.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for a in "$@"; do
  case "$a" in
  ref) timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?";;
  ab)  PIPSGPU_GROUP=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nogroup.json 2> gpurun_out/bench_nogroup.err; echo "nogroup rc=$?"; cat gpurun_out/bench_nogroup.json
       PIPSGPU_GROUP_FRAC=0.8 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_frac08.json 2> gpurun_out/bench_frac08.err; echo "frac0.8 rc=$?"; cat gpurun_out/bench_frac08.json;;
  esac
done
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
for a in "$@"; do
  if [ "$a" == "full" ]; then
    # the 4th factorisation of the run, every factor kernel launch (counts per factorisation are in the launch list)
    NF=$(grep -c "pivot_kernel\|small_fronts_kernel\|big_panel_kernel\|front_update_tiles_kernel\|border_group_kernel" gpurun_out/launches.csv)
    PER=$((NF / 6))
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pivot_kernel|small_fronts_kernel|big_panel_kernel|front_update_tiles_kernel|border_group_kernel" -s $((3 * PER)) -c $PER -f -o gpurun_out/prof_factor $B > gpurun_out/ncu_factor.log 2>&1; echo "full factor rc=$? (per factorisation: $PER launches)"
    ncu -i gpurun_out/prof_factor.ncu-rep --page raw --csv > gpurun_out/prof_factor_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_factor.ncu-rep --page details > gpurun_out/prof_factor_details.txt 2>/dev/null
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"psolve_stream_kernel|dense_solve_kernel" -s 40 -c 3 -f -o gpurun_out/prof_solve $B > gpurun_out/ncu_solve.log 2>&1; echo "full solve rc=$?"
    ncu -i gpurun_out/prof_solve.ncu-rep --page raw --csv > gpurun_out/prof_solve_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_solve.ncu-rep --page details > gpurun_out/prof_solve_details.txt 2>/dev/null
    for f in gpurun_out/prof_factor.ncu-rep gpurun_out/prof_solve.ncu-rep; do
      SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 20000000 ]; then rm -f $f; fi
    done
  fi
done
du -sh gpurun_out
