#!/bin/bash
# round 2, session 13: e2e path without host staging (device-side diagonal assembly, per-solve pinned buffers), plug-in multi-RHS timing,
# ncu evidence with bounded output (reports are reduced to text on the box), default bench line (C3, 512 blocks, 1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s13_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s13_pytest_gpu.log
tail -4 gpurun_out/s13_pytest_gpu.log
run() { name=$1; shift; timeout 1800 python bench.py "$@" > gpurun_out/s13_bench_$name.json 2> gpurun_out/s13_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s13_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "plugin", d.get("plugin_multi_rhs"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s13_bench_$name.err; }
run c2 --config C2 --steps 5 --warmup 3
run c3_64 --config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline
# group launches of one C3 factorisation: keep the details page of every launch as text, drop the report
timeout 900 ncu --set full --clock-control none -k regex:border_group -s 44 -c 24 -o gpurun_out/s13_group_c3 python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s13_ncu_group.log 2>&1; echo "ncu group rc=$?"
ncu -i gpurun_out/s13_group_c3.ncu-rep --page raw --csv > gpurun_out/s13_group_c3_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/s13_group_c3_raw.csv')))
hdr=rows[0]; data=rows[2:]
keep=['ID','Kernel Name','launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_red.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
idx=[hdr.index(k) for k in keep if k in hdr]
with open('gpurun_out/s13_group_c3_launches.csv','w') as f:
    w=csv.writer(f); w.writerow([hdr[i] for i in idx]); w.writerow([rows[1][i] for i in idx])
    for r in data: w.writerow([r[i] for i in idx])
print('group launches captured:', len(data))
PY
rm -f gpurun_out/s13_group_c3.ncu-rep gpurun_out/s13_group_c3_raw.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ldlt_diag|dgemm_nt" -s 60 -c 6 -o gpurun_out/s13_root_c4 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s13_ncu_root.log 2>&1; echo "ncu root rc=$?"
run default --steps 3 --warmup 3
du -sh gpurun_out; ls -la gpurun_out | grep s13_ | tail -14
