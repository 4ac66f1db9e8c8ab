#!/bin/bash
# A/B session 3 (1 GPU): batched panel kernels, k-major staging, flag-skipped group members, default BIG_U = 64.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log; echo "t=$(( $(date +%s) - T0 ))"
PIPSGPU_GROUP=0 PIPSGPU_BIG_U=128 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_big128_group0.log 2>&1; echo "pytest(BIG_U=128,GROUP=0) rc=$?" | tee -a gpurun_out/pytest_gpu_big128_group0.log
tail -4 gpurun_out/pytest_gpu_big128_group0.log; echo "t=$(( $(date +%s) - T0 ))"
: > gpurun_out/ab.jsonl
run() {   # name, env assignments...
  local name="$1"; shift
  local out
  out=$(env "$@" timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$name.err)
  echo "$name $out" >> gpurun_out/ab.jsonl
  echo "$name $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); s=d["stages_ms"]; k=d["kernels_ms"]
    print("step %.2f e2e %.2f leaf %.2f psolve %.3f pack %.3f root %.2f dense %.3f pool %d" % (d["ms_per_step"], d["e2e"]["value"]*1e3, s["leaf_ldlt_schur"], k["psolve"], k["pack_solve"], s["root_ldlt"], k["dense_solve"], d["pools"]["solve_pool"]))
except Exception as e:
    print("FAILED", e)') t=$(( $(date +%s) - T0 ))"
}
run default PIPSGPU_DUMMY=1
run group0 PIPSGPU_GROUP=0
run frac04 PIPSGPU_GROUP_FRAC=0.4
run frac08 PIPSGPU_GROUP_FRAC=0.8
run big128 PIPSGPU_BIG_U=128
run big48 PIPSGPU_BIG_U=48
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$? t=$(( $(date +%s) - T0 ))"
PIPSGPU_GROUP=0 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_group0.csv $B > gpurun_out/bench_ncu_group0.log 2>&1; echo "ncu list group0 rc=$? t=$(( $(date +%s) - T0 ))"
du -sh gpurun_out
