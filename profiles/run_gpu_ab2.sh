#!/bin/bash
# A/B session 2 (1 GPU): border-group kernel reading the persistent W21 scratch, division-free streaming solve.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log; echo "t=$(( $(date +%s) - T0 ))"
PIPSGPU_BIG_U=64 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_big64.log 2>&1; echo "pytest(BIG_U=64) rc=$?" | tee -a gpurun_out/pytest_gpu_big64.log
tail -4 gpurun_out/pytest_gpu_big64.log; echo "t=$(( $(date +%s) - T0 ))"
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
run big64 PIPSGPU_BIG_U=64
run big64_group0 PIPSGPU_BIG_U=64 PIPSGPU_GROUP=0
run big64_frac04 PIPSGPU_BIG_U=64 PIPSGPU_GROUP_FRAC=0.4
run big64_frac08 PIPSGPU_BIG_U=64 PIPSGPU_GROUP_FRAC=0.8
run big64_frac03 PIPSGPU_BIG_U=64 PIPSGPU_GROUP_FRAC=0.3
run big48 PIPSGPU_BIG_U=48
export PIPSGPU_BIG_U=64
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_big64.csv $B > gpurun_out/bench_ncu_big64.log 2>&1; echo "ncu list big64 rc=$? t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psolve_stream_kernel" -s 40 -c 1 -f -o gpurun_out/prof_solve $B > gpurun_out/ncu_solve.log 2>&1; echo "full solve rc=$? t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"border_group_kernel" -s 3 -c 1 -f -o gpurun_out/prof_group $B > gpurun_out/ncu_group.log 2>&1; echo "full group rc=$? t=$(( $(date +%s) - T0 ))"
for f in gpurun_out/prof_group.ncu-rep gpurun_out/prof_solve.ncu-rep; do
  SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 30000000 ]; then rm -f $f; fi
done
du -sh gpurun_out
