#!/bin/bash
# Final 1-GPU session of the round: parity on the default configuration and on the three switchable candidates
# (PIPSGPU_XFORM=1, PIPSGPU_PANEL=staged, PIPSGPU_GROUP_PIPE=1), the official bench line, per-switch bench lines, launch
# list and ncu --set full captures of the configuration that passed.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "t=$(( $(date +%s) - T0 ))"; }
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC0=$?; echo "pytest(default) rc=$RC0" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log; el
NEW="PIPSGPU_XFORM=1 PIPSGPU_PANEL=staged PIPSGPU_GROUP_PIPE=1"
env $NEW timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_new.log 2>&1; RC1=$?; echo "pytest(new: $NEW) rc=$RC1" | tee -a gpurun_out/pytest_gpu_new.log
tail -3 gpurun_out/pytest_gpu_new.log; el
if [ "$RC1" != "0" ]; then
  for v in PIPSGPU_XFORM=1 PIPSGPU_PANEL=staged PIPSGPU_GROUP_PIPE=1; do
    env $v timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$v.log 2>&1; echo "pytest($v) rc=$?" | tee -a gpurun_out/pytest_gpu_$v.log
    tail -3 gpurun_out/pytest_gpu_$v.log; el
  done
fi
: > gpurun_out/ab.jsonl
run() {   # name, bench args, env assignments...
  local name="$1"; shift
  local args="$1"; shift
  local out
  out=$(env "$@" timeout 300 python bench.py $args 2> gpurun_out/ab_$name.err)
  echo "$name $out" >> gpurun_out/ab.jsonl
  echo "$out" > gpurun_out/bench_$name.json
  echo "$name $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); s=d["stages_ms"]; k=d["kernels_ms"]
    print("step %.2f e2e %.2f leaf %.2f psolve %.3f pack %.3f root %.2f dense %.3f" % (d["ms_per_step"], d["e2e"]["value"]*1e3, s["leaf_ldlt_schur"], k["psolve"], k["pack_solve"], s["root_ldlt"], k["dense_solve"]))
except Exception as e:
    print("FAILED", e)') $(el)"
}
run default "--steps 5 --warmup 3" PIPSGPU_DUMMY=1
S="--steps 3 --warmup 3 --no-cpu-baseline"
run xform "$S" PIPSGPU_XFORM=1
run staged "$S" PIPSGPU_PANEL=staged
run pipe "$S" PIPSGPU_GROUP_PIPE=1
run new "$S" $NEW
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
if [ "$RC1" == "0" ]; then export $NEW; echo "profiling the NEW configuration"; else echo "profiling the DEFAULT configuration"; fi
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$? $(el)"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"psolve_stream_kernel" -s 40 -c 1 -f -o gpurun_out/prof_solve $B > gpurun_out/ncu_solve.log 2>&1; echo "full solve rc=$? $(el)"
# the heaviest level of the factorisation: launches 8..10 of the tile / panel kernels of the 4th factorisation
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"front_update_tiles_kernel|big_panel_kernel" -s 60 -c 8 -f -o gpurun_out/prof_factor $B > gpurun_out/ncu_factor.log 2>&1; echo "full factor rc=$? $(el)"
for f in gpurun_out/prof_factor.ncu-rep gpurun_out/prof_solve.ncu-rep; do
  SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 30000000 ]; then rm -f $f; fi
done
du -sh gpurun_out
