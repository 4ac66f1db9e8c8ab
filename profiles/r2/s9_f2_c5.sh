#!/bin/bash
# round 2, session 9: full gpu suite (incl. sparse root tests), C5 line, group_frac sweep on the 64x64 kernel, ncu evidence captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s9_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s9_pytest_gpu.log
tail -30 gpurun_out/s9_pytest_gpu.log
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/s9_bench_$name.json 2> gpurun_out/s9_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s9_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s9_bench_$name.err; }
C3="--config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
C4="--config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
run c5 --config C5 --steps 3 --warmup 3 --no-cpu-baseline
for f in 0.4 0.3 0.2; do
PIPSGPU_GROUP_FRAC=$f run c3_gf$f $C3
PIPSGPU_GROUP_FRAC=$f run c4_gf$f $C4
done
# ncu evidence: group kernel (C3, 64 blocks), psolve + tiles (C2)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:border_group -s 60 -c 1 -o gpurun_out/s9_group_c3 python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s9_ncu_group.log 2>&1; echo "ncu group rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psolve_stream -s 20 -c 1 -o gpurun_out/s9_psolve_c2 python bench.py --config C2 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s9_ncu_psolve.log 2>&1; echo "ncu psolve rc=$?"
ls -la gpurun_out | grep s9_ | head -30
