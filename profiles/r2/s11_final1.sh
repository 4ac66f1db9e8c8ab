#!/bin/bash
# round 2, session 11: new IPM-with-device-outer-solve tests, group-kernel occupancy variants, relax_mid on C2/C4, ncu of the heavy group launch + launch list of C2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s11_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s11_pytest_gpu.log
tail -12 gpurun_out/s11_pytest_gpu.log
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/s11_bench_$name.json 2> gpurun_out/s11_bench_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s11_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, {k:round(v,2) for k,v in d["kernels_ms"].items()}, (d.get("check") or {}).get("solve_residual_rel"))
except Exception as e:
    print("$name parse failed", e)
PY
tail -3 gpurun_out/s11_bench_$name.err; }
C2="--config C2 --steps 5 --warmup 3 --no-cpu-baseline --no-check"
C3="--config C3 --blocks-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
C4="--config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-check"
for occ in 5 6; do
PIPSGPU_GROUP_OCC=$occ run c2_occ$occ $C2
PIPSGPU_GROUP_OCC=$occ run c3_occ$occ $C3
PIPSGPU_GROUP_OCC=$occ run c4_occ$occ $C4
done
PIPSGPU_RELAX_MID=0.25 run c2_rm025 $C2
PIPSGPU_RELAX_MID=0.25 run c4_rm025 $C4
PIPSGPU_RELAX_MID=0.4 run c3_rm04 $C3
PIPSGPU_GROUP_BY_LEVEL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:border_group -s 3 -c 1 -o gpurun_out/s11_group_c3 python bench.py --config C3 --blocks-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/s11_ncu_group.log 2>&1; echo "ncu group rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/s11_launches_c2.csv python bench.py --config C2 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/s11_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ls -la gpurun_out | grep s11_ | tail -8
