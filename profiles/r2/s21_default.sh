#!/bin/bash
# round 2, session 21: the line the driver runs - python bench.py (C3, 512 blocks, 1 GPU, with the CPU baseline) - on the final tree
mkdir -p gpurun_out
timeout 1800 python bench.py > gpurun_out/s21_bench_default.json 2> gpurun_out/s21_bench_default.err; echo "bench default rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s21_bench_default.json").read().strip().splitlines()[-1])
print("default ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), d["steps"], d["warmup"], {k:round(v,2) for k,v in d["stages_ms"].items()}, d["check"]["solve_residual_rel"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline"]["kernel"][:30], round(d["roofline"]["frac"],3), d["clocks"])
PY
tail -3 gpurun_out/s21_bench_default.err
