#!/bin/bash
# last seconds of the round's GPU budget: parity + one bench line of the two newest candidates
N="PIPSGPU_GROUP_LISTS=1 PIPSGPU_BP_THREADS=128"
S="--steps 3 --warmup 3 --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stages_ms"]["leaf_ldlt_schur"], d["kernels_ms"])'
env $N timeout 40 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
env $N timeout 40 python bench.py $S 2>/dev/null | tee gpurun_out/bench_lists_bp128.json | python -c "$P"
env PIPSGPU_GROUP_LISTS=1 timeout 40 python bench.py $S 2>/dev/null | tee gpurun_out/bench_lists.json | python -c "$P"
