#!/bin/bash
# round 2, session 18 (final tree) (gpurun --gpus 8): the headline configuration - C3, 512 blocks, strong scaling over 8 B200 with NCCL
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/s18_smi.txt 2>&1
nproc > gpurun_out/s18_host.txt; free -g >> gpurun_out/s18_host.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/s18_bench_c3_8gpu.json 2> gpurun_out/s18_bench_c3_8gpu.err; echo "bench 8gpu rc=$?"
cut -c1-1400 gpurun_out/s18_bench_c3_8gpu.json; tail -5 gpurun_out/s18_bench_c3_8gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/s18_bench_c3_8gpu.json").read().strip().splitlines()[-1])
    print("8gpu ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), {k:round(v,2) for k,v in d["stages_ms"].items()}, d["check"])
except Exception as e:
    print("parse failed", e)
PY
