#!/bin/bash
# round 2, session 17 (gpurun --gpus 2): rank -> CPU-affinity binding in bench.py; 2-GPU NCCL tests with the final tree
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s17_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/s17_pytest_2gpu.log 2>&1; echo "pytest(two_gpus) rc=$?" | tee -a gpurun_out/s17_pytest_2gpu.log
tail -4 gpurun_out/s17_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config C3 --total-blocks 128 --steps 3 --warmup 3 > gpurun_out/s17_bench_c3_128_2gpu.json 2> gpurun_out/s17_bench_c3_128_2gpu.err; echo "bench 2gpu rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/s17_bench_c3_128_2gpu.json").read().strip().splitlines()[-1])
    print("2gpu ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]*1e3,2), d["config"].get("rank0_cpu_binding"), d["config"]["symbolic_setup_s"], d["check"]["solve_residual_rel"])
except Exception as e:
    print("parse failed", e)
PY
tail -4 gpurun_out/s17_bench_c3_128_2gpu.err; head -8 gpurun_out/s17_topo.txt
