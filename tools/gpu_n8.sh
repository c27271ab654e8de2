#!/bin/bash
# the bench line on N GPUs of one box (default 8): torchrun, one rank per GPU
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 ${@:2} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -v "NCCL INFO" gpurun_out/bench_n$N.err | grep -v "^$" | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
for k in ("value","ms_per_step","e2e","frame_512","frame_1080_seq","grid_512","full_training_step","frozen_body_params_step","weights_in_sync","clocks"):
    v = d.get(k)
    if isinstance(v, dict): v = {a: b for a, b in v.items() if a not in ("workload", "includes", "what")}
    print(k, v)
PY
