#!/bin/bash
# Multi-GPU bench line only (the scaling point the driver runs): gpurun --gpus N -- bash tools/gpu_n8.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["frame_512"]["ms"], d["full_training_step"])
PY
wc -l gpurun_out/bench_n$N.json
