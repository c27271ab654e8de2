#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the bench line and the inference configurations under torchrun.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
cat gpurun_out/bench_n$N.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/bench_inference.py > gpurun_out/inference_n$N.json 2> gpurun_out/inference_n$N.err; echo "inference rc=$?"
tail -3 gpurun_out/inference_n$N.err | cut -c1-300
cat gpurun_out/inference_n$N.json
timeout 300 python tools/bench_inference.py > gpurun_out/inference_n1.json 2> gpurun_out/inference_n1.err; echo "inference n1 rc=$?"
tail -3 gpurun_out/inference_n1.err | cut -c1-300
cat gpurun_out/inference_n1.json
