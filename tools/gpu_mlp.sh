#!/bin/bash
# MLP iteration loop on the GPU box: kernel parity tests, then kernel timings (train + inference).
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp" 2>&1 | tail -15
timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -5
timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -2
