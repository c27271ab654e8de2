#!/bin/bash
# MLP iteration loop on the GPU box: kernel parity tests, then kernel timings (train + inference) and, when a
# trace build exists (AN_MLP_TRACE=1 AN_LIB_PATH=tools/_variants/libtrace.so python -m anim_nerf_b200._build), its timeline.
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp" 2>&1 | tail -15
timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -5
timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -2
if [ -f tools/_variants/libtrace.so ]; then
  AN_LIB_PATH=$PWD/tools/_variants/libtrace.so timeout 120 python tools/trace_mlp.py --train 2>&1 | tail -40
fi
