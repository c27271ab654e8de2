#!/bin/bash
# TMEM-operand MLP forward: parity tests, then timing against the shared-memory-operand kernel (AN_MLP_SS=1)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp" 2>&1 | tail -15
echo "--- TS inference"; timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -2
echo "--- SS inference"; AN_MLP_SS=1 timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -2
echo "--- train + bwd"; timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -5
