#!/bin/bash
# A/B two library builds on the MLP kernels: parity tests on the default build, then timings of both.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_regularizers_gpu.py -x -q -k "mlp or sigma or second" 2>&1 | tail -4
for lib in "" $@; do
  echo "== lib: ${lib:-default}"
  AN_LIB_PATH=${lib:+$PWD/anim-nerf_b200/$lib} timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -4
  AN_LIB_PATH=${lib:+$PWD/anim-nerf_b200/$lib} timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -1
done
