#!/bin/bash
# forward-kernel organisations A/B: parity of impl 2 (forward, stash images, backward through its stash), then timings
timeout 300 env AN_MLP_FWD_IMPL=2 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp" 2>&1 | tail -4
for impl in 0 2; do
  echo "== impl $impl"
  AN_MLP_FWD_IMPL=$impl timeout 120 python tools/profile_mlp.py 1048576 --train 2>&1 | tail -1
  AN_MLP_FWD_IMPL=$impl timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -1
done
