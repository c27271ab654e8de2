#!/bin/bash
# A/B of a library variant (tools/_variants/<name>.so, built with AN_LIB_PATH=... python -m anim_nerf_b200._build) against
# the in-tree library: MLP parity tests on the variant, then MLP kernel timings of both.
V=$PWD/tools/_variants/${1:-libshare.so}
AN_LIB_PATH=$V timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "mlp" 2>&1 | tail -3
for i in 1 2; do
echo "--- variant $1"; AN_LIB_PATH=$V timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -3
echo "--- in-tree";   timeout 120 python tools/profile_mlp.py 1048576 --bwd 2>&1 | tail -3
done
