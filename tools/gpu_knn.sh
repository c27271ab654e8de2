#!/bin/bash
# KNN: parity (seeded / unseeded / exhaustive) + microbenchmark on the cfg2 batch; with a variant library as $1
# (tools/_variants/<name>.so, e.g. built with AN_NVCC_EXTRA="-DAN_KNN_STATS") candidate statistics are printed too
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_render_gpu.py -x -q -k "knn or sample_fine or fine_pass" 2>&1 | tail -5
timeout 600 python tools/bench_knn.py 2>&1 | tee gpurun_out/bench_knn.txt | tail -8
if [ -n "$1" ]; then AN_LIB_PATH=$PWD/tools/_variants/$1 timeout 600 python tools/bench_knn.py 2>&1 | tail -6; fi
