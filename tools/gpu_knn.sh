#!/bin/bash
# KNN variants: parity (seeded / unseeded / exhaustive) + microbenchmark on the cfg2 batch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "knn or sample_fine" 2>&1 | tail -8
timeout 600 python tools/bench_knn.py --variants ${1:-1,3:0,3:4,3:8,3:12,3:16,2:8,4:8} 2>&1 | tee gpurun_out/bench_knn.txt | tail -30
