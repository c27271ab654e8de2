#!/bin/bash
# ncu --set full of the KNN search kernel inside the microbenchmark (variant spec $1, default 3:16):
# launches 0-3 = coarse pass, 4-7 = fine pass unseeded, 8-11 = fine pass seeded
mkdir -p gpurun_out
V=${1:-3:16}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_search' -s 4 -c 12 -o gpurun_out/knn_prof -f python tools/bench_knn.py --variants $V --check 0 --reps 2 > gpurun_out/ncu_knn.log 2>&1
tail -3 gpurun_out/ncu_knn.log | cut -c1-200
