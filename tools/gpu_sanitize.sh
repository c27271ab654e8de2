#!/bin/bash
# compute-sanitizer passes over the small-size kernel tests: memcheck (out-of-bounds / misaligned) on every kernel family,
# racecheck (shared-memory hazards) on the non-tensor kernels.  Logs under gpurun_out/.
mkdir -p gpurun_out
SEL='raygen or sample_coarse or searchsorted or sample_fine or composite or knn_unpose_fixture or mlp_forward or mlp_backward or body_tables'
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_train_rays_gpu.py -x -q -k "$SEL or draws or identity" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_memcheck.log | tail -8
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "composite or sample_fine or searchsorted or knn_unpose_fixture" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | tail -8
