#!/bin/bash
# compute-sanitizer passes over the small-size kernel tests: memcheck (out-of-bounds / misaligned) on every kernel family,
# racecheck (shared-memory hazards) on the non-tensor kernels.  Logs under gpurun_out/.
mkdir -p gpurun_out
SEL='raygen or sample_coarse or searchsorted or sample_fine or composite or knn_unpose_fixture or knn_seeded or mlp_forward or mlp_backward or body_tables or render_loss or compact_valid or rays_sample'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_train_rays_gpu.py -x -q -k "$SEL or draws or identity" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_memcheck.log | tail -8
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_mesh_gpu.py tests/test_inference_gpu.py -x -q -k "equals_oracle or extract_mesh or density_grid" > gpurun_out/sanitize_memcheck_mesh.log 2>&1
echo "memcheck mesh/lattice rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_memcheck_mesh.log | tail -8
# (no -x: under the sanitizer the CPU oracle's first parallel torch op can come out ~1e-3 off on one thread's chunk of rows --
#  tools/_variants/comp_probe3.py showed the device result identical on a rerun and equal to every later oracle run -- so a value
#  assertion may fail once; what this pass is read for is the RACECHECK SUMMARY line)
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_mesh_gpu.py -q -k "composite or sample_fine or searchsorted or knn_unpose_fixture or render_loss or equals_oracle" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | tail -8
