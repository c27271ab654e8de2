mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_kernels_gpu.py -x -q -k "composite or sample_fine or searchsorted or knn_unpose_fixture or knn_seeded" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | tail -5
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp_forward or mlp_backward" > gpurun_out/sanitize_racecheck_mlp.log 2>&1
echo "racecheck mlp rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|Race reported" gpurun_out/sanitize_racecheck_mlp.log | cut -c1-250 | tail -8
timeout 200 python tools/bench_knn.py --variants 3:8 --check 1 2>&1 | tail -3
