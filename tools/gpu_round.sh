#!/bin/bash
# One GPU-box visit for the round record: parity tests, bench, ncu launch list of the bench
# command, and one `ncu --set full` capture of each of our hot kernels inside a bench step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench.json
SHORT="--steps 2 --warmup 3 --no-cpu-baseline --no-full-step --no-frozen-step --no-gpu-eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py $SHORT > gpurun_out/bench_ncu.log 2>&1
# full capture: skip the warm-up launches of each kernel (-s: 3 eager steps x 21 matching launches), take one step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_fwd_tc|mlp_bwd|knn_search|knn_classify|knn_unpose_bwd|composite|sample_fine|body_tables_bwd|rays_sample|ray_point_grad' -s 63 -c 21 -o gpurun_out/step_full -f python bench.py $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -40
