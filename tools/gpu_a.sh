#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_inference_gpu.py tests/test_render_gpu.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_a.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["launch"])
print(d["kernel_ms_per_step"]); print(d["frame_512"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_' -s 12 -c 4 -o gpurun_out/knn_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_knn.log 2>&1
tail -2 gpurun_out/ncu_knn.log | cut -c1-300
