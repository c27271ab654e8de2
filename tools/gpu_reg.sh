#!/bin/bash
# regulariser path: parity tests (with their printed error tables), then the whole suite and the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_regularizers_gpu.py -q -rA > gpurun_out/pytest_reg.log 2>&1; echo "reg rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_reg.log | tail -3
grep -E "rel err|losses ref|^\{|^E  " gpurun_out/pytest_reg.log | cut -c1-1200 | head -40
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reg.json 2> gpurun_out/bench_reg.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_reg.err | cut -c1-400
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_reg.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["full_training_step"])
print(d["kernel_ms_per_step"])
PY
