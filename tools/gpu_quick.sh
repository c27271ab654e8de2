#!/bin/bash
# full GPU parity suite + bench summary
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench3.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["launch"])
print(d["kernel_ms_per_step"])
print(d["roofline"]["frac"], d["roofline"]["achieved"])
PY
