#!/bin/bash
# same-box A/B of an environment switch on the bench headline: tools/gpu_env_ab.sh AN_SOME_SWITCH
V=${1:?name of the environment variable}
for i in 1 2 3; do
  for val in "" 1; do
    env $V=$val timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-full-step --no-frozen-step > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python -c "import json; d=json.load(open('gpurun_out/ab.json')); print('$V=[$val]', round(d['ms_per_step'],4), 'ms/step')"
  done
done
