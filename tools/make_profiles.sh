#!/bin/bash
# Turn one tools/gpu_round.sh visit (gpurun_out/) into the committed records under profiles/ (round tag $1, e.g. r02).
R=${1:-r02}
python tools/launch_summary.py gpurun_out/launches.csv 2 > profiles/${R}_launches_step.txt
gzip -9 -c gpurun_out/launches.csv > profiles/${R}_launches_full.csv.gz
python tools/ncu_traffic.py gpurun_out/step_full.ncu-rep > profiles/ncu_traffic.json
: > profiles/${R}_ncu_full_summary.txt
for k in mlp_fwd_tc mlp_bwd_dgrad mlp_bwd_wgrad knn_search knn_classify knn_unpose_bwd body_tables_bwd composite_fwd composite_bwd sample_fine rays_sample_kernel ray_point_grad; do
  python tools/ncu_summary.py gpurun_out/step_full.ncu-rep 12 $k >> profiles/${R}_ncu_full_summary.txt 2>/dev/null
done
# tracked export of the capture the summaries come from: every raw metric of every captured launch
ncu -i gpurun_out/step_full.ncu-rep --page raw --csv | gzip -9 > profiles/${R}_ncu_raw_page.csv.gz
cp gpurun_out/bench.json profiles/${R}_bench.json
grep -E "PASSED|FAILED|passed|failed|max abs err|PSNR|rel L2|loss " gpurun_out/pytest_gpu.log | cut -c1-400 > profiles/${R}_pytest_gpu.txt
cuobjdump -sass anim-nerf_b200/libanimnerf_b200.so > /tmp/all.sass
python - <<'PY'
import re
txt = open('/tmp/all.sass').read()
want = {'mlp_fwd_tc_kernelILi1': 'sass_mlp_fwd.txt', 'mlp_bwd_dgrad_kernel': 'sass_mlp_dgrad.txt', 'mlp_bwd_wgrad_kernel': 'sass_mlp_wgrad.txt'}
for p in re.split(r'\n\s*Function : ', txt)[1:]:
    name = p.split('\n', 1)[0]
    for k, f in want.items():
        if k in name:
            body = re.sub(r'\s*/\* 0x[0-9a-f]{16} \*/', '', 'Function : ' + p)
            open('profiles/' + f, 'w').write("# cuobjdump -sass libanimnerf_b200.so (sm_100a), instruction encodings stripped\n" + body)
PY
ls -la profiles
