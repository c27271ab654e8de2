"""Where the time of the cfg4 density-grid query goes: per-entry-point CUDA-event times + the torch glue around them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import _lib, synthetic, inference
from anim_nerf_b200.system import AnimNeRFSystem
dev = "cuda"
sysm = AnimNeRFSystem(body_model_data=synthetic.make_smpl_dict(0), n_samples=64, n_importance=64).to(dev)
for name, seed in (("nerf", 10), ("nerf_fine", 11)):
    getattr(sysm.anim_nerf, name).load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
an = sysm.anim_nerf
pa, tp = synthetic.make_body_params(1, seed=1)
an.setup_frame({k: torch.from_numpy(v).to(dev) for k, v in pa.items()}, {k: torch.from_numpy(v).to(dev) for k, v in tp.items()}, None)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
buf = torch.empty(N, N, N, device=dev)
def run():
    inference.query_density_grid(an, N, out=buf)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1)
t = _lib.enable_timing(True)
run(); torch.cuda.synchronize()
_lib.enable_timing(False)
ours = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in t.items()}
print("total %.2f ms; entry points: %s; glue (lattice generation, relu, copies): %.2f ms" % (total, {k: round(v, 2) for k, v in ours.items()}, total - sum(ours.values())))
center = (an.verts.max(dim=1)[0] + an.verts.min(dim=1)[0]) / 2.0
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
for a in range(0, N, 64):
    pts = inference.grid_slab_points(N, (-1.2, 1.2), (-1.2, 1.2), (-1.2, 1.2), center[0], a, min(N, a + 64), dev)
g1.record(); torch.cuda.synchronize()
print("lattice generation alone: %.2f ms" % g0.elapsed_time(g1))
