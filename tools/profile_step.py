"""Kernel-level breakdown of one cfg2 training step (shipped config: optim_body_params=True), torch kernels included:
torch.profiler over a few eager steps, grouped by kernel name.   python tools/profile_step.py [--frozen]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import anim_nerf_b200  # noqa
from anim_nerf_b200 import synthetic
from anim_nerf_b200.system import AnimNeRFSystem

frozen = "--frozen" in sys.argv
dev = torch.device("cuda", 0)
data, host, params, tmpl = bench.build_batch(0)
sysm = AnimNeRFSystem(body_model_data=data, n_samples=64, n_importance=64, num_frames=16, optim_body_params=not frozen).to(dev)
bench._load_nerfs(sysm)
sysm.init_body_model_params(params)
sysm.volume_renderer.device_rng = True
(opt,), _ = sysm.configure_optimizers()
flat = sysm.flat_grads
fi = torch.arange(16, device=dev)
tmpl_d = {k: v.to(dev) for k, v in tmpl.items()}
params_d = {k: v.to(dev) for k, v in params.items()}
b = {k: v.to(dev) for k, v in host.items()}
mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss


def step():
    flat.zero()
    p = params_d if frozen else sysm.body_model_params(fi)
    out = sysm(b["rays"], p, tmpl_d, perturb=1.0)
    loss = (mse(out["rgbs"], b["rgbs"]) + mse(out["rgbs_fine"], b["rgbs"]) + 0.1 * (l1(out["alphas"], b["alphas"]) + l1(out["alphas_fine"], b["alphas"])))
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] \
    if hasattr(prof.key_averages()[0], "device_type") else []
if not rows:
    rows = [(e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
rows.sort(key=lambda r: -r[1])
tot = 0.0
for k, t, c in rows[:45]:
    print("%9.1f us  x%5.1f  %s" % (t, c, k[:110]))
print("kernels/step:", sum(c for _, _, c in rows))
