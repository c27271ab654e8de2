"""Per-kernel CUDA-event breakdown of one 512x512 inference frame (cfg3) and of the 512^3 grid query (cfg4).
    python tools/profile_frame.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import anim_nerf_b200  # noqa
from anim_nerf_b200 import synthetic, inference, _lib
from anim_nerf_b200.system import AnimNeRFSystem

dev = torch.device("cuda", 0)
data, host, params, tmpl = bench.build_batch(0)
sysm = AnimNeRFSystem(body_model_data=data, n_samples=64, n_importance=64, num_frames=16, optim_body_params=False).to(dev)
bench._load_nerfs(sysm)
vr, an = sysm.volume_renderer, sysm.anim_nerf
p1 = {k: v[:1].to(dev) for k, v in params.items()}
t1 = {k: v[:1].to(dev) for k, v in tmpl.items()}
for H in (512, 1080):
    cam = synthetic.make_camera(H, H)
    cam_d = [torch.from_numpy(cam[k])[None].to(dev) for k in ("c2w", "focal", "c")]
    f = lambda: inference.render_frame(vr, an, cam_d[0], cam_d[1], cam_d[2], H, H, p1, t1)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t = _lib.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    _lib.enable_timing(False)
    print("frame %d^2: %.3f ms/frame; kernels:" % (H, e0.elapsed_time(e1) / 5),
          {k: round(sum(a.elapsed_time(b) for a, b in v) / 5, 3) for k, v in sorted(t.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1]))})
