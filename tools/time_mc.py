"""Marching-cubes timing: a 512^3 sphere shell and a 512^3 blob field (count + scan + emit, incl. the D2H of the totals)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import mesh, _lib
N = 512
ax = torch.arange(N, device="cuda", dtype=torch.float32)
vol = torch.sqrt((ax.view(N, 1, 1) - 255.3) ** 2 + (ax.view(1, N, 1) - 256.1) ** 2 + (ax.view(1, 1, N) - 254.8) ** 2) - 180.2
for _ in range(2):
    v, f = mesh.marching_cubes(vol, 0.0)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5):
    v, f = mesh.marching_cubes(vol, 0.0)
torch.cuda.synchronize()
print("512^3 sphere: %.3f ms, %d vertices, %d faces" % ((time.time() - t0) / 5 * 1e3, v.shape[0], f.shape[0]))
t = _lib.enable_timing(True)
mesh.marching_cubes(vol, 0.0); torch.cuda.synchronize()
print({k: round(sum(a.elapsed_time(b) for a, b in e), 3) for k, e in t.items()})
