"""Driver for ncu captures / quick timing of the MLP kernels on n random canonical points.
    python tools/profile_mlp.py [n] [--train] [--bwd]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import ops, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
train = "--train" in sys.argv or "--bwd" in sys.argv
bwd = "--bwd" in sys.argv
dev = "cuda"
w = synthetic.make_nerf_weights(10)
ws = [torch.from_numpy(w[k + ".weight"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
bs = [torch.from_numpy(w[k + ".bias"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
packed = ops.mlp_pack(ws, bs)
xc = torch.rand(n, 3, device=dev) * 2 - 1
sigma = torch.empty(n, device=dev); rgb = torch.empty(n, 3, device=dev)
stash = ops.mlp_stash(n, dev) if train else None
gs = torch.randn(n, device=dev); grgb = torch.randn(n, 3, device=dev)

def run():
    ops.mlp_fwd(packed, xc, sigma, rgb, stash=stash)
    if bwd:
        ops.mlp_bwd(packed, stash, xc, rgb, gs, grgb)

for _ in range(2):
    run()
torch.cuda.synchronize()
from anim_nerf_b200 import _lib
t = _lib.enable_timing(True)
for _ in range(5):
    run()
torch.cuda.synchronize()
for k, v in t.items():
    ms = np.mean([a.elapsed_time(b) for a, b in v])
    print("%-20s %8.3f ms  %7.1f TFLOP/s (1.18 MFLOP/pt)" % (k, ms, n * 1179904 / ms / 1e9))
