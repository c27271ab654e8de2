import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import ops, synthetic, _lib
n = 1 << 20
dev = "cuda"
w = synthetic.make_nerf_weights(10)
ws = [torch.from_numpy(w[k + ".weight"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
bs = [torch.from_numpy(w[k + ".bias"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
packed = ops.mlp_pack(ws, bs)
xc = torch.rand(n, 3, device=dev) * 2 - 1
sigma = torch.empty(n, device=dev); rgb = torch.empty(n, 3, device=dev)
stash = ops.mlp_stash(n, dev)
lib = _lib.load()
lib.an_debug_set.argtypes = [ctypes.c_int]
for train in (True, False):
    for flags in (0, 1, 2, 3, 4, 7, 8, 16, 32, 8 + 32, 63):
        lib.an_debug_set(flags)
        for _ in range(2):
            ops.mlp_fwd(packed, xc, sigma, rgb, stash=stash if train else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.mlp_fwd(packed, xc, sigma, rgb, stash=stash if train else None)
        e1.record(); torch.cuda.synchronize()
        print("train=%d flags=%2d  %.3f ms   (1 nomask, 2 nomaskSTG, 4 noTMAstore, 8 noSTTM, 16 noheads, 32 nobiasLDG)" % (train, flags, e0.elapsed_time(e1) / 5))
