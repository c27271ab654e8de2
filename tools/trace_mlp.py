"""In-kernel timeline of the MLP forward kernel (needs a library built with AN_MLP_TRACE=1).
    AN_MLP_TRACE=1 python -m anim_nerf_b200._build --force ; python tools/trace_mlp.py [n] [--train]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import ops, synthetic, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 148 * 256 * 8
train = "--train" in sys.argv
dev = "cuda"
w = synthetic.make_nerf_weights(10)
ws = [torch.from_numpy(w[k + ".weight"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
bs = [torch.from_numpy(w[k + ".bias"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
packed = ops.mlp_pack(ws, bs)
xc = torch.rand(n, 3, device=dev) * 2 - 1
sigma = torch.empty(n, device=dev); rgb = torch.empty(n, 3, device=dev)
stash = ops.mlp_stash(n, dev) if train else None
lib = _lib.load()
buf = torch.zeros(4 * 65536, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.mlp_fwd(packed, xc, sigma, rgb, stash=stash)
torch.cuda.synchronize()
lib.an_debug_trace_fwd.argtypes = [ctypes.c_void_p]
assert lib.an_debug_trace_fwd(ctypes.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.mlp_fwd(packed, xc, sigma, rgb, stash=stash); e1.record()
torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1), "n", n, "train", train)
h = buf.cpu().numpy().astype(np.uint64)
ev = np.concatenate([h[r * 65536 + 1: r * 65536 + 1 + int(h[r * 65536])] for r in range(4)]); cnt = len(ev)
clk = (ev >> np.uint64(24)).astype(np.int64); role = ((ev >> np.uint64(20)) & np.uint64(15)).astype(int)
e = ((ev >> np.uint64(16)) & np.uint64(15)).astype(int); a = ((ev >> np.uint64(8)) & np.uint64(255)).astype(int); b = (ev & np.uint64(255)).astype(int)
t0 = clk.min(); clk -= t0
print("events", cnt, "span clk", clk.max())
# epilogue tiles: per layer g: wait-for-acc (ev0->ev1), store drain (ev1->ev2), body (ev2->ev3)
for t in (0, 1):
    m = role == 2 + t
    c, ee, aa = clk[m], e[m], a[m]
    o = np.argsort(c, kind="stable"); c, ee, aa = c[o], ee[o], aa[o]
    wait = {}; drain = {}; body = {}; gap = {}; loop = {}; stw = {}; nbar = {}
    last = {}; prev_end = None
    for ci, ei, gi in zip(c, ee, aa):
        if ei == 0:
            if prev_end is not None: gap.setdefault(gi, []).append(ci - prev_end)
            last[0] = ci
        elif ei == 1: wait.setdefault(gi, []).append(ci - last[0]); last[1] = ci
        elif ei == 2: drain.setdefault(gi, []).append(ci - last[1]); last[2] = ci
        elif ei == 3: body.setdefault(gi, []).append(ci - last[2]); prev_end = ci
        elif ei == 4: loop.setdefault(gi, []).append(ci - last[2]); last[4] = ci
        elif ei == 5: stw.setdefault(gi, []).append(ci - last[4]); last[5] = ci
        elif ei == 6: nbar.setdefault(gi, []).append(ci - last[5])
    print("tile %d  layer: wait_acc / store_drain / epilogue_body / gap-before(prologue at g=0)   [median clk]" % t)
    for g in range(10):
        print("   g=%d  %8.0f %8.0f %8.0f %8.0f   | block loop %6.0f  st-wait %5.0f  fence+named-bar %5.0f" % (
            g, np.median(wait.get(g, [0])), np.median(drain.get(g, [0])), np.median(body.get(g, [0])), np.median(gap.get(g, [0])),
            np.median(loop.get(g, [0])), np.median(stw.get(g, [0])), np.median(nbar.get(g, [0]))))
# MMA issuer: wait for act (ev0->ev1), per-chunk full waits
m = role == 1
c, ee, aa, bb = clk[m], e[m], a[m], b[m]
o = np.argsort(c, kind="stable"); c, ee, aa, bb = c[o], ee[o], aa[o], bb[o]
wact = {}; wfull = {}; issue = {}
prev = None
for ci, ei, gi, bi in zip(c, ee, aa, bb):
    if ei == 1: wact.setdefault((gi, bi), []).append(ci - prev)
    if ei == 2: wfull.setdefault((gi, bi >> 3), []).append(ci - prev)       # time inside the full-barrier wait
    if ei == 3 and prev is not None: issue.setdefault((gi, bi >> 3), []).append(ci - prev)   # previous step's issue + commit
    prev = ci
print("MMA issuer [median clk]: wait for act[g,t] | per ring step: wait for the weights (full barrier) | issue 4 MMAs + commit")
for g in range(10):
    print("   g=%d  act: %7.0f %7.0f   full-wait: %7.0f %7.0f   issue: %7.0f %7.0f" % (
        g, np.median(wact[(g, 0)]), np.median(wact[(g, 1)]), np.median(wfull[(g, 0)]), np.median(wfull[(g, 1)]),
        np.median(issue.get((g, 0), [0])), np.median(issue.get((g, 1), [0]))))
tot_iter = len(wact[(0, 0)])
print("iterations traced:", tot_iter, " clk/iteration:", clk.max() / tot_iter)
