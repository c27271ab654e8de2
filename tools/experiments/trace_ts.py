"""In-kernel timeline of the TMEM-operand MLP forward (needs a library built with AN_MLP_TRACE=1):
    AN_MLP_TRACE=1 AN_LIB_PATH=tools/_variants/libtrace.so python -m anim_nerf_b200._build
    AN_LIB_PATH=tools/_variants/libtrace.so python tools/trace_ts.py [first_iter]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import anim_nerf_b200  # noqa
from anim_nerf_b200 import ops, synthetic, _lib

n = 74 * 256 * 8
dev = "cuda"
w = synthetic.make_nerf_weights(10)
ws = [torch.from_numpy(w[k + ".weight"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
bs = [torch.from_numpy(w[k + ".bias"]).to(dev) for k in synthetic.NERF_LAYER_NAMES]
packed = ops.mlp_pack(ws, bs)
xc = torch.rand(n, 3, device=dev) * 2 - 1
sigma = torch.empty(n, device=dev); rgb = torch.empty(n, 3, device=dev)
lib = _lib.load()
buf = torch.zeros(4 * 65536, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.mlp_fwd(packed, xc, sigma, rgb)
torch.cuda.synchronize()
lib.an_debug_trace_ts.argtypes = [ctypes.c_void_p]
assert lib.an_debug_trace_ts(ctypes.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.mlp_fwd(packed, xc, sigma, rgb); e1.record()
torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1), "n", n)
h = buf.cpu().numpy().astype(np.uint64)
ev = np.concatenate([h[r * 65536 + 1: r * 65536 + 1 + int(h[r * 65536])] for r in range(4)])
clk = (ev >> np.uint64(24)).astype(np.int64); role = ((ev >> np.uint64(20)) & np.uint64(15)).astype(int)
e = ((ev >> np.uint64(16)) & np.uint64(15)).astype(int); a = ((ev >> np.uint64(8)) & np.uint64(255)).astype(int); b = (ev & np.uint64(255)).astype(int)
clk -= clk.min()
print("events", len(ev), "span clk", clk.max())
# one steady-state iteration of the MMA issuer: from the 4th bar_enc wait to the 5th
m1 = (role == 1) & (e == 5)
starts = np.sort(clk[m1])
k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lo, hi = starts[k], starts[k + 1]
print("iteration %d: %d clk" % (k, hi - lo))
names = {0: {0: "issue"}, 1: {0: "wait_a", 1: "got_a", 3: "wait_full", 2: "got_full", 4: "commit_acc", 5: "wait_enc", 6: "got_enc"},
         2: {0: "wait_acc", 1: "got_acc", 2: "ld_done", 3: "st+arrive"}, 3: {0: "wait_acc", 1: "got_acc", 2: "ld_done", 3: "st+arrive"}}
sel = (clk >= lo) & (clk < hi) & (role != 0)
o = np.argsort(clk[sel], kind="stable")
for c, r, ee, aa, bb in zip(clk[sel][o], role[sel][o], e[sel][o], a[sel][o], b[sel][o]):
    if aa > 2 and aa < 8: continue
    print("%8d  %s%-4s %-10s g=%d %s" % (c - lo, "      " * (r - 1), "R%d" % r, names[r].get(ee, ee), aa, ("h=%d kc=%d" % (bb >> 3, bb & 7)) if r == 1 and ee in (2, 3) else "h=%d" % bb))
