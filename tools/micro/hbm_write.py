"""Pure-write and pure-read HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (what bounds the stash / dY stores)."""
import torch
n = 1 << 30           # 4 GiB of fp32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
w = t(lambda: a.fill_(1.0)); print("fill  (write only) %.1f GB/s" % (4 * n / w / 1e6))
z = t(lambda: a.zero_());    print("zero  (memset)     %.1f GB/s" % (4 * n / z / 1e6))
r = t(lambda: a.sum());      print("sum   (read only)  %.1f GB/s" % (4 * n / r / 1e6))
c = t(lambda: b.copy_(a));   print("copy  (read+write) %.1f GB/s" % (8 * n / c / 1e6))
