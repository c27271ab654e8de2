"""HBM write-only / read-only / copy bandwidth on this GPU (torch fill_, sum, copy_ over 4 GiB; best of 10)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def best(fn, nbytes):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / (min(t) * 1e-3) / 1e9
print("write-only (fill_)   %.0f GB/s" % best(lambda: a.fill_(1.0), 4 * n))
print("read-only  (sum)     %.0f GB/s" % best(lambda: a.sum(), 4 * n))
print("copy (read+write)    %.0f GB/s" % best(lambda: b.copy_(a), 8 * n))
