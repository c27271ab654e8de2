"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: isolate one
training step (the launches between two consecutive body_joints_kernel launches, i.e. from the
per-frame table builder of one step to the next) and print per-kernel totals and shares.
    python tools/launch_summary.py gpurun_out/launches.csv [step_index_from_end] [launches_per_step] > profiles/rNN_launches_step.txt"""
import collections, csv, sys
path = sys.argv[1]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = []
with open(path, newline="") as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rd:
        rows.append((r[ki], float(r[vi].replace(",", "")) / 1e3))
marks = [i for i, (k, _) in enumerate(rows) if k.startswith("body_joints_kernel")]
# training steps have the same launch count; the 512^2 frames at the end of the run are shorter
seg = [(a, b) for a, b in zip(marks[:-1], marks[1:])]
lens = collections.Counter(b - a for a, b in seg)
L = max(lens, key=lambda n: (lens[n] > 1, n))
if len(sys.argv) > 3:          # explicit segment length: pick the render-only step / the step with regularisers
    L = min(lens, key=lambda n: abs(n - int(sys.argv[3])))
steps = [s for s in seg if s[1] - s[0] == L]
a, b = steps[-min(back, len(steps))]
agg = collections.OrderedDict()
for k, us in rows[a:b]:
    k = k.split("(")[0][:96]
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + us)
tot = sum(t for _, t in agg.values())
print("# ncu launch list, one training step of the bench command (segment %d..%d of %d launches)" % (a, b, len(rows)))
print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_ms_per_step, not absolutes")
print("launches in step: %d   sum of kernel time: %.1f us" % (b - a, tot))
print("%-96s %5s %12s %7s" % ("kernel", "n", "total_us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-96s %5d %12.1f %7.3f" % (k, n, t, t / tot))
