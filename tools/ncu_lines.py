"""Per-CUDA-source-line summary of one kernel of an .ncu-rep (needs -lineinfo + --import-source on):
instructions executed, stall samples, threads per instruction.
    python tools/ncu_lines.py rep.ncu-rep kernel-regex [launch-skip] [n_top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
iL, iS = 0, 1
col = {n: i for i, n in enumerate(h)}
iI, iT, iSm = col["Instructions Executed"], col["Thread Instructions Executed"], col["# Samples"]
agg = {}
cur = None
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    if r[iL] == "Line No":
        continue
    if r[iL]:
        cur = (int(r[iL]), r[iS])
        agg.setdefault(cur, [0.0, 0.0, 0.0])
    if cur is None or not r[2]:
        continue
    f = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
    a = agg[cur]
    a[0] += f(r[iI]); a[1] += f(r[iT]); a[2] += f(r[iSm])
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print("total warp-instructions %.3g, samples %d" % (ti, ts))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[5]) if len(sys.argv) > 5 else 2])[:ntop]:
    print("%5d  inst %5.1f%%  samples %5.1f%%  thr/inst %4.1f  %s" % (ln, 100 * a[0] / ti, 100 * a[2] / ts, a[1] / max(a[0], 1), src.strip()[:110]))
