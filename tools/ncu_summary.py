"""Summarise an .ncu-rep: key raw metrics per kernel + top stall locations (SASS).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top] [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum.per_second", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct"]
names = []
for r in rows[2:]:
    nm = r[hdr.index("Kernel Name")]
    names.append(nm)
    if want and want not in nm:
        continue
    print("=" * 100)
    for h, u, v in zip(hdr, rows[1], r):
        if any(h == k or (k in h and len(k) > 25) for k in keys):
            print("  %-75s %s %s" % (h, v, u))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for nm, blk in zip(names, blocks[1:]):
    if want is None and nm != names[0]:
        continue
    if want and want not in nm:
        continue
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    h = lines[1]
    idx = {x: i for i, x in enumerate(h)}
    data = [l for l in lines[2:] if len(l) == len(h)]
    f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
    tot = sum(f(r[idx["# Samples"]]) for r in data) or 1
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    agg = {s: sum(f(r[idx[s]]) for r in data) for s in stalls}
    print("---- %s" % nm[:60])
    print("stall totals:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -f(r[idx["# Samples"]]))[:ntop]:
        s = sorted(((x, f(r[idx[x]])) for x in stalls), key=lambda kv: -kv[1])[:2]
        print("%6.1f%% %-80s %s" % (100 * f(r[idx["# Samples"]]) / tot, r[idx["Source"]][:80], s))
