"""DRAM traffic per launch of our kernels from an `ncu --set full` capture:
    python tools/ncu_traffic.py gpurun_out/step_full.ncu-rep > profiles/ncu_traffic.json
Writes {C-ABI entry point: {"bytes_per_launch": mean of dram__bytes_read.sum + dram__bytes_write.sum over the
captured launches (bench.py's roofline averages the coarse and the fine launch the same way), "launches": n,
"kernel": device kernel name}}.  bench.py reads the file for `roofline.traffic`."""
import csv, io, json, subprocess, sys

ENTRY = {"mlp_fwd_tc_kernel": "an_mlp_fwd", "mlp_bwd_dgrad_kernel": "an_mlp_bwd_dgrad", "mlp_bwd_wgrad_kernel": "an_mlp_bwd_wgrad",
         "knn_search_kernel": "an_knn_unpose_fwd", "composite_fwd_kernel": "an_composite_fwd",
         "composite_bwd_kernel": "an_composite_bwd", "sample_fine_merge_kernel": "an_sample_fine_merge_fwd"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
acc = {}
for r in rows[2:]:
    name = r[ik]
    for k, entry in ENTRY.items():
        if k in name:
            b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
            acc.setdefault(entry, {"kernel": k, "v": []})["v"].append(b)
out = {e: {"bytes_per_launch": sum(d["v"]) / len(d["v"]), "launches": len(d["v"]), "kernel": d["kernel"]} for e, d in acc.items()}
out["_source"] = "ncu --set full --clock-control none capture of `bench.py --steps 2 --warmup 3` (tools/gpu_round.sh): %s" % rep
print(json.dumps(out, indent=1))
