"""ncu --csv metric rows of the bench workload's solve launches -> profiles/roofline_r02.json entry.
    python scripts/ncu_roofline.py <ncu.csv> <signature> <source-label> [--peaks=issue_peak.json] [--out=profiles/roofline_r02.json]
                                   [--library="<ged_version() of the profiled build>"]
The csv is the log of   ncu --csv --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,... -k regex:ged_solve_kernel -s 3 -c 3"""
import csv
import json
import os
import sys

path, sig, source = sys.argv[1], sys.argv[2], sys.argv[3]
out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_r02.json"))
peaks_file = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--peaks=")), None)
library = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--library=")), None)
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
col = {name: hdr.index(name) for name in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value")}
launches = {}
for r in rows[1:]:
    if r == hdr:
        continue
    d = launches.setdefault(r[col["ID"]], {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
per_class = []
for k in sorted(launches, key=int):
    d = launches[k]
    per_class.append({"kernel": d["kernel"], "grid": d["grid"], "block": d["block"],
                      "warp_instructions": d.get("smsp__inst_executed.sum"),
                      "alu_warp_instructions": d.get("sm__inst_executed_pipe_alu.sum"),
                      "fp64_warp_instructions": d.get("sm__inst_executed_pipe_fp64.sum"),
                      "fma_warp_instructions": d.get("sm__inst_executed_pipe_fma.sum"),
                      "duration_ms_alone": d.get("gpu__time_duration.sum", 0) / 1e6,
                      "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "warps_active_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "dram_bytes": (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))})
entry = {"source": source, "library": library, "per_class": per_class,
         "warp_instructions_per_step": sum(c["warp_instructions"] for c in per_class),
         "alu_warp_instructions_per_step": sum(c["alu_warp_instructions"] for c in per_class),
         "dram_bytes_per_step": sum(c["dram_bytes"] for c in per_class)}
doc = {}
if os.path.exists(out):
    doc = json.load(open(out))
doc.setdefault("workloads", {})[sig] = entry
if peaks_file:
    doc["issue_peaks"] = json.load(open(peaks_file))
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(entry, indent=1))
