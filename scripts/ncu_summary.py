"""Summarise an .ncu-rep (raw metrics + per-source-line instruction/stall shares).  Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum', 'sm__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__cycles_active.avg']
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h:70s} {units[i]:14s} {val[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg, smp, stall = {}, {}, {}
names = None
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        names = r
        continue
    if len(r) < 10 or names is None or r[2] != '-':
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur, ln, r[1][:100])
    agg[key] = agg.get(key, 0) + int(r[names.index("Instructions Executed")] or 0)
    smp[key] = smp.get(key, 0) + int(r[names.index("# Samples")] or 0)
    for i, n in enumerate(names):
        if n.startswith("stall_") and "Not Issued" not in n:
            stall[n] = stall.get(n, 0) + int(r[i] or 0)
tot, ts = sum(agg.values()), sum(smp.values())
print(f"\nsource lines by executed instructions (total {tot:.3e}, samples {ts}):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{100 * v / tot:5.1f}% inst {100 * smp[k] / max(ts, 1):5.1f}% smp  {k[0][:16]}:{k[1]:4d} {k[2]}")
st = sum(stall.values())
print("\nstall reasons:", ", ".join(f"{k[6:]} {100 * v / st:.1f}%" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
