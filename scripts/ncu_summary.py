"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote: per kernel launch duration, DRAM
bytes, L2 hit rate, achieved occupancy, issue-slot utilisation, top stall reasons.  Usage: ncu_summary.py rep [> profiles/x.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum"]
print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised by the profiler)")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"\n== {d.get('Kernel Name')}  (launch id {d.get('ID')})")
    for w in want:
        if w in d and d[w] != "":
            print(f"  {w:70s} {d[w]} {u[w]}")
    stalls = []
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                stalls.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))
