"""Per-kernel summary of an `ncu --set full` report: time, DRAM traffic, issue / occupancy, instruction count, top stall reasons.
usage: ncu_summary.py <report.ncu-rep> [kernel-name-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else None
# optional: --traffic <workload> <frames> <source text>  writes profiles/threshold_traffic.json[workload] from the LAST matching kernel
traffic = None
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    traffic = (sys.argv[i + 1], int(sys.argv[i + 2]), sys.argv[i + 3])
    filt = filt if filt and not filt.startswith("--") else "threshold"
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.split("\n")))
h = rows[0]
col = {c: i for i, c in enumerate(h)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
units = rows[1]
for r in rows[2:]:
    if len(r) < len(h):
        continue
    name = r[col["Kernel Name"]]
    if filt and filt not in name:
        continue
    print(f"Kernel Name  {name[:150]}")
    for k in KEYS:
        if k in col:
            print(f"  {k:<70} {r[col[k]]} {units[col[k]]}")
    stalls = []
    for c, i in col.items():
        if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, nme in sorted(stalls, reverse=True)[:6]:
        print(f"  stall {nme:<40} {v:.3f} warps per issue-active cycle")
    print()
if traffic:
    import json, os
    last = [r for r in rows[2:] if len(r) >= len(h) and filt in r[col["Kernel Name"]]][-1]
    def _bytes(k):
        v, u = float(last[col[k]]), units[col[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    total = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "threshold_traffic.json")
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        d = {}
    d[traffic[0]] = {"bytes_per_frame": total / traffic[1], "source": traffic[2]}
    json.dump(d, open(path, "w"), indent=1)
    print("threshold traffic", traffic[0], total / traffic[1], "bytes per frame")
