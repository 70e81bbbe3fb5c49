"""Key metrics of every launch in one or more ncu reports as one CSV (tools only):
    python tools/ncu_summary.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
cols = []
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        cols.append((f"{name} [{rep.split('/')[-1]}#{d['ID']}]", {m: (d.get(m, ""), u.get(m, "")) for m in METRICS}))
with open(sys.argv[1], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric"] + [c[0] for c in cols])
    for m in METRICS:  # ncu picks the unit per report, so every value carries its own
        w.writerow([m] + [(c[1][m][0] + " " + c[1][m][1]).strip() for c in cols])
print("wrote", sys.argv[1], len(cols), "launches")
