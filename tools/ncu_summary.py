"""Turn gpurun_out/*.ncu-rep and launch CSVs into the text summaries kept under profiles/.

usage: python tools/ncu_summary.py <report.ncu-rep> <launches.csv> <out.md> [title]
Runs here (no GPU): ncu -i only reads the report."""
import collections
import csv
import io
import subprocess
import sys

rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]
title = sys.argv[4] if len(sys.argv) > 4 else rep

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_sample_count",
]

lines = ["# " + title, ""]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) >= 3:
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += ["## ncu --set full: `%s`" % name[:120], "", "| metric | unit | value |", "|---|---|---|"]
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                lines.append("| %s | %s | %s |" % (h, u, v))
        lines.append("")

if launches and launches != "-":
    rws = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rws[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rws[1:]:
        n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines += ["## launch list (ncu --metrics gpu__time_duration.sum, cold-cache, serialised: compare shares)", "",
              "| kernel | launches | total %s | share |" % rws[1][ui], "|---|---|---|---|"]
    for n, a in agg.items():
        lines.append("| %s | %d | %.0f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / tot))
    lines.append("")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            agg.append((cur_file, int(r[0]), r[1], int(r[6]), int(r[7]), int(r[8])))
        except ValueError:
            pass
if agg:
    ts = sum(a[3] for a in agg) or 1
    ti = sum(a[4] for a in agg) or 1
    agg.sort(key=lambda a: -(a[4] / ti + a[3] / ts))
    lines += ["## hottest source lines (share of stall samples / of executed warp instructions)", "",
              "| file:line | samples | instructions | threads/inst | source |", "|---|---|---|---|---|"]
    for a in agg[:16]:
        lines.append("| %s:%d | %.1f%% | %.1f%% | %.1f | `%s` |" % (a[0], a[1], 100 * a[3] / ts, 100 * a[4] / ti,
                                                                  a[5] / max(1, a[4]), a[2].strip()[:90].replace("|", "\\|")))
    lines.append("")
open(out, "w").write("\n".join(lines))
print("wrote", out)
