"""Summarise `ncu --set full` reports as markdown tables.  usage: python tools/ncu_summary.py rep1.ncu-rep [rep2 ...]
(reads each report with `ncu -i <rep> --page raw --csv`; one table per captured kernel)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("## %s: unreadable" % rep)
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("## %s  (`%s`)\n" % (d.get("Kernel Name", "?")[:80], rep.split("/")[-1]))
        print("| metric | value |\n|---|---|")
        for k in WANT:
            if k in d and d[k] != "":
                print("| `%s` | %s %s |" % (k, d[k], u.get(k, "")))
        print()
