"""Prints a compact per-kernel table from an ncu report (`ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def col(n): return hdr.index(n) if n in hdr else None
M = [("us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("blk", "launch__block_size"),
     ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("Minst", "smsp__inst_executed.sum"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("tensor%", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
     ("tc%", "sm__inst_executed_pipe_tensor.sum"), ("st_lsb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("st_ssb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"), ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
     ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"), ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"), ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2hit", "lts__t_sector_hit_rate.pct")]
print("%-44s" % "kernel" + "".join("%9s" % s for s, _ in M))
for r in rows[2:]:
    k = r[col("Kernel Name")].replace("void ", "")[:42]
    vals = []
    for s, n in M:
        c = col(n)
        v = r[c] if c is not None else "na"
        try:
            f = float(v.replace(",", ""))
            if s == "Minst": f /= 1e6
            v = "%.1f" % f
        except ValueError:
            pass
        vals.append(v[:8])
    print("%-44s" % k + "".join("%9s" % v for v in vals))
