#!/usr/bin/env python
"""Write a tracked markdown summary of one ncu report: key raw metrics + the source-level stall digest.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/ncu_x_r2.md "title" [min_exec]"""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
min_exec = sys.argv[4] if len(sys.argv) > 4 else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
lines = ["# " + title, "", "kernel: `%s`" % name[:200], "", "| metric | unit | value |", "|---|---|---|"]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        lines.append("| %s | %s | %s |" % (k, units[i], r[i]))
dig = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_src.py"), rep] + ([min_exec] if min_exec else []),
                     capture_output=True, text=True).stdout
lines += ["", "Source-level digest (`ncu --page source`, warp-state samples):", "", "```", dig.rstrip(), "```", ""]
open(out, "w").write("\n".join(lines))
print("wrote", out)
