#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one row per captured launch with the roofline-relevant counters.
usage: ncu_summary.py <report.ncu-rep> <title> > profiles/<name>.md"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_op_red.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [h.index(w) for w in want if w in h]
print(f"# {title}\n")
print("`ncu --set full --clock-control none --import-source on`; replays are cold-cache and serialised "
      "(report itself: gpurun_out/, scratch).\n")
print("| # | " + " | ".join(h[i] + " [" + u[i] + "]" for i in idx) + " |")
print("|---|" + "---|" * len(idx))
for n, r in enumerate(data):
    print(f"| {n} | " + " | ".join((r[i][:60] if h[i] == "Kernel Name" else r[i][:14]) for i in idx) + " |")
