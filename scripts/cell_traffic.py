#!/usr/bin/env python
"""Round-1 helper: profiles/cell_traffic.json + a markdown summary from one `ncu --set full` capture of the five
SINGLE cell launches of a decoder step.  Since round 2 the dominant kernel is the grouped launch (`cell_group_kernel`);
its capture is `scripts/gpu_run.sh <tag> ncu_cell`, summarised by scripts/ncu_summary.py, and profiles/cell_traffic.json
holds its per-launch DRAM reads and writes.  usage: cell_traffic.py <prof.ncu-rep> <tag> [workload]"""
import csv, io, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, data = rows[0], rows[1], rows[2:]
def col(name):
    return h.index(name)
def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m[unit]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
per = [to_bytes(r[rd], u[rd]) + to_bytes(r[wr], u[wr]) for r in data]
out = {"workload": workload, "dram_bytes_per_step": sum(per), "per_launch": per,
       "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} ({tag}); cold-cache replays"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "cell_traffic.json"), "w"), indent=1)
want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [h.index(w) for w in want if w in h]
with open(os.path.join(root, "profiles", f"{tag}_cell_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the fused ConvLSTM cell kernels\n\n")
    f.write("Five launches = levels 0-4 of one decoder step (BASELINE.json configs[1]: B=8, 256x256); scripts/gpu_check.sh.\n")
    f.write("Source report: gpurun_out (scratch, not tracked).  ncu replays are cold-cache and serialised.\n\n")
    f.write("| level | " + " | ".join(h[i] + " [" + u[i] + "]" for i in idx) + " |\n|---|" + "---|" * len(idx) + "\n")
    for l, r in enumerate(data):
        f.write(f"| {l} | " + " | ".join(r[i][-40:] if h[i] == "Kernel Name" else r[i][:12] for i in idx) + " |\n")
    f.write(f"\nDRAM traffic per decoder step (read + write, 5 launches): {sum(per)/1e6:.1f} MB "
            f"(algorithmic, SURVEY 8d definition: 72.9 MB).\n")
print(json.dumps(out))
