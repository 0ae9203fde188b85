#!/bin/bash
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -k regex:"cell_group|upsample|class_stop" -c 60 --csv --log-file $OUT/decoder_launches.csv python scripts/decoder_probe.py 8 256 256 10 3 > $OUT/ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2m/decoder_launches.csv")) if len(r) > 10]
h = rows[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iid = h.index("ID")
cur = {}
out = []
for r in rows[1:]:
    key = (r[iid], r[ik][:40])
    cur.setdefault(key, {})[r[im]] = r[iv]
for (i, k), m in cur.items():
    print(i, k, "us=%.1f" % (float(m.get("gpu__time_duration.sum", "0").replace(",", "")) / 1e3 if float(m.get("gpu__time_duration.sum","0").replace(",","")) > 1000 else float(m.get("gpu__time_duration.sum","0").replace(",",""))),
          "rd=%s wr=%s tensor%%=%s" % (m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"), m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
PY
