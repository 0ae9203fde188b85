#!/bin/bash
TAG=${1:-caps}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export RSIS_B200_PIPELINE=2
run() {
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench.json"))
    print("split [$RSIS_B200_PIPE_SPLIT] caps [$RSIS_B200_PIPE_CAPS]: value %.0f masks/s  %.3f ms/pass  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("split [$RSIS_B200_PIPE_SPLIT] caps [$RSIS_B200_PIPE_CAPS]: bench failed", e); print(open("$OUT/bench.err").read()[-800:])
PY
}
for sp in "1,1,0,0,0" "1,0,0,0,0" "0,1,0,0,0" "1,1,1,0,0"; do
  export RSIS_B200_PIPE_SPLIT=$sp; export RSIS_B200_PIPE_CAPS=""; run
done
export RSIS_B200_PIPE_SPLIT="1,1,0,0,0"; export RSIS_B200_PIPE_CAPS="0,0,32,40,72"; run
