#!/bin/bash
TAG=${1:-caps}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export RSIS_B200_PIPELINE=2
run() {
  timeout 300 python bench.py --workload ${WL:-cfg2} --steps ${ST:-30} --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench.json"))
    print("${WL:-cfg2} caps [$RSIS_B200_PIPE_CAPS]: value %.0f masks/s  %.3f ms/pass  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("caps [$RSIS_B200_PIPE_CAPS]: bench failed", e); print(open("$OUT/bench.err").read()[-800:])
PY
}
for caps in "0,0,24,40,80" "0,0,32,32,80" "0,0,32,48,64" "0,0,48,40,56" "0,0,16,32,96" "0,8,20,44,64" "0,0,32,40,64" "0,0,28,36,72"; do
  export RSIS_B200_PIPE_CAPS=$caps; run
done
export WL=cfg5 ST=8
export RSIS_B200_PIPE_CAPS=""; run
export RSIS_B200_PIPE_CAPS="0,0,32,40,72"; run
export WL=cfg3 ST=10
export RSIS_B200_PIPE_CAPS=""; run
export RSIS_B200_PIPE_CAPS="0,0,32,40,72"; run
