#!/bin/bash
# Wavefront decoder schedule A/B: parity of the end-to-end tests and bench numbers for RSIS_B200_PIPELINE = 0 / 1 / 2.
TAG=${1:-pipe}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for mode in 0 1 2; do
  export RSIS_B200_PIPELINE=$mode
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "e2e or sharding or cfg3 or cfg1" > $OUT/pytest_$mode.log 2>&1
  echo "== mode $mode pytest exit $? : $(tail -1 $OUT/pytest_$mode.log)"; grep -E "^(FAILED|ERROR)" $OUT/pytest_$mode.log | head -5
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$mode.json 2> $OUT/bench_$mode.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$mode.json"))
    print("mode $mode: value %.0f masks/s  %.3f ms/pass  e2e %.0f  launches/step %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["launches_per_step"]))
except Exception as e:
    print("mode $mode: bench failed", e); print(open("$OUT/bench_$mode.err").read()[-1500:])
PY
done
