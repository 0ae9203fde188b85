#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_parity.log 2>&1; echo "parity exit $?"; tail -5 $OUT/pytest_parity.log
timeout 300 python scripts/decoder_probe.py > $OUT/decoder_probe.txt 2>&1; cat $OUT/decoder_probe.txt
RSIS_B200_CELL_ROWS=0 timeout 300 python scripts/decoder_probe.py 8 256 256 10 3,2 > $OUT/decoder_probe_rows0.txt 2>&1; head -3 $OUT/decoder_probe_rows0.txt; grep -E "grouped|group of" $OUT/decoder_probe_rows0.txt
for mode in 3 2; do
  RSIS_B200_PIPELINE=$mode timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_p${mode}.json 2> $OUT/bench_p${mode}.err
  echo "pipeline=$mode: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_p${mode}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" 2>&1 | cut -c1-300)"
done
