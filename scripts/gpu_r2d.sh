#!/bin/bash
# Round 2, call D: grouped-wavefront decoder + N-dependent planner model: parity tests, A/B bench.
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_parity.log 2>&1; echo "parity exit $?"; tail -5 $OUT/pytest_parity.log
for mode in 3 2; do for mm in 1 0; do
  RSIS_B200_PIPELINE=$mode RSIS_B200_MMA_MODEL=$mm timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_p${mode}_m${mm}.json 2> $OUT/bench_p${mode}_m${mm}.err
  echo "pipeline=$mode mma_model=$mm: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_p${mode}_m${mm}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])" 2>&1 | cut -c1-300)"
done; done
RSIS_B200_MMA_MODEL=1 timeout 200 python scripts/enc_conv_probe.py --iters 3 > $OUT/enc_conv_m1.txt 2>&1
RSIS_B200_MMA_MODEL=0 timeout 200 python scripts/enc_conv_probe.py --iters 3 > $OUT/enc_conv_m0.txt 2>&1
paste $OUT/enc_conv_m1.txt $OUT/enc_conv_m0.txt | awk '{print $1, $7, "us (new model) vs", $18, "us (old)"}'
