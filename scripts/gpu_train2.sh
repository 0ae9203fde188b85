#!/bin/bash
# wgrad-UMMA bring-up: primitive tests first (own processes), then whole-step parity + training bench + launch list.
TAG=${1:-train2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for grp in conv_wgrad conv_dgrad global_maxpool bn_train_bwd; do
  timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -k $grp > $OUT/$grp.log 2>&1
  echo "== $grp exit $? : $(tail -1 $OUT/$grp.log)"
  grep -E "^(FAILED|ERROR)|Error|assert |timed out" $OUT/$grp.log | head -12
done
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "train_step or decoder_bptt" > $OUT/train_step.log 2>&1
echo "== train_step exit $? : $(tail -1 $OUT/train_step.log)"; grep -E "^(FAILED|ERROR)|AssertionError" $OUT/train_step.log | head
grep -h "median" gpurun_out/grad_parity_*.json
timeout 900 python bench_train.py --steps 5 --warmup 3 --cpu-steps 0 > $OUT/bench_train_auto.json 2> $OUT/bench_train_auto.err; cat $OUT/bench_train_auto.json; tail -3 $OUT/bench_train_auto.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/train_launches.csv \
    python bench_train.py --steps 1 --warmup 3 --cpu-steps 0 > $OUT/ncu_train.log 2>&1
python scripts/summarize_launches.py $OUT/train_launches.csv > $OUT/train_launches.md 2>&1; head -30 $OUT/train_launches.md
