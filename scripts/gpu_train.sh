#!/bin/bash
# GPU-box pass for the training step: whole-step gradient parity, training bench (both backward families), ncu launch list.
TAG=${1:-train}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "train_step or decoder_bptt" > $OUT/train_step.log 2>&1
echo "== train_step exit $? : $(tail -1 $OUT/train_step.log)"; grep -E "^(FAILED|ERROR)|AssertionError" $OUT/train_step.log | head
cat gpurun_out/grad_parity_*.json 2>/dev/null | grep -E "median|families|simt|auto" | head -30
RSIS_B200_BWD_IMPL=simt timeout 600 python bench_train.py --steps 5 --warmup 3 --cpu-steps 0 > $OUT/bench_train_simt.json 2> $OUT/bench_train_simt.err; cat $OUT/bench_train_simt.json; tail -3 $OUT/bench_train_simt.err
timeout 900 python bench_train.py --steps 5 --warmup 3 --cpu-steps 1 > $OUT/bench_train_auto.json 2> $OUT/bench_train_auto.err; cat $OUT/bench_train_auto.json; tail -3 $OUT/bench_train_auto.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/train_launches.csv \
    python bench_train.py --steps 1 --warmup 3 --cpu-steps 0 > $OUT/ncu_train.log 2>&1
python scripts/summarize_launches.py $OUT/train_launches.csv > $OUT/train_launches.md 2>&1; head -40 $OUT/train_launches.md
