#!/bin/bash
TAG=${1:-train4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q > $OUT/backward.log 2>&1
echo "== backward suite exit $? : $(tail -1 $OUT/backward.log)"; grep -E "^(FAILED|ERROR)|Error" $OUT/backward.log | head
timeout 900 python bench_train.py --steps 10 --warmup 3 --cpu-steps 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; cat $OUT/bench_train.json; tail -5 $OUT/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/train_launches.csv \
    python bench_train.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph > $OUT/ncu_train.log 2>&1
python scripts/summarize_launches.py $OUT/train_launches.csv > $OUT/train_launches.md 2>&1; head -24 $OUT/train_launches.md
