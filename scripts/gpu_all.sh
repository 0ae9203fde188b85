#!/bin/bash
# Whole GPU suite + training bench (graph) + soft-IoU roofline.
TAG=${1:-all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log; grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head -20
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -4 $OUT/smoke.log
timeout 900 python bench_train.py --steps 10 --warmup 3 --cpu-steps 0 > $OUT/bench_train.json 2> $OUT/bench_train.err; cat $OUT/bench_train.json; tail -5 $OUT/bench_train.err
