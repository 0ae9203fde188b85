#!/bin/bash
# Round-end evidence pass (final code): whole GPU suite, smoke, bench.py (both arms), training bench, launch lists.
TAG=${1:-final3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log; grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
cat $OUT/bench.json | cut -c1-1300; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_ref.json
timeout 900 python bench_train.py --steps 10 --warmup 3 --cpu-steps 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; cat $OUT/bench_train.json | cut -c1-1700; tail -3 $OUT/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches.md 2>&1; head -12 $OUT/launches.md
