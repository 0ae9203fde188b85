#!/bin/bash
# Round-end evidence pass: whole GPU suite, smoke, bench.py (both arms), training bench, ncu launch lists (inference
# pass + training step) and --set full captures of the fused cell kernel and the tcgen05 weight-gradient kernel.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log; grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
timeout 900 python bench_train.py --steps 10 --warmup 3 --cpu-steps 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; cat $OUT/bench_train.json; tail -3 $OUT/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches.md 2>&1; head -12 $OUT/launches.md
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $OUT/train_launches.csv \
    python bench_train.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph > $OUT/ncu_train.log 2>&1
python scripts/summarize_launches.py $OUT/train_launches.csv > $OUT/train_launches.md 2>&1; head -16 $OUT/train_launches.md
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_umma_kernel<\(bool\)1|cell_swap_kernel' -s 20 -c 5 -o $OUT/prof_cell -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_cell.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:wgrad_umma_kernel' -s 400 -c 4 -o $OUT/prof_wgrad -f \
    python bench_train.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph > $OUT/ncu_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:soft_iou_partial_kernel' -s 6 -c 2 -o $OUT/prof_iou -f \
    python scripts/iou_probe.py > $OUT/ncu_iou.log 2>&1
ls -la $OUT
