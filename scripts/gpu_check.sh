#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list + one full capture of the cell kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh <tag> [quick]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$2" != "quick" ]; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref.json 2>> $OUT/bench.err
  cat $OUT/bench_ref.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k 'regex:conv_umma_kernel<\(bool\)1|cell_swap_kernel' -s 20 -c 5 -o $OUT/prof_cell -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_cell.log 2>&1
  ls -la $OUT
fi
