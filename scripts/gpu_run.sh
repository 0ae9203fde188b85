#!/bin/bash
# One parameterised GPU session (run under gpurun from the repo root):  bash scripts/gpu_run.sh <tag> <step> [<step> ...]
# steps: tests | smoke | bench | bench_ref | bench_train | launches | ncu_cell | ncu_levels | pass_trace | sanitize | decoder_probe
TAG=${1:?tag}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for step in "$@"; do
  case $step in
    tests)   timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
             tail -3 $OUT/pytest.log; grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head ;;
    smoke)   timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log ;;
    bench)   timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
             cut -c1-1500 $OUT/bench.json; tail -2 $OUT/bench.err ;;
    bench_ref) timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_ref.json ;;
    bench_train) timeout 900 python bench_train.py --steps 10 --warmup 3 --cpu-steps 1 > $OUT/bench_train.json 2> $OUT/bench_train.err
             cut -c1-1700 $OUT/bench_train.json; tail -3 $OUT/bench_train.err ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
                 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-torch-gpu > $OUT/ncu_bench.log 2>&1
             python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches.md 2>&1; head -14 $OUT/launches.md ;;
    ncu_cell) timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
                 -k "regex:cell_group_kernel" -s 8 -c 3 -o $OUT/prof_cell -f \
                 python scripts/decoder_probe.py 8 256 256 10 3 > $OUT/ncu_cell.log 2>&1; ls -la $OUT/prof_cell.ncu-rep ;;
    sanitize) SUB='conv2d_matches_oracle or convlstm_cell or e2e_b2_64x64_t3 or mask_head or class_stop'
             timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "$SUB" \
                 > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $OUT/sanitizer_memcheck.log
             timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "convlstm_cell_hoisted" \
                 > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 $OUT/sanitizer_racecheck.log
             timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "convlstm_cell_hoisted or e2e_b2_64x64_t3" \
                 > $OUT/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 $OUT/sanitizer_synccheck.log ;;
    decoder_probe) timeout 300 python scripts/decoder_probe.py > $OUT/decoder_probe.txt 2>&1; cat $OUT/decoder_probe.txt ;;
    ncu_levels) # the grouped cell launch of ONE level at a time (all 148 CTAs) and of all five, one full capture each
             for lv in 0 1 2 3 4 all; do
               spec="-;-@$lv"; [ $lv = all ] && spec="-;-"
               timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
                 -k "regex:cell_group_kernel" -s 13 -c 1 -o $OUT/prof_level_$lv -f python scripts/group_tune.py "$spec" > $OUT/ncu_level_$lv.log 2>&1
             done; ls -la $OUT/*.ncu-rep ;;
    pass_trace) # needs the -DRSIS_DEBUG_TIMING build at build/librsis_dbg.so (scripts/pass_trace.py)
             TRACE_LOG=$OUT/trace_shapes.txt RSIS_B200_LIB=$PWD/build/librsis_dbg.so timeout 300 python scripts/pass_trace.py > $OUT/pass_trace.txt 2> $OUT/pass_trace.err
             tail -2 $OUT/pass_trace.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la $OUT
