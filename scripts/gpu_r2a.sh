#!/bin/bash
# Round 2, call A: tcgen05 shape/mode probe, sanitizer passes over a test subset, encoder ncu capture.
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
timeout 120 build/mma_probe > $OUT/mma_probe.txt 2>&1; echo "probe exit $?"; cat $OUT/mma_probe.txt | head -80
timeout 300 python scripts/enc_conv_probe.py --iters 3 > $OUT/enc_conv_times.txt 2>&1; cat $OUT/enc_conv_times.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_umma -c 24 \
    -o $OUT/enc_prof -f python scripts/enc_conv_probe.py --iters 2 > $OUT/enc_ncu.log 2>&1; echo "ncu exit $?"
SUB='conv2d_matches_oracle or convlstm_cell_hoisted or convlstm_cell_matches_reference or e2e_b2_64x64_t3 or mask_head or class_stop'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SUB" \
    > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 $OUT/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "convlstm_cell_hoisted or e2e_b2_64x64_t3" \
    > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -5 $OUT/sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "convlstm_cell_hoisted or e2e_b2_64x64_t3" \
    > $OUT/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; tail -5 $OUT/sanitizer_synccheck.log
ls -la $OUT
