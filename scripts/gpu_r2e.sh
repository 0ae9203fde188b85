#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout 300 python scripts/decoder_probe.py > $OUT/decoder_probe.txt 2>&1; cat $OUT/decoder_probe.txt
RSIS_B200_PRINT_PLAN=1 timeout 300 python scripts/decoder_probe.py 8 256 256 2 3 2>&1 | grep -E "rsis (plan|group)" | sort | uniq -c | sort -rn | head -40 > $OUT/plans.txt; cat $OUT/plans.txt
