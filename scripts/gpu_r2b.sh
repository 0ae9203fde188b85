#!/bin/bash
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 120 build/mma_probe > $OUT/mma_probe.txt 2>&1; echo "probe exit $?"
timeout 200 python scripts/hoist_probe.py > $OUT/hoist_plain.txt 2>&1; cat $OUT/hoist_plain.txt
RSIS_B200_SPLITK=0 timeout 200 python scripts/hoist_probe.py > $OUT/hoist_nosplit.txt 2>&1; cat $OUT/hoist_nosplit.txt
RSIS_B200_PDL=0 timeout 200 python scripts/hoist_probe.py > $OUT/hoist_nopdl.txt 2>&1; cat $OUT/hoist_nopdl.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/hoist_probe.py > $OUT/hoist_memcheck.txt 2>&1; grep -v "^=====" $OUT/hoist_memcheck.txt | tail -8
RSIS_B200_PRINT_PLAN=1 timeout 200 python scripts/hoist_probe.py 3,32,32,16,16,24 1 2>&1 | grep "rsis plan" > $OUT/hoist_plan.txt; cat $OUT/hoist_plan.txt
