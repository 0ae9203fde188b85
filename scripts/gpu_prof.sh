#!/bin/bash
# ncu evidence for one tag: launch list of a short bench run + one `--set full` capture of the cell kernels.
# Usage (under gpurun, repo root): bash scripts/gpu_prof.sh <tag> [kernel-regex] [skip] [count]
TAG=${1:-r1}
KRE=${2:-'conv_umma_kernel<\(bool\)1>'}
SKIP=${3:-20}
CNT=${4:-5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$KRE" -s $SKIP -c $CNT -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
