#!/bin/bash
# usage: bash scripts/quick_bench.sh <outfile-tag>; prints ms/step, e2e ms/step, cell step us and per-level us
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$1.json 2> gpurun_out/$1.err; tail -2 gpurun_out/$1.err
python -c "import json,sys; d=json.loads(open('gpurun_out/$1.json').read()); print('$1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['step_us'],1), [round(l['us'],1) for l in d['roofline']['per_level']])"
