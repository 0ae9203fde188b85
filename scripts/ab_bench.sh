# bash scripts/ab_bench.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...: the bench's resident / e2e numbers under each environment
OUT=gpurun_out/${1:-ab}; shift
mkdir -p $OUT
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-torch-gpu > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python -c "import json,sys; d=json.loads(open('$OUT/bench_$i.json').read().strip().splitlines()[-1]); print('$envs', '| masks/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'cell us', round(d['roofline']['step_us'],1), 'frac', round(d['roofline']['frac'],4))"
done
