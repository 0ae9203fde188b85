# bash scripts/quick_bench.sh <tag>: parity tests, the encoder per-convolution times, the bench without its slow legs
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_parity.log 2>&1; echo "parity exit $?"; tail -3 $OUT/pytest_parity.log
timeout 300 python scripts/enc_conv_probe.py --iters 3 > $OUT/enc_conv.txt 2>&1; cat $OUT/enc_conv.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-torch-gpu > $OUT/bench.json 2> $OUT/bench.err
python -c "import json,sys; d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_us'])"
