#!/bin/bash
# GPU-box pass for the training path: per-primitive backward parity (one pytest process per group so a faulting kernel
# cannot poison the rest), whole-step gradient parity, then the existing forward suite.
TAG=${1:-bwd}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for grp in conv_wgrad conv_dgrad bn_train_bwd maxpool_bwd upsample_bilinear_bwd lstm_gates global_maxpool class_stop_heads_bwd decoder_bptt train_step; do
  timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -k $grp > $OUT/$grp.log 2>&1
  echo "== $grp exit $? : $(tail -1 $OUT/$grp.log)"
  grep -E "^(FAILED|ERROR)|Error|assert " $OUT/$grp.log | head -12
done
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_backward.py > $OUT/forward.log 2>&1
echo "== forward suite exit $? : $(tail -1 $OUT/forward.log)"
