#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02b
for t in conv_forward_backward conv_bias dcn_forward dla34_training; do
  echo "=== $t" >> $out.train.txt
  timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k $t -s --tb=short 2>&1 | tail -n 70 >> $out.train.txt
done
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_parity_e2e_gpu.py tests/test_conv_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py -q -m gpu -s --tb=short 2>&1 | grep -v "^$" | tail -n 80 > $out.parity.txt
grep -n "passed\|failed" $out.train.txt; tail -n 40 $out.parity.txt
