#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02y.txt; : > $o
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_train_gpu.py -q -x -k "upsample or dw or up" 2>&1 | tail -n 4 >> $o
for impl in 3 4; do
  echo "== CNB_DW_DECONV_IMPL=$impl" >> $o
  CNB_DW_DECONV_IMPL=$impl timeout 300 python tools/up_bench.py >> $o 2>&1
done
cat $o
