#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02t.txt; : > $o
CNB_CONV_FP=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -n 8 >> $o
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
echo "== layers default (no fp)" >> $o
timeout 300 python tools/tma_layers_bench.py c64_128 off64 off128 c128_64o off256 off512 c64 c128 >> $o 2>&1
echo "== layers CNB_CONV_FP=1" >> $o
CNB_CONV_FP=1 timeout 300 python tools/tma_layers_bench.py c64_128 off64 off128 c128_64o off256 off512 c64 c128 >> $o 2>&1
echo "== dcn" >> $o
timeout 300 python tools/dcn_bench.py d64 d128 >> $o 2>&1
cat $o
