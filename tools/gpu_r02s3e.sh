#!/bin/bash
# plain mode: box ring depth
mkdir -p gpurun_out
o=gpurun_out/r02s3e.txt; : > $o
CNB_CONV_FP=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -n 4 >> $o
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
for nfp in 2 3 4 6; do
echo "== CNB_CONV_FP=1 boxes=$nfp" >> $o
CNB_CONV_FP_BOXES=$nfp CNB_CONV_FP=1 timeout 300 python tools/tma_layers_bench.py off128 c128_64o off64 c64_128 off256 c64 >> $o 2>&1
done
echo "== CNB_CONV_FP=0" >> $o
CNB_CONV_FP=0 timeout 300 python tools/tma_layers_bench.py off128 c128_64o off64 c64_128 off256 c64 >> $o 2>&1
timeout 300 python tools/dcn_bench.py d64 d128 >> $o 2>&1
cat $o
