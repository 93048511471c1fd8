#!/bin/bash
# plain mode stage skipping (library built with -DCNB_DCN_EXPERIMENTS): 1 no LDS, 2 (with 1) no STTM, 4 no MMAs, 64 no stores, 128 polling waits, 256 no box loads
mkdir -p gpurun_out
o=gpurun_out/r02s3f2.txt; : > $o
for dbg in 0 128 327 455; do
  echo "== CNB_DCN_DEBUG=$dbg" >> $o
  CNB_DCN_DEBUG=$dbg CNB_CONV_FP=1 timeout 120 python tools/tma_layers_bench.py off128 off64 c64_128 >> $o 2>&1
done
cat $o
