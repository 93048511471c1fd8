#!/bin/bash
# conv_rows stage-skipping (library built with -DCNB_ROWS_EXPERIMENTS): which role bounds the thin layers
mkdir -p gpurun_out
o=gpurun_out/r02s3d.txt; : > $o
for dbg in 0 1 2 4 3 5 6 7; do
  echo "== CNB_ROWS_DEBUG=$dbg (1 no input copies, 2 no stores, 4 no MMAs)" >> $o
  CNB_ROWS_DEBUG=$dbg timeout 120 python tools/tma_layers_bench.py c64_128 off64 c16 c32 >> $o 2>&1
done
cat $o
