#!/bin/bash
o=gpurun_out/r02m.dcn_dbg.txt; : > $o
for dbg in 0 63 15 16 1 2 4 8 32; do
  echo "== CNB_DCN_DEBUG=$dbg" >> $o
  CNB_DCN_DEBUG=$dbg timeout 300 python tools/dcn_bench.py d64 d128 >> $o 2>&1
done
cat $o
