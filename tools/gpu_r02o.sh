#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02o.txt; : > $o
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k dcn 2>&1 | tail -n 3 >> $o
for m in smooth random; do
  echo "== offsets=$m impl=fp" >> $o
  DCN_BENCH_OFFSETS=$m timeout 300 python tools/dcn_bench.py >> $o 2>&1
done
echo "== bf16 blend, smooth" >> $o
CNB_DCN_BLEND=bf16 timeout 300 python tools/dcn_bench.py >> $o 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 >> $o
cat $o
