#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02z2.txt; : > $o
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k dcn 2>&1 | tail -n 3 >> $o
for m in smooth far:1.2 far:4; do
  echo "== offsets=$m impl=fp" >> $o
  DCN_BENCH_OFFSETS=$m timeout 300 python tools/dcn_bench.py d64 d128 d256 d256_128 >> $o 2>&1
done
for i in 1 2; do
timeout 300 python tools/profile_layers.py 32 2>&1 | grep "^dcn\|^B=" > gpurun_out/r02z.layers.txt
echo "== in-network dcn layers: $(awk '/^dcn/{s+=$(NF-4)} END{print s}' gpurun_out/r02z.layers.txt) us total" >> $o
done
cat $o
