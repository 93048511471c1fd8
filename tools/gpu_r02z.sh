#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02z.txt; : > $o
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k dcn 2>&1 | tail -n 3 >> $o
for bw in e r; do for m in smooth random; do
  echo "== boxw=$bw offsets=$m" >> $o
  CNB_DCN_BOXW=$bw DCN_BENCH_OFFSETS=$m timeout 300 python tools/dcn_bench.py d64 d128 d256 d256_128 >> $o 2>&1
done; done
for bw in e r; do
  CNB_DCN_BOXW=$bw timeout 300 python tools/profile_layers.py 32 2>&1 | grep "^dcn\|^B=" > gpurun_out/r02z.layers_$bw.txt
  echo "== in-network dcn layers boxw=$bw: $(awk '/^dcn/{s+=$(NF-4)} END{print s}' gpurun_out/r02z.layers_$bw.txt) us total" >> $o
done
cat $o
