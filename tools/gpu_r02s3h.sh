#!/bin/bash
# decode: ring depth 2 (70 KB per CTA, 3 CTAs per SM) against 4 (105 KB, 2 CTAs per SM)
mkdir -p gpurun_out
o=gpurun_out/r02s3h.txt; : > $o
timeout 300 python -m pytest tests/test_conv_gpu.py -q -x -k "dcnv2_matches" 2>&1 | tail -n 2 >> $o
CNB_DECODE_NST=2 timeout 600 python -m pytest tests/test_decode_gpu.py -q -x 2>&1 | tail -n 3 >> $o
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
echo "== default (4 stages, 2 CTAs/SM, now 72 registers)" >> $o
timeout 200 python tools/decode_bench.py >> $o 2>&1
echo "== CNB_DECODE_NST=2 (3 CTAs/SM)" >> $o
CNB_DECODE_NST=2 timeout 200 python tools/decode_bench.py >> $o 2>&1
echo "== CNB_DECODE_NST=2 CNB_DECODE_CTAS=296" >> $o
CNB_DECODE_NST=2 CNB_DECODE_CTAS=296 timeout 200 python tools/decode_bench.py >> $o 2>&1
echo "== CNB_DECODE_NST=2 CNB_DECODE_WFLUSH=24" >> $o
CNB_DECODE_NST=2 CNB_DECODE_WFLUSH=24 timeout 200 python tools/decode_bench.py >> $o 2>&1
cat $o
