#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02h
timeout 600 python -m pytest tests/test_next_rows_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu -k "next or encoding or soft_nms or tta or dcn_forward or hourglass" -s --tb=short 2>&1 | grep -v "^$" | tail -n 30 > $out.next.txt; tail -n 12 $out.next.txt
timeout 300 python tools/profile_layers.py 32 > $out.layers.txt 2>&1; head -n 3 $out.layers.txt; grep -i "head\|dcn " $out.layers.txt | head -n 20
# sanitizers on small cases of every tensor-core / streaming kernel family
SAN="compute-sanitizer --print-limit 5 --error-exitcode 0"
for tool in memcheck racecheck; do
  echo "=== $tool" > $out.san_$tool.txt
  for sel in "tests/test_conv_gpu.py -k fused_center_head" "tests/test_train_gpu.py -k test_conv_forward_backward" "tests/test_train_gpu.py -k dcn_forward" "tests/test_conv_gpu.py -k dcn" "tests/test_decode_gpu.py -k golden" "tests/test_conv_gpu.py -k rows"; do
    echo "--- $sel" >> $out.san_$tool.txt
    timeout 900 $SAN --tool $tool python -m pytest $sel -q -m gpu -x --tb=line 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error:" | head -n 12 >> $out.san_$tool.txt
  done
done
cat $out.san_memcheck.txt $out.san_racecheck.txt
