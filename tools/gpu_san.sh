#!/bin/bash
# compute-sanitizer memcheck + racecheck over the GPU test selections of the hand-written kernels
mkdir -p gpurun_out
o=gpurun_out
SAN="compute-sanitizer --print-limit 5 --error-exitcode 0"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool" > $o/final.san_$tool.txt
  for sel in "tests/test_conv_gpu.py -k fused_center_head" "tests/test_train_gpu.py -k test_conv_forward_backward" "tests/test_train_gpu.py -k dcn_forward" "tests/test_conv_gpu.py -k dcnv2_matches" "tests/test_conv_gpu.py -k dcn_module" "tests/test_conv_gpu.py -k upsample" "tests/test_decode_gpu.py -k golden" "tests/test_conv_gpu.py -k rows"; do
    echo "--- pytest $sel" >> $o/final.san_$tool.txt
    timeout 900 $SAN --tool $tool python -m pytest $sel -q -m gpu -x --tb=line 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Program hit|Invalid|error:" | head -n 12 | cut -c1-240 >> $o/final.san_$tool.txt
  done
done
cat $o/final.san_memcheck.txt $o/final.san_racecheck.txt
