#!/bin/bash
# repeat the conv / DCN parity selection to catch intermittent failures
mkdir -p gpurun_out
o=gpurun_out/flaky.txt; : > $o
for i in $(seq 1 14); do
  CNB_PDL=0 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "test_conv_matches_torch or dcn" -p no:cacheprovider > /tmp/f.txt 2>&1
  rc=$?
  echo "run $i PDL=0 rc=$rc $(tail -n 1 /tmp/f.txt)" >> $o
  if [ $rc -ne 0 ]; then tail -n 60 /tmp/f.txt | cut -c1-220 >> $o; fi
done
for i in $(seq 1 6); do
  timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "test_conv_matches_torch or dcn" -p no:cacheprovider > /tmp/f.txt 2>&1
  rc=$?
  echo "run $i PDL=1 rc=$rc $(tail -n 1 /tmp/f.txt)" >> $o
  if [ $rc -ne 0 ]; then tail -n 60 /tmp/f.txt | cut -c1-220 >> $o; fi
done
cat $o
