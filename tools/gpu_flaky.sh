#!/bin/bash
# repeat parity selections to catch intermittent failures (full child output kept on failure)
mkdir -p gpurun_out
o=gpurun_out/flaky.txt; : > $o
for i in $(seq 1 10); do
  timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -p no:cacheprovider > /tmp/f.txt 2>&1
  rc=$?
  echo "run $i test_conv_gpu rc=$rc $(tail -n 1 /tmp/f.txt)" >> $o
  if [ $rc -ne 0 ]; then tail -n 80 /tmp/f.txt | cut -c1-220 >> $o; fi
done
for i in $(seq 1 4); do
  timeout 600 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py tests/test_train_gpu.py -x -q -m gpu -p no:cacheprovider > /tmp/f.txt 2>&1
  rc=$?
  echo "run $i model/engine/train rc=$rc $(tail -n 1 /tmp/f.txt)" >> $o
  if [ $rc -ne 0 ]; then tail -n 80 /tmp/f.txt | cut -c1-220 >> $o; fi
done
cat $o
