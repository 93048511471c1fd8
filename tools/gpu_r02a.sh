#!/bin/bash
# round-2 first GPU visit: new training / parity tests (one process per group: a trapped kernel kills its context), then
# the existing GPU suite and a bench line.
mkdir -p gpurun_out
out=gpurun_out/r02a
for t in conv_forward_backward stem_conv conv_bias batchnorm maxpool depthwise dcn_forward dla34_training; do
  echo "=== $t" >> $out.train.txt
  timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k $t -s --tb=short 2>&1 | tail -n 60 >> $out.train.txt
done
echo "=== conv_forward_backward with CNB_WGRAD_SWAP=1" >> $out.train.txt
CNB_WGRAD_SWAP=1 timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k conv_forward_backward -s --tb=line 2>&1 | tail -n 30 >> $out.train.txt
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_parity_e2e_gpu.py -q -m gpu -s --tb=short 2>&1 | tail -n 80 > $out.parity.txt
timeout 1500 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_train_gpu.py --deselect tests/test_dropin_gpu.py --deselect tests/test_parity_e2e_gpu.py 2>&1 | tail -n 60 > $out.tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $out.bench.json 2> $out.bench.err
tail -n 3 $out.bench.err
grep -c passed $out.train.txt; tail -n 5 $out.tests.txt; tail -n 15 $out.parity.txt; cat $out.bench.json | cut -c1-600
