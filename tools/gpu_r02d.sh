#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02d
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_e2e_gpu.py -q -m gpu -k "dla34 or e2e or benchmark" -s --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $out.tests.txt
timeout 600 python tools/profile_train.py 2 > $out.train_kernels.txt 2>&1
cat $out.tests.txt | tail -n 30; head -n 50 $out.train_kernels.txt
