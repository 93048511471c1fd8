#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02e
for t in "dcn_forward or depthwise" "dla34 or reduces" "graphed"; do
  echo "=== $t" >> $out.tests.txt
  timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "$t" -s --tb=short 2>&1 | grep -v "^$" | tail -n 30 >> $out.tests.txt
done
timeout 900 python -m pytest tests/test_parity_e2e_gpu.py -q -m gpu -s --tb=short 2>&1 | grep -v "^$" | tail -n 30 >> $out.tests.txt
timeout 600 python tools/profile_train.py 2 > $out.train_kernels.txt 2>&1
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > $out.bench_train.json 2> $out.bench_train.err; tail -n 3 $out.bench_train.err
cat $out.tests.txt | tail -n 60; head -n 32 $out.train_kernels.txt | cut -c1-150
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e.bench_train.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["train"]["launch"], d["gpu_launches"])
PY
