#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02f
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k fused_center_head -s --tb=short 2>&1 | grep -v "^$" | tail -n 30 > $out.head.txt
if grep -q failed $out.head.txt; then
  echo "=== CNB_HEAD_SWAP=1" >> $out.head.txt
  CNB_HEAD_SWAP=1 timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k fused_center_head -s --tb=line 2>&1 | grep -v "^$" | tail -n 20 >> $out.head.txt
fi
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "dcn_forward or graphed or dla34" -s --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $out.train.txt
timeout 600 python tools/profile_train.py 2 > $out.train_kernels.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $out.bench.json 2> $out.bench.err; tail -n 3 $out.bench.err
cat $out.head.txt; tail -n 25 $out.train.txt; head -n 14 $out.train_kernels.txt | cut -c1-150
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f.bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
t=d["train"]; print("train", t.get("value"), t.get("ms_per_step"), t.get("error"))
PY
