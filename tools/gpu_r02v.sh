#!/bin/bash
# DCN blend modes: accuracy vs torchvision fp32, parity tests, speed
mkdir -p gpurun_out
o=gpurun_out/r02v.txt; : > $o
for m in fp32 wbf16 bf16; do
  echo "== accuracy CNB_DCN_BLEND=$m" >> $o
  CNB_DCN_BLEND=$m timeout 300 python tools/dcn_debug.py 2>&1 | tail -n 9 >> $o
done
echo "== parity tests wbf16" >> $o
CNB_DCN_BLEND=wbf16 timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_parity_e2e_gpu.py -q -x -s 2>&1 | grep -i "rel\|passed\|failed\|overlap\|error" | tail -n 30 >> $o
echo "== parity tests fp32 (same prints)" >> $o
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_e2e_gpu.py -q -x -s 2>&1 | grep -i "rel\|passed\|failed\|overlap\|error" | tail -n 30 >> $o
for m in fp32 wbf16; do
  echo "== speed CNB_DCN_BLEND=$m" >> $o
  CNB_DCN_BLEND=$m timeout 300 python tools/dcn_bench.py >> $o 2>&1
done
CNB_DCN_BLEND=wbf16 timeout 900 python bench.py --steps 20 --warmup 3 --no-train-leg --no-cpu-baseline > gpurun_out/r02v.bench.json 2> gpurun_out/r02v.bench.err; echo "bench exit $?" >> $o
python - >> $o <<'PY'
import json
d=json.loads(open("gpurun_out/r02v.bench.json").read().strip().splitlines()[-1])
print("bench wbf16", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; dcn ms", d["roofline"].get("dcn_ms_per_step"))
PY
cat $o
