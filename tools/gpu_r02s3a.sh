#!/bin/bash
# plain 3x3 mode with per-group partial accumulators: correctness, reproducibility, layer times, bench
mkdir -p gpurun_out
o=gpurun_out/r02s3a.txt; : > $o
CNB_CONV_FP=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -n 8 >> $o
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
echo "== reproducibility CNB_CONV_FP=1" >> $o
CNB_CONV_FP=1 timeout 300 python tools/conv_race_hunt.py 1500 >> $o 2>&1
for v in 0 1 2; do
echo "== layers CNB_CONV_FP=$v" >> $o
CNB_CONV_FP=$v timeout 300 python tools/tma_layers_bench.py off128 c128_64o off64 off256 c64 >> $o 2>&1
done
echo "== default suite (conv, model)" >> $o
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -n 5 >> $o
for v in 0 default 2; do
  if [ $v = default ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/r02s3a_bench_$v.json 2> gpurun_out/r02s3a_bench_$v.err
  else CNB_CONV_FP=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/r02s3a_bench_$v.json 2> gpurun_out/r02s3a_bench_$v.err; fi
  echo "bench CNB_CONV_FP=$v: $(python -c "import json;d=json.loads(open('gpurun_out/r02s3a_bench_$v.json').read().strip().splitlines()[-1]);print(round(d['value'],1), round(d['ms_per_step'],3))")" >> $o
done
cat $o
