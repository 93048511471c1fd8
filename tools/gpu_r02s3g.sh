#!/bin/bash
# footprint kernel with 5 sampler groups (960 threads, 64 registers)
mkdir -p gpurun_out
o=gpurun_out/r02s3g.txt; : > $o
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -n 4 >> $o
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
timeout 300 python tools/dcn_bench.py >> $o 2>&1
timeout 300 python tools/tma_layers_bench.py off128 >> $o 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/r02s3g_bench.json 2> gpurun_out/r02s3g_bench.err
echo "bench: $(python -c "import json;d=json.loads(open('gpurun_out/r02s3g_bench.json').read().strip().splitlines()[-1]);print(round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['dcn_ms_per_step'])")" >> $o
cat $o
