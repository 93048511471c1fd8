#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02s.txt; : > $o
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 >> $o
echo "== dcn" >> $o
timeout 300 python tools/dcn_bench.py >> $o 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-train-leg --no-cpu-baseline > gpurun_out/r02s.bench.json 2> gpurun_out/r02s.bench.err; echo "bench exit $?" >> $o
python - >> $o <<'PY'
import json
d=json.loads(open("gpurun_out/r02s.bench.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "; frac", round(d["roofline"]["frac"],4))
PY
cat $o
