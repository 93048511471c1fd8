#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02q2.txt; : > $o
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 >> $o
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02q.bench.json 2> gpurun_out/r02q.bench.err; echo "bench exit $?" >> $o
python - >> $o <<'PY'
import json
d=json.loads(open("gpurun_out/r02q.bench.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "; frac", round(d["roofline"]["frac"],4), "train", d.get("train",{}).get("value"), d.get("train",{}).get("ms_per_step"))
PY
cat $o
