#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02c
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "dla34 or reduces" -s --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $out.train.txt
timeout 900 python -m pytest tests/test_parity_e2e_gpu.py -q -m gpu -s --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $out.parity.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $out.bench.json 2> $out.bench.err; tail -n 5 $out.bench.err
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > $out.bench_train.json 2> $out.bench_train.err; tail -n 5 $out.bench_train.err
cat $out.train.txt | tail -n 25; tail -n 25 $out.parity.txt
python - <<'PY'
import json
for f in ("gpurun_out/r02c.bench.json","gpurun_out/r02c.bench_train.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
        print(" train:", json.dumps(d.get("train"))[:900])
        if "roofline_loss" in d: print(" loss:", d["roofline_loss"]["achieved"], d["roofline_loss"]["frac"]); print(" dechost", d.get("decode_host_entry"))
    except Exception as e: print(f, "ERR", e)
PY
