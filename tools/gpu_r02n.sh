#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02n.txt; : > $o
for m in smooth random; do for impl in ws fp; do
  echo "== offsets=$m impl=$impl" >> $o
  DCN_BENCH_OFFSETS=$m CNB_DCN_IMPL=$impl timeout 300 python tools/dcn_bench.py >> $o 2>&1
done; done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_e2e_gpu.py tests/test_engine_gpu.py -x -q 2>&1 | tail -n 5 >> $o
timeout 900 python bench.py --steps 20 --warmup 3 --no-train-leg > gpurun_out/r02n.bench.json 2> gpurun_out/r02n.bench.err; echo "bench exit $?" >> $o
CNB_DCN_IMPL=ws timeout 900 python bench.py --steps 20 --warmup 3 --no-train-leg --no-cpu-baseline > gpurun_out/r02n.bench_ws.json 2> gpurun_out/r02n.bench_ws.err; echo "bench ws exit $?" >> $o
python - >> $o <<'PY'
import json
for f in ("bench","bench_ws"):
    try:
        d=json.loads(open(f"gpurun_out/r02n.{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "; dcn ms", d["roofline"].get("dcn_ms_per_step"), "frac", round(d["roofline"]["frac"],4))
    except Exception as e: print(f, "ERR", e)
PY
cat $o
