#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29605 bench.py --gpus 2 --steps 20 --warmup 3 > $o/final.bench2.json 2> $o/final.bench2.err; echo "bench N=2 exit $?"; tail -n 2 $o/final.bench2.err
timeout 600 $TR --master-port 29606 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $o/final.ref2.json 2> $o/final.ref2.err; echo "ref N=2 exit $?"; cut -c1-300 $o/final.ref2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final.bench2.json").read().strip().splitlines()[-1])
t=d.get("train") or {}
print("bench2", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "| train:", t.get("value"), t.get("ms_per_step"), (t.get("allreduce") or {}).get("exposed_ms_per_step"))
PY
