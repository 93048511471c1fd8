#!/bin/bash
# 2-GPU visit: training leg with the NCCL bucket all-reduce (graph and eager), this library's peer-memory all-reduce, the
# default line at N=2.
mkdir -p gpurun_out
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29601 bench.py --gpus 2 --mode train --steps 10 --warmup 3 --no-graph > $o/r02x.train2_eager.json 2> $o/r02x.train2_eager.err; echo "train eager exit $?"; tail -n 2 $o/r02x.train2_eager.err
timeout 600 $TR --master-port 29602 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > $o/r02x.train2_graph.json 2> $o/r02x.train2_graph.err; echo "train graph exit $?"; tail -n 2 $o/r02x.train2_graph.err
timeout 600 $TR --master-port 29603 tools/p2p_check.py > $o/r02x.p2p_check.txt 2>&1; echo "p2p check exit $?"; tail -n 12 $o/r02x.p2p_check.txt
CNB_GRAD_COMM=p2p timeout 600 $TR --master-port 29604 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > $o/r02x.train2_p2p.json 2> $o/r02x.train2_p2p.err; echo "train p2p exit $?"; tail -n 2 $o/r02x.train2_p2p.err
timeout 900 $TR --master-port 29605 bench.py --gpus 2 --steps 20 --warmup 3 > $o/r02x.bench2.json 2> $o/r02x.bench2.err; echo "bench N=2 exit $?"; tail -n 2 $o/r02x.bench2.err
python - <<'PY'
import json
for f in ("train2_eager","train2_graph","train2_p2p","bench2"):
    try:
        d=json.loads(open(f"gpurun_out/r02x.{f}.json").read().strip().splitlines()[-1])
        t=d.get("train") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],2), "| train:", t.get("value"), t.get("ms_per_step"), (t.get("allreduce") or {}).get("exposed_ms_per_step"), t.get("launch"), t.get("error"))
    except Exception as e: print(f, "ERR", e)
PY
