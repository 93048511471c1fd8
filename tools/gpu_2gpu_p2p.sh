#!/bin/bash
# 2-GPU visit: this library's peer-memory all-reduce against NCCL's (check tool), then the training leg with each.
mkdir -p gpurun_out
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29603 tools/p2p_check.py > $o/r02k.p2p_check.txt 2>&1; echo "p2p check exit $?"; grep -v "^\[rank\|^W1\|^E1\|^  " $o/r02k.p2p_check.txt | tail -n 14
CNB_GRAD_COMM=p2p timeout 600 $TR --master-port 29604 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > $o/r02k.train2_p2p.json 2> $o/r02k.train2_p2p.err; echo "train p2p exit $?"; tail -n 3 $o/r02k.train2_p2p.err
timeout 600 $TR --master-port 29602 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > $o/r02k.train2_nccl.json 2> $o/r02k.train2_nccl.err; echo "train nccl exit $?"; tail -n 2 $o/r02k.train2_nccl.err
python - <<'PY'
import json
for f in ("train2_p2p","train2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02k.{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],2), d.get("allreduce"))
    except Exception as e: print(f, "ERR", e)
PY
