"""2+ GPUs (torchrun): this library's peer-memory gradient all-reduce (csrc/p2p_allreduce.cu) against NCCL's, on the flat
gradient buffer of DLA-34 + ctdet heads, then a few training steps with each:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from centernet_pytorch_lightning_b200.trainer import FlatTrainer, ctdet_training_step  # noqa: E402
from centernet_pytorch_lightning_b200.utils.synthetic import ctdet_targets  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
model, head = bench.seeded_weights(bench.TRAIN)
model, head = model.to(dev).train(), head.to(dev).train()
for mc in ("1", "0"):
    os.environ["CNB_P2P_MULTICAST"] = mc
    tr = FlatTrainer([model, head], comm="p2p", world_size=world)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    for step in range(3):
        tr.zero_grad()
        local = torch.randn(tr.numel, generator=g).to(dev)
        want = local.clone()
        dist.all_reduce(want)
        tr.flat_g.copy_(local)
        for p in reversed(tr.params):
            p._cnb_ready()
        tr.finish_backward()
        torch.cuda.synchronize()
        err = (tr.flat_g - want).abs().max().item()
        if rank == 0:
            print(f"multicast={mc} (mapped: {bool(tr._p2p.multicast)}) step {step}: max |p2p - nccl| = {err:.3e} "
                  f"(max |sum| {want.abs().max().item():.2f}); buckets launched on readiness: {len(tr.launch_log)}")
        assert err <= 1e-5 * world, err
    del tr
x = torch.rand(4, 3, 256, 256, device=dev)
tgt = {k: v.to(dev) for k, v in ctdet_targets(4, 80, 64, 64, n_obj=10, seed=rank).items()}
for comm in ("nccl", "p2p"):
    model, head = bench.seeded_weights(bench.TRAIN)
    model, head = model.to(dev).train(), head.to(dev).train()
    tr = FlatTrainer([model, head], lr=1e-4, comm=comm, world_size=world)
    losses = [ctdet_training_step(model, head, tr, x, tgt).item() for _ in range(4)]
    w = tr.flat_p.clone()
    ref = w.clone()
    dist.broadcast(ref, 0)
    if rank == 0:
        print(f"training with comm={comm}: losses {[round(v, 4) for v in losses]}; parameters identical across ranks: "
              f"{bool((w - ref).abs().max() == 0)}")
    assert (w - ref).abs().max().item() == 0
dist.destroy_process_group()
