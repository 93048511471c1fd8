"""Per-kernel GPU time of one training step (configs[2], batch 16) through torch.profiler (CUPTI):
    python tools/profile_train.py [steps] > profiles/<tag>_train_kernels.txt
Prints the kernels sorted by total device time and the host wall time per step next to the device-busy time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from centernet_pytorch_lightning_b200.trainer import FlatTrainer, ctdet_training_step  # noqa: E402
from centernet_pytorch_lightning_b200.utils.synthetic import ctdet_targets  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
B, R = bench.TRAIN["batch"], bench.TRAIN["res"]
model, head = bench.seeded_weights(bench.TRAIN)
model, head = model.to(dev).train(), head.to(dev).train()
tr = FlatTrainer([model, head], lr=1e-4)
x = torch.rand(B, 3, R, R, device=dev)
tgt = {k: v.to(dev) for k, v in ctdet_targets(B, 80, R // 4, R // 4, n_obj=32, seed=1).items()}
for _ in range(3):
    ctdet_training_step(model, head, tr, x, tgt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ctdet_training_step(model, head, tr, x, tgt)
host_ms = (time.perf_counter() - t0) * 1e3 / steps           # enqueue time only (no sync inside)
torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) * 1e3 / steps
from centernet_pytorch_lightning_b200.trainer import GraphedCtdetStep  # noqa: E402
gstep = GraphedCtdetStep(model, head, tr, B, R)
for _ in range(2):
    gstep(x, tgt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    gstep(x, tgt)
e1.record()
torch.cuda.synchronize()
print(f"graph replay of the whole step: {e0.elapsed_time(e1) / 5:.2f} ms/step")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        ctdet_training_step(model, head, tr, x, tgt)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev) / steps / 1e3
print(f"batch {B}: wall {wall_ms:.2f} ms/step, host enqueue {host_ms:.2f} ms/step, device busy {tot:.2f} ms/step "
      f"({sum(e.count for e in ev) // steps} device activities/step)")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:45]:
    print(f"{e.device_time_total / steps / 1e3:9.3f} ms  x{e.count // steps:5d}  {e.key[:110]}")
