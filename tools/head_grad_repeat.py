"""Is the training path of the heads reproducible on identical inputs?  One backbone forward (train mode), then the
multi-pose heads + loss + backward repeated on the SAME detached feature map; prints, per head, the largest relative
difference of the first conv's weight gradient from the first repetition (fp32 atomics give ~1e-6; a lost hand-off in a
backward kernel would give percents): python tools/head_grad_repeat.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200.utils.synthetic import randomize_  # noqa: E402
from oracle import task_torch  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
ns = task_torch.namespace("b200")
torch.manual_seed(0)
task = task_torch.MultiPoseTask(ns, "dla_34")
randomize_(task.backbone.state_dict(), 3, offset_gain=0.02)
task = task.to(dev).train()
x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
_, tgt = task_torch.task_inputs("pose", B=2, H=32, W=32)
tgt = {k: v.to(dev) for k, v in tgt.items()}
feats = []
for _ in range(3):
    feats.append(task.backbone(x)[0].detach().clone())
print("backbone (train mode) run-to-run: max |f1 - f0| / max |f0| =",
      f"{((feats[1] - feats[0]).abs().max() / feats[0].abs().max()).item():.3e}",
      f"{((feats[2] - feats[0]).abs().max() / feats[0].abs().max()).item():.3e}")
feat = feats[0]
first, worst, losses = {}, {}, []
for it in range(reps):
    task.zero_grad(set_to_none=True)
    outs = [task.heads[0](feat)]
    loss, _ = task.loss(outs, tgt)
    loss.backward()
    torch.cuda.synchronize()
    losses.append(loss.item())
    for name, p in task.heads[0].named_parameters():
        g = p.grad.detach().float().clone()
        if it == 0:
            first[name] = g
            worst[name] = 0.0
        else:
            worst[name] = max(worst[name], ((g - first[name]).norm() / (first[name].norm() + 1e-30)).item())
print("loss min/max over repetitions:", min(losses), max(losses))
for name, w in worst.items():
    if name.endswith("fc.0.weight") or w > 1e-4:
        print(f"  {name}: worst rel difference from repetition 0: {w:.3e}")
