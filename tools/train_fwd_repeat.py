"""Run-to-run spread of the train-mode backbone forward (batch-statistics BatchNorm; its sums use fp32 atomics, so the
last bits differ between runs): rel-L2 and max differences of the output feature map, for the tiny test shape and a
larger one: python tools/train_fwd_repeat.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200.models import create_model  # noqa: E402
from centernet_pytorch_lightning_b200.utils.synthetic import randomize_  # noqa: E402

dev = torch.device("cuda:0")
for (B, hw, gain) in ((2, 128, 0.02), (2, 128, 0.0), (8, 256, 0.02), (16, 512, 0.02)):
    torch.manual_seed(0)
    m = create_model("dla_34")
    randomize_(m.state_dict(), 3, offset_gain=gain)
    m = m.to(dev).train()
    x = torch.rand(B, 3, hw, hw, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        fs = [m(x)[0].float().clone() for _ in range(4)]
    ref = fs[0]
    rl2 = [((f - ref).norm() / ref.norm()).item() for f in fs[1:]]
    mx = [((f - ref).abs().max() / ref.abs().max()).item() for f in fs[1:]]
    print(f"B={B} {hw}x{hw} offset_gain={gain}: rel-L2 vs run 0: " + " ".join(f"{v:.2e}" for v in rl2) + "   max/max: " + " ".join(f"{v:.2e}" for v in mx))
    m.eval()
    with torch.no_grad():
        e = [m(x)[0].float().clone() for _ in range(2)]
    print(f"   eval mode: identical = {torch.equal(e[0], e[1])}")
