"""Times the depthwise bilinear up-sampling (+ add) at the DLA-34 IDAUp geometries: python tools/up_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
for (c, h, f) in [(64, 64, 2), (128, 32, 2), (256, 16, 2), (64, 32, 4), (64, 16, 8)]:
    x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
    add = torch.randn(B, h * f, h * f, c, device=dev).to(torch.bfloat16)
    w = torch.randn(c, 1, 2 * f, 2 * f, device=dev)
    wt = ops.relayout_dw_weights(w, f)
    run = lambda: ops.dw_deconv_up(x, wt, f, add=add)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    mb = (x.numel() + 2 * add.numel()) * 2 / 1e6
    print(f"C={c:4d} {h}x{h} f={f}: {us:7.1f} us  {mb / us * 1e-3 * 1e3:6.2f} TB/s ({mb:.0f} MB)" .replace("TB/s", "GB/ms"))
