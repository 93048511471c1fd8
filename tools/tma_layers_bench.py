"""Times single conv_tma layers of the DLA-34 schedule (B=32): python tools/tma_layers_bench.py [case ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops  # noqa: E402

CASES = {  # name: (Ci, Co, H, k)
    "c256": (256, 256, 32, 3), "c128": (128, 128, 64, 3), "c512": (512, 512, 16, 3), "head": (64, 768, 128, 3),
    "off128": (128, 27, 64, 3), "off256": (256, 27, 32, 3), "root448": (448, 128, 64, 1), "c64": (64, 64, 64, 3),
    "c64_128": (64, 64, 128, 3), "off64": (64, 27, 128, 3), "off512": (512, 27, 16, 3), "c128_64o": (128, 64, 64, 3),
    "c16": (16, 16, 512, 3), "c32": (32, 32, 256, 3),
}
dev = torch.device("cuda:0")
B = 32
for name in (sys.argv[1:] or list(CASES)):
    ci, co, hw, k = CASES[name]
    x = torch.randn(B, hw, hw, ci, device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weights(torch.randn(co, ci, k, k, device=dev) * 0.05)
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    mode = 2 if co % 8 else 0
    run = lambda: ops.conv2d(x, w, co, k, 1, k // 2, sc, sh, act=1, out_mode=mode)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    tf = 2.0 * B * hw * hw * co * ci * k * k / us * 1e-6
    kblocks = (B * hw * hw / 128) * ((co + 255) // 256) * (max(ci, 64) * k * k / 64) / 148
    print(f"{name:8s} {us:8.1f} us  {tf:7.1f} TF/s  {us * 1e-6 * 1.965e9 / kblocks:7.0f} clk per K block per SM")
