"""Times the row-window conv kernel on the thin DLA-34 layers: python tools/rows_bench.py [case ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops  # noqa: E402

CASES = {  # name: (Ci, Co, H, k, stride)
    "stem": (8, 16, 512, 7, 1), "c16": (16, 16, 512, 3, 1), "c16s2": (16, 32, 512, 3, 2),
    "c32s2": (32, 64, 256, 3, 2), "c64": (64, 64, 128, 3, 1), "c64_27": (64, 27, 128, 3, 1),
}
dev = torch.device("cuda:0")
B = 32
for name in (sys.argv[1:] or list(CASES)):
    ci, co, hw, k, s = CASES[name]
    x = torch.randn(B, hw, hw, ci, device=dev).to(torch.bfloat16)
    w_kw = 8 if (ci == 8 and k == 7) else 0
    w = ops.pack_conv_weights(torch.randn(co, ci, k, k, device=dev) * 0.05, kw_pad=w_kw or None)
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    mode = 2 if co % 8 else 0
    run = lambda: ops.conv2d(x, w, co, k, s, k // 2, sc, sh, act=1, out_mode=mode, w_kw=w_kw)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    ho = hw // s
    units = B * ho * (ho // 128)
    print(f"{name:7s} {us:8.1f} us   {us * 1e-6 * 1.965e9 / (units / 148):7.0f} clk/unit  "
          f"(ACC={os.environ.get('CNB_ROWS_ACC', '-')} DEPTH={os.environ.get('CNB_ROWS_DEPTH', '-')})")
