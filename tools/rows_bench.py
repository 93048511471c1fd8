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
def s2d_case():
    x = torch.randn(B, 3, 512, 512, device=dev)
    wpk, geom = ops.pack_stem_s2d_weights(torch.randn(16, 3, 7, 7, device=dev) * 0.05)
    sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    x4 = ops.to_nhwc_bf16(x, c_pad=4)
    for fn, label in ((lambda: ops.stem_s2d(x4, wpk, geom, 16, sc, sh), "stem_s2d"),
                      (lambda: ops.to_nhwc_bf16(x, c_pad=4), "to_nhwc4"), (lambda: ops.to_nhwc_bf16(x, c_pad=8), "to_nhwc8")):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{label:9s} {e0.elapsed_time(e1) / 5 * 1e3:8.1f} us")


for name in (sys.argv[1:] or list(CASES) + ["s2d"]):
    if name == "s2d":
        s2d_case()
        continue
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
