"""Times single DCNv2 layers of the DLA-34 schedule (B=32): python tools/dcn_bench.py [case ...]
DCN_BENCH_OFFSETS=smooth (default: per-tap displacement + small per-pixel noise, what a network produces) | random
(independent N(0, 0.5) per pixel and tap: neighbouring pixels sample unrelated positions)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops  # noqa: E402

CASES = {"d64": (64, 64, 128), "d128_64": (128, 64, 64), "d128": (128, 128, 64), "d256": (256, 256, 32),
         "d256_128": (256, 128, 32), "d256_64": (256, 64, 32), "d512": (512, 256, 16)}
dev = torch.device("cuda:0")
B = 32
for name in (sys.argv[1:] or list(CASES)):
    ci, co, hw = CASES[name]
    x = torch.randn(B, hw, hw, ci, device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weights(torch.randn(co, ci, 3, 3, device=dev) * 0.05)
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    mode = os.environ.get("DCN_BENCH_OFFSETS", "smooth")
    if mode.startswith("far"):   # far:S -- a smooth field with per-tap displacements of std S pixels (what the deep layers of
        sdev = float(mode.split(":")[1]) if ":" in mode else 1.2   # the randomly initialised network produce)
        low = torch.randn(B, max(hw // 16, 1), max(hw // 16, 1), 32, device=dev) * sdev
        om = torch.nn.functional.interpolate(low.permute(0, 3, 1, 2), size=(hw, hw), mode="bilinear", align_corners=False)
        om = om.permute(0, 2, 3, 1).contiguous()
    elif mode == "random":
        om = torch.randn(B, hw, hw, 32, device=dev) * 0.5
    else:
        om = (torch.rand(1, 1, 1, 32, device=dev) * 2 - 1) * 0.8 + torch.randn(B, hw, hw, 32, device=dev) * 0.05
    # DCN_BENCH_COLD=1: rotate over enough copies of the inputs that none is left in the 126 MB L2 when it is reused
    ncopy = 1
    if os.environ.get("DCN_BENCH_COLD"):
        ncopy = max(2, int(300e6 / (x.numel() * 2 + om.numel() * 4)) + 1)
    xs = [x] + [x.clone() for _ in range(ncopy - 1)]
    oms = [om] + [om.clone() for _ in range(ncopy - 1)]
    run = lambda i: ops.dcnv2(xs[i % ncopy], oms[i % ncopy], w, co, sc, sh, act=1)
    for i in range(3):
        run(i)
    reps = max(10, ncopy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    kblocks = (B * hw * hw / 128) * (ci * 9 / 64) / 148
    print(f"{name:8s} {us:8.1f} us  {us * 1e-6 * 1.965e9 / kblocks:7.0f} clk per K block per SM")
