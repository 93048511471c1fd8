"""Decode-only micro-benchmark: ctdet_decode on resident [32,80,128,128] maps, L2 flushed
between launches, CUDA-event timed: python tools/decode_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200.decode import ctdet_decode  # noqa: E402
from centernet_pytorch_lightning_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for kind in ("uniform", "bumps", "saturated"):
    if kind == "saturated":
        g = torch.Generator().manual_seed(3)
        lg = torch.randn(32, 80, 16, 16, generator=g) * 30
        lg = torch.kron(lg, torch.ones(8, 8)) + torch.randn(32, 80, 128, 128, generator=g)
        heat = torch.sigmoid(lg).to(dev)
        wh = (torch.rand(32, 2, 128, 128, generator=g) * 30).to(dev)
        reg = torch.rand(32, 2, 128, 128, generator=g).to(dev)
    else:
        heat, wh, reg = synthetic.ctdet_maps(32, 80, 128, 128, seed=1, kind=kind)
        heat, wh, reg = (torch.from_numpy(a).to(dev) for a in (heat, wh, reg))
    for _ in range(3):
        ctdet_decode(heat, wh, reg)
    # three ways of making the reads cold: (w) write a 256 MB buffer -- leaves L2 full of DIRTY lines whose write-back
    # then competes with the scan's reads; (r) read a 256 MB buffer -- clean lines; (x) rotate three distinct copies of
    # the maps (3 x 168 MB, each larger than L2), no flush at all
    copies = [(heat, wh, reg)] + [(heat.clone(), wh.clone(), reg.clone()) for _ in range(2)]
    for mode in ("w", "r", "x"):
        ts = []
        for it in range(12):
            if mode == "w":
                flush.fill_(1)
            elif mode == "r":
                flush.sum()
            h, w_, r_ = copies[it % 3] if mode == "x" else copies[0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctdet_decode(h, w_, r_)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = ts[2:]
        ms = sorted(ts)[len(ts) // 2]
        print(f"{kind:10s} cold-by-{mode} decode ms median {ms:.4f} min {min(ts):.4f}  {167.9e6 / ms / 1e6:.0f} GB/s (events around memset+kernels)")
