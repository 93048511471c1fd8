"""Repeat conv layers (CTA-pair kernel, footprint plain mode, rows kernel) until an output differs from the first run."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
g = torch.Generator().manual_seed(5)
for (B, ci, co, hw, k) in ((8, 256, 256, 32, 3), (8, 128, 27, 64, 3), (8, 128, 64, 64, 3), (8, 64, 64, 128, 3), (4, 512, 512, 16, 3), (8, 64, 768, 64, 3)):
    x = ops.to_nhwc_bf16(torch.randn(B, ci, hw, hw, generator=g).to(dev))
    w = ops.pack_conv_weights((torch.randn(co, ci, k, k, generator=g) * 0.05).to(dev))
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    mode = 2 if co % 8 else 0
    first = ops.conv2d(x, w, co, k, 1, k // 2, sc, sh, act=1, out_mode=mode).clone()
    bad = 0
    for it in range(iters):
        y = ops.conv2d(x, w, co, k, 1, k // 2, sc, sh, act=1, out_mode=mode)
        if not torch.equal(y, first):
            bad += 1
    print(f"conv {ci}->{co} @{hw}x{hw}: {bad} differing runs in {iters}")
