"""Repeat one DCN layer until an output differs from the first run; describe where (tile, row in tile, channel chunk)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops
dev = torch.device("cuda:0")
B, Ci, Co, H = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (8, 64, 64, 128)))
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3000
g = torch.Generator().manual_seed(11)
x = ops.to_nhwc_bf16(torch.randn(B, Ci, H, H, generator=g).to(dev))
om = (torch.randn(B, H, H, 32, generator=g) * 0.7).to(dev)
wpk = ops.pack_conv_weights((torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).to(dev))
bias = torch.zeros(Co, device=dev)
first = ops.dcnv2(x, om, wpk, Co, None, bias, act=0).clone()
nbad = 0
for it in range(iters):
    y = ops.dcnv2(x, om, wpk, Co, None, bias, act=0)
    if not torch.equal(y, first):
        nbad += 1
        d = (y != first)
        idx = d.nonzero()
        n, yy, xx, cc = idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]
        tiles = torch.unique(torch.stack([n, yy // 8, xx // 16], 1), dim=0)
        rows = torch.unique((yy % 8) * 16 + xx % 16)
        print(f"iter {it}: {idx.shape[0]} elements differ in {tiles.shape[0]} tile(s) {tiles[:4].tolist()}; rows in tile {rows[:20].tolist()} "
              f"({rows.numel()} rows); channels {torch.unique(cc)[:16].tolist()} ({torch.unique(cc).numel()} ch); "
              f"max |diff| {(y.float() - first.float()).abs().max().item():.4f}; tile linear idx {[(int(t[0]) * (H // 8) + int(t[1])) * (H // 16) + int(t[2]) for t in tiles[:4]]}")
        if nbad >= 6:
            break
print(f"{nbad} differing runs in {it + 1}")
