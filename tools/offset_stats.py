"""Offset statistics of the DCN layers in one DLA-34 step with the bench's seeded weights (how far the samples reach)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from centernet_pytorch_lightning_b200 import ops
dev = torch.device("cuda:0")
model, head = bench.seeded_weights(bench.CONFIGS[2])
model, head = model.to(dev), head.to(dev)
x = torch.rand(4, 3, 512, 512, device=dev)
_dcn = ops.dcnv2
def dcnv2(xv, om, wpk, Co, *a, **kw):
    v = ops.as_view(xv)
    off = om[..., :18].float()
    print(f"dcn {v.C:4d}->{Co:4d} @{v.H}x{v.W}: |offset| mean {off.abs().mean():.2f} std {off.std():.2f} max {off.abs().max():.1f}; "
          f"> 2: {100 * (off.abs() > 2).float().mean():.2f} %  > 3: {100 * (off.abs() > 3).float().mean():.2f} %  > 4: {100 * (off.abs() > 4).float().mean():.2f} %; "
          f"pixel-to-pixel |d offset| mean {(off[:, :, 1:] - off[:, :, :-1]).abs().mean():.2f}")
    return _dcn(xv, om, wpk, Co, *a, **kw)
ops.dcnv2 = dcnv2
with torch.no_grad():
    head(model(x)[-1], sigmoid=("heatmap",))
