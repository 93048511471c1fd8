"""Debug aid: DCNv2 kernel vs torchvision on controlled offsets; prints where the error sits."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops
from torchvision.ops import deform_conv2d
dev = torch.device("cuda:0")
def bf(t): return t.to(torch.bfloat16).float()
def run(B, Ci, Co, H, W, kind):
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(B, Ci, H, W, generator=g))
    w = bf(torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5)
    bias = torch.zeros(Co)
    om = torch.zeros(B, 27, H, W)
    if kind == "const": om[:, :18] = 0.3
    if kind == "small": om = torch.randn(B, 27, H, W, generator=g) * 0.5
    if kind == "big": om = torch.randn(B, 27, H, W, generator=g) * 1.5
    if kind == "far": om[:, :18] = 7.25
    o1, o2, m = torch.chunk(om, 3, dim=1)
    ref = deform_conv2d(x, torch.cat((o1, o2), 1), w, bias, padding=1, mask=torch.sigmoid(m))
    omn = torch.zeros(B, H, W, 32); omn[..., :27] = om.permute(0, 2, 3, 1)
    y = ops.dcnv2(ops.to_nhwc_bf16(x.to(dev)), omn.to(dev).contiguous(), ops.pack_conv_weights(w.to(dev)), Co, None, bias.to(dev), act=0)
    torch.cuda.synchronize()
    e = (y.permute(0, 3, 1, 2).float().cpu() - ref).abs()
    pe = e.amax(dim=1)   # [B,H,W]
    bad = (pe > 2e-2 * ref.abs().max())
    print(f"{kind:6s} B{B} Ci{Ci} Co{Co} {H}x{W}: max err {e.max():.4f} (ref max {ref.abs().max():.3f}), bad pixels {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        for n in range(B):
            rows = ["".join("#" if bad[n, yy, xx] else "." for xx in range(W)) for yy in range(H)]
            print("\n".join(rows)); print()
        ce = e.amax(dim=(0, 2, 3)); print("per-channel max err (first 16):", [round(float(v), 3) for v in ce[:16]])
for kind in ("zero", "const", "small", "big", "far"):
    run(1, 64, 64, 16, 16, kind)
run(2, 64, 64, 24, 24, "small")
run(1, 128, 64, 16, 32, "small")
