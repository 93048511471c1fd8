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
    rel = (e.double().pow(2).sum() / ref.double().pow(2).sum()).sqrt().item()
    print(f"{kind:6s} B{B} Ci{Ci} Co{Co} {H}x{W}: max err {e.max():.4f} (ref max {ref.abs().max():.3f}), rel-L2 {rel:.3e}, bad pixels {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        for n in range(B):
            rows = ["".join("#" if bad[n, yy, xx] else "." for xx in range(W)) for yy in range(H)]
            print("\n".join(rows)); print()
        ce = e.amax(dim=(0, 2, 3)); print("per-channel max err (first 16):", [round(float(v), 3) for v in ce[:16]])
for kind in ("zero", "const", "small", "big", "far"):
    run(1, 64, 64, 16, 16, kind)
run(2, 64, 64, 24, 24, "small")
run(1, 128, 64, 16, 32, "small")
run(2, 256, 128, 32, 32, "small")
run(2, 64, 64, 64, 64, "small")

# determinism stress: the kernel has no atomics, so repeated runs on the same input must be bit-identical; a table or
# ring race would show up here as a differing element sooner or later
g = torch.Generator().manual_seed(7)
for (B, Ci, Co, H, W) in ((8, 64, 64, 128, 128), (8, 128, 64, 64, 64), (8, 256, 256, 32, 32)):
    x = ops.to_nhwc_bf16(torch.randn(B, Ci, H, W, generator=g).to(dev))
    om = (torch.randn(B, H, W, 32, generator=g) * 0.7).to(dev)
    wpk = ops.pack_conv_weights((torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).to(dev))
    bias = torch.zeros(Co, device=dev)
    first = ops.dcnv2(x, om, wpk, Co, None, bias, act=0).clone()
    bad = 0
    for it in range(300):
        y = ops.dcnv2(x, om, wpk, Co, None, bias, act=0)
        if not torch.equal(y, first):
            bad += 1
    torch.cuda.synchronize()
    print(f"determinism B{B} Ci{Ci} Co{Co} {H}x{W}: {bad} of 300 repeats differ from the first run")
