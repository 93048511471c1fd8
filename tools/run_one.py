"""Tiny drivers for ncu captures: python tools/run_one.py decode|conv64|dcn64|head"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops  # noqa: E402
from centernet_pytorch_lightning_b200.decode import ctdet_decode  # noqa: E402
from centernet_pytorch_lightning_b200.utils import synthetic  # noqa: E402

what = sys.argv[1]
dev = torch.device("cuda:0")
B = 32
if what == "decode":
    heat, wh, reg = synthetic.ctdet_maps(B, 80, 128, 128, seed=1)
    heat, wh, reg = (torch.from_numpy(a).to(dev) for a in (heat, wh, reg))
    for _ in range(4):
        ctdet_decode(heat, wh, reg)
elif what == "decode_sat":
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(32, 80, 16, 16, generator=g) * 30
    lg = torch.kron(lg, torch.ones(8, 8)) + torch.randn(32, 80, 128, 128, generator=g)
    heat = torch.sigmoid(lg).to(dev)
    wh = (torch.rand(32, 2, 128, 128, generator=g) * 30).to(dev)
    reg = torch.rand(32, 2, 128, 128, generator=g).to(dev)
    for _ in range(4):
        ctdet_decode(heat, wh, reg)
elif what == "head1x1":
    mid = torch.randn(B, 128, 128, 768, device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weights(torch.randn(80, 256, 1, 1, device=dev) * 0.05)
    bias = torch.zeros(80, device=dev)
    for _ in range(4):
        ops.conv2d(ops.View(mid, 256, 0), w, 80, 1, 1, 0, None, bias, act=2, out_mode=1)
elif what == "up":
    x = torch.randn(B, 64, 64, 64, device=dev).to(torch.bfloat16)
    add = torch.randn(B, 128, 128, 64, device=dev).to(torch.bfloat16)
    wt = ops.relayout_dw_weights(torch.randn(64, 1, 4, 4, device=dev), 2)
    for _ in range(4):
        ops.dw_deconv_up(x, wt, 2, add=add)
elif what == "headfused":      # CenterHead (ctdet) through csrc/head_fused.cu at the benchmark shape
    from centernet_pytorch_lightning_b200.models.heads import CenterHead
    from centernet_pytorch_lightning_b200.utils.synthetic import randomize_
    head = CenterHead({"heatmap": 80, "width_height": 2, "regression": 2}, 64, 256).eval()
    randomize_(head.state_dict(), 1)
    head = head.to(dev)
    x = torch.randn(B, 128, 128, 64, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(4):
            head(x, sigmoid=("heatmap",))
elif what in ("wgrad64", "wgrad256", "col2im64", "im2col64"):   # backward kernels at config-3 shapes (batch 16)
    from centernet_pytorch_lightning_b200 import autograd_ops as ag
    from centernet_pytorch_lightning_b200.DCN.dcn_v2 import DCN
    Bt = 16
    if what.startswith("wgrad"):
        ci, hw = (64, 128) if what == "wgrad64" else (256, 32)
        conv = torch.nn.Conv2d(ci, ci, 3, 1, 1, bias=False).to(dev)
        x = torch.randn(Bt, hw, hw, ci, device=dev).to(torch.bfloat16).requires_grad_(True)
        for _ in range(4):
            y = ag.conv(x, conv)
            y.backward(torch.randn_like(y))
    else:
        m = DCN(64, 64, (3, 3), 1, 1).to(dev).train()
        with torch.no_grad():
            m.conv_offset_mask.weight.normal_(std=0.01)
            m.conv_offset_mask.bias.uniform_(-0.3, 0.3)
        x = torch.randn(Bt, 128, 128, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
        for _ in range(4):
            y = ag.dcn(x, m)
            y.backward(torch.randn_like(y))
elif what == "s2d":
    x4 = ops.to_nhwc_bf16(torch.randn(B, 3, 512, 512, device=dev), c_pad=4)
    wpk, geom = ops.pack_stem_s2d_weights(torch.randn(16, 3, 7, 7, device=dev) * 0.05)
    sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    for _ in range(4):
        ops.stem_s2d(x4, wpk, geom, 16, sc, sh)
else:
    ci, co, hw, k = {"conv64": (64, 64, 128, 3), "dcn64": (64, 64, 128, 3), "head": (64, 768, 128, 3),
                     "conv256": (256, 256, 32, 3), "conv128": (128, 128, 64, 3), "conv16": (16, 16, 512, 3), "stem": (8, 16, 512, 7), "off128": (128, 27, 64, 3)}[what]
    x = torch.randn(B, hw, hw, ci, device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weights(torch.randn(co, ci, k, k, device=dev) * 0.05)
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    om = ((torch.rand(1, 1, 1, 32, device=dev) * 2 - 1) * 0.8 + torch.randn(B, hw, hw, 32, device=dev) * 0.05) if what.startswith("dcn") else None
    for _ in range(4):
        if om is not None:
            ops.dcnv2(x, om, w, co, sc, sh, act=1)
        else:
            ops.conv2d(x, w, co, k, 1, k // 2, sc, sh, act=1, out_mode=2 if co % 8 else 0)
torch.cuda.synchronize()
