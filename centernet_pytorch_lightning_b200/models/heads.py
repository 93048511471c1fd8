"""`CenterHead` / `HeadConv` with the reference's constructor, parameter names and init
(CenterNet/models/heads.py:4-50: `<name>.fc.0` 3x3 conv + ReLU, `<name>.fc.2` 1x1 conv; heat-map bias
-2.19, other heads N(0, 0.001)).

Execution (eval): ONE kernel for all heads (csrc/head_fused.cu): implicit-GEMM 3x3 on tcgen05, bias + ReLU + bf16 in the
epilogue, the 256-channel intermediate written back to TENSOR MEMORY and consumed from there by the 1x1 GEMM
(tcgen05.mma with the A operand in TMEM), bias / sigmoid, NCHW fp32 out -- the intermediate never reaches HBM.
Geometries the fused kernel does not cover (input channels != 64, head_conv != 256, > 96 output channels) run the
3x3 convs of all heads as one GEMM and one 1x1 GEMM per head.  Training: models/exec_modes.py.
"""
import torch
from torch import nn

from .. import ops
from ..ops import View


class HeadConv(nn.Module):
    """3x3 conv (bias) -> ReLU -> 1x1 conv, parameters under `fc.0` / `fc.2` (heads.py:4-25)."""

    def __init__(self, out_channels: int, intermediate_channel: int, head_conv: int):
        super().__init__()
        self.out_channels = out_channels
        self.fc = nn.Sequential(nn.Conv2d(intermediate_channel, head_conv, 3, padding=1),
                                nn.ReLU(inplace=True),
                                nn.Conv2d(head_conv, out_channels, 1))

    def fill_fc_weights(self):
        """N(0, 0.001) weights, zero biases (heads.py:17-22)."""
        with torch.no_grad():
            for conv in (self.fc[0], self.fc[2]):
                conv.weight.normal_(std=0.001)
                conv.bias.zero_()

    def forward(self, x):
        return CenterHead.run_heads({"_": self}, x)["_"]


class CenterHead(nn.Module):
    def __init__(self, heads, intermediate_channel, head_conv):
        super().__init__()
        self.heads = heads
        for name, out_channel in heads.items():
            setattr(self, name, HeadConv(out_channel, intermediate_channel, head_conv))
        self.init_weights()
        self._packed = ops.PackCache()

    def init_weights(self):
        """Heat-map heads keep the default conv init with the last bias at -2.19 (focal-loss prior); every other
        head is N(0, 0.001) / zero bias (heads.py:44-50)."""
        for name in self.heads:
            head = getattr(self, name)
            if name.startswith("heatmap"):
                with torch.no_grad():
                    head.fc[2].bias.fill_(-2.19)
            else:
                head.fill_fc_weights()

    def invalidate_caches(self):
        self._packed.clear()

    def _apply(self, fn, *a, **k):
        self.invalidate_caches()
        return super()._apply(fn, *a, **k)

    @staticmethod
    def _pack(mods):
        with torch.no_grad():
            w3 = torch.cat([m.fc[0].weight for m in mods.values()], 0)
            b3 = torch.cat([m.fc[0].bias for m in mods.values()], 0).float().contiguous()
            one = [t for m in mods.values()
                   for t in (ops.pack_conv_weights(m.fc[2].weight), m.fc[2].bias.detach().float().contiguous())]
            w1cat = torch.cat([one[2 * i].reshape(-1) for i in range(len(mods))])   # [sum round16(c_out)][256], fused kernel
        return (ops.pack_conv_weights(w3), b3, w1cat, *one)     # flat tuple of tensors (ops.PackCache refreshes in place)

    @staticmethod
    def run_heads(mods, x, packed=None, sigmoid=()):
        """mods: name -> HeadConv.  x: NCHW fp32 tensor (optionally tagged with its NHWC bf16 twin by the
        backbone), NHWC bf16 tensor, or ops.View.  Returns name -> [B,C_out,H,W] fp32."""
        if isinstance(x, torch.Tensor) and x.dtype == torch.float32:
            v = getattr(x, "_cnb_nhwc", None)
            x = v if v is not None else ops.to_nhwc_bf16(x)
        x = ops.as_view(x)
        w3, b3, w1cat, *one = packed if packed is not None else CenterHead._pack(mods)
        hc = next(iter(mods.values())).fc[0].out_channels
        fused = ops.heads_fused(x, w3, b3, w1cat, [one[2 * i + 1] for i in range(len(mods))], hc,
                                [m.out_channels for m in mods.values()], [2 if n in sigmoid else 0 for n in mods])
        if fused is not None:
            return dict(zip(mods, fused))
        mid = ops.conv2d(x, w3, hc * len(mods), 3, 1, 1, None, b3, act=1)
        ret = {}
        for i, (name, m) in enumerate(mods.items()):
            w1, b1 = one[2 * i], one[2 * i + 1]
            ret[name] = ops.conv2d(View(mid, hc, i * hc), w1, m.out_channels, 1, 1, 0, None, b1,
                                   act=2 if name in sigmoid else 0, out_mode=1)
        return ret

    def forward(self, x, sigmoid=()):
        from . import exec_modes
        if exec_modes.precision() == "fp32-strict":
            out = exec_modes.run_heads(self, x, exec_modes.StrictBackend())
            return {k: (v.sigmoid_() if k in sigmoid else v) for k, v in out.items()}
        if self.training:
            out = exec_modes.run_heads(self, x, exec_modes.TrainBackend())
            return {k: (v.sigmoid() if k in sigmoid else v) for k, v in out.items()}
        mods = {name: getattr(self, name) for name in self.heads}
        params = [p for m in mods.values() for p in (m.fc[0].weight, m.fc[0].bias, m.fc[2].weight, m.fc[2].bias)]
        packed = self._packed.get("heads", params, lambda: self._pack(mods))
        return self.run_heads(mods, x, packed, sigmoid)
