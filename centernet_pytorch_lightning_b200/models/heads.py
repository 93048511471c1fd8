"""`CenterHead` / `HeadConv` with the reference's constructor, parameter names and init
(CenterNet/models/heads.py:4-50: `<name>.fc.0` 3x3 conv + ReLU, `<name>.fc.2` 1x1 conv; heat-map bias
-2.19, other heads N(0, 0.001)).

Execution: the 3x3 convs of ALL heads run as one tcgen05 implicit GEMM (Co = head_conv * n_heads, bias +
ReLU in the epilogue, NHWC bf16), then one 1x1 GEMM per head reads its channel slice and writes the
NCHW fp32 map the decode kernels and the losses consume.
"""
import torch
from torch import nn

from .. import ops
from ..ops import View


class HeadConv(nn.Module):
    def __init__(self, out_channels: int, intermediate_channel: int, head_conv: int):
        super().__init__()
        self.out_channels = out_channels
        self.fc = nn.Sequential(
            nn.Conv2d(intermediate_channel, head_conv, kernel_size=3, padding=1, bias=True),
            nn.ReLU(inplace=True),
            nn.Conv2d(head_conv, out_channels, kernel_size=1, stride=1, padding=0),
        )

    def fill_fc_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        return CenterHead.run_heads({"_": self}, x)["_"]


class CenterHead(nn.Module):
    def __init__(self, heads, intermediate_channel, head_conv):
        super().__init__()
        self.heads = heads
        for name, out_channel in heads.items():
            self.__setattr__(name, HeadConv(out_channel, intermediate_channel, head_conv))
        self.init_weights()
        self._packed = None

    def init_weights(self):
        for name in self.heads.keys():
            if name.startswith("heatmap"):
                self.__getattr__(name).fc[-1].bias.data.fill_(-2.19)
            else:
                self.__getattr__(name).fill_fc_weights()

    def _invalidate(self):
        self._packed = None

    def train(self, mode=True):
        self._invalidate()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    @staticmethod
    def _pack(mods):
        with torch.no_grad():
            w3 = torch.cat([m.fc[0].weight for m in mods.values()], 0)
            b3 = torch.cat([m.fc[0].bias for m in mods.values()], 0).float().contiguous()
            one = [(ops.pack_conv_weights(m.fc[2].weight), m.fc[2].bias.detach().float().contiguous())
                   for m in mods.values()]
        return ops.pack_conv_weights(w3), b3, one

    @staticmethod
    def run_heads(mods, x, packed=None, sigmoid=()):
        """mods: name -> HeadConv.  x: NCHW fp32 tensor (optionally tagged with its NHWC bf16 twin by the
        backbone), NHWC bf16 tensor, or ops.View.  Returns name -> [B,C_out,H,W] fp32."""
        if isinstance(x, torch.Tensor) and x.dtype == torch.float32:
            v = getattr(x, "_cnb_nhwc", None)
            x = v if v is not None else ops.to_nhwc_bf16(x)
        x = ops.as_view(x)
        w3, b3, one = packed if packed is not None else CenterHead._pack(mods)
        hc = next(iter(mods.values())).fc[0].out_channels
        mid = ops.conv2d(x, w3, hc * len(mods), 3, 1, 1, None, b3, act=1)
        ret = {}
        for i, (name, m) in enumerate(mods.items()):
            w1, b1 = one[i]
            ret[name] = ops.conv2d(View(mid, hc, i * hc), w1, m.out_channels, 1, 1, 0, None, b1,
                                   act=2 if name in sigmoid else 0, out_mode=1)
        return ret

    def forward(self, x, sigmoid=()):
        if self.training:
            raise NotImplementedError("centernet_b200 CenterHead: training-mode forward is not built yet; call .eval()")
        mods = {name: self.__getattr__(name) for name in self.heads.keys()}
        if self._packed is None:
            self._packed = self._pack(mods)
        return self.run_heads(mods, x, self._packed, sigmoid)
