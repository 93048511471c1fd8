"""Hourglass-104 behind the reference's plugin contract (CenterNet/models/backbones/large_hourglass.py:322-343):
`create_model("hourglass")` -> module with `out_channels = 256` whose `forward(x[B,3,H,W])` returns TWO feature maps
`[B,256,H/4,W/4]` (one per stack; centernet.py:16 `num_stacks = 2`), state-dict keys equal to the reference's
(`pre.0.conv.weight`, `kps.0.low2.low2.low1.0.skip.0.weight`, `cnvs_.0.1.running_mean`, ...).

The modules hold parameters only.  Execution is the package's kernel schedule over NHWC bf16: every `convolution` /
`residual` (large_hourglass.py:8-28, :51-93) is tcgen05 implicit-GEMM convolutions with the folded BatchNorm, skip add
and ReLU in the epilogue (1x1 stride-2 skip convs and 3x3 stride-2 convs through the same TMA im2col path), and
`nn.Upsample(scale_factor=2)` + `MergeUp` (:116-125, kp_module.forward :206-213) is ONE launch of the depthwise
up-sampling kernel with the nearest-neighbour taps and the fused add.
"""
import torch
from torch import nn

from ... import ops
from ...ops import View
from .pose_dla_dcn import _Compiled, _conv_bn_act


class convolution(nn.Module):          # large_hourglass.py:8-28
    def __init__(self, k, inp_dim, out_dim, stride=1, with_bn=True):
        super().__init__()
        pad = (k - 1) // 2
        self.conv = nn.Conv2d(inp_dim, out_dim, (k, k), padding=(pad, pad), stride=(stride, stride), bias=not with_bn)
        self.bn = nn.BatchNorm2d(out_dim) if with_bn else nn.Sequential()
        self.relu = nn.ReLU(inplace=True)


class residual(nn.Module):             # large_hourglass.py:51-93
    def __init__(self, k, inp_dim, out_dim, stride=1, with_bn=True):
        super().__init__()
        self.conv1 = nn.Conv2d(inp_dim, out_dim, (3, 3), padding=(1, 1), stride=(stride, stride), bias=False)
        self.bn1 = nn.BatchNorm2d(out_dim)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_dim, out_dim, (3, 3), padding=(1, 1), bias=False)
        self.bn2 = nn.BatchNorm2d(out_dim)
        self.skip = nn.Sequential(nn.Conv2d(inp_dim, out_dim, (1, 1), stride=(stride, stride), bias=False),
                                  nn.BatchNorm2d(out_dim)) if stride != 1 or inp_dim != out_dim else nn.Sequential()
        self.relu = nn.ReLU(inplace=True)


def _layer(inp_dim, out_dim, modules):                 # make_layer :96-100
    return nn.Sequential(residual(3, inp_dim, out_dim), *[residual(3, out_dim, out_dim) for _ in range(1, modules)])


def _layer_revr(inp_dim, out_dim, modules):            # make_layer_revr :103-108
    return nn.Sequential(*[residual(3, inp_dim, inp_dim) for _ in range(modules - 1)], residual(3, inp_dim, out_dim))


def _hg_layer(dim0, dim1, mod):                        # make_hg_layer :315-319 (stride-2 first block: no pooling)
    return nn.Sequential(residual(3, dim0, dim1, stride=2), *[residual(3, dim1, dim1) for _ in range(mod - 1)])


class kp_module(nn.Module):            # large_hourglass.py:144-213
    def __init__(self, n, dims, modules):
        super().__init__()
        self.n = n
        curr_mod, next_mod, curr_dim, next_dim = modules[0], modules[1], dims[0], dims[1]
        self.up1 = _layer(curr_dim, curr_dim, curr_mod)
        self.max1 = nn.Sequential()
        self.low1 = _hg_layer(curr_dim, next_dim, curr_mod)
        self.low2 = kp_module(n - 1, dims[1:], modules[1:]) if n > 1 else _layer(next_dim, next_dim, next_mod)
        self.low3 = _layer_revr(next_dim, curr_dim, curr_mod)
        self.up2 = nn.Upsample(scale_factor=2)


class HourglassNet(nn.Module):         # exkp :216-313 with the HourglassNet arguments of :322-340
    def __init__(self, num_stacks=2):
        super().__init__()
        n, dims, modules, cnv_dim = 5, [256, 256, 384, 384, 384, 512], [2, 2, 2, 2, 2, 4], 256
        self.nstack, self.out_channels = num_stacks, 256
        curr_dim = dims[0]
        self.pre = nn.Sequential(convolution(7, 3, 128, stride=2), residual(3, 128, 256, stride=2))
        self.kps = nn.ModuleList([kp_module(n, dims, modules) for _ in range(num_stacks)])
        self.cnvs = nn.ModuleList([convolution(3, curr_dim, cnv_dim) for _ in range(num_stacks)])
        self.inters = nn.ModuleList([residual(3, curr_dim, curr_dim) for _ in range(num_stacks - 1)])
        self.inters_ = nn.ModuleList([nn.Sequential(nn.Conv2d(curr_dim, curr_dim, (1, 1), bias=False), nn.BatchNorm2d(curr_dim))
                                      for _ in range(num_stacks - 1)])
        self.cnvs_ = nn.ModuleList([nn.Sequential(nn.Conv2d(cnv_dim, curr_dim, (1, 1), bias=False), nn.BatchNorm2d(curr_dim))
                                    for _ in range(num_stacks - 1)])
        self.relu = nn.ReLU(inplace=True)
        self._cc = None
        self._nearest = {}

    def invalidate_caches(self):
        self._cc = None
        self._nearest = {}

    def _apply(self, fn, *a, **k):
        self.invalidate_caches()
        return super()._apply(fn, *a, **k)

    # ---- engine ------------------------------------------------------------------------------------------------
    def _nearest_taps(self, C, device):
        """nn.Upsample(scale_factor=2) (nearest) as the 4x4 / stride 2 / pad 1 depthwise transposed filter whose only
        non-zero taps are the central 2x2 ones: out[2i + a] = in[i]."""
        key = (C, str(device))
        if key not in self._nearest:
            w = torch.zeros(C, 1, 4, 4, device=device)
            w[:, :, 1:3, 1:3] = 1.0
            self._nearest[key] = ops.relayout_dw_weights(w, 2)
        return self._nearest[key]

    def _conv(self, cc, c, x):
        return _conv_bn_act(cc, x, c.conv, c.bn if isinstance(c.bn, nn.BatchNorm2d) else None, act=1)

    def _res(self, cc, r, x):
        h = _conv_bn_act(cc, x, r.conv1, r.bn1, act=1)
        skip = _conv_bn_act(cc, x, r.skip[0], r.skip[1], act=0) if len(r.skip) else x
        return _conv_bn_act(cc, h, r.conv2, r.bn2, act=1, res=skip)

    def _seq(self, cc, seq, x):
        for r in seq:
            x = self._res(cc, r, x)
        return x

    def _kp(self, cc, m, x):
        up1 = self._seq(cc, m.up1, x)
        low1 = self._seq(cc, m.low1, x)
        low2 = self._kp(cc, m.low2, low1) if isinstance(m.low2, kp_module) else self._seq(cc, m.low2, low1)
        low3 = self._seq(cc, m.low3, low2)
        assert up1.coffset == 0 and up1.cstride == up1.C
        y = ops.dw_deconv_up(low3, self._nearest_taps(low3.C, low3.buf.device), 2, add=up1.buf)
        return View(y, low3.C, 0)

    def forward_nhwc_all(self, x):
        if self._cc is None:
            self._cc = _Compiled()
        cc = self._cc
        h = View(ops.to_nhwc_bf16(x, c_pad=8), 8, 0)
        inter = self._res(cc, self.pre[1], self._conv(cc, self.pre[0], h))
        outs = []
        for ind in range(self.nstack):
            cnv = self._conv(cc, self.cnvs[ind], self._kp(cc, self.kps[ind], inter))
            outs.append(cnv)
            if ind < self.nstack - 1:
                a = _conv_bn_act(cc, inter, self.inters_[ind][0], self.inters_[ind][1], act=0)
                inter = _conv_bn_act(cc, cnv, self.cnvs_[ind][0], self.cnvs_[ind][1], act=1, res=a)   # relu(a + b)
                inter = self._res(cc, self.inters[ind], inter)
        return outs

    def forward_nhwc(self, x):
        return self.forward_nhwc_all(x)[-1]

    def forward(self, x):
        from ..._lib import require_cuda
        require_cuda(x)
        if self.training:
            raise NotImplementedError("centernet_b200 HourglassNet: inference schedule only (training is built for DLA-34, "
                                      "the north-star configuration); call .eval()")
        outs = []
        for v in self.forward_nhwc_all(x):
            o = ops.to_nchw_f32(v)
            o._cnb_nhwc = v
            outs.append(o)
        return outs


def get_large_hourglass_net(num_layers):
    return HourglassNet()
