"""DLA-34 + DLAUp/IDAUp (DCNv2) backbone behind the reference's plugin contract
(CenterNet/models/backbones/pose_dla_dcn.py:532-581): an nn.Module with `.out_channels` whose
`forward(x[B,3,H,W] fp32)` returns `[Tensor[B,64,H/4,W/4]]` and whose state-dict keys/shapes equal the
reference's (386 entries, e.g. `base.level2.tree1.conv1.weight`,
`dla_up.ida_0.proj_1.conv.conv_offset_mask.weight`, `ida_up.up_2.weight`), so checkpoints load.

The modules below only *hold parameters*.  Execution is a flat schedule of sm_100a kernels over NHWC
bf16 buffers (tcgen05 implicit-GEMM convs with folded-BN/ReLU/residual epilogues, the DCNv2 sampler
kernel, depthwise bilinear up-sampling fused with the skip add, 2x2 max-pool), with every `torch.cat`
of the Root nodes (pose_dla_dcn.py:182) replaced by writing producers straight into slices of a
pre-allocated concat buffer.
"""
import math

import numpy as np
import os

import torch
from torch import nn

from ... import ops
from ..._lib import require_cuda as _lib_require_cuda
from ...DCN.dcn_v2 import DCN
from ...ops import View

BN_MOMENTUM = 0.1


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


def _conv(ci, co, k, stride=1):
    return nn.Conv2d(ci, co, kernel_size=k, stride=stride, padding=k // 2, bias=False)


# ---- parameter containers (names mirror the reference modules) ------------------------------------------
class BasicBlock(nn.Module):          # pose_dla_dcn.py:28-68
    def __init__(self, ci, co, stride=1):
        super().__init__()
        self.conv1, self.bn1 = _conv(ci, co, 3, stride), _bn(co)
        self.conv2, self.bn2 = _conv(co, co, 3), _bn(co)
        self.stride = stride


class Root(nn.Module):                # pose_dla_dcn.py:165-188
    def __init__(self, ci, co, residual):
        super().__init__()
        self.conv, self.bn = _conv(ci, co, 1), _bn(co)
        self.residual = residual


class Tree(nn.Module):                # pose_dla_dcn.py:191-265
    def __init__(self, levels, ci, co, stride=1, level_root=False, root_dim=0, root_residual=False):
        super().__init__()
        root_dim = root_dim or 2 * co
        if level_root:
            root_dim += ci
        if levels == 1:
            self.tree1 = BasicBlock(ci, co, stride)
            self.tree2 = BasicBlock(co, co, 1)
            self.root = Root(root_dim, co, root_residual)
        else:
            self.tree1 = Tree(levels - 1, ci, co, stride, root_residual=root_residual)
            self.tree2 = Tree(levels - 1, co, co, root_dim=root_dim + co, root_residual=root_residual)
        self.levels, self.level_root, self.root_dim = levels, level_root, root_dim
        self.ci, self.co, self.stride = ci, co, stride
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = nn.Sequential(_conv(ci, co, 1), _bn(co)) if ci != co else None


class DLA(nn.Module):                 # pose_dla_dcn.py:268-396
    def __init__(self, levels, channels):
        super().__init__()
        c = self.channels = channels
        self.base_layer = nn.Sequential(nn.Conv2d(3, c[0], 7, 1, 3, bias=False), _bn(c[0]), nn.ReLU(inplace=True))
        self.level0 = self._conv_level(c[0], c[0], levels[0])
        self.level1 = self._conv_level(c[0], c[1], levels[1], stride=2)
        self.level2 = Tree(levels[2], c[1], c[2], 2, level_root=False)
        self.level3 = Tree(levels[3], c[2], c[3], 2, level_root=True)
        self.level4 = Tree(levels[4], c[3], c[4], 2, level_root=True)
        self.level5 = Tree(levels[5], c[4], c[5], 2, level_root=True)

    @staticmethod
    def _conv_level(ci, co, n, stride=1):
        mods = []
        for i in range(n):
            mods += [_conv(ci, co, 3, stride if i == 0 else 1), _bn(co), nn.ReLU(inplace=True)]
            ci = co
        return nn.Sequential(*mods)


def _fill_bilinear(up):
    """Separable bilinear kernel into weight[:, 0] (pose_dla_dcn.py:424-432, resnet_dcn.py:109-118).  For the
    depthwise DLA up-samplers that is the whole weight; a dense ConvTranspose2d keeps its default init elsewhere."""
    k = up.weight.shape[2]
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    g = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float32)
    with torch.no_grad():
        up.weight[:, 0].copy_((g[:, None] * g[None, :]).expand_as(up.weight[:, 0]))


class DeformConv(nn.Module):          # pose_dla_dcn.py:435-454
    def __init__(self, ci, co):
        super().__init__()
        self.actf = nn.Sequential(_bn(co), nn.ReLU(inplace=True))
        self.conv = DCN(ci, co, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)


class IDAUp(nn.Module):               # pose_dla_dcn.py:457-488
    def __init__(self, o, channels, up_f):
        super().__init__()
        self.o = o
        self.factors = [int(f) for f in up_f]
        for i in range(1, len(channels)):
            f = self.factors[i]
            up = nn.ConvTranspose2d(o, o, f * 2, stride=f, padding=f // 2, output_padding=0, groups=o, bias=False)
            _fill_bilinear(up)
            setattr(self, f"proj_{i}", DeformConv(int(channels[i]), o))
            setattr(self, f"up_{i}", up)
            setattr(self, f"node_{i}", DeformConv(o, o))


class DLAUp(nn.Module):               # pose_dla_dcn.py:491-516
    def __init__(self, startp, channels, scales):
        super().__init__()
        self.startp = startp
        channels = [int(c) for c in channels]
        in_channels = list(channels)
        scales = np.array(scales, dtype=int)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, f"ida_{i}", IDAUp(channels[j], in_channels[j:], scales[j:] // scales[j]))
            scales[j + 1:] = scales[j]
            in_channels[j + 1:] = [channels[j]] * len(in_channels[j + 1:])


# ---- engine ------------------------------------------------------------------------------------------------
def fold_bn(bn, conv_bias=None):
    """eval-mode BatchNorm as (scale, shift) on the fp32 accumulator: y = acc*scale + shift."""
    with torch.no_grad():
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
        shift = (bn.bias - bn.running_mean * scale).float()
        if conv_bias is not None:
            shift = shift + conv_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


def _bn_params(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var)


class _Compiled(ops.PackCache):
    """Packed weights + folded BN per conv.  Every entry is stamped with (version, data_ptr) of the parameters and
    BN buffers it was built from (ops.param_key) and rebuilt when they change -- an optimizer step, a
    `load_state_dict` through ANY parent module, `nn.init.*_` -- like `DCN.packed()`."""

    def conv(self, conv, bn=None):
        def build():
            # the 3-channel 7x7 stem runs on an 8-channel padded input: packed 7x8 (zero column) so that the taps
            # (kw, kw+1) of one row form one K=16 tensor-core step (cnb_conv_desc.w_kw)
            w_kw = conv.kernel_size[1] + 1 if (conv.in_channels <= 8 and conv.kernel_size[1] % 2 == 1
                                               and conv.kernel_size[1] > 1) else 0
            wpk = ops.pack_conv_weights(conv.weight, kw_pad=w_kw or None)
            if bn is not None:
                scale, shift = fold_bn(bn, conv.bias)
            else:
                scale, shift = None, (conv.bias.detach().float().contiguous() if conv.bias is not None else None)
            return (wpk, scale, shift, w_kw)
        return self.get(id(conv), (conv.weight, conv.bias) + (_bn_params(bn) if bn is not None else ()), build)

    def stem_s2d(self, conv, bn):
        """space-to-depth packing of the 7x7 stem (ops.pack_stem_s2d_weights): two output pixels per GEMM row"""
        def build():
            wpk, geom = ops.pack_stem_s2d_weights(conv.weight)
            scale, shift = fold_bn(bn, conv.bias)
            return (wpk, geom, scale.repeat(2).contiguous(), shift.repeat(2).contiguous())
        return self.get(("s2d", id(conv)), (conv.weight, conv.bias) + _bn_params(bn), build)

    def deform(self, dc):
        return self.get(id(dc), (dc.conv.bias,) + _bn_params(dc.actf[0]), lambda: fold_bn(dc.actf[0], dc.conv.bias))

    def up(self, up, f):
        return self.get(id(up), (up.weight,), lambda: (ops.relayout_dw_weights(up.weight, f),))[0]


def _new(x, H, W, C):
    return torch.empty((x.B, H, W, C), dtype=torch.bfloat16, device=x.buf.device)


def _conv_bn_act(cc, x, conv, bn, act=1, res=None, out=None):
    wpk, scale, shift, w_kw = cc.conv(conv, bn)
    k, s = conv.kernel_size[0], conv.stride[0]
    y = ops.conv2d(x, wpk, conv.out_channels, k, s, k // 2, scale, shift, res=res, act=act, out=out, w_kw=w_kw)
    return y if isinstance(y, View) else View(y, conv.out_channels, 0)


def _block(cc, blk, x, residual, out=None):
    h = _conv_bn_act(cc, x, blk.conv1, blk.bn1, act=1)
    return _conv_bn_act(cc, h, blk.conv2, blk.bn2, act=1, res=residual, out=out)


def _tree(cc, t, x, cat=None, dst=None):
    """Tree.forward (pose_dla_dcn.py:252-265).  `cat`: pre-allocated Root concat buffer whose child slots
    beyond this tree's own (x2, x1[, bottom]) were already filled by the caller; `dst`: where the result goes."""
    x = ops.as_view(x)
    Ho, Wo = x.H // t.stride, x.W // t.stride
    if t.levels == 1:
        co = t.co
        if cat is None:
            cat = _new(x, Ho, Wo, t.root_dim)
        bottom = x
        if t.level_root:          # children.append(bottom): lives in the concat buffer right after (x2, x1)
            bottom = ops.maxpool2d(x, t.stride, out=View(cat, t.ci, 2 * co))
        elif t.downsample is not None:
            bottom = View(ops.maxpool2d(x, t.stride), t.ci, 0)
        residual = _conv_bn_act(cc, bottom, t.project[0], t.project[1], act=0) if t.project is not None else bottom
        x1 = _block(cc, t.tree1, x, residual, out=View(cat, co, co))
        _block(cc, t.tree2, x1, x1, out=View(cat, co, 0))
        res = View(cat, co, 0) if t.root.residual else None
        return _conv_bn_act(cc, View(cat, cat.shape[3], 0), t.root.conv, t.root.bn, act=1, res=res, out=dst)
    # levels == 2: tree2's Root consumes (x2b, x1b, [bottom,] x1); the outer `project` result is dead code in
    # the reference (Tree.forward overwrites its `residual` argument, :255) and is skipped here.
    assert cat is None
    inner = t.tree2
    cbuf = _new(x, Ho, Wo, inner.root_dim)
    off = 2 * t.co
    if t.level_root:
        ops.maxpool2d(x, t.stride, out=View(cbuf, t.ci, off))
        off += t.ci
    x1 = _tree(cc, t.tree1, x, dst=View(cbuf, t.co, off))
    return _tree(cc, inner, x1, cat=cbuf, dst=dst)


def _deform(cc, dc, x, out=None):
    scale, shift = cc.deform(dc)
    y = dc.conv.forward_nhwc(x, scale=scale, shift=shift, act=1, out=out)
    return y if isinstance(y, View) else View(y, dc.conv.out_channels, 0)


def _ida(cc, ida, layers, startp, endp):
    """IDAUp.forward (pose_dla_dcn.py:482-488): layers[i] = node(up(proj(layers[i])) + layers[i-1])."""
    for i in range(startp + 1, endp):
        j = i - startp
        f = ida.factors[j]
        p = _deform(cc, getattr(ida, f"proj_{j}"), layers[i])
        prev = layers[i - 1]
        assert prev.coffset == 0 and prev.cstride == prev.C
        u = ops.dw_deconv_up(p, cc.up(getattr(ida, f"up_{j}"), f), f, add=prev.buf)
        layers[i] = _deform(cc, getattr(ida, f"node_{j}"), View(u, ida.o, 0))


class DLASeg(nn.Module):              # pose_dla_dcn.py:532-570
    def __init__(self, base_name="dla34", pretrained=False, down_ratio=4, final_kernel=1, last_level=5,
                 out_channel=0):
        super().__init__()
        assert base_name == "dla34" and down_ratio in (2, 4, 8, 16)
        self.first_level = int(np.log2(down_ratio))
        self.last_level = last_level
        self.base = DLA([1, 1, 1, 2, 2, 1], [16, 32, 64, 128, 256, 512])
        ch = self.base.channels
        scales = [2 ** i for i in range(len(ch[self.first_level:]))]
        self.dla_up = DLAUp(self.first_level, ch[self.first_level:], scales)
        self.out_channels = out_channel or ch[self.first_level]
        self.ida_up = IDAUp(self.out_channels, ch[self.first_level:self.last_level],
                            [2 ** i for i in range(self.last_level - self.first_level)])
        self._cc = None
        if pretrained:
            raise RuntimeError("ImageNet weights for dla34 need a network fetch (pose_dla_dcn.py:380-396); "
                               "load a checkpoint with load_state_dict instead")

    # Packed / folded copies are stamped with the parameters' versions (ops.PackCache) and rebuild themselves; only
    # a device / dtype move (`_apply`) or a write through `.data` needs the explicit reset below.
    def invalidate_caches(self):
        self._cc = None
        for m in self.modules():
            if isinstance(m, DCN):
                m._packed.clear()

    def _apply(self, fn, *a, **k):
        self.invalidate_caches()
        return super()._apply(fn, *a, **k)

    def forward_nhwc(self, x):
        """x [B,3,H,W] fp32 (CUDA) -> View of the [B,H/4,W/4,64] bf16 NHWC feature map."""
        if self._cc is None:
            self._cc = _Compiled()
        cc, b = self._cc, self.base
        if x.shape[3] % 2 == 0 and os.environ.get("CNB_STEM_S2D", "1") != "0":
            # 3 -> 4 channels; pixel pairs are the 8-channel super-pixels of the space-to-depth stem
            wpk, geom, scale2, shift2 = cc.stem_s2d(b.base_layer[0], b.base_layer[1])
            h = View(ops.stem_s2d(ops.to_nhwc_bf16(x, c_pad=4), wpk, geom, b.base_layer[0].out_channels, scale2,
                                  shift2), b.base_layer[0].out_channels, 0)
        else:
            h = View(ops.to_nhwc_bf16(x, c_pad=8), 8, 0)                   # 3 -> 8 channels for the 7x7 stem
            h = _conv_bn_act(cc, h, b.base_layer[0], b.base_layer[1])
        feats = []
        for lvl in (b.level0, b.level1):
            for i in range(0, len(lvl), 3):
                h = _conv_bn_act(cc, h, lvl[i], lvl[i + 1])
            feats.append(h)
        for lvl in (b.level2, b.level3, b.level4, b.level5):
            h = _tree(cc, lvl, h)
            feats.append(h)
        # DLAUp.forward (pose_dla_dcn.py:510-516)
        layers = list(feats)
        outs = [layers[-1]]
        n = len(layers)
        for i in range(n - self.first_level - 1):
            _ida(cc, getattr(self.dla_up, f"ida_{i}"), layers, n - i - 2, n)
            outs.insert(0, layers[-1])
        y = outs[:self.last_level - self.first_level]     # the reference clones; nothing here mutates in place
        _ida(cc, self.ida_up, y, 0, len(y))
        return y[-1]

    def forward(self, x):
        """[B,3,H,W] fp32 -> [Tensor[B,64,H/4,W/4] fp32] (the plugin contract, pose_dla_dcn.py:561-570).
        eval: the fused NHWC bf16 inference schedule above; train: the same operators on the autograd tape with
        batch-statistics BatchNorm (exec_modes.TrainBackend); precision "fp32-strict": fp32 CUDA-core kernels."""
        from .. import exec_modes
        _lib_require_cuda(x)
        if exec_modes.precision() == "fp32-strict":
            if self.training:
                raise RuntimeError("fp32-strict is an evaluation mode (parity checks); call .eval()")
            return [exec_modes.run_dla_seg(self, x, exec_modes.StrictBackend())]
        if self.training:
            return [exec_modes.run_dla_seg(self, x, exec_modes.TrainBackend())]
        v = self.forward_nhwc(x)
        out = ops.to_nchw_f32(v)
        out._cnb_nhwc = v            # lets CenterHead skip the NCHW fp32 -> NHWC bf16 round trip
        return [out]


def get_pose_net(num_layers, down_ratio=4):
    """models/backbones/pose_dla_dcn.py:573-581 (without the ImageNet download)."""
    assert num_layers == 34, "only dla_34 is built (the reference's default and only tested DLA)"
    return DLASeg("dla34", pretrained=False, down_ratio=down_ratio, final_kernel=1, last_level=5)
