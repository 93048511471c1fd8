"""The two execution modes beside the fused inference engine, sharing ONE traversal of the module tree:

* `TrainBackend`  -- training: NHWC bf16 tensors on the autograd tape, every operator an
  `autograd_ops` Function (tcgen05 GEMMs for forward / dX / dW, train-mode BatchNorm kernels, DCNv2 on sampled columns);
* `StrictBackend` -- `fp32-strict`: NCHW fp32 tensors, every operator in fp32 on the CUDA cores (csrc/strict_f32.cu);
  the precision of the reference itself, for small-shape end-to-end parity (hard part 2).

The traversal restates the control flow of the reference modules (file:line under CenterNet/models/backbones/):
`DLA.forward` pose_dla_dcn.py:372-378, `Tree.forward` :252-265, `BasicBlock.forward` :49-68, `Root.forward` :179-188,
`IDAUp.forward` :482-488, `DLAUp.forward` :510-516, `DLASeg.forward` :561-570; `PoseResNet.forward`
resnet_dcn.py:236-249 / msra_resnet.py:194-207; `HeadConv.forward` heads.py:24-25.
"""
import os

import torch

from .. import _lib, autograd_ops as ag, ops

_PRECISION = os.environ.get("CNB_PRECISION", "bf16")


def set_precision(mode):
    """"bf16" (default: NHWC bf16 tcgen05 engine, fp32 accumulation) or "fp32-strict" (everything in fp32 on the CUDA
    cores; eval only).  Also settable with CNB_PRECISION."""
    global _PRECISION
    if mode not in ("bf16", "fp32-strict"):
        raise ValueError(f"unknown precision mode {mode!r}")
    _PRECISION = mode


def precision():
    return _PRECISION


# ---------------------------------------------------------------------------------------------------------------
class TrainBackend:
    """NHWC bf16 + autograd.  BatchNorm uses batch statistics and updates the running ones (module.training)."""
    training = True

    def input(self, x):
        return ops.to_nhwc_bf16(x, c_pad=8)

    def output(self, v):
        out = ag.to_nchw_f32(v)
        out._cnb_nhwc = v
        return out

    def conv_bn_act(self, x, conv, bn, act=1, res=None):
        return ag.bn_act(ag.conv(x, conv), bn, res=res, act=act)

    def conv_bias_act(self, x, conv, act=0, out_nchw=False):
        return ag.conv(x, conv, act=act, out_mode=1 if out_nchw else 0)

    def maxpool(self, x, k):
        assert k == 2, "training path: MaxPool2d(2, 2) only (DLA downsample)"
        return ag.maxpool2(x)

    def cat(self, xs):
        return torch.cat(xs, dim=3)

    def deform(self, x, dc):                       # DeformConv: DCN -> BN -> ReLU
        return ag.bn_act(ag.dcn(x, dc.conv), dc.actf[0], act=1)

    def dcn_bn_act(self, x, dcn, bn):
        return ag.bn_act(ag.dcn(x, dcn), bn, act=1)

    def up_add(self, x, up, f, add):
        return ag.dw_up(x, up, f, add=add)

    def head_input(self, x):
        if isinstance(x, torch.Tensor) and x.dtype == torch.float32:
            v = getattr(x, "_cnb_nhwc", None)
            return v if v is not None else ag.to_nhwc_bf16(x)
        return x


# ---------------------------------------------------------------------------------------------------------------
def _f32(t):
    return t.detach().float().contiguous() if t is not None else None


class StrictBackend:
    """NCHW fp32 on the CUDA cores; eval-mode BatchNorm folded exactly as ATen's CPU kernel does
    (w = gamma * invstd, b = beta - mean * w)."""
    training = False

    def input(self, x):
        _lib.require_cuda(x)
        return x.float().contiguous()

    def output(self, v):
        return v

    @staticmethod
    def _fold(bn, bias=None):
        with torch.no_grad():
            invstd = 1.0 / torch.sqrt(bn.running_var.float() + bn.eps)
            w = bn.weight.float() * invstd
            b = bn.bias.float() - bn.running_mean.float() * w
            if bias is not None:
                b = b + bias.float() * w
        return w.contiguous(), b.contiguous()

    @staticmethod
    def _conv(x, weight, scale, shift, res, stride, pad, act):
        B, Ci, H, W = x.shape
        Co, _, KH, KW = weight.shape
        Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
        y = torch.empty((B, Co, Ho, Wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cnb_strict_conv2d_f32(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(scale), _lib.ptr(shift),
                                                        _lib.ptr(res), _lib.ptr(y), B, Ci, H, W, Co, KH, KW, stride, pad,
                                                        act, _lib.stream_ptr(x.device)), "cnb_strict_conv2d_f32")
        return y

    def conv_bn_act(self, x, conv, bn, act=1, res=None):
        scale, shift = self._fold(bn, conv.bias)
        return self._conv(x, _f32(conv.weight), scale, shift, res, conv.stride[0], conv.padding[0], act)

    def conv_bias_act(self, x, conv, act=0, out_nchw=False):
        return self._conv(x, _f32(conv.weight), None, _f32(conv.bias), None, conv.stride[0], conv.padding[0], act)

    def maxpool(self, x, k, stride=None, pad=0):
        stride = stride or k
        B, C, H, W = x.shape
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((B, C, Ho, Wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cnb_strict_maxpool2d_f32(_lib.ptr(x), _lib.ptr(y), B * C, H, W, k, stride, pad,
                                                           _lib.stream_ptr(x.device)), "cnb_strict_maxpool2d_f32")
        return y

    def cat(self, xs):
        return torch.cat(xs, dim=1)

    def _dcn(self, x, dcn, scale, shift, act):
        B, C, H, W = x.shape
        om = self.conv_bias_act(x, dcn.conv_offset_mask)
        col = torch.empty((B, C * 9, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cnb_strict_dcn_im2col_f32(_lib.ptr(x), _lib.ptr(om), _lib.ptr(col), B, C, H, W,
                                                            _lib.stream_ptr(x.device)), "cnb_strict_dcn_im2col_f32")
        w = _f32(dcn.weight).reshape(dcn.weight.shape[0], C * 9, 1, 1)
        return self._conv(col, w, scale, shift, None, 1, 0, act)

    def deform(self, x, dc):
        scale, shift = self._fold(dc.actf[0], dc.conv.bias)
        return self._dcn(x, dc.conv, scale, shift, 1)

    def dcn_bn_act(self, x, dcn, bn):
        scale, shift = self._fold(bn, dcn.bias)
        return self._dcn(x, dcn, scale, shift, 1)

    def conv_transpose(self, x, up, scale=None, shift=None, add=None, act=0):
        B, Ci, H, W = x.shape
        K, s, p = up.kernel_size[0], up.stride[0], up.padding[0]
        Co = up.out_channels
        depthwise = int(up.groups == Ci and up.groups > 1)
        Ho, Wo = (H - 1) * s - 2 * p + K, (W - 1) * s - 2 * p + K
        y = torch.empty((B, Co, Ho, Wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cnb_strict_conv_transpose2d_f32(
                _lib.ptr(x), _lib.ptr(_f32(up.weight)), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(add), _lib.ptr(y), B, Ci,
                H, W, Co, K, s, p, depthwise, act, _lib.stream_ptr(x.device)), "cnb_strict_conv_transpose2d_f32")
        return y

    def up_add(self, x, up, f, add):
        return self.conv_transpose(x, up, add=add)

    def head_input(self, x):
        return x.float().contiguous()


# ---------------------------------------------------------------------------------------------------------------
# DLA-34 + DLAUp / IDAUp
# ---------------------------------------------------------------------------------------------------------------
def _block(be, blk, x, residual):
    h = be.conv_bn_act(x, blk.conv1, blk.bn1, act=1)
    return be.conv_bn_act(h, blk.conv2, blk.bn2, act=1, res=residual)


def _tree(be, t, x, children=None):
    children = [] if children is None else children
    bottom = be.maxpool(x, t.stride) if t.downsample is not None else x
    if t.levels == 1:
        residual = be.conv_bn_act(bottom, t.project[0], t.project[1], act=0) if t.project is not None else bottom
    elif t.project is not None and be.training:
        # Tree.forward computes `self.project(bottom)` here too and then never uses it (:255 overwrites nothing, the
        # inner tree ignores the argument): dead for the result, but in train mode its BatchNorm still updates the
        # running statistics -- reproduced for state-dict fidelity, off the tape.
        with torch.no_grad():
            be.conv_bn_act(bottom, t.project[0], t.project[1], act=0)
    if t.level_root:
        children.append(bottom)
    if t.levels == 1:
        x1 = _block(be, t.tree1, x, residual)
        x2 = _block(be, t.tree2, x1, x1)
        xs = [x2, x1] + children
        return be.conv_bn_act(be.cat(xs), t.root.conv, t.root.bn, act=1, res=xs[0] if t.root.residual else None)
    x1 = _tree(be, t.tree1, x)
    children.append(x1)
    return _tree(be, t.tree2, x1, children=children)


def _ida(be, ida, layers, startp, endp):
    for i in range(startp + 1, endp):
        j = i - startp
        p = be.deform(layers[i], getattr(ida, f"proj_{j}"))
        u = be.up_add(p, getattr(ida, f"up_{j}"), ida.factors[j], layers[i - 1])
        layers[i] = be.deform(u, getattr(ida, f"node_{j}"))


def run_dla_seg(net, x, be):
    """DLASeg.forward on backend `be`; returns the backend's feature tensor [B,64,H/4,W/4]-equivalent."""
    b = net.base
    h = be.conv_bn_act(be.input(x), b.base_layer[0], b.base_layer[1], act=1)
    feats = []
    for lvl in (b.level0, b.level1):
        for i in range(0, len(lvl), 3):
            h = be.conv_bn_act(h, lvl[i], lvl[i + 1], act=1)
        feats.append(h)
    for lvl in (b.level2, b.level3, b.level4, b.level5):
        h = _tree(be, lvl, h)
        feats.append(h)
    layers = list(feats)
    outs = [layers[-1]]
    n = len(layers)
    for i in range(n - net.first_level - 1):
        _ida(be, getattr(net.dla_up, f"ida_{i}"), layers, n - i - 2, n)
        outs.insert(0, layers[-1])
    y = outs[:net.last_level - net.first_level]
    _ida(be, net.ida_up, y, 0, len(y))
    return be.output(y[-1])


# ---------------------------------------------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------------------------------------------
def run_heads(head, x, be):
    """CenterHead.forward: name -> NCHW fp32 map."""
    v = be.head_input(x)
    ret = {}
    for name in head.heads:
        m = getattr(head, name)
        mid = be.conv_bias_act(v, m.fc[0], act=1)
        ret[name] = be.conv_bias_act(mid, m.fc[2], act=0, out_nchw=True)
    return ret
