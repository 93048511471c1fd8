"""Training path: `torch.autograd.Function` wrappers over the C ABI (include/centernet_b200.h), NHWC bf16 tensors.

The reference trains through autograd + cuDNN (`CenterNet/centernet.py:70-80`: `training_step` -> `self.loss` ->
Lightning calls `loss.backward()`).  Here every operator of the hot path carries its own backward built from this
library's kernels:

    conv            fwd  cnb_conv2d_fprop                     (tcgen05 implicit GEMM)
                    dX   cnb_conv2d_fprop on the flipped / phase-decomposed filter (+ cnb_depth_to_space2 for stride 2)
                    dW   cnb_conv2d_wgrad (tcgen05, pixels as the contraction) + cnb_conv_unpack_wgrad
                    db   cnb_channel_sum
    BatchNorm(train)+residual+ReLU   cnb_bn_train_fwd / cnb_bn_train_bwd  (batch statistics, running-stat update)
    DCNv2           cnb_dcnv2_im2col -> 1x1 GEMM;  backward: 1x1 GEMM (dcol), cnb_dcnv2_col2im (dX, d offset, d mask),
                    cnb_conv2d_wgrad on the columns
    depthwise up-sampling            cnb_dw_deconv_up / cnb_dw_deconv_bwd
    MaxPool2d(2,2)                   cnb_maxpool2d / cnb_maxpool2x2_bwd

autograd is used for what it is: the tape.  Gradient accumulation of multiply-used activations is autograd's own.
A parameter that carries `_cnb_grad` (a view of the trainer's flat fp32 gradient buffer, trainer.py) receives its
gradient directly there (no intermediate tensor) and its `_cnb_ready` callback fires -- that is what lets the
bucketed all-reduce start while the backward pass is still running.
"""
import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import ConvDesc

BF = torch.bfloat16


def _st(t):
    return _lib.stream_ptr(t.device)


def _deliver(param, grad_fn):
    """Hand a parameter gradient over.  grad_fn(out, accumulate) writes the gradient into `out` (fp32, the parameter's
    shape).  Returns the tensor autograd should see, or None when it went straight into the trainer's flat buffer."""
    flat = getattr(param, "_cnb_grad", None)
    if flat is not None:
        grad_fn(flat, 1)
        ready = getattr(param, "_cnb_ready", None)
        if ready is not None:
            ready()
        return None
    out = torch.empty_like(param, dtype=torch.float32)
    grad_fn(out, 0)
    return out


def _channel_sum(dy, C, out, accumulate, y_mask=None, cstride=None, coffset=0):
    M = dy.numel() // (cstride or dy.shape[-1])
    with torch.cuda.device(dy.device):
        _lib.check(_lib.lib().cnb_channel_sum(_lib.ptr(dy), cstride or dy.shape[-1], coffset, _lib.ptr(y_mask),
                                              _lib.ptr(out), accumulate, M, C, _st(dy)), "cnb_channel_sum")


def _pad_channels(c):
    """Channel count the tensor-core kernels accept as a GEMM K slab: 8, 16, 32 or a multiple of 64."""
    for v in (8, 16, 32):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


# ---------------------------------------------------------------------------------------------------------------
# packed copies of a conv's weights for the three GEMMs (rebuilt when the parameter changes: ops.PackCache)
# ---------------------------------------------------------------------------------------------------------------
def _conv_packs(mod, co_pad):
    """(w_fwd, w_dgrad, w_kw) for an nn.Conv2d.  Stride 1: flipped/transposed filter; stride 2 (3x3, pad 1): the
    four output phases as channel blocks of one 3x3 stride-1 filter (see include/centernet_b200.h)."""
    cache = mod.__dict__.setdefault("_cnb_packs", ops.PackCache())
    w = mod.weight
    Co, Ci, KH, KW = w.shape

    def build():
        with torch.no_grad():
            w_kw = KW + 1 if (Ci <= 8 and KW % 2 == 1 and KW > 1) else 0
            fwd = ops.pack_conv_weights(w, kw_pad=w_kw or None)
            wf = w.detach().float()
            if co_pad > Co:
                wf = torch.cat([wf, wf.new_zeros(co_pad - Co, Ci, KH, KW)], 0)
            if Ci < 8:
                dg = fwd.new_zeros(1)          # network input: no data gradient
            elif mod.stride[0] == 1:
                dg = ops.pack_conv_weights(wf.flip(2, 3).transpose(0, 1).contiguous())
            else:
                assert mod.stride[0] == 2 and KH == 3 and KW == 3 and mod.padding[0] == 1, \
                    "stride-2 data gradient is built for 3x3 / pad 1 convolutions"
                w3 = wf.new_zeros(2, 2, Ci, co_pad, 3, 3)
                for a in (0, 1):
                    for ty in range(3):
                        kh = 3 - 2 * ty + a
                        if not 0 <= kh <= 2:
                            continue
                        for b in (0, 1):
                            for tx in range(3):
                                kw = 3 - 2 * tx + b
                                if 0 <= kw <= 2:
                                    w3[a, b, :, :, ty, tx] = wf[:, :, kh, kw].t()
                dg = ops.pack_conv_weights(w3.reshape(4 * Ci, co_pad, 3, 3))
        return (fwd, dg, w_kw)
    return cache.get(("conv", co_pad), (w,), build)


def _wgrad(desc, x_buf, dy_buf, dy_cstride, dy_coffset, Co, Kacc):
    acc = torch.zeros(Co * Kacc, dtype=torch.float32, device=x_buf.device)
    with torch.cuda.device(x_buf.device):
        _lib.check(_lib.lib().cnb_conv2d_wgrad(desc, _lib.ptr(x_buf), _lib.ptr(dy_buf), dy_cstride, dy_coffset,
                                               _lib.ptr(acc), _st(x_buf)), "cnb_conv2d_wgrad")
    return acc


def _unpack_wgrad(acc, out, Co, Ci, Ci_pad, KH, KW, KWp, accumulate):
    with torch.cuda.device(acc.device):
        _lib.check(_lib.lib().cnb_conv_unpack_wgrad(_lib.ptr(acc), _lib.ptr(out), Co, Ci, Ci_pad, KH, KW, KWp, accumulate,
                                                    1.0, _st(acc)), "cnb_conv_unpack_wgrad")


class _ConvFn(Function):
    """y = act(conv(x, W) + b).  x: NHWC bf16 [B,H,W,Ci_pad] or an ops.View of a wider buffer (heads).
    out_mode 0: NHWC bf16; 1: NCHW fp32 (head maps); 2: NHWC fp32 with channel stride 32 (DCN offset/mask maps)."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod, act, out_mode, x_view):
        Co, Ci, KH, KW = weight.shape
        k, s, p = KH, mod.stride[0], mod.padding[0]
        co_pad = Co if out_mode == 0 else _pad_channels(Co)
        fwd, dg, w_kw = _conv_packs(mod, co_pad)
        xv = ops.View(x, x_view[0], x_view[1]) if x_view is not None else ops.as_view(x)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        y = ops.conv2d(xv, fwd, Co, k, s, p, None, b32, act=act, out_mode=out_mode, w_kw=w_kw)
        ctx.mod, ctx.act, ctx.out_mode, ctx.x_view, ctx.w_kw, ctx.co_pad = mod, act, out_mode, x_view, w_kw, co_pad
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, y if act == 1 else None)
        ctx.geom = (xv.B, xv.H, xv.W, xv.C, xv.cstride, xv.coffset, y.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        mod, out_mode, co_pad = ctx.mod, ctx.out_mode, ctx.co_pad
        weight, bias = mod.weight, mod.bias
        Co, Ci, KH, KW = weight.shape
        B, H, W, Cx, x_cs, x_co, yshape = ctx.geom
        k, s, p = KH, mod.stride[0], mod.padding[0]
        L = _lib.lib()
        dev = x.device
        # ---- dY as NHWC bf16 [B,Ho,Wo,co_pad]
        if out_mode == 0:
            dy = dy.contiguous()
            if ctx.act == 1:
                g = torch.empty_like(dy)
                with torch.cuda.device(dev):
                    _lib.check(L.cnb_relu_bwd(_lib.ptr(y), _lib.ptr(dy), _lib.ptr(g), dy.numel(), _st(dy)), "cnb_relu_bwd")
                dy = g
            Ho, Wo = dy.shape[1], dy.shape[2]
        elif out_mode == 1:
            assert ctx.act == 0
            Ho, Wo = dy.shape[2], dy.shape[3]
            dy = ops.to_nhwc_bf16(dy.float().contiguous(), c_pad=co_pad)
        else:
            assert ctx.act == 0
            Ho, Wo = dy.shape[1], dy.shape[2]
            g = torch.empty(dy.shape, dtype=BF, device=dev)
            dyc = dy.contiguous()
            with torch.cuda.device(dev):
                _lib.check(L.cnb_f32_to_bf16(_lib.ptr(dyc), None, _lib.ptr(g), dyc.numel(), dyc.shape[-1], Co, _st(dyc)),
                           "cnb_f32_to_bf16")
            dy = g
        assert dy.shape[-1] == co_pad, (dy.shape, co_pad)
        fwd, dg, w_kw = _conv_packs(mod, co_pad)
        # ---- dX
        dx = None
        if ctx.needs_input_grad[0]:
            assert ctx.x_view is None or (x_co == 0 and x_cs == Cx), "data gradient into a channel slice is not built"
            if s == 1:
                dx = ops.conv2d(dy, dg, Cx, k, 1, k - 1 - p, None, None, act=0)
            else:
                assert H % 2 == 0 and W % 2 == 0
                d4 = ops.conv2d(dy, dg, 4 * Cx, 3, 1, 1, None, None, act=0)
                dx = torch.empty((B, H, W, Cx), dtype=BF, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(L.cnb_depth_to_space2(_lib.ptr(d4), _lib.ptr(dx), B, Ho, Wo, Cx, _st(dx)),
                               "cnb_depth_to_space2")
        # ---- dW, db
        dw = db = None
        if ctx.needs_input_grad[1]:
            xv = ops.View(x, Cx, x_co)
            d = ops._make_desc(xv, Co, k, s, p, 1, Ho, Wo, 0, 0, 0, 0, None, w_kw)
            kwp = w_kw or KW
            acc = _wgrad(d, x, dy, co_pad, 0, Co, KH * kwp * Cx)
            dw = _deliver(weight, lambda out, a: _unpack_wgrad(acc, out, Co, Ci, Cx, KH, KW, kwp, a))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _deliver(bias, lambda out, a: _channel_sum(dy, Co, out, a, cstride=co_pad))
        return dx, dw, db, None, None, None, None


def conv(x, mod, act=0, out_mode=0, x_view=None):
    """nn.Conv2d `mod` applied to NHWC bf16 `x` (x_view = (C, coffset) selects a channel slice of a wider buffer)."""
    return _ConvFn.apply(x, mod.weight, mod.bias, mod, act, out_mode, x_view)


class _BnActFn(Function):
    """y = act(BatchNorm_train(z) (+ res)); updates the module's running statistics like nn.BatchNorm2d."""

    @staticmethod
    def forward(ctx, z, gamma, beta, res, bn, act):
        z = z.contiguous()
        C = z.shape[-1]
        M = z.numel() // C
        y = torch.empty_like(z)
        stats = torch.empty(4 * C, dtype=torch.float32, device=z.device)
        ws = torch.empty(2 * C, dtype=torch.float32, device=z.device)
        res_c = res.contiguous() if res is not None else None
        mom = bn.momentum if bn.momentum is not None else 0.1
        track = bn.track_running_stats and bn.running_mean is not None
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().cnb_bn_train_fwd(
                _lib.ptr(z), _lib.ptr(gamma.detach()), _lib.ptr(beta.detach()),
                _lib.ptr(bn.running_mean) if track else None, _lib.ptr(bn.running_var) if track else None,
                mom, bn.eps, _lib.ptr(res_c), act, _lib.ptr(y), _lib.ptr(stats), _lib.ptr(ws), M, C, _st(z)),
                "cnb_bn_train_fwd")
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        ctx.bn, ctx.act, ctx.has_res = bn, act, res is not None
        ctx.save_for_backward(z, y if act == 1 else None, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, y, stats = ctx.saved_tensors
        bn = ctx.bn
        C = z.shape[-1]
        M = z.numel() // C
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        dres = torch.empty_like(z) if (ctx.has_res and ctx.needs_input_grad[3]) else None
        ws = torch.empty(2 * C, dtype=torch.float32, device=z.device)
        flat_g, flat_b = getattr(bn.weight, "_cnb_grad", None), getattr(bn.bias, "_cnb_grad", None)
        direct = flat_g is not None and flat_b is not None
        dg = flat_g if direct else torch.empty(C, dtype=torch.float32, device=z.device)
        db = flat_b if direct else torch.empty(C, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().cnb_bn_train_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(z), _lib.ptr(stats), _lib.ptr(dz),
                                                   _lib.ptr(dres), _lib.ptr(dg), _lib.ptr(db), 1 if direct else 0,
                                                   _lib.ptr(ws), M, C, _st(z)), "cnb_bn_train_bwd")
        if direct:
            for p in (bn.weight, bn.bias):
                ready = getattr(p, "_cnb_ready", None)
                if ready is not None:
                    ready()
            dg = db = None
        if dres is None and ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy
        return dz, dg, db, dres, None, None


def bn_act(z, bn, res=None, act=1):
    return _BnActFn.apply(z, bn.weight, bn.bias, res, bn, act)


class _MaxPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool2d(x, 2)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        dx = torch.empty_like(x)
        dy = dy.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cnb_maxpool2x2_bwd(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), B, H, W, C, _st(x)),
                       "cnb_maxpool2x2_bwd")
        return dx


def maxpool2(x):
    return _MaxPool2Fn.apply(x)


class _DwUpFn(Function):
    """depthwise bilinear ConvTranspose2d(2f, stride f, pad f/2) + add (pose_dla_dcn.py:466-488)."""

    @staticmethod
    def forward(ctx, x, weight, add, mod, f):
        x = x.contiguous()
        wt = ops.relayout_dw_weights(weight, f)
        y = ops.dw_deconv_up(x, wt, f, add=add.contiguous() if add is not None else None)
        ctx.mod, ctx.f = mod, f
        ctx.save_for_backward(x, wt)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wt = ctx.saved_tensors
        f, weight = ctx.f, ctx.mod.weight
        B, H, W, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        acc = torch.zeros(4 * f * f * C, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            _lib.check(L.cnb_dw_deconv_bwd(_lib.ptr(x), _lib.ptr(wt), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(acc), B, H, W, C,
                                           f, _st(x)), "cnb_dw_deconv_bwd")

        def unpack(out, a):
            with torch.cuda.device(x.device):
                _lib.check(L.cnb_dw_deconv_unpack_wgrad(_lib.ptr(acc), _lib.ptr(out), C, f, a, _st(x)),
                           "cnb_dw_deconv_unpack_wgrad")
        dw = _deliver(weight, unpack) if ctx.needs_input_grad[1] else None
        return dx, dw, (dy if ctx.needs_input_grad[2] else None), None, None


def dw_up(x, mod, f, add=None):
    return _DwUpFn.apply(x, mod.weight, add, mod, f)


class _DcnFn(Function):
    """DCNv2 main op on materialised columns: z = conv1x1(im2col(x; om), W) + b  (pre-BatchNorm output)."""

    @staticmethod
    def forward(ctx, x, om, weight, bias, mod):
        x = x.contiguous()
        B, H, W, C = x.shape
        Co = weight.shape[0]
        L = _lib.lib()
        col = torch.empty((B, H, W, 9 * C), dtype=BF, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.cnb_dcnv2_im2col(_lib.ptr(x), _lib.ptr(om), om.shape[-1], _lib.ptr(col), B, H, W, C, _st(x)),
                       "cnb_dcnv2_im2col")
        w_main, w_t = _dcn_packs(mod)
        z = ops.conv2d(col, w_main, Co, 1, 1, 0, None, bias.detach().float().contiguous(), act=0)
        ctx.mod = mod
        ctx.save_for_backward(x, om, col)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, om, col = ctx.saved_tensors
        mod = ctx.mod
        weight, bias = mod.weight, mod.bias
        B, H, W, C = x.shape
        Co = weight.shape[0]
        dz = dz.contiguous()
        L = _lib.lib()
        w_main, w_t = _dcn_packs(mod)
        dx = dom = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dcol = ops.conv2d(dz, w_t, 9 * C, 1, 1, 0, None, None, act=0)
            dx_acc = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
            dom = torch.empty(om.shape, dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.check(L.cnb_dcnv2_col2im(_lib.ptr(x), _lib.ptr(om), om.shape[-1], _lib.ptr(dcol), _lib.ptr(dx_acc),
                                              _lib.ptr(dom), B, H, W, C, _st(x)), "cnb_dcnv2_col2im")
                dx = torch.empty_like(x)
                _lib.check(L.cnb_f32_to_bf16(_lib.ptr(dx_acc), None, _lib.ptr(dx), dx_acc.numel(), C, 0, _st(x)),
                           "cnb_f32_to_bf16")
        dw = db = None
        if ctx.needs_input_grad[2]:
            cv = ops.View(col, 9 * C, 0)
            d = ops._make_desc(cv, Co, 1, 1, 0, 1, H, W, 0, 0, 0, 0, None, 0)
            acc = _wgrad(d, col, dz, Co, 0, Co, 9 * C)
            dw = _deliver(weight, lambda out, a: _unpack_wgrad(acc, out, Co, C, C, 3, 3, 3, a))
        if ctx.needs_input_grad[3]:
            db = _deliver(bias, lambda out, a: _channel_sum(dz, Co, out, a))
        return dx, dom, dw, db, None


def _dcn_packs(mod):
    """(W as [Co][9*Ci] -- the forward pack, (tap, ci) order --, W^T as the 1x1 filter Co -> 9*Ci of the dcol GEMM)"""
    cache = mod.__dict__.setdefault("_cnb_train_packs", ops.PackCache())
    w = mod.weight

    def build():
        with torch.no_grad():
            Co, Ci = w.shape[0], w.shape[1]
            wt = w.detach().float().permute(2, 3, 1, 0).reshape(9 * Ci, Co, 1, 1).contiguous()
            return (ops.pack_conv_weights(w), ops.pack_conv_weights(wt))
    return cache.get("dcn", (w,), build)


def dcn(x, mod):
    """DCN.dcn_v2.DCN forward for training: offset/mask conv (fp32 NHWC out) -> sampled-column GEMM; returns the
    pre-activation NHWC bf16 map (bias included)."""
    om = conv(x, mod.conv_offset_mask, act=0, out_mode=2)
    return _DcnFn.apply(x, om, mod.weight, mod.bias, mod)


class _ToNchwF32Fn(Function):
    """NHWC bf16 -> NCHW fp32 (the backbone's output contract); backward converts the gradient back."""

    @staticmethod
    def forward(ctx, v):
        ctx.C = v.shape[-1]
        return ops.to_nchw_f32(v.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return ops.to_nhwc_bf16(dy.float().contiguous(), c_pad=ctx.C)


def to_nchw_f32(v):
    return _ToNchwF32Fn.apply(v)


class _ToNhwcBf16Fn(Function):
    """NCHW fp32 -> NHWC bf16 (a head fed with a plain NCHW tensor); backward converts the gradient back."""

    @staticmethod
    def forward(ctx, x):
        return ops.to_nhwc_bf16(x.float().contiguous())

    @staticmethod
    def backward(ctx, dv):
        return ops.to_nchw_f32(dv.contiguous())


def to_nhwc_bf16(x):
    return _ToNhwcBf16Fn.apply(x)
