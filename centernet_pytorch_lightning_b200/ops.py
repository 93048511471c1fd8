"""Thin torch-tensor wrappers over the conv-engine entry points of the C ABI (include/centernet_b200.h).

Activations are NHWC bf16 torch tensors; a `View` names a channel slice of a (possibly wider) NHWC
buffer so that `torch.cat` along channels (Root, pose_dla_dcn.py:182) never has to be materialised.
Everything is asynchronous on the current CUDA stream; torch owns all memory.
"""
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ConvDesc


@dataclass
class View:
    """Channels [coffset, coffset + C) of an NHWC bf16 buffer [B,H,W,cstride]."""
    buf: torch.Tensor
    C: int
    coffset: int = 0

    @property
    def B(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def cstride(self):
        return self.buf.shape[3]

    def tensor(self):
        return self.buf[..., self.coffset:self.coffset + self.C]


def as_view(x) -> View:
    if isinstance(x, View):
        return x
    assert x.dim() == 4 and x.dtype == torch.bfloat16 and x.is_contiguous()
    return View(x, x.shape[3], 0)


def _stream(t):
    return _lib.stream_ptr(t.device)


class LaunchProfiler:
    """Optional per-launch CUDA-event timing of the tensor-core kernels (bench.py's roofline leg).
    `with LaunchProfiler() as p: model(x)` then `p.summary()` -> (total_ms, total_flops, n_launches)."""
    active = None

    def __init__(self):
        self.records = []

    def __enter__(self):
        LaunchProfiler.active = self
        return self

    def __exit__(self, *exc):
        LaunchProfiler.active = None

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, start, kind, flops):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((kind, flops, start, e))

    def summary(self, kind=None):
        torch.cuda.synchronize()
        rec = [r for r in self.records if kind is None or r[0] == kind]
        return sum(r[2].elapsed_time(r[3]) for r in rec), sum(r[1] for r in rec), len(rec)


def to_nhwc_bf16(x, c_pad: Optional[int] = None):
    """[B,C,H,W] fp32 -> [B,H,W,c_pad] bf16 (channels zero-padded to a multiple of 8)."""
    _lib.require_cuda(x)
    x = x.float().contiguous()
    B, C, H, W = x.shape
    c_pad = c_pad or (C + 7) // 8 * 8
    y = torch.empty((B, H, W, c_pad), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cnb_nchw_f32_to_nhwc_bf16(_lib.ptr(x), _lib.ptr(y), B, C, H, W, c_pad, _stream(x)),
                   "cnb_nchw_f32_to_nhwc_bf16")
    return y


def to_nchw_f32(v):
    v = as_view(v)
    y = torch.empty((v.B, v.C, v.H, v.W), dtype=torch.float32, device=v.buf.device)
    with torch.cuda.device(v.buf.device):
        _lib.check(_lib.lib().cnb_nhwc_bf16_to_nchw_f32(_lib.ptr(v.buf), _lib.ptr(y), v.B, v.C, v.H, v.W,
                                                         v.cstride, v.coffset, _stream(y)),
                   "cnb_nhwc_bf16_to_nchw_f32")
    return y


def param_key(*tensors):
    """(version, data_ptr) of every parameter a packed / folded copy was built from.  `_version` is bumped by every
    in-place update autograd can see (optimizer steps, `load_state_dict` -- also through a parent module --,
    `nn.init.*_`); writes through `.data` are invisible to it: call the model's `invalidate_caches()` after those."""
    return tuple((t._version, t.data_ptr()) for t in tensors if t is not None)


class PackCache:
    """Packed-weight / folded-BN copies keyed by the owning module, each entry stamped with `param_key` of the
    parameters it was built from and rebuilt when that changes.  A rebuilt entry is written IN PLACE into the old
    tensors when shapes allow, so CUDA graphs that captured their addresses (engine.py) see the new weights."""

    def __init__(self):
        self.entries = {}

    def get(self, key, params, build):
        stamp = param_key(*params)
        hit = self.entries.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        new = build()
        if hit is not None:
            old = hit[1]
            same = len(old) == len(new) and all(
                (not isinstance(o, torch.Tensor) and o == n) or
                (isinstance(o, torch.Tensor) and isinstance(n, torch.Tensor) and o.shape == n.shape and o.dtype == n.dtype)
                for o, n in zip(old, new))
            if same:
                with torch.no_grad():
                    for o, n in zip(old, new):
                        if isinstance(o, torch.Tensor):
                            o.copy_(n)
                new = old
        self.entries[key] = (stamp, new)
        return new

    def clear(self):
        self.entries.clear()


def pack_conv_weights(w, kw_pad=None):
    """[Co,Ci,KH,KW] fp32 -> packed bf16 [Co_pad][Kpad] (K order kh,kw,ci; Ci padded to 8).
    `kw_pad` > KW appends zero filter columns first (pass the same value as `w_kw` to conv2d)."""
    _lib.require_cuda(w)
    w = w.detach().float()
    if kw_pad is not None and kw_pad > w.shape[3]:
        w = torch.nn.functional.pad(w, (0, kw_pad - w.shape[3]))
    w = w.contiguous()
    Co, Ci, KH, KW = w.shape
    L = _lib.lib()
    nbytes = L.cnb_conv_packed_weight_bytes(Co, Ci, KH, KW)
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(L.cnb_conv_pack_weights(_lib.ptr(w), _lib.ptr(out), Co, Ci, KH, KW, _stream(w)),
                   "cnb_conv_pack_weights")
    return out


def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _out_geometry(H, W, k, stride, pad, dil=1):
    (kh, kw), (ph, pw) = _pair(k), _pair(pad)
    return (H + 2 * ph - dil * (kh - 1) - 1) // stride + 1, (W + 2 * pw - dil * (kw - 1) - 1) // stride + 1


def _make_desc(x: View, Co, k, stride, pad, dil, Ho, Wo, out_mode, act, y_cstride, y_coffset, res: Optional[View],
               w_kw=0):
    d = ConvDesc()
    d.B, d.Hi, d.Wi, d.Ci = x.B, x.H, x.W, x.C
    (kh, kw), (ph, pw) = _pair(k), _pair(pad)
    d.Co, d.KH, d.KW = Co, kh, kw
    d.stride, d.pad, d.dil = stride, ph, dil
    d.pad_w1 = pw + 1 if pw != ph else 0
    d.Ho, d.Wo = Ho, Wo
    d.x_cstride, d.x_coffset = x.cstride, x.coffset
    d.y_cstride, d.y_coffset = y_cstride, y_coffset
    d.res_cstride, d.res_coffset = (res.cstride, res.coffset) if res is not None else (0, 0)
    d.act, d.out_nchw_f32 = act, out_mode
    d.w_kw = w_kw
    return d


def _alloc_out(x: View, Co, Ho, Wo, out_mode, out):
    dev = x.buf.device
    if out is not None:
        out = as_view(out) if out_mode == 0 else out
        return out
    if out_mode == 0:
        return View(torch.empty((x.B, Ho, Wo, Co), dtype=torch.bfloat16, device=dev), Co, 0)
    if out_mode == 1:
        return torch.empty((x.B, Co, Ho, Wo), dtype=torch.float32, device=dev)
    cs = (Co + 15) // 16 * 16
    return torch.empty((x.B, Ho, Wo, cs), dtype=torch.float32, device=dev)


def conv2d(x, wpk, Co, k, stride, pad, scale, shift, res=None, act=0, out_mode=0, out=None, dil=1, w_kw=0,
           flops=None):
    """y = act(conv(x, w) * scale + shift (+ res)).  out_mode 0: NHWC bf16 (returns the tensor, or writes the
    `out` View), 1: NCHW fp32, 2: NHWC fp32 (channel stride rounded up to 16).  `k` / `pad` may be (h, w) pairs
    (rectangular filter: the space-to-depth stem)."""
    x = as_view(x)
    res = as_view(res) if res is not None else None
    Ho, Wo = _out_geometry(x.H, x.W, k, stride, pad, dil)
    o = _alloc_out(x, Co, Ho, Wo, out_mode, out)
    if out_mode == 0:
        y_cs, y_co, y_ptr = o.cstride, o.coffset, _lib.ptr(o.buf)
    elif out_mode == 1:
        y_cs, y_co, y_ptr = 0, 0, _lib.ptr(o)
    else:
        y_cs, y_co, y_ptr = o.shape[3], 0, _lib.ptr(o)
    d = _make_desc(x, Co, k, stride, pad, dil, Ho, Wo, out_mode, act, y_cs, y_co, res, w_kw)
    prof = LaunchProfiler.active
    with torch.cuda.device(x.buf.device):
        t0 = prof.begin() if prof else None
        _lib.check(_lib.lib().cnb_conv2d_fprop(d, _lib.ptr(x.buf), _lib.ptr(wpk), _lib.ptr(scale), _lib.ptr(shift),
                                                _lib.ptr(res.buf) if res is not None else None, y_ptr,
                                                _stream(x.buf)), "cnb_conv2d_fprop")
        if prof:   # algorithmic FLOPs: 2*MACs with the true input channels (the 7x7 stem pads 3 -> 8)
            kh, kw = _pair(k)
            ci = 3 if (kh == 7 and x.C == 8) else x.C
            prof.end(t0, "conv", flops if flops is not None else 2.0 * x.B * Ho * Wo * Co * kh * kw * ci)
    if out_mode == 0:
        return o.buf if (out is None) else o
    return o



def stem_s2d_filter(w):
    """[Co,Ci<=4,K,K] -> ([2*Co, 8, K, KX] fp32, (K, KX, D)): the rectangular filter of the space-to-depth stem (see
    pack_stem_s2d_weights); plain tensor arithmetic, runs on any device."""
    Co, Ci, K, K2 = w.shape
    assert K == K2 and K % 2 == 1 and Ci <= 4
    p = K // 2
    D = (p + 1) // 2
    KX = 2 * D + 1
    w = w.detach().float()
    ws = torch.zeros((2, Co, 2, 4, K, KX), dtype=torch.float32, device=w.device)   # [par, co, h, ci, kh, ds+D]
    for par in range(2):
        for kw in range(K):
            t = par + kw - p
            ds, h = t // 2, t % 2          # floor division: t = 2*ds + h
            ws[par, :, h, :Ci, :, ds + D] = w[:, :, :, kw]
    return ws.reshape(2 * Co, 8, K, KX), (K, KX, D)


def pack_conv_weights_dgrad(w):
    """Packed weights of the DATA-GRADIENT convolution of a stride-1 conv: dX = conv(dY, W') with
    W'[ci, co, kh, kw] = W[co, ci, KH-1-kh, KW-1-kw] and padding K-1-pad -- the same tcgen05 kernel as the forward.
    (First piece of the training path, SURVEY 8b `conv2d_dgrad`; stride-2 dgrad and wgrad are not built.)"""
    return pack_conv_weights(w.detach().float().flip(2, 3).transpose(0, 1).contiguous())


def conv2d_dgrad(dy, wpk_dgrad, Ci, k, pad, out=None):
    """dX [B,H,W,Ci] bf16 = d(conv2d stride 1)/dX applied to dY [B,H,W,Co] bf16 (NHWC), fp32 accumulation."""
    return conv2d(dy, wpk_dgrad, Ci, k, 1, k - 1 - pad, None, None, act=0, out=out)


def pack_stem_s2d_weights(w):
    """Space-to-depth packing of a stride-1 KxK filter over <= 4 input channels (the DLA 7x7 stem,
    pose_dla_dcn.py:281-285).  The input [B,H,W,4] bf16 is read as [B,H,W/2,8]: one 16-byte "super-pixel" holds
    two neighbouring pixels.  Output pixel x = 2X + par reads input pixel 2X + par + (kw - K//2), i.e. super-pixel
    X + ds, half h with 2*ds + h = par + kw - K//2.  The returned filter is [2*Co, 8, K, KX] (KX = 2*ceil(K//2 / 2) + 1
    taps over super-pixels, output channel par*Co + co, input channel h*4 + ci) packed for `stem_s2d`: per filter row
    3 tensor-core steps produce 256 output pixels x Co channels instead of 4 steps per 128 pixels."""
    ws, (K, KX, D) = stem_s2d_filter(w)
    kxp = KX + 1                           # even tap count: pairs of super-pixel taps form one K=16 step
    return pack_conv_weights(ws, kw_pad=kxp), (K, KX, D, kxp)


def stem_s2d(x4, wpk, geom, Co, scale2, shift2, act=1):
    """x4 [B,H,W,4] bf16 (to_nhwc_bf16(x, c_pad=4), W even) -> [B,H,W,Co] bf16 = act(conv_KxK(x) * scale + shift).
    `scale2` / `shift2` are the per-channel vectors repeated twice (one copy per output-pixel parity)."""
    K, KX, D, kxp = geom
    B, H, W, _ = x4.shape
    xs = View(x4.view(B, H, W // 2, 8), 8, 0)
    y = conv2d(xs, wpk, 2 * Co, (K, KX), 1, (K // 2, D), scale2, shift2, act=act, w_kw=kxp,
               flops=2.0 * B * H * W * Co * K * K * 3)
    return y.view(B, H, W, Co)

def dcnv2(x, om, wpk, Co, scale, shift, act=0, out=None):
    """Modulated deformable 3x3 conv: x NHWC bf16, om [B,H,W,>=27] fp32 NHWC raw offset/mask channels."""
    x = as_view(x)
    assert om.dtype == torch.float32 and om.is_contiguous() and om.shape[:3] == (x.B, x.H, x.W)
    o = _alloc_out(x, Co, x.H, x.W, 0, out)
    d = _make_desc(x, Co, 3, 1, 1, 1, x.H, x.W, 0, act, o.cstride, o.coffset, None)
    prof = LaunchProfiler.active
    with torch.cuda.device(x.buf.device):
        t0 = prof.begin() if prof else None
        _lib.check(_lib.lib().cnb_dcnv2_fprop(d, _lib.ptr(x.buf), _lib.ptr(om), om.shape[3], _lib.ptr(wpk),
                                               _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(o.buf), _stream(x.buf)),
                   "cnb_dcnv2_fprop")
        if prof:
            prof.end(t0, "dcn", 2.0 * x.B * x.H * x.W * Co * 9 * x.C)
    return o.buf if out is None else o


def heads_fused(x, w3pk, bias3, w1cat, bias1, head_conv, c_out, act):
    """All CenterHead heads in one kernel (cnb_head_fused_fprop): x NHWC bf16 (64 channels) -> [NCHW fp32 map per head],
    or None when the geometry is not covered (the caller then runs the two-GEMM path).  CNB_HEAD_FUSED=0 disables."""
    import ctypes
    import os
    if os.environ.get("CNB_HEAD_FUSED", "1") == "0":
        return None
    x = as_view(x)
    d = _lib.HeadDesc()
    d.B, d.H, d.W, d.Ci = x.B, x.H, x.W, x.C
    d.x_cstride, d.x_coffset, d.head_conv, d.nheads = x.cstride, x.coffset, head_conv, len(c_out)
    if len(c_out) > 8:
        return None
    for i, (c, a) in enumerate(zip(c_out, act)):
        d.c_out[i], d.act[i] = c, a
    L = _lib.lib()
    if not L.cnb_head_fused_supported(d):
        return None
    dev = x.buf.device
    ys = [torch.empty((x.B, c, x.H, x.W), dtype=torch.float32, device=dev) for c in c_out]
    yp = (ctypes.c_void_p * len(ys))(*[y.data_ptr() for y in ys])
    bp = (ctypes.c_void_p * len(ys))(*[b.data_ptr() if b is not None else None for b in bias1])
    prof = LaunchProfiler.active
    with torch.cuda.device(dev):
        t0 = prof.begin() if prof else None
        _lib.check(L.cnb_head_fused_fprop(d, _lib.ptr(x.buf), _lib.ptr(w3pk), _lib.ptr(bias3), _lib.ptr(w1cat), bp, yp,
                                          _stream(x.buf)), "cnb_head_fused_fprop")
        if prof:
            prof.end(t0, "conv", 2.0 * x.B * x.H * x.W * (len(c_out) * head_conv * 9 * x.C + head_conv * sum(c_out)))
    return ys


def maxpool2d(x, k, out=None):
    x = as_view(x)
    Ho, Wo = x.H // k, x.W // k
    o = as_view(out) if out is not None else View(
        torch.empty((x.B, Ho, Wo, x.C), dtype=torch.bfloat16, device=x.buf.device), x.C, 0)
    with torch.cuda.device(x.buf.device):
        _lib.check(_lib.lib().cnb_maxpool2d(_lib.ptr(x.buf), _lib.ptr(o.buf), x.B, x.H, x.W, x.C, x.cstride,
                                             x.coffset, o.cstride, o.coffset, k, _stream(x.buf)), "cnb_maxpool2d")
    return o.buf if out is None else o


def maxpool2d_pad(x, k, stride, pad):
    """MaxPool2d(k, stride, pad) on a dense NHWC bf16 tensor (-inf padding)."""
    x = as_view(x)
    assert x.coffset == 0 and x.cstride == x.C, "maxpool2d_pad needs a dense NHWC input"
    Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
    y = torch.empty((x.B, Ho, Wo, x.C), dtype=torch.bfloat16, device=x.buf.device)
    with torch.cuda.device(x.buf.device):
        _lib.check(_lib.lib().cnb_maxpool2d_pad(_lib.ptr(x.buf), _lib.ptr(y), x.B, x.H, x.W, x.C, k, stride, pad,
                                                 _stream(x.buf)), "cnb_maxpool2d_pad")
    return y


def pack_deconv4x4s2_weights(w):
    """ConvTranspose2d(Ci, Co, 4, stride 2, padding 1) weight [Ci,Co,4,4] -> packed weights of the equivalent
    3x3 / stride 1 / pad 1 convolution with 4*Co output channels (channel block = output phase py*2+px):
    out[2y+py, 2x+px] = sum over the 2x2 input window {y-1+py, y+py} x {x-1+px, x+px} of in * W[kh, kw] with
    kh = 3 - 2*dy - py... written out: phase 0 uses kernel rows (3, 1) at input rows (y-1, y); phase 1 uses
    kernel rows (2, 0) at input rows (y, y+1); the same along x."""
    w = w.detach().float()
    Ci, Co = w.shape[0], w.shape[1]
    w3 = torch.zeros(4, Co, Ci, 3, 3, dtype=torch.float32, device=w.device)
    taps = {0: ((0, 3), (1, 1)), 1: ((1, 2), (2, 0))}      # phase -> ((3x3 tap index, transposed-kernel index), ...)
    for py in (0, 1):
        for px in (0, 1):
            for ty, kh in taps[py]:
                for tx, kw in taps[px]:
                    w3[py * 2 + px, :, :, ty, tx] = w[:, :, kh, kw].t()
    return pack_conv_weights(w3.reshape(4 * Co, Ci, 3, 3))


def deconv4x4s2(x, wpk, Co, scale4, shift4, act=1):
    """Dense ConvTranspose2d(4, stride 2, pad 1) + per-channel scale/shift (+ReLU): one 3x3 conv producing the
    four output phases as channel blocks, then a pixel shuffle.  scale4/shift4 are the per-output-channel vectors
    repeated once per phase ([4*Co], built once at pack time: nothing that feeds a kernel prologue is produced
    in-stream)."""
    x = as_view(x)
    assert scale4 is None or scale4.numel() == 4 * Co
    y4 = conv2d(x, wpk, 4 * Co, 3, 1, 1, scale4, shift4, act=act)
    y = torch.empty((x.B, 2 * x.H, 2 * x.W, Co), dtype=torch.bfloat16, device=x.buf.device)
    with torch.cuda.device(x.buf.device):
        _lib.check(_lib.lib().cnb_depth_to_space2(_lib.ptr(y4), _lib.ptr(y), x.B, x.H, x.W, Co, _stream(x.buf)),
                   "cnb_depth_to_space2")
    return y


def relayout_dw_weights(w, f):
    """[C,1,2f,2f] fp32 -> [(2f)^2, C] fp32 for `dw_deconv_up`."""
    w = w.detach().float().contiguous()
    C = w.shape[0]
    wt = torch.empty((4 * f * f, C), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().cnb_dw_deconv_relayout_weights(_lib.ptr(w), _lib.ptr(wt), C, f, _stream(w)),
                   "cnb_dw_deconv_relayout_weights")
    return wt


def dw_deconv_up(x, wt, f, add=None, out=None):
    """Depthwise ConvTranspose2d(2f, stride f, pad f/2) (+ add); x / add / out dense NHWC bf16."""
    x = as_view(x)
    assert x.coffset == 0 and x.cstride == x.C, "dw_deconv_up needs a dense NHWC input"
    y = out if out is not None else torch.empty((x.B, x.H * f, x.W * f, x.C), dtype=torch.bfloat16,
                                                device=x.buf.device)
    if add is not None:
        assert add.shape == y.shape and add.is_contiguous()
    with torch.cuda.device(x.buf.device):
        _lib.check(_lib.lib().cnb_dw_deconv_up(_lib.ptr(x.buf), _lib.ptr(wt), _lib.ptr(add), _lib.ptr(y), x.B, x.H,
                                                x.W, x.C, f, _stream(x.buf)), "cnb_dw_deconv_up")
    return y
