"""Drop-in for CenterNet/utils/decode.py.

The reference builds decode out of `_nms`, `_topk`, `_topk_channel`, `_gather_feat` and
`_transpose_and_gather_feat` (utils/decode.py:5-63).  Here those stages are fused inside the decode
kernels (csrc/decode.cu); the helper the task classes call directly is `sigmoid_clamped`
(centernet_detection.py:104, centernet_multi_pose.py:104,110) -> `cnb_sigmoid_clamped_{fwd,bwd}`.  The five
primitives are also importable by name (stand-alone kernels, csrc/decode_helpers.cu) at the end of this file.
"""
import torch

from .. import _lib


class _SigmoidClamped(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, clamp):
        _lib.require_cuda(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            _lib.check(L.cnb_sigmoid_clamped_fwd(_lib.ptr(xc), _lib.ptr(y), xc.numel(), clamp, 1 - clamp,
                                                 _lib.stream_ptr(x.device)), "cnb_sigmoid_clamped_fwd")
        ctx.clamp = clamp
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        L = _lib.lib()
        with torch.cuda.device(y.device):
            _lib.check(L.cnb_sigmoid_clamped_bwd(_lib.ptr(y), _lib.ptr(dy), _lib.ptr(dx), y.numel(),
                                                 ctx.clamp, 1 - ctx.clamp, _lib.stream_ptr(y.device)),
                       "cnb_sigmoid_clamped_bwd")
        return dx, None


def sigmoid_clamped(x, clamp=1e-4):
    """clamp(sigmoid(x), clamp, 1-clamp) -- utils/decode.py:43-45 (out of place; differentiable)."""
    if x.dtype != torch.float32:
        x = x.float()
    return _SigmoidClamped.apply(x, float(clamp))


# ---- the decode primitives by name (utils/decode.py:5-63) --------------------------------------------------------
# ctdet_decode / multi_pose_decode fuse all of these into one streaming pass (csrc/decode.cu) and do not call them;
# they are provided so that `from CenterNet.utils.decode import _nms, _topk, ...` keeps working after the module swap
# of INTEGRATION.md.  Same signatures, return shapes and dtypes as the reference; equal scores are ordered by ascending
# index (the reference leaves ties to torch.topk).
def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args), name)


def _nms(heat, kernel=3):
    """keep = (max_pool2d(heat, 3, 1, 1) == heat); return heat * keep   (utils/decode.py:5-10)."""
    if kernel != 3:
        raise NotImplementedError("centernet_b200 _nms: 3x3 only (the reference's only use)")
    _lib.require_cuda(heat)
    h = heat.float().contiguous()
    out = torch.empty_like(h)
    with torch.cuda.device(h.device):
        _call("cnb_nms3x3", _lib.ptr(h), _lib.ptr(out), h.numel() // (h.shape[-1] * h.shape[-2]), h.shape[-2],
              h.shape[-1], _lib.stream_ptr(h.device))
    return out


def _topk_rows(scores2d, K):
    rows, n = scores2d.shape
    vals = torch.empty((rows, K), dtype=torch.float32, device=scores2d.device)
    idx = torch.empty((rows, K), dtype=torch.int64, device=scores2d.device)
    with torch.cuda.device(scores2d.device):
        _call("cnb_topk_rows", _lib.ptr(scores2d), rows, n, K, _lib.ptr(vals), _lib.ptr(idx),
              _lib.stream_ptr(scores2d.device))
    return vals, idx


def _gather_feat(feat, ind, mask=None):
    """feat [B,N,C], ind [B,K] -> [B,K,C]   (utils/decode.py:48-56)."""
    _lib.require_cuda(feat, ind)
    B, N, C = feat.shape
    K = ind.shape[1]
    dtype = feat.dtype
    f = feat.float().contiguous()
    out = torch.empty((B, K, C), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _call("cnb_gather_feat", _lib.ptr(f), _lib.ptr(ind.long().contiguous()), _lib.ptr(out), B, C, N, K, 0,
              _lib.stream_ptr(feat.device))
    out = out.to(dtype) if dtype != torch.float32 else out
    if mask is not None:
        mask = mask.unsqueeze(2).expand_as(out)
        out = out[mask].view(-1, C)
    return out


def _transpose_and_gather_feat(feat, ind):
    """feat [B,C,H,W], ind [B,K] -> [B,K,C] without the NCHW -> NHWC copy of utils/decode.py:59-63."""
    _lib.require_cuda(feat, ind)
    B, C, H, W = feat.shape
    K = ind.shape[1]
    f = feat.float().contiguous()
    out = torch.empty((B, K, C), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _call("cnb_gather_feat", _lib.ptr(f), _lib.ptr(ind.long().contiguous()), _lib.ptr(out), B, C, H * W, K, 1,
              _lib.stream_ptr(feat.device))
    return out


def _topk_channel(scores, K=40):
    """per-channel top-K (utils/decode.py:31-40) -> scores, inds, ys, xs, each [B,C,K]."""
    _lib.require_cuda(scores)
    B, C, H, W = scores.shape
    vals, inds = _topk_rows(scores.float().contiguous().view(B * C, H * W), K)
    vals, inds = vals.view(B, C, K), inds.view(B, C, K)
    inds = inds % (H * W)
    ys = (inds / W).int().float()
    xs = (inds % W).int().float()
    return vals, inds, ys, xs


def _topk(scores, K=40):
    """utils/decode.py:13-28 -> (score [B,K], inds [B,K] int64, clses [B,K] int32, ys, xs [B,K] float)."""
    B, C, H, W = scores.shape
    topk_scores, topk_inds, topk_ys, topk_xs = _topk_channel(scores, K)
    topk_score, topk_ind = _topk_rows(topk_scores.reshape(B, C * K).contiguous(), K)
    topk_clses = (topk_ind / K).int()
    topk_inds = _gather_feat(topk_inds.view(B, -1, 1).float(), topk_ind).view(B, K).long()
    topk_ys = _gather_feat(topk_ys.view(B, -1, 1), topk_ind).view(B, K)
    topk_xs = _gather_feat(topk_xs.view(B, -1, 1), topk_ind).view(B, K)
    return topk_score, topk_inds, topk_clses, topk_ys, topk_xs
