"""Drop-in for CenterNet/utils/decode.py.

The reference builds decode out of `_nms`, `_topk`, `_topk_channel`, `_gather_feat` and
`_transpose_and_gather_feat` (utils/decode.py:5-63).  Here those stages are fused inside the decode
kernels (csrc/decode.cu); the only helper call sites outside decode need is `sigmoid_clamped`
(centernet_detection.py:104, centernet_multi_pose.py:104,110) which maps to
`cnb_sigmoid_clamped_{fwd,bwd}`.
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
