"""Drop-in for CenterNet/utils/losses.py: FocalLoss (:42-50, `_neg_loss` :14-39), RegL1Loss (:53-63),
RegWeightedL1Loss (:81-91) -- each a single fused forward+backward CUDA pass (csrc/losses.cu), no
host sync (the reference syncs at `if num_pos == 0`, losses.py:35) and no NCHW->NHWC gather copy
(utils/decode.py:59-63).
"""
import torch
from torch import nn

from .. import _lib


def _loss_workspace(device):
    L = _lib.lib()
    return _lib.workspace(device, L.cnb_focal_loss_workspace_bytes(0))


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        _lib.require_cuda(pred, gt)
        pred = pred.contiguous()
        gt = gt.contiguous().float()
        L = _lib.lib()
        stats = torch.empty(3, dtype=torch.float32, device=pred.device)
        graw = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        ws = _loss_workspace(pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(L.cnb_focal_loss_prob_fwd_bwd(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(graw),
                                                     _lib.ptr(stats), pred.numel(), _lib.ptr(ws),
                                                     ws.numel(), _lib.stream_ptr(pred.device)),
                       "cnb_focal_loss_prob_fwd_bwd")
        ctx.save_for_backward(graw, stats)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        graw, stats = ctx.saved_tensors
        return graw * (grad_out * stats[1]), None


class _FocalLogitsFn(torch.autograd.Function):
    """sigmoid_clamped + _neg_loss fused: one pass over (logits, gt) for loss AND dloss/dlogits."""

    @staticmethod
    def forward(ctx, logits, gt):
        _lib.require_cuda(logits, gt)
        logits = logits.contiguous()
        gt = gt.contiguous().float()
        L = _lib.lib()
        stats = torch.empty(3, dtype=torch.float32, device=logits.device)
        graw = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        ws = _loss_workspace(logits.device)
        with torch.cuda.device(logits.device):
            _lib.check(L.cnb_focal_loss_fwd_bwd(_lib.ptr(logits), _lib.ptr(gt), None, _lib.ptr(graw),
                                                _lib.ptr(stats), logits.numel(), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(logits.device)),
                       "cnb_focal_loss_fwd_bwd")
        ctx.save_for_backward(graw, stats)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        graw, stats = ctx.saved_tensors
        return graw * (grad_out * stats[1]), None


def focal_loss_with_logits(logits, gt):
    """== FocalLoss()(sigmoid_clamped(logits), gt) in one HBM pass (12 B/element fwd+bwd)."""
    return _FocalLogitsFn.apply(_f32(logits), gt)


def _f32(t):
    """The kernels read fp32: an fp16 / bf16 prediction (AMP) is cast here, outside the autograd Function, so the
    gradient is cast back to the input dtype by autograd itself."""
    return t if t.dtype == torch.float32 else t.float()


class FocalLoss(nn.Module):
    """`FocalLoss()(pred, gt)`; pred is the sigmoid-clamped heat map (losses.py:42-50)."""

    def forward(self, out, target):
        return _FocalFn.apply(_f32(out), target)


class _RegL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, mask, ind, target, per_channel):
        _lib.require_cuda(output, mask, ind, target)
        output = output.contiguous()
        B, C, H, W = output.shape
        M = ind.shape[1]
        ind = ind.contiguous().long()
        target = target.contiguous().float()
        mask = mask.contiguous().float() if per_channel else mask.contiguous().to(torch.uint8)
        loss = torch.empty(1, dtype=torch.float32, device=output.device)
        L = _lib.lib()
        with torch.cuda.device(output.device):
            _lib.check(L.cnb_reg_l1_fwd_bwd(_lib.ptr(output), _lib.ptr(mask), _lib.ptr(ind), _lib.ptr(target),
                                            None, _lib.ptr(loss), B, C, H, W, M, int(per_channel), None,
                                            _lib.stream_ptr(output.device)), "cnb_reg_l1_fwd_bwd")
        ctx.save_for_backward(output, mask, ind, target)
        ctx.per_channel = per_channel
        return loss[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        output, mask, ind, target = ctx.saved_tensors
        B, C, H, W = output.shape
        M = ind.shape[1]
        dout = torch.empty_like(output)
        loss = torch.empty(1, dtype=torch.float32, device=output.device)
        gs = grad_out.reshape(1).float().contiguous()
        L = _lib.lib()
        with torch.cuda.device(output.device):
            _lib.check(L.cnb_reg_l1_fwd_bwd(_lib.ptr(output), _lib.ptr(mask), _lib.ptr(ind), _lib.ptr(target),
                                            _lib.ptr(dout), _lib.ptr(loss), B, C, H, W, M,
                                            int(ctx.per_channel), _lib.ptr(gs),
                                            _lib.stream_ptr(output.device)), "cnb_reg_l1_fwd_bwd")
        return dout, None, None, None, None


class RegL1Loss(nn.Module):
    """losses.py:53-63: mask [B,M] bool expanded over channels."""

    def forward(self, output, mask, ind, target):
        return _RegL1Fn.apply(_f32(output), mask, ind, target, False)


class RegWeightedL1Loss(nn.Module):
    """losses.py:81-91: mask [B,M,C] float weights."""

    def forward(self, output, mask, ind, target):
        return _RegL1Fn.apply(_f32(output), mask, ind, target, True)
