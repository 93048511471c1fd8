"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the loss half of the hot path (numpy, float32 with
float64 accumulation noted where used).  Restates CenterNet/utils/losses.py and
utils/decode.py:43-45; pinned against the unmodified reference by oracle/make_golden.py
(fixtures tests/golden/losses.npz).  Tolerance for the CUDA kernels: 1e-4 relative (fp32 sums in a
different order), stated in tests/test_losses_gpu.py.
"""
import numpy as np

F32 = np.float32


def sigmoid_clamped(x, clamp=1e-4):
    """utils/decode.py:43-45."""
    y = (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)
    return np.clip(y, F32(clamp), F32(1 - clamp))


def neg_loss(pred, gt):
    """utils/losses.py:14-39 `_neg_loss` -> (loss, dloss/dpred)."""
    pred = pred.astype(np.float64)
    gt = gt.astype(np.float64)
    pos = gt == 1
    neg = gt < 1
    nw = (1 - gt) ** 4
    pos_loss = np.where(pos, np.log(pred) * (1 - pred) ** 2, 0.0)
    neg_loss_ = np.where(neg, np.log(1 - pred) * pred ** 2 * nw, 0.0)
    num_pos = pos.sum()
    dpos = np.where(pos, (1 - pred) ** 2 / pred - 2 * (1 - pred) * np.log(pred), 0.0)
    dneg = np.where(neg, (2 * pred * np.log(1 - pred) - pred ** 2 / (1 - pred)) * nw, 0.0)
    if num_pos == 0:
        return F32(-neg_loss_.sum()), (-(dneg)).astype(F32)
    return F32(-(pos_loss.sum() + neg_loss_.sum()) / num_pos), (-(dpos + dneg) / num_pos).astype(F32)


def focal_with_logits(logits, gt, clamp=1e-4):
    """FocalLoss()(sigmoid_clamped(x), gt) and its gradient w.r.t. the logits."""
    s = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    p = np.clip(s, clamp, 1 - clamp)
    loss, dp = neg_loss(p, gt)
    inside = (s >= clamp) & (s <= 1 - clamp)
    return loss, (dp.astype(np.float64) * np.where(inside, s * (1 - s), 0.0)).astype(F32)


def _gather(output, ind):
    B, C, H, W = output.shape
    f = output.reshape(B, C, H * W).transpose(0, 2, 1)
    return np.take_along_axis(f, ind[:, :, None].astype(np.int64), axis=1)   # [B,M,C]


def reg_l1(output, mask, ind, target, per_channel=False):
    """utils/losses.py:53-63 (mask [B,M] bool expanded over C) / :81-91 (mask [B,M,C] float).
    -> (loss, dloss/doutput)."""
    output = output.astype(np.float64)
    pred = _gather(output, ind)
    m = mask.astype(np.float64) if per_channel else np.broadcast_to(mask[:, :, None], pred.shape).astype(np.float64)
    diff = pred * m - target.astype(np.float64) * m
    den = m.sum() + 1e-4
    loss = np.abs(diff).sum() / den
    g = np.sign(diff) * m / den
    B, C, H, W = output.shape
    dout = np.zeros((B, C, H * W))
    for b in range(B):
        for c in range(C):
            np.add.at(dout[b, c], ind[b], g[b, :, c])
    return F32(loss), dout.reshape(B, C, H, W).astype(F32)
