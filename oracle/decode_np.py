"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the decode half of the hot path.

numpy (float32 / int64) restatement of the reference's decode algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module; the product package never does (it fails loudly without its CUDA library).

Pinned (oracle/make_golden.py, tests/test_oracle_golden.py) against
  * the reference's own known-answer test tests/test_sample_encode_decode.py:14-56
    (fixture tests/data/coco_annotation.json), and
  * outputs of the unmodified reference functions executed in the authoring container
    (fixtures under tests/golden/, generator script committed).

Tie rule.  ``torch.topk`` leaves the order of equal scores unspecified; this oracle (and the
CUDA kernels) define it as (score descending, flat index ascending).  With distinct scores the
result equals the reference's bit for bit; parity tests compare rows whose score is unique.
"""
import numpy as np

F32 = np.float32


def nms(heat: np.ndarray) -> np.ndarray:
    """utils/decode.py:5-10  ``_nms``: keep = (max_pool2d(3,1,1) == heat); heat * keep."""
    B, C, H, W = heat.shape
    pad = np.full((B, C, H + 2, W + 2), -np.inf, dtype=F32)
    pad[:, :, 1:-1, 1:-1] = heat
    hmax = pad[:, :, 1:-1, 1:-1].copy()
    for dy in range(3):
        for dx in range(3):
            np.maximum(hmax, pad[:, :, dy:dy + H, dx:dx + W], out=hmax)
    keep = (hmax == heat).astype(F32)
    return heat * keep


def _topk_rows(x: np.ndarray, K: int):
    """Row-wise top-K, (value desc, index asc).  x: [..., N] -> values, indices [..., K]."""
    order = np.argsort(-x, axis=-1, kind="stable")[..., :K]
    return np.take_along_axis(x, order, axis=-1), order.astype(np.int64)


def topk_channel(scores: np.ndarray, K: int):
    """utils/decode.py:31-40 ``_topk_channel``."""
    B, C, H, W = scores.shape
    vals, inds = _topk_rows(scores.reshape(B, C, H * W), K)
    inds = inds % (H * W)
    # (inds / width).int().float(): exact for H*W < 2**24 (float32 true division, then trunc)
    ys = (inds // W).astype(F32)
    xs = (inds % W).astype(F32)
    return vals, inds, ys, xs


def topk(scores: np.ndarray, K: int):
    """utils/decode.py:13-28 ``_topk``: per-class top-K, then top-K over the C*K survivors."""
    B, C, H, W = scores.shape
    vals, inds, ys, xs = topk_channel(scores, K)
    score, ind = _topk_rows(vals.reshape(B, C * K), K)
    clses = (ind // K).astype(np.int32)
    inds = np.take_along_axis(inds.reshape(B, C * K), ind, axis=1)
    ys = np.take_along_axis(ys.reshape(B, C * K), ind, axis=1)
    xs = np.take_along_axis(xs.reshape(B, C * K), ind, axis=1)
    return score, inds, clses, ys, xs


def transpose_and_gather(feat: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """utils/decode.py:59-63: NCHW -> [B, HW, C] rows gathered at ind [B, M] -> [B, M, C]."""
    B, C, H, W = feat.shape
    f = feat.reshape(B, C, H * W).transpose(0, 2, 1)
    return np.take_along_axis(f, ind[:, :, None].astype(np.int64), axis=1)


def ctdet_decode(heat, wh, reg=None, K=100):
    """decode/ctdet.py:6-38.  heat [B,C,H,W] (already sigmoided), wh/reg [B,2,H,W] -> [B,K,6]."""
    heat = np.ascontiguousarray(heat, dtype=F32)
    B = heat.shape[0]
    heat = nms(heat)
    scores, inds, clses, ys, xs = topk(heat, K)
    if reg is not None:
        r = transpose_and_gather(np.asarray(reg, F32), inds)
        xs = xs + r[:, :, 0]
        ys = ys + r[:, :, 1]
    else:
        xs = xs + F32(0.5)
        ys = ys + F32(0.5)
    w = transpose_and_gather(np.asarray(wh, F32), inds)
    hw, hh = w[:, :, 0] / F32(2), w[:, :, 1] / F32(2)
    det = np.stack([xs - hw, ys - hh, xs + hw, ys + hh, scores, clses.astype(F32)], axis=2)
    return det.astype(F32).reshape(B, K, 6)


def multi_pose_decode(heat, wh, kps, reg=None, hm_hp=None, hp_offset=None, K=100):
    """decode/multi_pose.py:7-96.  Output [B,K,57] = bbox(4) score(1) kps(34) cls(1) hm_score(17).

    Reproduces the reference's quirks: in-place ``kps[...,::2] += xs`` before the reg offset
    (:17-18), the sentinel arithmetic (:58-61), ``dist.min`` first-index ties (:68), and the
    ``hm_score.view(batch, K, num_joints)`` of a [B,J,K,1] tensor *without* permute (:90).
    """
    heat = np.ascontiguousarray(heat, dtype=F32)
    B = heat.shape[0]
    J = kps.shape[1] // 2
    heat = nms(heat)
    scores, inds, clses, ys, xs = topk(heat, K)
    k = transpose_and_gather(np.asarray(kps, F32), inds).copy()          # [B,K,2J]
    k[..., 0::2] += xs[:, :, None]
    k[..., 1::2] += ys[:, :, None]
    if reg is not None:
        r = transpose_and_gather(np.asarray(reg, F32), inds)
        xs = xs + r[:, :, 0]
        ys = ys + r[:, :, 1]
    else:
        xs = xs + F32(0.5)
        ys = ys + F32(0.5)
    w = transpose_and_gather(np.asarray(wh, F32), inds)
    hw, hh = w[:, :, 0] / F32(2), w[:, :, 1] / F32(2)
    bboxes = np.stack([xs - hw, ys - hh, xs + hw, ys + hh], axis=2).astype(F32)   # [B,K,4]
    if hm_hp is None:
        raise NameError("hm_score")  # the reference raises NameError at multi_pose.py:94
    thresh = F32(0.1)
    hm = nms(np.ascontiguousarray(hm_hp, dtype=F32))
    kp = k.reshape(B, K, J, 2).transpose(0, 2, 1, 3)                       # [B,J,K,2]
    hm_score, hm_inds, hm_ys, hm_xs = topk_channel(hm, K)                  # [B,J,K]
    if hp_offset is not None:
        o = transpose_and_gather(np.asarray(hp_offset, F32), hm_inds.reshape(B, -1))
        o = o.reshape(B, J, K, 2)
        hm_xs = hm_xs + o[..., 0]
        hm_ys = hm_ys + o[..., 1]
    else:
        hm_xs = hm_xs + F32(0.5)
        hm_ys = hm_ys + F32(0.5)
    mask = (hm_score > thresh).astype(F32)
    one = F32(1)
    hm_score = (one - mask) * F32(-1) + mask * hm_score
    hm_ys = (one - mask) * F32(-10000) + mask * hm_ys
    hm_xs = (one - mask) * F32(-10000) + mask * hm_xs
    # dist[b,j,k,m] between regressed joint of detection k and heat-map candidate m
    dx = kp[:, :, :, None, 0] - hm_xs[:, :, None, :]
    dy = kp[:, :, :, None, 1] - hm_ys[:, :, None, :]
    dist = np.sqrt(dx * dx + dy * dy).astype(F32)
    min_ind = np.argmin(dist, axis=3)                                      # first index on ties
    min_dist = np.take_along_axis(dist, min_ind[..., None], axis=3)[..., 0]
    sc = np.take_along_axis(hm_score, min_ind, axis=2)                     # [B,J,K]
    hx = np.take_along_axis(hm_xs, min_ind, axis=2)
    hy = np.take_along_axis(hm_ys, min_ind, axis=2)
    l = bboxes[:, None, :, 0]
    t = bboxes[:, None, :, 1]
    r_ = bboxes[:, None, :, 2]
    b_ = bboxes[:, None, :, 3]
    gate = ((hx < l) | (hx > r_) | (hy < t) | (hy > b_) | (sc < thresh)
            | (min_dist > np.maximum(b_ - t, r_ - l) * F32(0.3)))
    m = gate.astype(F32)
    sc = sc * (one - m)                                                    # [B,J,K]
    hm_score_out = sc.reshape(B, K, J)                                     # view w/o permute (:90)
    kx = (one - m) * hx + m * kp[..., 0]
    ky = (one - m) * hy + m * kp[..., 1]
    kout = np.stack([kx, ky], axis=-1).transpose(0, 2, 1, 3).reshape(B, K, 2 * J)
    det = np.concatenate([bboxes, scores[:, :, None], kout, clses.astype(F32)[:, :, None],
                          hm_score_out], axis=2)
    return det.astype(F32)
