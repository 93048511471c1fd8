"""Seeded synthetic inputs of the shapes SURVEY.md section 8(d) names (no dataset, no network).

Heat maps are *tie-free by construction* (a per-image random permutation of distinct values)
because ``torch.topk`` leaves the order of equal scores unspecified (SURVEY.md section 7,
hard part 3) and bit-exact index parity is only defined on distinct scores.
"""
import math

import numpy as np


def distinct_uniform_heat(B, C, H, W, seed=1234, lo=0.0, hi=1.0):
    """[B,C,H,W] float32, every image a permutation of N distinct values in (lo, hi)."""
    rng = np.random.default_rng(seed)
    N = C * H * W
    base = (lo + (hi - lo) * (np.arange(1, N + 1, dtype=np.float64) / (N + 1))).astype(np.float32)
    assert len(np.unique(base)) == N, "value grid collapses in float32; narrow the range"
    out = np.empty((B, N), dtype=np.float32)
    for b in range(B):
        out[b] = base[rng.permutation(N)]
    return out.reshape(B, C, H, W)


def gaussian_bump_heat(B, C, H, W, n_obj=64, seed=1234, floor=0.05):
    """COCO-shaped heat: <= n_obj Gaussian bumps per image (peak values distinct, < 1) drawn with
    max-composition on a distinct-valued noise floor <= ``floor`` (cf. utils/gaussian.py:41-58)."""
    rng = np.random.default_rng(seed)
    heat = distinct_uniform_heat(B, C, H, W, seed=seed + 1, lo=0.0, hi=floor)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        peaks = 0.3 + 0.69 * (rng.permutation(n_obj) + 1.0) / (n_obj + 1.0)
        for k in range(n_obj):
            c = int(rng.integers(0, C))
            cy, cx = int(rng.integers(0, H)), int(rng.integers(0, W))
            rad = int(rng.integers(1, 8))
            sigma = (2 * rad + 1) / 6.0
            g = (peaks[k] * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sigma * sigma))).astype(np.float32)
            g[(np.abs(xx - cx) > rad) | (np.abs(yy - cy) > rad)] = 0
            np.maximum(heat[b, c], g, out=heat[b, c])
    return heat


def ctdet_maps(B, C=80, H=128, W=128, seed=1234, kind="uniform"):
    """(heat, wh, reg) for ``ctdet_decode`` -- SURVEY.md section 8(d) config 2."""
    rng = np.random.default_rng(seed + 7)
    heat = (distinct_uniform_heat(B, C, H, W, seed) if kind == "uniform"
            else gaussian_bump_heat(B, C, H, W, seed=seed))
    wh = (30.0 * rng.random((B, 2, H, W))).astype(np.float32)
    reg = rng.random((B, 2, H, W)).astype(np.float32)
    return heat, wh, reg


def multi_pose_maps(B, J=17, H=128, W=128, seed=1234):
    """(heat, wh, kps, reg, hm_hp, hp_offset) for ``multi_pose_decode`` -- config 5."""
    rng = np.random.default_rng(seed + 11)
    heat = distinct_uniform_heat(B, 1, H, W, seed)
    hm_hp = distinct_uniform_heat(B, J, H, W, seed + 3)
    wh = (30.0 * rng.random((B, 2, H, W))).astype(np.float32)
    kps = (3.0 * rng.standard_normal((B, 2 * J, H, W))).astype(np.float32)
    reg = rng.random((B, 2, H, W)).astype(np.float32)
    hp_offset = rng.random((B, 2, H, W)).astype(np.float32)
    return heat, wh, kps, reg, hm_hp, hp_offset


def randomize_(sd, seed=0, offset_gain=0.05, offset_bias=0.3):
    """Seeded 'trained-looking' parameters in place: He-normal convs (variance preserving through the
    34 layers), non-trivial BN statistics (so folding is exercised) and non-zero DCN offset/mask weights
    (zero init would reduce DCN to 0.5 * conv).  Deterministic on CPU.

    DCN offsets are a discontinuous-gradient function of the features: with large random offset weights a
    bf16 network drifts from the fp32 one by >30 % through the 16 stacked DCNs (measured with a CPU
    emulation of per-layer bf16 rounding), with offset_gain=0 by 0.4 %.  offset_gain (std of the offset conv
    weights in units of 1/sqrt(fan_in)) defaults to a trained-like small value; offset_bias gives every tap
    a fixed fractional displacement so the bilinear sampler is fully exercised either way."""
    import torch

    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if k.endswith("num_batches_tracked"):
            continue
        if ".up_" in k:
            continue                                             # bilinear up-sampling kernels stay as initialised
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif "conv_offset_mask.weight" in k:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * (offset_gain / math.sqrt(fan_in)))
        elif "conv_offset_mask.bias" in k:
            v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * offset_bias)
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * math.sqrt(2.0 / fan_in))
        elif k.endswith(".weight"):                              # BN gamma
            v.copy_(0.7 + 0.6 * torch.rand(v.shape, generator=g))
        elif k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def ctdet_targets(B, C=80, H=128, W=128, n_obj=32, max_objs=128, seed=1234):
    """COCO-shaped training targets of SURVEY.md 8(d) config 3 as torch CPU tensors: `heatmap` [B,C,H,W] (Gaussian
    bumps composed by max, exactly 1 at object centres), `indices` [B,max_objs] int64, `regression_mask` bool,
    `width_height`, `regression` [B,max_objs,2] -- the dict `CenterNetDetection.loss` consumes
    (centernet_detection.py:97-130; encoded on the CPU by sample/ctdet.py:39-90 in the reference)."""
    import torch

    rng = np.random.default_rng(seed)
    heat = np.zeros((B, C, H, W), np.float32)
    ind = np.zeros((B, max_objs), np.int64)
    mask = np.zeros((B, max_objs), bool)
    wh = np.zeros((B, max_objs, 2), np.float32)
    reg = np.zeros((B, max_objs, 2), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        used = set()
        for k in range(min(n_obj, max_objs)):
            c = int(rng.integers(0, C))
            cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
            if (c, cy, cx) in used:
                continue
            used.add((c, cy, cx))
            w, h = 2 + 28 * rng.random(), 2 + 28 * rng.random()
            rad = max(1, int(min(w, h) / 3))
            sigma = (2 * rad + 1) / 6.0
            g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sigma * sigma)).astype(np.float32)
            g[(np.abs(xx - cx) > rad) | (np.abs(yy - cy) > rad)] = 0
            np.maximum(heat[b, c], g, out=heat[b, c])
            heat[b, c, cy, cx] = 1.0
            ind[b, k], mask[b, k] = cy * W + cx, True
            wh[b, k] = (w, h)
            reg[b, k] = rng.random(2)
    return {"heatmap": torch.from_numpy(heat), "indices": torch.from_numpy(ind), "regression_mask": torch.from_numpy(mask),
            "width_height": torch.from_numpy(wh), "regression": torch.from_numpy(reg)}
