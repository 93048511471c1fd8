"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the ctdet target encoding
(CenterNet/sample/ctdet.py:39-90, utils/gaussian.py:6-58).  Used to re-run the reference's only
known-answer test for the hot path (tests/test_sample_encode_decode.py:14-56) without the reference.
"""
import math

import numpy as np

F32 = np.float32


def gaussian_radius(det_size, min_overlap=0.7):
    """utils/gaussian.py:6-27."""
    height, width = det_size
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + math.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + math.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + math.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def gaussian2d(diameter, sigma):
    """utils/gaussian.py:30-38 (float32 like torch.arange/exp defaults)."""
    m = (diameter - 1.0) / 2.0
    ax = np.arange(-m, m + 1, dtype=F32)
    h = np.exp(-(ax[None, :] * ax[None, :] + ax[:, None] * ax[:, None]) / F32(2 * sigma * sigma)).astype(F32)
    h[h < np.finfo(F32).eps * h.max()] = 0
    return h


def draw_umich_gaussian(heatmap, center, radius):
    """utils/gaussian.py:41-58."""
    diameter = 2 * radius + 1
    g = gaussian2d(diameter, diameter / 6)
    x, y = int(center[0]), int(center[1])
    H, W = heatmap.shape
    left, right = min(x, radius), min(W - x, radius + 1)
    top, bottom = min(y, radius), min(H - y, radius + 1)
    mh = heatmap[y - top:y + bottom, x - left:x + right]
    mg = g[radius - top:radius + bottom, radius - left:radius + right]
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg, out=mh)
    return heatmap


def encode_ctdet(boxes_xywh, class_ids, input_hw=(512, 512), down_ratio=4, num_classes=80, max_objects=128):
    """sample/ctdet.py:39-90 -> dict(heatmap, regression_mask, indices, width_height, regression)."""
    oh, ow = input_hw[0] // down_ratio, input_hw[1] // down_ratio
    heat = np.zeros((num_classes, oh, ow), F32)
    wh = np.zeros((max_objects, 2), F32)
    reg = np.zeros((max_objects, 2), F32)
    mask = np.zeros(max_objects, bool)
    ind = np.zeros(max_objects, np.int64)
    for k, (box, cls) in enumerate(zip(boxes_xywh[:max_objects], class_ids)):
        bb = np.array([box[0], box[1], box[0] + box[2], box[1] + box[3]], F32) / F32(down_ratio)
        bb[0::2] = np.clip(bb[0::2], 0, ow - 1)
        bb[1::2] = np.clip(bb[1::2], 0, oh - 1)
        h, w = bb[3] - bb[1], bb[2] - bb[0]
        if h > 0 and w > 0:
            radius = max(0, int(gaussian_radius((math.ceil(h), math.ceil(w)))))
            ct = np.array([(bb[0] + bb[2]) / 2, (bb[1] + bb[3]) / 2], F32)
            ct_int = ct.astype(np.int32)
            draw_umich_gaussian(heat[cls], ct_int, radius)
            wh[k] = (w, h)
            ind[k] = ct_int[1] * ow + ct_int[0]
            reg[k] = ct - ct_int
            mask[k] = True
    return dict(heatmap=heat, regression_mask=mask, indices=ind, width_height=wh, regression=reg)
