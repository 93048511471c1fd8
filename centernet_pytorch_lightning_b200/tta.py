"""Test-time augmentation around the engine on the GPU: the pieces of `CenterNetDetection.test_step` /
`test_step_end` (centernet_detection.py:139-171, 188-204) that the reference runs as separate torch / numpy passes with a
`.cpu()` synchronisation in the middle (:188)."""
import ctypes

import torch

from . import _lib

MEAN = (0.408, 0.447, 0.470)      # centernet_detection.py:29-30
STD = (0.289, 0.274, 0.278)


def padding_for(height, width, padding=31):
    """centernet_detection.py:143-144: pad to ((size | padding) + 1), split evenly"""
    return ((width | padding) + 1 - width) // 2, ((height | padding) + 1 - height) // 2


def prologue(img, padding=31, mean=MEAN, std=STD, flip=False):
    """img [1,3,H,W] or [3,H,W] fp32 (already resized for scales != 1) -> ([1 or 2,3,Hp,Wp] normalised, zero-padded
    BEFORE normalisation like F.pad + VF.normalize, second copy h-flipped; meta padding [pad_lr, pad_tb])."""
    _lib.require_cuda(img)
    x = img.float().contiguous()
    x = x[0] if x.dim() == 4 else x
    C, H, W = x.shape
    pad_lr, pad_tb = padding_for(H, W, padding)
    out = torch.empty((2 if flip else 1, C, H + 2 * pad_tb, W + 2 * pad_lr), dtype=torch.float32, device=x.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cnb_tta_prologue(_lib.ptr(x), _lib.ptr(out), C, H, W, pad_lr, pad_tb, m, s, int(flip),
                                               _lib.stream_ptr(x.device)), "cnb_tta_prologue")
    return out, [pad_lr, pad_tb]


def flip_merge(t):
    """[2,C,H,W] -> [1,C,H,W] = (t[0:1] + hflip(t[1:2])) / 2   (centernet_detection.py:169-170)"""
    _lib.require_cuda(t)
    t = t.float().contiguous()
    _, C, H, W = t.shape
    out = torch.empty((1, C, H, W), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().cnb_tta_flip_merge(_lib.ptr(t), _lib.ptr(out), C, H, W, _lib.stream_ptr(t.device)),
                   "cnb_tta_flip_merge")
    return out


def ctdet_post(det, padding, scale, num_classes=80, down_ratio=4):
    """det [K,6] (one image) -> (rows [K,5] in image pixels grouped by class, counts [C], offsets [C]) on the device;
    class j+1 of the reference's `class_predictions` is rows[offsets[j]: offsets[j] + counts[j]]
    (centernet_detection.py:191-204)."""
    _lib.require_cuda(det)
    det = det.float().contiguous().view(-1, 6)
    K = det.shape[0]
    rows = torch.empty((K, 5), dtype=torch.float32, device=det.device)
    counts = torch.empty(num_classes, dtype=torch.int32, device=det.device)
    offsets = torch.empty(num_classes, dtype=torch.int32, device=det.device)
    with torch.cuda.device(det.device):
        _lib.check(_lib.lib().cnb_ctdet_post(_lib.ptr(det), _lib.ptr(rows), _lib.ptr(counts), _lib.ptr(offsets), K, num_classes,
                                             float(down_ratio), float(padding[0]), float(padding[1]), float(scale[0]),
                                             float(scale[1]), _lib.stream_ptr(det.device)), "cnb_ctdet_post")
    return rows, counts, offsets
