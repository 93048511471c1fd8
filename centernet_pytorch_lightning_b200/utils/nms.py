"""`soft_nms` / `soft_nms_39` (CenterNet/utils/nms.py:5-206, numba-JIT on the CPU in the reference) on the GPU:
`cnb_soft_nms` keeps the sequential semantics of the original (selection order, swap-with-last compaction, float64 overlap
arithmetic), one CTA per box list -- all classes of an image are processed by one launch (`soft_nms_lists`)."""
import numpy as np
import torch

from .. import _lib


def soft_nms_lists(boxes, counts, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """boxes [L,N,ncol] fp32 CUDA tensor (in place), counts [L] int32 -> kept [L] int32: after the call the first kept[l]
    rows of list l are the surviving boxes in selection order with decayed scores."""
    _lib.require_cuda(boxes, counts)
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.dim() == 3
    L, N, ncol = boxes.shape
    kept = torch.empty(L, dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.lib().cnb_soft_nms(_lib.ptr(boxes), _lib.ptr(counts.to(torch.int32).contiguous()), _lib.ptr(kept), L, N,
                                           ncol, float(sigma), float(Nt), float(threshold), int(method),
                                           _lib.stream_ptr(boxes.device)), "cnb_soft_nms")
    return kept


def _one(boxes, sigma, Nt, threshold, method, device):
    was_numpy = isinstance(boxes, np.ndarray)
    t = torch.as_tensor(boxes, dtype=torch.float32)
    if t.shape[0] == 0:
        return []
    dev = torch.device(device) if device is not None else (t.device if t.is_cuda else torch.device("cuda"))
    g = t.to(dev).contiguous().unsqueeze(0).clone()
    kept = soft_nms_lists(g, torch.tensor([t.shape[0]], dtype=torch.int32, device=dev), sigma, Nt, threshold, method)
    res = g[0].to(t.device)
    if was_numpy:
        boxes[...] = res.numpy().astype(boxes.dtype)      # the reference works in place on the caller's array
    else:
        boxes.copy_(res)
    return list(range(int(kept.item())))


def soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0, device=None):
    """boxes [N,5] = (x1,y1,x2,y2,score), modified in place like the reference; returns the keep indices."""
    return _one(boxes, sigma, Nt, threshold, method, device)


def soft_nms_39(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0, device=None):
    """the 39-column variant (box, score, 17 key points): rows are carried along whole."""
    return _one(boxes, sigma, Nt, threshold, method, device)
