"""`ctdet_decode` with the reference signature (CenterNet/decode/ctdet.py:6), executed by the fused
sm_100a kernel `plane_scan_kernel<.., FUSE_CTDET>` through the C ABI (`cnb_ctdet_decode`)."""
import torch

from .. import _lib


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ctdet_decode(heat, wh, reg=None, K=100):
    """heat [B,C,H,W] (already sigmoided), wh [B,2,H,W], reg [B,2,H,W] or None -> detections [B,K,6]
    = (x1, y1, x2, y2, score, class) in output-stride units, like decode/ctdet.py:6-38.

    Ties between equal scores are ordered by flat (class, y, x) index; the reference leaves them to
    torch.topk.  One kernel launch; asynchronous on the current CUDA stream.
    """
    _lib.require_cuda(heat, wh, reg)
    heat, wh, reg = _f32c(heat), _f32c(wh), _f32c(reg)
    B, C, H, W = heat.shape
    if wh.shape != (B, 2, H, W) or (reg is not None and reg.shape != (B, 2, H, W)):
        raise ValueError(f"ctdet_decode: wh/reg must be [B,2,H,W]={B, 2, H, W}")
    L = _lib.lib()
    nbytes = L.cnb_ctdet_decode_workspace_bytes(B, C, H, W, K)
    if nbytes == 0:
        raise _lib.CnbError(f"ctdet_decode: unsupported shape B={B} C={C} H={H} W={W} K={K}")
    ws = _lib.workspace(heat.device, nbytes)
    out = torch.empty((B, K, 6), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):
        rc = L.cnb_ctdet_decode(_lib.ptr(heat), _lib.ptr(wh), _lib.ptr(reg), _lib.ptr(out),
                                B, C, H, W, K, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(heat.device))
    _lib.check(rc, "cnb_ctdet_decode")
    return out
