"""`multi_pose_decode` with the reference signature (CenterNet/decode/multi_pose.py:7), executed by
`plane_scan_kernel` + `multi_pose_assoc_kernel` through the C ABI (`cnb_multi_pose_decode`)."""
import torch

from .. import _lib
from .ctdet import _f32c


def multi_pose_decode(heat, wh, kps, reg=None, hm_hp=None, hp_offset=None, K=100):
    """-> [B,K,3J+6] = bbox(4) score(1) keypoints(2J) class(1) hm_score(J); decode/multi_pose.py:7-96.

    Unlike the reference this does NOT modify `kps` in place (multi_pose.py:17-18 adds xs/ys into a
    gathered copy, so callers never observed that anyway).  `hm_hp=None` raises NameError exactly as
    the reference does at multi_pose.py:94.
    """
    if hm_hp is None:
        raise NameError("name 'hm_score' is not defined")
    _lib.require_cuda(heat, wh, kps, reg, hm_hp, hp_offset)
    heat, wh, kps, reg, hm_hp, hp_offset = map(_f32c, (heat, wh, kps, reg, hm_hp, hp_offset))
    B, C, H, W = heat.shape
    if C != 1:
        raise ValueError("multi_pose_decode: heat must have one class channel")
    J = kps.shape[1] // 2
    if hm_hp.shape != (B, J, H, W):
        raise ValueError(f"multi_pose_decode: hm_hp must be [B,J,H,W]={B, J, H, W}")
    L = _lib.lib()
    nbytes = L.cnb_multi_pose_decode_workspace_bytes(B, J, H, W, K)
    if nbytes == 0:
        raise _lib.CnbError(f"multi_pose_decode: unsupported shape B={B} J={J} H={H} W={W} K={K}")
    ws = _lib.workspace(heat.device, nbytes)
    out = torch.empty((B, K, 3 * J + 6), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):
        rc = L.cnb_multi_pose_decode(_lib.ptr(heat), _lib.ptr(wh), _lib.ptr(kps), _lib.ptr(reg),
                                     _lib.ptr(hm_hp), _lib.ptr(hp_offset), _lib.ptr(out),
                                     B, J, H, W, K, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr(heat.device))
    _lib.check(rc, "cnb_multi_pose_decode")
    return out
