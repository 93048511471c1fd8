"""ctypes binding of libcenternet_b200.so (the C ABI declared in include/centernet_b200.h).

There is deliberately NO CPU fallback: if the library is missing or cannot be loaded the product
path raises.  (The CPU oracle lives under oracle/ and is test infrastructure only.)
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcenternet_b200.so")

_lib = None


class CnbError(RuntimeError):
    pass


class ConvDesc(Structure):
    """Mirror of `cnb_conv_desc` (include/centernet_b200.h)."""
    _fields_ = [(n, c_int) for n in (
        "B", "Hi", "Wi", "Ci", "Co", "KH", "KW", "stride", "pad", "dil", "Ho", "Wo",
        "x_cstride", "x_coffset", "y_cstride", "y_coffset", "res_cstride", "res_coffset",
        "act", "out_nchw_f32", "w_kw", "pad_w1")]


class HeadDesc(Structure):
    """Mirror of `cnb_head_desc` (include/centernet_b200.h)."""
    _fields_ = [(n, c_int) for n in ("B", "H", "W", "Ci", "x_cstride", "x_coffset", "head_conv", "nheads")] + \
               [("c_out", c_int * 8), ("act", c_int * 8)]


def _declare(lib):
    P = c_void_p
    sigs = {
        "cnb_version": (c_int, []),
        "cnb_last_error": (c_char_p, []),
        "cnb_launch_count": (c_ulonglong, []),
        "cnb_ctdet_decode_workspace_bytes": (c_size_t, [c_int] * 5),
        "cnb_ctdet_decode": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
        "cnb_ctdet_decode_host": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int]),
        "cnb_multi_pose_decode_workspace_bytes": (c_size_t, [c_int] * 5),
        "cnb_multi_pose_decode": (c_int, [P] * 7 + [c_int] * 5 + [P, c_size_t, P]),
        "cnb_focal_loss_workspace_bytes": (c_size_t, [c_longlong]),
        "cnb_focal_loss_fwd_bwd": (c_int, [P, P, P, P, P, c_longlong, P, c_size_t, P]),
        "cnb_focal_loss_prob_fwd_bwd": (c_int, [P, P, P, P, c_longlong, P, c_size_t, P]),
        "cnb_sigmoid_clamped_fwd": (c_int, [P, P, c_longlong, c_float, c_float, P]),
        "cnb_sigmoid_clamped_bwd": (c_int, [P, P, P, c_longlong, c_float, c_float, P]),
        "cnb_reg_l1_fwd_bwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
        "cnb_conv_packed_weight_bytes": (c_size_t, [c_int] * 4),
        "cnb_conv_pack_weights": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
        "cnb_conv2d_fprop": (c_int, [POINTER(ConvDesc), P, P, P, P, P, P, P]),
        "cnb_head_fused_supported": (c_int, [POINTER(HeadDesc)]),
        "cnb_head_fused_fprop": (c_int, [POINTER(HeadDesc), P, P, P, P, P, P, P]),
        "cnb_dcnv2_fprop": (c_int, [POINTER(ConvDesc), P, P, c_int, P, P, P, P, P]),
        "cnb_maxpool2d": (c_int, [P, P] + [c_int] * 9 + [P]),
        "cnb_maxpool2d_pad": (c_int, [P, P] + [c_int] * 7 + [P]),
        "cnb_depth_to_space2": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
        "cnb_dw_deconv_relayout_weights": (c_int, [P, P, c_int, c_int, P]),
        "cnb_dw_deconv_up": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
        "cnb_nchw_f32_to_nhwc_bf16": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
        "cnb_nhwc_bf16_to_nchw_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        # training path
        "cnb_conv_wgrad_acc_elems": (c_size_t, [c_int] * 4),
        "cnb_conv2d_wgrad": (c_int, [POINTER(ConvDesc), P, P, c_int, c_int, P, P]),
        "cnb_conv_unpack_wgrad": (c_int, [P, P] + [c_int] * 7 + [c_float, P]),
        "cnb_bn_train_fwd": (c_int, [P, P, P, P, P, c_float, c_float, P, c_int, P, P, P, c_longlong, c_int, P]),
        "cnb_scale_shift_act": (c_int, [P, P, P, P, c_int, P, c_longlong, c_int, P]),
        "cnb_bn_train_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int, P, c_longlong, c_int, P]),
        "cnb_channel_sum": (c_int, [P, c_int, c_int, P, P, c_int, c_longlong, c_int, P]),
        "cnb_relu_bwd": (c_int, [P, P, P, c_longlong, P]),
        "cnb_add_bf16": (c_int, [P, P, P, c_longlong, P]),
        "cnb_f32_to_bf16": (c_int, [P, P, P, c_longlong, c_int, c_int, P]),
        "cnb_maxpool2x2_bwd": (c_int, [P, P, P] + [c_int] * 4 + [P]),
        "cnb_dw_deconv_bwd": (c_int, [P, P, P, P, P] + [c_int] * 5 + [P]),
        "cnb_dw_deconv_unpack_wgrad": (c_int, [P, P, c_int, c_int, c_int, P]),
        "cnb_dcnv2_im2col": (c_int, [P, P, c_int, P] + [c_int] * 4 + [P]),
        "cnb_dcnv2_col2im": (c_int, [P, P, c_int, P, P, P] + [c_int] * 4 + [P]),
        "cnb_adam_step": (c_int, [P, P, P, P, c_longlong, c_float, c_float, c_float, c_float, c_int, P, P, c_float, P]),
        "cnb_p2p_signal_bytes": (c_size_t, []),
        "cnb_p2p_next_step": (c_int, [P, P]),
        "cnb_p2p_allreduce": (c_int, [P, P, P, P, P, c_longlong, c_longlong, c_int, c_int, c_int, c_int, P]),
        "cnb_p2p_wait": (c_int, [P, P, c_int, c_int, P]),
        # fp32-strict mode
        "cnb_strict_conv2d_f32": (c_int, [P] * 6 + [c_int] * 10 + [P]),
        "cnb_strict_dcn_im2col_f32": (c_int, [P, P, P] + [c_int] * 4 + [P]),
        "cnb_strict_conv_transpose2d_f32": (c_int, [P] * 6 + [c_int] * 10 + [P]),
        "cnb_strict_maxpool2d_f32": (c_int, [P, P] + [c_int] * 6 + [P]),
        # callers either side of the path
        "cnb_ctdet_encode": (c_int, [P] * 8 + [c_int] * 6 + [P]),
        "cnb_soft_nms": (c_int, [P, P, P, c_int, c_int, c_int, c_double, c_double, c_double, c_int, P]),
        "cnb_tta_prologue": (c_int, [P, P] + [c_int] * 5 + [P, P, c_int, P]),
        "cnb_tta_flip_merge": (c_int, [P, P, c_int, c_int, c_int, P]),
        "cnb_ctdet_post": (c_int, [P, P, P, P, c_int, c_int] + [c_float] * 5 + [P]),
        # decode primitives by name
        "cnb_nms3x3": (c_int, [P, P, c_longlong, c_int, c_int, P]),
        "cnb_topk_rows": (c_int, [P, c_int, c_int, c_int, P, P, P]),
        "cnb_gather_feat": (c_int, [P, P, P] + [c_int] * 5 + [P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return sigs


def lib():
    """The loaded library; raises CnbError when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CnbError(
                f"{LIB_PATH} not found - build it with `python -m centernet_pytorch_lightning_b200.build` "
                "(there is no CPU fallback in the product path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().cnb_last_error().decode("utf-8", "replace")
        raise CnbError(f"{what} failed (status {rc}): {msg}")


def launch_count():
    return int(lib().cnb_launch_count())


# ---- torch glue ------------------------------------------------------------------------------------
_workspaces = {}
_retired = []      # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def workspace(device, nbytes):
    """Scratch tensor of at least nbytes for the CURRENT stream of `device` (one buffer per (device, stream): two
    streams never share scratch, so the entry points stay re-entrant per stream as include/centernet_b200.h says).
    Buffers grow geometrically; an outgrown buffer is retired, never freed, because an engine's captured graph
    (engine.py) may replay kernels that were recorded with its address."""
    import torch

    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired.append(ws)
        ws = torch.empty(max(int(nbytes), 1 << 20, 2 * ws.numel() if ws is not None else 0), dtype=torch.uint8,
                         device=device)
        _workspaces[key] = ws
    return ws


def stream_ptr(device=None):
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CnbError(
                "centernet_b200 ops run on CUDA tensors only (sm_100a kernels; the product path has no "
                "CPU fallback). Move the inputs to a B200 device.")
