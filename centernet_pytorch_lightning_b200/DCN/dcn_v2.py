"""`DCN.dcn_v2.DCN` -- modulated deformable convolution v2 with the constructor the reference uses
(pose_dla_dcn.py:441-449, resnet_dcn.py:202-210) and the DCNv2 parameter names
(`weight`, `bias`, `conv_offset_mask.{weight,bias}`), executed by the sm_100a kernels:

    om  = conv3x3(x; conv_offset_mask)                 -> cnb_conv2d_fprop (tcgen05, NHWC fp32 out)
    y   = sum_k W_k * sigmoid(m_k) * bilinear(x, p+k+d_k) -> cnb_dcnv2_fprop (sampler feeds tcgen05 A tile)

Training: `autograd_ops.dcn` (cnb_dcnv2_im2col + 1x1 GEMMs + cnb_dcnv2_col2im + cnb_conv2d_wgrad).
"""
import math

import torch
from torch import nn

from .. import ops


class DCN(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1,
                 deformable_groups=1):
        super().__init__()
        kh, kw = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if (kh, kw, stride, padding, dilation, deformable_groups) != (3, 3, 1, 1, 1, 1):
            raise NotImplementedError("centernet_b200 DCN implements the reference's only configuration: "
                                      "3x3, stride 1, padding 1, dilation 1, deformable_groups 1")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * kh * kw, kernel_size=(kh, kw),
                                          stride=stride, padding=padding, bias=True)
        self.reset_parameters()
        self._packed = ops.PackCache()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.in_channels * 9)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.zero_()
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def packed(self):
        """(w_main, w_offset, b_offset) packed bf16 / fp32, rebuilt (in place) when the parameters change."""
        com = self.conv_offset_mask
        return self._packed.get("w", (self.weight, com.weight, com.bias),
                                lambda: (ops.pack_conv_weights(self.weight), ops.pack_conv_weights(com.weight),
                                         com.bias.detach().float().contiguous()))

    def forward_nhwc(self, x, scale=None, shift=None, act=0, out=None):
        """x: NHWC bf16 tensor or ops.View.  scale/shift default to (1, bias); pass folded BN to fuse it."""
        w_main, w_off, b_off = self.packed()
        om = ops.conv2d(x, w_off, 27, 3, 1, 1, None, b_off, act=0, out_mode=2)
        if shift is None:
            shift = self.bias.detach().float()
        return ops.dcnv2(x, om, w_main, self.out_channels, scale, shift, act=act, out=out)

    def forward(self, x):
        """[B,C,H,W] fp32 -> [B,Co,H,W] fp32, the stand-alone module call of the DCNv2 API.  In training mode the op
        runs on the autograd tape (autograd_ops.dcn: sampled columns + tensor-core GEMMs, backward included)."""
        if self.training and torch.is_grad_enabled():
            from .. import autograd_ops as ag
            return ag.to_nchw_f32(ag.dcn(ag.to_nhwc_bf16(x), self))
        y = self.forward_nhwc(ops.to_nhwc_bf16(x))
        return ops.to_nchw_f32(y)
