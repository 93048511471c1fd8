"""`DCN.dcn_v2.DCN` -- modulated deformable convolution v2 with the constructor the reference uses
(pose_dla_dcn.py:441-449, resnet_dcn.py:202-210) and the DCNv2 parameter names
(`weight`, `bias`, `conv_offset_mask.{weight,bias}`), executed by the sm_100a kernels:

    om  = conv3x3(x; conv_offset_mask)                 -> cnb_conv2d_fprop (tcgen05, NHWC fp32 out)
    y   = sum_k W_k * sigmoid(m_k) * bilinear(x, p+k+d_k) -> cnb_dcnv2_fprop (sampler feeds tcgen05 A tile)

Forward only for now (inference path, configs 2/4/5); `requires_grad` training raises.
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
        self._packed = None

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.in_channels * 9)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.zero_()
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def packed(self):
        """(w_main, w_offset) packed bf16, re-packed when the parameters change."""
        key = (self.weight._version, self.conv_offset_mask.weight._version, self.weight.data_ptr())
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, ops.pack_conv_weights(self.weight), ops.pack_conv_weights(self.conv_offset_mask.weight))
        return self._packed[1], self._packed[2]

    def forward_nhwc(self, x, scale=None, shift=None, act=0, out=None):
        """x: NHWC bf16 tensor or ops.View.  scale/shift default to (1, bias); pass folded BN to fuse it."""
        w_main, w_off = self.packed()
        om = ops.conv2d(x, w_off, 27, 3, 1, 1, None, self.conv_offset_mask.bias.detach().float(), act=0, out_mode=2)
        if shift is None:
            shift = self.bias.detach().float()
        return ops.dcnv2(x, om, w_main, self.out_channels, scale, shift, act=act, out=out)

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad) and self.training:
            raise NotImplementedError("DCN backward is not implemented yet in centernet_b200 (inference only)")
        y = self.forward_nhwc(ops.to_nhwc_bf16(x))
        return ops.to_nchw_f32(y)
