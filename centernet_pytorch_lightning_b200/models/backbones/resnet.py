"""ResNet + up-sampling backbones behind the reference's plugin contract:

* `resdcn_N`  -- CenterNet/models/backbones/resnet_dcn.py:131-277 (`PoseResNet` with 3 x [DCN 3x3 -> BN -> ReLU ->
  ConvTranspose2d(4, stride 2) -> BN -> ReLU], channels 256/128/64, `out_channels = 64`),
* `res_N`     -- CenterNet/models/backbones/msra_resnet.py:103-260 (3 x [ConvTranspose2d(4, stride 2) -> BN -> ReLU],
  channels 256/256/256, `out_channels = 256`),

N in {18, 34} (BasicBlock) or {50, 101, 152} (Bottleneck).  Module / parameter names equal the reference's
(`conv1`, `bn1`, `layer1.0.conv1.weight`, `layer2.0.downsample.0.weight`, `deconv_layers.0.conv_offset_mask.weight`,
`deconv_layers.3.weight`, ...), so reference checkpoints load with `load_state_dict`.

The modules only hold parameters.  Execution is a flat schedule of the package's sm_100a kernels over NHWC bf16
buffers: tcgen05 implicit-GEMM convolutions with the folded BatchNorm / residual / ReLU in the epilogue, the DCNv2
sampler kernel, a padded 3x3/2 max-pool, and every dense ConvTranspose2d(4, 2, 1) as ONE 3x3 convolution that
produces the four output phases as channel blocks followed by a pixel shuffle (ops.deconv4x4s2).
"""
import torch
from torch import nn

from ... import ops
from ...DCN.dcn_v2 import DCN
from ...ops import View
from .pose_dla_dcn import BN_MOMENTUM, _bn_params, _fill_bilinear, fold_bn


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


class BasicBlock(nn.Module):          # resnet_dcn.py:36-65, msra_resnet.py:32-61
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = _bn(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _bn(planes)
        self.downsample, self.stride = downsample, stride


class Bottleneck(nn.Module):          # resnet_dcn.py:68-106, msra_resnet.py:64-100
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample, self.stride = downsample, stride


resnet_spec = {18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]), 50: (Bottleneck, [3, 4, 6, 3]),
               101: (Bottleneck, [3, 4, 23, 3]), 152: (Bottleneck, [3, 8, 36, 3])}


class PoseResNet(nn.Module):
    def __init__(self, block, layers, dcn):
        super().__init__()
        self.inplanes = 64
        self.dcn = dcn
        self.out_channels = 64 if dcn else 256
        self.deconv_with_bias = False
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.deconv_layers = self._make_deconv_layer([256, 128, 64] if dcn else [256, 256, 256])
        self._cache = None

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                       _bn(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _make_deconv_layer(self, num_filters):
        layers = []
        for planes in num_filters:
            if self.dcn:   # resnet_dcn.py:196-231
                fc = DCN(self.inplanes, planes, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)
                up = nn.ConvTranspose2d(planes, planes, 4, stride=2, padding=1, output_padding=0,
                                        bias=self.deconv_with_bias)
                _fill_bilinear(up)     # resnet_dcn.py:109-118: the same separable bilinear kernel in every channel
                layers += [fc, _bn(planes), nn.ReLU(inplace=True), up, _bn(planes), nn.ReLU(inplace=True)]
            else:          # msra_resnet.py:161-181
                layers += [nn.ConvTranspose2d(self.inplanes, planes, 4, stride=2, padding=1, output_padding=0,
                                              bias=self.deconv_with_bias), _bn(planes), nn.ReLU(inplace=True)]
            self.inplanes = planes
        return nn.Sequential(*layers)

    # ---- packed weights / folded BN: version-stamped entries (ops.PackCache) ------------------------------
    def invalidate_caches(self):
        self._cache = None
        for m in self.modules():
            if isinstance(m, DCN):
                m._packed.clear()

    def _apply(self, fn, *a, **k):
        self.invalidate_caches()
        return super()._apply(fn, *a, **k)

    # ---- engine ------------------------------------------------------------------------------------------------
    def _conv(self, conv, bn):
        def build():
            w_kw = conv.kernel_size[1] + 1 if (conv.in_channels <= 8 and conv.kernel_size[1] > 1) else 0
            return (ops.pack_conv_weights(conv.weight, kw_pad=w_kw or None), *fold_bn(bn, conv.bias), w_kw)
        return self._cache.get(id(conv), (conv.weight, conv.bias) + _bn_params(bn), build)

    def _cba(self, x, conv, bn, act=1, res=None):
        wpk, scale, shift, w_kw = self._conv(conv, bn)
        k, s = conv.kernel_size[0], conv.stride[0]
        y = ops.conv2d(x, wpk, conv.out_channels, k, s, k // 2, scale, shift, res=res, act=act, w_kw=w_kw)
        return View(y, conv.out_channels, 0)

    def _block(self, blk, x):
        residual = x if blk.downsample is None else self._cba(x, blk.downsample[0], blk.downsample[1], act=0)
        if isinstance(blk, Bottleneck):
            h = self._cba(x, blk.conv1, blk.bn1)
            h = self._cba(h, blk.conv2, blk.bn2)
            return self._cba(h, blk.conv3, blk.bn3, act=1, res=residual)
        h = self._cba(x, blk.conv1, blk.bn1)
        return self._cba(h, blk.conv2, blk.bn2, act=1, res=residual)

    def _up(self, x, up, bn):
        def build():   # scale/shift repeated once per output phase here, not per call
            scale, shift = fold_bn(bn, up.bias)
            return (ops.pack_deconv4x4s2_weights(up.weight), scale.repeat(4).contiguous(), shift.repeat(4).contiguous())
        wpk, scale4, shift4 = self._cache.get(id(up), (up.weight, up.bias) + _bn_params(bn), build)
        return View(ops.deconv4x4s2(x, wpk, up.out_channels, scale4, shift4, act=1), up.out_channels, 0)

    def forward_nhwc(self, x):
        """x [B,3,H,W] fp32 (CUDA) -> View of the [B,H/4,W/4,out_channels] bf16 NHWC feature map."""
        if self.training:
            raise NotImplementedError("centernet_b200 PoseResNet: training-mode forward/backward is not built yet "
                                      "(inference engine only); call .eval()")
        if self._cache is None:
            self._cache = ops.PackCache()
        h = View(ops.to_nhwc_bf16(x, c_pad=8), 8, 0)
        h = self._cba(h, self.conv1, self.bn1)
        h = View(ops.maxpool2d_pad(h, 3, 2, 1), 64, 0)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                h = self._block(blk, h)
        mods = list(self.deconv_layers)
        step = 6 if self.dcn else 3
        for i in range(0, len(mods), step):
            if self.dcn:
                fc, bn_fc, _, up, bn_up, _ = mods[i:i + 6]
                scale, shift = self._cache.get(id(fc), (fc.bias,) + _bn_params(bn_fc),
                                               lambda: fold_bn(bn_fc, fc.bias))
                h = View(fc.forward_nhwc(h, scale=scale, shift=shift, act=1), fc.out_channels, 0)
                h = self._up(h, up, bn_up)
            else:
                up, bn_up, _ = mods[i:i + 3]
                h = self._up(h, up, bn_up)
        return h

    def forward(self, x):
        v = self.forward_nhwc(x)
        out = ops.to_nchw_f32(v)
        out._cnb_nhwc = v            # lets CenterHead skip the NCHW fp32 -> NHWC bf16 round trip
        return [out]

    def init_weights(self, num_layers=None, pretrained=False):
        """The reference downloads ImageNet weights here (resnet_dcn.py:251-263, msra_resnet.py:222-243); this
        build has no network access: deconv / BN init only, backbone weights come from `load_state_dict`."""
        for m in self.deconv_layers.modules():
            if isinstance(m, nn.ConvTranspose2d) and not self.dcn:
                nn.init.normal_(m.weight, std=0.001)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def get_pose_net_dcn(num_layers):
    """resnet_dcn.py:274-280 (without the ImageNet download)."""
    block, layers = resnet_spec[num_layers]
    model = PoseResNet(block, layers, dcn=True)
    model.init_weights(num_layers)
    return model


def get_pose_net(num_layers):
    """msra_resnet.py:254-260 (without the ImageNet download)."""
    block, layers = resnet_spec[num_layers]
    model = PoseResNet(block, layers, dcn=False)
    model.init_weights(num_layers)
    return model
