"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

Works only in the authoring container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to pin the oracle restatements and to generate the fixtures
committed under ``tests/golden/``.  Nothing in the product package imports this file.

Shims:
  * ``CenterNet`` is registered as a namespace stub so that ``CenterNet/__init__.py:1-3``
    (which imports pytorch_lightning, absent here) is never executed;
  * ``DCN.dcn_v2.DCN`` (external ``tteepe/DCNv2`` extension, unpinned in
    ``requirements.txt:1``, source not vendored) is provided on top of
    ``torchvision.ops.deform_conv2d`` -- the published DCNv2 semantics
    (offset/mask conv -> chunk(3) -> cat(o1,o2), sigmoid(mask));
  * ``torch.utils.model_zoo.load_url`` returns ``{}`` (no network).
"""
import math
import os
import sys
import types

REF_ROOT = os.environ.get("CENTERNET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "CenterNet"))


def _install_dcn_shim():
    import torch
    from torch import nn
    from torchvision.ops import deform_conv2d

    class DCN(nn.Module):
        """Modulated deformable conv v2 with the public DCNv2 constructor
        (call sites: pose_dla_dcn.py:441-449, resnet_dcn.py:202-210)."""

        def __init__(self, in_channels, out_channels, kernel_size, stride, padding,
                     dilation=1, deformable_groups=1):
            super().__init__()
            kh, kw = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size,) * 2
            self.stride, self.padding, self.dilation = stride, padding, dilation
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
            self.bias = nn.Parameter(torch.zeros(out_channels))
            self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * kh * kw,
                                              kernel_size=(kh, kw), stride=stride,
                                              padding=padding, bias=True)
            stdv = 1.0 / math.sqrt(in_channels * kh * kw)
            self.weight.data.uniform_(-stdv, stdv)
            self.conv_offset_mask.weight.data.zero_()
            self.conv_offset_mask.bias.data.zero_()

        def forward(self, x):
            out = self.conv_offset_mask(x)
            o1, o2, mask = torch.chunk(out, 3, dim=1)
            offset = torch.cat((o1, o2), dim=1)
            mask = torch.sigmoid(mask)
            return deform_conv2d(x, offset, self.weight, self.bias, stride=self.stride,
                                 padding=self.padding, dilation=self.dilation, mask=mask)

    pkg = types.ModuleType("DCN")
    pkg.__path__ = []
    mod = types.ModuleType("DCN.dcn_v2")
    mod.DCN = DCN
    pkg.dcn_v2 = mod
    sys.modules["DCN"] = pkg
    sys.modules["DCN.dcn_v2"] = mod


def install():
    """Make ``import CenterNet.decode.ctdet`` etc. resolve to the reference sources."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if "CenterNet" not in sys.modules:
        stub = types.ModuleType("CenterNet")
        stub.__path__ = [os.path.join(REF_ROOT, "CenterNet")]
        sys.modules["CenterNet"] = stub
    if "DCN.dcn_v2" not in sys.modules:
        _install_dcn_shim()
    import torch.utils.model_zoo as model_zoo

    model_zoo.load_url = lambda *a, **k: {}


def ref_dlaseg():
    """Reference DLASeg('dla34') without the network fetch (pose_dla_dcn.py:573-581)."""
    install()
    from CenterNet.models.backbones.pose_dla_dcn import DLASeg

    return DLASeg("dla34", pretrained=False, down_ratio=4, final_kernel=1, last_level=5)


def ref_resnet_dcn(num_layers):
    install()
    from CenterNet.models.backbones import resnet_dcn

    block, layers = resnet_dcn.resnet_spec[num_layers]
    return resnet_dcn.PoseResNet(block, layers)


def ref_msra_resnet(num_layers):
    install()
    from CenterNet.models.backbones import msra_resnet

    block, layers = msra_resnet.resnet_spec[num_layers]
    return msra_resnet.PoseResNet(block, layers)
