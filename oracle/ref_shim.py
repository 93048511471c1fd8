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


# ---- the LightningModule classes themselves (centernet.py, centernet_detection.py, centernet_multi_pose.py) --------
def install_task_stubs():
    """Stand-ins for what `CenterNet/centernet_detection.py:1-26` imports besides torch and the hot-path modules:
    pytorch_lightning (a 30-line LightningModule: nn.Module + save_hyperparameters / hparams / log), imgaug,
    pycocotools, and the augmentation package `CenterNet.transforms` (imgaug-based, CPU dataloader side, not on the hot
    path; `transforms/sample.py:5` also fails to import on Python >= 3.10).  None of this touches the reference's
    task classes, whose source runs unmodified."""
    import collections
    import collections.abc
    import inspect

    import torch

    if not hasattr(collections, "Callable"):
        collections.Callable = collections.abc.Callable

    class _HParams(dict):
        __getattr__ = dict.__getitem__

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self):
            frame = inspect.currentframe().f_back
            args = inspect.getargvalues(frame)
            self.hparams = _HParams({k: args.locals[k] for k in args.args if k != "self"})

        def log(self, *a, **k):
            self.__dict__.setdefault("logged", []).append((a, k))

    def fake(name, **attrs):
        if name in sys.modules and not getattr(sys.modules[name], "_cnb_stub", False):
            return sys.modules[name]
        m = types.ModuleType(name)
        m._cnb_stub = True
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return a[0] if a else None

    pl = fake("pytorch_lightning", LightningModule=LightningModule, Trainer=_Any, seed_everything=lambda *a, **k: None)
    pl.callbacks = fake("pytorch_lightning.callbacks", ModelCheckpoint=_Any, LearningRateMonitor=_Any)
    pl.loggers = fake("pytorch_lightning.loggers", TensorBoardLogger=_Any)
    ia = fake("imgaug", seed=lambda *a, **k: None)
    ia.augmenters = fake("imgaug.augmenters", Augmenter=_Any, Identity=_Any, Sequential=_Any)
    ia.augmentables = fake("imgaug.augmentables", Keypoint=_Any, KeypointsOnImage=_Any, BoundingBox=_Any,
                           BoundingBoxesOnImage=_Any)
    pc = fake("pycocotools")
    pc.cocoeval = fake("pycocotools.cocoeval", COCOeval=_Any)
    pc.coco = fake("pycocotools.coco", COCO=_Any)
    tr = fake("CenterNet.transforms", CategoryIdToClass=_Any, ImageAugmentation=_Any)
    tr.sample = fake("CenterNet.transforms.sample", ComposeSample=_Any, MultiSampleTransform=_Any, PoseFlip=_Any)
    fake("CenterNet.sample.multi_pose", MultiPoseSample=_Any)   # sample/multi_pose.py:74 breaks on numpy 2


def ref_tasks():
    """(CenterNet, CenterNetDetection, CenterNetMultiPose) -- the reference's task classes, source unmodified, importing
    whatever `CenterNet.models`, `CenterNet.utils.losses`, ... currently resolve to in sys.modules (the reference's own
    modules after `install()`, or this repo's after `install_b200_swap()`)."""
    install()
    install_task_stubs()
    import importlib

    stub = sys.modules["CenterNet"]
    for name in ("CenterNet.centernet", "CenterNet.centernet_detection", "CenterNet.centernet_multi_pose"):
        sys.modules.pop(name, None)
    base = importlib.import_module("CenterNet.centernet")
    stub.CenterNet = base.CenterNet
    det = importlib.import_module("CenterNet.centernet_detection")
    pose = importlib.import_module("CenterNet.centernet_multi_pose")
    return base.CenterNet, det.CenterNetDetection, pose.CenterNetMultiPose


_SWAPPED = ("DCN", "DCN.dcn_v2", "CenterNet.models", "CenterNet.models.heads", "CenterNet.decode.ctdet",
            "CenterNet.decode.multi_pose", "CenterNet.utils.losses", "CenterNet.utils.decode")


def install_b200_swap():
    """INTEGRATION.md section 1, executed: the hot-path modules of the reference resolve to this repo's package."""
    install()
    import centernet_pytorch_lightning_b200 as b

    return b.install_swap()


def remove_b200_swap(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for name in ("CenterNet.centernet", "CenterNet.centernet_detection", "CenterNet.centernet_multi_pose"):
        sys.modules.pop(name, None)
