"""CPU: the drop-in is EXECUTED against the reference's own task classes (needs /root/reference; skipped on the GPU box).

  1. `CenterNetDetection` / `CenterNetMultiPose` (source unmodified, Lightning stubbed) == the restatement
     oracle/task_torch.py on the reference namespace, bit for bit, and both == tests/golden/task_*.npz -- the first link
     of the chain whose second link runs on the GPU (tests/test_dropin_gpu.py).
  2. With INTEGRATION.md's import swap installed, the same unmodified classes import and construct on this repo's
     modules: backbone / heads / criteria / decode functions are this package's, state-dict keys and shapes equal the
     reference's, the legacy-checkpoint remap (centernet.py:23-62) loads through them, every name the reference imports
     from the swapped modules resolves, and a CPU forward raises (no silent fallback).
"""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim, task_torch

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout (/root/reference)")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _clone(d):
    return {k: v.clone() for k, v in d.items()}


@pytest.mark.parametrize("kind", ["ctdet", "pose"])
def test_reference_tasks_equal_restatement_and_golden(kind):
    from oracle.make_golden_tasks import detection_results_array
    _, Det, Pose = ref_shim.ref_tasks()
    ns = task_torch.namespace("reference")
    ref = (Det if kind == "ctdet" else Pose)("res_18")
    task = (task_torch.DetectionTask if kind == "ctdet" else task_torch.MultiPoseTask)(ns, "res_18", build_backbone=False)
    gold = np.load(os.path.join(GOLD, "task_detection.npz" if kind == "ctdet" else "task_multi_pose.npz"))
    out, tgt = task_torch.task_inputs(kind)
    for k, v in out.items():
        assert np.array_equal(gold[f"out_{k}"], v.numpy())
    loss, stats = ref.loss([_clone(out)], tgt)
    loss2, stats2 = task.loss([_clone(out)], tgt)
    assert torch.equal(loss, loss2)
    for k in stats:
        assert float(stats[k]) == float(stats2[k]) == float(gold[f"stat_{k}"]), k
    one = {k: v[:1].clone() for k, v in out.items()}
    meta = {"padding": gold["meta_padding"].tolist(), "scale": gold["meta_scale"].tolist()}
    _, results = ref.test_step_end((0, [_clone(one)], [copy.deepcopy(meta)]))
    got = detection_results_array(results) if kind == "ctdet" else np.asarray(results, np.float32)
    assert np.array_equal(got, gold["results"])
    assert np.array_equal(task.decode(_clone(one)).numpy(), gold["decoded"])


@pytest.fixture
def swapped():
    saved = ref_shim.install_b200_swap()
    try:
        yield ref_shim.ref_tasks()
    finally:
        ref_shim.remove_b200_swap(saved)


def test_swap_constructs_reference_tasks_on_b200_modules(swapped):
    import centernet_pytorch_lightning_b200 as b
    _, Det, Pose = swapped
    det = Det("dla_34")
    assert isinstance(det.backbone, b.models.backbones.pose_dla_dcn.DLASeg)
    assert all(isinstance(h, b.models.heads.CenterHead) for h in det.heads)
    assert isinstance(det.criterion, b.utils.losses.FocalLoss) and isinstance(det.criterion_regression, b.utils.losses.RegL1Loss)
    assert det.head_conv == 256 and det.backbone.out_channels == 64
    # same parameter names / shapes as the reference-built task (checkpoints interchange)
    ref_backbone = ref_shim.ref_dlaseg()
    want = {k: tuple(v.shape) for k, v in ref_backbone.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in det.backbone.state_dict().items()}
    assert got == want
    assert sorted(det.heads.state_dict()) == sorted(
        f"0.{n}.fc.{i}.{p}" for n in ("heatmap", "width_height", "regression") for i in (0, 2) for p in ("weight", "bias"))
    pose = Pose("dla_34")
    assert isinstance(pose.criterion_keypoints, b.utils.losses.RegWeightedL1Loss)
    assert set(pose.heads[0].heads) == set(task_torch.POSE_HEADS)
    # the product path has no CPU fallback: the reference's own forward raises on CPU tensors
    with pytest.raises(b._lib.CnbError):
        det(torch.zeros(1, 3, 64, 64))
    with pytest.raises(b._lib.CnbError):
        det.loss([{"heatmap": torch.zeros(1, 80, 8, 8), "width_height": torch.zeros(1, 2, 8, 8),
                   "regression": torch.zeros(1, 2, 8, 8)}], task_torch.task_inputs("ctdet", B=1, H=8, W=8)[1])


def test_swap_leaves_no_dangling_import(swapped):
    """every name any reference file imports from the swapped modules resolves in the replacements"""
    from CenterNet.utils.decode import (_gather_feat, _nms, _topk, _topk_channel, _transpose_and_gather_feat,  # noqa: F401
                                        sigmoid_clamped)
    from CenterNet.utils.losses import FocalLoss, RegL1Loss, RegWeightedL1Loss  # noqa: F401
    from CenterNet.decode.ctdet import ctdet_decode  # noqa: F401
    from CenterNet.decode.multi_pose import multi_pose_decode  # noqa: F401
    from CenterNet.models import create_model  # noqa: F401
    from CenterNet.models.heads import CenterHead  # noqa: F401
    from DCN.dcn_v2 import DCN  # noqa: F401


def test_legacy_checkpoint_remap_through_the_swap(swapped, tmp_path):
    """centernet.py:23-62 `load_pretrained_weights` (xingyizhou/CenterNet key layout: `module.<backbone key>`,
    `module.hm.0.weight`, `module.wh.2.bias`, ...) run unmodified on this package's modules."""
    _, Det, _ = swapped
    det = Det("dla_34")
    legacy = {"hm": "heatmap", "wh": "width_height", "reg": "regression"}
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, v in det.backbone.state_dict().items():
        sd["module." + k] = torch.rand(v.shape, generator=g).to(v.dtype) if v.is_floating_point() else v.clone()
    for short, name in legacy.items():
        for i in (0, 2):
            for p in ("weight", "bias"):
                v = det.heads.state_dict()[f"0.{name}.fc.{i}.{p}"]
                sd[f"module.{short}.{i}.{p}"] = torch.rand(v.shape, generator=g)
    path = os.path.join(tmp_path, "legacy.pth")
    torch.save({"state_dict": sd}, path)
    det.load_pretrained_weights(path)
    for k, v in det.backbone.state_dict().items():
        assert torch.equal(v, sd["module." + k]), k
    for short, name in legacy.items():
        assert torch.equal(det.heads.state_dict()[f"0.{name}.fc.2.weight"], sd[f"module.{short}.2.weight"])
