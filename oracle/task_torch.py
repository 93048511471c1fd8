"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's task classes around the hot path, parameterised by the
namespace that provides the hot-path functions:

    CenterNet.__init__            CenterNet/centernet.py:10-21        (head_conv / num_stacks per arch, create_model)
    CenterNetDetection            CenterNet/centernet_detection.py:44-95 (heads, criteria), :97-130 loss,
                                  :183-187 decode step of test_step_end
    CenterNetMultiPose            CenterNet/centernet_multi_pose.py:36-95, :97-155 loss, :223-231 decode step

`namespace("reference")` binds the reference's own modules (CPU; needs /root/reference) and `namespace("b200")` this
repo's drop-ins (CUDA).  Why it exists: the reference cannot travel to the GPU box, so the drop-in is proven as a chain
-- (1) here, next to the reference: the unmodified LightningModules == this restatement on the reference namespace,
bit for bit (tests/test_dropin_swap.py, fixtures tests/golden/task_*.npz); (2) on the GPU: this restatement on the
b200 namespace == those fixtures (tests/test_dropin_gpu.py).  Nothing in the product package imports this file.
"""
import types

import torch
from torch import nn

CTDET_HEADS = lambda num_classes=80: {"heatmap": num_classes, "width_height": 2, "regression": 2}   # noqa: E731
POSE_HEADS = {"heatmap": 1, "width_height": 2, "regression": 2, "heatmap_keypoints": 17, "keypoints": 34,
              "heatmap_keypoints_offset": 2}


def namespace(kind):
    ns = types.SimpleNamespace()
    if kind == "reference":
        from oracle import ref_shim
        ref_shim.install()
        from CenterNet.decode.ctdet import ctdet_decode
        from CenterNet.decode.multi_pose import multi_pose_decode
        from CenterNet.models import create_model
        from CenterNet.models.heads import CenterHead
        from CenterNet.utils.decode import sigmoid_clamped
        from CenterNet.utils.losses import FocalLoss, RegL1Loss, RegWeightedL1Loss
    else:
        from centernet_pytorch_lightning_b200.decode.ctdet import ctdet_decode
        from centernet_pytorch_lightning_b200.decode.multi_pose import multi_pose_decode
        from centernet_pytorch_lightning_b200.models import create_model
        from centernet_pytorch_lightning_b200.models.heads import CenterHead
        from centernet_pytorch_lightning_b200.utils.decode import sigmoid_clamped
        from centernet_pytorch_lightning_b200.utils.losses import FocalLoss, RegL1Loss, RegWeightedL1Loss
    ns.__dict__.update(ctdet_decode=ctdet_decode, multi_pose_decode=multi_pose_decode, create_model=create_model,
                       CenterHead=CenterHead, sigmoid_clamped=sigmoid_clamped, FocalLoss=FocalLoss, RegL1Loss=RegL1Loss,
                       RegWeightedL1Loss=RegWeightedL1Loss)
    return ns


class _Task(nn.Module):
    def __init__(self, ns, arch, heads, build_backbone=True):
        super().__init__()
        self.ns, self.arch = ns, arch
        self.head_conv = 256 if "dla" in arch or "hourglass" in arch else 64     # centernet.py:15
        self.num_stacks = 2 if "hourglass" in arch else 1                        # :16
        self.down_ratio = 4
        out_channels = {"dla": 64, "resdcn": 64, "res": 256, "hourglass": 256}[arch.split("_")[0]]
        if build_backbone:
            self.backbone = ns.create_model(arch)
            out_channels = self.backbone.out_channels
        self.heads = nn.ModuleList([ns.CenterHead(heads, out_channels, self.head_conv) for _ in range(self.num_stacks)])

    def forward(self, x):
        return [head(out) for head, out in zip(self.heads, self.backbone(x))]


class DetectionTask(_Task):
    def __init__(self, ns, arch, hm_weight=1, wh_weight=0.1, off_weight=1, num_classes=80, build_backbone=True):
        super().__init__(ns, arch, CTDET_HEADS(num_classes), build_backbone)
        self.w = (hm_weight, wh_weight, off_weight)
        self.criterion, self.criterion_regression, self.criterion_width_height = ns.FocalLoss(), ns.RegL1Loss(), ns.RegL1Loss()

    def loss(self, outputs, target):
        hm_loss, wh_loss, off_loss = 0, 0, 0
        for output in outputs:
            output["heatmap"] = self.ns.sigmoid_clamped(output["heatmap"])
            hm_loss += self.criterion(output["heatmap"], target["heatmap"])
            wh_loss += self.criterion_width_height(output["width_height"], target["regression_mask"], target["indices"],
                                                   target["width_height"])
            off_loss += self.criterion_regression(output["regression"], target["regression_mask"], target["indices"],
                                                  target["regression"])
        loss = (self.w[0] * hm_loss + self.w[1] * wh_loss + self.w[2] * off_loss) / len(outputs)
        return loss, {"loss": loss, "hm_loss": hm_loss, "wh_loss": wh_loss, "off_loss": off_loss}

    def decode(self, output):
        return self.ns.ctdet_decode(output["heatmap"].sigmoid_(), output["width_height"], reg=output["regression"])


class MultiPoseTask(_Task):
    def __init__(self, ns, arch, hm_weight=1, wh_weight=0.1, off_weight=1, hp_weight=1, hm_hp_weight=1,
                 build_backbone=True):
        super().__init__(ns, arch, dict(POSE_HEADS), build_backbone)
        self.w = (hm_weight, wh_weight, off_weight, hp_weight, hm_hp_weight)
        self.criterion, self.criterion_heatmap_keypoints = ns.FocalLoss(), ns.FocalLoss()
        self.criterion_keypoints = ns.RegWeightedL1Loss()
        self.criterion_regression, self.criterion_width_height = ns.RegL1Loss(), ns.RegL1Loss()

    def loss(self, outputs, target):
        hm_loss = wh_loss = off_loss = kp_loss = hm_kp_loss = hm_offset_loss = 0
        for output in outputs:
            output["heatmap"] = self.ns.sigmoid_clamped(output["heatmap"])
            output["heatmap_keypoints"] = self.ns.sigmoid_clamped(output["heatmap_keypoints"])
            hm_loss += self.criterion(output["heatmap"], target["heatmap"])
            wh_loss += self.criterion_width_height(output["width_height"], target["regression_mask"], target["indices"],
                                                   target["width_height"])
            off_loss += self.criterion_regression(output["regression"], target["regression_mask"], target["indices"],
                                                  target["regression"])
            kp_loss += self.criterion_keypoints(output["keypoints"], target["keypoints_mask"], target["indices"],
                                                target["keypoints"])
            hm_kp_loss += self.criterion_heatmap_keypoints(output["heatmap_keypoints"], target["heatmap_keypoints"])
            hm_offset_loss += self.criterion_regression(output["heatmap_keypoints_offset"], target["heatmap_keypoints_mask"],
                                                        target["heatmap_keypoints_indices"],
                                                        target["heatmap_keypoints_offset"])
        hm_w, wh_w, off_w, hp_w, hm_hp_w = self.w
        loss = (hm_w * hm_loss + wh_w * wh_loss + off_w * off_loss + hp_w * kp_loss + hm_hp_w * hm_kp_loss
                + off_w * hm_offset_loss) / len(outputs)
        return loss, {"loss": loss, "hm_loss": hm_loss, "kp_loss": kp_loss, "hm_kp_loss": hm_kp_loss,
                      "hm_offset_loss": hm_offset_loss, "wh_loss": wh_loss, "off_loss": off_loss}

    def decode(self, output):
        return self.ns.multi_pose_decode(output["heatmap"].sigmoid_(), output["width_height"], output["keypoints"],
                                         reg=output["regression"], hm_hp=output["heatmap_keypoints"].sigmoid_(),
                                         hp_offset=output["heatmap_keypoints_offset"])


# ---- seeded task-level inputs shared by the fixture generator and the tests ----------------------------------------
def task_inputs(kind, B=2, H=32, W=32, M=16, seed=11):
    """Head maps (raw logits, tie-free heat planes) + targets for `loss`, as CPU tensors."""
    import numpy as np
    from centernet_pytorch_lightning_b200.utils import synthetic
    rng = np.random.default_rng(seed)
    T = torch.from_numpy
    C = 80 if kind == "ctdet" else 1

    def logits(c, s):   # distinct probabilities in (0.02, 0.98) -> logits
        p = synthetic.distinct_uniform_heat(B, c, H, W, seed=s, lo=0.02, hi=0.98).astype(np.float64)
        return np.log(p / (1 - p)).astype(np.float32)

    def gt_map(c, s):
        g = (np.random.default_rng(s).random((B, c, H, W)) ** 6).astype(np.float32)
        g[np.random.default_rng(s + 1).random((B, c, H, W)) > 0.995] = 1.0
        return g

    out = {"heatmap": T(logits(C, seed)), "width_height": T((8 * rng.random((B, 2, H, W))).astype(np.float32)),
           "regression": T(rng.random((B, 2, H, W)).astype(np.float32))}
    ind = rng.integers(0, H * W, size=(B, M)).astype(np.int64)
    tgt = {"heatmap": T(gt_map(C, seed + 2)), "regression_mask": T(rng.random((B, M)) > 0.3), "indices": T(ind),
           "width_height": T((8 * rng.random((B, M, 2))).astype(np.float32)),
           "regression": T(rng.random((B, M, 2)).astype(np.float32))}
    if kind == "pose":
        J = 17
        out.update({"heatmap_keypoints": T(logits(J, seed + 5)),
                    "keypoints": T((3 * rng.standard_normal((B, 2 * J, H, W))).astype(np.float32)),
                    "heatmap_keypoints_offset": T(rng.random((B, 2, H, W)).astype(np.float32))})
        tgt.update({"keypoints_mask": T((rng.random((B, M, 2 * J)) > 0.4).astype(np.float32)),
                    "keypoints": T((3 * rng.standard_normal((B, M, 2 * J))).astype(np.float32)),
                    "heatmap_keypoints": T(gt_map(J, seed + 7)),
                    "heatmap_keypoints_mask": T(rng.random((B, M * J)) > 0.5),
                    "heatmap_keypoints_indices": T(rng.integers(0, H * W, size=(B, M * J)).astype(np.int64)),
                    "heatmap_keypoints_offset": T(rng.random((B, M * J, 2)).astype(np.float32))})
    return out, tgt


def ctdet_loss_torch(output, target, hm_weight=1.0, wh_weight=0.1, off_weight=1.0):
    """CenterNetDetection.loss on CPU tensors in plain differentiable torch (centernet_detection.py:97-130 with
    utils/decode.py:43-45, utils/losses.py:14-63 written out) -- the oracle of the whole-step training test."""
    pred = torch.clamp(torch.sigmoid(output["heatmap"]), min=1e-4, max=1 - 1e-4)
    gt = target["heatmap"]
    pos, neg = gt.eq(1).float(), gt.lt(1).float()
    pos_loss = (torch.log(pred) * torch.pow(1 - pred, 2) * pos).sum()
    neg_loss = (torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg).sum()
    num_pos = pos.sum()
    hm = -neg_loss if num_pos == 0 else -(pos_loss + neg_loss) / num_pos

    def reg_l1(out, mask, ind, tgt):
        B, C = out.shape[:2]
        feat = out.permute(0, 2, 3, 1).contiguous().view(B, -1, C)
        p = feat.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))
        m = mask.unsqueeze(2).expand_as(p).float()
        return torch.nn.functional.l1_loss(p * m, tgt * m, reduction="sum") / (m.sum() + 1e-4)

    wh = reg_l1(output["width_height"], target["regression_mask"], target["indices"], target["width_height"])
    off = reg_l1(output["regression"], target["regression_mask"], target["indices"], target["regression"])
    return hm_weight * hm + wh_weight * wh + off_weight * off
