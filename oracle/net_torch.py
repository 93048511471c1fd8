"""TEST INFRASTRUCTURE ONLY -- CPU fp32 oracle for the network half of the hot path.

A *functional* PyTorch restatement (state-dict driven, no nn.Module classes) of the reference's
DLA-34 + DLAUp/IDAUp + CenterHead forward in eval mode:
  CenterNet/models/backbones/pose_dla_dcn.py:28-68 (BasicBlock), :165-188 (Root), :191-265 (Tree),
  :268-378 (DLA), :435-454 (DeformConv), :457-488 (IDAUp), :491-516 (DLAUp), :532-570 (DLASeg),
  CenterNet/models/heads.py:4-50, and the external DCN.dcn_v2.DCN (tteepe/DCNv2, unpinned,
  requirements.txt:1; restated on torchvision.ops.deform_conv2d -- "parity unpinned" by the reference).
It is the floating-point reference the tier allows ("keep a torch fp32 reference only for a
floating-point kernel") and the CPU baseline bench.py times.  Pinned bit-for-bit against the unmodified
reference modules by tests/test_oracle_net.py (runs where /root/reference exists).
"""
import math

import torch
import torch.nn.functional as F
from torchvision.ops import deform_conv2d


_TRAINING = False   # train-mode BatchNorm (batch statistics, running-stat update in `sd`): see `training(...)` below


class training:
    """`with net_torch.training():` runs the restatement like `module.train()`: BatchNorm normalises with batch
    statistics and updates the running ones in the state dict (pose_dla_dcn.py:40, momentum 0.1); gradients flow to
    every tensor of `sd` that requires grad (centernet.py:70-80)."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _TRAINING
        self.prev, _TRAINING = _TRAINING, self.on

    def __exit__(self, *exc):
        global _TRAINING
        _TRAINING = self.prev


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        _TRAINING, 0.1, 1e-5)


def _conv(sd, p, x, stride=1):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd.get(p + ".bias"), stride=stride, padding=w.shape[2] // 2)


def _block(sd, p, x, residual, stride):
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride)))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out))
    return F.relu(out + residual)


def _tree(sd, p, levels, x, stride, level_root, children=None):
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
    residual = bottom
    if (p + ".project.0.weight") in sd:   # (levels == 2: computed and never used, as in Tree.forward :254-255; in
        # train mode its BatchNorm still updates the running statistics)
        residual = _bn(sd, p + ".project.1", _conv(sd, p + ".project.0", bottom))
    if level_root:
        children.append(bottom)
    if levels == 1:
        x1 = _block(sd, p + ".tree1", x, residual, stride)
        x2 = _block(sd, p + ".tree2", x1, x1, 1)
        cat = torch.cat([x2, x1] + children, 1)
        return F.relu(_bn(sd, p + ".root.bn", _conv(sd, p + ".root.conv", cat)))
    x1 = _tree(sd, p + ".tree1", levels - 1, x, stride, False)
    children.append(x1)
    return _tree(sd, p + ".tree2", levels - 1, x1, 1, False, children)


def dcn(sd, p, x):
    """DCNv2 forward: om = conv(x); o1,o2,m = chunk(om,3); offset = cat(o1,o2); mask = sigmoid(m)."""
    om = F.conv2d(x, sd[p + ".conv_offset_mask.weight"], sd[p + ".conv_offset_mask.bias"], padding=1)
    o1, o2, m = torch.chunk(om, 3, dim=1)
    return deform_conv2d(x, torch.cat((o1, o2), 1), sd[p + ".weight"], sd[p + ".bias"], padding=1,
                         mask=torch.sigmoid(m))


def _deform(sd, p, x):
    return F.relu(_bn(sd, p + ".actf.0", dcn(sd, p + ".conv", x)))


def _ida(sd, p, layers, startp, endp):
    for i in range(startp + 1, endp):
        j = i - startp
        w = sd[f"{p}.up_{j}.weight"]
        f = w.shape[2] // 2
        up = F.conv_transpose2d(_deform(sd, f"{p}.proj_{j}", layers[i]), w, stride=f, padding=f // 2,
                                groups=w.shape[0])
        layers[i] = _deform(sd, f"{p}.node_{j}", up + layers[i - 1])


DLA34_LEVELS = [1, 1, 1, 2, 2, 1]


def dla34_seg_forward(sd, x, first_level=2, last_level=5):
    """DLASeg('dla34', down_ratio=4, last_level=5).forward -> [B,64,H/4,W/4]."""
    h = F.relu(_bn(sd, "base.base_layer.1", _conv(sd, "base.base_layer.0", x)))
    feats = []
    h = F.relu(_bn(sd, "base.level0.1", _conv(sd, "base.level0.0", h)))
    feats.append(h)
    h = F.relu(_bn(sd, "base.level1.1", _conv(sd, "base.level1.0", h, 2)))
    feats.append(h)
    for lvl in range(2, 6):
        h = _tree(sd, f"base.level{lvl}", DLA34_LEVELS[lvl], h, 2, lvl > 2)
        feats.append(h)
    layers = list(feats)
    outs = [layers[-1]]
    n = len(layers)
    for i in range(n - first_level - 1):
        _ida(sd, f"dla_up.ida_{i}", layers, n - i - 2, n)
        outs.insert(0, layers[-1])
    y = [outs[i].clone() for i in range(last_level - first_level)]
    _ida(sd, "ida_up", y, 0, len(y))
    return y[-1]


def center_head_forward(sd, x, names, prefix=""):
    """CenterHead.forward (heads.py:38-43): name -> conv1x1(relu(conv3x3(x)))."""
    out = {}
    for n in names:
        p = f"{prefix}{n}.fc"
        h = F.relu(F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1))
        out[n] = F.conv2d(h, sd[p + ".2.weight"], sd[p + ".2.bias"])
    return out


# ---- ResNet backbones ------------------------------------------------------------------------------------------
RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]), 50: ("bottle", [3, 4, 6, 3]),
               101: ("bottle", [3, 4, 23, 3]), 152: ("bottle", [3, 8, 36, 3])}


def _res_block(sd, p, x, kind, stride):
    residual = x
    if (p + ".downsample.0.weight") in sd:
        residual = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    if kind == "basic":      # resnet_dcn.py:47-65
        out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride)))
        out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out))
    else:                    # resnet_dcn.py:85-106 (stride on the 3x3)
        out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
        out = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, stride)))
        out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    return F.relu(out + residual)


def pose_resnet_forward(sd, x, num_layers, dcn_variant):
    """PoseResNet.forward of resnet_dcn.py:233-249 (dcn_variant) / msra_resnet.py:183-199 -> [B,C,H/4,W/4]."""
    kind, layers = RESNET_SPEC[num_layers]
    h = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, nblocks in enumerate(layers, start=1):
        for bi in range(nblocks):
            h = _res_block(sd, f"layer{li}.{bi}", h, kind, 2 if (li > 1 and bi == 0) else 1)
    for i in range(3):
        if dcn_variant:
            b = 6 * i
            h = F.relu(_bn(sd, f"deconv_layers.{b + 1}", dcn(sd, f"deconv_layers.{b}", h)))
            h = F.conv_transpose2d(h, sd[f"deconv_layers.{b + 3}.weight"], stride=2, padding=1)
            h = F.relu(_bn(sd, f"deconv_layers.{b + 4}", h))
        else:
            b = 3 * i
            h = F.conv_transpose2d(h, sd[f"deconv_layers.{b}.weight"], stride=2, padding=1)
            h = F.relu(_bn(sd, f"deconv_layers.{b + 1}", h))
    return h


# ---- Hourglass-104 ------------------------------------------------------------------------------------------------
def _hg_conv(sd, p, x, stride=1):
    """`convolution` (large_hourglass.py:8-28): conv (+ BN) + ReLU."""
    w = sd[p + ".conv.weight"]
    y = F.conv2d(x, w, sd.get(p + ".conv.bias"), stride=stride, padding=(w.shape[2] - 1) // 2)
    if (p + ".bn.weight") in sd:
        y = _bn(sd, p + ".bn", y)
    return F.relu(y)


def _hg_res(sd, p, x, stride=1):
    """`residual` (:51-93)."""
    h = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)))
    h = _bn(sd, p + ".bn2", F.conv2d(h, sd[p + ".conv2.weight"], padding=1))
    skip = x
    if (p + ".skip.0.weight") in sd:
        skip = _bn(sd, p + ".skip.1", F.conv2d(x, sd[p + ".skip.0.weight"], stride=stride))
    return F.relu(h + skip)


def _hg_seq(sd, p, x, first_stride=1):
    i = 0
    while (f"{p}.{i}.conv1.weight") in sd:
        x = _hg_res(sd, f"{p}.{i}", x, first_stride if i == 0 else 1)
        i += 1
    return x


def _hg_kp(sd, p, x):
    """`kp_module.forward` (:206-213); the stride-2 first block of low1 replaces pooling (make_hg_layer :315-319)."""
    up1 = _hg_seq(sd, p + ".up1", x)
    low1 = _hg_seq(sd, p + ".low1", x, first_stride=2)
    low2 = _hg_kp(sd, p + ".low2", low1) if (p + ".low2.up1.0.conv1.weight") in sd else _hg_seq(sd, p + ".low2", low1)
    low3 = _hg_seq(sd, p + ".low3", low2)
    return up1 + F.interpolate(low3, scale_factor=2)


def hourglass_forward(sd, x, nstack=2):
    """`exkp.forward` (:297-313) with the HourglassNet arguments -> [stack outputs [B,256,H/4,W/4]]."""
    inter = _hg_res(sd, "pre.1", _hg_conv(sd, "pre.0", x, stride=2), stride=2)
    outs = []
    for ind in range(nstack):
        cnv = _hg_conv(sd, f"cnvs.{ind}", _hg_kp(sd, f"kps.{ind}", inter))
        outs.append(cnv)
        if ind < nstack - 1:
            a = _bn(sd, f"inters_.{ind}.1", F.conv2d(inter, sd[f"inters_.{ind}.0.weight"]))
            b = _bn(sd, f"cnvs_.{ind}.1", F.conv2d(cnv, sd[f"cnvs_.{ind}.0.weight"]))
            inter = _hg_res(sd, f"inters.{ind}", F.relu(a + b))
    return outs


from centernet_pytorch_lightning_b200.utils.synthetic import randomize_  # noqa: E402,F401 (seeded weights shared with bench.py)
