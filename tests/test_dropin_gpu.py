"""-m gpu: second link of the drop-in chain (see tests/test_dropin_swap.py, oracle/task_torch.py).

The restatement of `CenterNetDetection` / `CenterNetMultiPose` (`loss`, decode step of `test_step_end`) runs on THIS
package's functions (sigmoid_clamped, FocalLoss, RegL1Loss, RegWeightedL1Loss, ctdet_decode, multi_pose_decode, all CUDA)
and must reproduce what the unmodified reference classes produced on the CPU for the same seeded inputs
(tests/golden/task_*.npz, written by oracle/make_golden_tasks.py): losses and every loss_stats entry to 1e-5 relative,
the gradient of the composed loss w.r.t. every head map to 1e-5 of its largest entry, decoded detections bit-exact,
and the post-processed `test_step_end` result equal.  Where /root/reference is present next to a GPU the unmodified
classes themselves are driven through INTEGRATION.md's import swap (forward -> loss -> backward -> test_step_end).
"""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim, task_torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(kind):
    g = np.load(os.path.join(GOLD, "task_detection.npz" if kind == "ctdet" else "task_multi_pose.npz"))
    out = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("out_")}
    tgt = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("tgt_")}
    return g, out, tgt


@pytest.mark.parametrize("kind", ["ctdet", "pose"])
def test_task_loss_and_decode_match_reference_golden(cuda_dev, kind):
    g, out, tgt = _load(kind)
    ns = task_torch.namespace("b200")
    task = (task_torch.DetectionTask if kind == "ctdet" else task_torch.MultiPoseTask)(ns, "dla_34", build_backbone=False)
    dev = cuda_dev
    leaves = {k: v.to(dev).requires_grad_(True) for k, v in out.items()}
    t = {k: v.to(dev) for k, v in tgt.items()}
    loss, stats = task.loss([{k: v * 1 for k, v in leaves.items()}], t)
    for k, v in stats.items():
        want = float(g[f"stat_{k}"])
        assert abs(float(v.detach()) - want) <= 1e-5 * abs(want) + 1e-7, (k, float(v.detach()), want)
    loss.backward()
    for k, v in leaves.items():
        want = g[f"grad_{k}"]
        err = np.abs(v.grad.cpu().numpy() - want).max()
        assert err <= 1e-5 * np.abs(want).max() + 1e-9, (k, err)
    one = {k: v[:1].to(dev).clone() for k, v in out.items()}
    det = task.decode(one).cpu().numpy()
    # `test_step_end` applies torch's own `.sigmoid_()` to the heat maps before decoding: ATen's CUDA sigmoid differs from
    # the CPU one by <= 1 ulp, so score columns are compared to 2e-7; boxes, classes, key-points (gathers at the
    # selected indices, no transcendental) are bit-exact
    score_cols = [4] if kind == "ctdet" else [4] + list(range(40, 57))
    other = [c for c in range(det.shape[-1]) if c not in score_cols]
    want = g["decoded"]
    assert np.array_equal(det[..., other], want[..., other]), "decoded boxes / classes / key-points differ from the reference's"
    assert np.abs(det[..., score_cols] - want[..., score_cols]).max() <= 2e-7
    # post-processing of test_step_end (centernet_detection.py:188-223 / centernet_multi_pose.py:232-262), restated
    pad, scale = g["meta_padding"], g["meta_scale"]
    d = det[0].copy()
    d[:, :4] = (d[:, :4] * np.float32(4) - np.concatenate([pad, pad])) / np.concatenate([scale, scale])
    if kind == "ctdet":
        rows = [np.concatenate([np.full((int((d[:, 5] == j).sum()), 1), j + 1, np.float32), d[d[:, 5] == j, :5]], 1)
                for j in range(80) if (d[:, 5] == j).any()]
        res = np.concatenate(rows, 0)
        assert res.shape == g["results"].shape and np.allclose(res, g["results"], rtol=0, atol=2e-7)
    else:
        pts = d[:, 5:39].reshape(-1, 17, 2)
        d[:, 5:39] = ((pts * np.float32(4) - pad) / scale).reshape(-1, 34)
        kth = len(d) - 20
        keep = d[:, 4] >= np.partition(d[:, 4], kth)[kth]
        assert d[keep].shape == g["results"].shape and np.allclose(d[keep], g["results"], rtol=0, atol=2e-7)


@pytest.mark.skipif(not ref_shim.available(), reason="the unmodified reference classes need /root/reference")
def test_reference_lightning_modules_through_the_swap(cuda_dev):
    """forward -> loss -> backward -> test_step_end of the reference's own classes on this package's kernels."""
    saved = ref_shim.install_b200_swap()
    try:
        _, Det, _ = ref_shim.ref_tasks()
        det = Det("dla_34").to(cuda_dev).train()
        x = torch.rand(2, 3, 128, 128, device=cuda_dev)
        _, tgt = task_torch.task_inputs("ctdet", B=2, H=32, W=32)
        outputs = det(x)
        loss, stats = det.loss(outputs, {k: v.to(cuda_dev) for k, v in tgt.items()})
        loss.backward()
        assert torch.isfinite(loss) and all(p.grad is not None for n, p in det.named_parameters() if "project" not in n)
        det.eval()
        with torch.no_grad():
            outs = det(x[:1])
        image_id, results = det.test_step_end((3, outs, [{"padding": [0.0, 0.0], "scale": [1.0, 1.0]}]))
        assert image_id == 3 and sum(len(v) for v in results.values()) == 100
    finally:
        ref_shim.remove_b200_swap(saved)
