"""-m gpu: the SURVEY.md 8(f) rows on the device vs fixtures produced by the unmodified reference on the CPU
(oracle/make_golden_next.py, oracle/make_golden_tasks.py): target encoding, soft-NMS, TTA prologue / flip merge / post."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "next_rows.npz"))


def test_target_encoding_matches_reference(cuda_dev, g):
    """indices / masks / sizes / offsets bit-exact; heat map: same support and peaks (== 1 exactly at the centres), values
    within 2 ulp of torch.exp at values < 1 (2.4e-7; a numpy restatement of the same arithmetic differs from torch by
    1.2e-7 already)."""
    from centernet_pytorch_lightning_b200.sample.ctdet import CenterDetectionSample
    enc = CenterDetectionSample()
    t = enc.encode_batch(g["enc_boxes"], g["enc_cls"], g["enc_counts"], (512, 512), cuda_dev)
    for k in ("regression_mask", "indices", "width_height", "regression"):
        assert np.array_equal(t[k].cpu().numpy(), g[f"enc_{k}"]), k
    got, want = t["heatmap"].cpu().numpy(), g["enc_heatmap"]
    assert np.array_equal(got == 1, want == 1) and np.array_equal(got > 0, want > 0)
    assert np.abs(got - want).max() <= 2.4e-7
    # the reference's per-sample signature
    anns = [{"bbox": [float(v) for v in g["enc_boxes"][1, k]], "class_id": int(g["enc_cls"][1, k])} for k in range(17)]
    _, one = enc(torch.zeros(3, 512, 512, device=cuda_dev), anns)
    assert np.array_equal(one["indices"].cpu().numpy(), g["enc_indices"][1])
    assert np.abs(one["heatmap"].cpu().numpy() - want[1]).max() <= 2.4e-7


@pytest.mark.parametrize("method", [0, 1, 2])
def test_soft_nms_matches_reference(cuda_dev, g, method):
    from centernet_pytorch_lightning_b200.utils.nms import soft_nms, soft_nms_lists
    boxes = g["nms_in"].copy()
    keep = soft_nms(boxes, Nt=0.5, method=method, device=cuda_dev)
    n = int(g[f"nms_keep_{method}"])
    assert len(keep) == n
    want = g[f"nms_out_{method}"]
    assert np.array_equal(boxes[:n, :4], want[:n, :4]), "kept boxes / selection order differ"
    assert np.abs(boxes[:n, 4] - want[:n, 4]).max() <= 1e-6          # exp() in float64 on both sides, stored as fp32
    # several lists in one launch
    L = torch.from_numpy(np.stack([g["nms_in"], g["nms_in"][::-1].copy()])).to(cuda_dev)
    kept = soft_nms_lists(L, torch.tensor([120, 60], dtype=torch.int32, device=cuda_dev), Nt=0.5, method=method)
    assert int(kept[0]) == n and 0 < int(kept[1]) <= 60


def test_tta_prologue_merge_post(cuda_dev, g):
    from centernet_pytorch_lightning_b200 import tta
    out, pad = tta.prologue(torch.from_numpy(g["tta_img"]).to(cuda_dev), flip=True)
    assert pad == g["tta_pad"].tolist() and out.shape == g["tta_out"].shape
    assert np.abs(out.cpu().numpy() - g["tta_out"]).max() <= 1e-6      # (v - mean) / std vs torchvision's sub_().div_()
    merged = tta.flip_merge(torch.from_numpy(g["merge_in"]).to(cuda_dev))
    assert np.array_equal(merged.cpu().numpy(), g["merge_out"])
    # post-processing vs the reference's own test_step_end (fixture of oracle/make_golden_tasks.py)
    t = np.load(os.path.join(GOLD, "task_detection.npz"))
    rows, counts, offsets = tta.ctdet_post(torch.from_numpy(t["decoded"][0]).to(cuda_dev), t["meta_padding"], t["meta_scale"])
    rows, counts, offsets = rows.cpu().numpy(), counts.cpu().numpy(), offsets.cpu().numpy()
    want = t["results"]                                  # rows (class_id 1-based, x1, y1, x2, y2, score) in class order
    assert counts.sum() == len(want)
    got = np.concatenate([np.concatenate([np.full((counts[c], 1), c + 1, np.float32), rows[offsets[c]:offsets[c] + counts[c]]], 1)
                          for c in range(80) if counts[c]], 0)
    assert np.array_equal(got, want)
