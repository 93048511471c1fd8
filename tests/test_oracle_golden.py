"""CPU (-m "not gpu"): the oracle restatements against (i) fixtures produced by the unmodified
reference (oracle/make_golden.py) and (ii) the reference's own known-answer test for this path,
tests/test_sample_encode_decode.py:14-56 with tests/data/coco_annotation.json."""
import os

import numpy as np
import pytest

from oracle import decode_np, encode_np, losses_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_ctdet_oracle_vs_reference_fixture():
    g = np.load(os.path.join(GOLD, "ctdet_decode.npz"))
    for tag in ("a", "b"):
        ref = g[f"out_{tag}"]
        got = decode_np.ctdet_decode(g[f"heat_{tag}"], g[f"wh_{tag}"], g[f"reg_{tag}"], K=ref.shape[1])
        assert np.array_equal(got, ref)


def test_multi_pose_oracle_vs_reference_fixture():
    g = np.load(os.path.join(GOLD, "multi_pose_decode.npz"))
    got = decode_np.multi_pose_decode(g["heat"], g["wh"], g["kps"], g["reg"], g["hm_hp"], g["hp_offset"])
    assert np.array_equal(got, g["out"])


def test_reference_known_answer_encode_decode():
    """Mirror of the reference's tests/test_sample_encode_decode.py: encode the 2-box COCO fixture,
    scatter wh/reg into dense maps, decode, keep score > 0.5, x4: sum of centres == sum of annotation
    centres within 1e-3; plus the exact values recorded from the reference run."""
    g = np.load(os.path.join(GOLD, "kat_encode_decode.npz"))
    boxes, cls = g["bboxes"], g["class_ids"]
    enc = encode_np.encode_ctdet(boxes, cls)
    assert list(enc["indices"][:2]) == [7277, 6113]
    np.testing.assert_allclose(enc["width_height"][:2], [[13.2625, 34.5025], [3.78, 8.935]], rtol=1e-6)
    np.testing.assert_allclose(enc["regression"][:2], [[0.83125, 0.65375], [0.9975, 0.52]], atol=1e-5)
    for k in ("indices", "width_height", "regression", "regression_mask"):
        assert np.array_equal(enc[k], g[f"t_{k}"])
    heat = g["t_heatmap"][None]                      # the reference's own encoded heat map
    _, C, H, W = heat.shape
    wh = np.zeros((1, H, W, 2), np.float32)
    reg = np.zeros((1, H, W, 2), np.float32)
    ind = enc["indices"]
    wh[0, ind // W, ind % W] = enc["width_height"]
    reg[0, ind // W, ind % W] = enc["regression"]
    det = decode_np.ctdet_decode(heat, wh.transpose(0, 3, 1, 2), reg.transpose(0, 3, 1, 2))[0]
    det = 4 * det[det[:, 4] > 0.5]
    assert det.shape[0] == 2
    centre = (det[:, :2] + det[:, 2:4]) / 2
    ann_centre = np.stack([boxes[:, 0] + boxes[:, 2] / 2, boxes[:, 1] + boxes[:, 3] / 2], 1)
    assert abs(centre.sum() - ann_centre.sum()) < 1e-3
    np.testing.assert_allclose(sorted(det[:, 0]), [384.43, 412.8], atol=1e-3)
    np.testing.assert_allclose(sorted(det[:, 3]), [207.95, 295.62], atol=1e-3)
    assert np.all(det[:, 4] == 4.0) and np.all(det[:, 5] == 4.0)   # score 1.0, class 1 (x4)


def test_losses_oracle_vs_reference_fixture():
    g = np.load(os.path.join(GOLD, "losses.npz"))
    p = losses_np.sigmoid_clamped(g["logits"])
    np.testing.assert_allclose(p, g["prob"], rtol=2e-7, atol=0)
    loss, dp = losses_np.neg_loss(g["prob"], g["gt"])
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    np.testing.assert_allclose(dp, g["dprob"], rtol=1e-4, atol=1e-7)
    loss, dx = losses_np.focal_with_logits(g["logits"], g["gt"])
    np.testing.assert_allclose(dx, g["dlogits"], rtol=1e-4, atol=1e-7)
    loss0, dp0 = losses_np.neg_loss(g["prob"], g["gt0"])     # num_pos == 0 branch
    assert abs(loss0 - g["loss0"]) <= 1e-5 * abs(g["loss0"])
    np.testing.assert_allclose(dp0, g["dprob0"], rtol=1e-4, atol=1e-7)
    l, gr = losses_np.reg_l1(g["r_out"], g["r_mask"], g["r_ind"], g["r_tgt"])
    assert abs(l - g["r_loss"]) < 1e-6
    np.testing.assert_allclose(gr, g["r_grad"], atol=1e-7)
    l, gr = losses_np.reg_l1(g["w_out"], g["w_mask"], g["r_ind"], g["w_tgt"], per_channel=True)
    assert abs(l - g["w_loss"]) < 1e-6
    np.testing.assert_allclose(gr, g["w_grad"], atol=1e-7)


def test_oracle_tie_rule_and_zero_fill():
    """The documented tie rule: (score desc, flat index asc); filler rows are zero-score, index asc."""
    heat = np.zeros((1, 2, 4, 4), np.float32)
    heat[0, 1, 2, 2] = 0.7
    heat[0, 0, 1, 1] = 0.7
    wh = np.zeros((1, 2, 4, 4), np.float32)
    det = decode_np.ctdet_decode(heat, wh, None, K=5)[0]
    assert list(det[:, 4]) == [np.float32(0.7), np.float32(0.7), 0, 0, 0]
    assert list(det[:, 5]) == [0, 1, 0, 0, 0]                       # class 0 first on the tie
    assert [tuple(r) for r in det[2:, :2]] == [(0.5, 0.5), (1.5, 0.5), (2.5, 0.5)]   # flat 0,1,2


def test_reference_still_agrees_when_available():
    """Only in the authoring container: re-run the unmodified reference against the oracle."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not present on this machine")
    import torch
    ref_shim.install()
    from CenterNet.decode.ctdet import ctdet_decode
    from centernet_pytorch_lightning_b200.utils import synthetic
    heat, wh, reg = synthetic.ctdet_maps(2, 80, 64, 64, seed=99)
    ref = ctdet_decode(torch.from_numpy(heat), torch.from_numpy(wh), torch.from_numpy(reg)).numpy()
    assert np.array_equal(decode_np.ctdet_decode(heat, wh, reg), ref)
