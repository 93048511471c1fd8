"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) in the
authoring container, and checks every oracle restatement against it on the way.

    python -m oracle.make_golden

The GPU box has no /root/reference; the committed fixtures travel instead.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import decode_np, encode_np, losses_np, ref_shim  # noqa: E402
from centernet_pytorch_lightning_b200.utils import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
T = torch.from_numpy


def main():
    ref_shim.install()
    from CenterNet.decode.ctdet import ctdet_decode
    from CenterNet.decode.multi_pose import multi_pose_decode
    from CenterNet.sample.ctdet import CenterDetectionSample
    from CenterNet.utils.decode import sigmoid_clamped
    from CenterNet.utils.losses import FocalLoss, RegL1Loss, RegWeightedL1Loss

    os.makedirs(GOLD, exist_ok=True)

    # ---- ctdet_decode ------------------------------------------------------------------------------
    out = {}
    for tag, (shape, kind, K) in {"a": ((2, 5, 16, 24), "uniform", 100), "b": ((1, 8, 32, 32), "bumps", 40)}.items():
        heat, wh, reg = synthetic.ctdet_maps(*shape, seed=77, kind=kind)
        ref = ctdet_decode(T(heat), T(wh), T(reg), K=K).numpy()
        assert np.array_equal(decode_np.ctdet_decode(heat, wh, reg, K=K), ref)
        out.update({f"heat_{tag}": heat, f"wh_{tag}": wh, f"reg_{tag}": reg, f"out_{tag}": ref})
    np.savez_compressed(os.path.join(GOLD, "ctdet_decode.npz"), **out)

    # ---- multi_pose_decode -------------------------------------------------------------------------
    heat, wh, kps, reg, hm_hp, hpo = synthetic.multi_pose_maps(2, 17, 32, 48, seed=78)  # >= K positive peaks per plane: tie-free
    ref = multi_pose_decode(T(heat), T(wh), T(kps.copy()), T(reg), T(hm_hp), T(hpo)).numpy()
    assert np.array_equal(decode_np.multi_pose_decode(heat, wh, kps, reg, hm_hp, hpo), ref)
    np.savez_compressed(os.path.join(GOLD, "multi_pose_decode.npz"), heat=heat, wh=wh, kps=kps, reg=reg,
                        hm_hp=hm_hp, hp_offset=hpo, out=ref)

    # ---- the reference's known-answer test (tests/test_sample_encode_decode.py) ----------------------
    with open("/root/reference/tests/data/coco_annotation.json") as fh:
        ann = json.load(fh)
    for a in ann:
        a["class_id"] = a["category_id"]   # CategoryIdToClass(range(0, 100)) is the identity map
    _, tgt = CenterDetectionSample()(torch.zeros(3, 512, 512), ann)
    enc = encode_np.encode_ctdet([a["bbox"] for a in ann], [a["class_id"] for a in ann])
    for k in enc:   # heat values off the centre differ by <= 1 ulp (torch.exp vs numpy exp)
        if k == "heatmap":
            assert np.abs(enc[k] - tgt[k].numpy()).max() <= 6e-8 and np.array_equal(enc[k] == 1, tgt[k].numpy() == 1)
        else:
            assert np.array_equal(enc[k], tgt[k].numpy()), k
    kat = dict(bboxes=np.array([a["bbox"] for a in ann], np.float64),
               class_ids=np.array([a["class_id"] for a in ann]),
               **{f"t_{k}": v.numpy() for k, v in tgt.items()})
    np.savez_compressed(os.path.join(GOLD, "kat_encode_decode.npz"), **kat)

    # ---- losses --------------------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    B, C, H, W, M = 2, 6, 16, 16, 12
    logits = (rng.standard_normal((B, C, H, W)) * 3).astype(np.float32)
    gt = (rng.random((B, C, H, W)) ** 4).astype(np.float32)
    gt[rng.random((B, C, H, W)) > 0.98] = 1.0
    x = T(logits.copy()).requires_grad_(True)
    p = sigmoid_clamped(x.clone())
    p.retain_grad()
    loss = FocalLoss()(p, T(gt))
    loss.backward()
    focal = dict(logits=logits, gt=gt, prob=p.detach().numpy(), loss=loss.detach().numpy(),
                 dprob=p.grad.numpy(), dlogits=x.grad.numpy())
    l2, g2 = losses_np.neg_loss(focal["prob"], gt)
    assert abs(l2 - focal["loss"]) < 1e-5 * abs(focal["loss"]) and np.allclose(g2, focal["dprob"], rtol=1e-4, atol=1e-7)
    l3, g3 = losses_np.focal_with_logits(logits, gt)
    assert abs(l3 - focal["loss"]) < 1e-5 * abs(focal["loss"]) and np.allclose(g3, focal["dlogits"], rtol=1e-4, atol=1e-7)
    # no positives branch (losses.py:35-36)
    gt0 = np.minimum(gt, 0.9).astype(np.float32)
    p0 = T(focal["prob"].copy()).requires_grad_(True)
    loss0 = FocalLoss()(p0, T(gt0))
    loss0.backward()
    focal.update(gt0=gt0, loss0=loss0.detach().numpy(), dprob0=p0.grad.numpy())

    outp = (rng.standard_normal((B, 2, H, W))).astype(np.float32)
    ind = rng.integers(0, H * W, size=(B, M)).astype(np.int64)
    ind[0, 1] = ind[0, 0]   # duplicate index: gradients add
    mask = rng.random((B, M)) > 0.4
    tgt_ = rng.standard_normal((B, M, 2)).astype(np.float32)
    o = T(outp.copy()).requires_grad_(True)
    l = RegL1Loss()(o, T(mask), T(ind), T(tgt_))
    l.backward()
    lo, go = losses_np.reg_l1(outp, mask, ind, tgt_)
    assert abs(lo - l.item()) < 1e-5 and np.allclose(go, o.grad.numpy(), atol=1e-6)
    focal.update(r_out=outp, r_ind=ind, r_mask=mask, r_tgt=tgt_, r_loss=l.detach().numpy(), r_grad=o.grad.numpy())

    outk = rng.standard_normal((B, 34, H, W)).astype(np.float32)
    maskk = (rng.random((B, M, 34)) > 0.5).astype(np.float32)
    tgtk = rng.standard_normal((B, M, 34)).astype(np.float32)
    o = T(outk.copy()).requires_grad_(True)
    l = RegWeightedL1Loss()(o, T(maskk), T(ind), T(tgtk))
    l.backward()
    lo, go = losses_np.reg_l1(outk, maskk, ind, tgtk, per_channel=True)
    assert abs(lo - l.item()) < 1e-5 and np.allclose(go, o.grad.numpy(), atol=1e-6)
    focal.update(w_out=outk, w_mask=maskk, w_tgt=tgtk, w_loss=l.detach().numpy(), w_grad=o.grad.numpy())
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **focal)
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print(" ", f, os.path.getsize(os.path.join(GOLD, f)), "bytes")


if __name__ == "__main__":
    main()
