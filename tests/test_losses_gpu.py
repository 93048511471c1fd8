"""-m gpu parity: fused loss kernels vs the CPU oracle and the reference-generated fixtures.
Tolerance: 1e-4 relative on the loss, 1e-4 relative / 1e-7 absolute on gradients (fp32 reductions
in a different order; north_star's floating-point bound is 1e-4)."""
import os

import numpy as np
import pytest
import torch

from centernet_pytorch_lightning_b200.utils.decode import sigmoid_clamped
from centernet_pytorch_lightning_b200.utils.losses import (FocalLoss, RegL1Loss, RegWeightedL1Loss,
                                                           focal_loss_with_logits)
from oracle import losses_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_focal_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "losses.npz"))
    x = _t(g["logits"], cuda_dev).requires_grad_(True)
    p = sigmoid_clamped(x)
    p.retain_grad()
    loss = FocalLoss()(p, _t(g["gt"], cuda_dev))
    loss.backward()
    np.testing.assert_allclose(p.detach().cpu().numpy(), g["prob"], rtol=1e-6, atol=1e-9)
    assert abs(loss.item() - g["loss"]) <= 1e-4 * abs(g["loss"])
    np.testing.assert_allclose(p.grad.cpu().numpy(), g["dprob"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["dlogits"], rtol=2e-4, atol=1e-7)
    # num_pos == 0 branch (losses.py:35-36)
    p0 = _t(g["prob"], cuda_dev).requires_grad_(True)
    l0 = FocalLoss()(p0, _t(g["gt0"], cuda_dev))
    l0.backward()
    assert abs(l0.item() - g["loss0"]) <= 1e-4 * abs(g["loss0"])
    np.testing.assert_allclose(p0.grad.cpu().numpy(), g["dprob0"], rtol=1e-4, atol=1e-7)


def test_focal_fused_logits_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "losses.npz"))
    x = _t(g["logits"], cuda_dev).requires_grad_(True)
    loss = focal_loss_with_logits(x, _t(g["gt"], cuda_dev))
    (2.0 * loss).backward()      # upstream gradient != 1
    assert abs(loss.item() - g["loss"]) <= 1e-4 * abs(g["loss"])
    np.testing.assert_allclose(x.grad.cpu().numpy(), 2.0 * g["dlogits"], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("shape", [(16, 80, 128, 128), (3, 7, 33, 31)])
def test_focal_large_vs_oracle(cuda_dev, shape):
    """config 3 per-GPU size (16 x 80 x 128 x 128 = 21 M elements) and an odd-sized tail case."""
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal(shape) * 2 - 2).astype(np.float32)
    gt = (rng.random(shape) ** 6).astype(np.float32)
    gt[rng.random(shape) > 0.999] = 1.0
    ref_loss, ref_grad = losses_np.focal_with_logits(logits, gt)
    x = _t(logits, cuda_dev).requires_grad_(True)
    loss = focal_loss_with_logits(x, _t(gt, cuda_dev))
    loss.backward()
    assert abs(loss.item() - ref_loss) <= 1e-4 * abs(ref_loss)
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref_grad, rtol=3e-4, atol=1e-9)
    # determinism: fixed-order reduction
    loss2 = focal_loss_with_logits(_t(logits, cuda_dev), _t(gt, cuda_dev))
    assert loss2.item() == loss.item()


def test_reg_l1_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "losses.npz"))
    o = _t(g["r_out"], cuda_dev).requires_grad_(True)
    l = RegL1Loss()(o, _t(g["r_mask"], cuda_dev), _t(g["r_ind"], cuda_dev), _t(g["r_tgt"], cuda_dev))
    (3.0 * l).backward()
    assert abs(l.item() - g["r_loss"]) <= 1e-5
    np.testing.assert_allclose(o.grad.cpu().numpy(), 3.0 * g["r_grad"], rtol=1e-5, atol=1e-7)
    o = _t(g["w_out"], cuda_dev).requires_grad_(True)
    l = RegWeightedL1Loss()(o, _t(g["w_mask"], cuda_dev), _t(g["r_ind"], cuda_dev), _t(g["w_tgt"], cuda_dev))
    l.backward()
    assert abs(l.item() - g["w_loss"]) <= 1e-5
    np.testing.assert_allclose(o.grad.cpu().numpy(), g["w_grad"], rtol=1e-5, atol=1e-7)


def test_reg_l1_config3_size(cuda_dev):
    rng = np.random.default_rng(4)
    B, C, H, W, M = 16, 2, 128, 128, 128
    out = rng.standard_normal((B, C, H, W)).astype(np.float32)
    ind = rng.integers(0, H * W, size=(B, M)).astype(np.int64)
    mask = rng.random((B, M)) > 0.5
    tgt = rng.standard_normal((B, M, C)).astype(np.float32)
    ref_l, ref_g = losses_np.reg_l1(out, mask, ind, tgt)
    o = _t(out, cuda_dev).requires_grad_(True)
    l = RegL1Loss()(o, _t(mask, cuda_dev), _t(ind, cuda_dev), _t(tgt, cuda_dev))
    l.backward()
    assert abs(l.item() - ref_l) <= 1e-4 * abs(ref_l)
    np.testing.assert_allclose(o.grad.cpu().numpy(), ref_g, rtol=1e-4, atol=1e-8)
    # empty mask: loss 0 / (0 + 1e-4) = 0, gradient 0
    l = RegL1Loss()(_t(out, cuda_dev), _t(np.zeros_like(mask), cuda_dev), _t(ind, cuda_dev), _t(tgt, cuda_dev))
    assert l.item() == 0.0


def test_loss_wrappers_cast_half_inputs_and_drop_bad_indices(cuda_dev):
    """(a) fp16 / bf16 predictions (AMP) are cast to fp32 instead of being reinterpreted -- same loss as the
    fp32 call on the rounded values, gradient returned in the input dtype; (b) an index outside [0, H*W) is dropped
    (no out-of-bounds read / atomicAdd), the loss equals the one without that object."""
    from centernet_pytorch_lightning_b200.utils.losses import FocalLoss, RegL1Loss
    g = torch.Generator().manual_seed(0)
    pred = torch.rand(2, 3, 16, 16, generator=g).clamp(1e-3, 1 - 1e-3)
    gt = torch.rand(2, 3, 16, 16, generator=g) ** 4
    gt[0, 1, 3, 4] = 1.0
    for dt in (torch.float16, torch.bfloat16):
        p16 = pred.to(dt).to(cuda_dev).requires_grad_(True)
        p32 = pred.to(dt).float().to(cuda_dev).requires_grad_(True)
        l16, l32 = FocalLoss()(p16, gt.to(cuda_dev)), FocalLoss()(p32, gt.to(cuda_dev))
        l16.backward()
        l32.backward()
        assert torch.equal(l16, l32) and p16.grad.dtype == dt
        assert torch.allclose(p16.grad.float(), p32.grad, rtol=1e-2, atol=1e-6)
    out = torch.randn(2, 2, 8, 8, generator=g).to(cuda_dev).requires_grad_(True)
    ind = torch.randint(0, 64, (2, 5), generator=g)
    mask = torch.ones(2, 5, dtype=torch.bool)
    tgt = torch.randn(2, 5, 2, generator=g)
    bad = ind.clone()
    bad[1, 2] = 64 + 1000                      # out of range
    good_mask = mask.clone()
    good_mask[1, 2] = False
    l_bad = RegL1Loss()(out, mask.to(cuda_dev), bad.to(cuda_dev), tgt.to(cuda_dev))
    l_bad.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(l_bad) and torch.isfinite(out.grad).all()
    num_ref = RegL1Loss()(out.detach(), good_mask.to(cuda_dev), ind.to(cuda_dev), tgt.to(cuda_dev)) * (good_mask.sum() * 2 + 1e-4)
    assert torch.allclose(l_bad * (good_mask.sum() * 2 + 1e-4), num_ref, rtol=1e-5)
