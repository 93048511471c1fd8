"""-m gpu parity: CUDA decode kernels (through the C ABI via the Python boundary) vs the CPU oracle.

Bar: bit-exact (integer/index work and the float arithmetic is replicated op for op).
"""
import numpy as np
import pytest
import torch

from centernet_pytorch_lightning_b200.decode import ctdet_decode, multi_pose_decode
from centernet_pytorch_lightning_b200.utils import synthetic
from oracle import decode_np

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _run_ctdet(dev, heat, wh, reg, K):
    out = ctdet_decode(_t(heat, dev), _t(wh, dev), _t(reg, dev), K=K)
    torch.cuda.synchronize()
    return out.cpu().numpy()


CT_SHAPES = [
    # B, C, H, W, K
    (2, 80, 64, 64, 100),     # config 1 head maps (res_18 @256)
    (2, 80, 128, 128, 100),   # config 2 geometry
    (1, 3, 40, 56, 40),       # ragged, K=40 (the _topk default)
    (3, 5, 17, 23, 100),      # W % 4 != 0 -> scalar path
    (1, 1, 8, 16, 100),       # tiny
    (2, 2, 136, 136, 100),    # (512|31)+1 = 544 padded input -> 136x136 maps
    (1, 4, 64, 512, 100),     # wide
]


@pytest.mark.parametrize("B,C,H,W,K", CT_SHAPES)
@pytest.mark.parametrize("kind", ["uniform", "bumps"])
def test_ctdet_matches_oracle(cuda_dev, B, C, H, W, K, kind):
    heat, wh, reg = synthetic.ctdet_maps(B, C, H, W, seed=B * 131 + C, kind=kind)
    for r in (reg, None):
        ref = decode_np.ctdet_decode(heat, wh, r, K=K)
        got = _run_ctdet(cuda_dev, heat, wh, r, K)
        assert got.shape == (B, K, 6)
        assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"


@pytest.mark.parametrize("rpt", ["1", "4", "16"])
def test_ctdet_banding_variants(cuda_dev, monkeypatch, rpt):
    """Different band splits (rows per CTA) must not change the result."""
    monkeypatch.setenv("CNB_DECODE_RPT", rpt)
    heat, wh, reg = synthetic.ctdet_maps(2, 6, 128, 128, seed=3)
    ref = decode_np.ctdet_decode(heat, wh, reg)
    assert np.array_equal(_run_ctdet(cuda_dev, heat, wh, reg, 100), ref)


def test_ctdet_sparse_zero_fill(cuda_dev):
    """Fewer than K positive peaks: remaining rows are zero-score entries in flat-index order
    (the reference's tests/test_sample_encode_decode.py situation: 2 objects, 98 filler rows)."""
    rng = np.random.default_rng(0)
    B, C, H, W = 2, 4, 32, 32
    heat = np.zeros((B, C, H, W), np.float32)
    for b in range(B):
        for _ in range(7):
            heat[b, rng.integers(C), rng.integers(1, H - 1), rng.integers(1, W - 1)] = rng.random() * 0.9 + 0.05
    heat[1, 0, 0, 0] = 0.5   # a positive peak at flat index 0 must not be reused as filler
    wh = rng.random((B, 2, H, W)).astype(np.float32) * 10
    reg = rng.random((B, 2, H, W)).astype(np.float32)
    ref = decode_np.ctdet_decode(heat, wh, reg)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    assert np.array_equal(got, ref)


def test_ctdet_plateaus_and_ties(cuda_dev):
    """Equal scores: plateaus survive the NMS (hmax == heat) and ties order by flat index."""
    rng = np.random.default_rng(1)
    B, C, H, W = 1, 3, 24, 32
    heat = (rng.integers(0, 6, size=(B, C, H, W)) / 8.0).astype(np.float32)   # heavy ties
    wh = rng.random((B, 2, H, W)).astype(np.float32)
    reg = rng.random((B, 2, H, W)).astype(np.float32)
    ref = decode_np.ctdet_decode(heat, wh, reg)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    assert np.array_equal(got, ref)


def test_ctdet_constant_plane_slow_path(cuda_dev):
    """Every pixel is a peak (constant planes) -> candidate overflow -> exact bitwise-search path."""
    B, C, H, W = 1, 2, 64, 64
    heat = np.full((B, C, H, W), 0.25, np.float32)
    heat[0, 1] = 0.75
    wh = np.ones((B, 2, H, W), np.float32)
    reg = np.zeros((B, 2, H, W), np.float32)
    ref = decode_np.ctdet_decode(heat, wh, reg)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    assert np.array_equal(got, ref)


def test_ctdet_top100_all_in_one_plane(cuda_dev):
    """Worst case for the two-level selection: the whole top-K lives in one class plane."""
    heat, wh, reg = synthetic.ctdet_maps(1, 8, 64, 64, seed=11)
    heat *= 0.5
    heat[0, 5] += 0.5
    ref = decode_np.ctdet_decode(heat, wh, reg)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    assert np.array_equal(got, ref)


def test_ctdet_golden_from_reference(cuda_dev):
    """Fixtures produced by the unmodified reference (oracle/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctdet_decode.npz"))
    for tag in ("a", "b"):
        heat, wh, reg, ref = g[f"heat_{tag}"], g[f"wh_{tag}"], g[f"reg_{tag}"], g[f"out_{tag}"]
        got = _run_ctdet(cuda_dev, heat, wh, reg, ref.shape[1])
        assert np.array_equal(got, ref)


def test_ctdet_full_size_properties(cuda_dev):
    """BASELINE config 2 size (B=32, 80x128x128): size-independent properties + oracle on 2 images."""
    B = 32
    heat, wh, reg = synthetic.ctdet_maps(B, 80, 128, 128, seed=2024)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    scores = got[:, :, 4]
    assert np.all(np.diff(scores, axis=1) <= 0), "scores must be sorted descending"
    # every reported score is the heat value at the reported (class, y, x) and is a 3x3 local max
    cls = got[:, :, 5].astype(np.int64)
    cx = (got[:, :, 0] + got[:, :, 2]) / 2
    cy = (got[:, :, 1] + got[:, :, 3]) / 2
    for b in (0, 13, 31):
        ref = decode_np.ctdet_decode(heat[b:b + 1], wh[b:b + 1], reg[b:b + 1])
        assert np.array_equal(got[b:b + 1], ref)
    # permutation invariance over the batch (images are independent units)
    perm = np.random.default_rng(0).permutation(B)
    got_p = _run_ctdet(cuda_dev, heat[perm], wh[perm], reg[perm], 100)
    assert np.array_equal(got_p, got[perm])
    assert np.isfinite(cx).all() and np.isfinite(cy).all() and (cls >= 0).all() and (cls < 80).all()


def test_ctdet_saturated_network_like_maps(cuda_dev):
    """Heat maps as a random-weight network emits them: sigmoid of large logits -> big plateaus of exactly
    1.0 (every plateau pixel survives the NMS) next to exact zeros; the top-K is then decided by flat index."""
    rng = np.random.default_rng(7)
    B, C, H, W = 3, 80, 128, 128
    logits = rng.standard_normal((B, C, H // 8, W // 8)).astype(np.float32) * 30.0
    logits = np.kron(logits, np.ones((8, 8), np.float32)) + rng.standard_normal((B, C, H, W)).astype(np.float32)
    heat = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32)
    assert (heat == 1.0).sum() > 10000
    wh = rng.random((B, 2, H, W)).astype(np.float32) * 30
    reg = rng.random((B, 2, H, W)).astype(np.float32)
    ref = decode_np.ctdet_decode(heat, wh, reg)
    got = _run_ctdet(cuda_dev, heat, wh, reg, 100)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("ctas", ["1", "3", "7"])
def test_ctdet_cta_ranges_span_images(cuda_dev, monkeypatch, ctas):
    """Few persistent CTAs: every CTA's chunk range crosses image (group) boundaries."""
    monkeypatch.setenv("CNB_DECODE_CTAS", ctas)
    for kind in ("uniform", "bumps"):
        heat, wh, reg = synthetic.ctdet_maps(5, 6, 64, 64, seed=17, kind=kind)
        ref = decode_np.ctdet_decode(heat, wh, reg)
        assert np.array_equal(_run_ctdet(cuda_dev, heat, wh, reg, 100), ref)


def test_ctdet_large_k(cuda_dev):
    heat, wh, reg = synthetic.ctdet_maps(2, 4, 64, 64, seed=23)
    for K in (1, 257, 512):
        ref = decode_np.ctdet_decode(heat, wh, reg, K=K)
        assert np.array_equal(_run_ctdet(cuda_dev, heat, wh, reg, K), ref)


MP_SHAPES = [(2, 17, 128, 128, 100), (3, 17, 64, 96, 100), (1, 5, 30, 41, 40), (2, 17, 136, 136, 100)]


@pytest.mark.parametrize("B,J,H,W,K", MP_SHAPES)
def test_multi_pose_matches_oracle(cuda_dev, B, J, H, W, K):
    heat, wh, kps, reg, hm_hp, hpo = synthetic.multi_pose_maps(B, J, H, W, seed=J + H)
    dev = cuda_dev
    for r, o in ((reg, hpo), (None, None)):
        ref = decode_np.multi_pose_decode(heat, wh, kps, r, hm_hp, o, K=K)
        out = multi_pose_decode(_t(heat, dev), _t(wh, dev), _t(kps, dev), _t(r, dev), _t(hm_hp, dev), _t(o, dev), K=K)
        got = out.cpu().numpy()
        assert got.shape == (B, K, 3 * J + 6)
        assert np.array_equal(got, ref), f"mismatches {(got != ref).sum()} max {np.abs(got - ref).max()}"


def test_multi_pose_low_scores_masked(cuda_dev):
    """hm_hp below the 0.1 threshold everywhere -> every joint falls back to the regressed location."""
    heat, wh, kps, reg, hm_hp, hpo = synthetic.multi_pose_maps(2, 17, 64, 64, seed=5)
    hm_hp = (hm_hp * 0.09).astype(np.float32)
    dev = cuda_dev
    ref = decode_np.multi_pose_decode(heat, wh, kps, reg, hm_hp, hpo)
    got = multi_pose_decode(_t(heat, dev), _t(wh, dev), _t(kps, dev), _t(reg, dev), _t(hm_hp, dev), _t(hpo, dev)).cpu().numpy()
    assert np.array_equal(got, ref)
    assert np.all(got[:, :, 40:] == 0)


def test_multi_pose_golden_from_reference(cuda_dev):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "multi_pose_decode.npz"))
    dev = cuda_dev
    args = [_t(g[k], dev) for k in ("heat", "wh", "kps", "reg", "hm_hp", "hp_offset")]
    got = multi_pose_decode(*args).cpu().numpy()
    assert np.array_equal(got, g["out"])


def test_multi_pose_requires_hm_hp(cuda_dev):
    heat, wh, kps, reg, hm_hp, hpo = synthetic.multi_pose_maps(1, 17, 16, 16, seed=5)
    dev = cuda_dev
    with pytest.raises(NameError):   # multi_pose.py:94
        multi_pose_decode(_t(heat, dev), _t(wh, dev), _t(kps, dev), _t(reg, dev), None, None)




def test_decode_primitives_by_name(cuda_dev):
    """`_nms`, `_topk`, `_topk_channel`, `_gather_feat`, `_transpose_and_gather_feat` (utils/decode.py:5-63) as
    stand-alone kernels vs the numpy oracle on tie-free maps: values, indices, classes bit-exact."""
    from centernet_pytorch_lightning_b200.utils.decode import (_gather_feat, _nms, _topk, _topk_channel,
                                                               _transpose_and_gather_feat)
    heat, wh, _ = synthetic.ctdet_maps(2, 7, 24, 40, seed=31)
    th = torch.from_numpy(heat).to(cuda_dev)
    n = _nms(th)
    assert np.array_equal(n.cpu().numpy(), decode_np.nms(heat))
    for K in (1, 40, 100):
        got = _topk(n, K=K)
        want = decode_np.topk(decode_np.nms(heat), K)
        for g, w, name in zip(got, want, ("score", "inds", "clses", "ys", "xs")):
            assert np.array_equal(g.cpu().numpy().astype(np.float64), np.asarray(w).astype(np.float64)), (K, name)
        assert got[1].dtype == torch.int64 and got[2].dtype == torch.int32
    gc = _topk_channel(n, K=30)
    wc = decode_np.topk_channel(decode_np.nms(heat), 30)
    for g, w in zip(gc, wc):
        assert np.array_equal(g.cpu().numpy().astype(np.float64), np.asarray(w).astype(np.float64))
    ind = torch.from_numpy(np.random.default_rng(0).integers(0, 24 * 40, size=(2, 17))).to(cuda_dev)
    tw = torch.from_numpy(wh).to(cuda_dev)
    got = _transpose_and_gather_feat(tw, ind).cpu().numpy()
    assert np.array_equal(got, decode_np.transpose_and_gather(wh, ind.cpu().numpy()))
    feat = tw.permute(0, 2, 3, 1).reshape(2, -1, 2).contiguous()
    assert np.array_equal(_gather_feat(feat, ind).cpu().numpy(), got)


def test_ctdet_decode_host_entry(cuda_dev):
    """`cnb_ctdet_decode_host`: host buffers in, host buffer out (H2D, decode, D2H inside the call)."""
    import ctypes
    from centernet_pytorch_lightning_b200 import _lib
    heat, wh, reg = synthetic.ctdet_maps(3, 80, 32, 32, seed=41)
    out = np.empty((3, 100, 6), np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    _lib.check(_lib.lib().cnb_ctdet_decode_host(P(heat), P(wh), P(reg), P(out), 3, 80, 32, 32, 100), "cnb_ctdet_decode_host")
    assert np.array_equal(out, decode_np.ctdet_decode(heat, wh, reg))
