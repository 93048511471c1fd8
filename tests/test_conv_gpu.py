"""-m gpu parity: tcgen05 implicit-GEMM convolution, DCNv2 and the memory-bound layer ops.

Floating point: operands are bf16 (rounded identically for the kernel and the reference), the
accumulation is fp32 in both; the reference is plain PyTorch fp32 (TF32 disabled in conftest) on
the bf16-rounded operands.  Tolerance: |err| <= 2e-2 * max|ref| for bf16 outputs (one bf16 rounding
of the result = 2^-9 relative, plus summation-order noise), 2e-3 for fp32 outputs.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from centernet_pytorch_lightning_b200 import ops

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _check(got, ref, tol):
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


CONV_CASES = [
    # B, Ci, Co, H, W, k, stride, relu, residual
    (2, 64, 64, 32, 32, 3, 1, True, False),
    (2, 64, 64, 32, 32, 3, 1, True, True),
    (1, 16, 16, 64, 64, 3, 1, True, False),      # Ci=16: 4 taps per 64-wide K block
    (2, 16, 32, 64, 64, 3, 2, True, False),      # stride 2
    (2, 128, 128, 16, 16, 3, 1, True, True),
    (1, 512, 512, 16, 16, 3, 1, False, False),   # 4 N tiles, 72 K blocks
    (2, 256, 128, 16, 16, 1, 1, True, False),    # 1x1 (Root)
    (1, 448, 128, 8, 8, 1, 1, True, True),       # Root with odd concat width
    (3, 64, 80, 20, 28, 3, 1, False, False),     # ragged M (tail tile) and Co=80
    (1, 64, 64, 24, 24, 3, 1, True, True),       # 5 M tiles: the last cluster (2 or 4 CTAs) runs past the end
    (1, 128, 256, 40, 40, 3, 1, True, False),    # 13 M tiles x N=256: multicast weight tile, several tiles per cluster
    (1, 8, 16, 40, 40, 7, 1, True, False),       # stem geometry: 7x7 on a channel-padded input
    (1, 32, 27, 24, 24, 3, 1, False, False),     # Co=27 (offset/mask conv) -> NHWC fp32
    # row-window kernel (conv_rows.cu): W_out % 128 == 0, Ci <= 64
    (1, 16, 16, 8, 128, 3, 1, True, False),      # one strip, H < ring depth
    (4, 16, 16, 128, 128, 3, 1, True, False),    # several rows per CTA: ring reuse, strips end mid-range
    (2, 16, 32, 16, 256, 3, 2, True, False),     # stride 2: phase planes
    (1, 32, 64, 24, 256, 3, 2, True, False),
    (2, 64, 64, 10, 128, 3, 1, True, True),      # 8 chunk planes, residual
    (1, 64, 27, 6, 256, 3, 1, False, False),     # offset/mask conv geometry, two strips per row, fp32 out
    (1, 8, 16, 12, 128, 7, 1, True, False),      # stem: taps paired along kw (weights packed 7x8)
    (2, 8, 16, 40, 256, 7, 1, True, False),
    # footprint kernel, plain 3x3 mode by default (dcn_fp.cu: Ci >= 128, N <= 32, >= 4 tiles per SM): box staging, four
    # issuing groups with one partial accumulator each, summed in the epilogue
    (20, 128, 27, 64, 64, 3, 1, False, False),
    (20, 128, 24, 64, 64, 3, 1, True, True),     # bf16 output with residual through the same mode
]


@pytest.mark.parametrize("B,Ci,Co,H,W,k,stride,relu,residual", CONV_CASES)
def test_conv_matches_torch(cuda_dev, B, Ci, Co, H, W, k, stride, relu, residual):
    g = torch.Generator(device="cpu").manual_seed(Ci * 1000 + Co + k)
    x = _bf(torch.randn(B, Ci, H, W, generator=g)).to(cuda_dev)
    w = _bf(torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(cuda_dev)
    scale = (0.5 + torch.rand(Co, generator=g)).to(cuda_dev)
    shift = torch.randn(Co, generator=g).to(cuda_dev)
    pad = k // 2
    ref = F.conv2d(x, w, stride=stride, padding=pad) * scale[None, :, None, None] + shift[None, :, None, None]
    Ho, Wo = ref.shape[2:]
    res = None
    if residual:
        res = _bf(torch.randn(B, Co, Ho, Wo, generator=g)).to(cuda_dev)
        ref = ref + res
    if relu:
        ref = ref.relu()
    xn = ops.to_nhwc_bf16(x)
    w_kw = 8 if (Ci == 8 and k == 7) else 0      # the stem packs its filter 7x8 (see cnb_conv_desc.w_kw)
    wpk = ops.pack_conv_weights(w, kw_pad=w_kw or None)
    fp32_out = Co % 8 != 0
    y = ops.conv2d(xn, wpk, Co, k, stride, pad, scale, shift,
                   res=ops.to_nhwc_bf16(res) if residual else None, act=1 if relu else 0,
                   out_mode=2 if fp32_out else 0, w_kw=w_kw)
    torch.cuda.synchronize()
    got = y[..., :Co].permute(0, 3, 1, 2)
    _check(got, ref, 2e-3 if fp32_out else 2e-2)


@pytest.mark.parametrize("B,H,W,k", [(2, 20, 256, 7), (1, 9, 512, 7), (2, 12, 64, 7), (1, 16, 256, 3),
                                     (1, 8, 256, 5), (1, 7, 30, 7)])
def test_stem_space_to_depth(cuda_dev, B, H, W, k):
    """The 7x7 stem on pixel pairs (ops.pack_stem_s2d_weights / stem_s2d: 7x5 filter over [B,H,W/2,8], two output
    pixels per GEMM row) against F.conv2d; W/2 % 128 == 0 runs on the row-window kernel, other widths on v1."""
    g = torch.Generator(device="cpu").manual_seed(W + k)
    Co = 16
    x = _bf(torch.randn(B, 3, H, W, generator=g)).to(cuda_dev)
    w = _bf(torch.randn(Co, 3, k, k, generator=g) / (3 * k * k) ** 0.5).to(cuda_dev)
    scale = (0.5 + torch.rand(Co, generator=g)).to(cuda_dev)
    shift = torch.randn(Co, generator=g).to(cuda_dev)
    ref = (F.conv2d(x, w, padding=k // 2) * scale[None, :, None, None] + shift[None, :, None, None]).relu()
    wpk, geom = ops.pack_stem_s2d_weights(w)
    x4 = ops.to_nhwc_bf16(x, c_pad=4)
    assert x4.shape == (B, H, W, 4)
    assert torch.equal(x4[..., :3].float(), x.permute(0, 2, 3, 1)) and not x4[..., 3].any()
    y = ops.stem_s2d(x4, wpk, geom, Co, scale.repeat(2), shift.repeat(2), act=1)
    assert y.shape == (B, H, W, Co)
    _check(y.permute(0, 3, 1, 2), ref, 2e-2)


@pytest.mark.parametrize("B,Ci,Co,H,W,k", [(2, 64, 128, 24, 24, 3), (1, 16, 32, 8, 128, 3), (1, 256, 64, 16, 16, 1)])
def test_conv_dgrad_matches_autograd(cuda_dev, B, Ci, Co, H, W, k):
    """Data gradient of a stride-1 convolution = the forward kernel on the flipped / transposed filter
    (ops.pack_conv_weights_dgrad) against torch.autograd on the same bf16-rounded operands."""
    g = torch.Generator(device="cpu").manual_seed(Ci + Co + k)
    x = _bf(torch.randn(B, Ci, H, W, generator=g)).to(cuda_dev).requires_grad_(True)
    w = _bf(torch.randn(Co, Ci, k, k, generator=g) / (Co * k * k) ** 0.5).to(cuda_dev)
    dy = _bf(torch.randn(B, Co, H, W, generator=g)).to(cuda_dev)
    (ref,) = torch.autograd.grad(F.conv2d(x, w, padding=k // 2), x, dy)
    dx = ops.conv2d_dgrad(ops.to_nhwc_bf16(dy), ops.pack_conv_weights_dgrad(w), Ci, k, k // 2)
    assert dx.shape == (B, H, W, Ci)
    _check(dx.permute(0, 3, 1, 2), ref, 2e-2)


def test_conv_concat_slices(cuda_dev):
    """Reading / writing channel slices of wider NHWC buffers (Root's torch.cat without the copy)."""
    g = torch.Generator().manual_seed(7)
    B, H, W = 2, 16, 16
    xa = _bf(torch.randn(B, 64, H, W, generator=g)).to(cuda_dev)
    w = _bf(torch.randn(32, 64, 3, 3, generator=g) / 24).to(cuda_dev)
    wide_in = torch.zeros(B, H, W, 160, dtype=torch.bfloat16, device=cuda_dev)
    wide_in[..., 96:160] = ops.to_nhwc_bf16(xa)
    wide_out = torch.full((B, H, W, 96), 7.0, dtype=torch.bfloat16, device=cuda_dev)
    ops.conv2d(ops.View(wide_in, 64, 96), ops.pack_conv_weights(w), 32, 3, 1, 1, None, None, act=0,
               out=ops.View(wide_out, 32, 64))
    torch.cuda.synchronize()
    ref = F.conv2d(xa, w, padding=1)
    _check(wide_out[..., 64:96].permute(0, 3, 1, 2), ref, 2e-2)
    assert torch.all(wide_out[..., :64] == 7.0)      # untouched outside the slice


def test_conv_nchw_f32_head_output(cuda_dev):
    g = torch.Generator().manual_seed(9)
    B, H, W = 2, 32, 32
    x = _bf(torch.randn(B, 256, H, W, generator=g)).to(cuda_dev)
    w = _bf(torch.randn(80, 256, 1, 1, generator=g) / 16).to(cuda_dev)
    bias = torch.randn(80, generator=g).to(cuda_dev)
    y = ops.conv2d(ops.to_nhwc_bf16(x), ops.pack_conv_weights(w), 80, 1, 1, 0, None, bias, act=0, out_mode=1)
    torch.cuda.synchronize()
    assert y.shape == (B, 80, H, W) and y.dtype == torch.float32
    _check(y, F.conv2d(x, w, bias), 2e-3)
    ys = ops.conv2d(ops.to_nhwc_bf16(x), ops.pack_conv_weights(w), 80, 1, 1, 0, None, bias, act=2, out_mode=1)
    _check(ys, torch.sigmoid(F.conv2d(x, w, bias)), 2e-3)


@pytest.mark.parametrize("B,Ci,Co,H,W,sliced", [(2, 64, 64, 24, 24, False), (1, 128, 64, 16, 20, False),
                                                 (1, 256, 256, 8, 8, False),
                                                 (1, 128, 128, 20, 24, False),   # BN=128: two taps per stage, 18 K blocks
                                                 (2, 64, 24, 12, 40, False),     # Co not a multiple of 16
                                                 (2, 64, 96, 24, 32, False),     # BN = 96: 10 tensor-memory stages fit, 8 are used (a multiple of the 4 groups)
                                                 (2, 64, 64, 24, 32, False),     # 8 x 16 pixel tiles (W % 16 == 0)
                                                 (1, 128, 64, 20, 16, False),    # ... last tile row cut by the image
                                                 (1, 128, 128, 16, 48, False),   # two slabs, BN = 128
                                                 (1, 64, 64, 16, 32, True),      # ... on a channel slice
                                                 (1, 64, 64, 16, 24, True)])     # 16-byte-aligned channel slice of a
                                                                                 # wider tensor: 128-bit corner loads
def test_dcnv2_matches_torchvision(cuda_dev, B, Ci, Co, H, W, sliced):
    """DCN numerics are 'parity unpinned' by the reference (external extension); the pin is
    torchvision.ops.deform_conv2d on CPU fp32 with the same bf16-rounded operands."""
    from torchvision.ops import deform_conv2d
    g = torch.Generator().manual_seed(Ci + Co)
    x = _bf(torch.randn(B, Ci, H, W, generator=g))
    w = _bf(torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5)
    bias = torch.randn(Co, generator=g)
    om = torch.randn(B, 27, H, W, generator=g) * 1.5          # offsets up to a few pixels, some out of range
    om[:, :18, 0, 0] = 50.0                                     # far outside -> zero contribution
    o1, o2, m = torch.chunk(om, 3, dim=1)
    ref = deform_conv2d(x, torch.cat((o1, o2), 1), w, bias, padding=1, mask=torch.sigmoid(m))
    om_nhwc = torch.zeros(B, H, W, 32)
    om_nhwc[..., :27] = om.permute(0, 2, 3, 1)
    xn = ops.to_nhwc_bf16(x.to(cuda_dev))
    if sliced:   # x lives at channel offset 8 of a [B,H,W,Ci+24] buffer
        wide = torch.randn(B, H, W, Ci + 24, device=cuda_dev).to(torch.bfloat16)
        wide[..., 8:8 + Ci] = xn
        xn = ops.View(wide, Ci, 8)
    y = ops.dcnv2(xn, om_nhwc.to(cuda_dev).contiguous(),
                  ops.pack_conv_weights(w.to(cuda_dev)), Co, None, bias.to(cuda_dev), act=0)
    torch.cuda.synchronize()
    _check(y.permute(0, 3, 1, 2).cpu(), ref, 2e-2)


@pytest.mark.parametrize("B,Ci,Co,H", [(8, 64, 64, 128), (8, 128, 64, 64), (4, 256, 256, 32)])
def test_dcnv2_repeats_are_bit_identical(cuda_dev, B, Ci, Co, H):
    """The footprint kernel is a web of mbarrier hand-offs (table, boxes, A ring in tensor memory, weight slots, two
    accumulators) with several tiles per CTA at these sizes and no atomics: a lost ordering anywhere shows up as a
    run-to-run difference.  It found one: the setup warps handed the offset staging buffer back to the TMA engine behind
    a bar.sync while their shared-memory loads were still in flight (one run in ~500 built a tile row of the table from
    the next tile's offsets; fixed with fence.proxy.async before the barrier, then 0 differences in 32 000 runs of
    tools/dcn_race_hunt.py; compute-sanitizer racecheck is clean on both DCN kernels but had not seen this one)."""
    g = torch.Generator().manual_seed(11)
    x = ops.to_nhwc_bf16(torch.randn(B, Ci, H, H, generator=g).to(cuda_dev))
    om = (torch.randn(B, H, H, 32, generator=g) * 0.7).to(cuda_dev)
    wpk = ops.pack_conv_weights((torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).to(cuda_dev))
    bias = torch.zeros(Co, device=cuda_dev)
    first = ops.dcnv2(x, om, wpk, Co, None, bias, act=0).clone()
    for _ in range(60):
        assert torch.equal(ops.dcnv2(x, om, wpk, Co, None, bias, act=0), first)


def test_plain_conv_mode_repeats_are_bit_identical(cuda_dev):
    """Plain 3x3 mode of the footprint kernel (default for Ci >= 128, N <= 32, >= 4 tiles per SM): the four sampler groups
    issue their own MMAs, whose order at the tensor core varies from run to run -- each group accumulates into its own
    partial accumulator and the epilogue adds them in group order, so the result must not (tools/conv_race_hunt.py is
    the long version; the first version of the mode, one shared accumulator, differed on every run)."""
    g = torch.Generator().manual_seed(13)
    B, Ci, Co, H = 20, 128, 27, 64
    x = ops.to_nhwc_bf16(torch.randn(B, Ci, H, H, generator=g).to(cuda_dev))
    wpk = ops.pack_conv_weights((torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).to(cuda_dev))
    sc, sh = torch.ones(Co, device=cuda_dev), torch.zeros(Co, device=cuda_dev)
    first = ops.conv2d(x, wpk, Co, 3, 1, 1, sc, sh, act=0, out_mode=2).clone()
    for _ in range(100):
        assert torch.equal(ops.conv2d(x, wpk, Co, 3, 1, 1, sc, sh, act=0, out_mode=2), first)


def test_dcn_module_matches_torchvision(cuda_dev):
    """The drop-in `DCN.dcn_v2.DCN` module end to end (offset conv + sampler + GEMM) with non-zero
    conv_offset_mask weights (zero init would make DCN == 0.5 * conv)."""
    from torchvision.ops import deform_conv2d
    from centernet_pytorch_lightning_b200.DCN.dcn_v2 import DCN
    torch.manual_seed(5)
    m = DCN(64, 64, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)
    m.conv_offset_mask.weight.data.normal_(0, 0.05)
    m.conv_offset_mask.bias.data.uniform_(-1, 1)
    m.bias.data.normal_()
    x = _bf(torch.randn(2, 64, 20, 20))
    with torch.no_grad():
        wq, omq = _bf(m.weight), _bf(m.conv_offset_mask.weight)
        om = F.conv2d(x, omq, m.conv_offset_mask.bias, padding=1)
        o1, o2, mk = torch.chunk(om, 3, dim=1)
        ref = deform_conv2d(x, torch.cat((o1, o2), 1), wq, m.bias, padding=1, mask=torch.sigmoid(mk))
        got = m.to(cuda_dev).eval()(x.to(cuda_dev))
    _check(got.cpu(), ref, 3e-2)


@pytest.mark.parametrize("C,H,W,f", [(64, 64, 64, 2), (64, 32, 48, 4), (128, 64, 80, 2), (256, 72, 64, 2)])
def test_upsample_tile_kernel(cuda_dev, C, H, W, f):
    """Depthwise ConvTranspose2d(2f, stride f, pad f/2) + skip add at the sizes that take the cp.async tile kernel
    (outputs of 128 rows and more; pose_dla_dcn.py:466-488), incl. tiles cut by the right / bottom image border."""
    g = torch.Generator().manual_seed(C + f)
    x = _bf(torch.randn(2, C, H, W, generator=g)).to(cuda_dev)
    w = torch.randn(C, 1, 2 * f, 2 * f, generator=g).to(cuda_dev)
    add = _bf(torch.randn(2, C, H * f, W * f, generator=g)).to(cuda_dev)
    ref = F.conv_transpose2d(x, w, stride=f, padding=f // 2, groups=C)
    wt = ops.relayout_dw_weights(w, f)
    got = ops.dw_deconv_up(ops.to_nhwc_bf16(x), wt, f, add=ops.to_nhwc_bf16(add))
    _check(got.permute(0, 3, 1, 2), ref + add, 1e-2)
    got = ops.dw_deconv_up(ops.to_nhwc_bf16(x), wt, f)
    _check(got.permute(0, 3, 1, 2), ref, 1e-2)


def test_maxpool_and_upsample(cuda_dev):
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(2, 64, 16, 24, generator=g)).to(cuda_dev)
    xn = ops.to_nhwc_bf16(x)
    got = ops.maxpool2d(xn, 2)
    assert torch.equal(got.permute(0, 3, 1, 2).float(), F.max_pool2d(x, 2, 2))
    for f in (2, 4):
        w = torch.randn(64, 1, 2 * f, 2 * f, generator=g).to(cuda_dev)
        add = _bf(torch.randn(2, 64, 16 * f, 24 * f, generator=g)).to(cuda_dev)
        ref = F.conv_transpose2d(x, w, stride=f, padding=f // 2, groups=64) + add
        got = ops.dw_deconv_up(xn, ops.relayout_dw_weights(w, f), f, add=ops.to_nhwc_bf16(add))
        _check(got.permute(0, 3, 1, 2), ref, 1e-2)
        got = ops.dw_deconv_up(xn, ops.relayout_dw_weights(w, f), f)
        _check(got.permute(0, 3, 1, 2), ref - add, 1e-2)


def test_layout_roundtrip(cuda_dev):
    x = _bf(torch.randn(2, 3, 10, 12)).to(cuda_dev)
    xn = ops.to_nhwc_bf16(x, c_pad=8)
    assert xn.shape == (2, 10, 12, 8)
    assert torch.equal(xn[..., :3].permute(0, 3, 1, 2).float(), x) and torch.all(xn[..., 3:] == 0)
    back = ops.to_nchw_f32(ops.View(xn, 3, 0))
    assert torch.equal(back, x)


def test_dense_deconv_and_padded_maxpool(cuda_dev):
    """ConvTranspose2d(4, stride 2, pad 1) as one 3x3 conv + pixel shuffle (resnet_dcn.py:212-220), and the
    ResNet stem's MaxPool2d(3, 2, 1)."""
    g = torch.Generator().manual_seed(21)
    for B, Ci, Co, H, W in [(2, 64, 32, 12, 20), (1, 128, 64, 8, 8)]:
        x = _bf(torch.randn(B, Ci, H, W, generator=g)).to(cuda_dev)
        w = _bf(torch.randn(Ci, Co, 4, 4, generator=g) / (Ci * 4) ** 0.5).to(cuda_dev)
        scale = (0.5 + torch.rand(Co, generator=g)).to(cuda_dev)
        shift = torch.randn(Co, generator=g).to(cuda_dev)
        ref = F.conv_transpose2d(x, w, stride=2, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
        got = ops.deconv4x4s2(ops.to_nhwc_bf16(x), ops.pack_deconv4x4s2_weights(w), Co, scale.repeat(4).contiguous(),
                              shift.repeat(4).contiguous(), act=1)
        torch.cuda.synchronize()
        assert got.shape == (B, 2 * H, 2 * W, Co)
        _check(got.permute(0, 3, 1, 2), ref.relu(), 2e-2)
    x = _bf(torch.randn(2, 64, 17, 30, generator=g)).to(cuda_dev)
    got = ops.maxpool2d_pad(ops.to_nhwc_bf16(x), 3, 2, 1)
    assert torch.equal(got.permute(0, 3, 1, 2).float(), F.max_pool2d(x, 3, 2, 1))


@pytest.mark.parametrize("heads,B,H,W", [({"heatmap": 80, "width_height": 2, "regression": 2}, 2, 32, 48),
                                          ({"heatmap": 1, "width_height": 2, "regression": 2, "heatmap_keypoints": 17,
                                            "keypoints": 34, "heatmap_keypoints_offset": 2}, 1, 40, 24),
                                          ({"heatmap": 80, "width_height": 2, "regression": 2}, 3, 128, 128)])
def test_fused_center_head_matches_torch(cuda_dev, heads, B, H, W):
    """csrc/head_fused.cu (3x3 -> ReLU -> 1x1 with the intermediate in tensor memory) vs PyTorch fp32 on bf16-rounded
    operands, and vs the two-GEMM path (CNB_HEAD_FUSED=0).  The intermediate is rounded to bf16 on both GPU paths:
    <= 1e-2 of the largest output vs fp32; the two GPU paths agree to fp32 accumulation order."""
    import os
    from centernet_pytorch_lightning_b200.models.heads import CenterHead
    from centernet_pytorch_lightning_b200.utils.synthetic import randomize_
    torch.manual_seed(3)
    head = CenterHead(heads, 64, 256).eval()
    randomize_(head.state_dict(), 17)
    with torch.no_grad():
        for p in head.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(B, 64, H, W, generator=torch.Generator().manual_seed(5)))
    with torch.no_grad():
        ref = {n: getattr(head, n).fc(x) for n in heads}       # plain nn.Sequential forward on the CPU
        head = head.to(cuda_dev)
        xg = ops.to_nhwc_bf16(x.to(cuda_dev))
        got = head(xg, sigmoid=("heatmap",))
        os.environ["CNB_HEAD_FUSED"] = "0"
        try:
            two = head(xg, sigmoid=("heatmap",))
        finally:
            os.environ.pop("CNB_HEAD_FUSED")
    torch.cuda.synchronize()
    for n in heads:
        want = torch.sigmoid(ref[n]) if n == "heatmap" else ref[n]
        assert got[n].shape == want.shape and got[n].dtype == torch.float32
        err = (got[n].cpu() - want).abs().max().item() / (want.abs().max().item() + 1e-12)
        dif = (got[n] - two[n]).abs().max().item() / (want.abs().max().item() + 1e-12)
        print(f"{n}: fused vs fp32 {err:.2e}, fused vs two-GEMM path {dif:.2e}")
        assert err <= 1e-2 and dif <= 2e-3
