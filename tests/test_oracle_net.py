"""CPU: the functional network oracle (oracle/net_torch.py) against the unmodified reference modules
(only where /root/reference exists), and against a committed fixture of its own output (everywhere)."""
import os

import numpy as np
import pytest
import torch

from oracle import net_torch, ref_shim
from centernet_pytorch_lightning_b200.utils.synthetic import randomize_

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HEADS = {"heatmap": 80, "width_height": 2, "regression": 2}


def _seeded_state(seed=0):
    from centernet_pytorch_lightning_b200.models import create_model
    from centernet_pytorch_lightning_b200.models.heads import CenterHead
    torch.manual_seed(seed)
    sd = {k: v.clone() for k, v in create_model("dla_34").state_dict().items()}
    hd = {k: v.clone() for k, v in CenterHead(HEADS, 64, 256).state_dict().items()}
    net_torch.randomize_(sd, seed)
    net_torch.randomize_(hd, seed + 1)
    return sd, hd


def test_state_dict_schema_matches_reference():
    if not ref_shim.available():
        pytest.skip("reference sources not present on this machine")
    from centernet_pytorch_lightning_b200.models import create_model
    a, b = create_model("dla_34").state_dict(), ref_shim.ref_dlaseg().state_dict()
    assert list(a.keys()) == list(b.keys()) and len(a) == 386
    assert all(a[k].shape == b[k].shape for k in a)
    assert all(torch.equal(a[k], b[k]) for k in a if ".up_" in k)     # bilinear init (fill_up_weights)


def test_net_oracle_matches_reference_modules():
    if not ref_shim.available():
        pytest.skip("reference sources not present on this machine")
    ref_shim.install()
    from CenterNet.models.heads import CenterHead as RefHead
    sd, hd = _seeded_state(3)
    ref = ref_shim.ref_dlaseg().eval()
    ref.load_state_dict(sd)
    rh = RefHead(HEADS, 64, 256).eval()
    rh.load_state_dict(hd)
    x = torch.rand(1, 3, 96, 128, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = ref(x)[-1]
        got = net_torch.dla34_seg_forward(sd, x)
        assert torch.equal(got, want)
        wh, gh = rh(want), net_torch.center_head_forward(hd, got, HEADS)
        assert all(torch.equal(wh[k], gh[k]) for k in HEADS)


def test_net_oracle_fixture():
    """Seeded weights are reproducible (CPU generator), so a digest of the oracle output pins it here and
    on the GPU box (same torch build)."""
    sd, hd = _seeded_state(3)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = net_torch.dla34_seg_forward(sd, x)
    g = np.load(os.path.join(GOLD, "dla34_oracle.npz"))
    np.testing.assert_allclose(y.numpy(), g["feat"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("arch,num_layers", [("res", 18), ("resdcn", 18), ("resdcn", 50)])
def test_resnet_schema_and_oracle_match_reference(arch, num_layers):
    """ResNet backbones (msra_resnet.py / resnet_dcn.py): same state-dict schema as the reference modules and the
    functional oracle reproduces their eval forward exactly."""
    if not ref_shim.available():
        pytest.skip("reference sources not present on this machine")
    from centernet_pytorch_lightning_b200.models import create_model
    mine = create_model(f"{arch}_{num_layers}")
    ref = (ref_shim.ref_msra_resnet if arch == "res" else ref_shim.ref_resnet_dcn)(num_layers).eval()
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert mine.out_channels == ref.out_channels
    sd = {k: v.clone() for k, v in a.items()}
    net_torch.randomize_(sd, 7)
    ref.load_state_dict(sd)
    x = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = ref(x)[-1]
        got = net_torch.pose_resnet_forward(sd, x, num_layers, arch == "resdcn")
    assert torch.equal(got, want)


@pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout")
def test_train_mode_oracle_equals_reference_dlaseg():
    """`net_torch.training()` (batch-statistics BatchNorm, gradients to every tensor of the state dict) vs the
    reference DLASeg in `.train()`: forward bit-equal, every parameter gradient within 1e-4, running statistics equal,
    and the parameters the reference leaves without gradient (`level{3,4}.project`, dead in Tree.forward :254-255) get
    none here either.  This pins the oracle tests/test_train_gpu.py compares the CUDA training path against."""
    ref = ref_shim.ref_dlaseg().train()
    randomize_(ref.state_dict(), 3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    y_ref = ref(x)[-1]
    (y_ref ** 2).sum().backward()
    with net_torch.training():
        y = net_torch.dla34_seg_forward(sd, x)
    (y ** 2).sum().backward()
    assert torch.equal(y, y_ref)
    dead = []
    for k, p in ref.named_parameters():
        g = sd[k].grad
        if p.grad is None:
            dead.append(k)
            assert g is None or float(g.abs().max()) == 0.0, k
        else:
            assert torch.allclose(g, p.grad, rtol=1e-4, atol=1e-6), k
    assert sorted(dead) == sorted(f"base.level{l}.project.{s}" for l in (3, 4) for s in ("0.weight", "1.weight", "1.bias"))
    for k, b in ref.named_buffers():
        if not k.endswith("num_batches_tracked"):
            assert torch.allclose(sd[k], b), k


@pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout")
def test_hourglass_oracle_equals_reference():
    """`net_torch.hourglass_forward` vs the reference HourglassNet (large_hourglass.py:322-343): both stack outputs
    bit-equal, and this package's containers carry the same 960 state-dict entries."""
    ref_shim.install()
    from CenterNet.models.backbones.large_hourglass import HourglassNet
    from centernet_pytorch_lightning_b200.models import create_model
    ref = HourglassNet().eval()
    randomize_(ref.state_dict(), 4)
    mine = create_model("hourglass")
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert mine.out_channels == ref.out_channels == 256
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        want = ref(x)
        got = net_torch.hourglass_forward({k: v.clone() for k, v in ref.state_dict().items()}, x)
    assert len(want) == len(got) == 2
    for a, b in zip(got, want):
        assert torch.equal(a, b)
