"""-m gpu: the graph-captured serving step (`CtdetEngine` / `MultiPoseEngine`) must reproduce the eager path
bit for bit (same kernels, same order) and stay correct across replays with new inputs."""
import numpy as np
import pytest
import torch

from centernet_pytorch_lightning_b200.decode import ctdet_decode, multi_pose_decode
from centernet_pytorch_lightning_b200.engine import CtdetEngine, MultiPoseEngine
from centernet_pytorch_lightning_b200.models import create_model
from centernet_pytorch_lightning_b200.models.heads import CenterHead
from centernet_pytorch_lightning_b200.utils.synthetic import randomize_
from oracle import decode_np

pytestmark = pytest.mark.gpu
CT = {"heatmap": 80, "width_height": 2, "regression": 2}
MP = {"heatmap": 1, "width_height": 2, "regression": 2, "heatmap_keypoints": 17, "keypoints": 34,
      "heatmap_keypoints_offset": 2}


def _build(arch, heads, head_conv, seed, dev):
    torch.manual_seed(seed)
    m = create_model(arch).eval()
    h = CenterHead(heads, m.out_channels, head_conv).eval()
    randomize_(m.state_dict(), seed)
    randomize_(h.state_dict(), seed + 1)
    return m.to(dev), h.to(dev)


@pytest.mark.parametrize("graphs", [True, False])
def test_ctdet_engine_matches_eager_and_oracle_decode(cuda_dev, graphs):
    m, h = _build("dla_34", CT, 256, 3, cuda_dev)
    B, H, W = 2, 128, 256
    eng = CtdetEngine(m, h, B, H, W, slots=2, graphs=graphs)
    g = torch.Generator().manual_seed(5)
    for it in range(3):                                  # replays with fresh inputs, alternating slots
        x = torch.rand(B, 3, H, W, generator=g).to(cuda_dev)
        slot = it % 2
        eng.input(slot).copy_(x)
        det = eng.run(slot).clone()
        maps = {k: v.clone() for k, v in eng.head_maps(slot).items()}
        with torch.no_grad():
            o = h(m(x)[-1], sigmoid=("heatmap",))
            want = ctdet_decode(o["heatmap"], o["width_height"], reg=o["regression"])
        torch.cuda.synchronize()
        assert all(torch.equal(maps[k], o[k]) for k in CT)
        assert torch.equal(det, want)
        ref = decode_np.ctdet_decode(maps["heatmap"].cpu().numpy(), maps["width_height"].cpu().numpy(),
                                     maps["regression"].cpu().numpy())
        assert np.array_equal(det.cpu().numpy(), ref)    # decode bit-exact on the engine's own head maps
    assert eng.launches_per_step > 50


def test_multi_pose_engine_matches_eager(cuda_dev):
    m, h = _build("dla_34", MP, 256, 9, cuda_dev)
    B, H, W = 1, 128, 128
    eng = MultiPoseEngine(m, h, B, H, W, slots=1)
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(1)).to(cuda_dev)
    eng.input(0).copy_(x)
    det = eng.run(0).clone()
    with torch.no_grad():
        o = h(m(x)[-1], sigmoid=("heatmap", "heatmap_keypoints"))
        want = multi_pose_decode(o["heatmap"], o["width_height"], o["keypoints"], reg=o["regression"],
                                 hm_hp=o["heatmap_keypoints"], hp_offset=o["heatmap_keypoints_offset"])
    torch.cuda.synchronize()
    assert det.shape == (B, 100, 57) and torch.equal(det, want)
    ref = decode_np.multi_pose_decode(*(o[k].cpu().numpy() for k in ("heatmap", "width_height", "keypoints", "regression",
                                                                     "heatmap_keypoints", "heatmap_keypoints_offset")))
    assert np.array_equal(det.cpu().numpy(), ref)


def test_engine_rejects_cpu(cuda_dev):
    from centernet_pytorch_lightning_b200 import _lib
    m = create_model("res_18").eval()
    h = CenterHead(CT, m.out_channels, 64).eval()
    with pytest.raises(_lib.CnbError):
        CtdetEngine(m, h, 1, 64, 64)
