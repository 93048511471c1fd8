"""Device-resident throughput of the other BASELINE.json configs through the graph-captured engine (1 GPU):
    python tools/bench_configs.py  ->  one line per config (images/s, ms/step)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200.engine import CtdetEngine, MultiPoseEngine  # noqa: E402
from centernet_pytorch_lightning_b200.models import create_model  # noqa: E402
from centernet_pytorch_lightning_b200.models.heads import CenterHead  # noqa: E402
from centernet_pytorch_lightning_b200.utils.synthetic import randomize_  # noqa: E402

CT = {"heatmap": 80, "width_height": 2, "regression": 2}
MP = {"heatmap": 1, "width_height": 2, "regression": 2, "heatmap_keypoints": 17, "keypoints": 34,
      "heatmap_keypoints_offset": 2}
CONFIGS = [  # name, arch, heads, head_conv, B, res, engine
    ("config1 res_18 ctdet B=2 256^2", "res_18", CT, 64, 2, 256, CtdetEngine),
    ("config2 dla_34 ctdet B=32 512^2", "dla_34", CT, 256, 32, 512, CtdetEngine),
    ("config4 resdcn_50 ctdet B=16 512^2", "resdcn_50", CT, 64, 16, 512, CtdetEngine),
    ("config5 dla_34 multi_pose B=32 512^2 (1 GPU)", "dla_34", MP, 256, 32, 512, MultiPoseEngine),
]
dev = torch.device("cuda:0")
for name, arch, heads, hc, B, res, Eng in CONFIGS:
    torch.manual_seed(0)
    m = create_model(arch).eval()
    h = CenterHead(heads, m.out_channels, hc).eval()
    randomize_(m.state_dict(), 0)
    randomize_(h.state_dict(), 1)
    eng = Eng(m.to(dev), h.to(dev), B, res, res, slots=1)
    x = torch.rand(B, 3, res, res, device=dev)
    eng.input(0).copy_(x)
    for _ in range(3):
        eng.run(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        eng.run(0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    det = eng.run(0)
    print(f"{name:48s} {B / ms * 1e3:9.1f} img/s  {ms:8.3f} ms/step  launches/step {eng.launches_per_step}  out {tuple(det.shape)}")
    del eng, m, h
    torch.cuda.empty_cache()
