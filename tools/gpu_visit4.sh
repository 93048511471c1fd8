bash tools_gpu_tests.sh tests/test_decode_gpu.py
timeout 120 python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from centernet_pytorch_lightning_b200.decode import ctdet_decode
from centernet_pytorch_lightning_b200.utils import synthetic
dev = torch.device('cuda:0')
for kind in ('uniform', 'bumps'):
    heat, wh, reg = synthetic.ctdet_maps(32, 80, 128, 128, seed=1, kind=kind)
    heat, wh, reg = (torch.from_numpy(a).to(dev) for a in (heat, wh, reg))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3): ctdet_decode(heat, wh, reg)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctdet_decode(heat, wh, reg); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts)//2]
    print(kind, 'decode ms median', ms, 'min', min(ts), 'GB/s', 167.9e6/ms/1e6)
PY
