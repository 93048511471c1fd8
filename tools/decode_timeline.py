"""Per-CTA phase timeline of decode_scan_kernel (CNB_DECODE_DEBUG=1): python tools/decode_timeline.py [kind]"""
import ctypes
import os
import sys

os.environ["CNB_DECODE_DEBUG"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import _lib  # noqa: E402
from centernet_pytorch_lightning_b200.decode import ctdet_decode  # noqa: E402
from centernet_pytorch_lightning_b200.utils import synthetic  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
dev = torch.device("cuda:0")
heat, wh, reg = synthetic.ctdet_maps(32, 80, 128, 128, seed=1, kind=kind)
heat, wh, reg = (torch.from_numpy(a).to(dev) for a in (heat, wh, reg))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ctdet_decode(heat, wh, reg)
flush.fill_(1)
ctdet_decode(heat, wh, reg)
torch.cuda.synchronize()
n = 296
buf = (ctypes.c_ulonglong * (n * 8))()
L = _lib.lib()
L.cnb_decode_debug_dump.restype = ctypes.c_int
L.cnb_decode_debug_dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert L.cnb_decode_debug_dump(buf, n * 8) == 0
a = np.array(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)
t0 = a[:, 0].min()
us = lambda v: (v - t0) / 1e3
print(f"{kind}: kernel span {us(a[:, 3].max()):.1f} us")
for name, col in (("start", 0), ("first chunk done", 1), ("last chunk scanned", 2), ("end (after merges)", 3)):
    v = us(a[:, col])
    print(f"  {name:22s} min {v.min():7.1f}  median {np.median(v):7.1f}  max {v.max():7.1f} us")
print(f"  flushes per CTA: mean {a[:, 5].mean():.1f} max {a[:, 5].max()};  time in flushes per CTA: mean {a[:, 4].mean() / 1e3:.1f} us max {a[:, 4].max() / 1e3:.1f} us")
print(f"  slot 6 (scan kernel: chunks with hits; stream kernel: ns the producer waited for a free ring slot): mean {a[:, 6].mean():.1f} max {a[:, 6].max()}")
print(f"  time blocked on tile data per CTA (thread 0): mean {a[:, 7].mean() / 1e3:.1f} us max {a[:, 7].max() / 1e3:.1f} us")
mer = us(a[:, 3]) - us(a[:, 2])
print(f"  tail after last scan (group accounting + merge): median {np.median(mer):.1f} max {mer.max():.1f} us; CTAs with tail > 5us: {(mer > 5).sum()}")
