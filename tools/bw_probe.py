"""Bandwidth probe: torch elementwise ops on the up-sampling's tensor sizes vs cnb_dw_deconv_up."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_pytorch_lightning_b200 import ops
dev = torch.device("cuda:0")
B = 32
a = torch.randn(B, 128, 128, 64, device=dev).to(torch.bfloat16)
b = torch.randn(B, 128, 128, 64, device=dev).to(torch.bfloat16)
x = torch.randn(B, 64, 64, 64, device=dev).to(torch.bfloat16)
wt = ops.relayout_dw_weights(torch.randn(64, 1, 4, 4, device=dev), 2)
out = torch.empty_like(a)
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
mb = a.numel() * 2 / 1e6
t = timeit(lambda: out.copy_(a)); print(f"copy          {t:6.1f} us  {2 * mb / t * 1e-3:5.2f} TB/s")
t = timeit(lambda: torch.add(a, b, out=out)); print(f"add (2r + 1w) {t:6.1f} us  {3 * mb / t * 1e-3:5.2f} TB/s")
t = timeit(lambda: ops.dw_deconv_up(x, wt, 2, add=a)); print(f"dw_deconv_up  {t:6.1f} us  {(2 * mb + x.numel() * 2 / 1e6) / t * 1e-3:5.2f} TB/s")
t = timeit(lambda: ops.dw_deconv_up(x, wt, 2)); print(f"dw_deconv_up (no add) {t:6.1f} us  {(mb + x.numel() * 2 / 1e6) / t * 1e-3:5.2f} TB/s")
