"""Per-launch table of the tensor-core kernels in one DLA-34 ctdet step (CUDA events, live).
    python tools/profile_layers.py [B]  ->  kind, geometry, ms, TFLOP/s, share"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from centernet_pytorch_lightning_b200 import ops  # noqa: E402
from centernet_pytorch_lightning_b200.decode import ctdet_decode  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
model, head = bench.seeded_weights(bench.CONFIGS[2])
model, head = model.to(dev), head.to(dev)
x = torch.rand(B, 3, 512, 512, device=dev)

# record geometry next to each launch
geo = []
_conv, _dcn, _heads = ops.conv2d, ops.dcnv2, ops.heads_fused


def conv2d(x, wpk, Co, k, stride, pad, *a, **kw):
    v = ops.as_view(x)
    geo.append(f"conv {v.C:4d}->{Co:4d} k{k} s{stride} @{v.H}x{v.W} cs{v.cstride} mode{kw.get('out_mode', 0)}")
    return _conv(x, wpk, Co, k, stride, pad, *a, **kw)


def dcnv2(x, om, wpk, Co, *a, **kw):
    v = ops.as_view(x)
    geo.append(f"dcn  {v.C:4d}->{Co:4d} k3 s1 @{v.H}x{v.W}")
    return _dcn(x, om, wpk, Co, *a, **kw)


def heads_fused(x, w3pk, bias3, w1cat, bias1, head_conv, c_out, act):
    v = ops.as_view(x)
    r = _heads(x, w3pk, bias3, w1cat, bias1, head_conv, c_out, act)
    if r is not None:
        geo.append(f"head {v.C:4d}->{head_conv}x{len(c_out)}->{sum(c_out)} 3x3+1x1 fused @{v.H}x{v.W}")
    return r


ops.conv2d, ops.dcnv2, ops.heads_fused = conv2d, dcnv2, heads_fused


def step():
    with torch.no_grad():
        o = head(model(x)[-1], sigmoid=("heatmap",))
        return ctdet_decode(o["heatmap"], o["width_height"], reg=o["regression"])


for _ in range(3):
    step()
torch.cuda.synchronize()
geo.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with ops.LaunchProfiler() as prof:
    e0.record()
    step()
    e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1)
rows = []
for g, (kind, flops, a, b) in zip(geo, prof.records):
    ms = a.elapsed_time(b)
    rows.append((ms, g, flops))
tc = sum(r[0] for r in rows)
print(f"B={B} step {total:.3f} ms (with events), tensor-core kernels {tc:.3f} ms in {len(rows)} launches")
for ms, g, fl in rows:
    print(f"{g:48s} {ms*1e3:9.1f} us {fl/ms/1e9:8.1f} TF/s {100*ms/tc:5.1f}%")
