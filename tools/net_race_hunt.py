"""Whole DLA-34 ctdet step repeated on one input: every head map and the detections must be bit-identical from run to run."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from centernet_pytorch_lightning_b200.decode import ctdet_decode
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
model, head = bench.seeded_weights(bench.CONFIGS[2])
model, head = model.to(dev), head.to(dev)
x = torch.rand(B, 3, 512, 512, device=dev)
def step():
    with torch.no_grad():
        o = head(model(x)[-1], sigmoid=("heatmap",))
        det = ctdet_decode(o["heatmap"], o["width_height"], reg=o["regression"])
    return [o["heatmap"].clone(), o["width_height"].clone(), o["regression"].clone(), det.clone()]
first = step()
bad = 0
for it in range(iters):
    cur = step()
    diff = [i for i, (a, b) in enumerate(zip(cur, first)) if not torch.equal(a, b)]
    if diff:
        bad += 1
        if bad <= 5:
            a, b = cur[diff[0]], first[diff[0]]
            print(f"iter {it}: outputs {diff} differ; first: {int((a != b).sum())} elements, max |diff| {(a - b).abs().max().item():.3e}")
print(f"{bad} differing runs in {iters} (B={B})")
