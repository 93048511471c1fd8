"""Generates tests/golden/next_rows.npz from the UNMODIFIED reference (CPU) for the SURVEY.md 8(f) rows:
  * target encoding: CenterNet/sample/ctdet.py:39-90 `CenterDetectionSample.__call__` on seeded random boxes;
  * soft-NMS: CenterNet/utils/nms.py:5-106 `soft_nms` (numba) for methods 0 / 1 / 2;
  * TTA prologue / flip merge: the torch / torchvision calls of centernet_detection.py:143-156, 167-171, executed as
    written there (F.pad -> VF.normalize -> cat(hflip); (a[0:1] + hflip(a[1:2])) / 2).
    python -m oracle.make_golden_next
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
import torchvision.transforms.functional as VF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref_shim.install()
    from CenterNet.sample.ctdet import CenterDetectionSample
    from CenterNet.utils.nms import soft_nms
    rng = np.random.default_rng(2024)
    rec = {}
    # ---- encoding: 3 images, up to 40 boxes each (some degenerate, some clipped at the border, several per class)
    B, M = 3, 40
    boxes = np.zeros((B, M, 4), np.float64)
    cls = np.zeros((B, M), np.int32)
    counts = np.array([40, 17, 1], np.int32)
    enc = CenterDetectionSample()
    outs = []
    for b in range(B):
        anns = []
        for k in range(counts[b]):
            x, y = rng.uniform(-20, 500), rng.uniform(-20, 500)
            w, h = rng.uniform(0, 200), rng.uniform(0, 200)
            if k % 11 == 3:
                w = 0.0                                    # degenerate: skipped by `if h > 0 and w > 0`
            boxes[b, k] = (x, y, w, h)
            cls[b, k] = rng.integers(0, 80) if k % 5 else 7
            anns.append({"bbox": [float(v) for v in boxes[b, k]], "class_id": int(cls[b, k])})
        _, t = enc(torch.zeros(3, 512, 512), anns)
        outs.append(t)
    rec.update(enc_boxes=boxes, enc_cls=cls, enc_counts=counts)
    for k in outs[0]:
        rec[f"enc_{k}"] = np.stack([o[k].numpy() for o in outs])
    # ---- soft-NMS
    N = 120
    ctr = rng.uniform(40, 400, size=(N, 2))
    wh = rng.uniform(10, 120, size=(N, 2))
    b5 = np.concatenate([ctr - wh / 2, ctr + wh / 2, rng.uniform(0.002, 1.0, size=(N, 1))], 1).astype(np.float32)
    rec["nms_in"] = b5
    for method in (0, 1, 2):
        work = b5.copy()
        keep = soft_nms(work, Nt=0.5, method=method)
        rec[f"nms_out_{method}"] = work
        rec[f"nms_keep_{method}"] = np.int32(len(keep))
    # ---- TTA prologue / flip merge
    img = torch.rand(1, 3, 150, 200, generator=torch.Generator().manual_seed(5))
    mean, std = [0.408, 0.447, 0.470], [0.289, 0.274, 0.278]
    ptb, plr = ((150 | 31) + 1 - 150) // 2, ((200 | 31) + 1 - 200) // 2
    s = F.pad(img, (plr, plr, ptb, ptb))
    s = VF.normalize(s, mean, std)
    s = torch.cat([s, VF.hflip(s)])
    rec.update(tta_img=img.numpy(), tta_out=s.numpy(), tta_pad=np.int32([plr, ptb]))
    pair = torch.rand(2, 5, 12, 20, generator=torch.Generator().manual_seed(6))
    rec.update(merge_in=pair.numpy(), merge_out=((pair[0:1] + VF.hflip(pair[1:2])) / 2).numpy())
    np.savez_compressed(os.path.join(GOLD, "next_rows.npz"), **rec)
    print("next_rows.npz", os.path.getsize(os.path.join(GOLD, "next_rows.npz")), "bytes")


if __name__ == "__main__":
    main()
