"""`CenterDetectionSample` (CenterNet/sample/ctdet.py:9-92) with the encoding done on the GPU, batched: Gaussian-splat
heat maps, centre indices, masks, sizes and sub-pixel offsets from box lists (`cnb_ctdet_encode`, csrc/next_rows.cu).

Why: the reference encodes on the CPU in the dataloader and ships an [80,128,128] fp32 heat map per image to the device
(5.2 MB per image, 84 MB per 16-image batch -- more than the image batch itself); a box list is a few hundred bytes.

    enc = CenterDetectionSample()                       # same constructor arguments as the reference
    target = enc.encode_batch(boxes, class_ids, counts, input_hw=(512, 512), device=dev)   # dict of [B,...] tensors
    img, target = enc(img, annotations)                 # the reference's per-sample call, computed on img's device
"""
import numpy as np
import torch

from .. import _lib


class CenterDetectionSample:
    def __init__(self, down_ratio=4, num_classes=80, max_objects=128, gaussian_type="umich"):
        if gaussian_type != "umich":
            raise NotImplementedError("centernet_b200 CenterDetectionSample: the 'umich' Gaussian (the reference default) only")
        self.down_ratio, self.num_classes, self.max_objects = down_ratio, num_classes, max_objects

    def encode_batch(self, boxes, class_ids, counts, input_hw, device):
        """boxes [B,M,4] COCO (x, y, w, h) in input pixels (float64), class_ids [B,M] int, counts [B] int (objects per
        image, <= M <= max_objects) -> dict(heatmap [B,C,H,W], regression_mask [B,max_objects] bool, indices int64,
        width_height, regression [B,max_objects,2])."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.CnbError("CenterDetectionSample.encode_batch runs on CUDA devices only (no CPU fallback)")
        boxes = torch.as_tensor(np.asarray(boxes, dtype=np.float64))
        B, M = boxes.shape[0], self.max_objects
        H, W = input_hw[0] // self.down_ratio, input_hw[1] // self.down_ratio
        bx = torch.zeros((B, M, 4), dtype=torch.float64)
        cl = torch.zeros((B, M), dtype=torch.int32)
        n_in = min(boxes.shape[1], M)
        bx[:, :n_in] = boxes[:, :n_in]
        cl[:, :n_in] = torch.as_tensor(np.asarray(class_ids, dtype=np.int32))[:, :n_in]
        cnt = torch.clamp(torch.as_tensor(np.asarray(counts, dtype=np.int32)), max=M)
        bx, cl, cnt = bx.to(device), cl.to(device), cnt.to(device)
        heat = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=device)
        ind = torch.empty((B, M), dtype=torch.int64, device=device)
        mask = torch.empty((B, M), dtype=torch.uint8, device=device)
        wh = torch.empty((B, M, 2), dtype=torch.float32, device=device)
        reg = torch.empty((B, M, 2), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().cnb_ctdet_encode(_lib.ptr(bx), _lib.ptr(cl), _lib.ptr(cnt), _lib.ptr(heat), _lib.ptr(ind),
                                                   _lib.ptr(mask), _lib.ptr(wh), _lib.ptr(reg), B, self.num_classes, H, W, M,
                                                   self.down_ratio, _lib.stream_ptr(device)), "cnb_ctdet_encode")
        return {"heatmap": heat, "regression_mask": mask.bool(), "indices": ind, "width_height": wh, "regression": reg}

    def __call__(self, img, target):
        """The reference's per-sample signature (sample/ctdet.py:39): img [3,H,W] (note: it reads `_, input_w, input_h =
        img.shape`, i.e. square inputs), target = list of annotation dicts with `bbox` and `class_id` / `category_id`."""
        _, input_w, input_h = img.shape
        n = min(len(target), self.max_objects)
        boxes = np.zeros((1, max(n, 1), 4), np.float64)
        cls = np.zeros((1, max(n, 1)), np.int32)
        for k in range(n):
            ann = target[k]
            boxes[0, k] = ann["bbox"]
            cls[0, k] = ann["class_id"] if "class_id" in ann else int(ann["category_id"]) - 1
        out = self.encode_batch(boxes, cls, [n], (input_h, input_w), img.device)
        return img, {k: v[0] for k, v in out.items()}
