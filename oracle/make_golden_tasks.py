"""Generates tests/golden/task_{detection,multi_pose}.npz by executing the UNMODIFIED reference task classes
(`CenterNetDetection`, `CenterNetMultiPose`: /root/reference/CenterNet/centernet_detection.py:97-130,175-225 and
centernet_multi_pose.py:97-155,215-264) on the CPU, through `oracle/ref_shim.ref_tasks()` (Lightning stub).

    python -m oracle.make_golden_tasks

Recorded: `loss(outputs, target)` -> (loss, loss_stats) and `test_step_end` (decode + rescale + per-class split / top-k)
on the seeded inputs of `oracle/task_torch.task_inputs`.  The restatement in oracle/task_torch.py is asserted equal on
the way.  The GPU box has no /root/reference; the fixtures travel.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, task_torch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _clone(d):
    return {k: v.clone() for k, v in d.items()}


def detection_results_array(results):
    """{class_id: [n,5]} -> [N,6] rows (class_id, x1, y1, x2, y2, score) in class order."""
    rows = [np.concatenate([np.full((len(v), 1), j, np.float32), v.astype(np.float32)], 1) for j, v in sorted(results.items())
            if len(v)]
    return np.concatenate(rows, 0) if rows else np.zeros((0, 6), np.float32)


def main():
    _, Det, Pose = ref_shim.ref_tasks()
    ns = task_torch.namespace("reference")
    meta = {"padding": [3.0, 5.0], "scale": [0.75, 0.75]}
    for kind, cls, Task in (("ctdet", Det, task_torch.DetectionTask), ("pose", Pose, task_torch.MultiPoseTask)):
        ref = cls("res_18")                                        # the backbone is irrelevant to loss / test_step_end
        task = Task(ns, "res_18", build_backbone=False)
        out, tgt = task_torch.task_inputs(kind)
        loss, stats = ref.loss([_clone(out)], tgt)
        loss2, stats2 = task.loss([_clone(out)], tgt)
        assert torch.equal(loss, loss2) and all(torch.equal(torch.as_tensor(stats[k]), torch.as_tensor(stats2[k])) for k in stats)
        rec = {f"out_{k}": v.numpy() for k, v in out.items()}
        rec.update({f"tgt_{k}": v.numpy() for k, v in tgt.items()})
        rec.update({f"stat_{k}": np.float32(float(v)) for k, v in stats.items()})
        # gradient of the composed loss w.r.t. every head map
        g_in = {k: v.clone().requires_grad_(True) for k, v in out.items()}
        lg, _ = ref.loss([{k: v * 1 for k, v in g_in.items()}], tgt)   # non-leaf copies: sigmoid_clamped is in place
        lg.backward()
        rec.update({f"grad_{k}": v.grad.numpy() for k, v in g_in.items()})
        # test_step_end on image 0 (the reference squeezes a batch of one)
        one = {k: v[:1].clone() for k, v in out.items()}
        det_direct = task.decode(_clone(one)).numpy()
        image_id, results = ref.test_step_end((7, [_clone(one)], [copy.deepcopy(meta)]))
        assert image_id == 7
        rec["decoded"] = det_direct
        rec["results"] = detection_results_array(results) if kind == "ctdet" else np.asarray(results, np.float32)
        rec["meta_padding"], rec["meta_scale"] = np.float32(meta["padding"]), np.float32(meta["scale"])
        name = "task_detection.npz" if kind == "ctdet" else "task_multi_pose.npz"
        np.savez_compressed(os.path.join(GOLD, name), **rec)
        print(name, os.path.getsize(os.path.join(GOLD, name)), "bytes; loss", float(loss), "results", rec["results"].shape)


if __name__ == "__main__":
    main()
