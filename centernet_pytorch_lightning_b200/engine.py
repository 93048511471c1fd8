"""Inference engine: the whole hot path of `CenterNetDetection.forward` + `ctdet_decode`
(centernet_detection.py:88-95, :183-188) captured once as a CUDA graph and replayed per batch.

One step launches ~95 kernels (layout change, 57 convolutions, 16 DCN pairs, pools, up-samplings, heads,
decode); issued one by one from Python the step is launch-bound on the host (about 20 ms of CPU time for
11 ms of GPU work), so the launch sequence is recorded into a graph -- streams and graphs instead of a
tracing compiler -- and replayed with one call.  Everything inside the graph is this package's sm_100a kernels
called through the C ABI; the graph only removes the per-launch host cost.

    eng = CtdetEngine(model, head, batch=32, height=512, width=512)      # warm-up + capture
    eng.input(0).copy_(x_pinned, non_blocking=True)                      # stage a batch (NCHW fp32)
    det = eng.run(0)                                                     # [B,K,6] on the device (static tensor)

Two input slots (each with its own graph, sharing one memory pool) let the host->device copy of batch i+1
overlap the replay of batch i.
"""
import torch

from . import _lib
from .decode import ctdet_decode, multi_pose_decode


class CtdetEngine:
    def __init__(self, model, head, batch, height=512, width=512, K=100, slots=2, device=None, graphs=True):
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.CnbError("CtdetEngine needs a CUDA device (sm_100a kernels; no CPU fallback)")
        self.model, self.head, self.K = model.eval(), head.eval(), K
        self.heat_name = "heatmap" if "heatmap" in head.heads else next(n for n in head.heads if n.startswith("heatmap"))
        self.inputs = [torch.zeros(batch, 3, height, width, device=self.device) for _ in range(slots)]
        self.outputs = [None] * slots
        self.maps = [None] * slots
        self.graphs = [None] * slots
        self.launches_per_step = None
        with torch.no_grad(), torch.cuda.device(self.device):
            self._step(self.inputs[0])          # warm-up: packs weights, sizes workspaces, sets func attributes
            n0 = _lib.launch_count()
            self.outputs[0], self.maps[0] = self._step(self.inputs[0])
            self.launches_per_step = _lib.launch_count() - n0   # this package's kernels in one step
            torch.cuda.synchronize(self.device)
            if graphs:
                pool = None
                for i in range(slots):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        self.outputs[i], self.maps[i] = self._step(self.inputs[i])
                    pool = g.pool()
                    self.graphs[i] = g
                torch.cuda.synchronize(self.device)

    def _heads(self, x, sigmoid):
        # the NHWC bf16 feature map goes straight to the heads; the NCHW fp32 copy that `model(x)` returns for
        # the reference's API (a 134 MB write per batch) is not needed here
        feat = self.model.forward_nhwc(x) if hasattr(self.model, "forward_nhwc") else self.model(x)[-1]
        return self.head(feat, sigmoid=sigmoid)

    def _step(self, x):
        o = self._heads(x, (self.heat_name,))
        det = ctdet_decode(o[self.heat_name], o["width_height"], reg=o.get("regression"), K=self.K)
        return det, o

    def input(self, slot=0):
        """The static NCHW fp32 device buffer of `slot`; copy the batch into it before `run(slot)`."""
        return self.inputs[slot]

    def run(self, slot=0):
        """Replay the captured step on the current stream; returns the slot's static [B,K,6] detections."""
        if self.graphs[slot] is not None:
            self.graphs[slot].replay()
        else:
            with torch.no_grad():
                self.outputs[slot], self.maps[slot] = self._step(self.inputs[slot])
        return self.outputs[slot]

    def head_maps(self, slot=0):
        """The head maps of the slot's last run (name -> [B,C,H,W] fp32; the heat map is already sigmoided)."""
        return self.maps[slot]


class MultiPoseEngine(CtdetEngine):
    """The same for `CenterNetMultiPose.forward` + `multi_pose_decode` (centernet_multi_pose.py:213-231): heads
    heatmap / width_height / regression / heatmap_keypoints / keypoints / heatmap_keypoints_offset ->
    [B,K,3J+6] detections (decode/multi_pose.py:7-96)."""

    def _step(self, x):
        o = self._heads(x, (self.heat_name, "heatmap_keypoints"))
        det = multi_pose_decode(o[self.heat_name], o["width_height"], o["keypoints"], reg=o.get("regression"),
                                hm_hp=o["heatmap_keypoints"], hp_offset=o.get("heatmap_keypoints_offset"), K=self.K)
        return det, o
