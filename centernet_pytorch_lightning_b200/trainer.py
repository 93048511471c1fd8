"""Data-parallel training step of the hot path (BASELINE config 3: DLA-34 ctdet, batch 16 per GPU, 8 x B200):

    forward (train-mode BatchNorm) -> sigmoid_clamped + FocalLoss + 2 x RegL1Loss -> backward
    -> gradient all-reduce (NCCL, NVLink / NVSwitch) overlapped with the rest of the backward pass -> Adam

what Lightning + DistributedDataParallel + torch.optim.Adam do around the reference's `training_step`
(CenterNet/centernet.py:70-80 training_step, :94-105 configure_optimizers; SURVEY.md section 2 rows 17-18: DDP is the
only parallelism, BatchNorm is NOT synchronised, 19.68 M parameters = 78.7 MB of fp32 gradients per step).

Layout: every trainable parameter is re-seated as a view of ONE flat fp32 buffer; a second flat buffer holds the
gradients, laid out in REVERSE registration order so that the gradients produced first by the backward pass form the
first contiguous bucket.  The backward kernels write their results straight into those views (autograd_ops._deliver),
each parameter reports `ready`, and a bucket's all-reduce is launched the moment its last gradient has been enqueued
-- on NCCL's own stream, ordered after the producing kernels by an event -- while the compute stream carries on with
the data / weight gradients of earlier layers.  One Adam kernel then updates the whole flat buffer
(cnb_adam_step, gradients scaled by 1/world_size).  `torch.distributed` is plumbing here: rendezvous, the NCCL
communicator and its stream.
"""
import torch

from . import _lib


class FlatTrainer:
    def __init__(self, modules, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, bucket_mb=8.0, process_group=None, world_size=None,
                 comm=None):
        """modules: iterable of nn.Modules (backbone, heads) already on their CUDA device.
        comm: "nccl" (default: torch.distributed all_reduce per bucket) or "p2p" (this library's own kernel over NVLink
        peer memory, csrc/p2p_allreduce.cu: the gradient buffer is symmetric memory; CNB_GRAD_COMM selects it too)."""
        import os
        self.comm = comm or os.environ.get("CNB_GRAD_COMM", "nccl")
        params, seen = [], set()
        for m in modules:
            for p in m.parameters():
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        if not params:
            raise ValueError("FlatTrainer: no trainable parameters")
        self.params = params
        self.device = params[0].device
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.pg = process_group
        self.step_dev = None          # device copies of (step, lr): set by GraphedCtdetStep so that a replayed graph sees them
        self.lr_dev = None
        if world_size is None:
            world_size = torch.distributed.get_world_size(process_group) if (
                torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.world = world_size
        # ---- flat buffers (16-byte aligned segments)
        order = list(reversed(params))                       # gradient-production order, approximately
        limit = int(bucket_mb * (1 << 20) / 4)
        offs, total = {}, 0
        self.buckets = []                                     # [start, end, n_params]: contiguous slices of the buffers
        start, count = 0, 0
        for p in order:
            offs[id(p)] = total
            total += (p.numel() + 3) // 4 * 4
            count += 1
            if total - start >= limit:
                total = (total + 31) // 32 * 32               # bucket ends on 32 floats: shards of 4 floats x <= 8 ranks
                self.buckets.append([start, total, count])
                start, count = total, 0
        if count:
            total = (total + 31) // 32 * 32
            self.buckets.append([start, total, count])
        self.numel = total
        dev = self.device
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self._p2p = None
        if self.comm == "p2p" and self.world > 1:
            self._p2p = _P2PComm(total, dev, self.pg, len(self.buckets))
            self.flat_g = self._p2p.grad
        else:
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in order:
                o, n = offs[id(p)], p.numel()
                view = self.flat_p[o:o + n].view_as(p)
                view.copy_(p.detach().float())
                p.data = view                                  # the parameter now lives in the flat buffer
                p._cnb_grad = self.flat_g[o:o + n].view_as(p)
        self._bucket_of = {}
        bi = 0
        for p in order:
            while offs[id(p)] >= self.buckets[bi][1]:
                bi += 1
            self._bucket_of[id(p)] = bi
        for p in order:
            p._cnb_ready = self._make_ready(self._bucket_of[id(p)])
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self.launch_log = []                                  # (bucket, n_launched_before_finish) per step, for tests

    # ------------------------------------------------------------------------------------------------------------
    def _make_ready(self, b):
        def ready():
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return ready

    def _launch(self, b):
        if self._launched[b]:
            return
        self._launched[b] = True
        if self._p2p is not None:
            s, e, _ = self.buckets[b]
            self._p2p.allreduce(b, s, e - s)
        elif self.world > 1:
            s, e, _ = self.buckets[b]
            # async: NCCL's stream waits for everything enqueued on the compute stream so far (the kernels that wrote
            # this bucket), the compute stream does not wait for NCCL until `finish_backward`
            self._works.append(torch.distributed.all_reduce(self.flat_g[s:e], group=self.pg, async_op=True))
        self.launch_log.append(b)

    def zero_grad(self):
        if self._p2p is not None:
            self._p2p.next_step()
        self.flat_g.zero_()
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self.launch_log = []

    def finish_backward(self):
        """Launch whatever has not been launched (parameters without a gradient this step never report ready: the dead
        `project` convolutions of Tree.forward, pose_dla_dcn.py:254-255), then make the compute stream wait."""
        overlapped = len(self.launch_log)
        for b in range(len(self.buckets)):
            self._launch(b)
        for w in self._works:
            w.wait()
        if self._p2p is not None:
            self._p2p.wait_all()
        return overlapped

    def optimizer_step(self):
        """Adam on the whole flat buffer (torch.optim.Adam defaults of centernet.py:95); gradients are averaged over
        the ranks here (sum all-reduce * 1/world)."""
        self.step_count += 1
        if self.step_dev is not None:
            self.step_dev.add_(1)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cnb_adam_step(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.flat_m),
                                                _lib.ptr(self.flat_v), self.numel, self.lr, self.betas[0], self.betas[1],
                                                self.eps, self.step_count, _lib.ptr(self.step_dev), _lib.ptr(self.lr_dev),
                                                1.0 / self.world, _lib.stream_ptr(self.device)), "cnb_adam_step")
        # the kernel wrote through raw pointers: tell autograd / the packed-weight caches that the parameters changed
        torch.autograd.graph.increment_version(self.params)

    def grads(self):
        """name-free access for tests: parameter -> its gradient view"""
        return {id(p): p._cnb_grad for p in self.params}


class _P2PComm:
    """Symmetric-memory gradient buffer + signal pads for csrc/p2p_allreduce.cu (torch.distributed._symmetric_memory is
    the plumbing: allocation, handle exchange, peer / multicast mappings)."""

    def __init__(self, numel, device, group, nslots):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        L = _lib.lib()
        self.device, self.nslots = device, nslots
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.grad = symm.empty(numel, dtype=torch.float32, device=device)
        self.grad.zero_()
        self.sig = symm.empty(L.cnb_p2p_signal_bytes() // 4, dtype=torch.int32, device=device)
        self.sig.zero_()
        hg, hs = symm.rendezvous(self.grad, group), symm.rendezvous(self.sig, group)
        self._handles = (hg, hs)
        self.bufs = (ctypes.c_void_p * self.world)(*[int(p) for p in hg.buffer_ptrs])
        self.sigs = (ctypes.c_void_p * self.world)(*[int(p) for p in hs.buffer_ptrs])
        import os
        try:
            mc = int(hg.multicast_ptr)          # 0 when the fabric / driver offers no multicast mapping (no NVLS)
        except Exception:
            mc = 0
        self.multicast = mc if os.environ.get("CNB_P2P_MULTICAST", "1") != "0" else 0
        self.counters = torch.zeros(64, dtype=torch.int32, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.stream = torch.cuda.Stream(device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def next_step(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cnb_p2p_next_step(_lib.ptr(self.epoch), _lib.stream_ptr(self.device)), "cnb_p2p_next_step")

    def allreduce(self, slot, off, n):
        """on the communication stream, ordered after everything enqueued on the compute stream so far"""
        import ctypes
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cnb_p2p_allreduce(self.bufs, self.sigs, ctypes.c_void_p(self.multicast) if self.multicast else None,
                                                    _lib.ptr(self.counters), _lib.ptr(self.epoch), off, n, self.rank,
                                                    self.world, slot, 16, ctypes.c_void_p(self.stream.cuda_stream)),
                       "cnb_p2p_allreduce")

    def wait_all(self):
        import ctypes
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cnb_p2p_wait(_lib.ptr(self.sig), _lib.ptr(self.epoch), self.nslots, self.world,
                                               ctypes.c_void_p(self.stream.cuda_stream)), "cnb_p2p_wait")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


def ctdet_training_step(model, head, trainer, x, target, weights=(1.0, 0.1, 1.0)):
    """One optimisation step of `CenterNetDetection` (centernet_detection.py:88-130 + centernet.py:70-80,94-95) on
    this package's kernels.  Returns the loss tensor (device, not synchronised)."""
    from .utils.decode import sigmoid_clamped
    from .utils.losses import FocalLoss, RegL1Loss
    trainer.zero_grad()
    out = head(model(x)[-1])
    hm = sigmoid_clamped(out["heatmap"])
    loss = weights[0] * FocalLoss()(hm, target["heatmap"]) \
        + weights[1] * RegL1Loss()(out["width_height"], target["regression_mask"], target["indices"], target["width_height"]) \
        + weights[2] * RegL1Loss()(out["regression"], target["regression_mask"], target["indices"], target["regression"])
    loss.backward()
    trainer.finish_backward()
    trainer.optimizer_step()
    return loss.detach()


class GraphedCtdetStep:
    """The whole optimisation step -- forward, losses, backward, bucketed all-reduce, Adam -- captured ONCE as a CUDA graph
    and replayed per batch (streams and graphs instead of a tracing compiler, like engine.CtdetEngine for inference).
    Issued eagerly the step is ~950 kernel launches of this library plus the autograd tape's bookkeeping: ~27 ms of host
    time for ~25 ms of device time at batch 16, so the host would bound it.  Everything in the step is capture-safe:
    kernels go to the current stream through the C ABI, tensor maps are encoded on the host at capture time for
    addresses that stay fixed (the allocations live in the graph's private pool), the packed-weight copies are rebuilt
    by pack kernels that are part of the graph, Adam's step count and learning rate are device scalars, and NCCL's
    all-reduce is captured as a cross-stream dependency.

        step = GraphedCtdetStep(model, head, trainer, batch=16, res=512)
        loss = step(x, target)          # x [B,3,H,W] fp32, target dict (device or pinned host tensors)
    """

    def __init__(self, model, head, trainer, batch, res, n_classes=80, max_objs=128, weights=(1.0, 0.1, 1.0), warmup=3):
        dev = trainer.device
        o = res // 4
        self.model, self.head, self.trainer, self.weights = model, head, trainer, weights
        self.x = torch.zeros(batch, 3, res, res, device=dev)
        self.target = {"heatmap": torch.zeros(batch, n_classes, o, o, device=dev),
                       "indices": torch.zeros(batch, max_objs, dtype=torch.int64, device=dev),
                       "regression_mask": torch.zeros(batch, max_objs, dtype=torch.bool, device=dev),
                       "width_height": torch.zeros(batch, max_objs, 2, device=dev),
                       "regression": torch.zeros(batch, max_objs, 2, device=dev)}
        trainer.step_dev = torch.full((1,), trainer.step_count, dtype=torch.int32, device=dev)
        trainer.lr_dev = torch.full((1,), trainer.lr, dtype=torch.float32, device=dev)
        self.graph = None
        self._warmup = warmup
        self.loss = None

    def set_lr(self, lr):
        self.trainer.lr = lr
        self.trainer.lr_dev.fill_(lr)

    def _load(self, x, target):
        self.x.copy_(x, non_blocking=True)
        for k, v in self.target.items():
            v.copy_(target[k], non_blocking=True)

    def _snapshot(self):
        tr = self.trainer
        bufs = [b for m in (self.model, self.head) for b in m.buffers()]
        return (tr.flat_p.clone(), tr.flat_m.clone(), tr.flat_v.clone(), tr.step_count, [b.clone() for b in bufs], bufs)

    def _restore(self, snap):
        tr = self.trainer
        p, m, v, step, saved, bufs = snap
        tr.flat_p.copy_(p)
        tr.flat_m.copy_(m)
        tr.flat_v.copy_(v)
        tr.step_count = step
        tr.step_dev.fill_(step)
        for b, s in zip(bufs, saved):
            b.copy_(s)
        torch.autograd.graph.increment_version(tr.params)

    def _capture(self):
        """Warm-up steps (they size caches / workspaces and bind the autograd thread's CUDA context) run on a side stream
        and are UNDONE afterwards -- parameters, Adam state, BatchNorm statistics and the step counter are restored --
        so that capturing costs no optimisation step; recording the graph executes nothing."""
        dev = self.trainer.device
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                ctdet_training_step(self.model, self.head, self.trainer, self.x, self.target, self.weights)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.loss = ctdet_training_step(self.model, self.head, self.trainer, self.x, self.target, self.weights)
        self.trainer.step_count = snap[3]      # recording advanced the host-side counter only
        self.graph = g

    def __call__(self, x, target):
        self._load(x, target)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self.trainer.step_count += 1
        torch.autograd.graph.increment_version(self.trainer.params)
        return self.loss
