#!/usr/bin/env python
"""bench.py -- images/sec of the hot path on synthetic 512x512 COCO-shaped batches (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1|2|4|5] [--mode infer|train]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Default (what the driver runs): BASELINE configs[1] -- DLA-34 backbone (+16 DCNv2) -> ctdet heads (sigmoid fused into
the heat-map head epilogue) -> fused ctdet_decode, batch 32 per GPU, 512x512; weak scaling, no data-path collective
(SURVEY.md 8e).  The same line also carries a short TRAINING leg (configs[2]: forward + focal/L1 + backward +
bucketed NCCL gradient all-reduce overlapped with the backward pass + Adam, batch 16 per GPU) under "train", so that the
driver's 1/2/4/8-GPU runs measure the path's one real collective.  `--mode train` makes that leg the headline;
`--config N` selects the other BASELINE inference configs (1: res_18 B=2 256^2, 4: resdcn_50 B=16, 5: DLA-34
multi_pose B=32 split over the ranks).

One JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks); `e2e` = the same through
the public API with pinned-host inputs copied H2D every step and the result read back D2H every step; `roofline` = the
tcgen05 conv/DCN kernel family (tensor bound) with per-launch CUDA-event timing; `roofline_decode` = the fused decode
(HBM bound); `roofline_loss` = the fused focal-loss kernel; `parity` = what the GPU tests assert for this precision
mode; `cpu_baseline` = the CPU oracle (a port of the reference's PyTorch path) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CT = {"heatmap": 80, "width_height": 2, "regression": 2}
MP = {"heatmap": 1, "width_height": 2, "regression": 2, "heatmap_keypoints": 17, "keypoints": 34,
      "heatmap_keypoints_offset": 2}
# per-image forward GFLOP (2*MAC) and the Sigma-max roofline ceiling (img/s/GPU) from SURVEY.md 8(d) / BASELINE.md 3
CONFIGS = {
    1: dict(arch="res_18", heads=CT, head_conv=64, batch=2, res=256, task="ctdet", gflop=11.36, ceiling=88692.0,
            split=False, name="ResNet-18 ctdet inference, batch=2, 256x256 synthetic, fwd+sigmoid+ctdet_decode (configs[0] on the GPU)"),
    2: dict(arch="dla_34", heads=CT, head_conv=256, batch=32, res=512, task="ctdet", gflop=66.20, ceiling=16288.0,
            split=False, name="DLA-34 ctdet inference, batch=32/GPU, 512x512 synthetic, fwd+sigmoid+ctdet_decode (configs[1])"),
    4: dict(arch="resdcn_50", heads=CT, head_conv=64, batch=16, res=512, task="ctdet", gflop=52.38, ceiling=17984.0,
            split=False, name="ResNet-50-DCN ctdet inference, batch=16/GPU, 512x512 synthetic, fwd+sigmoid+ctdet_decode (configs[3])"),
    5: dict(arch="dla_34", heads=MP, head_conv=256, batch=32, res=512, task="pose", gflop=80.48, ceiling=13255.0,
            split=True, name="DLA-34 multi_pose inference, batch=32 split over the ranks, 512x512 synthetic, fwd+sigmoid+multi_pose_decode (configs[4])"),
}
TRAIN = dict(arch="dla_34", heads=CT, head_conv=256, batch=16, res=512, gflop=3 * 66.20,
             name="DLA-34 ctdet training, batch=16/GPU, 512x512 synthetic, fwd + focal/L1 + bwd + gradient all-reduce + Adam (configs[2])")
PARITY = {
    "precision_mode": "bf16 operands, fp32 accumulation (tcgen05); heads / losses / decode in fp32",
    "decode": "bit-exact vs the numpy oracle and reference-generated goldens (tests/test_decode_gpu.py)",
    "losses": "<= 1e-5 relative vs reference-generated goldens (tests/test_losses_gpu.py, tests/test_dropin_gpu.py)",
    "network_bf16": "rel-L2 <= 3e-2 vs the fp32 oracle up to 192x256; at 2x512x512: <= 1.5e-2 with feature-independent DCN offsets, "
                    "<= 8e-2 with trained-like offsets (tests/test_model_gpu.py, tests/test_parity_e2e_gpu.py)",
    "end_to_end_fp32_strict": "fp32-strict mode (CUDA-core fp32 kernels): top-100 index set and order equal to the fp32 CPU oracle, "
                              "|dcoord| <= 1e-4 (tests/test_parity_e2e_gpu.py::test_fp32_strict_end_to_end)",
    "dcn": "pinned to torchvision.ops.deform_conv2d (the reference's tteepe/DCNv2 extension is not vendored: parity unpinned by the reference)",
    "training": "per-operator gradients vs PyTorch fp32 autograd; whole DLA-34 step vs the train-mode oracle pinned to the reference "
                "(tests/test_train_gpu.py)",
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons / power sampled DURING the timed region (B200_PROFILING.md's clocks line) through
    in-process NVML -- a polling `nvidia-smi` process takes the driver lock for long enough to stall kernel
    launches, which a launch-heavy step would feel."""

    def __init__(self, gpu_index, period_s=0.05):
        super().__init__(daemon=True)
        self.gpu, self.period, self.rows, self._stop_evt = gpu_index, period_s, [], threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop_evt.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((sm, mx, rs, pw))
                self._stop_evt.wait(self.period)
        except Exception as e:   # no NVML: the line says so instead of inventing clocks
            self.rows.append(("error", repr(e)))

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        rows = [r for r in self.rows if r[0] != "error"]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                    "error": self.rows[0][1] if self.rows else "no samples"}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake": 0x80}
        reasons = sorted(n for n, b in bits.items() if any(r[2] & b for r in rows))
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": rows[0][1], "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[3] for r in rows)}


# ------------------------------------------------------------------------------------------------------
def seeded_weights(cfg, seed=1234):
    import torch
    from centernet_pytorch_lightning_b200.models import create_model
    from centernet_pytorch_lightning_b200.models.heads import CenterHead
    from centernet_pytorch_lightning_b200.utils.synthetic import randomize_
    torch.manual_seed(seed)
    model = create_model(cfg["arch"])
    head = CenterHead(cfg["heads"], model.out_channels, cfg["head_conv"])
    randomize_(model.state_dict(), seed)
    randomize_(head.state_dict(), seed + 1)
    return model.eval(), head.eval()


def cpu_reference_step(cfg, sd, hd, x):
    """The reference's CPU path for this workload (oracle port): forward, sigmoid_, decode."""
    import torch
    from oracle import decode_np, net_torch
    with torch.no_grad():
        if cfg["arch"].startswith("dla"):
            feat = net_torch.dla34_seg_forward(sd, x)
        else:
            kind, n = cfg["arch"].split("_")
            feat = net_torch.pose_resnet_forward(sd, x, int(n), kind == "resdcn")
        o = net_torch.center_head_forward(hd, feat, cfg["heads"])
        heat = o["heatmap"].sigmoid_()
        if cfg["task"] == "ctdet":
            return decode_np.ctdet_decode(heat.numpy(), o["width_height"].numpy(), o["regression"].numpy())
        return decode_np.multi_pose_decode(heat.numpy(), o["width_height"].numpy(), o["keypoints"].numpy(), o["regression"].numpy(),
                                           o["heatmap_keypoints"].sigmoid_().numpy(), o["heatmap_keypoints_offset"].numpy())


def cpu_reference_train_step(sd, hd, x, tgt):
    """fwd + focal/L1 + bwd of the reference path on the CPU (oracle port; no optimizer: it is noise next to this)."""
    from oracle import net_torch, task_torch
    for d in (sd, hd):
        for v in d.values():
            if v.is_floating_point() and v.requires_grad:
                v.grad = None
    with net_torch.training():
        o = net_torch.center_head_forward(hd, net_torch.dla34_seg_forward(sd, x), CT)
    loss = task_torch.ctdet_loss_torch(o, tgt)
    loss.backward()
    return float(loss)


def time_cpu(cfg, sample_batch, reps, warmup, train=False):
    import torch
    from centernet_pytorch_lightning_b200.utils.synthetic import ctdet_targets
    torch.set_num_threads(os.cpu_count() or 1)
    model, head = seeded_weights(cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    hd = {k: v.clone() for k, v in head.state_dict().items()}
    x = torch.rand(sample_batch, 3, cfg["res"], cfg["res"], generator=torch.Generator().manual_seed(1))
    if train:
        for d in (sd, hd):
            for k, v in d.items():
                if v.is_floating_point() and "running" not in k:
                    v.requires_grad_(True)
        tgt = ctdet_targets(sample_batch, 80, cfg["res"] // 4, cfg["res"] // 4, n_obj=32, seed=2)
        fn = lambda: cpu_reference_train_step(sd, hd, x, tgt)   # noqa: E731
    else:
        fn = lambda: cpu_reference_step(cfg, sd, hd, x)         # noqa: E731
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads()


def metric_name(cfg, train):
    if train:
        return "images/sec @512x512 DLA-34 ctdet training (fwd + focal/L1 + bwd + allreduce + Adam)"
    if cfg is CONFIGS[2]:
        return "images/sec @512x512 DLA-34 ctdet (fwd + fused decode)"
    return f"images/sec {cfg['arch']} {cfg['task']} @{cfg['res']}x{cfg['res']} (fwd + fused decode)"


def run_reference(args, rank):
    if rank != 0:
        return
    train = args.mode == "train"
    cfg = TRAIN if train else CONFIGS[args.config]
    cfg = dict(cfg, task=cfg.get("task", "ctdet"))
    sample = 2
    ts, threads = time_cpu(cfg, sample, max(args.steps, 1), args.warmup, train=train)
    mean_t = sum(ts) / len(ts)
    value = sample / mean_t
    what = "fwd+loss+bwd" if train else "fwd+sigmoid+decode"
    line = {
        "impl": "reference", "metric": metric_name(CONFIGS.get(args.config) if not train else None, train), "value": value,
        "unit": "images/sec", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": f"{sample} images per step on host cores (CPU path does not batch-scale)"},
        "cpu_baseline": {"value": value, "unit": "images/sec", "cores": threads, "kind": "port",
                         "sample": f"oracle port of the reference CPU path (PyTorch fp32 + torchvision deform_conv2d + numpy decode), "
                                   f"{what}, {sample} images/step, {len(ts)} steps"},
        "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def sync_all(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(value, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def train_leg(args, rank, world, dev, steps, warmup, with_e2e=True):
    """configs[2]: one optimisation step per 16-image batch and rank.  Returns a dict (rank 0: full; others: None)."""
    import torch
    from centernet_pytorch_lightning_b200 import _lib
    from centernet_pytorch_lightning_b200.trainer import FlatTrainer, GraphedCtdetStep, ctdet_training_step
    from centernet_pytorch_lightning_b200.utils.synthetic import ctdet_targets

    B, R = TRAIN["batch"], TRAIN["res"]
    model, head = seeded_weights(TRAIN)
    model, head = model.to(dev).train(), head.to(dev).train()
    trainer = FlatTrainer([model, head], lr=1e-4, bucket_mb=8.0, world_size=world)
    g = torch.Generator().manual_seed(300 + rank)
    x_host = [torch.rand(B, 3, R, R, generator=g).pin_memory() for _ in range(2)]
    t_host = [{k: v.pin_memory() for k, v in ctdet_targets(B, 80, R // 4, R // 4, n_obj=32, seed=400 + rank + 7 * i).items()}
              for i in range(2)]
    x_dev = [t.to(dev) for t in x_host]
    t_dev = [{k: v.to(dev) for k, v in t.items()} for t in t_host]
    if args.no_graph:
        do_step = lambda x, t: ctdet_training_step(model, head, trainer, x, t)   # noqa: E731
    else:   # the public training API: the whole step (fwd, losses, bwd, all-reduce, Adam) replayed as one CUDA graph
        do_step = GraphedCtdetStep(model, head, trainer, B, R)
    losses = []
    launch_mode = "eager (autograd tape)" if args.no_graph else "CUDA graph replay of the whole step (trainer.GraphedCtdetStep)"
    try:
        losses.append(do_step(x_dev[0], t_dev[0]).clone())
        torch.cuda.synchronize()
    except Exception as e:   # a capture that the NCCL build at hand refuses must not cost the line: run the same step eagerly
        if args.no_graph:
            raise
        launch_mode = f"eager (graph capture failed: {repr(e)[:120]})"
        args.no_graph = True
        do_step = lambda x, t: ctdet_training_step(model, head, trainer, x, t)   # noqa: E731
        losses.append(do_step(x_dev[0], t_dev[0]).clone())
    for i in range(1, warmup):
        losses.append(do_step(x_dev[i % 2], t_dev[i % 2]).clone())
    sync_all(world)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses.append(do_step(x_dev[i % 2], t_dev[i % 2]).clone())
    e1.record()
    sync_all(world)
    launches = _lib.launch_count() - n0
    if not args.no_graph:      # replayed kernels are not counted by the library: count one eager twin of the step
        snap = do_step._snapshot()
        n1 = _lib.launch_count()
        ctdet_training_step(model, head, trainer, x_dev[0], t_dev[0])
        launches = (_lib.launch_count() - n1) * steps
        do_step._restore(snap)
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev, world)
    overlapped = len(trainer.launch_log)
    # the same steps without the collective (world forced to 1 on the reducer): the difference is the exposed all-reduce
    exposed_ms = None
    if world > 1:
        trainer.world = 1
        nocoll = do_step if args.no_graph else GraphedCtdetStep(model, head, trainer, B, R)
        for i in range(2):
            nocoll(x_dev[i % 2], t_dev[i % 2])
        sync_all(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            nocoll(x_dev[i % 2], t_dev[i % 2])
        b.record()
        sync_all(world)
        ms_nocoll = max_over_ranks(a.elapsed_time(b), dev, world)
        exposed_ms = (ms_total - ms_nocoll) / steps
        trainer.world = world
    # end to end: batch and targets from pinned host memory every step, loss read back every step
    e2e_value, h2d = None, None
    if with_e2e:
        sync_all(world)
        t0 = time.perf_counter()
        for i in range(steps):
            if args.no_graph:
                xs = x_host[i % 2].to(dev, non_blocking=True)
                ts = {k: v.to(dev, non_blocking=True) for k, v in t_host[i % 2].items()}
                loss = do_step(xs, ts)
            else:
                loss = do_step(x_host[i % 2], t_host[i % 2])   # pinned host -> the graph's static input buffers (H2D)
            float(loss)                                   # D2H + sync: the caller logs the loss (centernet.py:75)
        sync_all(world)
        t_e2e = max_over_ranks(time.perf_counter() - t0, dev, world)
        e2e_value = world * B * steps / t_e2e
        h2d = x_host[0].numel() * 4 + sum(v.numel() * v.element_size() for v in t_host[0].values())
    if rank != 0:
        return None
    value = world * B * steps / (ms_total / 1e3)
    ms_step = ms_total / steps
    pk = peaks()
    tf = TRAIN["gflop"] * 1e9 * B / (ms_step * 1e-3) / 1e12
    return {
        "metric": metric_name(None, True), "value": value, "unit": "images/sec", "ms_per_step": ms_step, "steps": steps,
        "warmup": warmup, "batch_per_gpu": B, "n_gpus": world, "workload": TRAIN["name"],
        "loss_first_last": [float(losses[0]), float(losses[-1])],
        "allreduce": {"bytes_per_step": trainer.numel * 4, "buckets": len(trainer.buckets),
                      "buckets_launched_during_backward": overlapped, "exposed_ms_per_step": exposed_ms,
                      "how": "NCCL all-reduce per 8 MB bucket of the flat fp32 gradient buffer, launched (async, NCCL stream) when the "
                             "bucket's last gradient kernel has been enqueued; exposed = step time with minus without the collective"},
        "tflops_algorithmic": tf, "frac_of_sustained_bf16_peak": tf / pk["tf_sust"],
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "how": "pinned host image batch + target dict -> H2D every step, loss scalar D2H + sync every step"},
        "gpu_launches": int(launches),
        "launch": launch_mode,
        "precision": "bf16 activations / operands, fp32 accumulation, fp32 master weights, gradients and Adam state",
    }


def loss_roofline(dev, pk):
    """Fused focal loss (sigmoid_clamped + _neg_loss, forward AND backward in one pass) on the config-3 map size."""
    import torch
    from centernet_pytorch_lightning_b200 import _lib
    n = 16 * 80 * 128 * 128
    logits = torch.randn(n, device=dev)
    gt = torch.rand(n, device=dev) ** 4
    gt[::997] = 1.0
    grad = torch.empty_like(logits)
    stats = torch.empty(3, device=dev)
    L = _lib.lib()
    ws = _lib.workspace(dev, L.cnb_focal_loss_workspace_bytes(n))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(8):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.cnb_focal_loss_fwd_bwd(_lib.ptr(logits), _lib.ptr(gt), None, _lib.ptr(grad), _lib.ptr(stats), n, _lib.ptr(ws),
                                            ws.numel(), _lib.stream_ptr(dev)), "cnb_focal_loss_fwd_bwd")
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    by = 12.0 * n
    return {"bound": "hbm", "kernel": "focal_fwd_bwd_kernel (sigmoid_clamped + _neg_loss forward and backward, one pass)", "achieved": by / (ms * 1e-3) / 1e9,
            "peak": pk["hbm"], "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None, "ms": ms,
            "bytes_per_launch": by, "elements": n, "l2": "flushed (256 MB write) before every timed launch",
            "note": "12 B/element algorithmic (read logits, gt; write the unnormalised gradient); autograd's backward scales it by "
                    "grad_out/num_pos in a second 8 B/element pass"}


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from centernet_pytorch_lightning_b200 import _lib, build, ops
    from centernet_pytorch_lightning_b200.decode import ctdet_decode, multi_pose_decode

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    _lib.lib()

    if args.mode == "train":
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        tr = train_leg(args, rank, world, dev, args.steps, args.warmup)
        clocks = sampler.stop() if sampler else None
        if rank == 0:
            cpu = None
            if world == 1 and not args.no_cpu_baseline:
                ts, threads = time_cpu(dict(TRAIN, task="ctdet"), 1, 1, 0, train=True)
                cpu = {"value": 1 / min(ts), "unit": "images/sec", "cores": threads, "kind": "port",
                       "sample": "1 image 512x512, fwd + focal/L1 + bwd on host cores, one run (oracle port of the reference CPU path)"}
            line = {"metric": tr["metric"], "value": tr["value"], "unit": "images/sec", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": TRAIN["name"], "global_batch": TRAIN["batch"] * world,
                               "parallelism": f"dp{world} (gradient all-reduce over NCCL, BatchNorm statistics per rank as in the reference)",
                               "l2": "two resident input/target sets (100 MB + 168 MB > 126 MB L2) alternate; GBs of activations per step"},
                    "e2e": tr["e2e"], "gpu_launches": tr["gpu_launches"], "clocks": clocks,
                    "roofline": {"bound": "tensor", "kernel": "whole training step (conv fprop/dgrad/wgrad + DCN GEMMs on tcgen05)",
                                 "achieved": tr["tflops_algorithmic"], "peak": peaks()["tf_sust"], "unit": "TFLOP/s",
                                 "frac": tr["frac_of_sustained_bf16_peak"], "traffic": None,
                                 "note": "algorithmic 3 x 66.2 GFLOP/img over the WHOLE step time (bandwidth-bound BatchNorm / DCN "
                                         "gather-scatter kernels included)"},
                    "cpu_baseline": cpu, "train": tr, "parity": PARITY}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    from centernet_pytorch_lightning_b200.engine import CtdetEngine, MultiPoseEngine

    cfg = CONFIGS[args.config]
    BATCH = cfg["batch"] // world if cfg["split"] else cfg["batch"]
    RES = cfg["res"]
    pose = cfg["task"] == "pose"
    ncol = 57 if pose else 6
    model, head = seeded_weights(cfg)
    model, head = model.to(dev), head.to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = [torch.rand(BATCH, 3, RES, RES, generator=g).pin_memory() for _ in range(2)]
    x_dev = [t.to(dev) for t in x_host]
    det_host = torch.empty(BATCH, 100, ncol).pin_memory()
    # the public serving API: the step (layout change -> backbone -> heads -> decode) captured as a CUDA graph
    Engine = MultiPoseEngine if pose else CtdetEngine
    eng = Engine(model, head, BATCH, RES, RES, K=100, slots=2, graphs=not args.no_graph)

    def step(x):   # eager launch sequence (per-launch profiling legs below)
        with torch.no_grad():
            return eng._step(x)

    # ---- device-resident throughput: inputs already in HBM, copied (D2D) into the engine's input slot ----
    for i in range(args.warmup):
        eng.input(i % 2).copy_(x_dev[i % 2])
        eng.run(i % 2)
    sync_all(world)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.2)
    sync_all(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.input(i % 2).copy_(x_dev[i % 2])
        det = eng.run(i % 2)
    e1.record()
    sync_all(world)
    launches = eng.launches_per_step * args.steps   # kernels inside the replayed graphs (counted at warm-up)
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clocks = sampler.stop() if sampler else None
    value = world * BATCH * args.steps / (ms_total / 1e3)

    # ---- end to end: pinned host input H2D every step (prefetched on a copy stream), detections D2H -------
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            eng.input(i % 2).copy_(x_host[i % 2], non_blocking=True)
            copied[i % 2].record(copy_stream)

    for ev in consumed:
        ev.record(main)
    n_e2e = args.steps
    prefetch(0)
    sync_all(world)
    t0 = time.perf_counter()
    for i in range(n_e2e):
        prefetch(i + 1)
        main.wait_event(copied[i % 2])
        det = eng.run(i % 2)
        consumed[i % 2].record(main)
        det_host.copy_(det, non_blocking=True)
        main.synchronize()                      # the caller reads this step's detections
    sync_all(world)
    e2e_value = world * BATCH * n_e2e / max_over_ranks(time.perf_counter() - t0, dev, world)
    h2d = BATCH * 3 * RES * RES * 4
    d2h = BATCH * 100 * ncol * 4

    # ---- training leg (all ranks: it contains the path's one collective) ------------------------------------
    train = None
    if args.config == 2 and not args.no_train_leg:
        del eng
        torch.cuda.empty_cache()
        try:
            train = train_leg(args, rank, world, dev, steps=max(3, min(args.steps, 6)), warmup=3, with_e2e=False)
        except Exception as e:   # the inference line must survive a failure of the extra leg
            train = {"error": repr(e)[:300]} if rank == 0 else None
        eng = Engine(model.eval(), head.eval(), BATCH, RES, RES, K=100, slots=1, graphs=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (rank 0): per-launch CUDA events, live -------------------
    pk = peaks()
    with ops.LaunchProfiler() as prof:
        for i in range(2):
            step(x_dev[i % 2])
    conv_ms, conv_fl, n_conv = prof.summary()
    conv_ms, conv_fl, n_conv = conv_ms / 2, conv_fl / 2, n_conv // 2
    tf = conv_fl / (conv_ms * 1e-3) / 1e12
    dcn_ms, dcn_fl, n_dcn = prof.summary("dcn")
    roofline = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM family: conv_tma_kernel (CTA pairs for N >= 192), conv_rows_kernel, dcn_fp_kernel, head_fused_kernel",
                "achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": tf / pk["tf_sust"],
                "traffic": None, "peak_source": pk["src"] + " (sustained bf16, kernel timed inside a long step)",
                "launches_per_step": n_conv, "kernel_ms_per_step": conv_ms, "flops_per_step": conv_fl,
                "share_of_step": conv_ms / (ms_total / args.steps),
                "dcn_ms_per_step": dcn_ms / 2, "dcn_tflops": (dcn_fl / 2) / (dcn_ms / 2 * 1e-3) / 1e12 if dcn_ms else None}
    try:   # DRAM traffic of the family per step, summed over the committed ncu --set full captures (profiles/traffic.json)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        roofline["traffic"] = tj.get("conv family per step [32,3,512,512]", {}).get("bytes")
    except Exception:
        tj = {}
    # decode alone on resident head maps
    _, o = step(x_dev[0])
    if pose:
        dec = lambda: multi_pose_decode(o["heatmap"], o["width_height"], o["keypoints"], reg=o["regression"],   # noqa: E731
                                        hm_hp=o["heatmap_keypoints"], hp_offset=o["heatmap_keypoints_offset"])
        dec_bytes = BATCH * (18 * 128 * 128 * 4 + 100 * 38 * 4 + 17 * 100 * 2 * 4 + 100 * 57 * 4)
        dec_kernel = "decode_stream_kernel + decode_merge_kernel + multi_pose_assoc_kernel"
    else:
        dec = lambda: ctdet_decode(o["heatmap"], o["width_height"], reg=o["regression"])   # noqa: E731
        hw = (RES // 4) ** 2
        dec_bytes = BATCH * (80 * hw * 4 + 100 * 16 + 100 * 24)
        dec_kernel = "decode_stream_kernel + decode_merge_kernel (fused nms+topk+gather: warp-autonomous streaming scan, per-image merge; timed together with the workspace memset)"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_decode(fn):
        for _ in range(3):
            fn()
        dts = []
        for _ in range(10):
            flush.fill_(1)                                 # evict L2 (256 MB > 126 MB)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            dts.append(a.elapsed_time(b))
        return statistics.median(dts)

    dec_ms = time_decode(dec)
    dec_gbs = dec_bytes / (dec_ms * 1e-3) / 1e9
    dec_traffic = tj.get("decode [32,80,128,128]", {}).get("bytes") if (args.config == 2) else None
    variants = {"network head maps of this step (random-init weights: heat logits saturate, plateaus of equal scores)":
                {"ms": dec_ms, "GB/s": dec_gbs, "frac": dec_gbs / pk["hbm"]}}
    if not pose:   # the decode micro-benchmark on standalone maps in two distributions
        from centernet_pytorch_lightning_b200.utils import synthetic
        for kind, label in (("uniform", "distinct-uniform (SURVEY 8d-i)"), ("bumps", "COCO-shaped Gaussian bumps (SURVEY 8d-ii)")):
            hm, wh, rg = (torch.from_numpy(t).to(dev) for t in synthetic.ctdet_maps(BATCH, 80, RES // 4, RES // 4, seed=1, kind=kind))
            ms = time_decode(lambda: ctdet_decode(hm, wh, reg=rg))
            variants[label] = {"ms": ms, "GB/s": dec_bytes / (ms * 1e-3) / 1e9, "frac": dec_bytes / (ms * 1e-3) / 1e9 / pk["hbm"]}
    roofline_decode = {"bound": "hbm", "kernel": dec_kernel,
                       "achieved": dec_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": dec_gbs / pk["hbm"],
                       "traffic": dec_traffic, "ms": dec_ms, "bytes_per_launch": dec_bytes, "peak_source": pk["src"],
                       "l2": "flushed (256 MB write) before every timed launch", "by_input": variants}
    # host-buffer C-ABI entry (cnb_ctdet_decode_host): H2D of the maps + decode + D2H inside the call
    decode_host = None
    if not pose:
        import ctypes
        hm, wh, rg = (o[k].cpu().numpy() for k in ("heatmap", "width_height", "regression"))
        outb = det_host.numpy()
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
        L = _lib.lib()
        hs = []
        for i in range(5):
            t0 = time.perf_counter()
            _lib.check(L.cnb_ctdet_decode_host(P(hm), P(wh), P(rg), P(outb), BATCH, 80, RES // 4, RES // 4, 100), "cnb_ctdet_decode_host")
            hs.append(time.perf_counter() - t0)
        decode_host = {"entry": "cnb_ctdet_decode_host", "ms": min(hs[1:]) * 1e3, "images_per_s": BATCH / min(hs[1:]),
                       "h2d_bytes": int(hm.nbytes + wh.nbytes + rg.nbytes), "d2h_bytes": int(outb.nbytes),
                       "note": "pageable numpy buffers: bound by the host-to-device copy of the heat maps"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ts, threads = time_cpu(cfg, 2, 2, 1)
        cpu = {"value": 2 / min(ts), "unit": "images/sec", "cores": threads, "kind": "port",
               "sample": f"2 images {RES}x{RES}, fwd+sigmoid+decode on host cores, best of 2 after 1 warm-up "
                         "(oracle port of the reference CPU path)"}

    line = {
        "metric": metric_name(cfg, False), "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if cfg["split"] else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["name"], "global_batch": BATCH * world, "parallelism": f"dp{world} (independent images, no collective)",
                   "l2": "inputs rotate over 2 resident buffers, copied D2D into the engine slot inside the "
                         "timed region; several GB of activations per step",
                   "launch": "CUDA graph replay (one graph per input slot)" if not args.no_graph else "eager launches",
                   "weights": "seeded random init (He-normal convs, perturbed BN stats, non-zero DCN offsets)",
                   "accumulate": "fp32 (TMEM)"},
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "pinned host batch -> H2D into the engine's input slot on a copy stream (prefetch depth 1) -> engine.run "
                       "(CUDA-graph replay of model -> CenterHead -> decode) -> D2H of the detections + sync every step"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "roofline_decode": roofline_decode, "roofline_loss": loss_roofline(dev, pk),
        "decode_host_entry": decode_host, "cpu_baseline": cpu, "parity": PARITY, "train": train,
        "img_per_s_ceiling_conv_roofline": cfg["ceiling"] * world,
        "frac_of_conv_roofline": value / (cfg["ceiling"] * world),
    }
    if pose:
        line["roofline_multi_pose"] = roofline_decode
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4, 5], help="BASELINE.json inference config (3 = --mode train)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the short training leg of the default line")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on one node
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
