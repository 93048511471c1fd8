"""CPU, world_size 2 over gloo: the N>1 plumbing of the hot path (batch sharding, detection gather,
sync_dist scalar reduction).  The decode itself is stood in for by the CPU oracle here -- this test is
about the host-side sharding logic, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centernet_pytorch_lightning_b200.dist import gather_detections, shard_range, sync_mean


def test_shard_range_partitions():
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from centernet_pytorch_lightning_b200.utils import synthetic
    from oracle import decode_np
    heat, wh, reg = synthetic.ctdet_maps(n_total, 4, 16, 16, seed=21)     # same seeded batch on every rank
    lo, hi = shard_range(n_total, rank, world)
    local = torch.from_numpy(decode_np.ctdet_decode(heat[lo:hi], wh[lo:hi], reg[lo:hi], K=20))
    full = gather_detections(local, n_total)
    stats = sync_mean({"val_loss": torch.tensor(float(rank + 1)), "hm_loss": torch.tensor(2.0 * rank)})
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["val_loss"].item(), stats["hm_loss"].item()]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 4])
def test_sharded_decode_equals_single_process(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    from centernet_pytorch_lightning_b200.utils import synthetic
    from oracle import decode_np
    heat, wh, reg = synthetic.ctdet_maps(n_total, 4, 16, 16, seed=21)
    want = decode_np.ctdet_decode(heat, wh, reg, K=20)
    assert np.array_equal(np.load(tmp_path / "full.npy"), want)
    np.testing.assert_allclose(np.load(tmp_path / "stats.npy"), [1.5, 1.0])


# ---- data-parallel training plumbing: flat gradient buffer, reverse-order buckets, all-reduce launched on readiness ----
def _trainer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from centernet_pytorch_lightning_b200.trainer import FlatTrainer
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(6)])   # 6 x (4096 + 64) parameters
    tr = FlatTrainer([net], bucket_mb=8000 * 4 / (1 << 20))                  # 2 layers per bucket
    params = list(net.parameters())
    assert all(p.data_ptr() >= tr.flat_p.data_ptr() for p in params)          # re-seated into the flat buffer
    tr.zero_grad()
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(p.shape, generator=g) for p in params]
    launched_when = []
    for p, gr in reversed(list(zip(params, grads))):                          # backward order
        if p is params[2]:
            continue                                                          # a parameter without gradient this step
        p._cnb_grad.add_(gr)
        p._cnb_ready()
        launched_when.append(len(tr.launch_log))
    overlapped = tr.finish_backward()
    got = torch.cat([p._cnb_grad.reshape(-1) for p in params])
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), got.numpy())
        np.save(os.path.join(out_dir, "meta.npy"), np.array([len(tr.buckets), overlapped, max(launched_when)]))
    dist.destroy_process_group()


def test_flat_trainer_buckets_allreduce_over_gloo(tmp_path):
    world = 2
    mp.spawn(_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(6)])
    params = list(net.parameters())
    want = []
    for i, p in enumerate(params):
        tot = torch.zeros(p.shape)
        for rank in range(world):
            g = torch.Generator().manual_seed(100 + rank)
            grads = [torch.randn(q.shape, generator=g) for q in params]
            if i != 2:
                tot += grads[i]
        want.append(tot.reshape(-1))
    np.testing.assert_allclose(np.load(tmp_path / "flat.npy"), torch.cat(want).numpy(), rtol=1e-6, atol=1e-6)
    n_buckets, overlapped, during = np.load(tmp_path / "meta.npy")
    assert n_buckets >= 3
    assert during >= n_buckets - 1          # all buckets but the one holding the gradient-less parameter went out early
    assert overlapped == during             # finish_backward only had to launch the stragglers
