"""Multi-GPU plumbing for the hot path: one process per GPU, images are independent units.

Inference / decode shard the batch across ranks with NO data-path collective; the only
exchanges are the gather of the tiny [B,K,6] / [B,K,57] detection rows to every rank and the scalar
all-reduce behind `self.log(..., sync_dist=True)` (centernet.py:87,90).  Works on NCCL (CUDA tensors)
and on gloo (CPU tensors; used by the world_size-2 tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local: this rank's [b_local, K, D] rows (b_local from shard_range) -> [n_total, K, D] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    K, D = local.shape[1], local.shape[2]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local batch does not match shard_range"
    pad = torch.zeros((cap, K, D), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)


def sync_mean(stats: dict, group=None) -> dict:
    """`self.log(name, value, sync_dist=True)`: mean over ranks of every scalar, one fused all-reduce."""
    keys = sorted(stats)
    if not keys:
        return {}
    t = torch.stack([torch.as_tensor(stats[k], dtype=torch.float32).reshape(()) for k in keys])
    first = stats[keys[0]]
    if isinstance(first, torch.Tensor):
        t = t.to(first.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t / dist.get_world_size(group)
    return {k: t[i] for i, k in enumerate(keys)}
