"""Data-parallel plumbing for the quantizer path: one process per GPU, batch sharded on dim 0.

The only coupled quantities of the path are linear sums over vectors:
  * per EMA stage, between assignment and finalize: all-reduce(sum) of `stats = [dw (K*D) | cnt (K)]`
    so that every rank applies the identical full-batch update and codebooks stay bit-identical
    without a broadcast (the reference's nn.DataParallel keeps only shard 0's statistics,
    scripts/train_ablation.py:189 -- this engine deliberately equals the single-process full-batch
    result instead);
  * gradients of the small encoder/decoder and of a standard-VQ codebook: averaged (DDP semantics).

Everything here works on CPU tensors with the `gloo` backend as well, which is how the host-side
logic is tested without a GPU (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as torch_dist

_GROUP = None
_ENABLED = False


def enable(group=None) -> None:
    """Turn on the per-stage EMA-statistics all-reduce (call after init_process_group)."""
    global _GROUP, _ENABLED
    if not torch_dist.is_available() or not torch_dist.is_initialized():
        raise RuntimeError("vqb200.dist.enable(): torch.distributed is not initialised")
    _GROUP = group
    _ENABLED = True


def disable() -> None:
    global _GROUP, _ENABLED
    _GROUP, _ENABLED = None, False


def enabled() -> bool:
    return _ENABLED and torch_dist.is_initialized() and world_size() > 1


def world_size() -> int:
    if not (_ENABLED and torch_dist.is_initialized()):
        return 1
    return torch_dist.get_world_size(_GROUP)


def rank() -> int:
    if not (_ENABLED and torch_dist.is_initialized()):
        return 0
    return torch_dist.get_rank(_GROUP)


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """In-place sum of the packed EMA statistics over the data-parallel group.  Enqueued on the
    current CUDA stream by NCCL (stream-ordered between ema_accumulate and ema_finalize)."""
    if enabled():
        torch_dist.all_reduce(stats, op=torch_dist.ReduceOp.SUM, group=_GROUP)
    return stats


def shard_bounds(n: int, rank_: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n` samples for a rank; sizes differ by at most one."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def average_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> int:
    """DDP-style gradient averaging over the group for parameters that have a gradient.
    Flattens into buckets sized for launch latency (NVSwitch gives full bandwidth to every peer, so
    bucket count -- not link count -- is what matters).  Returns the number of all-reduce calls."""
    if not enabled():
        return 0
    w = float(world_size())
    grads: List[torch.Tensor] = [p.grad for p in params if p is not None and p.grad is not None]
    calls = 0
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size, calls
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        torch_dist.all_reduce(flat, op=torch_dist.ReduceOp.SUM, group=_GROUP)
        flat.div_(w)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        calls += 1
        bucket, size = [], 0

    for g in grads:
        nbytes = g.numel() * g.element_size()
        if bucket and (size + nbytes > bucket_bytes or g.dtype != bucket[0].dtype):
            flush()
        bucket.append(g)
        size += nbytes
    flush()
    return calls
