"""Data-parallel plumbing for the quantizer path: one process per GPU, batch sharded on dim 0.

The only coupled quantities of the path are linear sums over vectors:
  * per EMA stage, between assignment and finalize: all-reduce(sum) of `stats = [dw (K*D) | cnt (K)]`
    so that every rank applies the identical full-batch update and codebooks stay bit-identical
    without a broadcast (the reference's nn.DataParallel keeps only shard 0's statistics,
    scripts/train_ablation.py:189 -- this engine deliberately equals the single-process full-batch
    result instead);
  * gradients of the small encoder/decoder and of a standard-VQ codebook: averaged (DDP semantics).

Transport of the per-stage exchange, chosen by `enable(peer=...)`:
  * "nccl" -- `all_reduce(stats)` between K3a and K3b (NCCL over NVLink / NVSwitch);
  * "peer" -- no collective launch at all: K3a accumulates into a CUDA-IPC symmetric buffer and the finalize
    kernels barrier + read every rank's slot over NVLink themselves, summing in rank order
    (csrc/peer.cu, `vqb200_ema_finalize_peer`); bit-identical codebooks on all ranks;
  * "auto" (default) -- "peer" when every rank of the group sits on the same host, owns a distinct CUDA device and
    the IPC mapping plus a handshake barrier succeed on ALL ranks; otherwise "nccl" (reason kept in
    `peer_status()`).

Everything except the peer transport works on CPU tensors with the `gloo` backend as well, which is how the
host-side logic is tested without a GPU (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as torch_dist

_GROUP = None
_ENABLED = False
_PEER: Optional["PeerExchange"] = None
_PEER_STATUS = "off"
_UNIFORM = False

FLAG_BYTES = 256                       # VQB200_MAX_PEERS x uint32, padded
DEFAULT_SLOT_BYTES = 16 << 20          # K*(D+1)*4 per stage: 266 KB at 1024 x 64, 8.4 MB at 16384 x 128


class PeerExchange:
    """Symmetric CUDA-IPC buffer `[flags | slot 0 | slot 1]` mapped by every rank of one node, plus the epoch /
    slot bookkeeping of `vqb200_ema_finalize_peer` (include/vqb200.h).  Construction is collective."""

    def __init__(self, group, device: torch.device, slot_bytes: int = DEFAULT_SLOT_BYTES):
        from . import _lib
        lib = _lib.load()
        self.lib = lib
        self.group = group
        self.device = torch.device(device)
        self.rank = torch_dist.get_rank(group)
        self.world = torch_dist.get_world_size(group)
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.epoch = 0
        self.uses = 0
        self.base: List[int] = [0] * self.world
        self._opened: List[int] = []
        self._own = 0
        if self.world > 16:
            raise RuntimeError("peer exchange supports at most 16 ranks of one node")
        total = FLAG_BYTES + 2 * self.slot_bytes
        handle = (ctypes.c_ubyte * 64)()
        own = ctypes.c_void_p()
        err = ""
        with torch.cuda.device(self.device):
            rc = lib.vqb200_peer_alloc(ctypes.c_size_t(total), ctypes.byref(own), handle)
            if rc != 0:
                err = f"peer_alloc rc={rc}: {_lib.last_error()}"
            else:
                self._own = int(own.value)
        me = {"host": socket.gethostname(), "pid": os.getpid(), "handle": bytes(handle), "err": err,
              "dev": _device_uuid(self.device)}
        infos: List[Optional[dict]] = [None] * self.world
        torch_dist.all_gather_object(infos, me, group=group)
        if not err:
            if any(i["err"] for i in infos):
                err = "a peer failed to allocate: " + "; ".join(i["err"] for i in infos if i["err"])
            elif len({i["host"] for i in infos}) != 1:
                err = "ranks span several hosts (peer memory is intra-node)"
            elif len({i["dev"] for i in infos}) != self.world:
                err = "two ranks share one CUDA device"
        if not err:
            with torch.cuda.device(self.device):
                for p, info in enumerate(infos):
                    if p == self.rank:
                        self.base[p] = self._own
                        continue
                    mapped = ctypes.c_void_p()
                    rc = lib.vqb200_peer_open((ctypes.c_ubyte * 64).from_buffer_copy(info["handle"]), ctypes.byref(mapped))
                    if rc != 0:
                        err = f"peer_open(rank {p}) rc={rc}: {_lib.last_error()}"
                        break
                    self.base[p] = int(mapped.value)
                    self._opened.append(int(mapped.value))
        oks: List[Optional[str]] = [None] * self.world
        torch_dist.all_gather_object(oks, err, group=group)
        bad = [f"rank {p}: {e}" for p, e in enumerate(oks) if e]
        if bad:
            self.close()
            raise RuntimeError("; ".join(bad))
        self._flags = (ctypes.c_void_p * self.world)(*self.base)
        self._slots = [(ctypes.c_void_p * self.world)(*[b + FLAG_BYTES + s * self.slot_bytes for b in self.base])
                       for s in range(2)]
        # handshake: one barrier round through the flag words proves the mapping works in both directions
        with torch.cuda.device(self.device):
            self.epoch += 1
            rc = lib.vqb200_peer_barrier(self._flags, self.rank, self.world, ctypes.c_uint32(self.epoch),
                                         ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"peer_barrier rc={rc}: {_lib.last_error()}")
            torch.cuda.synchronize(self.device)

    def fits(self, K: int, D: int) -> bool:
        return K * (D + 1) * 4 <= self.slot_bytes

    def next_slot(self, n_epochs: int = 1):
        """One use of the exchange: returns (first epoch, device pointer of MY slot, table of every rank's slot) and
        consumes `n_epochs` barrier epochs (1 for a finalize call, S for a single-launch ResidualVQ).  Slots alternate
        per USE: a rank overwrites a slot only after it has passed every barrier of the next use, and a peer signals
        those only after (stream order) it finished reading this one."""
        first = (self.epoch + 1) & 0xFFFFFFFF
        self.epoch = (self.epoch + n_epochs) & 0xFFFFFFFF
        self.uses += 1
        s = self.uses & 1
        return first, self._own + FLAG_BYTES + s * self.slot_bytes, self._slots[s]

    @property
    def flags(self):
        return self._flags

    def close(self) -> None:
        """Collective when the process group is still alive: unmap the peers' buffers, barrier, free the own one."""
        lib = self.lib
        if torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        for m in self._opened:
            lib.vqb200_peer_close(ctypes.c_void_p(m))
        self._opened = []
        try:
            if torch_dist.is_initialized():
                torch_dist.barrier(group=self.group)
        except Exception:
            pass
        if self._own:
            lib.vqb200_peer_free(ctypes.c_void_p(self._own))
            self._own = 0


def _device_uuid(device: torch.device) -> str:
    try:
        return str(torch.cuda.get_device_properties(device).uuid)
    except Exception:
        return f"{socket.gethostname()}:{device}"


def enable(group=None, peer: str = "auto", device: Optional[torch.device] = None,
           slot_bytes: int = DEFAULT_SLOT_BYTES, uniform_shards: bool = False) -> None:
    """Turn on the per-stage EMA-statistics exchange (call after init_process_group; collective).
    peer: "auto" | "peer" | "nccl" (see the module docstring).  `device` defaults to the current CUDA device.
    uniform_shards: the caller's promise that every rank feeds the quantizers tensors of the SAME shape in every
    step; it lets launch-bound shapes take the single-launch ResidualVQ kernel with the exchange inside
    (`vqb200_rvq_small_forward_peer`) -- a choice all ranks must make alike."""
    global _GROUP, _ENABLED, _PEER, _PEER_STATUS, _UNIFORM
    _UNIFORM = bool(uniform_shards)
    if not torch_dist.is_available() or not torch_dist.is_initialized():
        raise RuntimeError("vqb200.dist.enable(): torch.distributed is not initialised")
    if peer not in ("auto", "peer", "nccl"):
        raise ValueError("peer must be 'auto', 'peer' or 'nccl'")
    _close_peer()
    _GROUP = group
    _ENABLED = True
    _PEER_STATUS = "nccl (requested)" if peer == "nccl" else "nccl"
    if peer == "nccl" or torch_dist.get_world_size(group) < 2:
        return
    if not torch.cuda.is_available():
        if peer == "peer":
            raise RuntimeError("vqb200.dist.enable(peer='peer'): no CUDA device")
        _PEER_STATUS = "nccl (no CUDA device)"
        return
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    try:
        _PEER = PeerExchange(group, dev, slot_bytes)
        _PEER_STATUS = "peer"
    except RuntimeError as e:          # the same verdict on every rank (collective checks inside)
        _PEER = None
        if peer == "peer":
            raise
        _PEER_STATUS = f"nccl (peer memory unavailable: {e})"


def _close_peer() -> None:
    global _PEER
    if _PEER is not None:
        _PEER.close()
        _PEER = None


def disable() -> None:
    global _GROUP, _ENABLED, _PEER_STATUS, _UNIFORM
    _AGREED.clear()
    _GRAD_SETS.clear()
    _close_peer()
    _GROUP, _ENABLED, _PEER_STATUS, _UNIFORM = None, False, "off", False


def uniform_shards() -> bool:
    return _UNIFORM and enabled()


_AGREED = {}
_GRAD_SETS: dict = {}       # parameter list -> (this rank's grad mask, union over ranks); see average_gradients


def agree(key, local: bool) -> bool:
    """Collective AND of a per-rank, device-probed decision, cached per `key` (one tiny all-reduce and host read the
    first time a key is seen).  Ranks must ask for a new key in the same step -- which `uniform_shards` promises.
    Used for choices that change the barrier / slot pattern of the peer exchange (the single-launch ResidualVQ kernel):
    `cudaOccupancyMaxActiveClusters` differs between GPUs of one node, so two ranks can disagree near the size limit."""
    k = (id(_GROUP), key)
    if k not in _AGREED:
        if not enabled():
            return bool(local)
        dev = _PEER.device if _PEER is not None else (torch.device("cuda", torch.cuda.current_device())
                                                      if torch.cuda.is_available() and torch_dist.get_backend(_GROUP) == "nccl"
                                                      else torch.device("cpu"))
        t = torch.tensor([1 if local else 0], dtype=torch.int32, device=dev)
        torch_dist.all_reduce(t, op=torch_dist.ReduceOp.MIN, group=_GROUP)
        _AGREED[k] = bool(int(t.item()))
    return _AGREED[k]


def peer_exchange() -> Optional[PeerExchange]:
    """The active peer-memory exchange, or None when the NCCL transport is in use."""
    return _PEER if enabled() else None


def peer_status() -> str:
    return _PEER_STATUS


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


def reset_gradient_sets() -> None:
    """Forget the agreed gradient sets (collective in effect: the next average_gradients call of every rank re-agrees)."""
    _GRAD_SETS.clear()


def average_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> int:
    """DDP-style gradient averaging over the group for all parameters that require a gradient.
    Flattens into buckets sized for launch latency (NVSwitch gives full bandwidth to every peer, so
    bucket count -- not link count -- is what matters).  Returns the number of all-reduce calls."""
    if not enabled():
        return 0
    w = float(world_size())
    # Which parameters take part is agreed on ONCE per parameter list (a rank whose branch produced no gradient for a
    # parameter that other ranks have one for contributes zeros and receives the average; parameters without a gradient
    # on every rank -- EMA codebooks -- stay None as in the reference): bucket sizes and order are then identical on all
    # ranks.  A gradient that shows up later for a parameter outside the agreed set is an error, not a silent omission.
    plist = [p for p in params if p is not None and p.requires_grad]
    local = tuple(1 if p.grad is not None else 0 for p in plist)
    key = tuple(id(p) for p in plist)
    known = _GRAD_SETS.get(key)
    if known is None:
        dev = next((p.grad.device for p in plist if p.grad is not None), plist[0].device if plist else torch.device("cpu"))
        m = torch.tensor(local, dtype=torch.int32, device=dev)
        if m.numel():
            torch_dist.all_reduce(m, op=torch_dist.ReduceOp.MAX, group=_GROUP)
        known = (local, tuple(int(v) for v in m.tolist()))
        _GRAD_SETS[key] = known
    elif any(has and not take for has, take in zip(local, known[1])):
        raise RuntimeError("vqb200.dist.average_gradients: a parameter that had no gradient on any rank when the set was "
                           "agreed has one now on this rank (call vqb200.dist.reset_gradient_sets() on all ranks)")
    grads: List[torch.Tensor] = []
    for p, take in zip(plist, known[1]):
        if not take:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
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
