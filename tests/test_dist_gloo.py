"""world_size-2 gloo test of the data-parallel host logic (no GPU): sharding, the packed EMA-stats
all-reduce and gradient averaging.  The per-rank statistics are produced by the oracle (tests may use
it); the point is that reduce-then-finalize over shards equals the single-process full-batch update
(SURVEY.md §8e), which is the semantics the CUDA path implements with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vqb200
    from oracle import VQState, vq_forward
    vqb200.dist.enable()
    assert vqb200.dist.world_size() == world and vqb200.dist.rank() == rank

    rng = np.random.default_rng(0)
    K, D, B, T = 32, 8, 10, 3
    E = rng.standard_normal((K, D)).astype(np.float32)
    w = rng.standard_normal((K, D)).astype(np.float32)
    z = rng.standard_normal((B, D, T)).astype(np.float32)
    full = VQState(E.copy(), np.zeros(K, np.float32), w.copy(), 0.25, True, 0.99)
    vq_forward(z, full, True)                                   # single-process full-batch oracle

    lo, hi = vqb200.dist.shard_bounds(B)
    covered = torch.zeros(B)
    covered[lo:hi] = 1
    dist.all_reduce(covered)
    assert torch.equal(covered, torch.ones(B))                  # shards tile the batch exactly once

    def reduce(cnt, dw):                                        # what RVQ stage s does between K3a and K3b
        stats = torch.from_numpy(np.concatenate([dw.reshape(-1), cnt]))
        vqb200.dist.all_reduce_stats(stats)
        s = stats.numpy()
        return s[K * D:].copy(), s[:K * D].reshape(K, D).copy()

    mine = VQState(E.copy(), np.zeros(K, np.float32), w.copy(), 0.25, True, 0.99)
    vq_forward(z[lo:hi], mine, True, stats_reduce=reduce)
    np.testing.assert_allclose(mine.embedding, full.embedding, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mine.ema_cluster_size, full.ema_cluster_size, rtol=1e-6)

    # codebooks are bit-identical across ranks without a broadcast
    e = torch.from_numpy(mine.embedding.copy())
    gathered = [torch.empty_like(e) for _ in range(world)]
    dist.all_gather(gathered, e)
    assert all(torch.equal(g, gathered[0]) for g in gathered)

    # DDP-style gradient averaging
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(rank + 1))
    q = torch.nn.Parameter(torch.zeros(3))                      # no grad: skipped
    calls = vqb200.dist.average_gradients([p, q])
    assert calls == 1 and torch.allclose(p.grad, torch.full((5,), (1 + world) / 2))
    assert q.grad is None                                       # no rank has one: stays None (EMA codebooks)
    # a parameter that only SOME ranks have a gradient for: the others contribute zeros and receive the average
    u = torch.nn.Parameter(torch.zeros(4))
    v = torch.nn.Parameter(torch.zeros(2))
    v.grad = torch.ones(2)
    if rank == 0:
        u.grad = torch.full((4,), 2.0)
    calls = vqb200.dist.average_gradients([u, v])
    assert calls == 1 and torch.allclose(u.grad, torch.full((4,), 2.0 / world)) and torch.allclose(v.grad, torch.ones(2))
    u.grad = None if rank == 1 else u.grad                      # zero_grad(set_to_none=True): same agreed set, still fine
    assert vqb200.dist.average_gradients([u, v]) == 1
    q.grad = torch.ones(3)                                      # outside the agreed set of [p, q]: loud, not silently skipped
    try:
        vqb200.dist.average_gradients([p, q])
        raise AssertionError("expected a RuntimeError")
    except RuntimeError as e:
        assert "had no gradient on any rank" in str(e)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_stats_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_bounds_cover():
    import vqb200
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            b = [vqb200.dist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_peer_exchange_slot_and_epoch_bookkeeping():
    """Host-side protocol of the peer-memory exchange (dist.PeerExchange.next_slot) without a GPU: slots alternate per
    USE, a use consumes as many barrier epochs as it has stages, epochs never repeat and wrap at 2^32."""
    import ctypes
    import vqb200
    from vqb200.dist import PeerExchange, FLAG_BYTES
    px = PeerExchange.__new__(PeerExchange)
    px.epoch, px.uses, px.slot_bytes, px._own = 1, 0, 1024, 1 << 20           # epoch 1 = the handshake
    px.base = [1 << 20, 2 << 20]
    px._slots = [(ctypes.c_void_p * 2)(*[b + FLAG_BYTES + s * px.slot_bytes for b in px.base]) for s in range(2)]
    seen = []
    e, mine, tab = px.next_slot()             # a finalize call: one epoch
    seen.append((e, mine)); assert e == 2 and tab[0] == mine
    e, mine, tab = px.next_slot(4)            # a single-launch ResidualVQ with 4 stages: epochs 3..6
    seen.append((e, mine)); assert e == 3 and px.epoch == 6
    e, mine, tab = px.next_slot()
    seen.append((e, mine)); assert e == 7
    slots = [m for _, m in seen]
    assert slots[0] != slots[1] and slots[0] == slots[2]                       # alternate per use, not per epoch
    assert {m - px._own - FLAG_BYTES for m in slots} == {0, px.slot_bytes}
    px.epoch = 0xFFFFFFFE
    e, _, _ = px.next_slot(4)
    assert e == 0xFFFFFFFF and px.epoch == 2                                    # wraps; kernels compare (int)(seen - epoch) >= 0
    assert px.fits(3, 64) and not px.fits(4, 64)
