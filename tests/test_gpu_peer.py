"""EMA finalize fused with its all-reduce over peer memory (csrc/peer.cu).

* one GPU: two "ranks" emulated in ONE process (two streams, two flag / slot buffers): the kernels' barrier, the
  rank-ordered sum and the finalize arithmetic against `vqb200_ema_finalize` on the pre-summed statistics (bit-exact)
  and against the oracle's update (models/vqvae.py:46-50);
* two or more GPUs (skipped otherwise): real processes, CUDA-IPC mapped buffers, a ResidualVQ training run sharded
  over the ranks, against the NCCL transport and the single-process full-batch run.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _finalize_args(lib, st, cs, w, E, K, D, stream):
    from vqb200._lib import ptr
    return (ptr(cs), ptr(w), ptr(E), K, D, ctypes.c_double(0.99), ctypes.c_double(1e-5), ptr(st["ee"]), ptr(st["image"]),
            ptr(st["info"]), ptr(st["scratch"]), ctypes.c_void_p(stream.cuda_stream))


def _state(lib, K, D, dev):
    nbytes = int(lib.vqb200_codebook_image_bytes(K, D))
    raw = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)     # zeros: the image has alignment gaps nobody writes
    off = (-raw.data_ptr()) % 1024
    return {"ee": torch.empty(K, device=dev), "image": raw[off:off + nbytes], "raw": raw,
            "info": torch.zeros(4, device=dev), "scratch": torch.empty(K + 8, device=dev)}


@pytest.mark.parametrize("K,D,world", [(1024, 64, 2), (512, 64, 4), (300, 128, 3), (4096, 64, 4)])
def test_peer_finalize_in_process_ranks(K, D, world):
    _need_cuda()
    import vqb200
    from vqb200._lib import check, load, ptr
    from oracle import VQState
    lib = load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(K + D + world)
    # per-rank statistics: integer counts, sums of that many N(0,1) rows (any floats do)
    cnts = [torch.randint(0, 50, (K,), generator=g).float() for _ in range(world)]
    dws = [torch.randn(K, D, generator=g) * c[:, None].sqrt() for c in cnts]
    slots = [torch.cat([dw.reshape(-1), c]).to(dev) for dw, c in zip(dws, cnts)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    cs0 = torch.rand(K, generator=g) * 20
    E0 = torch.randn(K, D, generator=g)
    w0 = E0 * cs0[:, None]
    ranks = []
    for r in range(world):
        ranks.append({"cs": cs0.clone().to(dev), "w": w0.clone().to(dev), "E": E0.clone().to(dev),
                      "cnt": torch.empty(K, device=dev), "st": _state(lib, K, D, dev), "stream": torch.cuda.Stream(dev)})
    slot_tab = (ctypes.c_void_p * world)(*[s.data_ptr() for s in slots])
    flag_tab = (ctypes.c_void_p * world)(*[f.data_ptr() for f in flags])
    # two spinning kernels of one process must not wait on a lazily loaded third one: load everything first
    warm = torch.zeros(64, dtype=torch.int32, device=dev)
    check(lib.vqb200_peer_barrier((ctypes.c_void_p * 1)(warm.data_ptr()), 0, 1, ctypes.c_uint32(1), None), "peer_barrier")
    torch.cuda.synchronize()
    for epoch in (1, 2):                                   # second epoch: flags are reused, state has moved on
        for r, R in enumerate(ranks):
            check(lib.vqb200_ema_finalize_peer(slot_tab, flag_tab, r, world, ctypes.c_uint32(epoch), ptr(R["cnt"]),
                                               *_finalize_args(lib, R["st"], R["cs"], R["w"], R["E"], K, D, R["stream"])),
                  "ema_finalize_peer")
        torch.cuda.synchronize()
    # reference 1: the stand-alone finalize on the rank-ordered sum (what NCCL would have delivered), twice
    total = slots[0].clone()
    for s in slots[1:]:
        total = total + s
    ref = {"cs": cs0.clone().to(dev), "w": w0.clone().to(dev), "E": E0.clone().to(dev), "st": _state(lib, K, D, dev)}
    cur = torch.cuda.current_stream(dev)
    for epoch in (1, 2):
        check(lib.vqb200_ema_finalize(ptr(total), *_finalize_args(lib, ref["st"], ref["cs"], ref["w"], ref["E"], K, D, cur)),
              "ema_finalize")
    torch.cuda.synchronize()
    for r, R in enumerate(ranks):
        for key in ("cs", "w", "E"):
            assert torch.equal(R[key], ref[key]), f"rank {r}: {key} differs from finalize(sum of slots)"
        assert torch.equal(R["st"]["ee"], ref["st"]["ee"]) and torch.equal(R["st"]["image"], ref["st"]["image"])
        assert torch.equal(R["cnt"], total[K * D:])
        assert int(flags[r][:world].min()) == 2
    # reference 2: the oracle's update formulas on the summed statistics (models/vqvae.py:46-50)
    o = VQState(E0.numpy().copy(), cs0.numpy().copy(), w0.numpy().copy(), 0.25, True, 0.99)
    cnt_np = total[K * D:].cpu().numpy()
    dw_np = total[:K * D].reshape(K, D).cpu().numpy()
    for epoch in (1, 2):
        o.ema_cluster_size = (o.ema_cluster_size * np.float32(0.99) + np.float32(1 - 0.99) * cnt_np).astype(np.float32)
        o.ema_w = (o.ema_w * np.float32(0.99) + np.float32(1 - 0.99) * dw_np).astype(np.float32)
        n = o.ema_cluster_size.sum(dtype=np.float32)
        cl = ((o.ema_cluster_size + np.float32(1e-5)) / (n + np.float32(K * 1e-5)) * n).astype(np.float32)
        o.embedding = (o.ema_w / cl[:, None]).astype(np.float32)
    np.testing.assert_allclose(ranks[0]["E"].cpu().numpy(), o.embedding, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ranks[0]["cs"].cpu().numpy(), o.ema_cluster_size, rtol=1e-5)


def test_peer_argument_errors_are_loud():
    _need_cuda()
    from vqb200._lib import load, last_error
    lib = load()
    dev = torch.device("cuda:0")
    f = torch.zeros(64, dtype=torch.int32, device=dev)
    tab = (ctypes.c_void_p * 1)(f.data_ptr())
    assert lib.vqb200_peer_barrier(tab, 0, 17, ctypes.c_uint32(1), None) == -2 and "world" in last_error()
    assert lib.vqb200_peer_barrier(tab, 3, 2, ctypes.c_uint32(1), None) == -2
    assert lib.vqb200_peer_barrier(None, 0, 1, ctypes.c_uint32(1), None) == -1
    assert lib.vqb200_peer_open(None, None) == -1


# ------------------------------------------------------------------------------------------------------------------
# real ranks
# ------------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_worker(rank, world, port, tmp):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import vqb200
    S, K, D, B, T = 3, 512, 64, 4096, 10
    torch.manual_seed(5)
    z_full = 0.5 * torch.randn(B, D, T)
    init = vqb200.ResidualVQ(S, K, D, use_ema=True)
    with torch.no_grad():
        for l in init.layers:
            l.embedding.weight.normal_(0, 0.3)
            l.ema_w.copy_(l.embedding.weight)
            l.ema_cluster_size.fill_(1.0)
    sd = {k: v.clone() for k, v in init.state_dict().items()}
    lo, hi = vqb200.dist.shard_bounds(B, rank, world)
    results = {}
    for transport in ("peer", "nccl"):
        vqb200.dist.enable(peer=transport)
        assert vqb200.dist.peer_status().startswith(transport), vqb200.dist.peer_status()
        m = vqb200.ResidualVQ(S, K, D, use_ema=True).to(dev).train()
        m.load_state_dict(sd)
        for step in range(3):
            z = (z_full[lo:hi] * (1.0 + 0.1 * step)).to(dev).requires_grad_(True)
            loss, q, met = m(z)
            (loss + q.square().mean()).backward()
        torch.cuda.synchronize()
        results[transport] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        results[transport]["ppl"] = met["perplexity"].detach().clone()
        # every rank holds bit-identical codebooks without a broadcast
        for k, v in m.state_dict().items():
            gathered = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(gathered, v.contiguous())
            assert all(torch.equal(g, gathered[0]) for g in gathered), f"{transport}: {k} differs across ranks"
    vqb200.dist.disable()
    def rows_off(a, b):
        # the two transports sum in different orders (~1e-7): a near-tie may flip and move one vector between two
        # codes, so compare code rows and allow a handful of them to differ
        a, b = a.float().reshape(max(a.shape[0], 1) if a.dim() else 1, -1), b.float().reshape(max(b.shape[0], 1) if b.dim() else 1, -1)
        return int((~torch.isclose(a, b, rtol=1e-4, atol=1e-5)).any(1).sum().item())

    for k in results["peer"]:
        if k == "ppl":
            assert torch.allclose(results["peer"][k], results["nccl"][k], rtol=1e-3)
            continue
        bad = rows_off(results["peer"][k], results["nccl"][k])
        assert bad <= max(4, results["peer"][k].shape[0] // 16), f"peer vs nccl: {k}: {bad} rows differ"
    if rank == 0:
        # the single-process full-batch run (what the sharded run must equal up to fp32 summation order)
        m = vqb200.ResidualVQ(S, K, D, use_ema=True).to(dev).train()
        m.load_state_dict(sd)
        for step in range(3):
            loss, q, met = m((z_full * (1.0 + 0.1 * step)).to(dev))
        for k, v in m.state_dict().items():
            # a near-tie flip moves one vector between two codes at its stage and re-routes that row in every later
            # stage, so a few dozen code rows may differ slightly after 3 steps x 3 stages; a rank whose statistics
            # were dropped would shift EVERY row by ~1/world
            bad = rows_off(results["peer"][k], v)
            assert bad <= max(4, v.shape[0] // 16), f"sharded vs full batch: {k}: {bad} rows differ"
            if k.endswith("ema_cluster_size"):
                assert torch.allclose(results["peer"][k].float(), v.float(), rtol=0.05), f"sharded vs full batch: {k}"
    # ---- launch-bound shape: the single-launch ResidualVQ with the exchange inside (uniform shards) ----
    S2, K2, Bt = 4, 512, 256 * world                      # 256 vectors per rank: the whole-GPU kernel
    torch.manual_seed(6)                                  # rank 0 drew extra random numbers above
    z2_full = 0.5 * torch.randn(Bt, D, 1)
    init2 = vqb200.ResidualVQ(S2, K2, D, use_ema=True)
    with torch.no_grad():
        for l in init2.layers:
            l.embedding.weight.normal_(0, 0.3)
            l.ema_w.copy_(l.embedding.weight)
            l.ema_cluster_size.fill_(1.0)
    sd2 = {k: v.clone() for k, v in init2.state_dict().items()}
    lo2, hi2 = vqb200.dist.shard_bounds(Bt, rank, world)
    res2 = {}
    for tag, uniform in (("single_launch", True), ("multi_kernel", False)):
        vqb200.dist.enable(peer="peer", uniform_shards=uniform)
        m = vqb200.ResidualVQ(S2, K2, D, use_ema=True).to(dev).train()
        m.load_state_dict(sd2)
        before = vqb200._lib.launch_count()
        for step in range(3):
            z = (z2_full[lo2:hi2] * (1.0 + 0.1 * step)).to(dev).requires_grad_(True)
            loss, q, met = m(z)
            (loss + q.square().mean()).backward()
        torch.cuda.synchronize()
        launches = vqb200._lib.launch_count() - before
        if uniform:
            assert launches <= 3 * 3, f"single-launch path not taken: {launches} launches in 3 steps"
        res2[tag] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        res2[tag]["ppl"] = met["perplexity"].detach().clone()
        for k, v in m.state_dict().items():
            gathered = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(gathered, v.contiguous())
            if not all(torch.equal(g, gathered[0]) for g in gathered):
                d = (gathered[-1].float() - gathered[0].float()).abs()
                raise AssertionError(f"{tag}: {k} differs across ranks: {int((d > 0).sum())} of {d.numel()} elements, "
                                     f"max |diff| {float(d.max()):.3e}, rows {(d.reshape(d.shape[0], -1) > 0).any(1).nonzero().flatten()[:8].tolist()}")
    vqb200.dist.disable()
    for k in res2["single_launch"]:
        if k == "ppl":
            assert torch.allclose(res2["single_launch"][k], res2["multi_kernel"][k], rtol=2e-3)
            continue
        bad = rows_off(res2["single_launch"][k], res2["multi_kernel"][k])
        assert bad <= max(4, res2["single_launch"][k].shape[0] // 16), f"single-launch vs multi-kernel under DP: {k}: {bad} rows differ"
    open(os.path.join(tmp, f"ok{rank}"), "w").write("peer")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_exchange_real_ranks(world, tmp_path):
    _need_cuda()
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_rank_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _trainer_worker(rank, world, port, tmp):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "LOCAL_RANK": str(rank),
                       "WORLD_SIZE": str(world)})
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import vqb200
    from vqb200 import trainer
    vqb200.dist.enable(peer="auto")
    argv = ["--mode", "teacher", "--arch", "transformer", "--method", "hybrid", "--window", "10", "--epochs", "2",
            "--batch_size", "128", "--synthetic", "1000", "--data_root", os.path.join(tmp, "nodata"),
            "--ckpt_dir", os.path.join(tmp, "ck"), "--log_dir", os.path.join(tmp, "res"), "--name", "t"]
    args = trainer.build_parser().parse_args(argv)
    hist = trainer.train_one_seed(args, 42, dev)
    assert len(hist["train_loss"]) == 2 and all(v == v for v in hist["train_loss"])
    model = trainer.train_one_seed.last_model
    for k, v in model.state_dict().items():          # replicas stayed identical without any parameter broadcast
        if "num_batches_tracked" in k:
            continue
        gathered = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(gathered, v.contiguous())
        assert all(torch.equal(g, gathered[0]) for g in gathered), f"{k} differs across ranks"
    open(os.path.join(tmp, f"ok{rank}"), "w").write(vqb200.dist.peer_status())
    vqb200.dist.disable()
    dist.destroy_process_group()


def test_ddp_trainer_real_ranks(tmp_path):
    _need_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_trainer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").read_text() == "peer"
    assert (tmp_path / "ck" / "t_hybrid_teacher_seed_42_final.pth").exists()
