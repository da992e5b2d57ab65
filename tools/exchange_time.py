"""Per-stage EMA-statistics exchange under data parallelism: NCCL all-reduce between K3a and K3b vs the
peer-memory finalize (csrc/peer.cu).  Run with torchrun on >= 2 GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/exchange_time.py
Times (a) the exchange + finalize alone on idle GPUs and (b) a latency-class RVQ training step (cfg1-like shard)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vqb200
from vqb200._lib import load, ptr, check

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = load()


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world}
for K, D in [(512, 64), (1024, 64), (4096, 64)]:
    st = vqb200.functional.QuantizerState(K, D, dev)
    cs = torch.ones(K, device=dev); w = torch.randn(K, D, device=dev); E = torch.randn(K, D, device=dev)
    st.stats.normal_()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    tail = (ptr(cs), ptr(w), ptr(E), K, D, ctypes.c_double(0.99), ctypes.c_double(1e-5), ptr(st.ee), ptr(st.image), ptr(st.info),
            ptr(st.scratch), stream)

    vqb200.dist.enable(peer="nccl")
    def nccl_path():
        vqb200.dist.all_reduce_stats(st.stats)
        check(lib.vqb200_ema_finalize(ptr(st.stats), *tail), "ema_finalize")
    t_nccl = timeit(nccl_path)
    vqb200.dist.enable(peer="peer")
    px = vqb200.dist.peer_exchange()
    def peer_path():
        epoch, mine, slots = px.next_slot()
        check(lib.vqb200_ema_finalize_peer(slots, px.flags, px.rank, px.world, ctypes.c_uint32(epoch), ptr(st.cnt), *tail), "peer")
    t_peer = timeit(peer_path)
    out[f"exchange+finalize K={K} D={D}"] = {"nccl_us": t_nccl, "peer_us": t_peer}

# a latency-class sharded step: cfg1-like (K=1024, EMA), 40 960 vectors over the ranks
for transport in ("nccl", "peer"):
    vqb200.dist.enable(peer=transport)
    torch.manual_seed(3)
    m = vqb200.ResidualVQ(4, 1024, 64, use_ema=True).to(dev).train()
    with torch.no_grad():
        for l in m.layers:
            l.embedding.weight.normal_(0, 0.3); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1)
    z = 0.5 * torch.randn(4096 // world, 64, 10, device=dev)
    with torch.no_grad():
        out[f"rvq4 k1024 step, 40960 vectors total, {transport}"] = {"us": timeit(lambda: m(z), reps=50, warm=5)}
# cfg2-like sharded step: RVQ 4 x K=512 on 512 vectors PER RANK -- multi-kernel path with the peer finalize vs the
# single-launch kernel with the exchange inside (uniform shards)
for tag, uniform in (("multi-kernel + peer finalize", False), ("single launch, exchange inside", True)):
    vqb200.dist.enable(peer="peer", uniform_shards=uniform)
    torch.manual_seed(4)
    m = vqb200.ResidualVQ(4, 512, 64, use_ema=True).to(dev).train()
    with torch.no_grad():
        for l in m.layers:
            l.embedding.weight.normal_(0, 0.3); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1)
    z = 0.5 * torch.randn(512, 64, 1, device=dev)
    with torch.no_grad():
        out[f"rvq4 k512 step, 512 vectors per rank, {tag}"] = {"us": timeit(lambda: m(z), reps=100, warm=10)}
vqb200.dist.disable()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
