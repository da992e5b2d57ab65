"""Development: the two launch-bound configs (cfg1 EMA-VQ [4096,64,10] K=1024, cfg2 Hybrid N=512).
usage: small_cfg.py time            -> eager + CUDA-graph microseconds per step
       small_cfg.py trace cfg1|cfg2 -> 3 graph replays only (run under ncu for the kernel list)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
one = torch.ones((), device=dev)

def make(which):
    torch.manual_seed(42)
    if which == "cfg1":
        mod = vqb200.VectorQuantizer(1024, 64, use_ema=True)
        with torch.no_grad():
            mod.embedding.weight.normal_(0, 0.25); mod.ema_w.copy_(mod.embedding.weight); mod.ema_cluster_size.fill_(1.0)
        mod = mod.to(dev).train()
        z = torch.randn(4096, 64, 10, device=dev).requires_grad_(True)
        g = torch.randn_like(z)
    else:
        mod = vqb200.HybridVQ(64, [8, 5, 5, 5], vq_codebook_size=512).to(dev).train()
        z = torch.randn(512, 1, 64, device=dev).permute(0, 2, 1).requires_grad_(True)
        g = torch.randn(512, 64, 1, device=dev)
    return mod, z, g

def timeit(fn, reps, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

mode = sys.argv[1]
if mode == "time":
    out = {}
    for which in ("cfg1", "cfg2"):
        mod, z, g = make(which)
        def step():
            z.grad = None
            loss, q, _ = mod(z)
            torch.autograd.backward([q, loss], [g, one])
        l0 = vqb200._lib.launch_count()
        us = timeit(step, 50)
        out[which + "_eager_us"] = us
        out[which + "_launches"] = (vqb200._lib.launch_count() - l0) / 55
        gs = vqb200.GraphedQuantizerStep(mod, z.detach())
        out[which + "_graph_us"] = timeit(lambda: gs(z.detach(), g), 200)
    print(json.dumps(out))
else:
    mod, z, g = make(sys.argv[2])
    gs = vqb200.GraphedQuantizerStep(mod, z.detach())
    torch.cuda.synchronize()
    for _ in range(3): gs(z.detach(), g)
    torch.cuda.synchronize()
