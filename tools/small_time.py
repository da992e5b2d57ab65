import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
torch.manual_seed(0)
def timeit(fn, reps=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for (S, K, N) in [(4, 512, 512), (1, 1024, 512), (4, 512, 4096), (1, 1024, 10)]:
    m = vqb200.ResidualVQ(S, K, 64, use_ema=True).to(dev).train()
    with torch.no_grad():
        for l in m.layers: l.embedding.weight.normal_(0, 0.3); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1)
    z = torch.randn(N, 64, 1, device=dev)
    with torch.no_grad():
        us = timeit(lambda: m(z))
    gs = vqb200.GraphedQuantizerStep(m, z, with_backward=False)
    usg = timeit(lambda: gs(z))
    print(json.dumps({"S": S, "K": K, "N": N, "eager_us": us, "graph_us": usg}))
