"""Eager steps of the two latency-bound configs (cfg1 EMA-VQ N=40960, cfg2 Hybrid N=512) for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
torch.manual_seed(42)
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
one = torch.ones((), device=dev)
if which == "cfg2":
    m = vqb200.HybridVQ(64, [8, 5, 5, 5], vq_codebook_size=512).to(dev).train()
    z = torch.randn(512, 1, 64, device=dev).permute(0, 2, 1).requires_grad_(True)
    g = torch.randn(512, 64, 1, device=dev)
else:
    m = vqb200.VectorQuantizer(1024, 64, use_ema=True).to(dev).train()
    with torch.no_grad():
        m.embedding.weight.normal_(0, 0.3); m.ema_w.copy_(m.embedding.weight); m.ema_cluster_size.fill_(1)
    z = torch.randn(4096, 64, 10, device=dev).requires_grad_(True)
    g = torch.randn(4096, 64, 10, device=dev)
for _ in range(4):
    z.grad = None
    loss, q, _ = m(z)
    torch.autograd.backward([q, loss], [g, one])
torch.cuda.synchronize()
print("ok")
