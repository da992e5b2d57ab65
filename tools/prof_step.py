"""One RVQ training step at a reduced batch (for ncu --set full captures of the HBM-bound kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200, bench
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOADS["cfg3_rvq4_k1024_d64"]); cfg["B"] = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
mod, layers = bench.build_module(vqb200, torch, cfg, dev)
z = (0.5 * torch.randn(cfg["B"], 64, 10, device=dev)).requires_grad_(True)
g = torch.randn(cfg["B"], 64, 10, device=dev); one = torch.ones((), device=dev)
for _ in range(3):
    z.grad = None
    loss, q, met = mod(z)
    torch.autograd.backward([q, loss], [g, one])
torch.cuda.synchronize(); print("ok", float(loss))
