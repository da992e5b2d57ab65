"""Tiny driver for ncu: a few launches of the tcgen05 assignment kernel on the headline shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (200000, 10, 1024)))
perm = len(sys.argv) > 4 and sys.argv[4] == "perm"
D = int(sys.argv[5]) if len(sys.argv) > 5 else 64
torch.manual_seed(0)
W = 0.3 * torch.randn(K, D, device=dev)
z = (0.5 * torch.randn(B, T, D, device=dev)).permute(0, 2, 1) if perm else 0.5 * torch.randn(B, D, T, device=dev)
st = vqb200.QuantizerState(K, D, dev)
for _ in range(3):
    idx = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
