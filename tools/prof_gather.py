import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, K, D = 400000, 10, 1024, 64
W = 0.3 * torch.randn(K, D, device=dev); z = 0.5 * torch.randn(B, D, T, device=dev)
idx = torch.randint(0, K, (B, T), dtype=torch.int32, device=dev)
out = torch.empty_like(z); res = torch.empty_like(z); sse = torch.zeros(1, dtype=torch.float64, device=dev)
s = stream_ptr(dev); sB, sC, sT = z.stride()
for _ in range(2):
    check(lib.vqb200_vq_gather_st(ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(idx), K, ptr(out), None, None, 0, ptr(sse), s), "a")
    check(lib.vqb200_vq_gather_st(ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(idx), K, None, ptr(res), ptr(out), 1, ptr(sse), s), "b")
torch.cuda.synchronize(); print("ok")
