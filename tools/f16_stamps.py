"""Development: timeline of CTA 0 of the resident fp16 assignment kernel from its clock64 stamps (VQB200_TC_DEBUG=520)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VQB200_TC_DEBUG"] = os.environ.get("VQB200_TC_DEBUG", "520")
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = 1000000, 10, int(os.environ.get("STAMP_K", "1024"))
torch.manual_seed(0)
W = 0.3 * torch.randn(K, 64, device=dev)
z = 0.5 * torch.randn(B, 64, T, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
for _ in range(2):
    vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
torch.cuda.synchronize()
N = B * T
ws = st._assign_ws.view(torch.int32)
n2 = (N + 1) & ~1
c3 = ws[64 + 2 * n2: 64 + 2 * n2 + 5 * 1024 * 4 * 2].view(torch.int64).cpu().view(5, 1024, 4)
mma, e0, e1, cv, pr = c3[0], c3[1], c3[2], c3[3], c3[4]
t0 = int(mma[0, 0])
print("MMA commit stamps per slot (rel cycles, delta, unit):")
for g in range(2):
    for i in range(24, 40):
        k = g * 512 + i
        print(f"  slot{g} committed={int(mma[k,0])-t0:8d} d={int(mma[k,0]-mma[k-1,0]):6d} unit={int(mma[k,3])//16}/{int(mma[k,3])%16}")
for name, e in (("epi g0", e0), ("epi g1", e1)):
    print(name, "(wait_start, wait_end, done; job*16+tile):")
    for i in range(24, 44):
        print(f"  start={int(e[i,0])-t0:8d} waited={int(e[i,1]-e[i,0]):6d} work={int(e[i,2]-e[i,1]):6d} unit={int(e[i,3])//16}/{int(e[i,3])%16}")
print("converter (job, wait, work):")
for i in range(2, 14):
    print(f"  job={int(cv[i,3])} start={int(cv[i,0])-t0:8d} waited={int(cv[i,1]-cv[i,0]):6d} work={int(cv[i,2]-cv[i,1]):6d}")
print("producer raw issue (job, t):", [int(pr[i,0]) - t0 for i in range(2, 14)])
