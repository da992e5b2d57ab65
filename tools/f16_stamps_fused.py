"""Development: converter / epilogue timeline of the resident kernel in the fused-residual mode (build with
VQB200_NVCC_EXTRA=-DVQB200_K1_DEBUG, run with VQB200_TC_DEBUG=520)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VQB200_TC_DEBUG"] = os.environ.get("VQB200_TC_DEBUG", "520")
import torch, vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
from ctypes import c_size_t
lib = _lib.load()
dev = torch.device("cuda:0")
B, T, K, D = 1000000, 10, 1024, 64
torch.manual_seed(0)
Wp = 0.3 * torch.randn(K, D, device=dev); W = 0.2 * torch.randn(K, D, device=dev)
stp, st = vqb200.QuantizerState(K, D, dev), vqb200.QuantizerState(K, D, dev)
st.refresh(W)
r_in = 0.5 * torch.randn(B, D, T, device=dev)
idx_prev = vqb200.vq_assign(r_in, Wp, stp, _lib.ASSIGN_SIMT)
N = B * T
sB, sC, sT = r_in.stride()
s = stream_ptr(dev)
r_out = torch.empty(B, D, T, device=dev)
idx = torch.empty(B, T, dtype=torch.int32, device=dev)
ws = st.assign_workspace(N)
for _ in range(2):
    check(lib.vqb200_vq_assign_residual(ptr(r_in), B, D, T, sB, sC, sT, ptr(Wp), ptr(idx_prev), K, ptr(r_out), ptr(W), ptr(st.ee),
                                        ptr(st.image), ptr(st.info), K, ptr(idx), ptr(ws), c_size_t(ws.numel()), 2, s), "ar")
torch.cuda.synchronize()
wsi = ws.view(torch.int32)
n2 = (N + 1) & ~1
c3 = wsi[64 + 2 * n2: 64 + 2 * n2 + 5 * 1024 * 4 * 2].view(torch.int64).cpu().view(5, 1024, 4)
mma, e0, e1, cv, pr = c3
t0 = int(mma[0, 0])
print("converter (job, start, waited, work):")
for i in range(4, 16):
    print(f"  job={int(cv[i,3])} start={int(cv[i,0])-t0:8d} waited={int(cv[i,1]-cv[i,0]):6d} work={int(cv[i,2]-cv[i,1]):6d}")
for name, e in (("epi g0", e0),):
    for i in range(24, 34):
        print(f"  {name} start={int(e[i,0])-t0:8d} waited={int(e[i,1]-e[i,0]):6d} work={int(e[i,2]-e[i,1]):6d} unit={int(e[i,3])//16}/{int(e[i,3])%16}")
