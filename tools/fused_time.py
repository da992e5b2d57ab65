"""Development: time one RVQ stage >= 1 (vqb200_vq_assign_residual) at the bench shape."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
from ctypes import c_size_t
lib = _lib.load()
dev = torch.device("cuda:0")
B, T, K, D = (int(os.environ.get("FB", 1000000)), int(os.environ.get("FT", 10)), 1024, 64)
ALGO = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
Wp = 0.3 * torch.randn(K, D, device=dev); W = 0.2 * torch.randn(K, D, device=dev)
stp, st = vqb200.QuantizerState(K, D, dev), vqb200.QuantizerState(K, D, dev)
st.refresh(W)
r_in = 0.5 * torch.randn(B, D, T, device=dev)
idx_prev = vqb200.vq_assign(r_in, Wp, stp, _lib.ASSIGN_TC)
N = B * T
sB, sC, sT = r_in.stride()
s = stream_ptr(dev)
r_out = torch.empty(B, D, T, device=dev)
idx = torch.empty(B, T, dtype=torch.int32, device=dev)
ws = st.assign_workspace(N)
def call():
    check(lib.vqb200_vq_assign_residual(ptr(r_in), B, D, T, sB, sC, sT, ptr(Wp), ptr(idx_prev), K, ptr(r_out), ptr(W), ptr(st.ee),
                                        ptr(st.image), ptr(st.info), K, ptr(idx), ptr(ws), c_size_t(ws.numel()), ALGO, s), "ar")
for _ in range(3): call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): call()
b.record(); torch.cuda.synchronize()
print(json.dumps({"dbg": os.environ.get("VQB200_TC_DEBUG", "0"), "algo": ALGO, "fused_stage_ms": a.elapsed_time(b) / 10}))
