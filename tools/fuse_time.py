"""Time one RVQ stage >= 1: stand-alone residual kernel + assignment vs. the fused vqb200_vq_assign_residual."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
dev = torch.device("cuda:0")
B, T, K = (int(x) for x in sys.argv[1:4])
torch.manual_seed(0)
lib = _lib.load()
W0 = 0.3 * torch.randn(K, 64, device=dev)
W1 = 0.1 * torch.randn(K, 64, device=dev)
z = 0.5 * torch.randn(B, 64, T, device=dev)
st0 = vqb200.QuantizerState(K, 64, dev); st1 = vqb200.QuantizerState(K, 64, dev)
idx0 = vqb200.vq_assign(z, W0, st0, _lib.ASSIGN_TC)
st1.refresh(W1)
r1 = torch.empty_like(z); r2 = torch.empty_like(z)
idx1 = torch.empty((B, T), dtype=torch.int32, device=dev); idx2 = torch.empty_like(idx1)
ws = st1.assign_workspace(B * T)
sse = torch.zeros(1, dtype=torch.float64, device=dev)
sB, sC, sT = z.stride()
s = stream_ptr(dev)
def unfused():
    check(lib.vqb200_vq_gather_st(ptr(z), B, 64, T, sB, sC, sT, ptr(W0), ptr(idx0), K, None, ptr(r1), None, 0, ptr(sse), s), "g")
    check(lib.vqb200_vq_assign(ptr(r1), B, 64, T, sB, sC, sT, ptr(W1), ptr(st1.ee), ptr(st1.image), ptr(st1.info), K,
                               ptr(idx1), None, ptr(ws), ctypes.c_size_t(ws.numel()), _lib.ASSIGN_TC, s), "a")
def fused():
    check(lib.vqb200_vq_assign_residual(ptr(z), B, 64, T, sB, sC, sT, ptr(W0), ptr(idx0), K, ptr(r2), ptr(W1), ptr(st1.ee),
                                        ptr(st1.image), ptr(st1.info), K, ptr(idx2), ptr(ws), ctypes.c_size_t(ws.numel()),
                                        _lib.ASSIGN_TC, s), "f")
def timeit(f):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
tu, tf = timeit(unfused), timeit(fused)
dbg = int(os.environ.get("VQB200_TC_DEBUG", "0"))
ok = bool(torch.equal(r1, r2)) and bool(torch.equal(idx1, idx2)) if dbg == 0 else None
print(json.dumps({"dbg": dbg, "N": B * T, "K": K, "ms_unfused": tu, "ms_fused": tf, "identical": ok}))
