"""Per-kernel time breakdown of one EMA-VQ stage (development aid, run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
from ctypes import c_double, c_float, c_size_t

lib = _lib.load()
dev = torch.device("cuda:0")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main(B=1_000_000, T=10, K=1024, D=64, algo=0):
    N = B * T
    mod = vqb200.VectorQuantizer(K, D, use_ema=True).to(dev)
    with torch.no_grad():
        mod.embedding.weight.normal_(0, 0.25); mod.ema_w.copy_(mod.embedding.weight); mod.ema_cluster_size.fill_(1)
    st = mod._state(dev)
    W = mod.embedding.weight.detach()
    st.refresh(W)
    z = 0.5 * torch.randn(B, D, T, device=dev)
    g = torch.randn(B, D, T, device=dev)
    idx = torch.empty(B, T, dtype=torch.int32, device=dev)
    out = torch.empty_like(z); res = torch.empty_like(z); gz = torch.empty_like(z)
    m3 = torch.empty(3, device=dev)
    gl = torch.ones((), device=dev)
    s = stream_ptr(dev)
    sB, sC, sT = z.stride()
    ws = st.assign_workspace(N)
    r = {}
    r["assign"] = timeit(lambda: check(lib.vqb200_vq_assign(ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(st.ee), ptr(st.image), ptr(st.info), K, ptr(idx), None, ptr(ws), c_size_t(ws.numel()), algo, s), "a"))
    r["accumulate"] = timeit(lambda: check(lib.vqb200_ema_accumulate(ptr(z), B, D, T, sB, sC, sT, ptr(idx), None, K, ptr(st.stats), 0, s), "b"))
    r["finalize"] = timeit(lambda: check(lib.vqb200_ema_finalize(ptr(st.stats), ptr(mod.ema_cluster_size), ptr(mod.ema_w), ptr(W), K, D, c_double(0.99), c_double(1e-5), ptr(st.ee), ptr(st.image), ptr(st.info), ptr(st.scratch), s), "c"))
    r["gather_st(out)"] = timeit(lambda: check(lib.vqb200_vq_gather_st(ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(idx), K, ptr(out), None, None, 0, ptr(st.sse), s), "d"))
    r["gather_st(res+acc)"] = timeit(lambda: check(lib.vqb200_vq_gather_st(ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(idx), K, None, ptr(res), ptr(out), 1, ptr(st.sse), s), "d"))
    r["metrics"] = timeit(lambda: check(lib.vqb200_vq_metrics(ptr(st.cnt), K, N, ptr(st.sse), N * D, c_float(0.25), 1, ptr(m3), s), "e"))
    r["backward_input"] = timeit(lambda: check(lib.vqb200_vq_backward_input(ptr(g), sB, sC, sT, ptr(z), B, D, T, sB, sC, sT, ptr(W), ptr(idx), K, ptr(gl), c_float(0.5 / (N * D)), ptr(gz), s), "f"))
    r["histogram"] = timeit(lambda: check(lib.vqb200_vq_histogram(ptr(idx), N, K, ptr(st.cnt), s), "g"))
    r["copy(z->out) torch"] = timeit(lambda: out.copy_(z))
    gb = N * D * 4 / 1e6
    print(json.dumps({"N": N, "K": K, "D": D, "T": T, "ms": {k: round(v, 4) for k, v in r.items()}, "tensor_MB": gb}))


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
