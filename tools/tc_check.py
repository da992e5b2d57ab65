"""Development check of the tcgen05 assignment path: TC (algo=2) vs exact SIMT (algo=1)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import _lib

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


def case(B, T, K, regime="small", seed=0, time_it=False, perm=False, D=64, zmode="plain"):
    torch.manual_seed(seed)
    W = torch.randn(K, D, device=dev)
    if regime == "small":
        W *= 0.3
    elif regime == "init":
        W = (torch.rand(K, D, device=dev) * 2 - 1) / K
    elif regime == "degenerate":
        W[: K // 2] *= 1e5
    elif regime == "dup":
        W[K // 2:] = W[: K - K // 2]
    elif regime == "dead":               # a few live codes among huge dead ones (the reference after its first EMA steps)
        W[::3] *= 3e4
    elif regime == "tiny":
        W *= 1e-4
    st = vqb200.QuantizerState(K, D, dev)
    if perm:
        z = (0.5 * torch.randn(B, T, D, device=dev)).permute(0, 2, 1)
    else:
        z = 0.5 * torch.randn(B, D, T, device=dev)
    if zmode == "rowscales":             # every row at its own magnitude, 1e-18 .. 1e18
        z = z * (10.0 ** (36 * torch.rand(B, 1, T, device=dev) - 18))
    elif zmode == "huge":
        z = z * 1e17; W = W * 1e17
    elif zmode == "zeros":               # a third of the rows exactly zero, some components exactly zero
        z = z * (torch.rand(B, 1, T, device=dev) > 0.33) * (torch.rand_like(z) > 0.2)
    elif zmode == "nonfinite":           # NaN / Inf rows and one NaN code: both paths must agree (NaN wins, first index)
        z = z.clone()
        z.view(-1)[::977] = float("nan"); z.view(-1)[5::1999] = float("inf"); z.view(-1)[11::2999] = float("-inf")
        W = W.clone(); W[K // 3, 7] = float("nan")
    i_simt = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
    i_tc = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
    torch.cuda.synchronize()
    off = (-st._assign_ws.data_ptr()) % 1024 if D != 64 else 0
    ws = st._assign_ws[off:off + 256].view(torch.int32)
    flagged, err, two, wide, rr = int(ws[0]), int(ws[1]), int(ws[2]) + int(ws[3]), int(ws[4]), int(ws[5])
    mism = int((i_simt != i_tc).sum())
    out = dict(B=B, T=T, K=K, D=D, regime=regime, zmode=zmode, perm=perm, N=B * T, mismatches=mism, flagged=flagged,
               multi_groups=two, wide=wide, rerank_list=rr, err=err)
    if mism:
        bad = (i_simt != i_tc).reshape(-1).nonzero().reshape(-1)[:5]
        out["first_bad_rows"] = bad.tolist()
        out["simt"] = i_simt.reshape(-1)[bad].tolist()
        out["tc"] = i_tc.reshape(-1)[bad].tolist()
    if time_it:
        out["ms_tc"] = timeit(lambda: vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC))
        if D == 64:
            out["ms_tc_split"] = timeit(lambda: vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC_SPLIT))
        out["ms_simt"] = timeit(lambda: vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT), reps=2, warm=1)
        fl = 2.0 * B * T * K * D
        out["tc_TFLOPs"] = fl / (out["ms_tc"] * 1e-3) / 1e12
    print(json.dumps(out), flush=True)
    return mism


def edge_cases():
    total = 0
    total += case(1, 1, 1024)                          # a single row
    total += case(1, 129, 1024, perm=True)             # one row more than a tile
    total += case(13, 3, 777)                          # ragged everything
    total += case(4099, 10, 1030)                      # streaming kernel, 6 codes in the last tile
    total += case(20000, 10, 4099, "normal")
    total += case(20000, 1, 1025, perm=True)
    total += case(40960, 10, 1024, zmode="rowscales")
    total += case(40960, 10, 2048, zmode="rowscales")
    total += case(40960, 10, 1024, "normal", zmode="huge")
    total += case(40960, 10, 1024, zmode="zeros")
    total += case(40960, 10, 4096, zmode="zeros")
    total += case(40960, 10, 1024, zmode="nonfinite")
    total += case(40960, 10, 2048, zmode="nonfinite")
    total += case(40960, 1, 1024, "dead", zmode="rowscales", perm=True)
    total += case(300000, 10, 1024, "dup", zmode="rowscales")
    print("TOTAL MISMATCHES (edge)", total)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "edge":
        edge_cases()
        sys.exit(0)
    total = 0
    total += case(2, 128, 128)                 # one CTA tile, one code tile
    total += case(1, 256, 256, perm=True)
    total += case(100, 10, 1024)
    total += case(4096, 10, 1024, time_it=True)
    total += case(4096, 10, 1000, "normal")
    total += case(512, 1, 512, "init", perm=True)
    total += case(3686, 10, 1024, "degenerate")
    total += case(3686, 10, 1024, "dead")
    total += case(3686, 10, 1024, "tiny")
    total += case(3686, 10, 1023, "small")
    total += case(3686, 10, 5, "small")
    total += case(4096, 10, 512, "init")
    total += case(4096, 10, 256, "small")
    total += case(4096, 10, 128, "normal")
    total += case(50000, 1, 2048, "small", perm=True)
    total += case(3000, 7, 300, "dup")
    total += case(100000, 10, 1024, time_it=True)
    total += case(1000000, 10, 1024, time_it=True)
    total += case(1000000, 1, 4096, "normal", time_it=True, perm=True)
    total += case(300, 10, 256, D=128)
    total += case(4096, 10, 1024, "normal", D=128, time_it=True)
    total += case(1000, 1, 512, D=256, perm=True)
    total += case(3686, 10, 1024, "degenerate", D=256)
    total += case(1000000, 1, 4096, "normal", D=128, time_it=True, perm=True)
    total += case(500000, 1, 4096, "normal", D=256, time_it=True, perm=True)
    total += case(100000, 10, 2048, "small", D=256, time_it=True)
    total += case(700, 3, 300, "small", D=512)
    total += case(3686, 1, 1024, "degenerate", D=512, perm=True)
    total += case(262144, 1, 4096, "normal", D=512, time_it=True, perm=True)
    print("TOTAL MISMATCHES", total)
