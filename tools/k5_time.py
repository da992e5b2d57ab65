import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for B in (4096, 65536, 1048576):
    ze = 2.0 * torch.randn(B, 4, 10, device=dev)
    basis = torch.tensor([1, 8, 40, 200], dtype=torch.int32, device=dev)
    ms = timeit(lambda: vqb200.fsq_round(ze, basis, 1000))
    n = B * 10
    zl = torch.randn(B, 10, 10, device=dev)
    ms2 = timeit(lambda: vqb200.lfq_sign(zl, 0.1))
    zh, idx, m2 = vqb200.fsq_round(ze, basis, 1000)
    ref = (torch.round(ze).permute(0, 2, 1) * basis).sum(-1).long()
    zq, loss, idx2, m3 = vqb200.lfq_sign(zl, 0.1)
    ref2 = ((zl > 0).long().permute(0, 2, 1) * (2 ** torch.arange(10, device=dev))).sum(-1)
    print(json.dumps({"B": B, "fsq_us": ms * 1e3, "fsq_GBps": n * 40 / ms / 1e6, "lfq_us": ms2 * 1e3, "lfq_GBps": n * 88 / ms2 / 1e6,
                      "fsq_ok": bool(torch.equal(idx, ref) and torch.equal(zh, torch.round(ze)) and int(m2[0]) == torch.unique(ref).numel()),
                      "lfq_ok": bool(torch.equal(idx2, ref2) and int(m3[1]) == torch.unique(ref2).numel())}))
