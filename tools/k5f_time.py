"""FSQ / LFQ modules: one-pass kernels with the 1x1 projections fused in vs stock conv1d + elementwise kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
T = 10
torch.manual_seed(0)
z = (2.0 * torch.randn(B, 64, T, device=dev))
g = torch.randn(B, 64, T, device=dev)
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, mod in (("fsq", vqb200.FSQ([8, 5, 5, 5], 64, 64).to(dev)), ("lfq", vqb200.LFQ(64, 10).to(dev))):
    res = {"module": name, "N": B * T}
    for fused in (True, False):
        mod.fuse_projections = fused
        zz = z.clone().requires_grad_(True)
        def fwd():
            with torch.no_grad(): mod(z)
        def fwdbwd():
            loss, out, _ = mod(zz)
            if loss.requires_grad: torch.autograd.backward([out, loss], [g, torch.ones_like(loss)])
            else: out.backward(g)
        tf, tfb = timeit(fwd), timeit(fwdbwd)
        tag = "fused" if fused else "unfused"
        res[tag + "_fwd_ms"] = round(tf, 4); res[tag + "_fwdbwd_ms"] = round(tfb, 4)
        if fused:
            res["fwd_GBps"] = round(B * T * (8 * 64 + 4 * mod.project_in.out_channels + 8) / tf / 1e6, 1)
    print(json.dumps(res))
