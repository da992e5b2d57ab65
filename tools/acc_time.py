"""Development: ema_accumulate (K3a) timing + check against a torch index_add reference.  VQB200_ACC_PRIV=0 -> L2-reduction kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
lib = _lib.load()
dev = torch.device("cuda:0")
B, T, K, D = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (1000000, 10, 1024, 64)
torch.manual_seed(0)
z = 0.5 * torch.randn(B, D, T, device=dev)
idx = torch.randint(0, K, (B, T), dtype=torch.int32, device=dev)
stats = torch.zeros(K * D + K, device=dev)
s = stream_ptr(dev)
sB, sC, sT = z.stride()
def run():
    check(lib.vqb200_ema_accumulate(ptr(z), B, D, T, sB, sC, sT, ptr(idx), None, K, ptr(stats), 0, s), "acc")
run(); torch.cuda.synchronize()
rows = z.permute(0, 2, 1).reshape(-1, D)
ref = torch.zeros(K, D, device=dev, dtype=torch.float64).index_add_(0, idx.view(-1).long(), rows.double())
cnt = torch.bincount(idx.view(-1).long(), minlength=K).double()
got = stats[:K * D].view(K, D).double()
err = float((got - ref).abs().max() / ref.abs().max())
cerr = float((stats[K * D:].double() - cnt).abs().max())
for _ in range(2): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(json.dumps({"priv": os.environ.get("VQB200_ACC_PRIV", "default"), "N": B * T, "K": K, "D": D, "ms": ms, "GBps": B * T * D * 4 / ms / 1e6, "rel_err_sums": err, "count_err": cerr}))
