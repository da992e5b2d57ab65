import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = (int(x) for x in sys.argv[1:4])
ALGO = int(sys.argv[4]) if len(sys.argv) > 4 else _lib.ASSIGN_TC
torch.manual_seed(0)
W = 0.3 * torch.randn(K, 64, device=dev)
z = 0.5 * torch.randn(B, 64, T, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
for _ in range(3): vqb200.vq_assign(z, W, st, ALGO)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): vqb200.vq_assign(z, W, st, ALGO)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(json.dumps({"dbg": os.environ.get("VQB200_TC_DEBUG", "0"), "algo": ALGO, "N": B * T, "K": K, "ms": ms, "scores_per_s": B * T * K / (ms * 1e-3), "flagged": int(st._assign_ws.view(torch.int32)[0]), "wide": int(st._assign_ws.view(torch.int32)[4]), "rerank_list": int(st._assign_ws.view(torch.int32)[5])}))
