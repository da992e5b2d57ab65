import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = (int(x) for x in sys.argv[1:4])
torch.manual_seed(0)
W = 0.3 * torch.randn(K, 64, device=dev)
z = 0.5 * torch.randn(B, 64, T, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
a = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
torch.cuda.synchronize()
print("mismatch", int((a != b).sum()), st._assign_ws.view(torch.int32)[:8].tolist())
