import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = 1000000, 1, 4096
torch.manual_seed(0)
W = torch.randn(K, 64, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
z = (0.5 * torch.randn(B, T, 64, device=dev)).permute(0, 2, 1)
outs = []
for rep in range(3):
    b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC).clone()
    torch.cuda.synchronize()
    outs.append(b.view(-1))
d = [(int((outs[0] != o).sum())) for o in outs[1:]]
diff = (outs[0] != outs[1]).nonzero().view(-1)[:10].tolist()
print(os.environ["VQB200_TC_DEBUG"], "differs:", d, [((int(outs[0][i]) & 0xfffffff) >> 5, (int(outs[1][i]) & 0xfffffff) >> 5) for i in diff], "rt of rows", [(i % 256) // 128 for i in diff])
