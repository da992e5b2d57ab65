import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VQB200_TC_DEBUG"] = "8"
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
for (B, T, K, perm) in ((1000000, 1, 4096, True), (1000000, 1, 2048, True), (400000, 10, 4096, False), (1000000, 1, 1024, True)):
    torch.manual_seed(0)
    W = torch.randn(K, 64, device=dev)
    st = vqb200.QuantizerState(K, 64, dev)
    z = (0.5 * torch.randn(B, T, 64, device=dev)).permute(0, 2, 1) if perm else 0.5 * torch.randn(B, 64, T, device=dev)
    outs = []
    for rep in range(4):
        b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC).clone()
        torch.cuda.synchronize()
        outs.append(b.view(-1))
    d = [(int((outs[0] != o).sum())) for o in outs[1:]]
    diff = (outs[0] != outs[1]).nonzero().view(-1)[:6].tolist()
    print(B, T, K, "raw filter output differs from rep 0 in rows:", d, "e.g.", diff,
          [(hex(int(outs[0][i]) & 0xffffffff), hex(int(outs[1][i]) & 0xffffffff)) for i in diff[:4]])
