import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K = 1000000, 1, 4096
torch.manual_seed(0)
W = torch.randn(K, 64, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
z = (0.5 * torch.randn(B, T, 64, device=dev)).permute(0, 2, 1)
a = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
for rep in range(3):
    b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
    torch.cuda.synchronize()
    ws = st._assign_ws.view(torch.int32)
    flagged, widen = int(ws[0]), int(ws[4])
    n2 = B
    lst = ws[64:64 + flagged].long()
    wide = ws[64 + 4 * n2: 64 + 4 * n2 + 2 * widen].view(-1, 2)[:, 0].long()
    bad = (a != b).view(-1)
    special = torch.zeros(B, dtype=torch.bool, device=dev); special[lst] = True; special[wide] = True
    print(os.environ.get("VQB200_TC_DEBUG"), "rep", rep, "flagged", flagged, "wide", widen, "bad", int(bad.sum()), "bad among flagged/wide", int((bad & special).sum()),
          "bad others", int((bad & ~special).sum()), "kindbits left", int(((b.view(-1) >> 28) != 0).sum()))
