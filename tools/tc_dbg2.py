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
b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
torch.cuda.synchronize()
ws = st._assign_ws.view(torch.int32)
N = B * T
flagged = int(ws[0])
lst = ws[64:64 + flagged].cpu()
bad = (a != b).view(-1).nonzero().view(-1).cpu()
print("flagged", flagged, "bad", len(bad), "bad in flagged list:", int(torch.isin(bad, lst.long()).sum()))
print("bad rows", bad[:8].tolist(), "simt", a.view(-1)[bad[:8].to(dev)].tolist(), "tc", b.view(-1)[bad[:8].to(dev)].tolist())
print("list head", lst[:8].tolist(), "sorted?", bool((lst[1:] >= lst[:-1]).all()))
# raw filter output (kind bits) of the bad rows: second process-level call with the filter-only knob is not possible
# (the knob is read once), so decode from a fresh library load in a subprocess
import subprocess
code = r'''
import sys, os
sys.path.insert(0, %r)
os.environ["VQB200_TC_DEBUG"] = "8"
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
K = 4096
W = torch.randn(K, 64, device=dev)
st = vqb200.QuantizerState(K, 64, dev)
z = (0.5 * torch.randn(1000000, 1, 64, device=dev)).permute(0, 2, 1)
a = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
b = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
torch.cuda.synchronize()
rows = %r
raw = b.view(-1)[torch.tensor(rows, device=dev)].tolist()
ws = st._assign_ws.view(torch.int32)
n2 = 1000000
c2 = ws[64 + n2: 64 + 2 * n2]; c3 = ws[64 + 2 * n2: 64 + 3 * n2]
wide_n = int(ws[4]); wide = ws[64 + 4 * n2: 64 + 4 * n2 + 2 * wide_n].view(-1, 2).cpu()
for r, v in zip(rows, raw):
    v &= 0xffffffff
    w = wide[wide[:, 0] == r]
    print(r, "kind", v >> 28, "grp", v & 0xfffffff, "cand2", int(c2[r]), "cand3", int(c3[r]), "wide", w.tolist(), "exact", int(a.view(-1)[r]))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), bad[:10].tolist())
print(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout)
