"""Development: why does a row's filter verdict miss the exact arg min?  Recomputes the fp16 filter scores in torch."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
B, T, K, row = 1000000, 10, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 6778768
torch.manual_seed(0)
W = torch.randn(K, 64, device=dev) * 0.3
st = vqb200.QuantizerState(K, 64, dev)
z = 0.5 * torch.randn(B, 64, T, device=dev)
i_simt = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
i_tc = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
b, t = divmod(row, T)
x = z[b, :, t].double()
Wd = W.double()
S = Wd @ x - 0.5 * (Wd * Wd).sum(1)
# fp16 emulation with the kernel's scaling
m = x.abs().max().item()
import math
sx = 2.0 ** (10 - math.floor(math.log2(m)))
xh = (x.float() * sx).half().double() / sx
out = {"row": row, "simt": int(i_simt.view(-1)[row]), "tc": int(i_tc.view(-1)[row])}
Sa = torch.empty(K, dtype=torch.float64, device=dev)
for j in range(K // 128):
    Wt = W[j * 128:(j + 1) * 128]
    mm = Wt.abs().max().item()
    se = 2.0 ** (10 - math.floor(math.log2(mm)))
    Wh = (Wt * se).half().double() / se
    Sa[j * 128:(j + 1) * 128] = Wh @ xh - 0.5 * (Wt.double() ** 2).sum(1)
top = torch.topk(S, 6)
out["exact_top"] = [(int(i), float(v)) for v, i in zip(top.values, top.indices)]
topa = torch.topk(Sa, 6)
out["approx_top"] = [(int(i), float(v)) for v, i in zip(topa.values, topa.indices)]
out["max_abs_err"] = float((Sa - S).abs().max())
xn = float(x.norm()); emax = float(W.norm(dim=1).max()); nmin = float(W.norm(dim=1).min())
R = min(emax, 3 * xn + 2 * nmin); mag = xn * R
out["xn"], out["emax"], out["nmin"], out["thr"] = xn, emax, nmin, 2 * (1.0e-3 * mag + 4.2e-6 * (mag + 0.5 * R * R)) + 2.4e-7 * (xn * xn + R * R)
g = Sa.view(-1, 4).max(1).values
tg = torch.topk(g, 5)
out["group_top"] = [(int(i), float(v)) for v, i in zip(tg.values, tg.indices)]
print(json.dumps(out, indent=1))
