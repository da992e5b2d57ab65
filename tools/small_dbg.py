import sys, os
os.environ["VQB200_SMALL_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
S, K, N = (int(x) for x in sys.argv[1:4])
m = vqb200.ResidualVQ(S, K, 64, use_ema=True).to(dev).train()
with torch.no_grad():
    for l in m.layers: l.embedding.weight.normal_(0, 0.3); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1)
z = torch.randn(N, 64, 1, device=dev)
with torch.no_grad():
    for _ in range(4): m(z)
torch.cuda.synchronize()
ws = m.layers[0]._state(dev)._small_ws
sc = sum(K * 65 for _ in range(S)); sc = (sc + 3) & ~3
scr = sum(K + 8 for _ in range(S)); off = sc + ((scr + 3) & ~3) + 4
t = ws[off:off + 120].view(torch.int64).cpu().tolist()
prev = t[0]
for i, v in enumerate(t[1:40]):
    if v == 0: break
    print(i + 1, v - prev)
    prev = v
