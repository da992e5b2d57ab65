import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
S, K, N = (int(x) for x in sys.argv[1:4])
m = vqb200.ResidualVQ(S, K, 64, use_ema=True).to(dev).train()
with torch.no_grad():
    for l in m.layers: l.embedding.weight.normal_(0, 0.3); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1)
z = torch.randn(N, 64, 1, device=dev)
if os.environ.get('WARM'):
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(400): a @ a
with torch.no_grad():
    for _ in range(4): m(z)
torch.cuda.synchronize(); print("ok")
