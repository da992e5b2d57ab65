import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
N5, D5, K5, mode = int(sys.argv[1]), 64, int(sys.argv[2]), sys.argv[3]
z5 = torch.randn(N5, D5, 1, device=dev)
torch.manual_seed(5)
m5 = vqb200.VectorQuantizer(K5, D5, use_ema=True).to(dev)
m5.train(mode == "train")
with torch.no_grad():
    m5.embedding.weight.normal_(0, 1.0); m5.ema_w.copy_(m5.embedding.weight); m5.ema_cluster_size.fill_(1.0)
    if mode == "assign":
        st = m5._state(dev)
        for i in range(6):
            vqb200.vq_assign(z5, m5.embedding.weight, st)
        torch.cuda.synchronize(); print("assign x6 ok", flush=True)
    else:
        for i in range(6):
            m5(z5)
        torch.cuda.synchronize(); print(mode, "x6 ok", flush=True)
