import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import _lib
dev = torch.device("cuda:0")
N5, D5, K5 = int(sys.argv[1]), 64, int(sys.argv[2])
z5 = torch.randn(N5, D5, 1, device=dev)
torch.manual_seed(5)
m5 = vqb200.VectorQuantizer(K5, D5, use_ema=True).to(dev).train()
with torch.no_grad():
    m5.embedding.weight.normal_(0, 1.0); m5.ema_w.copy_(m5.embedding.weight); m5.ema_cluster_size.fill_(1.0)
    for i in range(3):
        l0 = _lib.launch_count()
        m5(z5); torch.cuda.synchronize()
        st = m5._state(dev)
        print("step", i, "ok; launches", _lib.launch_count() - l0, "ws", st._assign_ws.view(torch.int32)[:8].tolist(),
              "info", st.info.tolist(), "E absmax", float(m5.embedding.weight.abs().max()), flush=True)
