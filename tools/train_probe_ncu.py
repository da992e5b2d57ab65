import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import trainer
from models.vqvae import DualMotionVQVAE
dev = torch.device("cuda:0"); torch.manual_seed(0)
arch, method, bs = sys.argv[1], sys.argv[2], int(sys.argv[3])
m = DualMotionVQVAE(robot_input_dim=29, human_input_dim=126, hidden_dim=64, arch=arch, method=method, window_size=10).to(dev).train()
x = torch.randn(bs, 10, 29, device=dev)
opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=2e-4, weight_decay=1e-4, fused=True)
for _ in range(3):
    opt.zero_grad(set_to_none=True); trainer.teacher_loss(m(x_robot=x)["robot"], x).backward(); opt.step()
torch.cuda.synchronize(); print("ok")
