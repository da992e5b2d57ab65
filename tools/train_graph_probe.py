"""Where does a training step of the drop-in model spend its time?  CUDA-event timing of eager forward / backward /
optimizer and of a whole-step graph replay (transformer + hybrid, batch 512, window 10)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import trainer
from models.vqvae import DualMotionVQVAE

dev = torch.device("cuda:0")
torch.manual_seed(0)
res = {}
for arch, method, bs in (("transformer", "hybrid", 512), ("resnet_no_down", "ema", 4096)):
    m = DualMotionVQVAE(robot_input_dim=29, human_input_dim=126, hidden_dim=64, arch=arch, method=method, window_size=10).to(dev).train()
    x = torch.randn(bs, 10, 29, device=dev)
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-4, capturable=True, fused=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    def step():
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        loss = trainer.teacher_loss(m(x_robot=x)["robot"], x)
        ev[1].record()
        loss.backward()
        ev[2].record()
        opt.step()
        ev[3].record()
        return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    import time
    t0 = time.time()
    for _ in range(20): step()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / 20 * 1e3
    r = {"eager_wall_ms": wall, "fwd_ms": ev[0].elapsed_time(ev[1]), "bwd_ms": ev[1].elapsed_time(ev[2]), "opt_ms": ev[2].elapsed_time(ev[3])}
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            opt.zero_grad(set_to_none=True); trainer.teacher_loss(m(x_robot=x)["robot"], x).backward(); opt.step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    for mm in m.modules():
        if hasattr(mm, "invalidate_cache"): mm.invalidate_cache()
    opt.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        l = trainer.teacher_loss(m(x_robot=x)["robot"], x); l.backward(); opt.step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): g.replay()
    b.record(); torch.cuda.synchronize()
    r["graph_replay_ms"] = a.elapsed_time(b) / 20
    res[f"{arch}/{method}/B{bs}"] = r
print(json.dumps(res))
