"""Development: run the cfg5 grid points one by one (sync after each) to find a failing configuration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
N5 = int(sys.argv[1]) if len(sys.argv) > 1 else 4_194_304
for D5 in (64, 128, 256, 512):
    z5 = torch.randn(N5, D5, 1, device=dev)
    for K5 in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        torch.manual_seed(5)
        m5 = vqb200.VectorQuantizer(K5, D5, use_ema=True).to(dev).train()
        with torch.no_grad():
            m5.embedding.weight.normal_(0, 1.0); m5.ema_w.copy_(m5.embedding.weight); m5.ema_cluster_size.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m5(z5); torch.cuda.synchronize()
            a.record(); m5(z5); b.record(); torch.cuda.synchronize()
        print(f"D={D5} K={K5} ok {a.elapsed_time(b):.2f} ms", flush=True)
        del m5
    del z5
