"""ms per training step of the drop-in DualMotionVQVAE (teacher mode) through <pkg>/trainer.py: eager loop vs the
whole-step CUDA graph (--cuda_graph).  Median over epochs 2.. of the wall time of the training loop / steps (one host
sync per epoch)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import trainer

dev = torch.device("cuda:0")
out = {}
for arch, method, window, bs in (("transformer", "hybrid", 10, 512), ("resnet_no_down", "ema", 10, 4096)):
    n = bs * 10 * 10 // 9 + 16              # ~10 full batches per epoch after the 90 % split
    for tag, extra in (("eager", []), ("graph", ["--cuda_graph"])):
        with tempfile.TemporaryDirectory() as tmp:
            args = trainer.build_parser().parse_args(
                ["--mode", "teacher", "--arch", arch, "--method", method, "--window", str(window), "--epochs", "8",
                 "--batch_size", str(bs), "--synthetic", str(n), "--data_root", os.path.join(tmp, "x"),
                 "--ckpt_dir", os.path.join(tmp, "ck"), "--log_dir", os.path.join(tmp, "res")] + extra)
            trainer.train_one_seed(args, 1, dev)
        ms = sorted(trainer.train_one_seed.step_ms[2:])
        out[f"{arch}/{method}/B{bs}/{tag}"] = {"ms_per_step_median": ms[len(ms) // 2], "epochs": len(ms)}
print(json.dumps(out))
