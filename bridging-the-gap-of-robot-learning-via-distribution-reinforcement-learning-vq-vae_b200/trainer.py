"""Data-parallel trainer for the drop-in `DualMotionVQVAE`.

The reference trains with `nn.DataParallel` (`scripts/train_ablation.py:189`): one process scatters every batch,
replicates the module per GPU and -- for the EMA quantizers -- keeps only replica 0's codebook statistics.  This
launcher keeps that script's command line, loss recipe, checkpoint files and logs, but runs ONE PROCESS PER GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_ddp.py \\
        --mode teacher --arch transformer --method hybrid --window 10 --epochs 400 --batch_size 512

* every rank holds the whole (small) paired dataset on its GPU and takes its `vqb200.dist.shard_bounds` slice of each
  global batch -- no DataLoader workers, no host round trip per step;
* the quantizer sums its EMA statistics over all ranks inside the finalize kernels (peer memory, csrc/peer.cu; NCCL
  when peer mapping is unavailable), i.e. the full-batch update the single-GPU run would do;
* encoder / decoder / standard-VQ gradients are averaged with bucketed all-reduces (`vqb200.dist.average_gradients`);
* rank 0 writes `checkpoints/<name>_<method>_<mode>_seed_<s>_{last,best,final}.pth` and `results/log_*.json` in the
  reference's formats (`scripts/train_ablation.py:276-283,338-364`), so `export_motion.py` / `analyze_latent_space.py`
  load them unchanged.

Like the reference (`:319-324`) the global batch is `--batch_size x number of GPUs`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import dist as vq_dist

# loss weights of the reference recipe (scripts/train_ablation.py:50-55)
W_RECON, W_VQ, W_VEL, W_ALIGN = 1.0, 1.0, 0.5, 100.0
LR, WEIGHT_DECAY = 2e-4, 1e-4


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="vqb200 data-parallel trainer (CLI of scripts/train_ablation.py)")
    p.add_argument("--method", type=str, default="hybrid")
    p.add_argument("--arch", type=str, default="transformer")
    p.add_argument("--epochs", type=int, default=400)
    p.add_argument("--batch_size", type=int, default=256, help="per-GPU batch (the reference scales it by the GPU count)")
    p.add_argument("--seed", type=int, nargs="+", default=[42])
    p.add_argument("--window", type=int, default=64)
    p.add_argument("--patience", type=int, default=-1)
    p.add_argument("--mode", type=str, default="teacher", choices=["teacher", "student"])
    p.add_argument("--resume", action="store_true")
    p.add_argument("--teacher_ckpt", type=str, default=None)
    p.add_argument("--name", type=str, default=None, help="run name prefix (default: <arch>_W<window>)")
    p.add_argument("--data_root", type=str, default=os.path.join("data", "processed"))
    p.add_argument("--synthetic", type=int, default=0,
                   help="if > 0 and the .npy files are missing: train on N synthetic windows (rng seed 0)")
    p.add_argument("--exchange", type=str, default="auto", choices=["auto", "peer", "nccl"])
    p.add_argument("--cuda_graph", action="store_true",
                   help="single GPU: capture forward + backward + AdamW of a full batch in one CUDA graph and replay it")
    p.add_argument("--log_dir", type=str, default="results")
    p.add_argument("--ckpt_dir", type=str, default="checkpoints")
    return p


def load_paired(root: str, window: int, synthetic: int):
    """g1_train.npy / human_train.npy as [N, window, dim] float32 (scripts/train_ablation.py:84-99)."""
    r_path, h_path = os.path.join(root, "g1_train.npy"), os.path.join(root, "human_train.npy")
    if os.path.exists(r_path) and os.path.exists(h_path):
        r, h = np.load(r_path).astype(np.float32), np.load(h_path).astype(np.float32)
    elif synthetic > 0:
        rng = np.random.default_rng(0)
        r = rng.standard_normal((synthetic, window, 29)).astype(np.float32)
        h = rng.standard_normal((synthetic, window, 126)).astype(np.float32)
    else:
        raise FileNotFoundError(f"{r_path} / {h_path} missing (run the reference's process_data.py or pass --synthetic N)")
    n = min(len(r), len(h))
    return r[:n], h[:n]


def split_indices(n: int, seed: int):
    """90 / 10 train / validation split, the same on every rank."""
    perm = np.random.default_rng(seed).permutation(n)
    cut = int(0.9 * n)
    return perm[:cut], perm[cut:]


def epoch_batches(train_idx: np.ndarray, global_batch: int, seed: int, epoch: int) -> List[np.ndarray]:
    """Shuffled global batches of one epoch -- a pure function of (seed, epoch), identical on all ranks."""
    order = np.random.default_rng([seed, epoch]).permutation(train_idx)
    out = [order[i:i + global_batch] for i in range(0, len(order), global_batch)]
    # every rank must take part in every training step (the EMA finalize kernels barrier across ranks): a trailing
    # batch with fewer samples than ranks is dropped
    world = vq_dist.world_size()
    if vq_dist.uniform_shards():
        # equal shards on every rank (the promise behind enable(uniform_shards=True)): trim each batch to a multiple of
        # the world size -- at most world-1 samples of the ragged tail are skipped per epoch
        out = [b[:len(b) // world * world] for b in out]
    return [b for b in out if len(b) >= world]


def _shard(batch: np.ndarray) -> np.ndarray:
    lo, hi = vq_dist.shard_bounds(len(batch))
    return batch[lo:hi]


def teacher_loss(out_r: Dict[str, torch.Tensor], x_r: torch.Tensor) -> torch.Tensor:
    recon = out_r["recon"]
    vel = F.mse_loss(recon[:, :, 1:] - recon[:, :, :-1], x_r[:, :, 1:] - x_r[:, :, :-1])
    return W_RECON * F.mse_loss(recon, x_r) + W_VQ * out_r["loss_vq"] + W_VEL * vel


def _all_sum(values: List[float], device) -> List[float]:
    if not vq_dist.enabled():
        return values
    import torch.distributed as td
    t = torch.tensor(values, dtype=torch.float64, device=device)
    td.all_reduce(t)
    return [float(v) for v in t.tolist()]


def train_one_seed(args, seed: int, device: torch.device, model_factory=None) -> Dict[str, list]:
    rank, world = vq_dist.rank(), vq_dist.world_size()
    is_main = rank == 0
    torch.manual_seed(seed)
    np.random.seed(seed)
    if model_factory is None:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        from models.vqvae import DualMotionVQVAE
        model_factory = DualMotionVQVAE
    r_np, h_np = load_paired(args.data_root, args.window, args.synthetic)
    x_r_all = torch.from_numpy(r_np).to(device)
    x_h_all = torch.from_numpy(h_np).to(device)
    train_idx, val_idx = split_indices(len(r_np), seed)
    global_batch = args.batch_size * world

    model = model_factory(robot_input_dim=r_np.shape[-1], human_input_dim=h_np.shape[-1], hidden_dim=64, arch=args.arch,
                          method=args.method, window_size=args.window).to(device)      # same seed => same init everywhere
    name = args.name or f"{args.arch}_W{args.window}"
    run = f"{name}_{args.method}_{args.mode}_seed_{seed}"
    last_path = os.path.join(args.ckpt_dir, f"{run}_last.pth")
    log_path = os.path.join(args.log_dir, f"log_{name}_seed_{seed}.json")
    config = {"name": name, "method": args.method, "arch": args.arch, "mode": args.mode, "window": args.window,
              "epochs": args.epochs, "batch_size": args.batch_size, "patience": args.patience, "world_size": world}
    history: Dict[str, list] = {k: [] for k in ("train_loss", "val_loss", "val_recon", "val_align")}
    start_epoch, best = 0, float("inf")

    if args.resume and os.path.exists(last_path):
        ck = torch.load(last_path, map_location=device)
        model.load_state_dict(ck["model_state_dict"])
        if ck.get("config", {}).get("mode") == args.mode:
            start_epoch, best = ck["epoch"] + 1, ck.get("best_loss", float("inf"))
        if os.path.exists(log_path):
            try:
                history = json.load(open(log_path))
            except (OSError, ValueError):
                pass
    elif args.mode == "student":
        if not (args.teacher_ckpt and os.path.exists(args.teacher_ckpt)):
            raise ValueError("student mode requires a valid --teacher_ckpt")
        teacher = torch.load(args.teacher_ckpt, map_location=device)
        teacher = teacher.get("model_state_dict", teacher)
        sd = model.state_dict()
        sd.update({k: v for k, v in teacher.items() if "human_encoder" not in k})
        model.load_state_dict(sd)
        for n_, p in model.named_parameters():
            if "human_encoder" not in n_:
                p.requires_grad = False

    params = [p for p in model.parameters() if p.requires_grad]
    use_graph = bool(getattr(args, "cuda_graph", False)) and world == 1
    # fused=True: one multi-tensor kernel for all ~160 parameters instead of several tiny kernels per parameter (same update
    # rule as the reference's torch.optim.AdamW(lr=2e-4, weight_decay=1e-4), scripts/train_ablation.py:182)
    opt = torch.optim.AdamW(params, lr=LR, weight_decay=WEIGHT_DECAY, capturable=use_graph, fused=True)

    def forward_loss(x_r, x_h):
        if args.mode == "teacher":
            return teacher_loss(model(x_robot=x_r, x_human=None)["robot"], x_r)
        out = model(x_robot=x_r, x_human=x_h)
        return W_ALIGN * F.mse_loss(out["human"]["z_e"], out["robot"]["z_e"].detach())

    graph = None
    if use_graph:
        # Whole-step CUDA graph: forward + backward + AdamW of one full batch captured once and
        # replayed -- at batch 512 the transformer teacher is launch-bound end to end.  Every vqb200 call is
        # capturable (no host sync, metrics stay on the device).  Only full batches replay; the ragged tail of an epoch is
        # dropped.  Warm-up runs on a side stream and is rolled back so that the trajectory starts from the same state.
        model.train()
        xr_s = torch.empty((args.batch_size,) + tuple(x_r_all.shape[1:]), device=device)
        xh_s = torch.empty((args.batch_size,) + tuple(x_h_all.shape[1:]), device=device)
        first = torch.from_numpy(train_idx[:args.batch_size] if len(train_idx) >= args.batch_size else
                                 np.resize(train_idx, args.batch_size)).to(device)
        xr_s.copy_(x_r_all[first]); xh_s.copy_(x_h_all[first])
        saved_model = {k: v.clone() for k, v in model.state_dict().items()}
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                opt.zero_grad(set_to_none=True)
                forward_loss(xr_s, xh_s).backward()
                opt.step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        with torch.no_grad():
            model.load_state_dict(saved_model)
            for st_ in opt.state.values():          # Adam moments and step counters back to zero
                for v in st_.values():
                    if torch.is_tensor(v):
                        v.zero_()
        for m_ in model.modules():
            if hasattr(m_, "invalidate_cache"):
                m_.invalidate_cache()
        opt.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss_s = forward_loss(xr_s, xh_s)
            loss_s.backward()
            opt.step()
        with torch.no_grad():                       # capture does not execute, but be explicit about the state
            model.load_state_dict(saved_model)
        for m_ in model.modules():
            if hasattr(m_, "invalidate_cache"):
                m_.invalidate_cache()

    stale = 0
    t0 = time.time()
    train_one_seed.step_ms = []                  # wall ms per training step of every epoch (tools/train_step_time.py)
    for epoch in range(start_epoch, args.epochs):
        model.train()
        batches = epoch_batches(train_idx, global_batch, seed, epoch)
        if graph is not None:
            batches = [b for b in batches if len(b) == args.batch_size]
        total_t = torch.zeros((), dtype=torch.float64, device=device)        # no host sync inside the epoch
        t_epoch = time.time()
        for gb in batches:
            mine = torch.from_numpy(_shard(gb)).to(device, non_blocking=True)
            if graph is not None:
                torch.index_select(x_r_all, 0, mine, out=xr_s)
                torch.index_select(x_h_all, 0, mine, out=xh_s)
                graph.replay()
                total_t += loss_s.detach()
                continue
            x_r, x_h = x_r_all[mine], x_h_all[mine]
            opt.zero_grad(set_to_none=True)
            loss = forward_loss(x_r, x_h)
            loss.backward()
            vq_dist.average_gradients(params)
            opt.step()
            total_t += loss.detach()
        total = float(total_t)                     # the epoch's only host sync
        train_one_seed.step_ms.append((time.time() - t_epoch) * 1e3 / max(len(batches), 1))
        if graph is not None:
            # replays change weights behind Python's back (no tensor version bump): drop the codebook-derived caches
            for m_ in model.modules():
                if hasattr(m_, "invalidate_cache"):
                    m_.invalidate_cache()
        model.eval()
        v_sum, v_cnt = 0.0, 0
        with torch.no_grad():
            for i in range(0, len(val_idx), global_batch):
                mine = _shard(val_idx[i:i + global_batch])
                if len(mine) == 0:
                    continue
                sel = torch.from_numpy(mine).to(device)
                out = model(x_robot=x_r_all[sel], x_human=x_h_all[sel])
                if args.mode == "teacher":
                    v = F.mse_loss(out["robot"]["recon"], x_r_all[sel], reduction="sum")
                    v_cnt += out["robot"]["recon"].numel()
                else:
                    v = F.mse_loss(out["human"]["z_e"], out["robot"]["z_e"], reduction="sum")
                    v_cnt += out["human"]["z_e"].numel()
                v_sum += float(v)
        t_sum, v_sum, v_cnt = _all_sum([total, v_sum, float(v_cnt)], device)
        train_loss = t_sum / max(len(batches) * world, 1)
        val = v_sum / max(v_cnt, 1.0)
        history["train_loss"].append(train_loss)
        history["val_recon" if args.mode == "teacher" else "val_align"].append(val)
        improved = val < best
        if improved:
            best, stale = val, 0
        else:
            stale += 1
        if is_main:
            os.makedirs(args.ckpt_dir, exist_ok=True)
            os.makedirs(args.log_dir, exist_ok=True)
            ck = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                  "best_loss": best, "config": config}
            torch.save(ck, last_path)
            json.dump(history, open(log_path, "w"), indent=4)
            if improved:
                torch.save(ck, os.path.join(args.ckpt_dir, f"{run}_best.pth"))
            if epoch % 5 == 0 or epoch == args.epochs - 1:
                print(f"[{run}] epoch {epoch}: train {train_loss:.4f} | val {val:.4f} | {time.time() - t0:.0f}s", flush=True)
        if args.patience > 0 and stale >= args.patience:
            break
    if is_main:
        os.makedirs(args.ckpt_dir, exist_ok=True)
        os.makedirs(args.log_dir, exist_ok=True)
        torch.save(model.state_dict(), os.path.join(args.ckpt_dir, f"{run}_final.pth"))
        json.dump(history, open(os.path.join(args.log_dir, f"log_{name}_{args.mode}_seed_{seed}.json"), "w"), indent=4)
    train_one_seed.last_model = model          # for tests
    return history


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("vqb200 trainer: no CUDA device (the quantizer engine has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as td
        if not td.is_initialized():
            td.init_process_group("nccl", device_id=device)
        vq_dist.enable(peer=args.exchange, uniform_shards=True)
        if vq_dist.rank() == 0:
            print(f"vqb200 trainer: {world} ranks, EMA statistics exchange = {vq_dist.peer_status()}", flush=True)
    try:
        for seed in args.seed:
            train_one_seed(args, seed, device)
    finally:
        if world > 1:
            import torch.distributed as td
            vq_dist.disable()
            td.destroy_process_group()
    return 0
