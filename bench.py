#!/usr/bin/env python
"""bench.py -- throughput of the quantizer hot path (quantized vectors / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Headline workload (BASELINE.json configs[2], the config the "1/2/4/8 B200" metric is quoted on):
  ResidualVQ, S=4 stages, K=1024 codes, D=64, training mode (forward + EMA update + backward) on
  synthetic latents z ~ 0.5*N(0,1): 1 000 000 windows x 10 frames = 10 M vectors IN TOTAL, sharded on dim 0 over
  the N GPUs (strong scaling, the reference's own split of a fixed batch, scripts/train_ablation.py:189,319-328);
  the per-stage EMA statistics are exchanged over NVLink.  `--scaling weak` keeps 1 M windows PER GPU instead (a short
  weak-scaling measurement is also reported under other_workloads at N > 1).
One "step" = one full pass (fwd + EMA + bwd) over the batch.  One JSON line is printed by rank 0.
`verify` (N > 1, outside the timed region): codebook state bit-identical on all ranks + a sharded 2-step run compared
with the single-process full-batch run of the same small problem.

`value`     : whole-job vectors/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`       : same metric through the public nn.Module call with the step's input copied from pinned HOST
              memory and the result (loss, perplexity, int32 indices) read back to the host every step.
`roofline`  : dominant kernel (K1 fused distance+argmin) timed alone with CUDA events; algorithmic flops
              2*N*K*D per launch against the measured bf16 tensor peak (MEASURED_PEAKS.json).
`cpu_baseline`: the UNMODIFIED reference modules (oracle/_ref/models/vqvae.py, staged by build()) on this box's host
              cores with torch CPU, on a bounded sample of the same workload (kind "reference"; falls back to the
              numpy oracle port, kind "port", when the staged copy is missing).  `--impl reference` prints that arm
              as its own line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "quantized vectors/sec (VQ fwd+EMA+bwd)"
UNIT = "vectors/s"

WORKLOADS = {
    # name: (kind, S, K, D, windows B, T)
    "cfg3_rvq4_k1024_d64": dict(kind="rvq", S=4, K=1024, D=64, B=1_000_000, T=10),
    "cfg1_ema_k1024_d64": dict(kind="vq", S=1, K=1024, D=64, B=4096, T=10),
}
HEADLINE = "cfg3_rvq4_k1024_d64"


def workload_text(name, cfg, B_total, world):
    K, D, S, T = cfg["K"], cfg["D"], cfg["S"], cfg["T"]
    if cfg["kind"] == "rvq":
        return (f"{name}: ResidualVQ S={S} K={K} D={D} EMA training step (fwd+EMA+bwd), z [{B_total},{D},{T}] fp32 = "
                f"{B_total * T} vectors in total, sharded on dim 0 over {world} GPU(s), per-stage EMA stats summed across ranks")
    return f"{name}: VectorQuantizer K={K} D={D} EMA training step (fwd+EMA+bwd), z [{B_total},{D},{T}] fp32 in total"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                        bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def count_between(self, t0, t1):
        return sum(1 for t, _ in self.lines if t0 <= t <= t1)

    def stop(self, t0=None, t1=None, window=None):
        """Summary of the samples received between t0 and t1 (host clock; all samples when omitted)."""
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        if window:
            out["window"] = window
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference modules (torch CPU), or the numpy oracle port when they are not staged
# ---------------------------------------------------------------------------------------------
def cpu_reference_arm(cfg, steps, warmup, chunk_vectors=65536):
    """Times forward + EMA update + backward of the reference quantizer on all host cores on a bounded sample of the
    workload.  The reference materialises N x K fp32 matrices, so the sample is a chunk that fits (BASELINE.md §3:
    65 536 vectors; halved until one step takes < 8 s)."""
    try:
        from oracle.stage_ref import load_reference_vqvae
        ref = load_reference_vqvae()
    except Exception:
        return cpu_port_arm(cfg, steps, warmup)
    import torch
    K, D, S, T = cfg["K"], cfg["D"], cfg["S"], cfg["T"]
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)           # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every host core
    torch.manual_seed(1237)
    if cfg["kind"] == "rvq":
        mod = ref.ResidualVQ(S, K, D, use_ema=True)
        layers = list(mod.layers)
    else:
        mod = ref.VectorQuantizer(K, D, use_ema=True)
        layers = [mod]
    with torch.no_grad():
        for l in layers:
            l.embedding.weight.normal_(0, 0.25)
            l.ema_w.copy_(l.embedding.weight)
            l.ema_cluster_size.fill_(1.0)
    mod.train()
    chunk = chunk_vectors
    while True:
        Bs = max(1, min(cfg["B"], chunk // T))
        z = (0.5 * torch.randn(Bs, D, T)).requires_grad_(True)
        g = torch.randn(Bs, D, T)
        one = torch.ones(())

        def step():
            z.grad = None
            loss, q, _ = mod(z)
            torch.autograd.backward([q, loss], [g, one])

        t0 = time.perf_counter()
        step()
        probe = time.perf_counter() - t0
        if probe < 8.0 or chunk <= 8192:
            break
        chunk //= 2
    for _ in range(max(0, min(warmup, 2) - 1)):
        step()
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    n = Bs * T
    return dict(value=n / med, unit=UNIT, cores=cores, kind="reference",
                sample=f"{Bs} windows x T={T} = {n} vectors per step (chunk of the workload; the reference materialises "
                       f"N x K), median of {len(times)} steps, unmodified reference modules (models/vqvae.py) on torch "
                       f"{torch.__version__} CPU, {cores} threads",
                ms_per_step=med * 1e3)


def cpu_port_arm(cfg, steps, warmup, sample_windows=None):
    """Fallback when oracle/_ref is not staged: the numpy oracle port (numpy/OpenBLAS on all host cores)."""
    import numpy as np
    from oracle import VQState, rvq_forward, rvq_backward
    K, D, S, T = cfg["K"], cfg["D"], cfg["S"], cfg["T"]
    Bs = sample_windows or max(1, min(cfg["B"], 16384 // T))
    rng = np.random.default_rng(1237)
    stages = []
    for s in range(S):
        E = (0.25 * rng.standard_normal((K, D))).astype(np.float32)
        stages.append(VQState(E.copy(), np.ones(K, np.float32), E.copy(), 0.25, True, 0.99))
    z = (0.5 * rng.standard_normal((Bs, D, T))).astype(np.float32)
    g = rng.standard_normal((Bs, D, T)).astype(np.float32)

    def step():
        fw = rvq_forward(z, stages, True)
        rvq_backward(fw, stages, g, 1.0)

    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:  # pragma: no cover
        limiter = None
    for _ in range(max(1, min(warmup, 2))):
        step()
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    if limiter is not None:
        limiter.restore_original_limits()
    times.sort()
    med = times[len(times) // 2]
    n = Bs * T
    return dict(value=n / med, unit=UNIT, cores=os.cpu_count(), kind="port",
                sample=f"{Bs} windows x T={T} = {n} vectors per step (chunk of the workload; the reference "
                       f"materialises N x K), median of {steps} steps, numpy/OpenBLAS fp32 (oracle/_ref not staged)",
                ms_per_step=med * 1e3)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def build_module(vqb200, torch, cfg, dev, seed=42):
    K, D, S = cfg["K"], cfg["D"], cfg["S"]
    torch.manual_seed(seed)
    if cfg["kind"] == "rvq":
        mod = vqb200.ResidualVQ(S, K, D, use_ema=True)
        layers = list(mod.layers)
    else:
        mod = vqb200.VectorQuantizer(K, D, use_ema=True)
        layers = [mod]
    with torch.no_grad():
        for l in layers:                      # non-degenerate start, identical on every rank
            l.embedding.weight.normal_(0, 0.25)
            l.ema_w.copy_(l.embedding.weight)
            l.ema_cluster_size.fill_(1.0)
    return mod.to(dev).train(), layers


def bind_to_gpu_numa(torch, local):
    """Pin this process (and therefore its pinned-memory allocations, first touch) to the NUMA node of its GPU:
    with every rank on node 0 the end-to-end arm is bound by one socket's memory -> PCIe path.  Best effort."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devid = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = ids & set(os.sched_getaffinity(0)) or ids
        os.sched_setaffinity(0, allowed)
        return f"node {node} ({len(allowed)} cpus)"
    except Exception as ex:  # pragma: no cover
        return f"not bound ({type(ex).__name__})"


def state_checksum(torch, layers):
    """Bit-level checksum (sum of the int32 views) of every stage's embedding.weight / ema_w / ema_cluster_size."""
    vals = []
    for l in layers:
        for t in (l.embedding.weight.detach(), l.ema_w, l.ema_cluster_size):
            vals.append(t.contiguous().view(torch.int32).to(torch.int64).sum())
    return torch.stack(vals)


def verify_multi_gpu(torch, vqb200, dist, dev, rank, world, layers, exchange):
    """Correctness bits for the multi-GPU line (outside the timed region):
    (1) the codebook state the timed loop left behind is bit-identical on every rank;
    (2) a small problem run sharded for 2 steps equals rank 0's single-process full-batch run of the same problem
        (indices identical up to near-ties counted by value, codebooks to 1e-5 relative -- fp32 summation order)."""
    out = {}
    cs = state_checksum(torch, layers)
    gathered = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(gathered, cs)
    out["ranks_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
    # (2) small sharded-vs-single comparison: RVQ 4 x 1024, 8192 windows in total
    cfg = dict(kind="rvq", S=4, K=1024, D=64, B=8192, T=10)
    gen = torch.Generator(device=dev).manual_seed(4321)
    zfull = torch.randn((cfg["B"], cfg["D"], cfg["T"]), generator=gen, device=dev).mul_(0.5)
    lo, hi = vqb200.dist.shard_bounds(cfg["B"], rank, world)
    mod_s, lay_s = build_module(vqb200, torch, cfg, dev, seed=7)
    with torch.no_grad():
        for _ in range(2):
            mod_s(zfull[lo:hi].contiguous())
    idx_s = mod_s.last_indices.clone()                               # [S, B/world, T] of step 2
    w_sharded = torch.cat([l.embedding.weight.detach().reshape(-1) for l in lay_s])
    torch.cuda.synchronize()
    vqb200.dist.disable()
    ok = torch.ones(2, device=dev)
    if rank == 0:
        mod_f, lay_f = build_module(vqb200, torch, cfg, dev, seed=7)
        with torch.no_grad():
            for _ in range(2):
                mod_f(zfull)
        w_full = torch.cat([l.embedding.weight.detach().reshape(-1) for l in lay_f])
        rel = float(((w_sharded - w_full).abs().max() / w_full.abs().max()).item())
        flips = int((mod_f.last_indices[:, lo:hi] != idx_s).sum().item())
        out["sharded_vs_single"] = {"windows": cfg["B"], "steps": 2, "codebook_max_rel_err": rel,
                                    "index_flips_rank0_shard": flips, "rows_rank0_shard": int(idx_s.numel()),
                                    "ok": bool(rel < 1e-4 and flips <= idx_s.numel() // 1000)}
        ok[0] = 1.0 if out["sharded_vs_single"]["ok"] else 0.0
    vqb200.dist.enable(peer=exchange, uniform_shards=True)
    return out


def run_gpu(args):
    import torch
    import vqb200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(torch, local) if world > 1 else "single process: not bound"
    cfg = dict(WORKLOADS[args.workload])
    if args.windows:
        cfg["B"] = args.windows
    strong = args.scaling == "strong"
    B_total = cfg["B"] if strong else cfg["B"] * world
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        # equal shards let the launch-bound shapes take the single-launch kernels with the exchange inside
        vqb200.dist.enable(peer=args.exchange, uniform_shards=(B_total % world == 0))
    lo, hi = vqb200.dist.shard_bounds(B_total, rank, world) if world > 1 else (0, B_total)
    K, D, S, T = cfg["K"], cfg["D"], cfg["S"], cfg["T"]
    B = hi - lo                              # this rank's windows
    N = B * T
    N_total = B_total * T
    peaks = load_peaks()
    mod, layers = build_module(vqb200, torch, cfg, dev)
    gen = torch.Generator(device=dev).manual_seed(1237 + rank)
    z = torch.randn((B, D, T), generator=gen, device=dev).mul_(0.5)
    g = torch.randn((B, D, T), generator=gen, device=dev)
    one = torch.ones((), device=dev)

    def step(zin):
        zin.grad = None
        loss, q, met = mod(zin)
        torch.autograd.backward([q, loss], [g, one])
        return loss, met

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # before the warm-up: nvidia-smi needs ~0.1 s before its first line
    zr = z.requires_grad_(True)
    for _ in range(args.warmup):
        step(zr)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t_host0 = time.time()
    l0 = vqb200._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, met = step(zr)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = vqb200._lib.launch_count() - l0
    t_host1 = time.time()
    # clocks DURING the timed region; when that region is shorter than a few sampling periods (multi-GPU shards: 5 steps
    # of 3.4 ms), every rank keeps running the same step until rank 0 has seen samples under the same load
    need_more = 1 if (rank == 0 and sampler.count_between(t_host0, t_host1 + 0.02) < 2) else 0
    if world > 1:
        import torch.distributed as dist
        nm = torch.tensor([need_more], device=dev)
        dist.broadcast(nm, src=0)
        need_more = int(nm.item())
    clock_window = None
    if need_more:
        t_more0 = time.time()
        n_more = max(1, int(0.5 / max(ms / args.steps * 1e-3, 1e-4)))
        for _ in range(n_more):
            step(zr)
        barrier()
        t_host0, t_host1 = t_more0, time.time()
        clock_window = f"{n_more} more steps of the same workload right after the timed region (it was shorter than two sampling periods)"
    clocks = sampler.stop(t_host0, t_host1 + 0.02, clock_window) if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = N_total / (ms_per_step * 1e-3)
    verify = None
    if world > 1 and not args.no_verify:
        import torch.distributed as dist
        verify = verify_multi_gpu(torch, vqb200, dist, dev, rank, world, layers, args.exchange)

    # ---- end-to-end: pinned host input -> H2D -> fwd+EMA+bwd -> D2H of loss / perplexity / indices ----
    # Every step copies ITS input from pinned host memory and reads ITS result back; the copy of step i+1 is issued on
    # a second stream while step i computes (two device buffers), which is how a data loader would feed the module.
    e2e = None
    try:
        z_host = torch.empty((B, D, T), dtype=torch.float32, pin_memory=True)
        z_host.copy_(z.detach())
        idx_host = torch.empty((S, B, T), dtype=torch.int32, pin_memory=True)
        sc_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
        z_bufs = [torch.empty_like(z.detach()).requires_grad_(True) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream(dev)

        def issue_copy(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # the step that used this buffer has finished
                with torch.no_grad():
                    z_bufs[b].copy_(z_host, non_blocking=True)
                copied[b].record(copy_stream)

        def e2e_step(i, last):
            b = i & 1
            if not last:
                issue_copy(i + 1)
            main.wait_event(copied[b])
            loss, met = step(z_bufs[b])
            consumed[b].record(main)
            idx = mod.last_indices if cfg["kind"] == "rvq" else mod.last_indices.unsqueeze(0)
            idx_host.copy_(idx, non_blocking=True)
            sc_host.copy_(torch.stack([loss.detach(), met["perplexity"]]), non_blocking=True)

        for ev in consumed:
            ev.record(main)
        n_e2e = max(3, min(args.steps, 5))
        issue_copy(0)
        for i in range(2):
            e2e_step(i, False)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(2, 2 + n_e2e):
            e2e_step(i, i == 1 + n_e2e)
        a1.record()
        barrier()
        ems = a0.elapsed_time(a1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": N_total / (ems / n_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(z_host.numel() * 4), "d2h_bytes_per_step": int(idx_host.numel() * 4 + 8),
               "ms_per_step": ems / n_e2e, "steps": n_e2e,
               "h2d_GBps_per_rank": z_host.numel() * 4 / (ems / n_e2e * 1e-3) / 1e9, "numa": numa,
               "note": "per-rank bytes; H2D of step i+1 overlaps compute of step i (double-buffered); PCIe-bound"}
        del z_host, idx_host, z_bufs
    except Exception as ex:  # pragma: no cover
        e2e = {"error": repr(ex)}

    # ---- roofline of the dominant kernel (K1), timed alone ----
    roof = None
    if rank == 0:
        st = layers[0]._state(dev)
        w0 = layers[0].embedding.weight
        zz = z.detach()
        for _ in range(3):
            vqb200.vq_assign(zz, w0, st)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        k0.record()
        for _ in range(reps):
            vqb200.vq_assign(zz, w0, st)
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        flops = 2.0 * N * K * D
        bytes_alg = N * (4 * D + 4)
        t_tensor = flops / (peaks["bf16_tflops"] * 1e12)
        t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        if t_tensor >= t_hbm:
            roof = {"kernel": "vq_assign (K1 fused distance+argmin)", "bound": "tensor",
                    "achieved": flops / (kms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": (flops / (kms * 1e-3) / 1e12) / peaks["bf16_tflops"], "traffic": traffic,
                    "ms_per_launch": kms, "peak_source": peaks["source"] + ", burst bf16"}
        else:
            roof = {"kernel": "vq_assign (K1 fused distance+argmin)", "bound": "hbm",
                    "achieved": bytes_alg / (kms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (bytes_alg / (kms * 1e-3) / 1e9) / peaks["hbm_gbs"], "traffic": traffic,
                    "ms_per_launch": kms, "peak_source": peaks["source"]}
        # whole-step floor for context: max(bytes/BW, flops/P)
        step_flops = 2.0 * N * S * K * D
        step_bytes = N * (20 * D + 4 * S + 4)
        floor = max(step_flops / (peaks["bf16_tflops_sustained"] * 1e12), step_bytes / (peaks["hbm_gbs"] * 1e9))
        roof["step_floor_ms"] = floor * 1e3
        roof["step_frac_of_floor"] = floor * 1e3 / ms_per_step

    # ---- N > 1, strong scaling: a short weak-scaling measurement (1 M windows PER GPU) for other_workloads ----
    weak = None
    if world > 1 and strong and not args.no_extras:
        import torch.distributed as dist
        try:
            del z, g, zr
            torch.cuda.empty_cache()
            Bw = cfg["B"]
            zw = torch.randn((Bw, D, T), generator=gen, device=dev).mul_(0.5).requires_grad_(True)
            gw = torch.randn((Bw, D, T), generator=gen, device=dev)

            def wstep():
                zw.grad = None
                lw, qw, _ = mod(zw)
                torch.autograd.backward([qw, lw], [gw, one])
            for _ in range(3):
                wstep()
            barrier()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            for _ in range(3):
                wstep()
            w1.record()
            barrier()
            t = torch.tensor([w0.elapsed_time(w1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wms = float(t.item()) / 3
            weak = {"scaling": "weak", "windows_per_gpu": Bw, "ms_per_step": wms,
                    "vectors_per_s": Bw * T * world / (wms * 1e-3), "steps": 3}
            del zw, gw
        except Exception as ex:  # pragma: no cover
            weak = {"error": repr(ex)}

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    cpu = None
    extras = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_reference_arm(cfg, steps=3, warmup=1)
    if world == 1 and not args.no_extras:
        extras = run_extras(torch, vqb200, dev, peaks)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.workload, cfg, B_total, world),
                   "vectors_total": N_total, "vectors_per_gpu": N, "parallelism": f"dp{world}",
                   "stats_exchange": vqb200.dist.peer_status() if world > 1 else "none (single GPU)",
                   "l2": "inputs larger than L2 (2.56 GB per tensor per GPU)" if N * D * 4 > 256e6 else "L2 not flushed (small input)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "cpu_baseline": cpu,
        "loss": float(loss.item()), "perplexity": float(met["perplexity"].item()),
    }
    if verify is not None:
        out["verify"] = verify
    if extras:
        out["other_workloads"] = extras
    if weak is not None:
        out.setdefault("other_workloads", {})["cfg3_weak_scaling"] = weak
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_extras(torch, vqb200, dev, peaks):
    """Short measurements of the other BASELINE.json configs (parity-test cases, not the headline)."""
    res = {}

    def timeit(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    try:
        # cfg1: EMA-VQ K=1024, z [4096,64,10]
        cfg = WORKLOADS["cfg1_ema_k1024_d64"]
        mod, _ = build_module(vqb200, torch, cfg, dev)
        z = torch.randn(cfg["B"], cfg["D"], cfg["T"], device=dev).requires_grad_(True)
        g = torch.randn_like(z); one = torch.ones((), device=dev)

        def s1():
            z.grad = None
            loss, q, _ = mod(z)
            torch.autograd.backward([q, loss], [g, one])
        l0 = vqb200._lib.launch_count()
        ms = timeit(s1, 20)
        res["cfg1_ema_k1024_n40960"] = {"us_per_step": ms * 1e3, "vectors_per_s": cfg["B"] * cfg["T"] / (ms * 1e-3),
                                        "launches_per_step": (vqb200._lib.launch_count() - l0) / 23}
        # cfg2: Hybrid on the permuted [512,64,1] view
        torch.manual_seed(42)
        hy = vqb200.HybridVQ(64, [8, 5, 5, 5], vq_codebook_size=512).to(dev).train()
        zz = torch.randn(512, 1, 64, device=dev).permute(0, 2, 1).requires_grad_(True)
        g2 = torch.randn(512, 64, 1, device=dev)

        def s2():
            zz.grad = None
            loss, q, _ = hy(zz)
            torch.autograd.backward([q, loss], [g2, one])
        l0 = vqb200._lib.launch_count()
        ms = timeit(s2, 30)
        res["cfg2_hybrid_n512"] = {"us_per_step": ms * 1e3, "vectors_per_s": 512 / (ms * 1e-3),
                                   "vqb200_launches_per_step": (vqb200._lib.launch_count() - l0) / 33}
        gs = vqb200.GraphedQuantizerStep(hy, zz.detach())
        ms = timeit(lambda: gs(zz.detach(), g2), 50)
        res["cfg2_hybrid_n512_cuda_graph"] = {"us_per_step": ms * 1e3, "vectors_per_s": 512 / (ms * 1e-3)}
        g1s = vqb200.GraphedQuantizerStep(mod, z.detach())
        ms = timeit(lambda: g1s(z.detach(), g), 50)
        res["cfg1_ema_k1024_n40960_cuda_graph"] = {"us_per_step": ms * 1e3, "vectors_per_s": cfg["B"] * cfg["T"] / (ms * 1e-3)}
        # cfg4 (BASELINE.json configs[3]): FSQ / LFQ batch sweep 4096 -> 1 M windows x 10 frames.
        # "elementwise" = the round / sign + index-pack stage alone on the post-projection tensor (40 / 88 algorithmic
        # bytes per vector); "module" = the whole module with both 1x1 projections fused in (8D + 4d + 8 bytes forward,
        # 20D + 8d + 8 forward + backward).  Small batches are launch-bound: read them as microseconds.
        basis = torch.tensor([1, 8, 40, 200], dtype=torch.int32, device=dev)
        sweep = {}
        for B4 in (4096, 16384, 65536, 262144, 1048576):
            n = B4 * 10
            reps = 20 if B4 <= 65536 else 8
            row = {"vectors": n}
            ze = 2.0 * torch.randn(B4, 4, 10, device=dev)
            ms = timeit(lambda: vqb200.fsq_round(ze, basis, 1000), reps)
            row["fsq_elementwise"] = {"us": ms * 1e3, "vectors_per_s": n / (ms * 1e-3), "frac_of_hbm": n * 40 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
            zl = torch.randn(B4, 10, 10, device=dev)
            ms = timeit(lambda: vqb200.lfq_sign(zl, 0.1), reps)
            row["lfq_elementwise"] = {"us": ms * 1e3, "vectors_per_s": n / (ms * 1e-3), "frac_of_hbm": n * 88 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
            del ze, zl
            z4 = 2.0 * torch.randn(B4, 64, 10, device=dev)
            g4 = torch.randn(B4, 64, 10, device=dev)
            for name, m4, dq in (("fsq", vqb200.FSQ([8, 5, 5, 5], 64, 64).to(dev), 4), ("lfq", vqb200.LFQ(64, 10).to(dev), 10)):
                def f4():
                    with torch.no_grad():
                        m4(z4)
                ms = timeit(f4, reps)
                byts = n * (8 * 64 + 4 * dq + 8)
                row[f"{name}_module_fwd"] = {"us": ms * 1e3, "vectors_per_s": n / (ms * 1e-3),
                                             "frac_of_hbm": byts / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
                zz4 = z4.clone().requires_grad_(True)

                def fb4():
                    zz4.grad = None
                    loss, out, _ = m4(zz4)
                    if loss.requires_grad:
                        torch.autograd.backward([out, loss], [g4, one])
                    else:
                        out.backward(g4)
                ms = timeit(fb4, max(3, reps // 2))
                byts = n * (20 * 64 + 8 * dq + 8)
                row[f"{name}_module_fwdbwd"] = {"us": ms * 1e3, "vectors_per_s": n / (ms * 1e-3),
                                                "frac_of_hbm": byts / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
                del zz4
            del z4, g4
            sweep[f"B{B4}"] = row
        res["cfg4_fsq_lfq_batch_sweep"] = sweep
        # cfg5 (BASELINE.json configs[4]): EMA-VQ forward + EMA update (train mode, no backward) on N = 4 194 304 vectors
        # [N, D, 1], codebook N(0,1), the full K x D grid.  floor = max(flops / bf16 peak, (8D + 4) bytes / HBM) per
        # SURVEY.md §8(d); frac = floor / measured.  (One warm-up + 2 timed steps per point.)
        grid = {}
        N5 = 4_194_304
        for D5 in (64, 128, 256, 512):
            z5 = torch.randn(N5, D5, 1, device=dev)
            for K5 in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
                try:
                    torch.manual_seed(5)
                    m5 = vqb200.VectorQuantizer(K5, D5, use_ema=True).to(dev).train()
                    with torch.no_grad():
                        m5.embedding.weight.normal_(0, 1.0)
                        m5.ema_w.copy_(m5.embedding.weight)
                        m5.ema_cluster_size.fill_(1.0)

                    def s5():
                        with torch.no_grad():
                            m5(z5)
                    ms = timeit(s5, 2, warm=1)
                    fl = 2.0 * N5 * K5 * D5
                    floor_ms = max(fl / (peaks["bf16_tflops"] * 1e12), N5 * (8 * D5 + 4) / (peaks["hbm_gbs"] * 1e9)) * 1e3
                    grid[f"k{K5}_d{D5}"] = {"ms": ms, "vectors_per_s": N5 / (ms * 1e-3), "TFLOPs": fl / (ms * 1e-3) / 1e12,
                                            "bound": "tensor" if fl / (peaks["bf16_tflops"] * 1e12) * 1e3 >= floor_ms * 0.999 else "hbm",
                                            "frac_of_roofline": floor_ms / ms}
                    del m5
                except Exception as ex:  # pragma: no cover
                    grid[f"k{K5}_d{D5}"] = {"error": repr(ex)[:200]}
            del z5
            torch.cuda.empty_cache()
        res["cfg5_ema_vq_fwd_ema_n4m_grid"] = grid
    except Exception as ex:  # pragma: no cover
        res["error"] = repr(ex)
    return res


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(WORKLOADS[args.workload])
    K, D, S, B, T = cfg["K"], cfg["D"], cfg["S"], cfg["B"], cfg["T"]
    r = cpu_reference_arm(cfg, steps=max(1, args.steps), warmup=args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.workload, cfg, B if args.scaling == "strong" else B * world, world),
                   "parallelism": "host cpu", "sample": r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vqb200", choices=["vqb200", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--windows", type=int, default=0, help="override the number of windows B (development only)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="transport of the per-stage EMA statistics at N > 1: peer memory fused into the finalize "
                         "kernels (csrc/peer.cu) or an NCCL all-reduce; auto = peer when the node allows it")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the workload's windows are the TOTAL, sharded over the GPUs (BASELINE.json "
                         "configs[2]); weak: that many windows per GPU")
    ap.add_argument("--no-verify", action="store_true", help="skip the multi-GPU correctness bits (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vqb200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
