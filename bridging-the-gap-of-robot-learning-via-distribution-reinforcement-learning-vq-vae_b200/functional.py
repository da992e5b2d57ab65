"""Functional API + autograd Functions over the C ABI (include/vqb200.h).

Returns the 4-tuple `(quantized, loss, perplexity, indices)` style results BASELINE.json's north_star
names; the nn.Modules in quantizers.py adapt them to the reference's real 3-tuple
`(loss, quantized, metrics)` (models/vqvae.py:74-76).
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_float, c_int, c_int64, c_size_t
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import dist as _dist
from ._lib import check, ptr, stream_ptr


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("vqb200: input must be a CUDA tensor (the engine has no CPU fallback)")
    return t if t.dtype == torch.float32 else t.float()


class QuantizerState:
    """Non-persistent device state derived from one codebook: |E_k|^2, the bf16 tile image for the
    tcgen05 kernel, info flags, and reusable scratch.  Refreshed whenever `embedding.weight` changes
    (optimizer step, load_state_dict, EMA finalize)."""

    def __init__(self, K: int, D: int, device: torch.device):
        lib = _lib.load()
        self.K, self.D, self.device = K, D, device
        f32 = dict(dtype=torch.float32, device=device)
        self.ee = torch.empty(K, **f32)
        nbytes = int(lib.vqb200_codebook_image_bytes(K, D))
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        off = (-raw.data_ptr()) % 1024
        self._image_raw = raw
        self.image = raw[off:off + nbytes]
        self.info = torch.zeros(4, **f32)
        self.stats = torch.empty(K * (D + 1), **f32)          # [dw | cnt]
        self.scratch = torch.empty(K + 8, **f32)
        self.sse = torch.zeros(1, dtype=torch.float64, device=device)
        self._assign_ws: Optional[torch.Tensor] = None
        self._key = None

    @property
    def cnt(self) -> torch.Tensor:
        return self.stats[self.K * self.D:]

    def assign_workspace(self, N: int) -> torch.Tensor:
        need = int(_lib.load().vqb200_assign_workspace_bytes(N, self.D))
        if self._assign_ws is None or self._assign_ws.numel() < need:
            self._assign_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._assign_ws

    def _weight_key(self, weight: torch.Tensor):
        return (weight.data_ptr(), weight._version)

    def refresh(self, weight: torch.Tensor) -> None:
        key = self._weight_key(weight)
        if key != self._key:
            codebook_prepare(weight, self)
            self._key = key

    def mark_fresh(self, weight: torch.Tensor) -> None:
        self._key = self._weight_key(weight)

    def invalidate(self) -> None:
        self._key = None


def codebook_prepare(weight: torch.Tensor, state: QuantizerState) -> None:
    lib = _lib.load()
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    with torch.cuda.device(w.device):
        check(lib.vqb200_codebook_prepare(ptr(w), w.shape[0], w.shape[1], ptr(state.ee), ptr(state.image),
                                          ptr(state.info), stream_ptr(w.device)), "codebook_prepare")


def vq_assign(z: torch.Tensor, weight: torch.Tensor, state: Optional[QuantizerState] = None,
              algo: int = _lib.ASSIGN_AUTO, return_distance: bool = False):
    """Fused distance + argmin (models/vqvae.py:30-38).  z: [B,C,T] fp32 CUDA (any strides).
    Returns int32 indices [B,T] (and the winning fp32 distances if asked)."""
    lib = _lib.load()
    z = _f32(z)
    B, C, T = z.shape
    w = weight.detach()
    K, D = w.shape
    if C != D:
        raise RuntimeError(f"vqb200.vq_assign: channel dim {C} != embedding_dim {D}")
    if state is None:
        state = QuantizerState(K, D, z.device)
    state.refresh(w)
    idx = torch.empty((B, T), dtype=torch.int32, device=z.device)
    best = torch.empty((B, T), dtype=torch.float32, device=z.device) if return_distance else None
    ws = state.assign_workspace(B * T)
    sB, sC, sT = z.stride()
    with torch.cuda.device(z.device):
        check(lib.vqb200_vq_assign(ptr(z), B, C, T, sB, sC, sT, ptr(w), ptr(state.ee), ptr(state.image),
                                   ptr(state.info), K, ptr(idx), ptr(best), ptr(ws), c_size_t(ws.numel()),
                                   algo, stream_ptr(z.device)), "vq_assign")
    return (idx, best) if return_distance else idx


class RVQConfig:
    """Everything `_RVQFn` needs besides the differentiable tensors."""

    def __init__(self, states: Sequence[QuantizerState], ema_cluster_size: Sequence[Optional[torch.Tensor]],
                 ema_w: Sequence[Optional[torch.Tensor]], commitment_cost: float, use_ema: bool, decay: float,
                 training: bool, plain: bool, algo: int = _lib.ASSIGN_AUTO, eps: float = 1e-5):
        self.states = list(states)
        self.ema_cluster_size = list(ema_cluster_size)
        self.ema_w = list(ema_w)
        self.commitment_cost = float(commitment_cost)
        self.use_ema = bool(use_ema)
        self.decay = float(decay)
        self.training = bool(training)
        self.plain = bool(plain)          # True: a bare VectorQuantizer (S == 1, output is st_0 itself)
        self.algo = algo
        self.eps = eps


class _RVQFn(torch.autograd.Function):
    """S residual stages of VQ (S == 1: plain VectorQuantizer).  models/vqvae.py:28-76, 87-108.

    forward per stage s on residual r_s (r_0 = z):
        idx_s = argmin_k d(r_s, E_s)                          K1  vqb200_vq_assign
        training & EMA: stats = [sum x | count]               K3a vqb200_ema_accumulate
                        all-reduce(stats) over the DP group       (dist.all_reduce_stats)
                        decay / Laplace / normalise, E_s <- ...  K3b vqb200_ema_finalize   (in place)
        st_s = r_s + (E_s[idx_s] - r_s); r_{s+1} = r_s - st_s; out += st_s; sse_s
                                                              K2  vqb200_vq_gather_st
        loss_s, perplexity_s, dcr_s                               vqb200_vq_metrics
    backward (closed form of autograd):
        dL/dz   = g + g_loss*c*2/(N*D)*(z - E_0[idx_0])           K2b vqb200_vq_backward_input
        dL/dE_s = g_loss*2/(N*D)*sum_{idx_s=k}(E_s[k] - r_s)      K3 (mode 1) + vqb200_vq_backward_codebook
                  (standard VQ only; EMA codebooks get grad None like the reference)
    """

    @staticmethod
    def forward(ctx, z, cfg: RVQConfig, *weights):
        lib = _lib.load()
        z = _f32(z)
        B, C, T = z.shape
        S = len(weights)
        dev = z.device
        N = B * T
        f32 = dict(dtype=torch.float32, device=dev)
        idx = torch.empty((S, B, T), dtype=torch.int32, device=dev)
        m3 = torch.empty((S + 1, 3), **f32)          # row S: aggregate written by the single-launch kernels
        out = torch.empty((B, C, T), **f32)
        residuals = [z] + [torch.empty((B, C, T), **f32) for _ in range(S - 1)]
        ema_train = cfg.training and cfg.use_ema
        world = _dist.world_size() if ema_train else 1
        e0_snapshot = None
        Ks = (ctypes.c_int64 * S)(*[weights[s].shape[0] for s in range(S)])
        for s in range(S):
            if weights[s].shape[1] != C:
                raise RuntimeError(f"vqb200: channel dim {C} != embedding_dim {weights[s].shape[1]}")
        # launch-bound shapes: the whole stage loop in ONE cluster kernel (csrc/rvq_small.cu).  Not for standard-VQ
        # training (its codebook gradient needs the per-stage residuals) nor under data parallelism (the EMA
        # statistics must cross GPUs between assignment and update).
        fused = (cfg.algo == _lib.ASSIGN_AUTO and N > 0 and not _dist.enabled()
                 and (cfg.use_ema or not any(ctx.needs_input_grad[2:]))
                 and bool(lib.vqb200_rvq_small_eligible(N, C, S, Ks)))
        # under data parallelism the single-launch kernel runs with the exchange inside (csrc/rvq_small.cu,
        # grid_barrier_world) when the caller promised equal shards -- all ranks must choose alike
        px = _dist.peer_exchange() if ema_train else None
        fused_peer = False
        if (not fused and cfg.algo == _lib.ASSIGN_AUTO and N > 0 and px is not None and px.device == dev
                and _dist.uniform_shards()):
            with torch.cuda.device(dev):
                local_ok = (bool(lib.vqb200_rvq_small_peer_eligible(N, C, S, Ks))
                            and int(lib.vqb200_rvq_small_stats_floats(S, Ks)) * 4 <= px.slot_bytes)
            # the eligibility probe is device-local (co-resident clusters differ between GPUs): every rank must take
            # the same path, or slot parity and barrier pattern of the exchange diverge
            fused_peer = _dist.agree(("rvq_small_peer", N, C, S, tuple(int(k) for k in Ks)), local_ok)
            fused = fused_peer
        if fused:
            with torch.cuda.device(dev):
                stream = stream_ptr(dev)
                Wd = [weights[s].detach() for s in range(S)]
                Es = (ctypes.c_void_p * S)(*[w.data_ptr() for w in Wd])
                if ema_train:
                    Cs = (ctypes.c_void_p * S)(*[cfg.ema_cluster_size[s].data_ptr() for s in range(S)])
                    Ws = (ctypes.c_void_p * S)(*[cfg.ema_w[s].data_ptr() for s in range(S)])
                else:
                    Cs = Ws = None
                st0 = cfg.states[0]
                need = int(lib.vqb200_rvq_small_workspace_floats(S, Ks))
                ws = getattr(st0, "_small_ws", None)
                if ws is None or ws.numel() < need:
                    ws = torch.empty(need, **f32)
                    st0._small_ws = ws
                sse = torch.empty(S, dtype=torch.float64, device=dev)
                sB, sC, sT = z.stride()
                if fused_peer:
                    epoch0, _mine, slots = px.next_slot(S)
                    check(lib.vqb200_rvq_small_forward_peer(ptr(z), B, C, T, sB, sC, sT, S, Es, Cs, Ws, Ks, c_double(cfg.decay),
                                                            c_double(cfg.eps), c_float(cfg.commitment_cost), ptr(ws), ptr(sse),
                                                            ptr(idx), ptr(out), ptr(m3), slots, px.flags, px.rank, px.world,
                                                            ctypes.c_uint32(epoch0), N * px.world, stream),
                          "rvq_small_forward_peer")
                else:
                    check(lib.vqb200_rvq_small_forward(ptr(z), B, C, T, sB, sC, sT, S, Es, Cs, Ws, Ks, c_double(cfg.decay),
                                                       c_double(cfg.eps), c_float(cfg.commitment_cost),
                                                       1 if cfg.use_ema else 0, 1 if cfg.training else 0, ptr(ws), ptr(sse),
                                                       ptr(idx), ptr(out), ptr(m3), stream), "rvq_small_forward")
                for s in range(S):
                    cfg.states[s].invalidate()      # |E|^2 / tile image were not refreshed by the fused kernel
                if ema_train and ctx.needs_input_grad[0]:
                    e0_snapshot = Wd[0].clone()
        with torch.cuda.device(dev):
            stream = stream_ptr(dev)
            for s in range(S if not fused else 0):
                st = cfg.states[s]
                W = weights[s].detach()
                K, D = W.shape
                st.refresh(W)
                r = residuals[s]
                sB, sC, sT = r.stride()
                if N > 0 and s == 0:
                    ws = st.assign_workspace(N)
                    check(lib.vqb200_vq_assign(ptr(r), B, C, T, sB, sC, sT, ptr(W), ptr(st.ee), ptr(st.image),
                                               ptr(st.info), K, ptr(idx[s]), None, ptr(ws), c_size_t(ws.numel()),
                                               cfg.algo, stream), "vq_assign")
                elif N > 0:
                    # r_s = r_{s-1} - st_{s-1} (with the codebook stage s-1 has just updated) and its assignment in
                    # one call: on the tensor-core path the residual update rides inside the assignment kernel
                    ws = st.assign_workspace(N)
                    rp = residuals[s - 1]
                    pB, pC, pT = rp.stride()
                    Wp = weights[s - 1].detach()
                    check(lib.vqb200_vq_assign_residual(ptr(rp), B, C, T, pB, pC, pT, ptr(Wp), ptr(idx[s - 1]),
                                                        Wp.shape[0], ptr(r), ptr(W), ptr(st.ee), ptr(st.image),
                                                        ptr(st.info), K, ptr(idx[s]), ptr(ws), c_size_t(ws.numel()),
                                                        cfg.algo, stream), "vq_assign_residual")
                if ema_train and px is not None and px.device == dev and px.fits(K, D):
                    # K3a into this rank's slot of the symmetric buffer; K3b barriers and sums every rank's slot over
                    # NVLink itself (csrc/peer.cu): no collective launch between the two
                    epoch, my_slot, slots = px.next_slot()
                    check(lib.vqb200_ema_accumulate(ptr(r), B, C, T, sB, sC, sT, ptr(idx[s]), None, K,
                                                    ctypes.c_void_p(my_slot), 0, stream), "ema_accumulate")
                    check(lib.vqb200_ema_finalize_peer(slots, px.flags, px.rank, px.world, ctypes.c_uint32(epoch),
                                                       ptr(st.cnt), ptr(cfg.ema_cluster_size[s]), ptr(cfg.ema_w[s]),
                                                       ptr(W), K, D, c_double(cfg.decay), c_double(cfg.eps), ptr(st.ee),
                                                       ptr(st.image), ptr(st.info), ptr(st.scratch), stream),
                          "ema_finalize_peer")
                elif ema_train:
                    check(lib.vqb200_ema_accumulate(ptr(r), B, C, T, sB, sC, sT, ptr(idx[s]), None, K,
                                                    ptr(st.stats), 0, stream), "ema_accumulate")
                    _dist.all_reduce_stats(st.stats)
                    check(lib.vqb200_ema_finalize(ptr(st.stats), ptr(cfg.ema_cluster_size[s]), ptr(cfg.ema_w[s]),
                                                  ptr(W), K, D, c_double(cfg.decay), c_double(cfg.eps), ptr(st.ee),
                                                  ptr(st.image), ptr(st.info), ptr(st.scratch), stream),
                          "ema_finalize")
                if ema_train:
                    st.mark_fresh(W)
                    if s == 0 and ctx.needs_input_grad[0]:
                        e0_snapshot = W.clone()       # a later call may update E_0 before backward runs
                else:
                    check(lib.vqb200_vq_histogram(ptr(idx[s]), N, K, ptr(st.cnt), stream), "vq_histogram")
                if cfg.plain and S == 1:
                    # bare VectorQuantizer: out = st_0 itself (keeps the sign of zero of the reference's :63)
                    check(lib.vqb200_vq_gather_st(ptr(r), B, C, T, sB, sC, sT, ptr(W), ptr(idx[s]), K, ptr(out),
                                                  None, None, 0, ptr(st.sse), stream), "vq_gather_st")
                    check(lib.vqb200_vq_metrics(ptr(st.cnt), K, max(N * world, 1), ptr(st.sse), max(N * C, 1),
                                                c_float(cfg.commitment_cost), 1 if cfg.use_ema else 0, ptr(m3[s]),
                                                stream), "vq_metrics")
            if not fused and not (cfg.plain and S == 1):
                # all stages at once: out = ((0 + st_0) + st_1) + ... and the S loss sums, from z + indices + codebooks
                Es = (ctypes.c_void_p * S)(*[weights[s].detach().data_ptr() for s in range(S)])
                Is = (ctypes.c_void_p * S)(*[idx[s].data_ptr() for s in range(S)])
                sse = torch.empty(S, dtype=torch.float64, device=dev)
                sB, sC, sT = z.stride()
                rc = lib.vqb200_rvq_output_chain(ptr(z), B, C, T, sB, sC, sT, S, Es, Is, Ks, ptr(out), ptr(sse), None, stream)
                if rc == -5:        # VQB200_EWORKSPACE: arbitrary view, needs a running-residual scratch
                    scratch = torch.empty((B, C, T), **f32)
                    rc = lib.vqb200_rvq_output_chain(ptr(z), B, C, T, sB, sC, sT, S, Es, Is, Ks, ptr(out), ptr(sse),
                                                     ptr(scratch), stream)
                check(rc, "rvq_output_chain")
                for s in range(S):
                    st = cfg.states[s]
                    check(lib.vqb200_vq_metrics(ptr(st.cnt), weights[s].shape[0], max(N * world, 1), ptr(sse[s:s + 1]),
                                                max(N * C, 1), c_float(cfg.commitment_cost), 1 if cfg.use_ema else 0,
                                                ptr(m3[s]), stream), "vq_metrics")
        if fused:                       # the single-launch kernel has already reduced over the stages (row S)
            loss, ppl, dcr = m3[S, 0], m3[S, 1], m3[S, 2]
        elif S == 1:
            loss, ppl, dcr = m3[0, 0], m3[0, 1], m3[0, 2]
            if not cfg.plain:
                ppl, dcr = m3[:S, 1].mean(), m3[:S, 2].mean()
        else:
            loss, ppl, dcr = m3[:S, 0].sum(), m3[:S, 1].mean(), m3[:S, 2].mean()
        m3 = m3[:S]
        ctx.cfg = cfg
        ctx.S = S
        ctx.shape = (B, C, T)
        if cfg.use_ema:
            # the kernels update E_0 through raw pointers (autograd's version counter does not see it): backward must not
            # read a codebook that a later training forward has moved on, so it always gets its own copy
            if e0_snapshot is None and ctx.needs_input_grad[0]:
                e0_snapshot = weights[0].detach().clone()
            ctx.save_for_backward(z, idx, e0_snapshot if e0_snapshot is not None else weights[0].detach())
        else:
            ctx.save_for_backward(z, idx, *residuals[1:], *weights)
        ctx.mark_non_differentiable(ppl, dcr, idx, m3)
        return out, loss, ppl, dcr, idx, m3

    @staticmethod
    def backward(ctx, g_out, g_loss, *_unused):
        lib = _lib.load()
        cfg: RVQConfig = ctx.cfg
        S = ctx.S
        B, C, T = ctx.shape
        N = B * T
        saved = ctx.saved_tensors
        z, idx = saved[0], saved[1]
        dev = z.device
        if cfg.use_ema:
            e0 = saved[2]
            residuals = [z]
            weights: List[torch.Tensor] = []
        else:
            residuals = [z] + list(saved[2:2 + S - 1])
            weights = list(saved[2 + S - 1:])
            e0 = weights[0]
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        numel = max(N * C, 1)
        gz = None
        grads_w: List[Optional[torch.Tensor]] = [None] * S
        with torch.cuda.device(dev):
            stream = stream_ptr(dev)
            if ctx.needs_input_grad[0]:
                gz = torch.empty((B, C, T), dtype=torch.float32, device=dev)
                K0 = e0.shape[0]
                if g_out is not None:
                    g_out = _f32(g_out)
                    gs = g_out.stride()
                else:
                    gs = (0, 0, 0)
                sB, sC, sT = z.stride()
                check(lib.vqb200_vq_backward_input(ptr(g_out), gs[0], gs[1], gs[2], ptr(z), B, C, T, sB, sC, sT,
                                                   ptr(e0.detach()), ptr(idx[0]), K0, ptr(g_loss),
                                                   c_float(cfg.commitment_cost * 2.0 / numel), ptr(gz), stream),
                      "vq_backward_input")
            if not cfg.use_ema:
                for s in range(S):
                    if not ctx.needs_input_grad[2 + s]:
                        continue
                    W = weights[s].detach()
                    K, D = W.shape
                    r = residuals[s]
                    sB, sC, sT = r.stride()
                    stats = torch.empty(K * (D + 1), dtype=torch.float32, device=dev)
                    check(lib.vqb200_ema_accumulate(ptr(r), B, C, T, sB, sC, sT, ptr(idx[s]), ptr(W), K,
                                                    ptr(stats), 1, stream), "ema_accumulate(mode=1)")
                    gE = torch.empty_like(W)
                    check(lib.vqb200_vq_backward_codebook(ptr(stats), K, D, ptr(g_loss), c_float(2.0 / numel),
                                                          ptr(gE), stream), "vq_backward_codebook")
                    grads_w[s] = gE
        return (gz, None, *grads_w)


def rvq_quantize(z: torch.Tensor, weights: Sequence[torch.Tensor], cfg: RVQConfig):
    """-> (quantized [B,C,T], loss, perplexity, dcr, indices int32 [S,B,T], per-stage metrics [S,3])."""
    return _RVQFn.apply(z, cfg, *weights)


def vq_quantize(z: torch.Tensor, weight: torch.Tensor, cfg: RVQConfig):
    """Single-stage VQ; -> (quantized, loss, perplexity, dcr, indices int32 [B,T])."""
    q, loss, ppl, dcr, idx, _ = _RVQFn.apply(z, cfg, weight)
    return q, loss, ppl, dcr, idx[0]


# --------------------------------------------------------------------------------------------
# FSQ / LFQ elementwise stages
# --------------------------------------------------------------------------------------------
_UNIQ_WS = {}


def _unique_workspace(device: torch.device) -> torch.Tensor:
    # one per (device, stream): calls on different streams must not share the bitmap / ticket / entropy words
    key = (device.type, device.index, int(torch.cuda.current_stream(device).cuda_stream))
    ws = _UNIQ_WS.get(key)
    if ws is None:
        ws = torch.empty(int(_lib.load().vqb200_unique_workspace_bytes()), dtype=torch.uint8, device=device)
        _UNIQ_WS[key] = ws
    return ws


class _FSQFn(torch.autograd.Function):
    """z_hard = z + (round(z) - z) with identity gradient; int64 index pack; #unique on device
    (models/vqvae.py:127-147,152-154)."""

    @staticmethod
    def forward(ctx, z_e, basis, codebook_size):
        lib = _lib.load()
        z_e = _f32(z_e).contiguous()
        B, d, T = z_e.shape
        dev = z_e.device
        z_hard = torch.empty_like(z_e)
        idx = torch.empty((B, T), dtype=torch.int64, device=dev)
        m2 = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.vqb200_fsq_forward(ptr(z_e), B, d, T, ptr(basis), int(codebook_size), ptr(z_hard), ptr(idx),
                                         ptr(_unique_workspace(dev)), ptr(m2), stream_ptr(dev)), "fsq_forward")
        ctx.mark_non_differentiable(idx, m2)
        return z_hard, idx, m2

    @staticmethod
    def backward(ctx, g, *_):
        return g, None, None


def fsq_round(z_e: torch.Tensor, basis: torch.Tensor, codebook_size: int):
    """-> (z_hard [B,d,T], indices int64 [B,T], metrics = [perplexity(#unique), dcr])."""
    if basis.dtype != torch.int32:
        basis = basis.to(torch.int32)
    return _FSQFn.apply(z_e, basis.contiguous(), codebook_size)


class _LFQFn(torch.autograd.Function):
    """z_q = z_e + (sign(z_e) - z_e), entropy loss, int64 index pack, #unique (models/vqvae.py:171-191)."""

    @staticmethod
    def forward(ctx, z_e, weight):
        lib = _lib.load()
        z_e = _f32(z_e).contiguous()
        B, d, T = z_e.shape
        dev = z_e.device
        z_q = torch.empty_like(z_e)
        idx = torch.empty((B, T), dtype=torch.int64, device=dev)
        m3 = torch.empty(3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.vqb200_lfq_forward(ptr(z_e), B, d, T, c_float(weight), ptr(z_q), ptr(idx),
                                         ptr(_unique_workspace(dev)), ptr(m3), stream_ptr(dev)), "lfq_forward")
        ctx.save_for_backward(z_e)
        ctx.weight = float(weight)
        loss = m3[0]
        ctx.mark_non_differentiable(idx, m3)
        return z_q, loss, idx, m3

    @staticmethod
    def backward(ctx, g_zq, g_loss, *_):
        lib = _lib.load()
        (z_e,) = ctx.saved_tensors
        dev = z_e.device
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        if g_zq is not None:
            g_zq = _f32(g_zq).contiguous()
        g = torch.empty_like(z_e)
        with torch.cuda.device(dev):
            check(lib.vqb200_lfq_backward(ptr(z_e), ptr(g_zq), ptr(g_loss), z_e.numel(), c_float(ctx.weight),
                                          ptr(g), stream_ptr(dev)), "lfq_backward")
        return g, None


def lfq_sign(z_e: torch.Tensor, entropy_loss_weight: float = 0.1):
    """-> (z_q [B,d,T], loss, indices int64 [B,T], metrics = [loss, perplexity(#unique), dcr])."""
    return _LFQFn.apply(z_e, entropy_loss_weight)


class _ProjFusedFn(torch.autograd.Function):
    """Whole FSQ / LFQ module forward in one kernel (project_in -> round | sign -> project_out, indices, metrics,
    LFQ entropy loss) and its autograd in one more (models/vqvae.py:126-154, :170-194)."""

    @staticmethod
    def forward(ctx, z, w_in, b_in, w_out, b_out, is_lfq, basis, codebook_size, weight):
        lib = _lib.load()
        B, D, T = z.shape
        d = w_in.shape[0]
        dev = z.device
        out = torch.empty_like(z)
        z_e = torch.empty((B, d, T), dtype=torch.float32, device=dev)
        idx = torch.empty((B, T), dtype=torch.int64, device=dev)
        m = torch.empty(3 if is_lfq else 2, dtype=torch.float32, device=dev)
        wi, wo = w_in.detach().contiguous(), w_out.detach().contiguous()
        bi, bo = b_in.detach().contiguous(), b_out.detach().contiguous()
        with torch.cuda.device(dev):
            if is_lfq:
                check(lib.vqb200_lfq_fused_forward(ptr(z), B, D, T, ptr(wi), ptr(bi), ptr(wo), ptr(bo), d, c_float(weight),
                                                   ptr(out), ptr(z_e), ptr(idx), ptr(_unique_workspace(dev)), ptr(m),
                                                   stream_ptr(dev)), "lfq_fused_forward")
            else:
                check(lib.vqb200_fsq_fused_forward(ptr(z), B, D, T, ptr(wi), ptr(bi), ptr(wo), ptr(bo), d, ptr(basis),
                                                   int(codebook_size), ptr(out), ptr(z_e), ptr(idx),
                                                   ptr(_unique_workspace(dev)), ptr(m), stream_ptr(dev)), "fsq_fused_forward")
        ctx.save_for_backward(z, z_e, wi, wo)
        ctx.is_lfq, ctx.weight = bool(is_lfq), float(weight)
        # FSQ has no loss term (models/vqvae.py:137): the module returns its own constant 0; this slot is an unused,
        # uninitialised placeholder (no fill kernel)
        loss = m[0] if is_lfq else m.new_empty(())
        ctx.mark_non_differentiable(idx, m, z_e)
        return out, loss, idx, m, z_e

    @staticmethod
    def backward(ctx, g_out, g_loss, *_):
        lib = _lib.load()
        z, z_e, wi, wo = ctx.saved_tensors
        B, D, T = z.shape
        d = wi.shape[0]
        dev = z.device
        g_out = torch.zeros_like(z) if g_out is None else _f32(g_out).contiguous()
        if g_loss is not None:
            g_loss = g_loss.to(torch.float32).contiguous()
        elif ctx.is_lfq:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_z = torch.empty_like(z)
        grads = torch.empty(int(lib.vqb200_proj_fused_grad_floats(D, d)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.vqb200_proj_fused_backward(1 if ctx.is_lfq else 0, ptr(g_out), ptr(z), ptr(z_e), B, D, T, ptr(wi),
                                                 ptr(wo), d, ptr(g_loss) if ctx.is_lfq else None, c_float(ctx.weight),
                                                 ptr(g_z), ptr(grads), stream_ptr(dev)), "proj_fused_backward")
        g_wi = grads[:d * D].view(d, D, 1)
        g_bi = grads[d * D:d * D + d]
        g_wo = grads[d * D + d:2 * d * D + d].view(D, d, 1)
        g_bo = grads[2 * d * D + d:]
        return g_z, g_wi, g_bi, g_wo, g_bo, None, None, None, None


def proj_fused_eligible(z: torch.Tensor, d: int) -> bool:
    """True when the one-pass kernels apply: contiguous fp32 [B,64,T] (T <= 128) on CUDA and d <= 16."""
    if z.dim() != 3 or not z.is_cuda or z.dtype != torch.float32 or not z.is_contiguous() or z.shape[0] == 0:
        return False
    B, D, T = z.shape
    return z.data_ptr() % 16 == 0 and bool(_lib.load().vqb200_proj_fused_eligible(B, D, int(d), T))


def fsq_module_fused(z, project_in, project_out, basis, codebook_size):
    """-> (out [B,D,T], indices int64 [B,T], metrics [perplexity(#unique), dcr], z_e [B,d,T])."""
    if basis.dtype != torch.int32:
        basis = basis.to(torch.int32)
    out, _, idx, m2, z_e = _ProjFusedFn.apply(z, project_in.weight, project_in.bias, project_out.weight, project_out.bias,
                                              False, basis.contiguous(), codebook_size, 0.0)
    return out, idx, m2, z_e


def lfq_module_fused(z, project_in, project_out, entropy_loss_weight):
    """-> (out [B,D,T], loss, indices int64 [B,T], metrics [loss, perplexity(#unique), dcr], z_e [B,d,T])."""
    return _ProjFusedFn.apply(z, project_in.weight, project_in.bias, project_out.weight, project_out.bias,
                              True, None, 0, float(entropy_loss_weight))
