"""Drop-in quantizer modules: same constructors, forward contract and state_dict keys as the
reference's `models/vqvae.py:10-259`, computed by the vqb200 CUDA kernels.

Contract kept:
  forward(z: fp32 [B,C,T], any strides, CUDA) -> (loss: 0-dim tensor with grad,
                                                   quantized: contiguous [B,C,T] with straight-through grad,
                                                   metrics: dict[str -> 0-dim device tensor])
Indices (never returned by the reference) are exposed as `module.last_indices`.
CPU tensors are rejected: there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .functional import (QuantizerState, RVQConfig, rvq_quantize, fsq_round, lfq_sign, proj_fused_eligible,
                         fsq_module_fused, lfq_module_fused)


def _require_cuda(z: torch.Tensor, who: str) -> torch.Tensor:
    if not isinstance(z, torch.Tensor) or not z.is_cuda:
        raise RuntimeError(f"{who}: input must be a CUDA tensor -- vqb200 has no CPU fallback "
                           "(the reference's CPU path lives in /root/reference, the test oracle in oracle/)")
    if z.dim() != 3:
        raise RuntimeError(f"{who}: expected [B, C, T], got {tuple(z.shape)}")
    # the kernels compute in fp32 like the reference; a half / bf16 latent (autocast) is widened differentiably
    return z if z.dtype == torch.float32 else z.float()


class VectorQuantizer(nn.Module):
    """Standard / EMA VQ (reference: models/vqvae.py:10-76)."""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost=0.25, use_ema=False, decay=0.99):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.commitment_cost = commitment_cost
        self.use_ema = use_ema
        # same RNG call sequence as the reference (:19-26) so that equal seeds give equal initial state
        self.embedding = nn.Embedding(num_embeddings, embedding_dim)
        self.embedding.weight.data.uniform_(-1 / num_embeddings, 1 / num_embeddings)
        if use_ema:
            self.decay = decay
            self.register_buffer("ema_cluster_size", torch.zeros(num_embeddings))
            self.register_buffer("ema_w", torch.empty(num_embeddings, embedding_dim).normal_())
        self.assign_algo = _lib.ASSIGN_AUTO
        self.last_indices: Optional[torch.Tensor] = None
        self._states: Dict[tuple, QuantizerState] = {}

    # -- derived device state ------------------------------------------------------------
    def _state(self, device: torch.device) -> QuantizerState:
        key = (device.type, device.index)
        st = self._states.get(key)
        if st is None:
            st = QuantizerState(self.num_embeddings, self.embedding_dim, device)
            self._states[key] = st
        return st

    def invalidate_cache(self) -> None:
        """Call after writing `embedding.weight.data` behind autograd's back."""
        for st in self._states.values():
            st.invalidate()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_cache()

    def revive_dead_codes(self, inputs: torch.Tensor, threshold: float = 1e-3, seed: int = 0,
                          usage: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Opt-in codebook health op (NOT in the reference): every code whose usage
        (`ema_cluster_size` by default, or the given per-code tensor, e.g. last step's histogram) is below `threshold`
        is re-seeded from a row of `inputs` [B,C,T] chosen by a hash of (seed, code).  Returns the number of
        replaced codes as an int32 device tensor (no host sync)."""
        inputs = _require_cuda(inputs, "VectorQuantizer.revive_dead_codes")
        if usage is None:
            if not self.use_ema:
                raise RuntimeError("revive_dead_codes: pass `usage` for a non-EMA quantizer")
            usage = self.ema_cluster_size
        lib = _lib.load()
        B, C, T = inputs.shape
        if C != self.embedding_dim:
            raise RuntimeError(f"revive_dead_codes: channel dim {C} != embedding_dim {self.embedding_dim}")
        dev = inputs.device
        revived = torch.zeros(1, dtype=torch.int32, device=dev)
        sB, sC, sT = inputs.stride()
        w = self.embedding.weight.detach()
        with torch.cuda.device(dev):
            _lib.check(lib.vqb200_codebook_revive(
                _lib.ptr(inputs.detach()), B, C, T, sB, sC, sT, _lib.ptr(usage.detach().to(torch.float32).contiguous()),
                _lib.c_float(threshold), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(w),
                _lib.ptr(self.ema_cluster_size) if self.use_ema else None, _lib.ptr(self.ema_w) if self.use_ema else None,
                self.num_embeddings, _lib.ptr(revived), _lib.stream_ptr(dev)), "codebook_revive")
        self.invalidate_cache()
        return revived

    def _config(self, device, plain: bool) -> RVQConfig:
        return RVQConfig([self._state(device)],
                         [self.ema_cluster_size if self.use_ema else None],
                         [self.ema_w if self.use_ema else None],
                         self.commitment_cost, self.use_ema, self.decay if self.use_ema else 0.0,
                         self.training, plain, self.assign_algo)

    def forward(self, inputs):
        inputs = _require_cuda(inputs, "VectorQuantizer")
        w = self.embedding.weight
        if w.device != inputs.device:
            raise RuntimeError("VectorQuantizer: module and input live on different devices")
        cfg = self._config(inputs.device, plain=True)
        quantized, loss, ppl, dcr, idx, _ = rvq_quantize(inputs, [w], cfg)
        self.last_indices = idx[0]
        return loss, quantized, {"perplexity": ppl, "dcr": dcr}


class ResidualVQ(nn.Module):
    """Residual VQ, all stages driven from one autograd node (reference: models/vqvae.py:78-108)."""

    def __init__(self, num_quantizers, num_embeddings, embedding_dim, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantizer(num_embeddings, embedding_dim, **kwargs)
                                     for _ in range(num_quantizers)])
        self.last_indices: Optional[torch.Tensor] = None

    def forward(self, x):
        x = _require_cuda(x, "ResidualVQ")
        layers: List[VectorQuantizer] = list(self.layers)
        if not layers:
            raise RuntimeError("ResidualVQ: no quantizer layers")
        first = layers[0]
        dev = x.device
        cfg = RVQConfig([l._state(dev) for l in layers],
                        [l.ema_cluster_size if l.use_ema else None for l in layers],
                        [l.ema_w if l.use_ema else None for l in layers],
                        first.commitment_cost, first.use_ema, first.decay if first.use_ema else 0.0,
                        self.training, False, first.assign_algo)
        quantized, loss, ppl, dcr, idx, _ = rvq_quantize(x, [l.embedding.weight for l in layers], cfg)
        self.last_indices = idx
        for s, l in enumerate(layers):
            l.last_indices = idx[s]
        return loss, quantized, {"perplexity": ppl, "dcr": dcr}


class FSQ(nn.Module):
    """Finite scalar quantization with UNBOUNDED rounding, as the reference does it
    (models/vqvae.py:110-154): no tanh / level clamp; `levels` only define `_basis` and codebook_size."""

    def __init__(self, levels, input_dim, hidden_dim):
        super().__init__()
        self.levels = levels
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.fsq_dim = len(levels)
        self.project_in = nn.Conv1d(input_dim, self.fsq_dim, 1)
        self.project_out = nn.Conv1d(self.fsq_dim, input_dim, 1)
        self.register_buffer("_levels", torch.tensor(levels, dtype=torch.int32))
        basis = torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.int64), dim=0)
        self.register_buffer("_basis", basis.to(torch.int32))
        self.codebook_size = math.prod(levels)
        self.fuse_projections = True          # False: stock conv1d + the elementwise kernel (arbitrary layouts use it anyway)
        self.last_indices: Optional[torch.Tensor] = None
        self.last_z_e: Optional[torch.Tensor] = None

    def forward(self, z):
        z = _require_cuda(z, "FSQ")
        conv_ok = self.project_in.bias is not None and self.project_out.bias is not None
        if self.fuse_projections and conv_ok and proj_fused_eligible(z, self.fsq_dim):
            # one pass over z: both 1x1 projections ride inside the quantisation kernel
            z_out, idx, m2, z_e = fsq_module_fused(z, self.project_in, self.project_out, self._basis, self.codebook_size)
            self.last_indices, self.last_z_e = idx, z_e
            loss = torch.zeros((), dtype=torch.float32, device=z.device)
            return loss, z_out, {"perplexity": m2[0], "dcr": m2[1]}
        z_e = self.project_in(z)                                       # [B, d, T]
        self.last_z_e = z_e.detach()
        z_hard, idx, m2 = fsq_round(z_e, self._basis, self.codebook_size)
        z_out = self.project_out(z_hard)
        self.last_indices = idx
        loss = torch.zeros((), dtype=torch.float32, device=z.device)
        return loss, z_out, {"perplexity": m2[0], "dcr": m2[1]}


class LFQ(nn.Module):
    """Lookup-free (binary) quantization (reference: models/vqvae.py:156-194)."""

    def __init__(self, input_dim, codebook_dim=10, entropy_loss_weight=0.1):
        super().__init__()
        self.input_dim = input_dim
        self.codebook_dim = codebook_dim
        self.entropy_loss_weight = entropy_loss_weight
        self.codebook_size = 2 ** codebook_dim
        self.project_in = nn.Conv1d(input_dim, codebook_dim, 1)
        self.project_out = nn.Conv1d(codebook_dim, input_dim, 1)
        self.register_buffer("_basis", 2 ** torch.arange(codebook_dim))
        self.fuse_projections = True
        self.last_indices: Optional[torch.Tensor] = None
        self.last_z_e: Optional[torch.Tensor] = None

    def forward(self, z):
        z = _require_cuda(z, "LFQ")
        conv_ok = self.project_in.bias is not None and self.project_out.bias is not None
        if self.fuse_projections and conv_ok and proj_fused_eligible(z, self.codebook_dim):
            out, loss, idx, m3, z_e = lfq_module_fused(z, self.project_in, self.project_out, self.entropy_loss_weight)
            self.last_indices, self.last_z_e = idx, z_e
            return loss, out, {"perplexity": m3[1], "dcr": m3[2]}
        z_e = self.project_in(z)
        self.last_z_e = z_e.detach()
        z_q, loss, idx, m3 = lfq_sign(z_e, self.entropy_loss_weight)
        out = self.project_out(z_q)
        self.last_indices = idx
        return loss, out, {"perplexity": m3[1], "dcr": m3[2]}


class HybridVQ(nn.Module):
    """FSQ base + 4-stage EMA residual VQ on what FSQ misses (reference: models/vqvae.py:199-241)."""

    def __init__(self, hidden_dim, fsq_levels=[8, 5, 5, 5], vq_codebook_size=1024):
        super().__init__()
        self.fsq = FSQ(levels=fsq_levels, input_dim=hidden_dim, hidden_dim=hidden_dim)
        self.vq = ResidualVQ(num_quantizers=4, num_embeddings=vq_codebook_size, embedding_dim=hidden_dim,
                             commitment_cost=0.25, use_ema=True)

    def forward(self, z):
        z = _require_cuda(z, "HybridVQ")
        _, z_fsq, m_fsq = self.fsq(z)
        residual = z - z_fsq
        loss_vq, z_vq, m_vq = self.vq(residual)
        z_out = z_fsq + z_vq
        return loss_vq, z_out, {"perplexity": m_fsq["perplexity"], "dcr": m_fsq["dcr"],
                                "rvq_ppl": m_vq["perplexity"]}


class IdentityVQ(nn.Module):
    """method='ae': pass-through (reference: models/vqvae.py:243-259)."""

    def forward(self, z):
        dev = z.device
        return (torch.tensor(0.0, device=dev), z,
                {"perplexity": torch.tensor(1.0, device=dev), "dcr": torch.tensor(0.0, device=dev)})
