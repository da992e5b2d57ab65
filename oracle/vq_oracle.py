"""numpy fp32 restatement of the reference quantizers -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to /root/reference).
All arithmetic is done in float32 with the reference's association order wherever that
order is visible in the Python source; library reductions (GEMM, sum) are free to differ
in summation order, exactly as MKL / cuBLAS differ from each other.

Pinned against the unmodified reference modules through tests/golden/*.npz
(generator: tests/golden/make_golden.py).  Not importable from the product path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

F = np.float32


def _f(x) -> np.float32:
    return np.float32(x)


# --------------------------------------------------------------------------------------
# VectorQuantizer                                                  models/vqvae.py:10-76
# --------------------------------------------------------------------------------------
@dataclass
class VQState:
    """Parameters / buffers of one `VectorQuantizer` (models/vqvae.py:12-26)."""
    embedding: np.ndarray                      # [K, D]  nn.Embedding weight      (:19-20)
    ema_cluster_size: Optional[np.ndarray] = None   # [K]    buffer, only if use_ema   (:24)
    ema_w: Optional[np.ndarray] = None              # [K, D] buffer, only if use_ema   (:25-26)
    commitment_cost: float = 0.25
    use_ema: bool = False
    decay: float = 0.99

    def copy(self) -> "VQState":
        return VQState(self.embedding.copy(),
                       None if self.ema_cluster_size is None else self.ema_cluster_size.copy(),
                       None if self.ema_w is None else self.ema_w.copy(),
                       self.commitment_cost, self.use_ema, self.decay)

    @property
    def K(self) -> int:
        return int(self.embedding.shape[0])

    @property
    def D(self) -> int:
        return int(self.embedding.shape[1])


def vq_distances(flat: np.ndarray, E: np.ndarray) -> np.ndarray:
    """d[n,k] = fl(fl(|x_n|^2 + |E_k|^2) - 2*(x_n . E_k))     models/vqvae.py:34-36."""
    flat = np.asarray(flat, F)
    E = np.asarray(E, F)
    xx = np.sum(flat * flat, axis=1, keepdims=True, dtype=F)
    ee = np.sum(E * E, axis=1, dtype=F)
    return (xx + ee) - _f(2.0) * (flat @ E.T)


def _argmin_rows(dist: np.ndarray) -> np.ndarray:
    """First index of the row minimum; a NaN wins (torch.argmin)  models/vqvae.py:38."""
    return np.argmin(dist, axis=1).astype(np.int64)   # numpy has the same NaN/tie rule


def vq_forward(z: np.ndarray, st: VQState, training: bool = True,
               force_indices: Optional[np.ndarray] = None,
               stats_reduce: Optional[Callable] = None,
               keep_distances: bool = False) -> Dict[str, np.ndarray]:
    """`VectorQuantizer.forward`                               models/vqvae.py:28-76.

    z: [B, C, T] fp32.  Mutates `st` in place exactly like the module mutates its buffers
    (EMA update happens BEFORE the codeword gather, :43-52).

    force_indices: teacher-forcing hook for the parity harness (SURVEY.md §8c rule 2): use
      these assignments instead of the oracle's own argmin.
    stats_reduce: hook `(cnt, dw) -> (cnt, dw)` that emulates the per-stage data-parallel
      all-reduce of the EMA statistics.
    """
    z = np.asarray(z, F)
    B, C, T = z.shape
    K, D = st.K, st.D
    assert C == D
    x = np.ascontiguousarray(z.transpose(0, 2, 1))            # :30
    flat = x.reshape(-1, D)                                   # :32
    N = flat.shape[0]

    dist = vq_distances(flat, st.embedding)                   # :34-36
    idx = _argmin_rows(dist)                                  # :38
    if force_indices is not None:
        idx = np.asarray(force_indices, np.int64).reshape(-1)
    cnt = np.bincount(idx, minlength=K).astype(F)             # encodings.sum(0)  :44 / :71

    if training and st.use_ema:                               # :43-50
        dw = np.zeros((K, D), np.float64)
        np.add.at(dw, idx, flat.astype(np.float64))           # == encodings.t() @ flat  :45
        dw = dw.astype(F)
        cnt_upd = cnt
        if stats_reduce is not None:
            cnt_upd, dw = stats_reduce(cnt_upd, dw)
        st.ema_cluster_size[...] = st.ema_cluster_size * _f(st.decay) + cnt_upd * _f(1 - st.decay)  # :46
        st.ema_w[...] = st.ema_w * _f(st.decay) + dw * _f(1 - st.decay)                              # :47
        n = np.sum(st.ema_cluster_size, dtype=F)                                                     # :48
        cluster = (st.ema_cluster_size + _f(1e-5)) / (n + _f(K * 1e-5)) * n                          # :49
        st.embedding[...] = st.ema_w / cluster[:, None]                                              # :50

    q = st.embedding[idx].reshape(x.shape)                    # one_hot @ E (post-update) :52
    diff = q - x
    mse = np.mean(diff * diff, dtype=F)                       # F.mse_loss :56 / :59-60
    c = _f(st.commitment_cost)
    loss = c * mse if st.use_ema else mse + c * mse           # :57 / :61
    st_val = x + (q - x)                                      # straight-through value :63

    p = cnt / _f(N)                                           # torch.mean(encodings, 0) :66
    perplexity = np.exp(-np.sum(p * np.log(p + _f(1e-10)), dtype=F))   # :67
    active = _f(np.count_nonzero(cnt > 0))                    # :71
    dcr = _f(1.0) - active / _f(K)                            # :72

    out = {
        "loss": F(loss),
        "quantized": np.ascontiguousarray(st_val.transpose(0, 2, 1)),   # :76
        "perplexity": F(perplexity), "dcr": F(dcr),
        "indices": idx.reshape(B, T), "counts": cnt,
        "q": q, "x": x,                                       # cache for vq_backward
    }
    if keep_distances:
        out["distances"] = dist
    return out


def vq_backward(cache: Dict[str, np.ndarray], st: VQState, g_quantized: np.ndarray,
                g_loss: float = 1.0) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Closed form of what autograd computes for models/vqvae.py:52-63.

    dL/dx = g + g_loss*c*(2/(N*D))*(x - q)            (EMA and standard)
    dL/dE[k] = g_loss*(2/(N*D))*sum_{n: idx_n=k}(q_n - x_n)   (standard VQ only; EMA: None)
    Returns (grad_z [B,C,T], grad_embedding [K,D] | None).
    """
    x, q = cache["x"], cache["q"]
    numel = _f(x.size)
    g = np.asarray(g_quantized, F).transpose(0, 2, 1)         # to [B,T,C]
    coef = _f(g_loss) * _f(st.commitment_cost) * (_f(2.0) / numel)
    gx = g + coef * (x - q)
    gE = None
    if not st.use_ema:
        idx = cache["indices"].reshape(-1)
        acc = np.zeros((st.K, st.D), np.float64)
        contrib = (_f(g_loss) * (_f(2.0) / numel)) * (q - x).reshape(-1, st.D)
        np.add.at(acc, idx, contrib.astype(np.float64))
        gE = acc.astype(F)
    return np.ascontiguousarray(gx.transpose(0, 2, 1)), gE


# --------------------------------------------------------------------------------------
# ResidualVQ                                                      models/vqvae.py:78-108
# --------------------------------------------------------------------------------------
def rvq_forward(z: np.ndarray, stages: Sequence[VQState], training: bool = True,
                force_indices: Optional[Sequence[np.ndarray]] = None,
                stats_reduce: Optional[Callable] = None) -> Dict[str, object]:
    """`ResidualVQ.forward`                                    models/vqvae.py:87-108."""
    residual = np.asarray(z, F)
    out = None
    total = _f(0.0)
    ppl: List[np.float32] = []
    dcr: List[np.float32] = []
    caches = []
    for s, stg in enumerate(stages):                          # :94
        fi = None if force_indices is None else force_indices[s]
        r = vq_forward(residual, stg, training, fi, stats_reduce)   # :95
        residual = residual - r["quantized"]                  # :96
        out = (_f(0.0) + r["quantized"]) if out is None else out + r["quantized"]   # :88,:97
        total = F(total + r["loss"])                          # :98
        ppl.append(r["perplexity"]); dcr.append(r["dcr"])     # :99-100
        caches.append(r)
    return {
        "loss": F(total), "quantized": out,
        "perplexity": F(np.mean(np.asarray(ppl, F), dtype=F)),     # :104
        "dcr": F(np.mean(np.asarray(dcr, F), dtype=F)),            # :105
        "indices": np.stack([c["indices"] for c in caches], 0),    # [S,B,T]
        "stage": caches,
    }


def rvq_backward(fw: Dict[str, object], stages: Sequence[VQState], g_quantized: np.ndarray,
                 g_loss: float = 1.0) -> Tuple[np.ndarray, List[Optional[np.ndarray]]]:
    """Reverse-mode sweep through models/vqvae.py:94-98 (what autograd does).

    Because each stage's straight-through output has identity Jacobian w.r.t. its input,
    d r_{s+1} / d r_s = 0 and only stage 0's commitment term reaches z.
    """
    caches = fw["stage"]
    g_out = np.asarray(g_quantized, F)
    G_next = np.zeros_like(g_out)                             # dL/d r_S (unused residual)
    gEs: List[Optional[np.ndarray]] = [None] * len(stages)
    for s in range(len(stages) - 1, -1, -1):
        g_st = g_out - G_next                                 # from `out +=` and `residual -=`
        g_through, gEs[s] = vq_backward(caches[s], stages[s], g_st, g_loss)
        G_next = G_next + g_through                           # direct path + through the layer
    return G_next, gEs


# --------------------------------------------------------------------------------------
# 1x1 convolutions used by FSQ / LFQ                     models/vqvae.py:118-119,164-165
# --------------------------------------------------------------------------------------
def conv1x1(x: np.ndarray, weight: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """nn.Conv1d(cin, cout, 1): x [B,cin,T], weight [cout,cin,1], bias [cout] -> [B,cout,T]."""
    w = np.asarray(weight, F)[:, :, 0]
    return (np.einsum("oc,bct->bot", w, np.asarray(x, F), dtype=F, optimize=False)
            + np.asarray(bias, F)[None, :, None]).astype(F)


# --------------------------------------------------------------------------------------
# FSQ                                                            models/vqvae.py:110-154
# --------------------------------------------------------------------------------------
def fsq_basis(levels: Sequence[int]) -> np.ndarray:
    """_basis = cumprod([1] + levels[:-1]) as int32              models/vqvae.py:122."""
    return np.cumprod(np.asarray([1] + list(levels[:-1]), np.int64)).astype(np.int32)


def fsq_quantize(z_e: np.ndarray, levels: Sequence[int]) -> Dict[str, object]:
    """Elementwise stage of FSQ on the post-`project_in` tensor z_e [B,d,T].

    Unbounded half-to-even rounding (no tanh / clamp), float multiply-sum with `_basis`
    then truncation to int64                                models/vqvae.py:127-147,152-154.
    """
    z_e = np.asarray(z_e, F)
    basis = fsq_basis(levels).astype(F)
    zt = z_e.transpose(0, 2, 1)                               # [B,T,d]  :127
    z_hard = zt + (np.rint(zt) - zt)                          # _round_ste value :153-154
    prod = z_hard * basis                                     # :141
    s = prod[..., 0]
    for i in range(1, prod.shape[-1]):
        s = s + prod[..., i]
    with np.errstate(invalid="ignore"):
        indices = np.trunc(s).astype(np.int64)                # .long()
    unique = int(np.unique(indices).size)                     # :142
    size = int(np.prod(levels))
    return {
        "z_hard": np.ascontiguousarray(z_hard.transpose(0, 2, 1)),   # [B,d,T]
        "indices": indices,                                    # [B,T] int64
        "unique": unique,
        "perplexity": F(float(unique)),                        # :146
        "dcr": F(1.0 - (unique / size)),                       # :144,:147
    }


def fsq_forward(z: np.ndarray, levels: Sequence[int], w_in, b_in, w_out, b_out, z_e=None) -> Dict[str, object]:
    """`FSQ.forward` incl. both 1x1 projections                 models/vqvae.py:125-150.
    z_e: optional override of the post-`project_in` tensor (bit-exactness of the rounding is only
    defined from that tensor onward, SURVEY.md §7 "FSQ/LFQ bit-exactness")."""
    z_e = conv1x1(z, w_in, b_in) if z_e is None else np.asarray(z_e, F)   # :126
    r = fsq_quantize(z_e, levels)
    r["z_e"] = z_e
    r["quantized"] = conv1x1(r["z_hard"], w_out, b_out)       # :133-134
    r["loss"] = F(0.0)                                        # :136
    return r


# --------------------------------------------------------------------------------------
# LFQ                                                            models/vqvae.py:156-194
# --------------------------------------------------------------------------------------
def lfq_quantize(z_e: np.ndarray, entropy_loss_weight: float = 0.1) -> Dict[str, object]:
    """Elementwise stage of LFQ on the post-`project_in` tensor z_e [B,d,T]
                                                            models/vqvae.py:171-191."""
    z_e = np.asarray(z_e, F)
    d = z_e.shape[1]
    z_q = np.where(z_e > 0, _f(1.0), _f(-1.0))                # :171
    z_st = z_e + (z_q - z_e)                                  # :172
    prob = (_f(1.0) / (_f(1.0) + np.exp(-z_e))).astype(F)     # sigmoid :175
    ent = -(prob * np.log(prob + _f(1e-6)) + (_f(1.0) - prob) * np.log(_f(1.0) - prob + _f(1e-6)))  # :176
    loss = -np.mean(ent, dtype=F) * _f(entropy_loss_weight)   # :177
    bits = (z_st > 0).astype(np.int64).transpose(0, 2, 1)     # [B,T,d] :184
    basis = (2 ** np.arange(d)).astype(np.int64)              # :167
    indices = (bits * basis).sum(-1)                          # :185
    unique = int(np.unique(indices).size)                     # :186
    return {
        "z_q": z_st.astype(F), "indices": indices, "unique": unique,
        "loss": F(loss),
        "perplexity": F(float(unique)),                        # :190
        "dcr": F(1.0 - (unique / (2 ** d))),                   # :188,:191
    }


def lfq_backward_ze(z_e: np.ndarray, g_zq: np.ndarray, g_loss: float = 1.0,
                    entropy_loss_weight: float = 0.1) -> np.ndarray:
    """dL/dz_e of models/vqvae.py:171-177 in closed form.

    Straight-through identity plus the derivative of -w*mean(H_b(sigmoid(z_e))).
    """
    z_e = np.asarray(z_e, np.float64)
    p = 1.0 / (1.0 + np.exp(-z_e))
    dlt = 1e-6
    dH = -(np.log(p + dlt) + p / (p + dlt) - np.log(1 - p + dlt) - (1 - p) / (1 - p + dlt))
    g = np.asarray(g_zq, np.float64) + g_loss * (-entropy_loss_weight / z_e.size) * dH * p * (1 - p)
    return g.astype(F)


def lfq_forward(z: np.ndarray, w_in, b_in, w_out, b_out, entropy_loss_weight: float = 0.1) -> Dict[str, object]:
    """`LFQ.forward` incl. both 1x1 projections                 models/vqvae.py:169-194."""
    z_e = conv1x1(z, w_in, b_in)                              # :170
    r = lfq_quantize(z_e, entropy_loss_weight)
    r["z_e"] = z_e
    r["quantized"] = conv1x1(r["z_q"], w_out, b_out)          # :179
    return r


# --------------------------------------------------------------------------------------
# HybridVQ                                                       models/vqvae.py:199-241
# --------------------------------------------------------------------------------------
def hybrid_forward(z: np.ndarray, levels: Sequence[int], w_in, b_in, w_out, b_out,
                   stages: Sequence[VQState], training: bool = True,
                   force_indices: Optional[Sequence[np.ndarray]] = None,
                   stats_reduce: Optional[Callable] = None, z_e=None) -> Dict[str, object]:
    """`HybridVQ.forward`: FSQ base + RVQ on the residual       models/vqvae.py:219-241."""
    z = np.asarray(z, F)
    f = fsq_forward(z, levels, w_in, b_in, w_out, b_out, z_e)  # :221
    residual = z - f["quantized"]                             # :224
    r = rvq_forward(residual, stages, training, force_indices, stats_reduce)   # :228
    return {
        "loss": r["loss"],                                    # returns loss_vq only :241
        "quantized": f["quantized"] + r["quantized"],         # :231
        "perplexity": f["perplexity"], "dcr": f["dcr"],       # :236-237
        "rvq_ppl": r["perplexity"],                           # :238
        "fsq": f, "rvq": r, "residual": residual,
    }
