"""Parity at the sizes the bench actually times -- GPU tests through the nn.Module / C ABI.

* cfg3 at the bench size: ResidualVQ S=4, K=1024 on z [1 000 000, 64, 10] (2.56 GB, byte offsets beyond 2^31), one
  training step; 65 536 sampled rows are re-done by the oracle stage by stage (teacher-forced on the engine's upstream
  indices, pre-update codebooks for the assignment, post-update codebooks for the gather -- models/vqvae.py:43-52,94-98).
* the K = 65 536 end of the cfg5 sweep against the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import vq_distances, check_indices, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _need(gb):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory")


def test_cfg3_bench_size_sampled():
    _need(40)
    import vqb200
    S, K, D, B, T = 4, 1024, 64, 1_000_000, 10
    N = B * T
    torch.manual_seed(42)
    mod = vqb200.ResidualVQ(S, K, D, use_ema=True).to(DEV).train()
    with torch.no_grad():
        for l in mod.layers:                       # the bench's start state (bench.py::build_module)
            l.embedding.weight.normal_(0, 0.25)
            l.ema_w.copy_(l.embedding.weight)
            l.ema_cluster_size.fill_(1.0)
    gen = torch.Generator(device=DEV).manual_seed(1237)
    z = torch.randn((B, D, T), generator=gen, device=DEV).mul_(0.5)
    assert z.numel() * 4 > 2 ** 31
    for _ in range(2):                             # two steps: the second one starts from a trained-on-data codebook
        E_pre = [l.embedding.weight.detach().clone() for l in mod.layers]
        cs_pre = [float(l.ema_cluster_size.sum()) for l in mod.layers]
        with torch.no_grad():
            loss, q, met = mod(z)
    idx = mod.last_indices                         # [S, B, T] int32
    assert idx.shape == (S, B, T) and int(idx.min()) >= 0 and int(idx.max()) < K
    E_post = [l.embedding.weight.detach() for l in mod.layers]
    for s, l in enumerate(mod.layers):             # conservation: sum(cs') = decay*sum(cs) + (1-decay)*N
        expect = 0.99 * cs_pre[s] + 0.01 * N
        assert abs(float(l.ema_cluster_size.sum()) - expect) / expect < 1e-5
    # ---- 65 536 sampled rows, including the first and the last ones ----
    rs = np.random.default_rng(7)
    rows = np.unique(np.concatenate([rs.integers(0, N, 65536 - 64), np.arange(32), np.arange(N - 32, N)]))
    rt = torch.from_numpy(rows).to(DEV)
    b, t = rt // T, rt % T
    x = z[b, :, t].cpu().numpy()                   # [n, D]
    q_eng = q[b, :, t].cpu().numpy()
    idx_eng = idx[:, b, t].cpu().numpy().astype(np.int64)       # [S, n]
    r = x.copy()
    out = np.zeros_like(x)
    total_flips = 0
    for s in range(S):
        d = vq_distances(r, E_pre[s].cpu().numpy())
        ref = np.argmin(d, axis=1)
        flips, bad, bad_rows = check_indices(idx_eng[s], ref, d)
        assert bad == 0, f"stage {s}: {bad} index mismatches beyond the 1e-6 near-tie rule (rows {rows[bad_rows[:5]]})"
        total_flips += flips
        qs = E_post[s].cpu().numpy()[idx_eng[s]]   # gather from the UPDATED codebook with the engine's indices
        st = r + (qs - r)
        out = out + st
        r = r - st
    assert_close(q_eng, out, 1e-5, "quantized rows at the bench size")
    print(f"cfg3 bench size: {len(rows)} sampled rows, {total_flips} benign flips, loss {float(loss):.6f}, "
          f"perplexity {float(met['perplexity']):.2f}")


def test_cfg5_k65536_parity():
    """The K = 65 536 corner of the cfg5 sweep (timed in bench.py) against the oracle on 2 048 vectors."""
    _need(4)
    from test_gpu_parity import _mods, _fresh_state, _load_vq, _vq_step
    vq = _mods()
    K, D, N = 65536, 64, 2048
    st = _fresh_state(K, D, True, 11, "normal")
    mod = vq.VectorQuantizer(K, D, use_ema=True).to(DEV)
    _load_vq(mod, st)
    rng = np.random.default_rng(1240)
    z = rng.standard_normal((N, D, 1)).astype(np.float32)
    _vq_step(mod, st, z, rng.standard_normal((N, D, 1)).astype(np.float32), 1.0, True)


@pytest.mark.parametrize("B,T,K,regime,zmode,perm", [
    (1, 1, 1024, "small", "plain", False),            # a single row
    (1, 129, 1024, "small", "plain", True),           # one row more than a row tile
    (13, 3, 777, "small", "plain", False),            # ragged everything
    (4099, 10, 1030, "small", "plain", False),        # streaming kernel, 6 codes in the last code tile
    (20000, 1, 1025, "small", "plain", True),
    (40960, 10, 1024, "small", "rowscales", False),   # every row at its own magnitude, 1e-18 .. 1e18
    (40960, 10, 2048, "small", "rowscales", False),
    (40960, 10, 1024, "normal", "huge", False),       # 1e17-sized data and codes
    (40960, 10, 1024, "small", "zeros", False),       # exactly-zero rows and components
    (40960, 10, 1024, "small", "nonfinite", False),   # NaN / Inf rows, one NaN code
    (40960, 10, 2048, "small", "nonfinite", False),
    (40960, 1, 1024, "dead", "rowscales", True),      # a few live codes among 3e4-times larger dead ones
    (100000, 10, 1024, "dup", "rowscales", False),    # duplicated codes: ties everywhere
])
def test_tensor_core_assignment_equals_exact_kernel_on_edge_regimes(B, T, K, regime, zmode, perm):
    """The tcgen05 filter + exact finish must give the SAME indices as the exact CUDA-core kernel (which the other
    tests hold to the oracle), bit for bit, whatever the data looks like: the filter may only ever hand rows over."""
    import vqb200
    from vqb200 import _lib
    dev = torch.device("cuda:0")
    D = 64
    torch.manual_seed(0)
    W = torch.randn(K, D, device=dev)
    if regime == "small":
        W *= 0.3
    elif regime == "dup":
        W[K // 2:] = W[: K - K // 2]
    elif regime == "dead":
        W[::3] *= 3e4
    z = (0.5 * torch.randn(B, T, D, device=dev)).permute(0, 2, 1) if perm else 0.5 * torch.randn(B, D, T, device=dev)
    if zmode == "rowscales":
        z = z * (10.0 ** (36 * torch.rand(B, 1, T, device=dev) - 18))
    elif zmode == "huge":
        z, W = z * 1e17, W * 1e17
    elif zmode == "zeros":
        z = z * (torch.rand(B, 1, T, device=dev) > 0.33) * (torch.rand_like(z) > 0.2)
    elif zmode == "nonfinite":
        z = z.clone()
        z.view(-1)[::977] = float("nan"); z.view(-1)[5::1999] = float("inf"); z.view(-1)[11::2999] = float("-inf")
        W = W.clone(); W[K // 3, 7] = float("nan")
    st = vqb200.QuantizerState(K, D, dev)
    i_exact = vqb200.vq_assign(z, W, st, _lib.ASSIGN_SIMT)
    i_tc = vqb200.vq_assign(z, W, st, _lib.ASSIGN_TC)
    torch.cuda.synchronize()
    assert int(st._assign_ws[:8].view(torch.int32)[1]) == 0          # error word of the kernel
    assert torch.equal(i_exact, i_tc)
