"""GPU parity tests: the CUDA engine (through the C ABI) against the numpy oracle, the golden
fixtures recorded from the unmodified reference, and size-independent properties at full size.

Tolerances are BASELINE.json's: FSQ/LFQ indices bit-exact; VQ/RVQ indices equal except where the
oracle's two candidate fp32 distances differ by < 1e-6 relative; values within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import (VQState, vq_forward, vq_backward, rvq_forward, rvq_backward, fsq_quantize,
                    lfq_quantize, lfq_backward_ze, hybrid_forward, vq_distances, check_indices, assert_close)
from _golden import load, vq_state_from

pytestmark = pytest.mark.gpu
# the stock torch.nn 1x1 convolutions around FSQ/LFQ must run in true fp32 for CPU-recorded fixtures
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
TOL = 1e-5
DEV = "cuda:0"


def _mods():
    import vqb200
    return vqb200


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=DEV)


def N_(t):
    return t.detach().cpu().numpy()


def _load_vq(mod, st: VQState):
    with torch.no_grad():
        mod.embedding.weight.copy_(T(st.embedding))
        if st.use_ema:
            mod.ema_cluster_size.copy_(T(st.ema_cluster_size))
            mod.ema_w.copy_(T(st.ema_w))


def _cmp_state(mod, st: VQState, what):
    assert_close(N_(mod.embedding.weight), st.embedding, TOL, what + ".embedding", rows=True)
    if st.use_ema:
        assert_close(N_(mod.ema_cluster_size), st.ema_cluster_size, TOL, what + ".ema_cluster_size")
        assert_close(N_(mod.ema_w), st.ema_w, TOL, what + ".ema_w", rows=True)


def _vq_step(mod, st, z_np, g_np, g_loss, training, z_tensor=None, check_grad=True):
    """One engine step vs the oracle teacher-forced on the engine's indices.  Returns #flips."""
    mod.train(training)
    pre = st.copy()
    free = vq_forward(z_np, pre, training, keep_distances=True)       # oracle's own choice + distances
    z = (T(z_np) if z_tensor is None else z_tensor).requires_grad_(True)
    loss, q, met = mod(z)
    idx = N_(mod.last_indices).astype(np.int64)
    flips, bad, rows = check_indices(idx, free["indices"], free["distances"])
    assert bad == 0, f"{bad} non-benign index flips (of {flips}) at rows {rows[:8]}"
    ref = vq_forward(z_np, st, training, force_indices=idx)
    assert q.is_contiguous() and q.shape == z.shape
    assert_close(N_(q), ref["quantized"], TOL, "quantized")
    assert_close(N_(loss), ref["loss"], TOL, "loss")
    assert_close(N_(met["perplexity"]), ref["perplexity"], TOL, "perplexity")
    assert_close(N_(met["dcr"]), ref["dcr"], 1e-6, "dcr")
    _cmp_state(mod, st, "state")
    if check_grad:
        mod.zero_grad()
        (loss * g_loss + (q * T(g_np)).sum()).backward()
        gz, gE = vq_backward(ref, st, g_np, g_loss)
        assert_close(N_(z.grad), gz, TOL, "grad_z")
        if st.use_ema:
            assert mod.embedding.weight.grad is None
        else:
            assert_close(N_(mod.embedding.weight.grad), gE, TOL, "grad_embedding")
    return flips, ref


# ----------------------------------------------------------------------------------------------
# golden fixtures (recorded from the reference)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["vq_std_small", "vq_ema_fresh", "vq_ema_k1024_perm", "vq_std_ragged"])
def test_vq_golden(name):
    vq = _mods()
    g = load(name)
    use_ema = bool(g["use_ema"])
    st = vq_state_from(g, "init.", use_ema)
    mod = vq.VectorQuantizer(int(g["K"]), int(g["D"]), use_ema=use_ema).to(DEV)
    _load_vq(mod, st)
    total = 0
    for s in range(int(g["steps"])):
        p = f"s{s}."
        z_np = g[p + "z"]
        zt = None
        if name.endswith("_perm"):      # present the permuted T'=1 view exactly like the transformer encoder
            zt = T(np.ascontiguousarray(z_np.transpose(0, 2, 1))).permute(0, 2, 1)
            assert zt.stride(1) == 1 and not zt.is_contiguous() or z_np.shape[2] == 1
        flips, ref = _vq_step(mod, st, z_np, g[p + "g"], 1.7, True, zt)
        total += flips
        if total == 0:                  # no flip so far: compare straight to the reference's outputs
            np.testing.assert_array_equal(N_(mod.last_indices), g[p + "indices"])
            assert_close(N_(mod.embedding.weight), g[p + "after.embedding"], TOL, "E vs reference", rows=True)
            assert_close(ref["quantized"], g[p + "quantized"], TOL, "quantized vs reference")
    if "eval.z" in g:
        before = N_(mod.embedding.weight).copy()
        _vq_step(mod, st, g["eval.z"], None, 1.0, False, check_grad=False)
        np.testing.assert_array_equal(before, N_(mod.embedding.weight))


def _rvq_modules(vq, g, prefix, S, K, D, use_ema):
    mod = vq.ResidualVQ(S, K, D, use_ema=use_ema).to(DEV)
    stages = [vq_state_from(g, f"{prefix}layers.{i}.", use_ema) for i in range(S)]
    for l, st in zip(mod.layers, stages):
        _load_vq(l, st)
    return mod, stages


def _rvq_step(mod, stages, z_np, g_np, g_loss, training=True, z_tensor=None):
    mod.train(training)
    z = (T(z_np) if z_tensor is None else z_tensor).requires_grad_(True)
    loss, q, met = mod(z)
    idx = N_(mod.last_indices).astype(np.int64)
    # stage-wise flip check: replay the oracle forced on the engine's earlier stages
    S = len(stages)
    probe = [s.copy() for s in stages]
    residual = np.asarray(z_np, np.float32)
    flips_total = 0
    for s in range(S):
        flat = np.ascontiguousarray(residual.transpose(0, 2, 1)).reshape(-1, stages[s].D)
        dist = vq_distances(flat, probe[s].embedding)
        flips, bad, rows = check_indices(idx[s], np.argmin(dist, 1), dist)
        assert bad == 0, f"stage {s}: {bad} non-benign flips"
        flips_total += flips
        r = vq_forward(residual, probe[s], training, force_indices=idx[s])
        residual = residual - r["quantized"]
    ref = rvq_forward(z_np, stages, training, force_indices=list(idx))
    assert_close(N_(q), ref["quantized"], TOL, "quantized")
    assert_close(N_(loss), ref["loss"], TOL, "loss")
    assert_close(N_(met["perplexity"]), ref["perplexity"], TOL, "perplexity")
    assert_close(N_(met["dcr"]), ref["dcr"], 1e-6, "dcr")
    for l, st in zip(mod.layers, stages):
        _cmp_state(l, st, "stage")
    if g_np is not None:
        mod.zero_grad()
        (loss * g_loss + (q * T(g_np)).sum()).backward()
        gz, gEs = rvq_backward(ref, stages, g_np, g_loss)
        assert_close(N_(z.grad), gz, TOL, "grad_z")
        for l, st, gE in zip(mod.layers, stages, gEs):
            if st.use_ema:
                assert l.embedding.weight.grad is None
            else:
                assert_close(N_(l.embedding.weight.grad), gE, TOL, "grad_embedding")
    return flips_total, ref


@pytest.mark.parametrize("name", ["rvq_ema", "rvq_std"])
def test_rvq_golden(name):
    vq = _mods()
    g = load(name)
    S, K, D, use_ema = int(g["S"]), int(g["K"]), int(g["D"]), bool(g["use_ema"])
    mod, stages = _rvq_modules(vq, g, "init.", S, K, D, use_ema)
    total = 0
    for s in range(int(g["steps"])):
        p = f"s{s}."
        flips, ref = _rvq_step(mod, stages, g[p + "z"], g[p + "g"], 0.9)
        total += flips
        if total == 0:
            np.testing.assert_array_equal(N_(mod.last_indices), g[p + "indices"])
            assert_close(ref["quantized"], g[p + "quantized"], TOL, "vs reference")
            assert_close(N_(mod.layers[-1].embedding.weight), g[p + f"after.layers.{S-1}.embedding"], TOL, "E", rows=True)


@pytest.mark.parametrize("name", ["fsq_module", "fsq_module_x30", "fsq_crafted"])
def test_fsq_golden_bit_exact(name):
    vq = _mods()
    g = load(name)
    levels = [int(v) for v in g["levels"]]
    D = int(g["D"])
    mod = vq.FSQ(levels, D, D).to(DEV)
    mod.load_state_dict({k[6:]: T(v, dtype=None) for k, v in g.items() if k.startswith("state.")})
    # elementwise kernel on the reference's own post-projection tensor: bit exact
    z_hard, idx, m2 = vq.fsq_round(T(g["z_e"]), mod._basis, mod.codebook_size)
    np.testing.assert_array_equal(N_(idx), g["indices"])
    np.testing.assert_array_equal(N_(z_hard), g["z_hard"])
    assert float(m2[0]) == float(g["perplexity"])
    assert_close(N_(m2[1]), g["dcr"], 1e-6, "dcr")
    # whole module incl. cuDNN 1x1 projections + gradients
    z = T(g["z"]).requires_grad_(True)
    loss, q, met = mod(z)
    (loss + (q * T(g["g"])).sum()).backward()
    assert float(loss) == 0.0
    assert_close(N_(q), g["quantized"], 2e-5, "quantized")
    assert_close(N_(z.grad), g["grad_z"], 2e-5, "grad_z")
    for k in ("project_in.weight", "project_in.bias", "project_out.weight", "project_out.bias"):
        assert_close(N_(dict(mod.named_parameters())[k].grad), g["grad." + k], 5e-5, k)


@pytest.mark.parametrize("name", ["lfq_module", "lfq_crafted"])
def test_lfq_golden_bit_exact(name):
    vq = _mods()
    g = load(name)
    w = float(g["entropy_loss_weight"])
    ze = T(g["z_e"]).requires_grad_(True)
    z_q, loss, idx, m3 = vq.lfq_sign(ze, w)
    np.testing.assert_array_equal(N_(idx), g["indices"])
    np.testing.assert_array_equal(N_(z_q), g["z_q"])
    assert_close(N_(loss), g["loss"], TOL, "loss")
    assert float(m3[1]) == float(g["perplexity"])
    assert_close(N_(m3[2]), g["dcr"], 1e-6, "dcr")
    (loss * float(g["g_loss"]) + (z_q * T(g["g_zq"])).sum()).backward()
    assert_close(N_(ze.grad), g["grad_z_e"], TOL, "grad_z_e")
    mod = vq.LFQ(int(g["D"]), int(g["d"]), w).to(DEV)
    mod.load_state_dict({k[6:]: T(v, dtype=None) for k, v in g.items() if k.startswith("state.")})
    z = T(g["z"]).requires_grad_(True)
    loss, q, met = mod(z)
    (loss * 1.3 + (q * T(g["g"])).sum()).backward()
    assert_close(N_(q), g["quantized"], 2e-5, "quantized")
    assert_close(N_(z.grad), g["grad_z"], 2e-5, "grad_z")
    assert float(met["perplexity"]) == float(g["perplexity"])


@pytest.mark.parametrize("name", ["hybrid_perm", "hybrid_t10"])
def test_hybrid_golden(name):
    vq = _mods()
    g = load(name)
    D, K, S = int(g["D"]), int(g["K"]), int(g["S"])
    mod = vq.HybridVQ(D, [8, 5, 5, 5], vq_codebook_size=K).to(DEV)
    mod.load_state_dict({k[5:]: T(v, dtype=None) for k, v in g.items() if k.startswith("init.")})
    stages = [VQState(g[f"init.vq.layers.{i}.embedding.weight"].copy(), g[f"init.vq.layers.{i}.ema_cluster_size"].copy(),
                      g[f"init.vq.layers.{i}.ema_w"].copy(), 0.25, True, 0.99) for i in range(S)]
    levels = [8, 5, 5, 5]
    mod.train()
    clean = True
    for s in range(int(g["steps"])):
        p = f"s{s}."
        z_np = g[p + "z"]
        if name.endswith("_perm"):
            z = T(np.ascontiguousarray(z_np.transpose(0, 2, 1))).permute(0, 2, 1).requires_grad_(True)
        else:
            z = T(z_np).requires_grad_(True)
        loss, q, met = mod(z)
        idx = N_(mod.vq.last_indices).astype(np.int64)
        z_e = N_(mod.fsq.last_z_e)      # the engine's own post-project_in tensor (fused kernel)
        np.testing.assert_array_equal(N_(mod.fsq.last_indices), fsq_quantize(z_e, levels)["indices"])
        ref = hybrid_forward(z_np, levels, g["init.fsq.project_in.weight"], g["init.fsq.project_in.bias"],
                             g["init.fsq.project_out.weight"], g["init.fsq.project_out.bias"], stages, True,
                             force_indices=list(idx), z_e=z_e)
        assert_close(N_(q), ref["quantized"], 2e-5, p + "quantized")
        assert_close(N_(loss), ref["loss"], 2e-5, p + "loss")
        assert_close(N_(met["rvq_ppl"]), ref["rvq_ppl"], TOL, p + "rvq_ppl")
        assert float(met["perplexity"]) == float(ref["perplexity"])
        clean = clean and np.array_equal(idx, g[p + "indices"])
        if clean:
            assert_close(N_(q), g[p + "quantized"], 2e-5, p + "quantized vs reference")
            assert_close(N_(loss), g[p + "loss"], 2e-5, p + "loss vs reference")
            mod.zero_grad()
            (loss + (q * T(g[p + "g"])).sum()).backward()
            assert_close(N_(z.grad), g[p + "grad_z"], 5e-5, p + "grad_z vs reference")
            assert_close(N_(mod.fsq.project_in.weight.grad), g[p + "grad.fsq.project_in.weight"], 1e-4, "grad project_in")


@pytest.mark.parametrize("name,method", [("model_resnet_no_down_ema", "ema"), ("model_resnet_no_down_hybrid", "hybrid")])
def test_whole_model_golden(name, method):
    """Drop-in DualMotionVQVAE loads the reference's state_dict strictly and reproduces its eval outputs."""
    from models.vqvae import DualMotionVQVAE
    g = load(name)
    m = DualMotionVQVAE(human_input_dim=12, robot_input_dim=7, hidden_dim=16, codebook_size=64,
                        arch="resnet_no_down", method=method, n_layers=2, window_size=10).to(DEV)
    sd = {k[6:]: T(v, dtype=None) for k, v in g.items() if k.startswith("state.")}
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        o = m(x_robot=T(g["x_robot"]), x_human=T(g["x_human"]))
    # encoders run on cuDNN here vs MKL in the fixture: allow conv-level noise, then the quantizer may
    # legitimately flip on near-ties, so compare the latent first and outputs loosely
    assert_close(N_(o["robot"]["z_e"]), g["robot.z_e"], 1e-4, "z_e")
    assert_close(N_(o["robot"]["recon"]), g["robot.recon"], 5e-3, "recon")
    assert_close(N_(o["human"]["retargeted"]), g["human.retargeted"], 5e-3, "retargeted")
    assert_close(N_(o["robot"]["loss_vq"]), g["robot.loss_vq"], 1e-3, "loss_vq")
    assert set(o["robot"]["metrics"]) == {k.split(".")[-1] for k in g if k.startswith("robot.metrics.")}


# ----------------------------------------------------------------------------------------------
# oracle parity at the BASELINE.json shapes the oracle can still finish in seconds
# ----------------------------------------------------------------------------------------------
def _fresh_state(K, D, use_ema, seed, regime):
    rng = np.random.default_rng(seed)
    if regime == "init":            # U(+-1/K) init, ema_w ~ N(0,1), cs = 0  (exact ties, then 1e5-sized codes)
        E = rng.uniform(-1 / K, 1 / K, (K, D)).astype(np.float32)
    elif regime == "normal":
        E = rng.standard_normal((K, D)).astype(np.float32)
    else:                           # "small": 0.3 * N(0,1)
        E = (0.3 * rng.standard_normal((K, D))).astype(np.float32)
    cs = np.zeros(K, np.float32) if use_ema else None
    w = rng.standard_normal((K, D)).astype(np.float32) if use_ema else None
    if use_ema and regime != "init":
        cs = rng.uniform(0.5, 20.0, K).astype(np.float32)
        w = (E * cs[:, None]).astype(np.float32)
    return VQState(E, cs, w, 0.25, use_ema, 0.99)


@pytest.mark.parametrize("regime", ["init", "normal", "small"])
@pytest.mark.parametrize("algo", ["simt", "auto"])
def test_cfg1_shape_three_steps(regime, algo):
    """cfg1: z_e [3686,64,10] (N=36 860), K=1024, EMA, three consecutive training steps."""
    vq = _mods()
    K, D, B, Tt = 1024, 64, 3686, 10
    st = _fresh_state(K, D, True, 5, regime)
    mod = vq.VectorQuantizer(K, D, use_ema=True).to(DEV)
    mod.assign_algo = vq._lib.ASSIGN_SIMT if algo == "simt" else vq._lib.ASSIGN_AUTO
    _load_vq(mod, st)
    rng = np.random.default_rng(17)
    flips = 0
    for step in range(3):
        z = rng.standard_normal((B, D, Tt)).astype(np.float32)
        g = rng.standard_normal((B, D, Tt)).astype(np.float32)
        f, _ = _vq_step(mod, st, z, g, 1.0, True)
        flips += f
    print(f"cfg1 {regime}/{algo}: benign flips {flips}")


def test_cfg2_hybrid_three_steps():
    """cfg2: HybridVQ on z [512,64,1] presented as the permuted view, 3 training steps, fwd+bwd."""
    vq = _mods()
    torch.manual_seed(42)
    mod = vq.HybridVQ(64, [8, 5, 5, 5], vq_codebook_size=512).to(DEV)
    sd = {k: N_(v) for k, v in mod.state_dict().items()}
    stages = [VQState(sd[f"vq.layers.{i}.embedding.weight"].copy(), sd[f"vq.layers.{i}.ema_cluster_size"].copy(),
                      sd[f"vq.layers.{i}.ema_w"].copy(), 0.25, True, 0.99) for i in range(4)]
    rng = np.random.default_rng(1236)
    mod.train()
    for step in range(3):
        z_np = np.ascontiguousarray(rng.standard_normal((512, 1, 64)).astype(np.float32).transpose(0, 2, 1))
        z = T(np.ascontiguousarray(z_np.transpose(0, 2, 1))).permute(0, 2, 1).requires_grad_(True)
        loss, q, met = mod(z)
        idx = N_(mod.vq.last_indices).astype(np.int64)
        z_e = N_(mod.fsq.last_z_e)      # the engine's own post-project_in tensor (fused kernel)
        np.testing.assert_array_equal(N_(mod.fsq.last_indices), fsq_quantize(z_e, [8, 5, 5, 5])["indices"])
        ref = hybrid_forward(z_np, [8, 5, 5, 5], sd["fsq.project_in.weight"], sd["fsq.project_in.bias"],
                             sd["fsq.project_out.weight"], sd["fsq.project_out.bias"], stages, True,
                             force_indices=list(idx), z_e=z_e)
        # the engine's stage-0 choice must be the oracle's up to benign flips (same residual up to conv noise)
        assert_close(N_(q), ref["quantized"], 5e-5, "quantized")
        assert_close(N_(loss), ref["loss"], 5e-5, "loss")
        assert_close(N_(met["rvq_ppl"]), ref["rvq_ppl"], TOL, "rvq_ppl")
        for i in range(4):
            assert_close(N_(mod.vq.layers[i].embedding.weight), stages[i].embedding, 5e-5, f"E{i}", rows=True)
        (loss + (q * T(rng.standard_normal((512, 64, 1)).astype(np.float32))).sum()).backward()
        assert z.grad is not None and torch.isfinite(z.grad).all()
        mod.zero_grad()


def test_cfg3_rvq_chunk():
    """cfg3 per-GPU chunk the oracle can hold: RVQ S=4, K=1024, D=64 on [2048,64,10]."""
    vq = _mods()
    S, K, D, B, Tt = 4, 1024, 64, 2048, 10
    mod = vq.ResidualVQ(S, K, D, use_ema=True).to(DEV)
    stages = [_fresh_state(K, D, True, 100 + s, "small") for s in range(S)]
    for l, st in zip(mod.layers, stages):
        _load_vq(l, st)
    rng = np.random.default_rng(1237)
    for step in range(2):
        z = (0.5 * rng.standard_normal((B, D, Tt))).astype(np.float32)
        g = rng.standard_normal((B, D, Tt)).astype(np.float32)
        flips, _ = _rvq_step(mod, stages, z, g, 1.0)
        print("cfg3 chunk step", step, "benign flips", flips)


@pytest.mark.parametrize("B,Tt,S,K,training", [(7, 1, 4, 512, True), (100, 3, 4, 512, True), (512, 1, 4, 512, True),
                                                 (1000, 1, 3, 700, True), (1024, 1, 2, 1024, True),
                                                 (1025, 1, 4, 512, True), (300, 10, 4, 512, True),
                                                 (512, 1, 4, 512, False), (97, 2, 8, 333, True), (210, 10, 3, 333, True)])
def test_single_launch_rvq_shapes(B, Tt, S, K, training):
    """K4 (rvq_small.cu) on both of its kernels: the whole-GPU cooperative variant (N <= 1024, S*K <= 3072) and the
    one-cluster variant (larger N / more codes), two steps each against the oracle, plus the multi-kernel path on the
    same state (indices must be identical: all three are exact fp32 with the same summation order)."""
    vq = _mods()
    D = 64
    mod = vq.ResidualVQ(S, K, D, use_ema=True).to(DEV)
    twin = vq.ResidualVQ(S, K, D, use_ema=True).to(DEV)
    for l in twin.layers:
        l.assign_algo = vq._lib.ASSIGN_SIMT          # forces the multi-kernel path
    stages = [_fresh_state(K, D, True, 300 + s, "small") for s in range(S)]
    for l, l2, st in zip(mod.layers, twin.layers, stages):
        _load_vq(l, st)
        _load_vq(l2, st)
    rng = np.random.default_rng(77 + B)
    for step in range(2):
        z = (0.5 * rng.standard_normal((B, D, Tt))).astype(np.float32)
        g = rng.standard_normal((B, D, Tt)).astype(np.float32)
        twin.train(training)
        with torch.no_grad():
            twin(T(z))
        _rvq_step(mod, stages, z, g, 1.0, training)
        a, b = N_(mod.last_indices), N_(twin.last_indices)
        if step == 0:       # identical codebooks going in: stage 0 must agree exactly (same arithmetic, same order)
            np.testing.assert_array_equal(a[0], b[0])
        # later stages / steps see codebooks that differ by float-atomic summation order (~1e-7): near-ties may flip,
        # and a flipped row moves one vector between two codes -- the oracle check above is the parity statement
        assert (a != b).mean() < 2e-3, f"single-launch vs multi-kernel indices differ on {(a != b).mean():.2%} of rows"


def _proj_oracle_grads(z, z_q, g, g_ze, w_in, w_out):
    """Autograd of out = W_out z_q + b_out, z_e = W_in z + b_in with straight-through z_q (float64)."""
    z, z_q, g, g_ze = (np.asarray(a, np.float64) for a in (z, z_q, g, g_ze))
    return {"g_z": np.einsum("jc,bjt->bct", np.asarray(w_in, np.float64)[:, :, 0], g_ze),
            "w_in": np.einsum("bjt,bct->jc", g_ze, z)[:, :, None], "b_in": g_ze.sum((0, 2)),
            "w_out": np.einsum("bct,bjt->cj", g, z_q)[:, :, None], "b_out": g.sum((0, 2))}


@pytest.mark.parametrize("levels,B,Tt", [([8, 5, 5, 5], 4099, 10), ([8, 5, 5, 5], 512, 1), ([7, 5, 5, 5, 5, 3], 333, 7),
                                         ([3] * 10, 1000, 10), ([2] * 16, 257, 3)])
def test_fsq_fused_projections(levels, B, Tt):
    """FSQ module with both 1x1 projections inside the kernel vs the oracle: indices bit-exact
    from the engine's own z_e, z_e / output / gradients within tolerance, and the unfused path agrees."""
    from oracle import fsq_forward, conv1x1
    vq = _mods()
    D, d = 64, len(levels)
    torch.manual_seed(5 + d)
    mod = vq.FSQ(levels, D, D).to(DEV)
    rng = np.random.default_rng(B + d)
    z_np = (2.0 * rng.standard_normal((B, D, Tt))).astype(np.float32)
    g_np = rng.standard_normal((B, D, Tt)).astype(np.float32)
    sd = {k: N_(v) for k, v in mod.state_dict().items()}
    z = T(z_np).requires_grad_(True)
    loss, out, met = mod(z)
    assert mod.last_z_e is not None and mod.last_z_e.shape == (B, d, Tt)
    (out * T(g_np)).sum().backward()
    z_e = N_(mod.last_z_e)
    assert_close(z_e, conv1x1(z_np, sd["project_in.weight"], sd["project_in.bias"]), TOL, "z_e")
    ref = fsq_forward(z_np, levels, sd["project_in.weight"], sd["project_in.bias"], sd["project_out.weight"],
                      sd["project_out.bias"], z_e=z_e)
    np.testing.assert_array_equal(N_(mod.last_indices), ref["indices"])
    assert_close(N_(out), ref["quantized"], TOL, "out")
    assert float(loss) == 0.0
    assert float(met["perplexity"]) == float(len(np.unique(ref["indices"])))
    g_ze = np.einsum("cj,bct->bjt", sd["project_out.weight"][:, :, 0].astype(np.float64), g_np.astype(np.float64))
    og = _proj_oracle_grads(z_np, ref["z_hard"], g_np, g_ze, sd["project_in.weight"], sd["project_out.weight"])
    assert_close(N_(z.grad), og["g_z"], 1e-4, "g_z")
    assert_close(N_(mod.project_in.weight.grad), og["w_in"], 1e-4, "g project_in.weight")
    assert_close(N_(mod.project_in.bias.grad), og["b_in"], 1e-4, "g project_in.bias")
    assert_close(N_(mod.project_out.weight.grad), og["w_out"], 1e-4, "g project_out.weight")
    assert_close(N_(mod.project_out.bias.grad), og["b_out"], 1e-4, "g project_out.bias")
    # stock conv1d + elementwise kernel path: same numbers up to conv rounding
    mod.fuse_projections = False
    _, out_u, _ = mod(T(z_np))
    same = N_(mod.last_indices) == ref["indices"]
    assert same.mean() > 0.999
    assert_close(N_(out_u)[same.nonzero()[0]], N_(out)[same.nonzero()[0]], 1e-4, "fused vs unfused")


@pytest.mark.parametrize("cd,B,Tt", [(10, 4099, 10), (10, 512, 1), (4, 300, 7), (16, 1000, 3)])
def test_lfq_fused_projections(cd, B, Tt):
    from oracle import lfq_quantize, lfq_backward_ze, conv1x1
    vq = _mods()
    D = 64
    torch.manual_seed(11 + cd)
    mod = vq.LFQ(D, codebook_dim=cd).to(DEV)
    rng = np.random.default_rng(B + cd)
    z_np = (2.0 * rng.standard_normal((B, D, Tt))).astype(np.float32)
    g_np = rng.standard_normal((B, D, Tt)).astype(np.float32)
    sd = {k: N_(v) for k, v in mod.state_dict().items()}
    z = T(z_np).requires_grad_(True)
    loss, out, met = mod(z)
    ((out * T(g_np)).sum() + 3.0 * loss).backward()
    z_e = N_(mod.last_z_e)
    assert_close(z_e, conv1x1(z_np, sd["project_in.weight"], sd["project_in.bias"]), TOL, "z_e")
    ref = lfq_quantize(z_e, 0.1)
    np.testing.assert_array_equal(N_(mod.last_indices), ref["indices"])
    assert_close(N_(out), conv1x1(ref["z_q"], sd["project_out.weight"], sd["project_out.bias"]), TOL, "out")
    assert abs(float(loss) - float(ref["loss"])) <= 1e-5 * max(1.0, abs(float(ref["loss"])))
    assert float(met["perplexity"]) == float(len(np.unique(ref["indices"])))
    g_zq = np.einsum("cj,bct->bjt", sd["project_out.weight"][:, :, 0].astype(np.float64), g_np.astype(np.float64))
    g_ze = lfq_backward_ze(z_e, g_zq, 3.0, 0.1)
    og = _proj_oracle_grads(z_np, ref["z_q"], g_np, g_ze, sd["project_in.weight"], sd["project_out.weight"])
    assert_close(N_(z.grad), og["g_z"], 1e-4, "g_z")
    assert_close(N_(mod.project_in.weight.grad), og["w_in"], 1e-4, "g project_in.weight")
    assert_close(N_(mod.project_in.bias.grad), og["b_in"], 1e-4, "g project_in.bias")
    assert_close(N_(mod.project_out.weight.grad), og["w_out"], 1e-4, "g project_out.weight")
    assert_close(N_(mod.project_out.bias.grad), og["b_out"], 1e-4, "g project_out.bias")


@pytest.mark.parametrize("B,Tt,K,contig", [(4000, 10, 1024, True), (2311, 7, 300, True), (5003, 1, 512, True),
                                            (3000, 10, 256, False), (100, 10, 64, True)])
def test_assign_residual_entry_point(B, Tt, K, contig):
    """vqb200_vq_assign_residual (residual update fused into the tensor-core assignment for contiguous D=64 input)
    against the two stand-alone C-ABI calls and against the oracle's residual arithmetic (models/vqvae.py:94-98):
    the residual must be bit-identical, the indices equal up to provably tied distances."""
    import ctypes
    vq = _mods()
    from vqb200 import _lib
    from vqb200._lib import ptr, stream_ptr, check
    lib = _lib.load()
    D = 64
    rng = np.random.default_rng(B + K)
    W0 = (0.3 * rng.standard_normal((K, D))).astype(np.float32)
    W1 = (0.1 * rng.standard_normal((K, D))).astype(np.float32)
    z_np = (0.5 * rng.standard_normal((B, D, Tt))).astype(np.float32)
    z = T(z_np) if contig else T(np.ascontiguousarray(z_np.transpose(0, 2, 1))).transpose(1, 2)   # channel-last view
    w0, w1 = T(W0), T(W1)
    st0, st1 = vq.QuantizerState(K, D, torch.device(DEV)), vq.QuantizerState(K, D, torch.device(DEV))
    idx0 = vq.vq_assign(z, w0, st0, _lib.ASSIGN_SIMT)
    st1.refresh(w1)
    ws = st1.assign_workspace(B * Tt)
    r_ref, r_fused = torch.empty((B, D, Tt), device=DEV), torch.empty((B, D, Tt), device=DEV)
    idx_fused = torch.empty((B, Tt), dtype=torch.int32, device=DEV)
    sse = torch.zeros(1, dtype=torch.float64, device=DEV)
    sB, sC, sT = z.stride()
    s = stream_ptr(torch.device(DEV))
    check(lib.vqb200_vq_gather_st(ptr(z), B, D, Tt, sB, sC, sT, ptr(w0), ptr(idx0), K, None, ptr(r_ref), None, 0,
                                  ptr(sse), s), "gather_st")
    idx_ref = vq.vq_assign(r_ref, w1, st1, _lib.ASSIGN_SIMT)
    check(lib.vqb200_vq_assign_residual(ptr(z), B, D, Tt, sB, sC, sT, ptr(w0), ptr(idx0), K, ptr(r_fused), ptr(w1),
                                        ptr(st1.ee), ptr(st1.image), ptr(st1.info), K, ptr(idx_fused), ptr(ws),
                                        ctypes.c_size_t(ws.numel()), _lib.ASSIGN_AUTO, s), "assign_residual")
    torch.cuda.synchronize()
    assert torch.equal(r_fused, r_ref)
    # oracle arithmetic: st = x + (q - x); r = x - st, all fp32
    q = W0[N_(idx0).astype(np.int64)].transpose(0, 2, 1)
    r_or = z_np - (z_np + (q - z_np))
    assert np.array_equal(N_(r_fused), r_or)
    d = vq_distances(np.ascontiguousarray(r_or.transpose(0, 2, 1)).reshape(-1, D), W1)
    flips, bad, rows = check_indices(N_(idx_fused).reshape(-1).astype(np.int64), N_(idx_ref).reshape(-1).astype(np.int64), d)
    assert bad == 0, f"{bad} non-benign flips of {flips} at rows {rows[:8]}"


@pytest.mark.parametrize("K,D", [(512, 64), (2048, 128), (4096, 64), (1000, 24), (16384, 256), (1024, 512)])
def test_cfg5_points(K, D):
    """cfg5 sweep points on a 16 384-vector chunk (the oracle materialises N x K)."""
    vq = _mods()
    N = 16384 if K * D <= 2048 * 128 else 4096
    st = _fresh_state(K, D, True, 7, "normal")
    mod = vq.VectorQuantizer(K, D, use_ema=True).to(DEV)
    _load_vq(mod, st)
    rng = np.random.default_rng(1239)
    z = rng.standard_normal((N, D, 1)).astype(np.float32)
    _vq_step(mod, st, z, rng.standard_normal((N, D, 1)).astype(np.float32), 1.0, True)


# ----------------------------------------------------------------------------------------------
# layouts, ragged / empty inputs, NaN rule, error behaviour
# ----------------------------------------------------------------------------------------------
def test_strided_views_agree():
    vq = _mods()
    K, D, B, Tt = 256, 32, 37, 9
    torch.manual_seed(0)
    mod = vq.VectorQuantizer(K, D).to(DEV).eval()
    with torch.no_grad():
        mod.embedding.weight.normal_()
    base = torch.randn(B, D, Tt, device=DEV)
    _, q0, _ = mod(base)
    i0 = mod.last_indices.clone()
    big = torch.zeros(B, D + 3, Tt + 5, device=DEV)
    big[:, :D, :Tt] = base
    views = {
        "btc_permuted": base.permute(0, 2, 1).contiguous().permute(0, 2, 1),     # [B,T,C] memory, strides (T*C, 1, C)
        "padded": big[:, :D, :Tt],                                               # arbitrary strides
        "clone": base.clone(),
    }
    for name, v in views.items():
        assert torch.equal(v, base)
        _, q, _ = mod(v)
        assert torch.equal(mod.last_indices, i0), name
        assert torch.equal(q, q0), name
        assert q.is_contiguous()


def test_rvq_strided_views_agree():
    """ResidualVQ on arbitrary views (generic strided kernels + scratch-based output chain) must equal the
    contiguous fast path bit for bit, in eval and in EMA training mode, forward and backward."""
    import copy
    vq = _mods()
    S, K, D, B, Tt = 3, 128, 32, 41, 6
    torch.manual_seed(0)
    ref_mod = vq.ResidualVQ(S, K, D, use_ema=True).to(DEV)
    with torch.no_grad():
        for l in ref_mod.layers:
            l.embedding.weight.normal_(0, 0.5); l.ema_w.copy_(l.embedding.weight); l.ema_cluster_size.fill_(1.0)
    base = torch.randn(B, D, Tt, device=DEV)
    g = torch.randn(B, D, Tt, device=DEV)
    big = torch.zeros(B, D + 5, Tt + 3, device=DEV)
    big[:, :D, :Tt] = base
    views = {"contiguous": base.clone(),
             "btc_permuted": base.permute(0, 2, 1).contiguous().permute(0, 2, 1),
             "padded": big[:, :D, :Tt]}
    results = {}
    for name, v in views.items():
        for train in (False, True):
            m = copy.deepcopy(ref_mod).train(train)
            x = v.detach().clone().as_strided(v.shape, v.stride()) if name == "contiguous" else v.detach()
            x = x.requires_grad_(True)
            loss, q, met = m(x)
            torch.autograd.backward([q, loss], [g, torch.ones((), device=DEV)])
            results[(name, train)] = (q.detach(), loss.detach(), m.last_indices.clone(), x.grad.detach().clone(),
                                      m.layers[-1].embedding.weight.detach().clone())
    for train in (False, True):
        q0, l0, i0, g0, e0 = results[("contiguous", train)]
        for name in ("btc_permuted", "padded"):
            q, l, i, gz, e = results[(name, train)]
            assert torch.equal(i, i0), (name, train)
            if train:       # float-atomic order of the EMA sums differs between launches: last-bit differences in E
                assert torch.allclose(q, q0, rtol=1e-5, atol=1e-6), (name, train)
            else:
                assert torch.equal(q, q0), (name, train)
            assert torch.allclose(l, l0, rtol=1e-6), (name, train)
            assert torch.allclose(gz, g0, rtol=1e-6, atol=1e-9), (name, train)
            assert torch.allclose(e, e0, rtol=1e-5, atol=1e-7), (name, train)      # atomics order only


@pytest.mark.parametrize("B,Tt", [(0, 10), (1, 1), (1, 10), (7, 3), (129, 1)])
def test_ragged_and_empty(B, Tt):
    vq = _mods()
    K, D = 100, 20
    st = _fresh_state(K, D, True, 3, "normal")
    mod = vq.VectorQuantizer(K, D, use_ema=True).to(DEV)
    _load_vq(mod, st)
    rng = np.random.default_rng(B * 31 + Tt)
    z = rng.standard_normal((B, D, Tt)).astype(np.float32)
    if B == 0:
        mod.eval()
        loss, q, met = mod(T(z))
        assert q.shape == (0, D, Tt)
        return
    _vq_step(mod, st, z, rng.standard_normal((B, D, Tt)).astype(np.float32), 1.0, True)


def test_nan_and_tie_rules():
    """torch.argmin: ties -> lowest index; a NaN distance wins."""
    vq = _mods()
    K, D = 8, 4
    mod = vq.VectorQuantizer(K, D).to(DEV).eval()
    E = np.zeros((K, D), np.float32)
    E[2] = E[5] = [1, 0, 0, 0]         # duplicate codes: exact tie -> 2
    E[1] = [5, 5, 5, 5]
    with torch.no_grad():
        mod.embedding.weight.copy_(T(E))
    z = np.zeros((3, D, 1), np.float32)
    z[0, :, 0] = [1, 0, 0, 0]
    z[1, :, 0] = [0, 0, 0, 0]          # ties between all zero codes -> 0
    z[2, :, 0] = [np.nan, 0, 0, 0]     # every distance NaN -> 0
    mod(T(z))
    ref = vq_forward(z, VQState(E.copy()), False)
    np.testing.assert_array_equal(N_(mod.last_indices), ref["indices"])
    E2 = E.copy(); E2[6, 0] = np.nan    # one NaN code: it wins for every finite row
    with torch.no_grad():
        mod.embedding.weight.copy_(T(E2))
    mod(T(z))
    ref = vq_forward(z, VQState(E2.copy()), False)
    np.testing.assert_array_equal(N_(mod.last_indices), ref["indices"])


def test_short_work_list_split_merge_keeps_nan_and_tie_rules():
    """Mid-size N: the unproven rows of the tensor-core path are re-done by the exact kernel with the codebook sweep
    split over several CTAs and merged through 64-bit atomicMin keys; the merge must keep torch.argmin's rules
    (first minimum, a NaN distance wins) -- compare with the unsplit exact kernel on crafted rows."""
    vq = _mods()
    from vqb200 import _lib
    torch.manual_seed(9)
    K, D, B, Tt = 1024, 64, 500, 10
    w = 0.3 * torch.randn(K, D, device=DEV)
    w[700] = w[3]; w[901] = w[3]; w[512] = w[130]          # duplicate codes: exact ties across different splits
    z = 0.5 * torch.randn(B, D, Tt, device=DEV)
    zt = z.permute(0, 2, 1)                                # [B,T,D] view of the same storage
    zt[0, 0] = w[3]; zt[1, 1] = w[901]; zt[2, 2] = w[512]  # rows that ARE duplicated codes
    zt[3, 3, 5] = float("nan"); zt[4, 4] = float("inf"); zt[5, 5] = 0.0
    zt[6, 6] = 0.5 * (w[10] + w[20])                       # near-tie between two codes
    st = vq.QuantizerState(K, D, torch.device(DEV))
    exact = vq.vq_assign(z, w, st, _lib.ASSIGN_SIMT)
    auto = vq.vq_assign(z, w, st, _lib.ASSIGN_AUTO)
    assert torch.equal(exact, auto)
    assert int(auto[0, 0]) == 3 and int(auto[1, 1]) == 3 and int(auto[2, 2]) == 130
    # every row unproven (all-equal codebook): the whole batch goes through the split merge
    w2 = w[:1].expand(K, D).contiguous()
    st2 = vq.QuantizerState(K, D, torch.device(DEV))
    assert int(vq.vq_assign(z, w2, st2, _lib.ASSIGN_AUTO).abs().max()) == 0


def test_called_twice_per_step_like_student_mode():
    """The shared quantizer is invoked twice per forward (robot + human branch) and the EMA state
    mutates between the calls; backward of the first call must use ITS codebook."""
    vq = _mods()
    K, D, B, Tt = 64, 16, 32, 1
    st = _fresh_state(K, D, True, 9, "normal")
    mod = vq.VectorQuantizer(K, D, use_ema=True).to(DEV).train()
    _load_vq(mod, st)
    rng = np.random.default_rng(5)
    z1 = rng.standard_normal((B, D, Tt)).astype(np.float32)
    z2 = rng.standard_normal((B, D, Tt)).astype(np.float32)
    a = T(z1).requires_grad_(True); b = T(z2).requires_grad_(True)
    l1, q1, _ = mod(a); i1 = N_(mod.last_indices).astype(np.int64)
    l2, q2, _ = mod(b); i2 = N_(mod.last_indices).astype(np.int64)
    (l1 + l2 + q1.sum() + q2.sum()).backward()
    r1 = vq_forward(z1, st, True, force_indices=i1)
    g1, _ = vq_backward(r1, st, np.ones_like(z1), 1.0)
    r2 = vq_forward(z2, st, True, force_indices=i2)
    g2, _ = vq_backward(r2, st, np.ones_like(z2), 1.0)
    assert_close(N_(a.grad), g1, TOL, "grad of first call")
    assert_close(N_(b.grad), g2, TOL, "grad of second call")
    assert_close(N_(mod.ema_cluster_size).sum(), st.ema_cluster_size.sum(), TOL, "cs")


def test_errors_are_loud():
    vq = _mods()
    mod = vq.VectorQuantizer(16, 8).to(DEV)
    with pytest.raises(RuntimeError):
        mod(torch.randn(2, 8, 3))                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        mod(torch.randn(2, 7, 3, device=DEV))           # channel mismatch
    with pytest.raises(RuntimeError):
        mod(torch.randn(2, 8, device=DEV))              # not [B,C,T]


# ----------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full sizes
# ----------------------------------------------------------------------------------------------
def test_full_size_properties_rvq():
    """cfg3-sized shard (N = 2.5 M here to bound test time; the bench runs 10 M): conservation laws of
    the EMA update, idempotence of quantisation, residual energy decreasing stage by stage."""
    vq = _mods()
    S, K, D, B, Tt = 4, 1024, 64, 250_000, 10
    torch.manual_seed(1)
    mod = vq.ResidualVQ(S, K, D, use_ema=True).to(DEV).train()
    with torch.no_grad():
        for l in mod.layers:
            l.embedding.weight.normal_(0, 0.3)
            l.ema_cluster_size.fill_(1.0)
            l.ema_w.copy_(l.embedding.weight)
    z = 0.5 * torch.randn(B, D, Tt, device=DEV)
    cs_before = [l.ema_cluster_size.sum().item() for l in mod.layers]
    loss, q, met = mod(z)
    N = B * Tt
    idx = mod.last_indices
    assert idx.shape == (S, B, Tt) and int(idx.min()) >= 0 and int(idx.max()) < K
    for s, l in enumerate(mod.layers):     # sum(cs') = decay*sum(cs) + (1-decay)*N
        expect = 0.99 * cs_before[s] + 0.01 * N
        assert abs(l.ema_cluster_size.sum().item() - expect) / expect < 1e-5
        hist = torch.bincount(idx[s].reshape(-1).long(), minlength=K).float()
        assert int(hist.sum()) == N
    # quantized == sum of the (updated) codewords the indices point to, up to the fp32 ST chain
    recon = sum(l.embedding.weight[idx[s].long()] for s, l in enumerate(mod.layers))    # [B,T,D]
    assert_close(N_(q[:4096]), N_(recon.permute(0, 2, 1)[:4096]), 1e-5, "out == sum of codewords")
    # stage-0 assignment is a true nearest neighbour under the *pre-update* codebook: spot-check rows
    mod.eval()
    loss2, q2, _ = mod(z)
    res = z - q2
    assert float((res ** 2).mean()) < float((z ** 2).mean())
    # idempotence (eval): quantising a codeword of stage 0 returns that codeword, loss 0
    l0 = mod.layers[0].eval()
    cw = l0.embedding.weight[:K].t().unsqueeze(0).contiguous()                          # [1, D, K]
    _, qq, _ = l0(cw)
    d_self = ((qq - cw) ** 2).sum(1)
    assert float(d_self.max()) <= 1e-10 or torch.equal(l0.last_indices.reshape(-1).cpu(), torch.arange(K, dtype=torch.int32))


def test_full_size_fsq_lfq_sweep():
    """cfg4 end points: B = 4096 and 1 048 576 windows of [64,10]; kernel indices re-derived with torch
    integer ops must match bit for bit, unique counts must equal torch.unique."""
    vq = _mods()
    for B in (4096, 1_048_576):
        z_e = 2.0 * torch.randn(B, 4, 10, device=DEV)
        basis = torch.tensor([1, 8, 40, 200], dtype=torch.int32, device=DEV)
        z_hard, idx, m2 = vq.fsq_round(z_e, basis, 1000)
        ref_idx = (torch.round(z_e).permute(0, 2, 1) * basis).sum(-1).long()
        assert torch.equal(idx, ref_idx)
        assert torch.equal(z_hard, torch.round(z_e))
        assert int(m2[0]) == torch.unique(ref_idx).numel()
        z_l = torch.randn(B, 10, 10, device=DEV)
        z_q, loss, idx, m3 = vq.lfq_sign(z_l, 0.1)
        ref = ((z_l > 0).long().permute(0, 2, 1) * (2 ** torch.arange(10, device=DEV))).sum(-1)
        assert torch.equal(idx, ref)
        assert int(m3[1]) == torch.unique(ref).numel()
        assert torch.equal(z_q, torch.where(z_l > 0, 1.0, -1.0))
