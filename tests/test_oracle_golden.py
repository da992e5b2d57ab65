"""Pin the numpy oracle against fixtures recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import (VQState, vq_forward, vq_backward, rvq_forward, rvq_backward, fsq_forward,
                    fsq_quantize, lfq_forward, lfq_quantize, lfq_backward_ze, hybrid_forward,
                    check_indices, assert_close, vq_distances)
from _golden import load, vq_state_from

TOL = 1e-5


def _check_vq_step(g, p, st, training, bwd=True, g_loss=1.7):
    z = g[p + "z"]
    has_d = (p + "distances") in g
    ref_idx = g[p + "indices"]
    # oracle's own argmin vs the reference's
    pre = st.copy()
    r_free = vq_forward(z, pre, training, keep_distances=True)
    dist_ref = g[p + "distances"] if has_d else r_free["distances"]
    if has_d:
        assert_close(r_free["distances"], dist_ref, 2e-6, p + "distances")
    flips, bad, _ = check_indices(r_free["indices"], ref_idx, dist_ref)
    assert bad == 0, f"{p}: {bad} non-benign index flips of {flips}"
    # teacher-forced on the reference's indices for the value comparison
    r = vq_forward(z, st, training, force_indices=ref_idx)
    assert_close(r["quantized"], g[p + "quantized"], TOL, p + "quantized")
    assert_close(r["loss"], g[p + "loss"], TOL, p + "loss")
    assert_close(r["perplexity"], g[p + "perplexity"], TOL, p + "perplexity")
    assert_close(r["dcr"], g[p + "dcr"], 1e-6, p + "dcr")
    assert_close(st.embedding, g[p + "after.embedding"], TOL, p + "embedding", rows=True)
    if st.use_ema:
        assert_close(st.ema_cluster_size, g[p + "after.ema_cluster_size"], TOL, p + "ema_cluster_size")
        assert_close(st.ema_w, g[p + "after.ema_w"], TOL, p + "ema_w", rows=True)
    if bwd and (p + "grad_z") in g:
        gz, gE = vq_backward(r, st, g[p + "g"], g_loss)
        assert_close(gz, g[p + "grad_z"], TOL, p + "grad_z")
        if (p + "grad_embedding") in g:
            assert_close(gE, g[p + "grad_embedding"], TOL, p + "grad_embedding")
        else:
            assert gE is None
    return flips


@pytest.mark.parametrize("name", ["vq_std_small", "vq_ema_fresh", "vq_ema_k1024_perm", "vq_std_ragged"])
def test_vq_against_reference(name):
    g = load(name)
    use_ema = bool(g["use_ema"])
    st = vq_state_from(g, "init.", use_ema)
    total_flips = 0
    for s in range(int(g["steps"])):
        total_flips += _check_vq_step(g, f"s{s}.", st, True)
    if "eval.z" in g:
        before = st.copy()
        _check_vq_step(g, "eval.", st, False, bwd=False)
        np.testing.assert_array_equal(before.embedding, st.embedding)   # eval never mutates
    print(name, "benign flips:", total_flips)


@pytest.mark.parametrize("name", ["rvq_ema", "rvq_std"])
def test_rvq_against_reference(name):
    g = load(name)
    S, use_ema = int(g["S"]), bool(g["use_ema"])
    stages = [vq_state_from(g, f"init.layers.{i}.", use_ema) for i in range(S)]
    g_loss = 0.9
    for s in range(int(g["steps"])):
        p = f"s{s}."
        ref_idx = g[p + "indices"]
        # free-running oracle: stage-by-stage flip check against the reference's distances
        free = rvq_forward(g[p + "z"], [x.copy() for x in stages], True)
        for i in range(S):
            flips, bad, _ = check_indices(free["indices"][i], ref_idx[i], g[p + "distances"][i])
            assert bad == 0
            if flips:      # later stages see a different residual on flipped rows: stop there
                break
        r = rvq_forward(g[p + "z"], stages, True, force_indices=list(ref_idx))
        assert_close(r["quantized"], g[p + "quantized"], TOL, p + "quantized")
        assert_close(r["loss"], g[p + "loss"], TOL, p + "loss")
        assert_close(r["perplexity"], g[p + "perplexity"], TOL, p + "perplexity")
        assert_close(r["dcr"], g[p + "dcr"], 1e-6, p + "dcr")
        gz, gEs = rvq_backward(r, stages, g[p + "g"], g_loss)
        assert_close(gz, g[p + "grad_z"], TOL, p + "grad_z")
        for i in range(S):
            assert_close(stages[i].embedding, g[p + f"after.layers.{i}.embedding"], TOL, "E", rows=True)
            if use_ema:
                assert_close(stages[i].ema_w, g[p + f"after.layers.{i}.ema_w"], TOL, "ema_w", rows=True)
                assert_close(stages[i].ema_cluster_size, g[p + f"after.layers.{i}.ema_cluster_size"], TOL, "cs")
                assert gEs[i] is None
            else:
                assert_close(gEs[i], g[p + f"grad_embedding.{i}"], TOL, f"grad_embedding.{i}")


def test_rvq_input_gradient_closed_form():
    """only stage 0's commitment term reaches the input."""
    g = load("rvq_ema")
    stages = [vq_state_from(g, f"init.layers.{i}.", True) for i in range(int(g["S"]))]
    r = rvq_forward(g["s0.z"], stages, True, force_indices=list(g["s0.indices"]))
    c0 = r["stage"][0]
    coef = np.float32(0.9 * 0.25 * 2.0 / c0["x"].size)
    expect = g["s0.g"] + (coef * (c0["x"] - c0["q"])).transpose(0, 2, 1)
    assert_close(expect, g["s0.grad_z"], TOL, "closed form")


@pytest.mark.parametrize("name", ["fsq_module", "fsq_module_x30", "fsq_crafted"])
def test_fsq_against_reference(name):
    g = load(name)
    levels = [int(v) for v in g["levels"]]
    # elementwise stage on the reference's own post-projection tensor: bit exact
    q = fsq_quantize(g["z_e"], levels)
    np.testing.assert_array_equal(q["indices"], g["indices"])
    np.testing.assert_array_equal(q["z_hard"], g["z_hard"])
    assert float(q["perplexity"]) == float(g["perplexity"])
    assert_close(q["dcr"], g["dcr"], 1e-6, "dcr")
    # whole module (projection results may differ in the last ulp across BLAS back ends)
    r = fsq_forward(g["z"], levels, g["state.project_in.weight"], g["state.project_in.bias"],
                    g["state.project_out.weight"], g["state.project_out.bias"])
    assert_close(r["z_e"], g["z_e"], 1e-5, "z_e")
    assert float(r["loss"]) == float(g["loss"]) == 0.0
    np.testing.assert_array_equal(g["state._basis"], np.array([1, 8, 40, 200], np.int32))
    if name == "fsq_module_x30":
        assert g["indices"].min() < 0 or g["indices"].max() >= 1000   # unbounded by design


@pytest.mark.parametrize("name", ["lfq_module", "lfq_crafted"])
def test_lfq_against_reference(name):
    g = load(name)
    w = float(g["entropy_loss_weight"])
    q = lfq_quantize(g["z_e"], w)
    np.testing.assert_array_equal(q["indices"], g["indices"])
    np.testing.assert_array_equal(q["z_q"], g["z_q"])
    assert_close(q["loss"], g["loss"], TOL, "loss")
    assert float(q["perplexity"]) == float(g["perplexity"])
    assert_close(q["dcr"], g["dcr"], 1e-6, "dcr")
    gz = lfq_backward_ze(g["z_e"], g["g_zq"], float(g["g_loss"]), w)
    assert_close(gz, g["grad_z_e"], TOL, "grad_z_e")
    r = lfq_forward(g["z"], g["state.project_in.weight"], g["state.project_in.bias"],
                    g["state.project_out.weight"], g["state.project_out.bias"], w)
    assert_close(r["quantized"], g["quantized"], 1e-5, "quantized")


@pytest.mark.parametrize("name", ["hybrid_perm", "hybrid_t10"])
def test_hybrid_against_reference(name):
    g = load(name)
    levels = [int(v) for v in g["levels"]]
    S = int(g["S"])
    stages = [VQState(g[f"init.vq.layers.{i}.embedding.weight"].copy(),
                      g[f"init.vq.layers.{i}.ema_cluster_size"].copy(),
                      g[f"init.vq.layers.{i}.ema_w"].copy(), 0.25, True, 0.99) for i in range(S)]
    for s in range(int(g["steps"])):
        p = f"s{s}."
        r = hybrid_forward(g[p + "z"], levels, g["init.fsq.project_in.weight"], g["init.fsq.project_in.bias"],
                           g["init.fsq.project_out.weight"], g["init.fsq.project_out.bias"], stages, True,
                           force_indices=list(g[p + "indices"]))
        assert_close(r["quantized"], g[p + "quantized"], TOL, p + "quantized")
        assert_close(r["loss"], g[p + "loss"], TOL, p + "loss")
        assert float(r["perplexity"]) == float(g[p + "perplexity"])
        assert_close(r["dcr"], g[p + "dcr"], 1e-6, p + "dcr")
        assert_close(r["rvq_ppl"], g[p + "rvq_ppl"], TOL, p + "rvq_ppl")
        for i in range(S):
            assert_close(stages[i].embedding, g[p + f"after.vq.layers.{i}.embedding.weight"], TOL, "E", rows=True)
            assert_close(stages[i].ema_w, g[p + f"after.vq.layers.{i}.ema_w"], TOL, "ema_w", rows=True)
        # the free-running oracle reproduces the reference's stage-0 indices up to benign flips
        dist0 = vq_distances(np.ascontiguousarray(r["residual"].transpose(0, 2, 1)).reshape(-1, int(g["D"])),
                             g["init.vq.layers.0.embedding.weight"] if s == 0 else g[f"s{s-1}.after.vq.layers.0.embedding.weight"])
        _, bad, _ = check_indices(np.argmin(dist0, 1), g[p + "indices"][0], g[p + "distances"][0])
        assert bad == 0


def test_known_degenerate_regime_is_covered():
    """Fresh-init EMA step produces codebook rows ~1e5 in magnitude."""
    g = load("vq_ema_fresh")
    assert np.abs(g["s0.after.embedding"]).max() > 1e3
