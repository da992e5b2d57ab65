#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the reference tree is not present on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz + *.json

It imports `/root/reference/models/vqvae.py` untouched (torch CPU, fp32), drives each quantizer
on seeded inputs, and records inputs, initial state, outputs, post-step buffers and autograd
gradients.  Indices (which the reference never returns, SURVEY.md "three things" #1) are
captured with forward-pre-hooks that evaluate the reference's own distance expression
(models/vqvae.py:34-38) on the layer input with the pre-update codebook.
"""
import json
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("VQ_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)

import numpy as np
import torch

from models.vqvae import (VectorQuantizer, ResidualVQ, FSQ, LFQ, HybridVQ,  # noqa: E402
                          DualMotionVQVAE)

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def npy(t):
    return t.detach().cpu().numpy().copy()


class IndexTap:
    """Records (indices, distances) of every VectorQuantizer call via pre-hooks."""

    def __init__(self, module):
        self.records = []
        self.handles = []
        for m in module.modules():
            if isinstance(m, VectorQuantizer):
                self.handles.append(m.register_forward_pre_hook(self._hook))

    def _hook(self, mod, args):
        with torch.no_grad():
            x = args[0].permute(0, 2, 1).contiguous().view(-1, mod.embedding_dim)
            w = mod.embedding.weight
            d = (torch.sum(x ** 2, dim=1, keepdim=True) + torch.sum(w ** 2, dim=1)
                 - 2 * torch.matmul(x, w.t()))
            self.records.append((npy(torch.argmin(d, dim=1)), npy(d)))

    def pop(self):
        r, self.records = self.records, []
        return r


def vq_state(prefix, m, out):
    out[prefix + "embedding"] = npy(m.embedding.weight)
    if m.use_ema:
        out[prefix + "ema_cluster_size"] = npy(m.ema_cluster_size)
        out[prefix + "ema_w"] = npy(m.ema_w)


def run_vq(name, K, D, B, T, use_ema, steps, seed, permuted=False, scale=1.0, with_eval=False,
           keep_dist=True):
    torch.manual_seed(seed)
    m = VectorQuantizer(K, D, use_ema=use_ema)
    tap = IndexTap(m)
    out = {"K": K, "D": D, "use_ema": int(use_ema), "steps": steps,
           "commitment_cost": m.commitment_cost, "decay": 0.99}
    vq_state("init.", m, out)
    m.train()
    for s in range(steps):
        if permuted:
            z = (torch.randn(B, T, D) * scale).permute(0, 2, 1)
        else:
            z = torch.randn(B, D, T) * scale
        z = z.clone().requires_grad_(True) if not permuted else z.detach().requires_grad_(True)
        g = torch.randn(B, D, T)
        loss, q, met = m(z)
        (idx, dist), = tap.pop()
        m.zero_grad()
        (loss * 1.7 + (q * g).sum()).backward()
        p = f"s{s}."
        out[p + "z"] = npy(z); out[p + "g"] = npy(g)
        out[p + "loss"] = npy(loss); out[p + "quantized"] = npy(q)
        out[p + "perplexity"] = npy(met["perplexity"]); out[p + "dcr"] = npy(met["dcr"])
        out[p + "indices"] = idx.reshape(B, T)
        if keep_dist:
            out[p + "distances"] = dist
        out[p + "grad_z"] = npy(z.grad)
        if m.embedding.weight.grad is not None:
            out[p + "grad_embedding"] = npy(m.embedding.weight.grad)
        vq_state(p + "after.", m, out)
    if with_eval:
        m.eval()
        z = torch.randn(B, D, T) * scale
        with torch.no_grad():
            loss, q, met = m(z)
        (idx, dist), = tap.pop()
        out["eval.z"] = npy(z); out["eval.loss"] = npy(loss); out["eval.quantized"] = npy(q)
        out["eval.perplexity"] = npy(met["perplexity"]); out["eval.dcr"] = npy(met["dcr"])
        out["eval.indices"] = idx.reshape(B, T)
        if keep_dist:
            out["eval.distances"] = dist
        vq_state("eval.after.", m, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_rvq(name, S, K, D, B, T, use_ema, steps, seed):
    torch.manual_seed(seed)
    m = ResidualVQ(S, K, D, use_ema=use_ema)
    tap = IndexTap(m)
    out = {"S": S, "K": K, "D": D, "use_ema": int(use_ema), "steps": steps}
    for i, l in enumerate(m.layers):
        vq_state(f"init.layers.{i}.", l, out)
    m.train()
    for s in range(steps):
        z = (torch.randn(B, D, T) * 0.7).requires_grad_(True)
        g = torch.randn(B, D, T)
        loss, q, met = m(z)
        recs = tap.pop()
        m.zero_grad()
        (loss * 0.9 + (q * g).sum()).backward()
        p = f"s{s}."
        out[p + "z"] = npy(z); out[p + "g"] = npy(g)
        out[p + "loss"] = npy(loss); out[p + "quantized"] = npy(q)
        out[p + "perplexity"] = npy(met["perplexity"]); out[p + "dcr"] = npy(met["dcr"])
        out[p + "indices"] = np.stack([r[0].reshape(B, T) for r in recs], 0)
        out[p + "distances"] = np.stack([r[1] for r in recs], 0)
        out[p + "grad_z"] = npy(z.grad)
        for i, l in enumerate(m.layers):
            vq_state(p + f"after.layers.{i}.", l, out)
            if l.embedding.weight.grad is not None:
                out[p + f"grad_embedding.{i}"] = npy(l.embedding.weight.grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def proj_state(prefix, m, out):
    for k, v in m.state_dict().items():
        out[prefix + k] = npy(v)


def run_fsq(name, D, B, T, seed, scale, crafted=False):
    torch.manual_seed(seed)
    levels = [8, 5, 5, 5]
    if crafted:
        D = 4
    m = FSQ(levels, D, D)
    if crafted:   # identity projection so that the rounding sees hand-picked values
        with torch.no_grad():
            m.project_in.weight.copy_(torch.eye(4).unsqueeze(-1)); m.project_in.bias.zero_()
        vals = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.0, -0.0, 0.49999997, 0.50000006,
                             3.5, -3.5, 1e-30, -1e-30, 7.5, 8.5, 1234.5, -1234.5, 8388607.5, 16777216.0,
                             -4.5, 4.5, 2.4999998, 0.99999994, 30000.7, -30000.2, 1e6 + 0.5, 5.5, 6.5, -6.5,
                             100.5, 101.5, -100.5, -101.5, 0.25, 0.75, -0.25, -0.75, 3.0, -3.0])
        z = vals.view(1, 4, 10).repeat(3, 1, 1)
        z[1] = z[1].flip(-1); z[2] = -z[2]
        B, T = 3, 10
    else:
        z = torch.randn(B, D, T) * scale
    z = z.clone().requires_grad_(True)
    g = torch.randn(B, D, T)
    z_e = m.project_in(z)
    loss, q, met = m(z)
    m.zero_grad()
    (loss + (q * g).sum()).backward()
    out = {"levels": np.asarray(levels), "D": D}
    proj_state("state.", m, out)
    with torch.no_grad():
        zt = z_e.permute(0, 2, 1)
        z_hard = zt + (torch.round(zt) - zt)
        idx = (z_hard * m._basis).sum(dim=-1).long()
    out.update({"z": npy(z), "g": npy(g), "z_e": npy(z_e), "z_hard": npy(z_hard.permute(0, 2, 1)),
                "indices": npy(idx), "loss": npy(loss), "quantized": npy(q),
                "perplexity": npy(met["perplexity"]), "dcr": npy(met["dcr"]),
                "grad_z": npy(z.grad),
                "grad.project_in.weight": npy(m.project_in.weight.grad),
                "grad.project_in.bias": npy(m.project_in.bias.grad),
                "grad.project_out.weight": npy(m.project_out.weight.grad),
                "grad.project_out.bias": npy(m.project_out.bias.grad)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_lfq(name, D, B, T, seed, scale, crafted=False):
    torch.manual_seed(seed)
    d = 10
    if crafted:
        D = d
    m = LFQ(D, codebook_dim=d)
    if crafted:
        with torch.no_grad():
            m.project_in.weight.copy_(torch.eye(d).unsqueeze(-1)); m.project_in.bias.zero_()
        z = torch.randn(4, d, 7) * 3
        z[0, :, 0] = 0.0; z[1, ::2, 1] = -0.0; z[2, :, 2] = 1e-38; z[3, :, 3] = -1e-38
        z[0, :, 4] = 40.0; z[1, :, 5] = -40.0; z[2, :, 6] = torch.arange(d).float() - 4.5
        B, T = 4, 7
    else:
        z = torch.randn(B, D, T) * scale
    z = z.clone().requires_grad_(True)
    g = torch.randn(B, D, T)
    z_e = m.project_in(z).detach().requires_grad_(True)
    # module-level pass
    loss, q, met = m(z)
    m.zero_grad()
    (loss * 1.3 + (q * g).sum()).backward()
    # elementwise-level gradient w.r.t. z_e (for the K5 backward kernel)
    g_zq = torch.randn(B, d, T)
    zq = torch.where(z_e > 0, torch.tensor(1.0), torch.tensor(-1.0))
    zq = z_e + (zq - z_e).detach()
    prob = torch.sigmoid(z_e)
    ent = -(prob * torch.log(prob + 1e-6) + (1 - prob) * torch.log(1 - prob + 1e-6))
    l2 = -ent.mean() * m.entropy_loss_weight
    (l2 * 1.3 + (zq * g_zq).sum()).backward()
    with torch.no_grad():
        bits = (zq > 0).int().permute(0, 2, 1)
        idx = (bits * m._basis).sum(dim=-1)
    out = {"D": D, "d": d, "entropy_loss_weight": m.entropy_loss_weight}
    proj_state("state.", m, out)
    out.update({"z": npy(z), "g": npy(g), "z_e": npy(z_e), "z_q": npy(zq), "indices": npy(idx),
                "loss": npy(loss), "quantized": npy(q),
                "perplexity": npy(met["perplexity"]), "dcr": npy(met["dcr"]),
                "grad_z": npy(z.grad), "g_zq": npy(g_zq), "grad_z_e": npy(z_e.grad), "g_loss": 1.3,
                "grad.project_in.weight": npy(m.project_in.weight.grad),
                "grad.project_in.bias": npy(m.project_in.bias.grad),
                "grad.project_out.weight": npy(m.project_out.weight.grad),
                "grad.project_out.bias": npy(m.project_out.bias.grad)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_hybrid(name, D, K, B, T, steps, seed, permuted):
    torch.manual_seed(seed)
    m = HybridVQ(D, [8, 5, 5, 5], vq_codebook_size=K)
    tap = IndexTap(m)
    out = {"D": D, "K": K, "S": 4, "steps": steps, "levels": np.asarray([8, 5, 5, 5])}
    proj_state("init.", m, out)
    m.train()
    for s in range(steps):
        z = torch.randn(B, T, D).permute(0, 2, 1) if permuted else torch.randn(B, D, T)
        z = z.detach().requires_grad_(True)
        g = torch.randn(B, D, T)
        loss, q, met = m(z)
        recs = tap.pop()
        m.zero_grad()
        (loss + (q * g).sum()).backward()
        p = f"s{s}."
        out[p + "z"] = npy(z); out[p + "g"] = npy(g)
        out[p + "loss"] = npy(loss); out[p + "quantized"] = npy(q)
        for k in ("perplexity", "dcr", "rvq_ppl"):
            out[p + k] = npy(met[k])
        out[p + "indices"] = np.stack([r[0].reshape(B, T) for r in recs], 0)
        out[p + "distances"] = np.stack([r[1] for r in recs], 0)
        out[p + "grad_z"] = npy(z.grad)
        for k in ("project_in.weight", "project_in.bias", "project_out.weight", "project_out.bias"):
            out[p + "grad.fsq." + k] = npy(dict(m.fsq.named_parameters())[k].grad)
        proj_state(p + "after.", m, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_state_dict_keys():
    table = {}
    for arch in ("transformer", "resnet_no_down", "resnet", "simple"):
        for method in ("standard", "ema", "rvq", "fsq", "lfq", "hybrid", "ae"):
            torch.manual_seed(0)
            m = DualMotionVQVAE(human_input_dim=126, robot_input_dim=29, hidden_dim=64,
                                arch=arch, method=method, window_size=10)
            table[f"{arch}/{method}"] = [[k, list(v.shape), str(v.dtype).replace("torch.", "")]
                                         for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(table, f)


def run_model(name, arch, method, seed, window):
    """Whole-model known answer (eval mode so BatchNorm/Dropout are deterministic)."""
    torch.manual_seed(seed)
    m = DualMotionVQVAE(human_input_dim=12, robot_input_dim=7, hidden_dim=16, codebook_size=64,
                        arch=arch, method=method, n_layers=2, window_size=window)
    out = {}
    # a couple of training steps first so that EMA buffers / BN stats are non-trivial
    m.train()
    for s in range(2):
        xr = torch.randn(6, window, 7); xh = torch.randn(6, window, 12)
        m(x_robot=xr, x_human=xh)
    for k, v in m.state_dict().items():
        out["state." + k] = npy(v)
    m.eval()
    xr = torch.randn(5, window, 7); xh = torch.randn(5, window, 12)
    with torch.no_grad():
        o = m(x_robot=xr, x_human=xh)
    out.update({"x_robot": npy(xr), "x_human": npy(xh),
                "robot.recon": npy(o["robot"]["recon"]), "robot.loss_vq": npy(o["robot"]["loss_vq"]),
                "robot.z_e": npy(o["robot"]["z_e"]),
                "human.retargeted": npy(o["human"]["retargeted"]), "human.loss_vq": npy(o["human"]["loss_vq"]),
                "human.z_e": npy(o["human"]["z_e"])})
    for br in ("robot", "human"):
        for k, v in o[br]["metrics"].items():
            out[f"{br}.metrics.{k}"] = npy(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    run_vq("vq_std_small", K=64, D=16, B=4, T=10, use_ema=False, steps=2, seed=11)
    run_vq("vq_ema_fresh", K=128, D=32, B=8, T=10, use_ema=True, steps=3, seed=12, with_eval=True)
    run_vq("vq_ema_k1024_perm", K=1024, D=64, B=256, T=1, use_ema=True, steps=2, seed=13,
           permuted=True, keep_dist=False)
    run_vq("vq_std_ragged", K=37, D=24, B=3, T=7, use_ema=False, steps=1, seed=14, with_eval=True)
    run_rvq("rvq_ema", S=4, K=64, D=16, B=6, T=10, use_ema=True, steps=3, seed=21)
    run_rvq("rvq_std", S=3, K=32, D=8, B=5, T=4, use_ema=False, steps=1, seed=22)
    run_fsq("fsq_module", D=16, B=5, T=10, seed=31, scale=1.0)
    run_fsq("fsq_module_x30", D=16, B=5, T=10, seed=32, scale=30.0)
    run_fsq("fsq_crafted", D=4, B=3, T=10, seed=33, scale=1.0, crafted=True)
    run_lfq("lfq_module", D=16, B=5, T=10, seed=41, scale=1.0)
    run_lfq("lfq_crafted", D=10, B=4, T=7, seed=42, scale=1.0, crafted=True)
    run_hybrid("hybrid_perm", D=16, K=64, B=32, T=1, steps=3, seed=51, permuted=True)
    run_hybrid("hybrid_t10", D=16, K=32, B=4, T=10, steps=2, seed=52, permuted=False)
    run_state_dict_keys()
    run_model("model_resnet_no_down_ema", "resnet_no_down", "ema", 61, 10)
    run_model("model_resnet_no_down_hybrid", "resnet_no_down", "hybrid", 62, 10)
    sizes = {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE))}
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
