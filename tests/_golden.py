"""Helpers shared by the CPU (oracle) and GPU (engine) golden tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def vq_state_from(g, prefix, use_ema, commitment_cost=0.25, decay=0.99):
    from oracle import VQState
    return VQState(g[prefix + "embedding"].copy(),
                   g[prefix + "ema_cluster_size"].copy() if use_ema else None,
                   g[prefix + "ema_w"].copy() if use_ema else None,
                   commitment_cost, bool(use_ema), decay)
