"""Experiment table consumed by the reference's reporting scripts
(reference: models/experiment_config.py:9-18 -- ids/names/arch/method only, no codebook sizes)."""

_ROWS = (
    ("simple_ema", "Baseline(Simple)", "simple", "ema"),
    ("resnet_ema", "ResNet+EMA", "resnet", "ema"),
    ("resnet_rvq", "ResNet+RVQ", "resnet", "rvq"),
    ("resnet_fsq", "FSQ", "resnet", "fsq"),
    ("resnet_lfq", "LFQ", "resnet", "lfq"),
    ("resnet_hybrid", "Ours(Dual-Enc+Hybrid)", "resnet", "hybrid"),
)

EXPERIMENTS = [dict(id=i, name=n, arch=a, method=m) for i, n, a, m in _ROWS]
