"""CUDA-graph replay of a quantizer training step for the launch-bound shapes (cfg1 / cfg2).

Every vqb200 entry point is asynchronous and sync-free (metrics stay on the device, FSQ/LFQ count unique codes on the
device), so forward + EMA update + backward of any quantizer module can be captured once and replayed: the ~45 small
launches of a Hybrid step (plus the stock 1x1 convolutions) collapse into one graph launch.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


class GraphedQuantizerStep:
    """step(z, g) -> (loss, quantized, metrics, grad_z); all returned tensors are static buffers that the next
    call overwrites, and so are the parameters' `.grad` (each replay overwrites them: zero_grad(set_to_none) semantics).  `module` keeps its normal semantics (EMA buffers advance on every replay in train mode)."""

    def __init__(self, module: torch.nn.Module, example_z: torch.Tensor, with_backward: bool = True, warmup: int = 3):
        if not example_z.is_cuda:
            raise RuntimeError("GraphedQuantizerStep: CUDA tensors only (no CPU fallback)")
        self.module = module
        self.with_backward = with_backward
        dev = example_z.device
        # keep the caller's memory format (e.g. the transformer's permuted T'=1 view) for the static input
        self.z = torch.empty_strided(example_z.shape, example_z.stride(), dtype=torch.float32, device=dev)
        self.z.copy_(example_z.detach())
        self.z.requires_grad_(with_backward)
        self.g = torch.zeros(example_z.shape, dtype=torch.float32, device=dev)
        self._one = torch.ones((), device=dev)
        state = {k: v.clone() for k, v in module.state_dict().items()}       # warm-up must not advance EMA state
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            module.load_state_dict(state)
        self.z.grad = None
        self._invalidate()            # capture the refresh of |E|^2 / tile image too: replays never run Python
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.quantized, self.metrics = self._run()
        self.grad_z = self.z.grad
        with torch.no_grad():
            module.load_state_dict(state)                                    # capture itself does not execute
        self._invalidate()

    def _invalidate(self):
        for m in self.module.modules():
            if hasattr(m, "invalidate_cache"):
                m.invalidate_cache()

    def _run(self):
        loss, q, met = self.module(self.z)
        if self.with_backward:
            self.z.grad = None
            # like optimizer.zero_grad(set_to_none=True): every replay WRITES the parameter gradients (into buffers of
            # the graph's pool) instead of accumulating into older ones -- five fewer add kernels per Hybrid step
            for p in self.module.parameters():
                p.grad = None
            torch.autograd.backward([q, loss], [self.g, self._one])
        return loss, q, met

    def __call__(self, z: torch.Tensor, g: Optional[torch.Tensor] = None):
        with torch.no_grad():
            self.z.copy_(z)
            if g is not None:
                self.g.copy_(g)
        self.graph.replay()
        return self.loss, self.quantized, self.metrics, self.grad_z
