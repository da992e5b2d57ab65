"""vqb200 -- B200-native vector-quantization engine (host side).

Python mirror of the reference's quantizer layer (`models/vqvae.py:10-259`) on top of the C ABI in
`include/vqb200.h` / `lib/libvqb200.so` (hand-written sm_100a CUDA).  Import as `vqb200` through the
shim at the repository root (this directory's name is not a Python identifier).
"""
from . import _lib  # noqa: F401
from .functional import (  # noqa: F401
    vq_assign, vq_quantize, rvq_quantize, fsq_round, lfq_sign, codebook_prepare, QuantizerState,
)
from .quantizers import (  # noqa: F401
    VectorQuantizer, ResidualVQ, FSQ, LFQ, HybridVQ, IdentityVQ,
)
from . import dist  # noqa: F401
from . import tokens  # noqa: F401
from . import trainer  # noqa: F401
from .graphs import GraphedQuantizerStep  # noqa: F401

__all__ = ["VectorQuantizer", "ResidualVQ", "FSQ", "LFQ", "HybridVQ", "IdentityVQ",
           "vq_assign", "vq_quantize", "rvq_quantize", "fsq_round", "lfq_sign", "codebook_prepare",
           "QuantizerState", "dist", "tokens", "trainer", "GraphedQuantizerStep"]
