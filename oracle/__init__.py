"""CPU oracle for the quantizer hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy (fp32) restatement of the reference's quantizer algorithms
(`/root/reference/models/vqvae.py:10-259`).  It exists to *check* the CUDA path:
only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product path (the `vqb200` package and the drop-in
`models/vqvae.py`) never imports anything from here and fails loudly when the CUDA
library is missing.

Pinning: the reference ships no tests and no golden vectors for this path,
so the oracle is pinned against outputs of the *unmodified reference modules* imported in
the build container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`).  See
`tests/test_oracle_golden.py`.
"""
from .vq_oracle import (  # noqa: F401
    vq_distances, vq_forward, vq_backward, rvq_forward, rvq_backward,
    fsq_quantize, fsq_forward, lfq_quantize, lfq_forward, lfq_backward_ze,
    hybrid_forward, conv1x1, VQState,
)
from .compare import (  # noqa: F401
    check_indices, rel_err, assert_close,
)
