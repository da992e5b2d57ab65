"""ctypes binding of libvqb200.so (the C ABI declared in include/vqb200.h).

The library is the ONLY compute path: if it cannot be loaded the import of any quantizer op raises.
There is no CPU / eager fallback.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

from . import build as _build

_P = c_void_p
_LIB = None

ASSIGN_AUTO, ASSIGN_SIMT, ASSIGN_TC, ASSIGN_TC_SPLIT = 0, 1, 2, 3

_SIGNATURES = {
    # name: (restype, [argtypes])
    "vqb200_abi_version": (c_int, []),
    "vqb200_last_error_string": (c_char_p, []),
    "vqb200_launch_count": (c_int64, []),
    "vqb200_codebook_image_bytes": (c_size_t, [c_int64, c_int64]),
    "vqb200_codebook_prepare": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P]),
    "vqb200_assign_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "vqb200_vq_assign": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                 _P, _P, _P, _P, c_int64, _P, _P, _P, c_size_t, c_int, _P]),
    "vqb200_ema_accumulate": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                      _P, _P, c_int64, _P, c_int, _P]),
    "vqb200_ema_finalize": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_double, c_double, _P, _P, _P, _P, _P]),
    "vqb200_peer_alloc": (c_int, [c_size_t, _P, _P]),
    "vqb200_peer_open": (c_int, [_P, _P]),
    "vqb200_peer_close": (c_int, [_P]),
    "vqb200_peer_free": (c_int, [_P]),
    "vqb200_peer_barrier": (c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, _P]),
    "vqb200_ema_finalize_peer": (c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, _P, _P, _P, _P,
                                         c_int64, c_int64, c_double, c_double, _P, _P, _P, _P, _P]),
    "vqb200_vq_histogram": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "vqb200_vq_gather_st": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                    _P, _P, c_int64, _P, _P, _P, c_int, _P, _P]),
    "vqb200_vq_assign_residual": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                          _P, _P, c_int64, _P, _P, _P, _P, _P, c_int64, _P, _P, c_size_t, c_int, _P]),
    "vqb200_proj_fused_eligible": (c_int, [c_int64, c_int64, c_int64, c_int64]),
    "vqb200_proj_fused_grad_floats": (c_size_t, [c_int64, c_int64]),
    "vqb200_fsq_fused_forward": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64,
                                         _P, _P, _P, _P, _P, _P]),
    "vqb200_lfq_fused_forward": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int64, c_float,
                                         _P, _P, _P, _P, _P, _P]),
    "vqb200_proj_fused_backward": (c_int, [c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P, c_float,
                                           _P, _P, _P]),
    "vqb200_token_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    "vqb200_tokens_pack": (c_int, [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P]),
    "vqb200_tokens_unpack": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P]),
    "vqb200_tokens_decode": (c_int, [_P, c_int64, _P, _P, _P, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, _P]),
    "vqb200_codebook_revive": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, c_float,
                                       ctypes.c_uint64, _P, _P, _P, c_int64, _P, _P]),
    "vqb200_rvq_output_chain": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                        ctypes.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "vqb200_rvq_small_eligible": (c_int, [c_int64, c_int64, ctypes.c_int32, _P]),
    "vqb200_rvq_small_workspace_floats": (c_size_t, [ctypes.c_int32, _P]),
    "vqb200_rvq_small_forward": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                         ctypes.c_int32, _P, _P, _P, _P, c_double, c_double, c_float, c_int, c_int,
                                         _P, _P, _P, _P, _P, _P]),
    "vqb200_rvq_small_peer_eligible": (c_int, [c_int64, c_int64, ctypes.c_int32, _P]),
    "vqb200_rvq_small_stats_floats": (c_size_t, [ctypes.c_int32, _P]),
    "vqb200_rvq_small_forward_peer": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                              ctypes.c_int32, _P, _P, _P, _P, c_double, c_double, c_float,
                                              _P, _P, _P, _P, _P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                              c_int64, _P]),
    "vqb200_vq_metrics": (c_int, [_P, c_int64, c_int64, _P, c_int64, c_float, c_int, _P, _P]),
    "vqb200_vq_backward_input": (c_int, [_P, c_int64, c_int64, c_int64,
                                         _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                         _P, _P, c_int64, _P, c_float, _P, _P]),
    "vqb200_vq_backward_codebook": (c_int, [_P, c_int64, c_int64, _P, c_float, _P, _P]),
    "vqb200_unique_workspace_bytes": (c_size_t, []),
    "vqb200_fsq_forward": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int64, _P, _P, _P, _P, _P]),
    "vqb200_lfq_forward": (c_int, [_P, c_int64, c_int64, c_int64, c_float, _P, _P, _P, _P, _P]),
    "vqb200_lfq_backward": (c_int, [_P, _P, _P, c_int64, c_float, _P, _P]),
}


def exported_symbols():
    """Every symbol include/vqb200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen the in-tree library (building it with nvcc first if the .so is absent or stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("VQB200_LIB_PATH") or _build.LIB      # development: A/B builds of the same sources
    own = path == _build.LIB
    if build_if_missing and own and (not os.path.exists(path) or os.environ.get("VQB200_REBUILD") == "1"
                                     or (_build.have_nvcc() and _build.needs_build())):
        _build.build_locked(force=os.environ.get("VQB200_REBUILD") == "1")
    if not os.path.exists(path):
        raise RuntimeError(f"vqb200: CUDA library {path} is missing and could not be built; "
                           "there is no CPU fallback (run `python __graft_entry__.py build`)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.vqb200_abi_version() != 1:
        raise RuntimeError("vqb200: ABI version mismatch")
    _LIB = lib
    return lib


def last_error():
    return load().vqb200_last_error_string().decode("utf-8", "replace")


def launch_count():
    return int(load().vqb200_launch_count())


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"vqb200.{what} failed (rc={rc}): {last_error()}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL).  Anything that is not on a CUDA device is rejected:
    the engine has no CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("vqb200: tensors must live on a CUDA device (no CPU fallback)")
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
