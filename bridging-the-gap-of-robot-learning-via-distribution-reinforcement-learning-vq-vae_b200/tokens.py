"""Token export and the decode-only path.

The reference never materialises tokens: `scripts/deployment/export_motion.py:25-83` pushes every sliding window
through encoder -> quantizer -> decoder with B = 1.  Here the quantizer's device-side results (indices / FSQ digits)
are bit-packed into a compact token stream on the GPU, written to a small self-describing file, and turned back into
the quantized latent `[B, C, T]` for the decoder without touching the encoder:

    tok = vqb200.tokens.encode(model.quantizer, z_e)        # runs the quantizer (eval), packs on the device
    vqb200.tokens.save("motion.vqtok", tok)
    tok = vqb200.tokens.load("motion.vqtok", device="cuda")
    z_q = vqb200.tokens.decode(model.quantizer, tok)        # == quantizer(z_e)[1] to 1e-6
    recon = model.robot_decoder(z_q)

Token layout per latent vector (little-endian bit stream, whole bytes): S codebook indices of `code_bits` bits, then
d signed FSQ digits of `digit_bits` bits.  hybrid (FSQ d=4 + 4 x K=512): 4*9 + 4*8 = 68 bits -> 9 bytes instead of
256 bytes of fp32 latent.  FSQ digits are stored instead of the mixed-radix index because the reference's rounding is
unbounded (`models/vqvae.py:127-131`) and that index is not invertible.
"""
from __future__ import annotations

import ctypes
import json
import struct
from dataclasses import dataclass, asdict
from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .quantizers import VectorQuantizer, ResidualVQ, FSQ, LFQ, HybridVQ

MAGIC = b"VQTK1\n"


@dataclass
class TokenSpec:
    method: str          # "vq" | "rvq" | "fsq" | "lfq" | "hybrid"
    B: int
    C: int
    T: int
    S: int               # codebook indices per token
    code_bits: int
    d: int               # FSQ digits per token
    digit_bits: int
    bytes_per_token: int


@dataclass
class TokenBatch:
    spec: TokenSpec
    data: torch.Tensor   # uint8 [B*T, bytes_per_token]
    saturated: Optional[torch.Tensor] = None   # int32 device scalar: 1 if a field did not fit its width


def _bits_for(K: int) -> int:
    return max(1, (int(K) - 1).bit_length())


def spec_for(module, B: int, C: int, T: int, digit_bits: int = 8) -> TokenSpec:
    """Token layout of one quantizer module (no GPU needed)."""
    if isinstance(module, HybridVQ):
        layers = list(module.vq.layers)
        m, S, cb, d, db = "hybrid", len(layers), _bits_for(max(l.num_embeddings for l in layers)), module.fsq.fsq_dim, digit_bits
    elif isinstance(module, ResidualVQ):
        layers = list(module.layers)
        m, S, cb, d, db = "rvq", len(layers), _bits_for(max(l.num_embeddings for l in layers)), 0, 0
    elif isinstance(module, VectorQuantizer):
        m, S, cb, d, db = "vq", 1, _bits_for(module.num_embeddings), 0, 0
    elif isinstance(module, FSQ):
        m, S, cb, d, db = "fsq", 0, 0, module.fsq_dim, digit_bits
    elif isinstance(module, LFQ):
        if module.codebook_dim > 31:
            raise RuntimeError("vqb200.tokens: LFQ codebook_dim > 31 does not fit an int32 code")
        m, S, cb, d, db = "lfq", 1, module.codebook_dim, 0, 0
    else:
        raise RuntimeError(f"vqb200.tokens: no token format for {type(module).__name__}")
    bits = S * cb + d * db
    if bits > 256:
        raise RuntimeError(f"vqb200.tokens: {bits} bits per token exceed the 256-bit maximum")
    return TokenSpec(m, int(B), int(C), int(T), S, cb, d, db, (bits + 7) // 8)


def _fields(module, z: torch.Tensor):
    """Run the quantizer in eval mode (no state change) and collect (codes int32 [S,N] | None, z_e [B,d,T] | None)."""
    was_training = module.training
    module.eval()
    try:
        with torch.no_grad():
            module(z)
    finally:
        module.train(was_training)
    B, _, T = z.shape
    codes = z_e = None
    if isinstance(module, HybridVQ):
        codes, z_e = module.vq.last_indices.reshape(-1, B * T), module.fsq.last_z_e
    elif isinstance(module, ResidualVQ):
        codes = module.last_indices.reshape(-1, B * T)
    elif isinstance(module, VectorQuantizer):
        codes = module.last_indices.reshape(1, B * T)
    elif isinstance(module, FSQ):
        z_e = module.last_z_e
    elif isinstance(module, LFQ):
        codes = module.last_indices.reshape(1, B * T).to(torch.int32)
    if codes is not None:
        codes = codes.to(torch.int32).contiguous()
    if z_e is not None:
        z_e = z_e.to(torch.float32).contiguous()
    return codes, z_e


def encode(module, z: torch.Tensor, digit_bits: int = 8) -> TokenBatch:
    """Quantize `z` [B,C,T] (eval semantics: no EMA update) and bit-pack the result on the device."""
    if not z.is_cuda:
        raise RuntimeError("vqb200.tokens.encode: input must be a CUDA tensor -- there is no CPU path")
    lib = _lib.load()
    B, C, T = z.shape
    spec = spec_for(module, B, C, T, digit_bits)
    codes, z_e = _fields(module, z)
    dev = z.device
    data = torch.empty((B * T, spec.bytes_per_token), dtype=torch.uint8, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.vqb200_tokens_pack(ptr(codes), spec.S, spec.code_bits, ptr(z_e), spec.d, spec.digit_bits, B, T,
                                     ptr(data), ptr(ovf), stream_ptr(dev)), "tokens_pack")
    return TokenBatch(spec, data, ovf)


def unpack(tok: TokenBatch):
    """-> (codes int32 [S, B*T] | None, digits fp32 [B, d, T] | None) on the token batch's device."""
    lib = _lib.load()
    sp = tok.spec
    dev = tok.data.device
    if not tok.data.is_cuda:
        raise RuntimeError("vqb200.tokens.unpack: tokens must live on a CUDA device (load(..., device='cuda'))")
    codes = torch.empty((sp.S, sp.B * sp.T), dtype=torch.int32, device=dev) if sp.S else None
    digits = torch.empty((sp.B, sp.d, sp.T), dtype=torch.float32, device=dev) if sp.d else None
    with torch.cuda.device(dev):
        check(lib.vqb200_tokens_unpack(ptr(tok.data.contiguous()), sp.S, sp.code_bits, sp.d, sp.digit_bits, sp.B, sp.T,
                                       ptr(codes), ptr(digits), stream_ptr(dev)), "tokens_unpack")
    return codes, digits


def decode(module, tok: TokenBatch) -> torch.Tensor:
    """Tokens -> quantized latent [B,C,T] (what `module(z)[1]` returned at encode time), ready for the decoder."""
    lib = _lib.load()
    sp = tok.spec
    _validate_spec(sp, "vqb200.tokens.decode")
    _check_module(module, sp)
    codes, digits = unpack(tok)
    dev = tok.data.device
    out = torch.empty((sp.B, sp.C, sp.T), dtype=torch.float32, device=dev)
    if sp.method == "lfq":
        # the code IS the sign pattern: z_q = +-1 per bit, then the stock 1x1 project_out (:179)
        bits = (codes.reshape(sp.B, sp.T, 1) >> torch.arange(module.codebook_dim, device=dev, dtype=torch.int32)) & 1
        z_q = (bits.to(torch.float32) * 2.0 - 1.0).permute(0, 2, 1).contiguous()
        with torch.no_grad():
            return module.project_out(z_q)
    weights = []
    if sp.method == "hybrid":
        weights = [l.embedding.weight.detach().contiguous() for l in module.vq.layers]
        fsq = module.fsq
    elif sp.method == "rvq":
        weights, fsq = [l.embedding.weight.detach().contiguous() for l in module.layers], None
    elif sp.method == "vq":
        weights, fsq = [module.embedding.weight.detach().contiguous()], None
    else:
        fsq = module
    S = len(weights)
    if S != sp.S:
        raise RuntimeError(f"vqb200.tokens.decode: module has {S} codebooks, tokens carry {sp.S}")
    Es = (ctypes.c_void_p * max(S, 1))(*[w.data_ptr() for w in weights])
    Ks = (ctypes.c_int64 * max(S, 1))(*[w.shape[0] for w in weights])
    w_out = fsq.project_out.weight.detach().contiguous() if fsq is not None else None
    b_out = fsq.project_out.bias.detach().contiguous() if fsq is not None else None
    with torch.cuda.device(dev):
        check(lib.vqb200_tokens_decode(ptr(codes), S, Es, Ks, ptr(digits), sp.d, ptr(w_out), ptr(b_out),
                                       sp.B, sp.C, sp.T, ptr(out), stream_ptr(dev)), "tokens_decode")
    return out


def _validate_spec(sp: TokenSpec, where: str) -> None:
    """A token header is untrusted input: the device kernels index with these fields."""
    ints = (sp.B, sp.C, sp.T, sp.S, sp.code_bits, sp.d, sp.digit_bits, sp.bytes_per_token)
    if any((not isinstance(v, int)) or v < 0 for v in ints) or sp.method not in ("vq", "rvq", "fsq", "lfq", "hybrid"):
        raise RuntimeError(f"{where}: malformed token header")
    if sp.code_bits > 31 or sp.digit_bits > 32 or sp.S > 64 or sp.d > 64:
        raise RuntimeError(f"{where}: token header out of range (code_bits={sp.code_bits}, digit_bits={sp.digit_bits}, S={sp.S}, d={sp.d})")
    want = (sp.S * sp.code_bits + sp.d * sp.digit_bits + 7) // 8
    if sp.bytes_per_token != want:
        raise RuntimeError(f"{where}: header says {sp.bytes_per_token} bytes per token, its layout needs {want}")


def _check_module(module, sp: TokenSpec) -> None:
    """The tokens must belong to a module of this shape before any kernel indexes its weights with them."""
    if sp.method == "hybrid":
        dims, d = [l.embedding_dim for l in module.vq.layers], module.fsq.fsq_dim
        po = module.fsq.project_out.weight
    elif sp.method == "rvq":
        dims, d, po = [l.embedding_dim for l in module.layers], 0, None
    elif sp.method == "vq":
        dims, d, po = [module.embedding_dim], 0, None
    elif sp.method == "fsq":
        dims, d, po = [], module.fsq_dim, module.project_out.weight
    else:
        return
    if any(c != sp.C for c in dims):
        raise RuntimeError(f"vqb200.tokens.decode: tokens are for C={sp.C}, module has embedding_dim {dims}")
    if d != sp.d:
        raise RuntimeError(f"vqb200.tokens.decode: tokens carry {sp.d} FSQ digits, module has {d}")
    if po is not None and (po.shape[0] != sp.C or po.shape[1] != sp.d):
        raise RuntimeError(f"vqb200.tokens.decode: project_out is {tuple(po.shape)}, tokens need ({sp.C}, {sp.d}, 1)")


# ---- file format: MAGIC | u32 header length | JSON header (TokenSpec) | payload (B*T*bytes_per_token bytes) --------
def save(path: str, tok: TokenBatch) -> None:
    if tok.saturated is not None and int(tok.saturated.item()) != 0:
        raise RuntimeError("vqb200.tokens.save: a field did not fit its bit width (raise digit_bits)")
    header = json.dumps(asdict(tok.spec)).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        f.write(tok.data.detach().cpu().numpy().tobytes())


def load(path: str, device="cpu") -> TokenBatch:
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise RuntimeError(f"{path}: not a vqb200 token file")
        (hl,) = struct.unpack("<I", f.read(4))
        spec = TokenSpec(**json.loads(f.read(hl).decode()))
        payload = f.read()
    _validate_spec(spec, path)
    n = spec.B * spec.T * spec.bytes_per_token
    if len(payload) != n:
        raise RuntimeError(f"{path}: payload has {len(payload)} bytes, header promises {n}")
    data = torch.frombuffer(bytearray(payload), dtype=torch.uint8).reshape(spec.B * spec.T, spec.bytes_per_token)
    return TokenBatch(spec, data.to(device))
