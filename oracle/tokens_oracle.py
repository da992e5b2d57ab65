"""Numpy restatement of the vqb200 token bit layout (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

There is no reference implementation of tokens: the reference never materialises them
(scripts/deployment/export_motion.py:25-83).  The layout is this repository's own (include/vqb200.h, "token export"),
so this file pins the CUDA pack/unpack kernels against an independent, obviously-correct bit loop.
"""
import numpy as np


def pack(codes, digits, code_bits: int, digit_bits: int):
    """codes int [S,N] or None, digits int [N,d] or None -> uint8 [N, bytes]; little-endian bit stream."""
    S = 0 if codes is None else codes.shape[0]
    d = 0 if digits is None else digits.shape[1]
    N = codes.shape[1] if S else digits.shape[0]
    nbits = S * code_bits + d * digit_bits
    nbytes = (nbits + 7) // 8
    out = np.zeros((N, nbytes), np.uint8)
    for n in range(N):
        acc, pos = 0, 0
        for s in range(S):
            acc |= (int(codes[s, n]) & ((1 << code_bits) - 1)) << pos
            pos += code_bits
        for j in range(d):
            acc |= (int(digits[n, j]) & ((1 << digit_bits) - 1)) << pos
            pos += digit_bits
        out[n] = np.frombuffer(acc.to_bytes(nbytes, "little"), np.uint8)
    return out


def unpack(tokens, S: int, code_bits: int, d: int, digit_bits: int):
    N = tokens.shape[0]
    codes = np.zeros((S, N), np.int64)
    digits = np.zeros((N, d), np.int64)
    for n in range(N):
        acc, pos = int.from_bytes(tokens[n].tobytes(), "little"), 0
        for s in range(S):
            codes[s, n] = (acc >> pos) & ((1 << code_bits) - 1)
            pos += code_bits
        for j in range(d):
            v = (acc >> pos) & ((1 << digit_bits) - 1)
            digits[n, j] = v - (1 << digit_bits) if v >> (digit_bits - 1) else v
            pos += digit_bits
    return codes, digits


# ---- opt-in codebook revive (include/vqb200.h "codebook health"): same hash, numpy --------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def revive(z, E, usage, threshold: float, seed: int, ema_cluster_size=None, ema_w=None):
    """z [B,C,T]; every code with usage < threshold takes row splitmix64(seed + k*golden) mod N of z."""
    B, C, T = z.shape
    rows = np.ascontiguousarray(z.transpose(0, 2, 1)).reshape(B * T, C)
    E = E.copy()
    cs = None if ema_cluster_size is None else ema_cluster_size.copy()
    w = None if ema_w is None else ema_w.copy()
    n_rev = 0
    for k in range(E.shape[0]):
        if usage[k] < threshold:
            n = splitmix64((seed + k * 0x9E3779B97F4A7C15) & _M64) % (B * T)
            E[k] = rows[n]
            if w is not None:
                w[k] = rows[n]
            if cs is not None:
                cs[k] = 1.0
            n_rev += 1
    return E, cs, w, n_rev
