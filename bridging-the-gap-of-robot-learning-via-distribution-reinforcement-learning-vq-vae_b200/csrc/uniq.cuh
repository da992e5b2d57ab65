// vqb200 -- on-device unique-code counting shared by the FSQ / LFQ kernels (replaces the host-synchronising
// torch.unique().numel() of models/vqvae.py:142 and :186): a bitmap over a window around 0, a hash set for codes
// outside it and a last-CTA finalize.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace vqb200 {

// ---- unique-code workspace ----------------------------------------------------------------
//   [0]   u32 unique count      [4] u32 ticket      [8] u32 hash overflow flag
//   [16]  f64 entropy sum (LFQ)
//   [64 .. 64+UNIQ_BITMAP_BYTES)           bitmap for codes in [-UNIQ_HALF, UNIQ_HALF)
//   [.. + UNIQ_HASH_SLOTS*8)               open-addressing set for codes outside the window (0 = empty)
constexpr long long UNIQ_HALF = 1LL << 20;
constexpr size_t UNIQ_BITMAP_BYTES = (size_t)(2 * UNIQ_HALF) / 8;      // 256 KiB
constexpr int UNIQ_HASH_SLOTS = 1 << 16;
constexpr size_t UNIQ_WS_BYTES = 64 + UNIQ_BITMAP_BYTES + (size_t)UNIQ_HASH_SLOTS * 8;

struct UniqWs {
  unsigned* count; unsigned* ticket; unsigned* overflow; double* ent;
  unsigned* bitmap; unsigned long long* hash;
  __host__ __device__ explicit UniqWs(void* ws) {
    unsigned char* b = reinterpret_cast<unsigned char*>(ws);
    count = reinterpret_cast<unsigned*>(b); ticket = count + 1; overflow = count + 2;
    ent = reinterpret_cast<double*>(b + 16);
    bitmap = reinterpret_cast<unsigned*>(b + 64);
    hash = reinterpret_cast<unsigned long long*>(b + 64 + UNIQ_BITMAP_BYTES);
  }
};

__device__ __forceinline__ void unique_insert(const UniqWs& w, long long code) {
  if (code >= -UNIQ_HALF && code < UNIQ_HALF) {
    const unsigned bit = (unsigned)(code + UNIQ_HALF);
    unsigned* word = w.bitmap + (bit >> 5);
    const unsigned m = 1u << (bit & 31);
    if (!(*reinterpret_cast<volatile unsigned*>(word) & m)) {      // skip the atomic once the bit is visible
      const unsigned old = atomicOr(word, m);
      if (!(old & m)) atomicAdd(w.count, 1u);
    }
  } else {
    const unsigned long long key = (unsigned long long)code;      // never 0: 0 lies inside the window
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    unsigned slot = (unsigned)(h >> 40) & (UNIQ_HASH_SLOTS - 1);
    for (int probe = 0; probe < UNIQ_HASH_SLOTS; ++probe) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(w.hash + slot);
      if (cur == key) return;
      if (cur == 0ull) {
        cur = atomicCAS(w.hash + slot, 0ull, key);
        if (cur == 0ull) { atomicAdd(w.count, 1u); return; }
        if (cur == key) return;
      }
      slot = (slot + 1) & (UNIQ_HASH_SLOTS - 1);
    }
    atomicExch(w.overflow, 1u);                                     // set is full: metrics become NaN
  }
}

// returns true in exactly one thread of the last CTA to finish
__device__ __forceinline__ bool last_block_done(const UniqWs& w) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(w.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  return s_last && threadIdx.x == 0;
}

__device__ __forceinline__ long long trunc_to_i64(float s) {
  // x86 cvttss2si semantics of torch's CPU .long(): NaN / out of range -> INT64_MIN
  if (!(fabsf(s) < 9.2233720368547758e18f)) return LLONG_MIN;
  return (long long)s;
}

constexpr long long Q_LOCAL_HALF = 1LL << 15;                 // CTA-local bitmap window [-2^15, 2^15)
constexpr int Q_LOCAL_WORDS = (int)(2 * Q_LOCAL_HALF / 32);   // 2048 words = 8 KiB of shared memory

}  // namespace vqb200
