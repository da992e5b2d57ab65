// vqb200 -- tcgen05 / TMEM wrappers and UMMA descriptors shared by the tensor-core assignment kernels.
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace vqb200 {
namespace tcc {
using namespace ptx;

constexpr int TILE_M = 128;                 // rows per MMA (accumulator lanes)
constexpr int BN = 128;                     // codes per accumulator (== IMG_TILE_CODES)

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
               " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
// mbarrier arrive that ptxas cannot hoist above the math producing (a, b): the values are first stored to a shared
// sink word, and the arrive (release semantics) is ordered after that store.  A bare arrive was measured to be scheduled
// before the last chunk's math of the epilogue (mis-assigned rows, run to run different, see assign_f16.cu); an
// always-true predicate on the values "works" until the bit pattern it excludes turns up (it did: a lost arrive = hang).
__device__ __forceinline__ void arrive_after(uint32_t bar, uint32_t sink, float a, float b) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(sink), "r"(__float_as_uint(a) ^ __float_as_uint(b)) : "memory");
  mbar_arrive(bar);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 128-byte rows, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address        [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset  [16,30)  (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset   [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);


__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {      // a -> low half, b -> high half
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// (score & ~127) | column in ONE alu op: lop3 with LUT (a & c) | b
__device__ __forceinline__ float pack_col(float s, uint32_t col, uint32_t mask) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(r) : "r"(__float_as_uint(s)), "r"(col), "r"(mask));
  return __uint_as_float(r);
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Running top-2 of one 32-column chunk of scores: s = acc + nh (nh = -|E|^2/2), column index packed into the low
// 7 mantissa bits, 3.5 ALU ops per score.  `base` = first column of the chunk inside the 128-code tile.
#ifndef VQB200_TOP2_VARIANT
#define VQB200_TOP2_VARIANT 0      // measured on B200 (vq_assign, 10 M x 1024, one variant per build): V = 0 4.26 ms,
                                   // V = 1 4.37 ms, V = 2 4.54 ms -- the epilogue warps are bound by their own instruction
                                   // count (4.5 / 5.0 / 5.5 per score), not by the alu pipe, so V = 0 stays
#endif
constexpr int TOP2_VARIANT = VQB200_TOP2_VARIANT;
// V selects where the runner-up arithmetic runs (an experiment kept for the record: moving min() from the alu pipe to
// add/sub on the fma pipe is legal because t2 only feeds the inequality g1 - g2 > thr, whose budget covers a few ulps):
//   V = 0: lo = min(p0,p1), m = min(t1,hi)                      3.5 alu + 1 fma ops per score
//   V = 1: m = (t1 + hi) - max(t1,hi)  on the fma pipe          3.0 alu + 2 fma
//   V = 2: additionally lo = (p0 + p1) - hi                     2.5 alu + 3 fma
// (-inf) - (-inf) = NaN is harmless: max.f32 drops NaN operands, exactly what a -inf candidate would have done.
template <int V = TOP2_VARIANT>
__device__ __forceinline__ void top2_chunk(const uint32_t (&cur)[32], const float4* nh, int base, uint32_t mask,
                                           float& t1, float& t2) {
#pragma unroll
  for (int e4 = 0; e4 < 8; ++e4) {
    const float4 h = nh[e4];
    const int col = base + e4 * 4;
    const float p0 = pack_col(__uint_as_float(cur[e4 * 4 + 0]) + h.x, col + 0, mask);
    const float p1 = pack_col(__uint_as_float(cur[e4 * 4 + 1]) + h.y, col + 1, mask);
    const float p2 = pack_col(__uint_as_float(cur[e4 * 4 + 2]) + h.z, col + 2, mask);
    const float p3 = pack_col(__uint_as_float(cur[e4 * 4 + 3]) + h.w, col + 3, mask);
    {
      const float hi = fmaxf(p0, p1);
      const float lo = (V >= 2) ? __fsub_rn(__fadd_rn(p0, p1), hi) : fminf(p0, p1);
      const float tn = fmaxf(t1, hi);
      const float m = (V >= 1) ? __fsub_rn(__fadd_rn(t1, hi), tn) : fminf(t1, hi);
      t2 = fmax3(t2, lo, m);
      t1 = tn;
    }
    {
      const float hi = fmaxf(p2, p3);
      const float lo = (V >= 2) ? __fsub_rn(__fadd_rn(p2, p3), hi) : fminf(p2, p3);
      const float tn = fmaxf(t1, hi);
      const float m = (V >= 1) ? __fsub_rn(__fadd_rn(t1, hi), tn) : fminf(t1, hi);
      t2 = fmax3(t2, lo, m);
      t1 = tn;
    }
  }
}

// Rigorous bound on the error of one score for a D-dim dot product done as 3 split-bf16 MMAs:
// split remainder (3*2^-18) + fp32 accumulation over 3*D products (truncation assumed) + rounding of the
// -|E|^2/2 add + 7 index bits packed into the mantissa.  mag = |x| * max|E|.
__device__ __forceinline__ float score_error_bound(float mag, float emax, int D) {
  const float acc = 3.0f * (float)D * 1.2e-7f;
  // + 1e-6 on both terms: the runner-up may be formed by add/sub on the fma pipe (top2_chunk<V > 0>)
  return (1.15e-5f + acc + 1.6e-5f + 1.0e-6f) * mag + 1.8e-5f * (0.5f * emax * emax);
}

}  // namespace tcc
}  // namespace vqb200
