// vqb200 -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>

#include "../../include/vqb200.h"

namespace vqb200 {

// ------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the ABI)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  fail(int code, const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int  sm_count();
int  current_device();           // cudaGetDevice() & 63

// cudaFuncSetAttribute / occupancy answers are per DEVICE: a call site remembers what it has configured per device
// (a static zero-initialised instance per site; racing first calls configure twice, which is harmless).
struct PerDevice {
  std::atomic<size_t> v[64];
  std::atomic<size_t>& here() { return v[current_device()]; }
};

#define VQ_CHECK_ARG(cond, code, ...) do { if (!(cond)) return ::vqb200::fail((code), __VA_ARGS__); } while (0)
#define VQ_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return ::vqb200::cuda_fail(e__, #expr); } while (0)
#define VQ_LAUNCH_CHECK(name) do { ::vqb200::count_launch(); cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return ::vqb200::cuda_fail(e__, name); } while (0)

// ------------------------------------------------------------------------------------------
// [B,C,T] fp32 view with element strides.  Vector n = b*T + t, component k at
// p[b*sB + k*sC + t*sT].
// ------------------------------------------------------------------------------------------
enum ZMode : int {
  Z_ROW = 0,   // sC == 1: every vector is D contiguous floats (T'=1 permuted view, [N,D] workspaces)
  Z_BCT = 1,   // fully contiguous [B,C,T] with T > 1: a sample is one C*T slab
  Z_GEN = 2    // anything else
};

struct ZView {
  const float* p;
  long long B, C, T, sB, sC, sT;
  long long N;
  int mode;
  __host__ __device__ __forceinline__ long long row_base(long long n) const {
    long long b = n / T, t = n - b * T;
    return b * sB + t * sT;
  }
};

inline ZView make_zview(const float* p, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT) {
  ZView v;
  v.p = p; v.B = B; v.C = C; v.T = T; v.sB = sB; v.sC = sC; v.sT = sT; v.N = B * T;
  if (sC == 1) v.mode = Z_ROW;
  else if (sT == 1 && sC == T && sB == C * T) v.mode = Z_BCT;
  else v.mode = Z_GEN;
  return v;
}

// ------------------------------------------------------------------------------------------
// Cooperative tile loader: brings rows [n0, n0+rows) of the view into shared memory through a
// caller-supplied store functor st(row, k, value).  Global accesses are coalesced for Z_ROW
// (k fastest) and Z_BCT (memory order of the C*T slabs the tile touches).
// ------------------------------------------------------------------------------------------
template <class Store>
__device__ __forceinline__ void load_rows(const ZView& z, long long n0, int rows, int D,
                                          int tid, int nthreads, Store st) {
  if (z.mode == Z_ROW) {
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(z.p) & 15) == 0) && ((z.sB & 3) == 0) && ((z.sT & 3) == 0 || z.T == 1)) {
      const int d4 = D >> 2;
      for (int i = tid; i < rows * d4; i += nthreads) {
        int r = i / d4, q = i - r * d4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(z.p + z.row_base(n0 + r)) + q);
        st(r, 4 * q + 0, v.x); st(r, 4 * q + 1, v.y); st(r, 4 * q + 2, v.z); st(r, 4 * q + 3, v.w);
      }
    } else {
      for (int i = tid; i < rows * D; i += nthreads) {
        int r = i / D, k = i - r * D;
        st(r, k, __ldg(z.p + z.row_base(n0 + r) + k));
      }
    }
  } else if (z.mode == Z_BCT) {
    // memory order of the C*T slabs the tile touches; 32-bit index math (a tile spans < 2^31 elements)
    const unsigned T = (unsigned)z.T;
    const unsigned slab = (unsigned)D * T;
    const long long b_lo = n0 / z.T;
    const long long b_hi = (n0 + rows - 1) / z.T;
    const unsigned total = (unsigned)(b_hi - b_lo + 1) * slab;
    const float* base = z.p + b_lo * (long long)slab;
    const int row0 = (int)(b_lo * z.T - n0);    // row id (relative to the tile) of (b_lo, t=0); <= 0
    for (unsigned i = tid; i < total; i += nthreads) {
      const unsigned bl = i / slab;
      const unsigned rem = i - bl * slab;
      const unsigned k = rem / T;
      const unsigned t = rem - k * T;
      const int r = row0 + (int)(bl * T + t);
      if (r >= 0 && r < rows) st(r, (int)k, __ldg(base + i));
    }
  } else {
    for (int i = tid; i < rows * D; i += nthreads) {
      int r = i / D, k = i - r * D;
      st(r, k, __ldg(z.p + z.row_base(n0 + r) + (long long)k * z.sC));
    }
  }
}

// ------------------------------------------------------------------------------------------
// argmin ordering of torch.argmin: smaller distance wins, a NaN beats every number, ties go to
// the lower index.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool dist_less(float a, float b) {       // a strictly better than b
  return (a < b) || ((a != a) && (b == b));
}
__device__ __forceinline__ bool dist_same(float a, float b) {
  return (a == b) || ((a != a) && (b != b));
}
__device__ __forceinline__ bool cand_better(float da, int ia, float db, int ib) {
  return dist_less(da, db) || (dist_same(da, db) && ia < ib);
}

// (distance, index) -> 64-bit key whose unsigned order is the cand_better order: NaN first, then ascending
// distance (-0 == +0), ties by ascending index.
__device__ __forceinline__ unsigned long long cand_key(float d, int i) {
  unsigned u;
  if (d != d) u = 0u;
  else {
    if (d == 0.f) d = 0.f;
    u = __float_as_uint(d);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (u == 0u) u = 1u;                           // keep 0 for NaN only (-NaN bit patterns never reach here)
  }
  return ((unsigned long long)u << 32) | (unsigned)i;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector reduction to global memory (sm_90+): one L2 atomic transaction per 4 floats.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

inline int grid_for(long long work_items, int per_block, int max_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace vqb200
