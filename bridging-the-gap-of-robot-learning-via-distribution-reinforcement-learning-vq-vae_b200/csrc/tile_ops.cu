// vqb200 -- HBM-bound row-tile kernels for the contiguous layouts ([B,C,T] contiguous, or [N,D] rows).
//
// A tile is a run of WHOLE samples, i.e. one contiguous byte range of the tensor, so every global access
// is a coalesced 16-byte stream; the channel-major <-> row transposition happens in shared memory:
//   1. stream the tile (and the second operand: running RVQ sum or upstream gradient) into smem,
//   2. walk it row-wise (lanes along the channel dim: codeword rows E[idx] are read as coalesced
//      128-byte lines, smem is read with stride T -- at most 2-way bank conflicts for even T),
//   3. stream the results back.
// Used by vq_gather_st (K2), vq_backward_input (K2b) and ema_accumulate (K3a); the generic strided
// kernels in gather.cu / ema.cu remain the fallback for arbitrary views.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace vqb200 {

constexpr int TILE_ELEMS = 4096;          // gather / backward: 16 KiB per tile buffer (small tiles -> many CTAs per SM)
constexpr int ACC_TILE_ELEMS = 8192;      // accumulate: 32 KiB tiles (fewer histogram flushes)
constexpr int TILE_NT = 256;

struct TileGeom {
  int D, T;              // channels, time steps (T == 1 for row-major [N,D])
  int rows_per_tile;     // multiple of T
  int dshift;            // log2(D) if D is a power of two, else -1
  long long N;           // total rows
  long long ntiles;
};

inline bool tile_geom(const ZView& z, TileGeom& g, int tile_elems = TILE_ELEMS) {
  const bool rows_contig = z.mode == Z_ROW && z.T == 1 && z.sB == z.C;
  const bool bct = z.mode == Z_BCT;
  if (!rows_contig && !bct) return false;
  if ((z.C & 3) || (reinterpret_cast<uintptr_t>(z.p) & 15)) return false;
  const long long slab = z.C * z.T;
  if (slab > tile_elems || slab <= 0) return false;
  g.D = (int)z.C; g.T = (int)z.T;
  if (z.T > 256) return false;
  g.rows_per_tile = (int)min((tile_elems / slab) * z.T, (256 / z.T) * z.T);
  g.dshift = -1;
  for (int sh = 0; sh < 16; ++sh) if ((1 << sh) == g.D) g.dshift = sh;
  g.N = z.N;
  g.ntiles = (z.N + g.rows_per_tile - 1) / g.rows_per_tile;
  return true;
}

__device__ __forceinline__ void tile_copy_in(float* dst, const float* __restrict__ src, int n, int tid) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < (n >> 2); i += TILE_NT) d4[i] = __ldg(s4 + i);
}
__device__ __forceinline__ void tile_copy_out(float* __restrict__ dst, const float* src, int n, int tid) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < (n >> 2); i += TILE_NT) d4[i] = s4[i];
}

enum GatherMode : int { GM_PLAIN = 0, GM_RVQ = 1, GM_BACKWARD = 2 };

// GM_PLAIN   : o1[...] = x + (q - x)                                   (out)
// GM_RVQ     : o1[...] = x - st (next residual, optional), o2 = (init ? o2 : 0) + st   (running sum)
// GM_BACKWARD: o1[...] = in2 + scale*(x - q)                           (in2 = upstream gradient, may be null)
__global__ void __launch_bounds__(TILE_NT)
gather_tile_kernel(const float* __restrict__ z, const float* __restrict__ E, const int32_t* __restrict__ idx,
                   int K, TileGeom g, int mode, float* __restrict__ o1, float* __restrict__ o2,
                   const float* __restrict__ in2, int accum_init, const float* __restrict__ g_loss, float coef,
                   double* __restrict__ sse) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                               // [TILE_ELEMS]
  float* Y = smem + TILE_ELEMS;                  // [TILE_ELEMS] second operand / second result
  __shared__ int s_off[256];
  __shared__ int s_code[256];
  const int tid = threadIdx.x;
  const int D = g.D, T = g.T;
  const float scale = (mode == GM_BACKWARD) ? __fmul_rn(g_loss ? __ldg(g_loss) : 1.0f, coef) : 0.f;
  const bool need_y_in = (mode == GM_RVQ && o2 && accum_init) || (mode == GM_BACKWARD && in2);
  float part = 0.f;
  for (long long tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const long long r0 = tile * g.rows_per_tile;
    const int rows = (int)min((long long)g.rows_per_tile, g.N - r0);
    const int n = rows * D;
    const long long e0 = r0 * D;                 // whole samples: element offset of the tile
    __syncthreads();
    tile_copy_in(X, z + e0, n, tid);
    if (need_y_in) tile_copy_in(Y, (mode == GM_BACKWARD ? in2 : o2) + e0, n, tid);
    if (tid < rows) {
      const int b = tid / T, t = tid - b * T;
      s_off[tid] = b * D * T + t;
      int k = __ldg(idx + r0 + tid);
      s_code[tid] = min(max(k, 0), K - 1);
    }
    __syncthreads();
    for (int i = tid; i < n; i += TILE_NT) {
      const int r = g.dshift >= 0 ? (i >> g.dshift) : (i / D);
      const int k = i - r * D;
      const int a = s_off[r] + k * T;
      const float x = X[a];
      const float q = __ldg(E + (size_t)s_code[r] * D + k);
      if (mode == GM_BACKWARD) {
        X[a] = fmaf(scale, __fsub_rn(x, q), need_y_in ? Y[a] : 0.f);
      } else {
        const float diff = __fsub_rn(q, x);
        const float st = __fadd_rn(x, diff);
        part = fmaf(diff, diff, part);
        if (mode == GM_PLAIN) X[a] = st;
        else {
          X[a] = __fsub_rn(x, st);
          if (o2) Y[a] = __fadd_rn(need_y_in ? Y[a] : 0.f, st);
        }
      }
    }
    __syncthreads();
    if (o1) tile_copy_out(o1 + e0, X, n, tid);
    if (mode == GM_RVQ && o2) tile_copy_out(o2 + e0, Y, n, tid);
  }
  if (sse) {
    __shared__ double red[TILE_NT / 32];
    double p = warp_sum((double)part);
    if ((tid & 31) == 0) red[tid >> 5] = p;
    __syncthreads();
    if (tid < 32) {
      double v = tid < TILE_NT / 32 ? red[tid] : 0.0;
      v = warp_sum(v);
      if (tid == 0 && v != 0.0) atomicAdd(sse, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bulk-TMA pipelined variant of gather_tile_kernel: tiles are prefetched STAGES-1 ahead with
// cp.async.bulk (completion on mbarriers), transformed in place in shared memory and written back
// with cp.async.bulk shared->global (bulk async-groups), so loads, math and stores of different
// tiles overlap inside one CTA and the SM always has >= 2 tiles of reads in flight.
// ------------------------------------------------------------------------------------------
constexpr int BULK_STAGES = 3;

template <bool TWO>      // TWO: second operand/result buffer Y (RVQ running sum, or upstream gradient)
__global__ void __launch_bounds__(TILE_NT)
gather_bulk_kernel(const float* __restrict__ z, const float* __restrict__ E, const int32_t* __restrict__ idx,
                   int K, TileGeom g, int mode, float* __restrict__ o1, float* __restrict__ o2,
                   const float* __restrict__ in2, int accum_init, const float* __restrict__ g_loss, float coef,
                   double* __restrict__ sse) {
  using namespace ptx;
  extern __shared__ __align__(128) float smem[];
  constexpr int STAGES = 2;                            // 64 KiB (TWO) / 32 KiB per CTA: 3 / 6 CTAs per SM
  constexpr int STAGE_FLOATS = (TWO ? 2 : 1) * TILE_ELEMS;
  __shared__ uint64_t full[STAGES];
  __shared__ int s_off[256];
  __shared__ int s_code[2][256];
  const int tid = threadIdx.x;
  const int D = g.D, T = g.T;
  const float scale = (mode == GM_BACKWARD) ? __fmul_rn(g_loss ? __ldg(g_loss) : 1.0f, coef) : 0.f;
  const bool y_in = TWO && ((mode == GM_RVQ && accum_init) || mode == GM_BACKWARD);
  const float* ysrc = (mode == GM_BACKWARD) ? in2 : o2;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  if (tid < g.rows_per_tile) { const int b = tid / T, t = tid - b * T; s_off[tid] = b * D * T + t; }
  __syncthreads();

  const long long my_tiles = (g.ntiles > blockIdx.x) ? (g.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_rows = [&](long long i) {
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    return (int)min((long long)g.rows_per_tile, g.N - r0);
  };
  auto issue_load = [&](long long i) {           // thread 0 only
    const int s = (int)(i % STAGES);
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    const uint32_t bytes = (uint32_t)tile_rows(i) * D * 4;
    float* X = smem + (size_t)s * STAGE_FLOATS;
    mbar_expect_tx(smem_u32(full + s), y_in ? 2 * bytes : bytes);
    bulk_g2s(smem_u32(X), z + r0 * D, bytes, smem_u32(full + s));
    if (y_in) bulk_g2s(smem_u32(X + TILE_ELEMS), ysrc + r0 * D, bytes, smem_u32(full + s));
  };
  if (tid == 0) for (long long i = 0; i < my_tiles && i < STAGES - 1; ++i) issue_load(i);
  if (my_tiles > 0 && tid < tile_rows(0)) {
    const int k = __ldg(idx + (long long)blockIdx.x * g.rows_per_tile + tid);
    s_code[0][tid] = min(max(k, 0), K - 1);
  }
  float part = 0.f;
  for (long long i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % STAGES);
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    const int rows = tile_rows(i);
    const int n = rows * D;
    float* X = smem + (size_t)s * STAGE_FLOATS;
    float* Y = X + TILE_ELEMS;
    // prefetch the codes of the next tile into registers (stored after the math)
    int next_code = 0;
    const bool have_next = (i + 1 < my_tiles) && tid < tile_rows(i + 1);
    if (have_next) next_code = __ldg(idx + (blockIdx.x + (i + 1) * gridDim.x) * (long long)g.rows_per_tile + tid);
    if (tid == 0 && i + STAGES - 1 < my_tiles) {
      bulk_wait_read<0>();                       // the store that last read stage (i-1)%STAGES has drained it
      issue_load(i + STAGES - 1);
    }
    __syncthreads();                             // s_code[i&1] written (previous iteration / prologue)
    mbar_wait(smem_u32(full + s), (uint32_t)((i / STAGES) & 1), nullptr, 0);
    const int* code = s_code[i & 1];
    // one thread = 4 consecutive dims of one row (16-byte codeword loads), U row-quads in flight per thread
    constexpr int U = 4;
    const int d4 = D >> 2, n4 = rows * d4;
    for (int e0 = tid; e0 < n4; e0 += TILE_NT * U) {
      int a[U]; float4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * TILE_NT;
        if (e < n4) {
          const int r = e / d4, k = (e - r * d4) * 4;
          a[u] = s_off[r] + k * T;
          q[u] = __ldg(reinterpret_cast<const float4*>(E + (size_t)code[r] * D + k));
        } else { a[u] = -1; q[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (a[u] < 0) continue;
        const float qv[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int ai = a[u] + c * T;
          const float x = X[ai];
          if (mode == GM_BACKWARD) {
            X[ai] = fmaf(scale, __fsub_rn(x, qv[c]), y_in ? Y[ai] : 0.f);
          } else {
            const float diff = __fsub_rn(qv[c], x);
            const float st = __fadd_rn(x, diff);
            part = fmaf(diff, diff, part);
            if (mode == GM_PLAIN) X[ai] = st;
            else {
              X[ai] = __fsub_rn(x, st);
              if (TWO) Y[ai] = __fadd_rn(y_in ? Y[ai] : 0.f, st);
            }
          }
        }
      }
    }
    if (have_next) s_code[(i + 1) & 1][tid] = min(max(next_code, 0), K - 1);
    fence_proxy_async();                         // generic-proxy smem writes -> visible to the bulk store
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)n * 4;
      if (o1) bulk_s2g(o1 + r0 * D, smem_u32(X), bytes);
      if (TWO && mode == GM_RVQ && o2) bulk_s2g(o2 + r0 * D, smem_u32(Y), bytes);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
  if (sse) {
    __shared__ double red[TILE_NT / 32];
    double p = warp_sum((double)part);
    if ((tid & 31) == 0) red[tid >> 5] = p;
    __syncthreads();
    if (tid < 32) {
      double v = tid < TILE_NT / 32 ? red[tid] : 0.0;
      v = warp_sum(v);
      if (tid == 0 && v != 0.0) atomicAdd(sse, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// RVQ output chain (models/vqvae.py:94-98 replayed for all stages at once).  Every stage's codebook is
// updated exactly once per step, so after the last stage the whole value chain
//     r_0 = z;  st_s = r_s + (E_s[idx_s] - r_s);  out = ((0 + st_0) + st_1) + ...;  r_{s+1} = r_s - st_s
// can be recomputed bit-identically from z, the indices and the final codebooks.  This kernel does that in
// one pass (read z, write out, S codeword gathers per element from L2) and also produces the S loss sums,
// which removes the per-stage read-modify-write of the running sum from the stage kernels.
// Same bulk-TMA tile pipeline as gather_bulk_kernel.
// ------------------------------------------------------------------------------------------
constexpr int CHAIN_MAX_S = 8;
constexpr int CHAIN_STAGES = 2;
struct ChainArgs {
  const float* E[CHAIN_MAX_S];
  const int32_t* idx[CHAIN_MAX_S];
  double* sse[CHAIN_MAX_S];
  int K[CHAIN_MAX_S];
  int S;
};

template <int S>
__global__ void __launch_bounds__(TILE_NT)
rvq_chain_kernel(const float* __restrict__ z, ChainArgs ca, TileGeom g, float* __restrict__ out) {
  using namespace ptx;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t full[CHAIN_STAGES];
  __shared__ int s_off[256];
  __shared__ int s_code[2][CHAIN_MAX_S][64];         // rows_per_tile <= 64 for this kernel
  const int tid = threadIdx.x;
  const int D = g.D, T = g.T;
  if (tid == 0) {
    for (int s = 0; s < CHAIN_STAGES; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  if (tid < g.rows_per_tile) { const int b = tid / T, t = tid - b * T; s_off[tid] = b * D * T + t; }
  __syncthreads();
  const long long my_tiles = (g.ntiles > blockIdx.x) ? (g.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_rows = [&](long long i) {
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    return (int)min((long long)g.rows_per_tile, g.N - r0);
  };
  auto issue_load = [&](long long i) {
    const int s = (int)(i % CHAIN_STAGES);
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    const uint32_t bytes = (uint32_t)tile_rows(i) * D * 4;
    mbar_expect_tx(smem_u32(full + s), bytes);
    bulk_g2s(smem_u32(smem + (size_t)s * TILE_ELEMS), z + r0 * D, bytes, smem_u32(full + s));
  };
  auto load_codes = [&](long long i, int buf) {      // all threads: S x rows codes of tile i
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    const int rows = tile_rows(i);
    for (int e = tid; e < S * rows; e += TILE_NT) {
      const int s = e / rows, r = e - s * rows;
      const int k = __ldg(ca.idx[s] + r0 + r);
      s_code[buf][s][r] = min(max(k, 0), ca.K[s] - 1);
    }
  };
  if (tid == 0) for (long long i = 0; i < my_tiles && i < CHAIN_STAGES - 1; ++i) issue_load(i);
  if (my_tiles > 0) load_codes(0, 0);
  float part[S];
#pragma unroll
  for (int s = 0; s < S; ++s) part[s] = 0.f;
  for (long long i = 0; i < my_tiles; ++i) {
    const int st = (int)(i % CHAIN_STAGES);
    const long long r0 = (blockIdx.x + i * gridDim.x) * g.rows_per_tile;
    const int rows = tile_rows(i);
    const int n = rows * D;
    float* X = smem + (size_t)st * TILE_ELEMS;
    if (tid == 0 && i + CHAIN_STAGES - 1 < my_tiles) {
      bulk_wait_read<0>();
      issue_load(i + CHAIN_STAGES - 1);
    }
    __syncthreads();                             // codes of this tile are in s_code[i & 1]
    if (i + 1 < my_tiles) load_codes(i + 1, (int)((i + 1) & 1));
    mbar_wait(smem_u32(full + st), (uint32_t)((i / CHAIN_STAGES) & 1), nullptr, 0);
    const int (*code)[64] = s_code[i & 1];
    // one thread = 4 consecutive dims of one row: codeword rows are fetched as 16-byte vectors (4x fewer
    // load instructions), U row-quads in flight per thread to cover the L2 latency of the S gathers
    constexpr int U = (S <= 2) ? 4 : ((S <= 4) ? 3 : 2);
    const int d4 = D >> 2, n4 = rows * d4;
    for (int e0 = tid; e0 < n4; e0 += TILE_NT * U) {
      int a[U], rr[U], kk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * TILE_NT;
        if (e < n4) {
          rr[u] = e / d4;
          kk[u] = (e - rr[u] * d4) * 4;
          a[u] = s_off[rr[u]] + kk[u] * T;
        } else { a[u] = -1; rr[u] = 0; kk[u] = 0; }
      }
      float4 q[U][S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          q[u][s] = (a[u] >= 0) ? __ldg(reinterpret_cast<const float4*>(ca.E[s] + (size_t)code[s][rr[u]] * D + kk[u]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (a[u] < 0) continue;
        float r[4] = {X[a[u]], X[a[u] + T], X[a[u] + 2 * T], X[a[u] + 3 * T]};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float qv[4] = {q[u][s].x, q[u][s].y, q[u][s].z, q[u][s].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float diff = __fsub_rn(qv[c], r[c]);
            const float stv = __fadd_rn(r[c], diff);
            part[s] = fmaf(diff, diff, part[s]);
            acc[c] = __fadd_rn(acc[c], stv);
            r[c] = __fsub_rn(r[c], stv);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) X[a[u] + c * T] = acc[c];
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(out + r0 * D, smem_u32(X), (uint32_t)n * 4);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
  __shared__ double red[TILE_NT / 32];
  for (int s = 0; s < S; ++s) {
    double p = warp_sum((double)part[s]);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = p;
    __syncthreads();
    if (tid < 32) {
      double v = tid < TILE_NT / 32 ? red[tid] : 0.0;
      v = warp_sum(v);
      if (tid == 0 && v != 0.0) atomicAdd(ca.sse[s], v);
    }
  }
}

// ema_accumulate on a tile: cnt via smem histogram, dw via 16-byte vector reductions
__global__ void __launch_bounds__(TILE_NT)
accumulate_tile_kernel(const float* __restrict__ z, const int32_t* __restrict__ idx, const float* __restrict__ E,
                       int K, TileGeom g, float* __restrict__ dw, float* __restrict__ cnt, int mode, int use_hist) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;
  int* hist = reinterpret_cast<int*>(smem + ACC_TILE_ELEMS);
  __shared__ int s_off[256];
  __shared__ int s_code[256];
  const int tid = threadIdx.x;
  const int D = g.D, T = g.T, d4 = D >> 2;
  if (use_hist) for (int k = tid; k < K; k += TILE_NT) hist[k] = 0;
  for (long long tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const long long r0 = tile * g.rows_per_tile;
    const int rows = (int)min((long long)g.rows_per_tile, g.N - r0);
    const int n = rows * D;
    __syncthreads();
    tile_copy_in(X, z + r0 * D, n, tid);
    if (tid < rows) {
      const int b = tid / T, t = tid - b * T;
      s_off[tid] = b * D * T + t;
      const int k = __ldg(idx + r0 + tid);
      s_code[tid] = k;
      if ((unsigned)k < (unsigned)K) { if (use_hist) atomicAdd(hist + k, 1); else atomicAdd(cnt + k, 1.0f); }
    }
    __syncthreads();
    for (int i = tid; i < rows * d4; i += TILE_NT) {
      const int r = i / d4, q = i - r * d4;
      const int k = s_code[r];
      if ((unsigned)k >= (unsigned)K) continue;
      const int a = s_off[r] + 4 * q * T;
      float4 v = make_float4(X[a], X[a + T], X[a + 2 * T], X[a + 3 * T]);
      if (mode == 1) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + q);
        v = make_float4(e.x - v.x, e.y - v.y, e.z - v.z, e.w - v.w);
      }
      red_add_v4(dw + (size_t)k * D + 4 * q, v.x, v.y, v.z, v.w);
    }
  }
  if (use_hist) {
    __syncthreads();
    for (int k = tid; k < K; k += TILE_NT) { const int h = hist[k]; if (h) atomicAdd(cnt + k, (float)h); }
  }
}

static int tile_grid(const TileGeom& g, int per_sm) {
  return (int)max(1LL, min(g.ntiles, (long long)sm_count() * per_sm));
}

// returns 1 if handled, 0 if the layout is not eligible (caller falls back), < 0 / > 0 on error
int try_gather_tile(const ZView& z, const float* E, const int32_t* idx, int K, int mode, float* o1, float* o2,
                    const float* in2, int accum_init, const float* g_loss, float coef, double* sse,
                    cudaStream_t stream) {
  TileGeom g;
  if (!tile_geom(z, g)) return 0;
  if ((reinterpret_cast<uintptr_t>(o1) & 15) || (reinterpret_cast<uintptr_t>(o2) & 15) ||
      (reinterpret_cast<uintptr_t>(in2) & 15) || (reinterpret_cast<uintptr_t>(E) & 15)) return 0;
  const bool two = (mode == GM_RVQ && o2) || (mode == GM_BACKWARD && in2);
  static const bool use_bulk = !(getenv("VQB200_NO_BULK") && atoi(getenv("VQB200_NO_BULK")));
  if (use_bulk && (g.rows_per_tile * g.D * 4) % 16 == 0) {
    const size_t smem = (size_t)2 * (two ? 2 : 1) * TILE_ELEMS * sizeof(float);
    static PerDevice configured_;
    std::atomic<size_t>& configured = configured_.here();
    if (!configured.load()) {
      cudaError_t e1 = cudaFuncSetAttribute(gather_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            2 * 2 * TILE_ELEMS * (int)sizeof(float));
      cudaError_t e2 = cudaFuncSetAttribute(gather_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            CHAIN_STAGES * TILE_ELEMS * (int)sizeof(float));
      if (e1 != cudaSuccess || e2 != cudaSuccess) return cuda_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute(gather_bulk_kernel)");
      configured.store(1);
    }
    const int grid = tile_grid(g, two ? 3 : 6);
    if (two) gather_bulk_kernel<true><<<grid, TILE_NT, smem, stream>>>(z.p, E, idx, K, g, mode, o1, o2, in2, accum_init, g_loss, coef, sse);
    else gather_bulk_kernel<false><<<grid, TILE_NT, smem, stream>>>(z.p, E, idx, K, g, mode, o1, o2, in2, accum_init, g_loss, coef, sse);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gather_bulk_kernel");
    return 1;
  }
  const size_t smem = (two ? 2 : 1) * TILE_ELEMS * sizeof(float);
  gather_tile_kernel<<<tile_grid(g, 8), TILE_NT, smem, stream>>>(z.p, E, idx, K, g, mode, o1, o2, in2, accum_init,
                                                                g_loss, coef, sse);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gather_tile_kernel");
  return 1;
}

// returns 1 if handled, 0 if the layout is not eligible
int try_rvq_chain(const ZView& z, int S, const float* const* E, const int32_t* const* idx, const int* K,
                  double* const* sse, float* out, cudaStream_t stream) {
  TileGeom g;
  if (S < 1 || S > CHAIN_MAX_S || !tile_geom(z, g)) return 0;
  if (g.rows_per_tile > 64 || (g.rows_per_tile * g.D * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return 0;
  ChainArgs ca;
  ca.S = S;
  for (int s = 0; s < S; ++s) {
    if (reinterpret_cast<uintptr_t>(E[s]) & 15) return 0;
    ca.E[s] = E[s]; ca.idx[s] = idx[s]; ca.K[s] = K[s]; ca.sse[s] = sse[s];
  }
  for (int s = S; s < CHAIN_MAX_S; ++s) { ca.E[s] = nullptr; ca.idx[s] = nullptr; ca.K[s] = 1; ca.sse[s] = nullptr; }
  const size_t smem = (size_t)CHAIN_STAGES * TILE_ELEMS * sizeof(float);
  const int grid = tile_grid(g, 6);
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, TILE_NT, smem, stream>>>(z.p, ca, g, out);
    return cudaSuccess;
  };
  cudaError_t le = cudaSuccess;
  switch (S) {
    case 1: le = launch(rvq_chain_kernel<1>); break;
    case 2: le = launch(rvq_chain_kernel<2>); break;
    case 3: le = launch(rvq_chain_kernel<3>); break;
    case 4: le = launch(rvq_chain_kernel<4>); break;
    case 5: le = launch(rvq_chain_kernel<5>); break;
    case 6: le = launch(rvq_chain_kernel<6>); break;
    case 7: le = launch(rvq_chain_kernel<7>); break;
    default: le = launch(rvq_chain_kernel<8>); break;
  }
  if (le != cudaSuccess) return cuda_fail(le, "rvq_chain_kernel setup");
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rvq_chain_kernel");
  return 1;
}

int try_accumulate_tile(const ZView& z, const int32_t* idx, const float* E, int K, float* dw, float* cnt, int mode,
                        cudaStream_t stream) {
  TileGeom g;
  if (!tile_geom(z, g, ACC_TILE_ELEMS)) return 0;
  if ((reinterpret_cast<uintptr_t>(dw) & 15) || (mode == 1 && (reinterpret_cast<uintptr_t>(E) & 15))) return 0;
  const int use_hist = K <= 8192 ? 1 : 0;
  const size_t smem = ACC_TILE_ELEMS * sizeof(float) + (use_hist ? (size_t)K * sizeof(int) : 0);
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (smem > configured.load()) {
    cudaError_t e = cudaFuncSetAttribute(accumulate_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(accumulate_tile_kernel)");
    configured.store(smem);
  }
  accumulate_tile_kernel<<<tile_grid(g, 4), TILE_NT, smem, stream>>>(z.p, idx, E, K, g, dw, cnt, mode, use_hist);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "accumulate_tile_kernel");
  return 1;
}

}  // namespace vqb200
