// vqb200 K5f: FSQ / LFQ with their 1x1 projections fused in.
//
// Replaces, in ONE pass over z, models/vqvae.py:126-154 (FSQ.forward: project_in -> round -> index/metrics ->
// project_out) and :170-194 (LFQ.forward: project_in -> sign, entropy loss, index/metrics -> project_out), and in
// one more pass their autograd backward (input gradient + the four projection parameter gradients).  The unfused
// path moves z_e / z_q through HBM twice more and runs three launches per direction; here the algorithmic bytes are
// 8D + 4d + 8 per vector forward (read z, write out, write z_e and the int64 index) and 12D + 4d backward.
//
// Layout: contiguous [B, 64, T]; a tile is a run of whole samples (<= 128 rows) = one contiguous byte range streamed
// by bulk-TMA (cp.async.bulk + mbarrier in, cp.async.bulk shared->global out).  TWO THREADS PER ROW (b,t), adjacent
// lanes, 32 channels each: the D -> d projection is a thread-local dot product plus ONE shuffle, the d -> D projection
// is thread-local; the projection weights are broadcast from shared memory
// 16 bytes at a time ([channel][DQ] layout).  A thread walks its row's channels ROTATED by its sample index inside the
// tile (channel (c + bl) mod 64 at step c): element (row r, channel c) of the [B,C,T] tile sits at word
// bl*64*T + t + c*T, so the 32 rows of a warp then touch banks r + c*T -- conflict-free for every T, where the
// straight walk is 4-way conflicted at T = 10 and 32-way at T = 1.  (The summation order of a projection therefore
// depends on the row's position in the tile; the results differ in the last bit only.)  The backward kernel adds a
// second phase with one thread per channel that forms the parameter-gradient outer products over the tile's rows.
#include "common.cuh"
#include "ptx.cuh"
#include "uniq.cuh"

namespace vqb200 {

constexpr int F_D = 64;                   // channel count the fused kernels are built for
constexpr int F_ROWS = 128;               // rows per tile; two threads per row (even / odd steps of the channel walk)
constexpr int F_TILE = F_ROWS * F_D;      // 8192 floats = 32 KiB
constexpr int F_NT = 2 * F_ROWS;
constexpr int F_STAGES_FWD = 2;           // forward: two tiles in flight per CTA; backward: one (two operands per tile)
constexpr int F_MAX_DQ = 16;

struct FusedParams {
  const float* z;            // [B,64,T] input
  const float* g_out;        // backward: upstream gradient wrt `out`
  float* out;                // forward: projected output; backward: gradient wrt z
  float* z_e;                // [B,d,T] pre-quantisation latent (written forward, read backward)
  long long* idx;            // [B,T]
  const float* W_in; const float* b_in;     // [d,64], [d]
  const float* W_out; const float* b_out;   // [64,d], [64]
  const int32_t* basis;      // FSQ
  double codebook_size;      // FSQ
  float weight;              // LFQ entropy loss weight
  const float* g_loss;       // LFQ backward: upstream gradient of the loss (device scalar, may be null = 1)
  float* grads;              // backward: [d*64 | d | 64*d | 64] = dW_in, db_in, dW_out, db_out (zeroed by the caller)
  long long B; int d, T;
  int samples_per_tile; long long ntiles;
  void* ws; float* outm;
};

__device__ __forceinline__ float fsq_round_st(float z) { return __fadd_rn(z, __fsub_rn(rintf(z), z)); }   // :130-131
__device__ __forceinline__ float lfq_sign_st(float z) { return __fadd_rn(z, __fsub_rn((z > 0.f) ? 1.f : -1.f, z)); }   // :172-174

// shared-memory copies of the projection weights, [channel][DQ] so that one 16-byte broadcast load feeds 4 FMAs
template <int DQ>
__device__ __forceinline__ void stage_weights(const FusedParams& p, float* sWin, float* sWout, int tid) {
  for (int i = tid; i < F_D * DQ; i += F_NT) {
    const int c = i / DQ, j = i - c * DQ;
    sWin[i] = (j < p.d) ? __ldg(p.W_in + j * F_D + c) : 0.f;       // W_in[j][c]
    sWout[i] = (j < p.d) ? __ldg(p.W_out + c * p.d + j) : 0.f;     // W_out[c][j]
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <bool IS_LFQ, int DQ>
__global__ void __launch_bounds__(F_NT, 2)
fused_forward_kernel(const FusedParams p) {
  using namespace ptx;
  extern __shared__ __align__(128) float smem[];      // [F_STAGES_FWD][F_TILE]
  __shared__ uint64_t full[F_STAGES_FWD];
  __shared__ unsigned lbm[Q_LOCAL_WORDS];
  __shared__ __align__(16) float sWin[F_D * DQ], sWout[F_D * DQ];
  __shared__ float sBout[F_D];
  const UniqWs w(p.ws);
  const int tid = threadIdx.x;
  const int d = p.d, T = p.T;
  const int slab = F_D * T;

  stage_weights<DQ>(p, sWin, sWout, tid);
  if (tid < F_D) sBout[tid] = __ldg(p.b_out + tid);
  float bin[DQ], fb[DQ];
#pragma unroll
  for (int j = 0; j < DQ; ++j) {
    bin[j] = (j < d) ? __ldg(p.b_in + j) : 0.f;
    fb[j] = (!IS_LFQ && j < d) ? (float)__ldg(p.basis + j) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < F_STAGES_FWD; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  for (int i = tid; i < Q_LOCAL_WORDS; i += F_NT) lbm[i] = 0u;
  const int row = tid >> 1, h = tid & 1;          // this thread's row inside every tile, and its half of the walk
  const int bl = row / T, t = row - bl * T;
  __syncthreads();

  const long long my_tiles = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_samples = [&](long long i) {
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    return (int)min((long long)p.samples_per_tile, p.B - b0);
  };
  auto issue_load = [&](long long i) {           // thread 0 only
    const int s = (int)(i % F_STAGES_FWD);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const uint32_t bytes = (uint32_t)tile_samples(i) * slab * 4;
    mbar_expect_tx(smem_u32(full + s), bytes);
    bulk_g2s(smem_u32(smem + (size_t)s * F_TILE), p.z + b0 * slab, bytes, smem_u32(full + s));
  };
  if (tid == 0) for (long long i = 0; i < my_tiles && i < F_STAGES_FWD - 1; ++i) issue_load(i);

  float ent = 0.f;
  for (long long i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % F_STAGES_FWD);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const int ns = tile_samples(i);
    const int rows = ns * T;
    float* X = smem + (size_t)s * F_TILE;
    if (tid == 0 && i + F_STAGES_FWD - 1 < my_tiles) {
      bulk_wait_read<0>();                       // the store that last read the other stage has drained it
      issue_load(i + F_STAGES_FWD - 1);
    }
    mbar_wait(smem_u32(full + s), (uint32_t)((i / F_STAGES_FWD) & 1), nullptr, 0);
    {
      const bool active = row < rows;             // warp-uniform up to the last partial pair: the shuffle needs every lane
      float* px = X + (active ? bl * slab + t : 0);
      float ze[DQ], zh[DQ];
#pragma unroll
      for (int j = 0; j < DQ; ++j) ze[j] = 0.f;
#pragma unroll
      for (int i = 0; i < F_D / 2; ++i) {         // z_e = W_in z (+ b_in below): this thread's 32 channels
        const int cr = (2 * i + h + bl) & (F_D - 1);
        const float x = active ? px[cr * T] : 0.f;
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(sWin + cr * DQ + j4 * 4);
          ze[j4 * 4 + 0] = fmaf(wv.x, x, ze[j4 * 4 + 0]); ze[j4 * 4 + 1] = fmaf(wv.y, x, ze[j4 * 4 + 1]);
          ze[j4 * 4 + 2] = fmaf(wv.z, x, ze[j4 * 4 + 2]); ze[j4 * 4 + 3] = fmaf(wv.w, x, ze[j4 * 4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        ze[j] += __shfl_xor_sync(0xffffffffu, ze[j], 1);          // the partner's 32 channels (bit-identical in both)
        ze[j] += bin[j];
        zh[j] = IS_LFQ ? lfq_sign_st(ze[j]) : fsq_round_st(ze[j]);
      }
      if (active) {
#pragma unroll
      for (int i = 0; i < F_D / 2; ++i) {         // out = W_out z_q + b_out, in place over the input tile
        const int cr = (2 * i + h + bl) & (F_D - 1);
        float o = sBout[cr];
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(sWout + cr * DQ + j4 * 4);
          o = fmaf(wv.x, zh[j4 * 4 + 0], o); o = fmaf(wv.y, zh[j4 * 4 + 1], o);
          o = fmaf(wv.z, zh[j4 * 4 + 2], o); o = fmaf(wv.w, zh[j4 * 4 + 3], o);
        }
        px[cr * T] = o;
      }
      const long long b = b0 + bl;
      long long code = 0;
      float sidx = 0.f;
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        if (j < d) {
          if ((j & 1) == h) {                     // the pair splits the per-component work
            p.z_e[(b * d + j) * T + t] = ze[j];
            if (IS_LFQ) {
              // entropy term: only its mean enters the loss (1e-5 tolerance) -> MUFU-based fast intrinsics
              const float pr = __fdividef(1.f, 1.f + __expf(-ze[j]));
              const float q = 1.f - pr;
              ent -= fmaf(pr, __logf(pr + 1e-6f), q * __logf(q + 1e-6f));
            }
          }
          if (IS_LFQ) {
            if (zh[j] > 0.f) code |= (1LL << j);
          } else {
            const float pj = __fmul_rn(zh[j], fb[j]);                 // :135 float multiply-sum, then truncate
            sidx = (j == 0) ? pj : __fadd_rn(sidx, pj);
          }
        }
      }
      if (!IS_LFQ) code = trunc_to_i64(sidx);
      if (h == 0) {
        p.idx[b * T + t] = code;
        if (code >= -Q_LOCAL_HALF && code < Q_LOCAL_HALF) {
          const unsigned bit = (unsigned)(code + Q_LOCAL_HALF);
          const unsigned m = 1u << (bit & 31);
          if (!(lbm[bit >> 5] & m)) atomicOr(&lbm[bit >> 5], m);
        } else {
          unique_insert(w, code);
        }
      }
      }
    }
    fence_proxy_async();                         // generic-proxy smem writes -> visible to the bulk store
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(p.out + b0 * slab, smem_u32(X), (uint32_t)ns * slab * 4);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();

  // ---- metrics: merge the CTA-local bitmap, reduce the entropy, last CTA finalises ----
  __syncthreads();
  {
    unsigned* gbm = w.bitmap + (unsigned)((UNIQ_HALF - Q_LOCAL_HALF) >> 5);
    for (int i = tid; i < Q_LOCAL_WORDS; i += F_NT) { const unsigned v = lbm[i]; if (v) atomicOr(gbm + i, v); }
  }
  if (IS_LFQ) {
    __shared__ double red[F_NT / 32];
    double pe = warp_sum((double)ent);
    if ((tid & 31) == 0) red[tid >> 5] = pe;
    __syncthreads();
    if (tid < 32) {
      double v = tid < F_NT / 32 ? red[tid] : 0.0;
      v = warp_sum(v);
      if (tid == 0) atomicAdd(w.ent, v);
    }
  }
  __shared__ bool s_is_last;
  __shared__ unsigned s_local_total;
  __threadfence();
  __syncthreads();
  if (tid == 0) { s_is_last = (atomicAdd(w.ticket, 1u) == gridDim.x - 1); s_local_total = 0u; }
  __syncthreads();
  if (s_is_last) {
    __threadfence();
    const volatile unsigned* gbm = w.bitmap + (unsigned)((UNIQ_HALF - Q_LOCAL_HALF) >> 5);
    unsigned c = 0;
    for (int i = tid; i < Q_LOCAL_WORDS; i += F_NT) c += __popc(gbm[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0 && c) atomicAdd(&s_local_total, c);
  }
  __syncthreads();
  if (s_is_last && tid == 0) {
    const unsigned u = atomicAdd(w.count, 0u) + s_local_total;
    const bool ovf = atomicAdd(w.overflow, 0u) != 0u;
    if (IS_LFQ) {
      const double sum = atomicAdd(w.ent, 0.0);
      const float mean = (float)(sum / ((double)p.B * T * d));
      p.outm[0] = __fmul_rn(-mean, p.weight);
      p.outm[1] = ovf ? NAN : (float)u;
      p.outm[2] = ovf ? NAN : (float)(1.0 - (double)u / exp2((double)d));
    } else {
      p.outm[0] = ovf ? NAN : (float)u;
      p.outm[1] = ovf ? NAN : (float)(1.0 - (double)u / p.codebook_size);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: g_z = W_in^T g_ze,  g_ze = W_out^T g_out (+ LFQ entropy term);  dW_out = sum g_out (x) z_q,
// db_out = sum g_out,  dW_in = sum g_ze (x) z,  db_in = sum g_ze.  (autograd of :126-154 / :170-194; the rounding
// and the sign are straight-through)
//   phase A1 (thread = row):      g_ze, z_q of the row -> shared memory
//   phase B  (thread = channel):  outer products over the tile's rows into register accumulators
//   phase A2 (thread = row):      g_z = W_in^T g_ze, in place over the g_out tile -> bulk store
// ------------------------------------------------------------------------------------------
template <bool IS_LFQ, int DQ>
__global__ void __launch_bounds__(F_NT, 2)
fused_backward_kernel(const FusedParams p) {
  using namespace ptx;
  extern __shared__ __align__(128) float smem[];      // [g_out tile | z tile]
  __shared__ uint64_t full;
  __shared__ __align__(16) float sWin[F_D * DQ], sWout[F_D * DQ];
  __shared__ __align__(16) float sGze[F_ROWS * DQ], sZq[F_ROWS * DQ];
  __shared__ int sOff[F_ROWS];
  __shared__ __align__(16) float sZe[F_ROWS * F_MAX_DQ];   // z_e slab of the tile ([sample][d][T], as in global memory)
  __shared__ float sBin[F_MAX_DQ];
  float* G = smem;
  const float* X = smem + F_TILE;
  const int tid = threadIdx.x;
  const int d = p.d, T = p.T;
  const int slab = F_D * T;

  stage_weights<DQ>(p, sWin, sWout, tid);
  const int row = tid >> 1, h = tid & 1;
  const int bl = row / T, t = row - bl * T;
  if (h == 0) sOff[row] = bl * slab + t;
  if (tid < F_MAX_DQ) sBin[tid] = 0.f;
  const float lscale = IS_LFQ ? (p.g_loss ? __ldg(p.g_loss) : 1.f) * (-p.weight / (float)((double)p.B * T * d)) : 0.f;
  if (tid == 0) { mbar_init(smem_u32(&full), 1); fence_barrier_init(); }
  __syncthreads();

  // phase B role: channel cB, rows [hB*32, hB*32+32)
  const int cB = tid & (F_D - 1), hB = tid >> 6;
  float a_win[DQ], a_wout[DQ], a_bout = 0.f, a_bin[DQ];
#pragma unroll
  for (int j = 0; j < DQ; ++j) { a_win[j] = 0.f; a_wout[j] = 0.f; a_bin[j] = 0.f; }

  const long long my_tiles = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (long long i = 0; i < my_tiles; ++i) {
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const int ns = (int)min((long long)p.samples_per_tile, p.B - b0);
    const int rows = ns * T;
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)ns * slab * 4;
      bulk_wait_read<0>();                       // the previous tile's store has drained the buffer
      mbar_expect_tx(smem_u32(&full), 2 * bytes);
      bulk_g2s(smem_u32(G), p.g_out + b0 * slab, bytes, smem_u32(&full));
      bulk_g2s(smem_u32(G + F_TILE), p.z + b0 * slab, bytes, smem_u32(&full));
    }
    // the tile's z_e slab (any size / alignment): plain loads that overlap the bulk copies above
    for (int e = tid; e < ns * d * T; e += F_NT) sZe[e] = __ldg(p.z_e + b0 * d * T + e);
    mbar_wait(smem_u32(&full), (uint32_t)(i & 1), nullptr, 0);
    __syncthreads();
    // ---- A1 ----
    float gze[DQ];
#pragma unroll
    for (int j = 0; j < DQ; ++j) gze[j] = 0.f;
    const bool active = row < rows;
    {
      const float* pg = G + (active ? bl * slab + t : 0);
#pragma unroll
      for (int i = 0; i < F_D / 2; ++i) {
        const int cr = (2 * i + h + bl) & (F_D - 1);
        const float g = active ? pg[cr * T] : 0.f;
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(sWout + cr * DQ + j4 * 4);
          gze[j4 * 4 + 0] = fmaf(wv.x, g, gze[j4 * 4 + 0]); gze[j4 * 4 + 1] = fmaf(wv.y, g, gze[j4 * 4 + 1]);
          gze[j4 * 4 + 2] = fmaf(wv.z, g, gze[j4 * 4 + 2]); gze[j4 * 4 + 3] = fmaf(wv.w, g, gze[j4 * 4 + 3]);
        }
      }
      const float* pze = sZe + (active ? bl : 0) * d * T + t;
      float zq[DQ];
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        const float ze = (active && j < d) ? pze[j * T] : 0.f;
        zq[j] = (j < d) ? (IS_LFQ ? lfq_sign_st(ze) : fsq_round_st(ze)) : 0.f;
        if (IS_LFQ && j < d && (j & 1) == h) {    // d(-w * mean H_b(sigmoid(z_e)))/dz_e, SURVEY row a14 (the pair splits the components)
          // MUFU-based intrinsics: the term is scaled by w/M and added to the straight-through gradient, its ~1e-6
          // relative error is far inside the 1e-5 tolerance
          const float dl = 1e-6f;
          const float pr = __fdividef(1.f, 1.f + __expf(-ze));
          const float q = 1.f - pr;
          const float dH = -(__logf(pr + dl) + __fdividef(pr, pr + dl) - __logf(q + dl) - __fdividef(q, q + dl));
          gze[j] = fmaf(lscale * dH, pr * q, gze[j]);
        }
        gze[j] += __shfl_xor_sync(0xffffffffu, gze[j], 1);        // the partner's 32 channels and its entropy terms
        if (j >= d || !active) gze[j] = 0.f;
        if (h == 0) a_bin[j] += gze[j];
      }
      if (active && h == 0) {
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          *reinterpret_cast<float4*>(sGze + row * DQ + j4 * 4) = make_float4(gze[j4 * 4], gze[j4 * 4 + 1], gze[j4 * 4 + 2], gze[j4 * 4 + 3]);
          *reinterpret_cast<float4*>(sZq + row * DQ + j4 * 4) = make_float4(zq[j4 * 4], zq[j4 * 4 + 1], zq[j4 * 4 + 2], zq[j4 * 4 + 3]);
        }
      }
    }
    __syncthreads();
    // ---- B ----
    {
      const int r_end = min(rows, hB * 32 + 32);
      for (int r = hB * 32; r < r_end; ++r) {
        const int off = sOff[r] + cB * T;
        const float gv = G[off], xv = X[off];
        a_bout += gv;
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          const float4 gz = *reinterpret_cast<const float4*>(sGze + r * DQ + j4 * 4);
          const float4 zq = *reinterpret_cast<const float4*>(sZq + r * DQ + j4 * 4);
          a_win[j4 * 4 + 0] = fmaf(gz.x, xv, a_win[j4 * 4 + 0]); a_win[j4 * 4 + 1] = fmaf(gz.y, xv, a_win[j4 * 4 + 1]);
          a_win[j4 * 4 + 2] = fmaf(gz.z, xv, a_win[j4 * 4 + 2]); a_win[j4 * 4 + 3] = fmaf(gz.w, xv, a_win[j4 * 4 + 3]);
          a_wout[j4 * 4 + 0] = fmaf(gv, zq.x, a_wout[j4 * 4 + 0]); a_wout[j4 * 4 + 1] = fmaf(gv, zq.y, a_wout[j4 * 4 + 1]);
          a_wout[j4 * 4 + 2] = fmaf(gv, zq.z, a_wout[j4 * 4 + 2]); a_wout[j4 * 4 + 3] = fmaf(gv, zq.w, a_wout[j4 * 4 + 3]);
        }
      }
    }
    __syncthreads();
    // ---- A2 ----
    if (active) {
      float* pg = G + bl * slab + t;
#pragma unroll
      for (int i = 0; i < F_D / 2; ++i) {
        const int cr = (2 * i + h + bl) & (F_D - 1);
        float o = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < DQ / 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(sWin + cr * DQ + j4 * 4);
          o = fmaf(wv.x, gze[j4 * 4 + 0], o); o = fmaf(wv.y, gze[j4 * 4 + 1], o);
          o = fmaf(wv.z, gze[j4 * 4 + 2], o); o = fmaf(wv.w, gze[j4 * 4 + 3], o);
        }
        pg[cr * T] = o;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(p.out + b0 * slab, smem_u32(G), (uint32_t)ns * slab * 4);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();

  // ---- parameter gradients: registers -> global (one atomic per element per thread; four threads per channel) ----
  float* g_win = p.grads;                    // [d][64]
  float* g_bin = g_win + d * F_D;            // [d]
  float* g_wout = g_bin + d;                 // [64][d]
  float* g_bout = g_wout + F_D * d;          // [64]
  if (my_tiles > 0) {
    atomicAdd(g_bout + cB, a_bout);
#pragma unroll
    for (int j = 0; j < DQ; ++j) {
      if (j < d) {
        atomicAdd(g_win + j * F_D + cB, a_win[j]);
        atomicAdd(g_wout + cB * d + j, a_wout[j]);
        atomicAdd(sBin + j, a_bin[j]);
      }
    }
  }
  __syncthreads();
  if (tid < d && my_tiles > 0) atomicAdd(g_bin + tid, sBin[tid]);
}

static bool fused_geom(const void* a, const void* b, int64_t B, int64_t D, int64_t d, int64_t T, FusedParams& p) {
  if (D != F_D || d < 1 || d > F_MAX_DQ || T < 1 || B < 1) return false;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return false;
  const long long slab = D * T;
  if (slab > F_TILE) return false;
  p.samples_per_tile = (int)(F_TILE / slab);
  p.ntiles = (B + p.samples_per_tile - 1) / p.samples_per_tile;
  p.B = B; p.d = (int)d; p.T = (int)T;
  return true;
}

template <bool IS_LFQ, bool BWD>
static int launch_fused(const FusedParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)(BWD ? 2 : F_STAGES_FWD) * F_TILE * sizeof(float);
#define VQ_FUSED_CASE(DQ)                                                                                             \
  do {                                                                                                                \
    auto kern = BWD ? fused_backward_kernel<IS_LFQ, DQ> : fused_forward_kernel<IS_LFQ, DQ>;                           \
    static PerDevice occ_;                       /* resident CTAs per SM of this instantiation, per device (0 = unknown) */ \
    std::atomic<size_t>& occ = occ_.here();                                                                           \
    int per_sm_now = (int)occ.load();                                                                                 \
    if (per_sm_now == 0) {                                                                                            \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fsq_lfq_fused)");                               \
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_now, kern, F_NT, smem);                               \
      if (e != cudaSuccess || per_sm_now < 1) return cuda_fail(e, "occupancy(fsq_lfq_fused)");                        \
      occ.store((size_t)per_sm_now);                                                                                  \
    }                                                                                                                 \
    const int grid_now = (int)max(1LL, min(p.ntiles, (long long)sm_count() * per_sm_now));                            \
    kern<<<grid_now, F_NT, smem, stream>>>(p);                                                                            \
  } while (0)
  if (p.d <= 4) VQ_FUSED_CASE(4);
  else if (p.d <= 8) VQ_FUSED_CASE(8);
  else if (p.d <= 12) VQ_FUSED_CASE(12);
  else VQ_FUSED_CASE(16);
#undef VQ_FUSED_CASE
  VQ_LAUNCH_CHECK("fsq_lfq_fused_kernel");
  return VQB200_OK;
}

}  // namespace vqb200

using namespace vqb200;

extern "C" {

int vqb200_proj_fused_eligible(int64_t B, int64_t D, int64_t d, int64_t T) {
  FusedParams p;
  return fused_geom(nullptr, nullptr, B, D, d, T, p) ? 1 : 0;
}

size_t vqb200_proj_fused_grad_floats(int64_t D, int64_t d) { return (size_t)(2 * D * d + D + d); }

int vqb200_fsq_fused_forward(const float* z, int64_t B, int64_t D, int64_t T, const float* W_in, const float* b_in,
                             const float* W_out, const float* b_out, int64_t d, const int32_t* basis,
                             int64_t codebook_size, float* out, float* z_e, int64_t* idx, void* workspace,
                             float* out2, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(z && W_in && b_in && W_out && b_out && basis && out && z_e && idx && workspace && out2, VQB200_EINVAL,
               "fsq_fused_forward: null pointer");
  VQ_CHECK_ARG(codebook_size > 0, VQB200_EINVAL, "fsq_fused_forward: codebook_size must be positive");
  FusedParams p = {};
  VQ_CHECK_ARG(fused_geom(z, out, B, D, d, T, p), VQB200_EUNSUPPORTED,
               "fsq_fused_forward: needs contiguous 16-byte aligned [B,64,T] with T <= 128 and d <= 16");
  p.z = z; p.out = out; p.z_e = z_e; p.idx = (long long*)idx; p.W_in = W_in; p.b_in = b_in; p.W_out = W_out; p.b_out = b_out;
  p.basis = basis; p.codebook_size = (double)codebook_size; p.ws = workspace; p.outm = out2;
  VQ_CUDA(cudaMemsetAsync(workspace, 0, UNIQ_WS_BYTES, stream));
  return launch_fused<false, false>(p, stream);
}

int vqb200_lfq_fused_forward(const float* z, int64_t B, int64_t D, int64_t T, const float* W_in, const float* b_in,
                             const float* W_out, const float* b_out, int64_t d, float entropy_loss_weight,
                             float* out, float* z_e, int64_t* idx, void* workspace, float* out3,
                             vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(z && W_in && b_in && W_out && b_out && out && z_e && idx && workspace && out3, VQB200_EINVAL,
               "lfq_fused_forward: null pointer");
  FusedParams p = {};
  VQ_CHECK_ARG(fused_geom(z, out, B, D, d, T, p), VQB200_EUNSUPPORTED,
               "lfq_fused_forward: needs contiguous 16-byte aligned [B,64,T] with T <= 128 and d <= 16");
  p.z = z; p.out = out; p.z_e = z_e; p.idx = (long long*)idx; p.W_in = W_in; p.b_in = b_in; p.W_out = W_out; p.b_out = b_out;
  p.weight = entropy_loss_weight; p.ws = workspace; p.outm = out3;
  VQ_CUDA(cudaMemsetAsync(workspace, 0, UNIQ_WS_BYTES, stream));
  return launch_fused<true, false>(p, stream);
}

int vqb200_proj_fused_backward(int is_lfq, const float* g_out, const float* z, const float* z_e, int64_t B, int64_t D,
                               int64_t T, const float* W_in, const float* W_out, int64_t d, const float* g_loss,
                               float entropy_loss_weight, float* g_z, float* grads, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(g_out && z && z_e && W_in && W_out && g_z && grads, VQB200_EINVAL, "proj_fused_backward: null pointer");
  FusedParams p = {};
  VQ_CHECK_ARG(fused_geom(z, g_z, B, D, d, T, p) && (reinterpret_cast<uintptr_t>(g_out) & 15) == 0, VQB200_EUNSUPPORTED,
               "proj_fused_backward: needs contiguous 16-byte aligned [B,64,T] with T <= 128 and d <= 16");
  p.z = z; p.g_out = g_out; p.out = g_z; p.z_e = const_cast<float*>(z_e); p.W_in = W_in; p.W_out = W_out;
  p.g_loss = g_loss; p.weight = entropy_loss_weight; p.grads = grads;
  VQ_CUDA(cudaMemsetAsync(grads, 0, vqb200_proj_fused_grad_floats(D, d) * sizeof(float), stream));
  return is_lfq ? launch_fused<true, true>(p, stream) : launch_fused<false, true>(p, stream);
}

}  // extern "C"
