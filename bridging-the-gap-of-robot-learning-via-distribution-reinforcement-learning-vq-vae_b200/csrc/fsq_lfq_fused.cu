// vqb200 K5f: FSQ / LFQ with their 1x1 projections fused in.
//
// Replaces, in ONE pass over z, models/vqvae.py:126-154 (FSQ.forward: project_in -> round -> index/metrics ->
// project_out) and :170-194 (LFQ.forward: project_in -> sign, entropy loss, index/metrics -> project_out), and in
// one more pass their autograd backward (input gradient + the four projection parameter gradients).  The unfused
// path moves z_e / z_q through HBM twice more and runs three launches per direction; here the algorithmic bytes are
// 8D + 4d + 8 per vector forward (read z, write out, write z_e and the int64 index) and 12D + 4d backward.
//
// Layout: contiguous [B, 64, T]; a tile is a run of whole samples = one contiguous byte range streamed by bulk-TMA
// (cp.async.bulk + mbarrier in, cp.async.bulk shared->global out), two stages per CTA.  A row (b,t) is handled by
// LPR = 64/CPT lanes of one warp (lane kq owns channels kq, kq+LPR, ...: a stride-T walk over shared memory that is
// bank-conflict free for odd T and T = 2*odd): the D -> d projection is a CPT-term partial dot product
// per lane plus an xor-butterfly over the lanes of the row (every lane ends with bit-identical sums), the d -> D
// projection is thread-local.  Projection weights live in registers.
#include "common.cuh"
#include "ptx.cuh"
#include "uniq.cuh"

namespace vqb200 {

constexpr int F_D = 64;                   // channel count the fused kernels are built for
constexpr int F_TILE_FWD = 8192;          // forward: 32 KiB tiles (128 rows), 2 stages
constexpr int F_TILE_BWD = 4096;          // backward: two operands per stage -> 16 KiB tiles
constexpr int F_NT = 256;
constexpr int F_STAGES = 2;
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

template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <bool IS_LFQ, int DQ, int CPT>
__global__ void __launch_bounds__(F_NT)
fused_forward_kernel(const FusedParams p) {
  using namespace ptx;
  constexpr int LPR = F_D / CPT;                 // lanes per row
  constexpr int RPP = F_NT / LPR;                // rows per pass
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t full[F_STAGES];
  __shared__ unsigned lbm[Q_LOCAL_WORDS];
  const UniqWs w(p.ws);
  const int tid = threadIdx.x;
  const int d = p.d, T = p.T;
  const int kq = tid % LPR, rsub = tid / LPR;
  const int slab = F_D * T;

  float win[DQ][CPT], wout[CPT][DQ], bin[DQ], bout[CPT], fb[DQ];
#pragma unroll
  for (int j = 0; j < DQ; ++j) {
    bin[j] = (j < d) ? __ldg(p.b_in + j) : 0.f;
    fb[j] = (!IS_LFQ && j < d) ? (float)__ldg(p.basis + j) : 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      win[j][c] = (j < d) ? __ldg(p.W_in + j * F_D + kq + c * LPR) : 0.f;
      wout[c][j] = (j < d) ? __ldg(p.W_out + (kq + c * LPR) * d + j) : 0.f;
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) bout[c] = __ldg(p.b_out + kq + c * LPR);

  if (tid == 0) {
    for (int s = 0; s < F_STAGES; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  for (int i = tid; i < Q_LOCAL_WORDS; i += F_NT) lbm[i] = 0u;
  __syncthreads();

  const long long my_tiles = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_samples = [&](long long i) {
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    return (int)min((long long)p.samples_per_tile, p.B - b0);
  };
  auto issue_load = [&](long long i) {           // thread 0 only
    const int s = (int)(i % F_STAGES);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const uint32_t bytes = (uint32_t)tile_samples(i) * slab * 4;
    mbar_expect_tx(smem_u32(full + s), bytes);
    bulk_g2s(smem_u32(smem + (size_t)s * F_TILE_FWD), p.z + b0 * slab, bytes, smem_u32(full + s));
  };
  if (tid == 0) for (long long i = 0; i < my_tiles && i < F_STAGES - 1; ++i) issue_load(i);

  float ent = 0.f;
  for (long long i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % F_STAGES);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const int ns = tile_samples(i);
    const int rows = ns * T;
    float* X = smem + (size_t)s * F_TILE_FWD;
    if (tid == 0 && i + F_STAGES - 1 < my_tiles) {
      bulk_wait_read<0>();                       // the store that last read the other stage has drained it
      issue_load(i + F_STAGES - 1);
    }
    mbar_wait(smem_u32(full + s), (uint32_t)((i / F_STAGES) & 1), nullptr, 0);
    for (int r0 = 0; r0 < rows; r0 += RPP) {     // warp-uniform trip count: the butterfly needs every lane
      const int r = r0 + rsub;
      const bool active = r < rows;
      const int bl = active ? r / T : 0, t = active ? r - bl * T : 0;
      float* px = X + bl * slab + t + kq * T;
      float x[CPT], ze[DQ], zh[DQ];
#pragma unroll
      for (int c = 0; c < CPT; ++c) x[c] = active ? px[c * LPR * T] : 0.f;
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) a = fmaf(win[j][c], x[c], a);
        ze[j] = row_sum<LPR>(a) + bin[j];
        zh[j] = IS_LFQ ? lfq_sign_st(ze[j]) : fsq_round_st(ze[j]);
      }
      if (IS_LFQ && active && kq < d) {
        // entropy term of component kq (one lane per component); only its mean enters the loss (1e-5 tolerance)
        // -> MUFU-based fast intrinsics
        float zk = ze[0];
#pragma unroll
        for (int j = 1; j < DQ; ++j) zk = (kq == j) ? ze[j] : zk;
        const float pr = __fdividef(1.f, 1.f + __expf(-zk));
        const float q = 1.f - pr;
        ent -= fmaf(pr, __logf(pr + 1e-6f), q * __logf(q + 1e-6f));
      }
      if (active) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          float o = bout[c];
#pragma unroll
          for (int j = 0; j < DQ; ++j) o = fmaf(wout[c][j], zh[j], o);
          px[c * LPR * T] = o;
        }
        if (kq == 0) {
          const long long b = b0 + bl;
          long long code = 0;
          float sidx = 0.f;
#pragma unroll
          for (int j = 0; j < DQ; ++j) {
            if (j < d) {
              p.z_e[(b * d + j) * T + t] = ze[j];
              if (IS_LFQ) {
                if (zh[j] > 0.f) code |= (1LL << j);
              } else {
                const float pj = __fmul_rn(zh[j], fb[j]);                 // :135 float multiply-sum, then truncate
                sidx = (j == 0) ? pj : __fadd_rn(sidx, pj);
              }
            }
          }
          if (!IS_LFQ) code = trunc_to_i64(sidx);
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
// ------------------------------------------------------------------------------------------
template <bool IS_LFQ, int DQ, int CPT>
__global__ void __launch_bounds__(F_NT)
fused_backward_kernel(const FusedParams p) {
  using namespace ptx;
  constexpr int LPR = F_D / CPT;
  constexpr int RPP = F_NT / LPR;
  constexpr int STAGE_FLOATS = 2 * F_TILE_BWD;    // [g_out tile | z tile]
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t full[F_STAGES];
  __shared__ float sacc[2 * F_D * F_MAX_DQ + F_D + F_MAX_DQ];
  const int tid = threadIdx.x;
  const int d = p.d, T = p.T;
  const int kq = tid % LPR, rsub = tid / LPR;
  const int slab = F_D * T;
  const int nacc = 2 * F_D * d + F_D + d;

  float win[DQ][CPT], wout[CPT][DQ];
#pragma unroll
  for (int j = 0; j < DQ; ++j) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      win[j][c] = (j < d) ? __ldg(p.W_in + j * F_D + kq + c * LPR) : 0.f;
      wout[c][j] = (j < d) ? __ldg(p.W_out + (kq + c * LPR) * d + j) : 0.f;
    }
  }
  float a_win[DQ][CPT] = {}, a_wout[CPT][DQ] = {}, a_bout[CPT] = {}, a_bin[DQ] = {};
  const float lscale = IS_LFQ ? (p.g_loss ? __ldg(p.g_loss) : 1.f) * (-p.weight / (float)((double)p.B * T * d)) : 0.f;

  if (tid == 0) {
    for (int s = 0; s < F_STAGES; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  for (int i = tid; i < nacc; i += F_NT) sacc[i] = 0.f;
  __syncthreads();

  const long long my_tiles = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_samples = [&](long long i) {
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    return (int)min((long long)p.samples_per_tile, p.B - b0);
  };
  auto issue_load = [&](long long i) {           // thread 0 only
    const int s = (int)(i % F_STAGES);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const uint32_t bytes = (uint32_t)tile_samples(i) * slab * 4;
    float* G = smem + (size_t)s * STAGE_FLOATS;
    mbar_expect_tx(smem_u32(full + s), 2 * bytes);
    bulk_g2s(smem_u32(G), p.g_out + b0 * slab, bytes, smem_u32(full + s));
    bulk_g2s(smem_u32(G + F_TILE_BWD), p.z + b0 * slab, bytes, smem_u32(full + s));
  };
  if (tid == 0) for (long long i = 0; i < my_tiles && i < F_STAGES - 1; ++i) issue_load(i);

  for (long long i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % F_STAGES);
    const long long b0 = (blockIdx.x + i * gridDim.x) * p.samples_per_tile;
    const int ns = tile_samples(i);
    const int rows = ns * T;
    float* G = smem + (size_t)s * STAGE_FLOATS;
    const float* X = G + F_TILE_BWD;
    if (tid == 0 && i + F_STAGES - 1 < my_tiles) {
      bulk_wait_read<0>();
      issue_load(i + F_STAGES - 1);
    }
    mbar_wait(smem_u32(full + s), (uint32_t)((i / F_STAGES) & 1), nullptr, 0);
    for (int r0 = 0; r0 < rows; r0 += RPP) {
      const int r = r0 + rsub;
      const bool active = r < rows;
      const int bl = active ? r / T : 0, t = active ? r - bl * T : 0;
      const int off = bl * slab + t + kq * T;
      float g[CPT], x[CPT], gze[DQ], zq[DQ];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { g[c] = active ? G[off + c * LPR * T] : 0.f; x[c] = active ? X[off + c * LPR * T] : 0.f; }
      const float* pze = p.z_e + ((b0 + bl) * d) * T + t;
      float eterm = 0.f;                          // LFQ: lane kq computes the entropy-gradient term of component kq
      if (IS_LFQ && active && kq < d) {           // d(-w * mean H_b(sigmoid(z_e)))/dz_e, SURVEY row a14
        const float zk = __ldg(pze + kq * T);
        const float dl = 1e-6f;
        const float pr = 1.f / (1.f + expf(-zk));
        const float q = 1.f - pr;
        const float dH = -(logf(pr + dl) + pr / (pr + dl) - logf(q + dl) - q / (q + dl));
        eterm = lscale * dH * (pr * q);
      }
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        const float ze = (active && j < d) ? __ldg(pze + j * T) : 0.f;
        zq[j] = (j < d) ? (IS_LFQ ? lfq_sign_st(ze) : fsq_round_st(ze)) : 0.f;
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) a = fmaf(wout[c][j], g[c], a);
        float gj = row_sum<LPR>(a);
        if (IS_LFQ) gj += __shfl_sync(0xffffffffu, eterm, j, LPR);
        gze[j] = (active && j < d) ? gj : 0.f;
      }
      if (active) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          float o = 0.f;
#pragma unroll
          for (int j = 0; j < DQ; ++j) o = fmaf(win[j][c], gze[j], o);
          G[off + c * LPR * T] = o;                // gradient wrt z, in place over the g_out tile
          a_bout[c] += g[c];
#pragma unroll
          for (int j = 0; j < DQ; ++j) { a_wout[c][j] = fmaf(g[c], zq[j], a_wout[c][j]); a_win[j][c] = fmaf(gze[j], x[c], a_win[j][c]); }
        }
        if (kq == 0) {
#pragma unroll
          for (int j = 0; j < DQ; ++j) a_bin[j] += gze[j];
        }
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

  // ---- parameter gradients: registers -> shared (CTA) -> global (one atomic per element per CTA) ----
  float* s_win = sacc;                       // [d][64]
  float* s_bin = s_win + d * F_D;            // [d]
  float* s_wout = s_bin + d;                 // [64][d]
  float* s_bout = s_wout + F_D * d;          // [64]
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int ch = kq + c * LPR;
    atomicAdd(s_bout + ch, a_bout[c]);
#pragma unroll
    for (int j = 0; j < DQ; ++j) {
      if (j < d) { atomicAdd(s_win + j * F_D + ch, a_win[j][c]); atomicAdd(s_wout + ch * d + j, a_wout[c][j]); }
    }
  }
  if (kq == 0) {
#pragma unroll
    for (int j = 0; j < DQ; ++j) if (j < d) atomicAdd(s_bin + j, a_bin[j]);
  }
  __syncthreads();
  for (int i = tid; i < nacc; i += F_NT) { const float v = sacc[i]; if (v != 0.f) atomicAdd(p.grads + i, v); }
}

static bool fused_geom(const void* a, const void* b, int64_t B, int64_t D, int64_t d, int64_t T, FusedParams& p,
                       int tile_elems = F_TILE_BWD) {
  if (D != F_D || d < 1 || d > F_MAX_DQ || T < 1 || B < 1) return false;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return false;
  const long long slab = D * T;
  if (slab > F_TILE_BWD) return false;
  p.samples_per_tile = (int)(tile_elems / slab);
  p.ntiles = (B + p.samples_per_tile - 1) / p.samples_per_tile;
  p.B = B; p.d = (int)d; p.T = (int)T;
  return true;
}

template <bool IS_LFQ, bool BWD>
static int launch_fused(const FusedParams& p, cudaStream_t stream) {
  const int per_sm = BWD ? 3 : 3;
  const int grid = (int)max(1LL, min(p.ntiles, (long long)sm_count() * per_sm));
  const size_t smem = (size_t)F_STAGES * (BWD ? 2 * F_TILE_BWD : F_TILE_FWD) * sizeof(float);
#define VQ_FUSED_CASE(DQ, CPT)                                                                                        \
  do {                                                                                                                \
    auto kern = BWD ? fused_backward_kernel<IS_LFQ, DQ, CPT> : fused_forward_kernel<IS_LFQ, DQ, CPT>;                 \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fsq_lfq_fused)");                                 \
    kern<<<grid, F_NT, smem, stream>>>(p);                                                                            \
  } while (0)
  if (p.d <= 4) VQ_FUSED_CASE(4, 4);
  else if (p.d <= 8) VQ_FUSED_CASE(8, 2);
  else if (p.d <= 12) VQ_FUSED_CASE(12, 2);
  else VQ_FUSED_CASE(16, 2);
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
  VQ_CHECK_ARG(fused_geom(z, out, B, D, d, T, p, F_TILE_FWD), VQB200_EUNSUPPORTED,
               "fsq_fused_forward: needs contiguous 16-byte aligned [B,64,T] with T <= 64 and d <= 16");
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
  VQ_CHECK_ARG(fused_geom(z, out, B, D, d, T, p, F_TILE_FWD), VQB200_EUNSUPPORTED,
               "lfq_fused_forward: needs contiguous 16-byte aligned [B,64,T] with T <= 64 and d <= 16");
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
               "proj_fused_backward: needs contiguous 16-byte aligned [B,64,T] with T <= 64 and d <= 16");
  p.z = z; p.g_out = g_out; p.out = g_z; p.z_e = const_cast<float*>(z_e); p.W_in = W_in; p.W_out = W_out;
  p.g_loss = g_loss; p.weight = entropy_loss_weight; p.grads = grads;
  VQ_CUDA(cudaMemsetAsync(grads, 0, vqb200_proj_fused_grad_floats(D, d) * sizeof(float), stream));
  return is_lfq ? launch_fused<true, true>(p, stream) : launch_fused<false, true>(p, stream);
}

}  // extern "C"
