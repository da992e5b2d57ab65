// vqb200 K4: single-launch ResidualVQ for the launch-bound shapes (BASELINE cfg2: 512 vectors, 4 x K=512).
//
// Replaces the whole Python loop of models/vqvae.py:94-100 (S x {distances, argmin, one-hot sums, EMA update,
// gather, loss, metrics}, ~25 ATen launches per stage in the reference, 7 vqb200 launches per stage in the
// multi-kernel path) by ONE kernel: a thread-block cluster of 16 CTAs (one GPC) owns the batch, keeps the running residual in
// shared memory across all stages, and orders the stage phases
//     assign + statistics  ->  EMA cluster sizes  ->  codebook update  ->  gather / residual / running sum
// with cluster barriers (release/acquire at cluster scope), so the update-then-gather ordering of the reference
// (:43-52) is kept without leaving the kernel.  Arithmetic is the exact fp32 formulation of assign_simt.cu
// (d = fl(fl(|x|^2 + |E|^2) - 2 x.E), first minimum, NaN wins) and of ema.cu / gather.cu.
// Eligible: D == 64, N <= 4096, K <= 4096, S <= 8, single process (no inter-GPU all-reduce inside the launch).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace vqb200 {
namespace small {

constexpr int D = 64;
constexpr int CLUSTER = 16;                     // non-portable cluster size (one GPC), opt-in below
constexpr int NT = 256;
constexpr int MAX_S = 8;
constexpr int MAX_K = 4096;
constexpr int MAX_ROWS_PER_CTA = 256;           // N <= 4096
constexpr int LDR = D + 1;                      // padded residual rows: conflict-free column walks
constexpr int CHUNK = 128;                      // codes staged in shared memory per step (32 KiB)

struct Args {
  ZView z;
  int S;
  int training_ema;                             // 1: EMA statistics + codebook update between assign and gather
  int use_ema;                                  // loss formula: c*mse (EMA) vs mse + c*mse
  float commitment;
  float decay, one_minus_decay, eps;
  float* E[MAX_S];
  float* cs[MAX_S];
  float* w[MAX_S];
  float k_eps[MAX_S];
  int K[MAX_S];
  float* stats;                                 // [S][K_s*(D+1)] packed back to back (dw | cnt), zeroed in-kernel
  float* scratch;                               // [S][K_s + 8]: normalised cluster sizes, n
  double* sse;                                  // [S]
  int32_t* idx;                                 // [S][N]
  float* out;                                   // [B,C,T] contiguous
  float* m3;                                    // [S][3] loss, perplexity, dcr
  int stamps;                                   // development (VQB200_RVQ_STAMPS): CTA 0 prints per-phase globaltimer stamps
  // data-parallel ranks of one node (wide kernel only; world == 1: stats_of[0] == stats): every rank's statistics slot
  // and flag words as mapped into this process, the first barrier epoch of this call, vectors of ALL ranks
  int world, rank, peer_timeout_s;
  unsigned epoch0;
  long long n_total;
  const float* stats_of[VQB200_MAX_PEERS];
  unsigned* flags_of[VQB200_MAX_PEERS];
};

// [dw (K*D) | cnt (K)] of one stage, padded so that every stage's dw stays 16-byte aligned (vector reductions)
__host__ __device__ __forceinline__ long long stage_stats_floats(long long K) { return (K * (D + 1) + 3) & ~3LL; }
__device__ __forceinline__ long long stats_offset(const Args& a, int s) {
  long long o = 0;
  for (int i = 0; i < s; ++i) o += stage_stats_floats(a.K[i]);
  return o;
}
__device__ __forceinline__ long long scratch_offset(const Args& a, int s) {
  long long o = 0;
  for (int i = 0; i < s; ++i) o += a.K[i] + 8;
  return o;
}

__global__ void __launch_bounds__(NT, 1)
rvq_small_kernel(const Args a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float smem[];
  float* R = smem;                                        // [rows_c][LDR] running residual
  float* ee = R + MAX_ROWS_PER_CTA * LDR;                 // [MAX_K] |E_k|^2 of the current stage
  float* bestd = ee + MAX_K;                              // [8 warps][32]
  int* bestk = reinterpret_cast<int*>(bestd + 8 * 32);    // [8 warps][32]
  int* rowk = bestk + 8 * 32;                             // [MAX_ROWS_PER_CTA] code of each row (current stage)
  float* Ech = reinterpret_cast<float*>(rowk + MAX_ROWS_PER_CTA);   // [CHUNK][D] staged codebook chunk (16-byte aligned)
  __shared__ double red[8];
  __shared__ float s_n;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster.block_rank();
  const long long N = a.z.N;
  const int rows_per = (int)((N + CLUSTER - 1) / CLUSTER);
  const long long row0 = (long long)rank * rows_per;
  const int rows = (int)max(0LL, min((long long)rows_per, N - row0));
  const int C = (int)a.z.C, T = (int)a.z.T;

  // ---- load this CTA's rows, zero the statistics of all stages ----
  for (int i = tid; i < rows * D; i += NT) {
    const int r = i / D, k = i - r * D;
    R[r * LDR + k] = __ldg(a.z.p + a.z.row_base(row0 + r) + (long long)k * a.z.sC);
  }
  {
    const long long total = stats_offset(a, a.S);
    for (long long i = (long long)rank * NT + tid; i < total; i += (long long)CLUSTER * NT) a.stats[i] = 0.f;
    if (rank == 0 && tid < a.S) a.sse[tid] = 0.0;
  }
  cluster.sync();

  for (int s = 0; s < a.S; ++s) {
    const int K = a.K[s];
    float* __restrict__ E = a.E[s];
    float* dw = a.stats + stats_offset(a, s);
    float* cnt = dw + (long long)K * D;
    float* cl = a.scratch + scratch_offset(a, s);

    // ---- K1: exact fp32 distances + argmin.  lane = row (held in 64 registers); the codebook is staged through
    //      shared memory in 128-code chunks (coalesced loads, then warp-uniform broadcast reads); G warps share a
    //      32-row block and split the codes of every chunk between them ----
    {
      const int nb = (rows + 31) >> 5;                                  // 32-row blocks of this CTA
      const int G = nb <= 1 ? 8 : (nb == 2 ? 4 : (nb <= 4 ? 2 : 1));    // warps per row block
      const int P = 8 / G;                                              // row blocks per pass
      const int sub = warp % G;
      for (int pass = 0; pass * P < max(nb, 1); ++pass) {
        const int blk = pass * P + warp / G;
        const int r = blk * 32 + lane;
        const bool valid = blk < nb && r < rows;
        float x[D];
        float xx = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) { x[c] = valid ? R[r * LDR + c] : 0.f; xx = fmaf(x[c], x[c], xx); }
        float bd = INFINITY; int bk = INT_MAX;
        for (int k0 = 0; k0 < K; k0 += CHUNK) {
          const int kc = min(CHUNK, K - k0);
          __syncthreads();                                              // previous chunk fully consumed
          {
            // CHUNK*D/4 = 2048 float4 = 8 per thread: issue all loads before the first store (one L2 round trip)
            const float4* src = reinterpret_cast<const float4*>(E + (size_t)k0 * D);
            float4* dst = reinterpret_cast<float4*>(Ech);
            const int n4 = kc * (D / 4);
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = tid + u * NT; if (i < n4) t[u] = src[i]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = tid + u * NT; if (i < n4) dst[i] = t[u]; }
          }
          __syncthreads();
          if (pass == 0) {
            // |E_k|^2 of the chunk (needed once per stage): one code per thread, 16-byte reads rotated by the
            // lane id so that the 32 rows a warp touches hit different banks
            if (tid < kc) {
              const float4* e4 = reinterpret_cast<const float4*>(Ech + tid * D);
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < D / 4; ++c) {
                const float4 v = e4[(c + lane) & (D / 4 - 1)];
                acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
              }
              ee[k0 + tid] = acc;
            }
            __syncthreads();
          }
          if (blk < nb) {
            // four codes per iteration: four independent FMA chains (each still sums dims 0..63 in order)
            for (int kk = sub; kk < kc; kk += 4 * G) {
              const float4* e0 = reinterpret_cast<const float4*>(Ech + kk * D);     // warp-uniform: broadcast
              const float4* e1 = reinterpret_cast<const float4*>(Ech + min(kk + G, kc - 1) * D);
              const float4* e2 = reinterpret_cast<const float4*>(Ech + min(kk + 2 * G, kc - 1) * D);
              const float4* e3 = reinterpret_cast<const float4*>(Ech + min(kk + 3 * G, kc - 1) * D);
              float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
              for (int c = 0; c < D / 4; ++c) {
                const float4 v0 = e0[c], v1 = e1[c], v2 = e2[c], v3 = e3[c];
                d0 = fmaf(x[4 * c], v0.x, d0); d1 = fmaf(x[4 * c], v1.x, d1); d2 = fmaf(x[4 * c], v2.x, d2); d3 = fmaf(x[4 * c], v3.x, d3);
                d0 = fmaf(x[4 * c + 1], v0.y, d0); d1 = fmaf(x[4 * c + 1], v1.y, d1); d2 = fmaf(x[4 * c + 1], v2.y, d2); d3 = fmaf(x[4 * c + 1], v3.y, d3);
                d0 = fmaf(x[4 * c + 2], v0.z, d0); d1 = fmaf(x[4 * c + 2], v1.z, d1); d2 = fmaf(x[4 * c + 2], v2.z, d2); d3 = fmaf(x[4 * c + 2], v3.z, d3);
                d0 = fmaf(x[4 * c + 3], v0.w, d0); d1 = fmaf(x[4 * c + 3], v1.w, d1); d2 = fmaf(x[4 * c + 3], v2.w, d2); d3 = fmaf(x[4 * c + 3], v3.w, d3);
              }
              const float dots[4] = {d0, d1, d2, d3};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int kl = kk + u * G;
                if (kl < kc) {
                  const int k = k0 + kl;
                  const float d = __fsub_rn(__fadd_rn(xx, ee[k]), __fmul_rn(2.0f, dots[u]));
                  if (cand_better(d, k, bd, bk)) { bd = d; bk = k; }
                }
              }
            }
          }
        }
        bestd[warp * 32 + lane] = bd; bestk[warp * 32 + lane] = bk;
        __syncthreads();
        if (sub == 0 && valid) {
          for (int g2 = 1; g2 < G; ++g2) {
            const float od = bestd[(warp + g2) * 32 + lane]; const int ok = bestk[(warp + g2) * 32 + lane];
            if (cand_better(od, ok, bd, bk)) { bd = od; bk = ok; }
          }
          rowk[r] = bk;
          a.idx[(long long)s * N + row0 + r] = bk;
        }
        __syncthreads();
      }
    }

    // ---- K3a: statistics (counts always: they feed perplexity / dcr) ----
    for (int r = tid; r < rows; r += NT) atomicAdd(cnt + rowk[r], 1.0f);
    if (a.training_ema) {
      for (int i = tid; i < rows * (D / 4); i += NT) {
        const int r = i / (D / 4), q = i - r * (D / 4);
        const float* p = R + r * LDR + 4 * q;
        red_add_v4(dw + (size_t)rowk[r] * D + 4 * q, p[0], p[1], p[2], p[3]);
      }
    }
    if (a.training_ema) {
      __threadfence();
      cluster.sync();                                      // every CTA's statistics are in
      // ---- K3b step 1 (cluster rank 0): cs <- decay*cs + (1-decay)*cnt ; n ; normalised cluster sizes ----
      if (rank == 0) {
        float* cs = a.cs[s];
        double part = 0.0;
        for (int k = tid; k < K; k += NT) {
          const float v = fmaf(__ldcg(cnt + k), a.one_minus_decay, __fmul_rn(cs[k], a.decay));
          cs[k] = v;
          part += (double)v;
        }
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid < 32) {
          double v = tid < 8 ? red[tid] : 0.0;
          v = warp_sum(v);
          if (tid == 0) s_n = (float)v;
        }
        __syncthreads();
        const float n = s_n;
        for (int k = tid; k < K; k += NT)
          cl[k] = __fmul_rn(__fdiv_rn(__fadd_rn(cs[k], a.eps), __fadd_rn(n, a.k_eps[s])), n);
        __threadfence();
      }
      cluster.sync();
      // ---- K3b step 2 (all CTAs, code slices): w <- decay*w + (1-decay)*dw ; E <- w / cluster ----
      {
        float* wv = a.w[s];
        const int per = (K + CLUSTER - 1) / CLUSTER;
        const int k0 = rank * per, k1 = min(K, k0 + per);
        // 16-byte vectors, two in flight per thread (the loads are independent: batch them ahead of the math)
        const int q0 = k0 * (D / 4), q1 = k1 * (D / 4);
        for (int i = q0 + tid; i < q1; i += 2 * NT) {
          const int j = i + NT;
          const bool two = j < q1;
          const float4 d0 = __ldcg(reinterpret_cast<const float4*>(dw) + i);
          const float4 w0 = reinterpret_cast<const float4*>(wv)[i];
          const float c0 = __ldcg(cl + i / (D / 4));
          const float4 d1 = two ? __ldcg(reinterpret_cast<const float4*>(dw) + j) : d0;
          const float4 w1 = two ? reinterpret_cast<const float4*>(wv)[j] : w0;
          const float c1 = two ? __ldcg(cl + j / (D / 4)) : c0;
          float4 n0, n1, e0, e1;
          n0.x = fmaf(d0.x, a.one_minus_decay, __fmul_rn(w0.x, a.decay)); n0.y = fmaf(d0.y, a.one_minus_decay, __fmul_rn(w0.y, a.decay));
          n0.z = fmaf(d0.z, a.one_minus_decay, __fmul_rn(w0.z, a.decay)); n0.w = fmaf(d0.w, a.one_minus_decay, __fmul_rn(w0.w, a.decay));
          n1.x = fmaf(d1.x, a.one_minus_decay, __fmul_rn(w1.x, a.decay)); n1.y = fmaf(d1.y, a.one_minus_decay, __fmul_rn(w1.y, a.decay));
          n1.z = fmaf(d1.z, a.one_minus_decay, __fmul_rn(w1.z, a.decay)); n1.w = fmaf(d1.w, a.one_minus_decay, __fmul_rn(w1.w, a.decay));
          e0 = make_float4(__fdiv_rn(n0.x, c0), __fdiv_rn(n0.y, c0), __fdiv_rn(n0.z, c0), __fdiv_rn(n0.w, c0));
          e1 = make_float4(__fdiv_rn(n1.x, c1), __fdiv_rn(n1.y, c1), __fdiv_rn(n1.z, c1), __fdiv_rn(n1.w, c1));
          reinterpret_cast<float4*>(wv)[i] = n0; reinterpret_cast<float4*>(E)[i] = e0;
          if (two) { reinterpret_cast<float4*>(wv)[j] = n1; reinterpret_cast<float4*>(E)[j] = e1; }
        }
        __threadfence();
      }
      cluster.sync();                                      // the updated codebook is complete
    }

    // ---- K2: gather (post-update codebook), straight-through value, loss sum, running sum, next residual ----
    float part = 0.f;
    {
      constexpr int U = 4;                                 // independent codeword / running-sum loads in flight
      for (int i0 = tid; i0 < rows * D; i0 += NT * U) {
        float q[U], prev[U]; float* op[U]; int ri[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * NT;
          if (i < rows * D) {
            const int r = i / D, c = i - r * D;
            ri[u] = r * LDR + c;
            q[u] = __ldcg(E + (size_t)rowk[r] * D + c);
            const long long n = row0 + r;
            const long long b = n / T; const int t = (int)(n - b * T);
            op[u] = a.out + (b * C + c) * T + t;
            prev[u] = s > 0 ? *op[u] : 0.f;
          } else { ri[u] = -1; q[u] = 0.f; op[u] = nullptr; prev[u] = 0.f; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (ri[u] < 0) continue;
          const float x = R[ri[u]];
          const float diff = __fsub_rn(q[u], x);
          const float st = __fadd_rn(x, diff);
          part = fmaf(diff, diff, part);
          *op[u] = __fadd_rn(prev[u], st);
          R[ri[u]] = __fsub_rn(x, st);
        }
      }
    }
    {
      double p = warp_sum((double)part);
      __syncthreads();
      if (lane == 0) red[warp] = p;
      __syncthreads();
      if (tid < 32) {
        double v = tid < 8 ? red[tid] : 0.0;
        v = warp_sum(v);
        if (tid == 0 && v != 0.0) atomicAdd(a.sse + s, v);
      }
    }
    __syncthreads();
  }

  // ---- loss / perplexity / dcr of every stage (cluster rank 0, one warp per stage) ----
  __threadfence();
  cluster.sync();
  if (rank == 0 && warp < a.S) {
    const int s = warp;
    const int K = a.K[s];
    const float* cnt = a.stats + stats_offset(a, s) + (long long)K * D;
    const float Nf = (float)N;
    double ent = 0.0; int active = 0;
    for (int k0 = 0; k0 < K; k0 += 32 * 8) {
      float c8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int k = k0 + u * 32 + lane; c8[u] = (k < K) ? __ldcg(cnt + k) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = __fdiv_rn(c8[u], Nf);
        ent += (double)__fmul_rn(p, logf(__fadd_rn(p, 1e-10f)));
        active += (c8[u] > 0.f);
      }
    }
    ent = warp_sum(ent);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) active += __shfl_xor_sync(0xffffffffu, active, o);
    if (lane == 0) {
      const float mse = (float)(__ldcg(a.sse + s) / ((double)N * D));
      a.m3[s * 3 + 0] = a.use_ema ? __fmul_rn(a.commitment, mse) : __fadd_rn(mse, __fmul_rn(a.commitment, mse));
      a.m3[s * 3 + 1] = expf(-(float)ent);
      a.m3[s * 3 + 2] = __fsub_rn(1.0f, __fdiv_rn((float)active, (float)K));
    }
  }
  __syncthreads();
  if (rank == 0 && tid == 0) {                        // row S: sum of the stage losses, mean perplexity, mean dcr
    float l = 0.f, p = 0.f, d = 0.f;
    for (int s = 0; s < a.S; ++s) { l = __fadd_rn(l, a.m3[s * 3]); p = __fadd_rn(p, a.m3[s * 3 + 1]); d = __fadd_rn(d, a.m3[s * 3 + 2]); }
    a.m3[a.S * 3 + 0] = l;
    a.m3[a.S * 3 + 1] = __fdiv_rn(p, (float)a.S);
    a.m3[a.S * 3 + 2] = __fdiv_rn(d, (float)a.S);
  }
}

}  // namespace small

// ==========================================================================================
// Wide variant (the BASELINE cfg2 shape and its neighbours: N <= 64 x row blocks, sum_s ceil(K_s / 8) <= 384): the same
// algorithm spread over the whole GPU instead of one GPC.  Grid = row blocks x 8 code slices, 512 threads; the 8 CTAs
// that share a row block form a cluster; the number of row blocks is what cudaOccupancyMaxActiveClusters reports as
// co-resident (15 on a B200).  What makes it fast is the dependency chain, not the FLOPs:
//   * every CTA loads its code slice of EVERY stage (pre-update codebooks, known at launch) once, up front;
//   * per stage: register-tiled exact distances over the slice (4 rows x 4 codes per lane) -> ONE cluster barrier,
//     candidates merged through distributed shared memory (64-bit keys, cand_better order) -> EMA statistics as L2
//     reductions -> ONE grid barrier -> every CTA derives cs', n and the updated codewords of its rows' codes locally
//     (same formulas, same bits as the in-place update) and applies gather / residual / running sum to its private copy
//     of the row block;
//   * the in-place updates of (ema_cluster_size, ema_w, embedding) of ALL stages are deferred past the last grid
//     barrier (nobody reads the old values any more) and done as one flattened pass; CTA s reports stage s's metrics and
//     the last reporter adds the aggregate row m3[S];
//   * data-parallel ranks (world > 1): statistics live in NVLink peer slots, the per-stage grid barrier also spans the
//     ranks (grid_barrier_world) and every read of [dw | cnt] is a rank-ordered sum over all slots.
// The grid barrier is a monotonically increasing counter in the workspace with a bounded spin (a protocol bug prints
// and traps instead of hanging the GPU); the launch is cooperative, so all CTAs are co-resident.  Under Nsight Compute
// (which cannot launch cooperative + clustered kernels) the attribute is dropped; VQB200_RVQ_STAMPS=1 prints the phase
// timeline of CTA 0.
// ==========================================================================================
namespace wide {
using small::Args;
using small::stats_offset;
constexpr int D = 64, NT = 512, NW = NT / 32, CS = 8, MAX_RB = 16;    // code slices (cluster size), row blocks (runtime: as many 8-CTA
                                                        // clusters as the device can keep co-resident, at most 16)
constexpr int MAX_RPB = 64;                             // rows per block  (N <= 1024)
constexpr int MAX_SLICE_CODES = 384;                    // sum over stages of ceil(K_s / 8)  (S*K <= 3072)
constexpr int LDR = D + 4, LDE = D + 4;                 // row pitches: 16-byte aligned, bank-conflict-free tiles
constexpr unsigned SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& generation, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = (generation + 1u) * nblocks;
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++spins > SPIN_LIMIT) {
        printf("vqb200 rvq_wide_kernel: grid barrier timed out in CTA %d (generation %u: %u of %u arrivals)\n",
               (int)blockIdx.x, generation, v, target);
        __trap();
      }
    }
    __threadfence();
  }
  ++generation;
  __syncthreads();
}

// statistics summed over the ranks in RANK ORDER (identical bits on every rank); off = float offset inside a slot
__device__ __forceinline__ float sum_ranks(const Args& a, long long off) {
  float v = __ldcg(a.stats_of[0] + off);
  for (int p = 1; p < a.world; ++p) v = __fadd_rn(v, __ldcg(a.stats_of[p] + off));
  return v;
}
__device__ __forceinline__ float4 sum_ranks4(const Args& a, long long off4) {      // off4 in float4 units
  float4 v = __ldcg(reinterpret_cast<const float4*>(a.stats_of[0]) + off4);
  for (int p = 1; p < a.world; ++p) {
    const float4 o = __ldcg(reinterpret_cast<const float4*>(a.stats_of[p]) + off4);
    v.x = __fadd_rn(v.x, o.x); v.y = __fadd_rn(v.y, o.y); v.z = __fadd_rn(v.z, o.z); v.w = __fadd_rn(v.w, o.w);
  }
  return v;
}

// Grid barrier that also spans the data-parallel ranks: every CTA arrives on the local counter; CTA 0 waits for all
// local arrivals, publishes `epoch` into every peer's flag word over NVLink and waits for theirs (csrc/peer.cu
// protocol), then releases the local CTAs through a second word.  world == 1 callers use grid_barrier().
__device__ __forceinline__ void grid_barrier_world(const Args& a, unsigned* bar, unsigned& generation, unsigned nblocks,
                                                   unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = (generation + 1u) * nblocks;
    unsigned spins = 0;
    if (blockIdx.x == 0) {
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v >= target) break;
        if (++spins > SPIN_LIMIT) { printf("vqb200 rvq_wide_kernel: local arrivals timed out (%u of %u)\n", v, target); __trap(); }
      }
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(a.flags_of[p] + a.rank), "r"(epoch) : "memory");
      unsigned long long t0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (int p = 0; p < a.world; ++p) {
        unsigned polls = 0;
        while (true) {
          unsigned v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.flags_of[a.rank] + p) : "memory");
          if ((int)(v - epoch) >= 0) break;
          if ((++polls & 1023u) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > (unsigned long long)a.peer_timeout_s * 1000000000ull) {
              printf("vqb200 rvq_wide_kernel: rank %d waited %d s for rank %d (epoch %u, saw %u)\n", a.rank, a.peer_timeout_s, p, epoch, v);
              __trap();
            }
          }
        }
      }
      __threadfence_system();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 2), "r"(generation + 1u) : "memory");
    }
    spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 2) : "memory");
      if (v >= generation + 1u) break;
      if (++spins > (SPIN_LIMIT << 6)) { printf("vqb200 rvq_wide_kernel: release word timed out in CTA %d\n", (int)blockIdx.x); __trap(); }
    }
    __threadfence();
  }
  ++generation;
  __syncthreads();
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VQ_STAMP(i) do { if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) { stamp[(i)] = gtime(); if ((i) == 0) cyc0 = clock64(); } } while (0)

// One pass of exact distances for this CTA: warp tile = (8 * RI) rows x 16 codes, lane = RI rows x 4 codes (rows
// i*8 + rg, codes j*4 + cg).  Candidates go to ckey[] with 64-bit atomicMin (cand_better order).
template <int RI>
__device__ __forceinline__ void k1_tiles(const float* __restrict__ R, const float* __restrict__ Esl,
                                         const float* __restrict__ eesl, const float* __restrict__ xxs,
                                         unsigned long long* __restrict__ ckey, int rows, int ks, int k0s, int warp, int lane) {
  const int rg = lane >> 2, cg = lane & 3;
  const int code_blocks = (ks + 15) >> 4;
  const int tiles = ((rows + 8 * RI - 1) / (8 * RI)) * code_blocks;
  for (int t = warp; t < tiles; t += NW) {
    const int rblk = t / code_blocks, cblk = t - rblk * code_blocks;
    int rI[RI], kI[4];
#pragma unroll
    for (int i = 0; i < RI; ++i) rI[i] = rblk * (8 * RI) + i * 8 + rg;
#pragma unroll
    for (int j = 0; j < 4; ++j) kI[j] = cblk * 16 + j * 4 + cg;
    const float4* xp[RI]; const float4* ep[4];
#pragma unroll
    for (int i = 0; i < RI; ++i) xp[i] = reinterpret_cast<const float4*>(R + min(rI[i], rows - 1) * LDR);
#pragma unroll
    for (int j = 0; j < 4; ++j) ep[j] = reinterpret_cast<const float4*>(Esl + min(kI[j], ks - 1) * LDE);
    float acc[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D / 4; ++c) {
      float4 xv[RI], ev[4];
#pragma unroll
      for (int i = 0; i < RI; ++i) xv[i] = xp[i][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) ev[j] = ep[j][c];
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = acc[i][j];
          d = fmaf(xv[i].x, ev[j].x, d); d = fmaf(xv[i].y, ev[j].y, d);
          d = fmaf(xv[i].z, ev[j].z, d); d = fmaf(xv[i].w, ev[j].w, d);
          acc[i][j] = d;
        }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      unsigned long long best = ~0ull;
      const float xx = xxs[min(rI[i], rows - 1)];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kI[j] < ks) {
          const float dd = __fsub_rn(__fadd_rn(xx, eesl[kI[j]]), __fmul_rn(2.0f, acc[i][j]));
          const unsigned long long key = cand_key(dd, k0s + kI[j]);
          best = key < best ? key : best;
        }
      }
      unsigned long long o = __shfl_xor_sync(0xffffffffu, best, 1);
      best = o < best ? o : best;
      o = __shfl_xor_sync(0xffffffffu, best, 2);
      best = o < best ? o : best;
      if (cg == 0 && rI[i] < rows) atomicMin(ckey + rI[i], best);
    }
  }
}

__global__ void __launch_bounds__(NT, 1)
rvq_wide_kernel(const Args a, unsigned* __restrict__ barrier) {
  __shared__ unsigned long long stamp[64];
  long long cyc0 = 0;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float smem[];
  float* R = smem;                                  // [MAX_RPB][LDR] running residual (private copy of the row block)
  float* O = R + MAX_RPB * LDR;                     // [MAX_RPB][LDR] running sum of the straight-through values
  float* Es = O + MAX_RPB * LDR;                    // [MAX_SLICE_CODES][LDE] this CTA's code slice of every stage
  float* ees = Es + MAX_SLICE_CODES * LDE;            // [MAX_SLICE_CODES] |E_k|^2
  float* csn = ees + MAX_SLICE_CODES;               // [MAX_K] cs' of the current stage
  float* xxs = csn + small::MAX_K;                  // [MAX_RPB] |x|^2 of the current residual rows
  int* rowk = reinterpret_cast<int*>(xxs + MAX_RPB); // [MAX_RPB]
  unsigned long long* ckey = reinterpret_cast<unsigned long long*>(rowk + MAX_RPB);   // [MAX_RPB] this slice's candidates
  __shared__ double red[NW];
  __shared__ float s_n;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (a.stamps && tid < 64) stamp[tid] = 0;
  __syncthreads();
  VQ_STAMP(0);
  const int slice = (int)cluster.block_rank();      // code slice 0..7
  const int rb = blockIdx.x / CS;                   // row block 0..15
  const unsigned nblocks = gridDim.x;
  const int cta = blockIdx.x;
  unsigned generation = 0;
  const long long N = a.z.N;
  const int RB = (int)(gridDim.x / CS);
  const int rpb = (int)((N + RB - 1) / RB);
  const long long row0 = (long long)rb * rpb;
  const int rows = (int)max(0LL, min((long long)rpb, N - row0));
  const int C = (int)a.z.C, T = (int)a.z.T;

  // ---- prologue: rows, and this CTA's slice of every stage's (pre-update) codebook with |E|^2.  All global loads of
  //      a batch are issued before the first store: one L2 round trip per batch instead of one per element ----
  for (int i0 = tid; i0 < rows * D; i0 += NT * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * NT;
      const int r = i / D, k = i - r * D;
      v[u] = (i < rows * D) ? __ldg(a.z.p + a.z.row_base(row0 + r) + (long long)k * a.z.sC) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * NT;
      if (i < rows * D) { const int r = i / D, k = i - r * D; R[r * LDR + k] = v[u]; O[r * LDR + k] = 0.f; }
    }
  }
  VQ_STAMP(50);
  int sl_off[small::MAX_S + 1];                     // slice offsets (codes) inside Es per stage
  sl_off[0] = 0;
#pragma unroll
  for (int s = 0; s < small::MAX_S; ++s) sl_off[s + 1] = sl_off[s] + (s < a.S ? (a.K[s] + CS - 1) / CS : 0);
  {
    // the slices of all stages as ONE flattened list of float4 (the destination is contiguous in Es by construction;
    // a short slice -- K not a multiple of 8 -- leaves its tail of the reserved range untouched and unread)
    const int total4 = sl_off[a.S] * (D / 4);
    float4* dst = reinterpret_cast<float4*>(Es);
    for (int i0 = tid; i0 < total4; i0 += NT * 8) {
      float4 v[8]; bool ok[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NT;
        ok[u] = false;
        if (i < total4) {
          const int code = i / (D / 4);
          int st = 0;
#pragma unroll
          for (int q = 1; q < small::MAX_S; ++q) st += (q < a.S && code >= sl_off[q]) ? 1 : 0;
          const int per = (a.K[st] + CS - 1) / CS;
          const int k0 = min(a.K[st], slice * per), k1 = min(a.K[st], k0 + per);
          const int kl = code - sl_off[st];
          if (kl < k1 - k0) {
            ok[u] = true;
            v[u] = __ldg(reinterpret_cast<const float4*>(a.E[st] + (size_t)(k0 + kl) * D) + (i - code * (D / 4)));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NT;
        if (ok[u]) dst[(i / (D / 4)) * (LDE / 4) + (i & (D / 4 - 1))] = v[u];
      }
    }
  }
  __syncthreads();
  VQ_STAMP(51);
  for (int s = 0; s < a.S; ++s) {
    const int per = (a.K[s] + CS - 1) / CS;
    const int k0 = min(a.K[s], slice * per), k1 = min(a.K[s], k0 + per);
    for (int k = tid; k < k1 - k0; k += NT) {
      const float4* e4 = reinterpret_cast<const float4*>(Es + (size_t)(sl_off[s] + k) * LDE);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {             // same summation order as the narrow kernel / codebook_prepare
        const float4 v = e4[(c + lane) & (D / 4 - 1)];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
      }
      ees[sl_off[s] + k] = acc;
    }
  }
  __syncthreads();

  float n_of[small::MAX_S];                         // n = sum(cs') of every stage (for the deferred in-place updates)
#pragma unroll
  for (int i = 0; i < small::MAX_S; ++i) n_of[i] = 0.f;
  VQ_STAMP(53);

  for (int s = 0; s < a.S; ++s) {
    VQ_STAMP(1 + s * 8 + 0);
    const int K = a.K[s];
    float* dw = a.stats + stats_offset(a, s);
    float* cnt = dw + (long long)K * D;
    const int per = (K + CS - 1) / CS;
    const int k0s = min(K, slice * per), ks = min(K, k0s + per) - k0s;      // this CTA's codes [k0s, k0s + ks)
    const float* Esl = Es + (size_t)sl_off[s] * LDE;
    const float* eesl = ees + sl_off[s];

    // ---- K1: exact fp32 distances + argmin over this CTA's code slice.  Register-tiled: a warp owns a tile of 32 rows
    //      x 16 codes, a lane 4 rows x 4 codes (rows i*8 + rg, codes j*4 + cg: with the padded row pitches every 16-byte
    //      shared load of a warp touches 32 distinct banks), so one step of 4 dims costs 8 one-wavefront loads for 64
    //      FMAs per lane.  (The earlier lane = row scheme broadcast one code vector per 4 FMAs and was bound by shared
    //      memory bandwidth: 4.7 us per stage.)  Every (row, code) dot product is still ONE sequential fmaf chain over
    //      dims 0..63, so distances -- and ties -- are bit-identical to the other exact paths ----
    {
      for (int r = tid; r < rows; r += NT) {          // |x|^2 per row, the same fmaf chain as everywhere else
        const float* xr = R + r * LDR;
        float xx = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) xx = fmaf(xr[c], xr[c], xx);
        xxs[r] = xx;
        ckey[r] = cand_key(INFINITY, INT_MAX);
      }
      __syncthreads();
      if (s == 0) VQ_STAMP(40);
      // rows per lane: the smallest register tile that still gives every tile to a warp in one pass (more warps busy,
      // fewer padded rows); larger slices fall back to the 4 x 4 tile with the best FMA : shared-load ratio
      {
        const int code_blocks = (ks + 15) >> 4;
        if (((rows + 15) >> 4) * code_blocks <= NW) k1_tiles<2>(R, Esl, eesl, xxs, ckey, rows, ks, k0s, warp, lane);
        else k1_tiles<4>(R, Esl, eesl, xxs, ckey, rows, ks, k0s, warp, lane);
      }
      if (s == 0) VQ_STAMP(41);
    }
    VQ_STAMP(1 + s * 8 + 1);
    cluster.sync();                                   // all 8 slices of this row block have their candidates
    VQ_STAMP(1 + s * 8 + 2);
    for (int r = tid; r < rows; r += NT) {
      unsigned long long best = ~0ull;
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        const unsigned long long v = cluster.map_shared_rank(ckey, c)[r];
        best = v < best ? v : best;
      }
      rowk[r] = (int)(best & 0xFFFFFFFFull);
    }
    __syncthreads();

    // ---- statistics (the row block's rows are dealt round-robin to the 8 CTAs of the cluster) ----
    for (int r = slice + CS * (tid / (D / 4)); r < rows; r += CS * (NT / (D / 4))) {
      const int q = tid % (D / 4);
      const int k = rowk[r];
      if (q == 0) { atomicAdd(cnt + k, 1.0f); a.idx[(long long)s * N + row0 + r] = k; }
      if (a.training_ema) {
        const float* p = R + r * LDR + 4 * q;
        red_add_v4(dw + (size_t)k * D + 4 * q, p[0], p[1], p[2], p[3]);
      }
    }

    float n_stage = 0.f;
    VQ_STAMP(1 + s * 8 + 3);
    if (a.training_ema) {
      if (a.world > 1) grid_barrier_world(a, barrier, generation, nblocks, a.epoch0 + (unsigned)s);
      else grid_barrier(barrier, generation, nblocks);
      VQ_STAMP(1 + s * 8 + 4);     // the statistics of stage s are complete (also orders ckey reuse)
      {
        const float* cs = a.cs[s];
        double part = 0.0;
        for (int k = tid; k < K; k += NT) {
          const float v = fmaf(sum_ranks(a, stats_offset(a, s) + (long long)K * D + k), a.one_minus_decay, __fmul_rn(__ldcg(cs + k), a.decay));
          csn[k] = v;
          part += (double)v;
        }
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid < 32) {
          double v = tid < NW ? red[tid] : 0.0;
          v = warp_sum(v);
          if (tid == 0) s_n = (float)v;
        }
        __syncthreads();
        n_stage = s_n;
      }
    } else {
      cluster.sync();                                 // eval / standard VQ: only ckey reuse needs ordering
    }

    VQ_STAMP(1 + s * 8 + 5);
    // ---- gather (post-update codeword), straight-through value, loss sum, running sum, next residual ----
    float part = 0.f;
    {
      // one item = 4 consecutive dims of one row (rows <= 64 -> at most NT * 4 items: a single batch); the codeword
      // pieces of every item are requested before the first one is used
      constexpr int U = 1024 / NT;
      const int items = rows * (D / 4);
      const float4* Eg4 = reinterpret_cast<const float4*>(a.E[s]);
      const long long dw4_off = stats_offset(a, s) / 4;
      const float4* w4 = reinterpret_cast<const float4*>(a.w[s]);
      float4 wv[U], dv[U]; float clv[U]; int rr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int it = tid + u * NT;
        rr[u] = -1; clv[u] = 1.f;
        wv[u] = make_float4(0.f, 0.f, 0.f, 0.f); dv[u] = wv[u];
        if (it < items) {
          const int r = it / (D / 4), qd = it - r * (D / 4);
          const int k = rowk[r];
          rr[u] = r * LDR + 4 * qd;
          if (a.training_ema) {
            dv[u] = sum_ranks4(a, dw4_off + (long long)k * (D / 4) + qd);
            wv[u] = __ldcg(w4 + (size_t)k * (D / 4) + qd);
            clv[u] = __fmul_rn(__fdiv_rn(__fadd_rn(csn[k], a.eps), __fadd_rn(n_stage, a.k_eps[s])), n_stage);
          } else {
            wv[u] = __ldcg(Eg4 + (size_t)k * (D / 4) + qd);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (rr[u] < 0) continue;
        float q[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
        if (a.training_ema) {
          const float d[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) q[e] = __fdiv_rn(fmaf(d[e], a.one_minus_decay, __fmul_rn(q[e], a.decay)), clv[u]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = R[rr[u] + e];
          const float diff = __fsub_rn(q[e], x);
          const float st = __fadd_rn(x, diff);
          part = fmaf(diff, diff, part);
          O[rr[u] + e] = __fadd_rn(O[rr[u] + e], st);
          R[rr[u] + e] = __fsub_rn(x, st);
        }
      }
    }
    {
      double p = warp_sum((double)part);
      __syncthreads();
      if (lane == 0) red[warp] = p;
      __syncthreads();
      if (tid < 32 && slice == 0) {                   // every slice computed the same sum: one of them reports it
        double v = tid < NW ? red[tid] : 0.0;
        v = warp_sum(v);
        if (tid == 0 && v != 0.0) atomicAdd(a.sse + s, v);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < small::MAX_S; ++i) if (i == s) n_of[i] = n_stage;
    VQ_STAMP(1 + s * 8 + 6);
  }

  // ---- output: the row block's rows are dealt to the 8 CTAs of the cluster ----
  for (int i = tid; i < rows * D; i += NT) {
    const int r = i / D, c = i - r * D;
    if ((r & (CS - 1)) != slice) continue;
    const long long n = row0 + r;
    const long long b = n / T; const int t = (int)(n - b * T);
    a.out[(b * C + c) * T + t] = O[r * LDR + c];
  }
  grid_barrier(barrier, generation, nblocks);
  VQ_STAMP(61);
  // every CTA has gathered with every stage's pre-update state: the in-place updates of ALL stages run here, off the
  // per-stage critical path, as one flattened pass (a single L2 round trip instead of one per stage).  Stage sp's codes
  // are split evenly over the CTAs; item = one float4 of one code.
  if (a.training_ema) {
    int ioff[small::MAX_S + 1];
    ioff[0] = 0;
#pragma unroll
    for (int sp = 0; sp < small::MAX_S; ++sp) {
      int cnt_codes = 0;
      if (sp < a.S) {
        const int per = (a.K[sp] + (int)nblocks - 1) / (int)nblocks;
        const int k0 = min(a.K[sp], cta * per);
        cnt_codes = min(a.K[sp], k0 + per) - k0;
      }
      ioff[sp + 1] = ioff[sp] + cnt_codes * (D / 4);
    }
    for (int it = tid; it < ioff[a.S]; it += NT) {
      int sp = 0;
#pragma unroll
      for (int q = 1; q < small::MAX_S; ++q) sp += (q < a.S && it >= ioff[q]) ? 1 : 0;
      const int Kp = a.K[sp];
      const long long so = stats_offset(a, sp);
      const int per = (Kp + (int)nblocks - 1) / (int)nblocks;
      const int k0 = min(Kp, cta * per);
      const int i = it - ioff[sp];
      const int k = k0 + i / (D / 4);
      const float n = n_of[sp];
      const size_t q4 = (size_t)k0 * (D / 4) + i;
      const float4 d0 = sum_ranks4(a, so / 4 + (long long)q4);
      const float4 w0 = __ldcg(reinterpret_cast<const float4*>(a.w[sp]) + q4);
      const float csv = fmaf(sum_ranks(a, so + (long long)Kp * D + k), a.one_minus_decay, __fmul_rn(__ldcg(a.cs[sp] + k), a.decay));
      const float cl = __fmul_rn(__fdiv_rn(__fadd_rn(csv, a.eps), __fadd_rn(n, a.k_eps[sp])), n);
      float4 n0;
      n0.x = fmaf(d0.x, a.one_minus_decay, __fmul_rn(w0.x, a.decay)); n0.y = fmaf(d0.y, a.one_minus_decay, __fmul_rn(w0.y, a.decay));
      n0.z = fmaf(d0.z, a.one_minus_decay, __fmul_rn(w0.z, a.decay)); n0.w = fmaf(d0.w, a.one_minus_decay, __fmul_rn(w0.w, a.decay));
      reinterpret_cast<float4*>(a.w[sp])[q4] = n0;
      reinterpret_cast<float4*>(a.E[sp])[q4] = make_float4(__fdiv_rn(n0.x, cl), __fdiv_rn(n0.y, cl), __fdiv_rn(n0.z, cl), __fdiv_rn(n0.w, cl));
    }
    __syncthreads();                                  // every thread has read the old cluster sizes of this CTA's codes
    for (int it = tid; it < ioff[a.S] / (D / 4); it += NT) {
      int sp = 0;
#pragma unroll
      for (int q = 1; q < small::MAX_S; ++q) sp += (q < a.S && it >= ioff[q] / (D / 4)) ? 1 : 0;
      const int Kp = a.K[sp];
      const int per = (Kp + (int)nblocks - 1) / (int)nblocks;
      const int k = min(Kp, cta * per) + (it - ioff[sp] / (D / 4));
      a.cs[sp][k] = fmaf(sum_ranks(a, stats_offset(a, sp) + (long long)Kp * D + k), a.one_minus_decay, __fmul_rn(__ldcg(a.cs[sp] + k), a.decay));
    }
  }
  VQ_STAMP(62);

  // ---- loss / perplexity / dcr: CTA s reports stage s (all threads, one block reduction) ----
  if (cta < a.S) {
    const int s = cta;
    const int K = a.K[s];
    const long long cnt_off = stats_offset(a, s) + (long long)K * D;
    const float Nf = (float)a.n_total;            // perplexity / dcr over the vectors of ALL ranks
    const double sse_s = (tid == 0) ? __ldcg(a.sse + s) : 0.0;
    double ent = 0.0; int active = 0;
    for (int k = tid; k < K; k += NT) {
      const float c1 = sum_ranks(a, cnt_off + k);
      const float p = __fdiv_rn(c1, Nf);
      ent += (double)__fmul_rn(p, logf(__fadd_rn(p, 1e-10f)));
      active += (c1 > 0.f);
    }
    ent = warp_sum(ent);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) active += __shfl_xor_sync(0xffffffffu, active, o);
    __shared__ int red_i[NW];
    if (lane == 0) { red[warp] = ent; red_i[warp] = active; }
    __syncthreads();
    if (tid == 0) {
      double e = 0.0; int act = 0;
      for (int w = 0; w < NW; ++w) { e += red[w]; act += red_i[w]; }
      const float mse = (float)(sse_s / ((double)N * D));
      a.m3[s * 3 + 0] = a.use_ema ? __fmul_rn(a.commitment, mse) : __fadd_rn(mse, __fmul_rn(a.commitment, mse));
      a.m3[s * 3 + 1] = expf(-(float)e);
      a.m3[s * 3 + 2] = __fsub_rn(1.0f, __fdiv_rn((float)act, (float)K));
      // the last of the S reporting CTAs adds row S: sum of the stage losses, mean perplexity, mean dcr
      __threadfence();
      if (atomicAdd(barrier + 1, 1u) == (unsigned)a.S - 1u) {
        __threadfence();
        float l = 0.f, p = 0.f, d = 0.f;
        for (int q = 0; q < a.S; ++q) {
          l = __fadd_rn(l, __ldcg(a.m3 + q * 3)); p = __fadd_rn(p, __ldcg(a.m3 + q * 3 + 1)); d = __fadd_rn(d, __ldcg(a.m3 + q * 3 + 2));
        }
        a.m3[a.S * 3 + 0] = l;
        a.m3[a.S * 3 + 1] = __fdiv_rn(p, (float)a.S);
        a.m3[a.S * 3 + 2] = __fdiv_rn(d, (float)a.S);
      }
    }
  }
  __syncthreads();
  VQ_STAMP(63);
  if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0) {
    printf("rvq_wide: %lld SM cycles in %llu ns; stamps (ns since stamp 0):", (long long)(clock64() - cyc0), gtime() - stamp[0]);
    for (int i = 1; i < 64; ++i) if (stamp[i]) printf(" [%d]=%llu", i, stamp[i] - stamp[0]);
    printf("\n");
  }
  cluster.sync();                                     // no CTA exits while its shared memory may still be read
}

constexpr size_t SMEM_BYTES = ((size_t)2 * MAX_RPB * LDR + (size_t)MAX_SLICE_CODES * (LDE + 1) + small::MAX_K + MAX_RPB) * sizeof(float) +
                              MAX_RPB * sizeof(int) + MAX_RPB * sizeof(unsigned long long) + 16;

}  // namespace wide
}  // namespace vqb200

using namespace vqb200;

extern "C" {

int vqb200_rvq_small_eligible(int64_t N, int64_t D, int32_t S, const int64_t* K) {
  if (N < 1 || N > (int64_t)small::CLUSTER * small::MAX_ROWS_PER_CTA || D != small::D || S < 1 || S > small::MAX_S) return 0;
  for (int s = 0; s < S; ++s) if (K[s] < 1 || K[s] > small::MAX_K) return 0;
  return 1;
}

size_t vqb200_rvq_small_workspace_floats(int32_t S, const int64_t* K) {
  size_t n = 0;
  for (int s = 0; s < S; ++s) n += (size_t)small::stage_stats_floats(K[s]) + (size_t)K[s] + 8;
  return n + 16;
}

}  // extern "C"

namespace vqb200 {
// number of co-resident 8-CTA clusters of the wide kernel on the current device (0: no cooperative launch / disabled)
static int wide_row_blocks(int* out) {
  static const int wide_off = [] { const char* e = getenv("VQB200_RVQ_SMALL_NARROW"); return e ? atoi(e) : 0; }();
  // per device: 0 = not probed, 1 = no cooperative launch, 2 + n = n co-resident clusters (GPC floor-planning decides)
  static PerDevice probe_;
  std::atomic<size_t>& probe = probe_.here();
  if (probe.load() == 0) {
    int dev = 0, v = 0;
    const bool coop = cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v;
    if (!coop) { cudaGetLastError(); probe.store(1); }
    else {
      VQ_CUDA(cudaFuncSetAttribute(wide::rvq_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide::SMEM_BYTES));
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(wide::MAX_RB * wide::CS); q.blockDim = dim3(wide::NT); q.dynamicSmemBytes = wide::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = wide::CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, wide::rvq_wide_kernel, &q) != cudaSuccess) { cudaGetLastError(); nc = 0; }
      int max_rb = nc > wide::MAX_RB ? wide::MAX_RB : (nc < 0 ? 0 : nc);
      if (const char* e = getenv("VQB200_RVQ_WIDE_MAXRB")) max_rb = std::min(max_rb, atoi(e));     // development knob
      if (getenv("VQB200_RVQ_WIDE_VERBOSE")) fprintf(stderr, "vqb200: rvq_wide: %d co-resident 8-CTA clusters, using up to %d\n", nc, max_rb);
      probe.store(2 + (size_t)max_rb);
    }
  }
  *out = (wide_off || probe.load() < 2) ? 0 : (int)probe.load() - 2;
  return VQB200_OK;
}

static bool wide_shape_ok(int max_rb, long long N, int32_t S, const int64_t* K) {
  long long slice_codes = 0;
  for (int s = 0; s < S; ++s) slice_codes += (K[s] + wide::CS - 1) / wide::CS;
  return max_rb >= 4 && N <= (long long)max_rb * wide::MAX_RPB && slice_codes <= wide::MAX_SLICE_CODES;
}

static int rvq_small_forward_impl(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                             int32_t S, float* const* E, float* const* ema_cluster_size, float* const* ema_w,
                             const int64_t* K, double decay, double eps, float commitment_cost, int use_ema,
                             int training, float* workspace, double* sse, int32_t* idx, float* out, float* m3,
                             const float* const* peer_stats, uint32_t* const* peer_flags, int rank, int world,
                             uint32_t epoch0, int64_t n_total, vqb200_stream_t stream_) {
  using namespace small;
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(z && E && K && workspace && sse && idx && out && m3, VQB200_EINVAL, "rvq_small_forward: null pointer");
  VQ_CHECK_ARG(vqb200_rvq_small_eligible(B * T, C, S, K), VQB200_EUNSUPPORTED,
               "rvq_small_forward: shape not eligible (needs D == 64, N <= 4096, K <= 4096, S <= 8)");
  const int train_ema = (training && use_ema) ? 1 : 0;
  VQ_CHECK_ARG(!train_ema || (ema_cluster_size && ema_w), VQB200_EINVAL, "rvq_small_forward: EMA buffers required");
  Args a;
  a.z = make_zview(z, B, C, T, sB, sC, sT);
  a.S = S; a.training_ema = train_ema; a.use_ema = use_ema ? 1 : 0; a.commitment = commitment_cost;
  a.decay = (float)decay; a.one_minus_decay = (float)(1.0 - decay); a.eps = (float)eps;
  size_t stats_floats = 0;
  for (int s = 0; s < S; ++s) {
    VQ_CHECK_ARG(E[s] && (reinterpret_cast<uintptr_t>(E[s]) & 15) == 0, VQB200_EALIGN, "rvq_small_forward: codebook %d must be 16-byte aligned", s);
    a.E[s] = E[s];
    a.cs[s] = train_ema ? ema_cluster_size[s] : nullptr;
    a.w[s] = train_ema ? ema_w[s] : nullptr;
    VQ_CHECK_ARG(!train_ema || (a.cs[s] && a.w[s]), VQB200_EINVAL, "rvq_small_forward: EMA buffers of stage %d missing", s);
    VQ_CHECK_ARG(!train_ema || (reinterpret_cast<uintptr_t>(a.w[s]) & 15) == 0, VQB200_EALIGN, "rvq_small_forward: ema_w %d must be 16-byte aligned", s);
    a.K[s] = (int)K[s];
    a.k_eps[s] = (float)((double)K[s] * eps);
    stats_floats += (size_t)stage_stats_floats(K[s]);
  }
  for (int s = S; s < MAX_S; ++s) { a.E[s] = nullptr; a.cs[s] = nullptr; a.w[s] = nullptr; a.K[s] = 0; a.k_eps[s] = 0.f; }
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VQB200_EALIGN, "rvq_small_forward: workspace must be 16-byte aligned");
  a.stats = workspace;
  a.scratch = workspace + ((stats_floats + 3) & ~(size_t)3);
  a.sse = sse; a.idx = idx; a.out = out; a.m3 = m3;
  static const int stamps = [] { const char* e = getenv("VQB200_RVQ_STAMPS"); return e ? atoi(e) : 0; }();
  a.stamps = stamps;
  for (int p = 0; p < VQB200_MAX_PEERS; ++p) { a.stats_of[p] = nullptr; a.flags_of[p] = nullptr; }
  a.world = 1; a.rank = 0; a.epoch0 = 0; a.n_total = B * T; a.stats_of[0] = a.stats;
  static const int peer_timeout = [] { const char* e = getenv("VQB200_PEER_TIMEOUT_S"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 120; }();
  a.peer_timeout_s = peer_timeout;
  if (world > 1) {
    VQ_CHECK_ARG(train_ema, VQB200_EINVAL, "rvq_small_forward_peer: only the EMA training step exchanges statistics");
    VQ_CHECK_ARG(peer_stats && peer_flags && world <= VQB200_MAX_PEERS && rank >= 0 && rank < world, VQB200_ESHAPE,
                 "rvq_small_forward_peer: bad rank %d / world %d", rank, world);
    for (int p = 0; p < world; ++p) {
      VQ_CHECK_ARG(peer_stats[p] && peer_flags[p] && (reinterpret_cast<uintptr_t>(peer_stats[p]) & 15) == 0, VQB200_EALIGN,
                   "rvq_small_forward_peer: slot of rank %d missing or not 16-byte aligned", p);
      a.stats_of[p] = peer_stats[p];
      a.flags_of[p] = reinterpret_cast<unsigned*>(peer_flags[p]);
    }
    a.world = world; a.rank = rank; a.epoch0 = epoch0; a.n_total = n_total;
    a.stats = const_cast<float*>(peer_stats[rank]);       // this rank accumulates straight into its own slot
  }

  // ---- wide variant: the whole GPU instead of one GPC (see namespace wide) ----
  {
    int max_rb = 0;
    { const int rc = wide_row_blocks(&max_rb); if (rc != VQB200_OK) return rc; }
    const bool wide_ok = wide_shape_ok(max_rb, B * T, S, K);
    VQ_CHECK_ARG(world == 1 || wide_ok, VQB200_EUNSUPPORTED,
                 "rvq_small_forward_peer: only the whole-GPU variant runs under data parallelism (see vqb200_rvq_small_peer_eligible)");
    if (wide_ok) {
      const int rb_used = (int)std::min<long long>(max_rb, (B * T + 15) / 16);            // at least 16 rows per block
      size_t scratch_floats = 0;
      for (int s = 0; s < S; ++s) scratch_floats += (size_t)K[s] + 8;
      unsigned* barrier = reinterpret_cast<unsigned*>(a.scratch + scratch_floats);       // inside the +16 slack
      const size_t zero_bytes = (size_t)((reinterpret_cast<unsigned char*>(barrier) + 16) - reinterpret_cast<unsigned char*>(workspace));
      VQ_CUDA(cudaMemsetAsync(workspace, 0, zero_bytes, stream));                        // statistics + barrier words
      if (world > 1) VQ_CUDA(cudaMemsetAsync(a.stats, 0, stats_floats * sizeof(float), stream));   // ... of the peer slot
      VQ_CUDA(cudaMemsetAsync(sse, 0, (size_t)S * sizeof(double), stream));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((rb_used < 1 ? 1 : rb_used) * wide::CS);
      cfg.blockDim = dim3(wide::NT);
      cfg.dynamicSmemBytes = wide::SMEM_BYTES;
      cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = wide::CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      // Nsight Compute (2025.2) cannot launch a kernel that is both cooperative and clustered (the launch fails inside
      // the tool).  Under an injected profiler kernels are serialised, so the co-residency the cooperative attribute
      // guarantees follows from the occupancy probe above: launch the same kernel without the attribute there.
      static const bool injected = [] {
        const char* f = getenv("VQB200_RVQ_WIDE_NONCOOP");
        if (f) return atoi(f) != 0;
        return getenv("CUDA_INJECTION64_PATH") != nullptr || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
               getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr;
      }();
      cfg.attrs = attr; cfg.numAttrs = injected ? 1 : 2;
      VQ_CUDA(cudaLaunchKernelEx(&cfg, wide::rvq_wide_kernel, a, barrier));
      VQ_LAUNCH_CHECK("rvq_wide_kernel");
      return VQB200_OK;
    }
  }

  const size_t smem = ((size_t)MAX_ROWS_PER_CTA * LDR + MAX_K + 8 * 32 + (size_t)CHUNK * D) * sizeof(float) +
                      (8 * 32 + MAX_ROWS_PER_CTA) * sizeof(int);
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (!configured.load()) {
    VQ_CUDA(cudaFuncSetAttribute(rvq_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VQ_CUDA(cudaFuncSetAttribute(rvq_small_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured.store(1);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CLUSTER);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  VQ_CUDA(cudaLaunchKernelEx(&cfg, rvq_small_kernel, a));
  VQ_LAUNCH_CHECK("rvq_small_kernel");
  return VQB200_OK;
}
}  // namespace vqb200

extern "C" {

int vqb200_rvq_small_forward(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                             int32_t S, float* const* E, float* const* ema_cluster_size, float* const* ema_w,
                             const int64_t* K, double decay, double eps, float commitment_cost, int use_ema,
                             int training, float* workspace, double* sse, int32_t* idx, float* out, float* m3,
                             vqb200_stream_t stream) {
  return rvq_small_forward_impl(z, B, C, T, sB, sC, sT, S, E, ema_cluster_size, ema_w, K, decay, eps, commitment_cost, use_ema,
                                training, workspace, sse, idx, out, m3, nullptr, nullptr, 0, 1, 0, B * T, stream);
}

int vqb200_rvq_small_peer_eligible(int64_t N, int64_t D, int32_t S, const int64_t* K) {
  if (!vqb200_rvq_small_eligible(N, D, S, K)) return 0;
  int max_rb = 0;
  if (wide_row_blocks(&max_rb) != VQB200_OK) return 0;
  return wide_shape_ok(max_rb, N, S, K) ? 1 : 0;
}

size_t vqb200_rvq_small_stats_floats(int32_t S, const int64_t* K) {
  size_t n = 0;
  for (int s = 0; s < S; ++s) n += (size_t)small::stage_stats_floats(K[s]);
  return n;
}

int vqb200_rvq_small_forward_peer(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                                  int32_t S, float* const* E, float* const* ema_cluster_size, float* const* ema_w,
                                  const int64_t* K, double decay, double eps, float commitment_cost,
                                  float* workspace, double* sse, int32_t* idx, float* out, float* m3,
                                  const float* const* peer_stats, uint32_t* const* peer_flags, int32_t rank, int32_t world,
                                  uint32_t epoch0, int64_t n_total, vqb200_stream_t stream) {
  VQ_CHECK_ARG(world >= 1, VQB200_ESHAPE, "rvq_small_forward_peer: bad world %d", world);
  return rvq_small_forward_impl(z, B, C, T, sB, sC, sT, S, E, ema_cluster_size, ema_w, K, decay, eps, commitment_cost, 1, 1,
                                workspace, sse, idx, out, m3, peer_stats, peer_flags, rank, world, epoch0, n_total, stream);
}

}  // extern "C"
