// vqb200 K1 (CUDA-core variant): fused distance + argmin in exact fp32.
//
// Replaces models/vqvae.py:30-38 of the reference (permute/contiguous, |x|^2, |E|^2, the N x K
// matmul, the two broadcasts and argmin) without ever materialising the N x K matrix.  This is the
// exactness anchor of the engine: the tcgen05 path (assign_tc.cu) only *filters* candidates and
// re-ranks them with the arithmetic below, and falls back to this kernel (through a row list) for
// rows whose filter margin is not rigorous.
//
//   d[n,k] = fl(fl(|x_n|^2 + |E_k|^2) - 2*(x_n . E_k))   ;  idx[n] = first argmin_k d[n,k], NaN wins
//
// Tiling: CTA = 64 rows x 128 codes, 256 threads, 4x8 register tile per thread, z tile resident in
// shared memory for the whole codebook sweep, codebook streamed in 128x16 chunks (register prefetch).
// Short work lists (the latency configs: a few hundred unproven rows) would occupy a handful of SMs for a whole
// codebook sweep each; there the sweep is split over `splits` CTAs per row tile, which merge through one 64-bit
// atomicMin per row on a key (ordered distance bits << 32 | index) that encodes exactly the cand_better order.
#include "common.cuh"
#include <limits.h>

namespace vqb200 {

namespace simt {
constexpr int BM = 64, BN = 128, BK = 16, TM = 4, TN = 8;
constexpr int NT = (BM / TM) * (BN / TN);      // 256
constexpr int LDZ = BM + 4;
constexpr int LDE = BN + 4;
}  // namespace simt

__global__ void __launch_bounds__(256)
assign_keys_finalize_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ row_list,
                            const int32_t* __restrict__ row_count, int32_t* __restrict__ idx_out, long long key_cap) {
  const int total = *row_count;
  if (total > key_cap) return;                     // the sweep was not split (see vq_assign_simt_kernel)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    idx_out[row_list[i]] = (int32_t)(keys[i] & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(simt::NT, 2)
vq_assign_simt_kernel(ZView z, const float* __restrict__ E, const float* __restrict__ ee,
                      int K, int D, int Dp, int32_t* __restrict__ idx_out, float* __restrict__ best_out,
                      const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count,
                      int splits_req, unsigned long long* __restrict__ keys_req, long long key_cap) {
  using namespace simt;
  extern __shared__ __align__(16) float smem[];
  float* zs = smem;                    // [Dp][LDZ]   k-major z tile
  float* es = zs + (size_t)Dp * LDZ;   // [BK][LDE]   k-major codebook chunk
  float* xx = es + BK * LDE;           // [BM]
  float* ees = xx + BM;                // [BN]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long total = row_list ? (long long)(*row_count) : z.N;
  // the merge keys are indexed by list position and there are only key_cap of them: a longer list (known only on the
  // device) is swept unsplit, which is what a long list wants anyway
  const int splits = (total <= key_cap) ? splits_req : 1;
  unsigned long long* keys = (splits > 1) ? keys_req : nullptr;
  const long long ntiles = (total + BM - 1) / BM;
  const bool vecE = ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(E) & 15) == 0);

  const int Kc = (K + splits - 1) / splits;            // codes per split, a multiple of BN when splits > 1
  for (long long w = blockIdx.x; w < ntiles * splits; w += gridDim.x) {
    const long long tile = w / splits;
    const int c_begin = (int)(w - tile * splits) * Kc, c_end = min(K, c_begin + Kc);
    const long long n0 = tile * BM;
    const int rows = (int)min((long long)BM, total - n0);

    __syncthreads();                                   // previous tile fully consumed
    for (int i = tid; i < Dp * LDZ; i += NT) zs[i] = 0.f;
    __syncthreads();
    if (row_list) {
      for (int i = tid; i < rows * D; i += NT) {
        int r = i / D, k = i - r * D;
        long long n = row_list[n0 + r];
        zs[k * LDZ + r] = __ldg(z.p + z.row_base(n) + (long long)k * z.sC);
      }
    } else {
      load_rows(z, n0, rows, D, tid, NT, [&](int r, int k, float v) { zs[k * LDZ + r] = v; });
    }
    __syncthreads();
    if (tid < BM) {
      float s = 0.f;
      for (int k = 0; k < D; ++k) { float v = zs[k * LDZ + tid]; s = fmaf(v, v, s); }
      xx[tid] = s;
    }

    float best[TM]; int bidx[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) { best[i] = INFINITY; bidx[i] = INT_MAX; }

    for (int c0 = c_begin; c0 < c_end; c0 += BN) {
      float acc[TM][TN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

      // register prefetch of the first chunk
      float4 pre[2];
      auto fetch = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = (tid >> 2) + 64 * h, q = tid & 3;
          const int gc = c0 + c, gk = k0 + 4 * q;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gc < K) {
            const float* src = E + (size_t)gc * D + gk;
            if (vecE) { if (gk < D) v = __ldg(reinterpret_cast<const float4*>(src)); }
            else {
              if (gk + 0 < D) v.x = __ldg(src + 0);
              if (gk + 1 < D) v.y = __ldg(src + 1);
              if (gk + 2 < D) v.z = __ldg(src + 2);
              if (gk + 3 < D) v.w = __ldg(src + 3);
            }
          }
          pre[h] = v;
        }
      };
      auto stash = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = (tid >> 2) + 64 * h, q = tid & 3;
          es[(4 * q + 0) * LDE + c] = pre[h].x;
          es[(4 * q + 1) * LDE + c] = pre[h].y;
          es[(4 * q + 2) * LDE + c] = pre[h].z;
          es[(4 * q + 3) * LDE + c] = pre[h].w;
        }
      };
      fetch(0);
      for (int k0 = 0; k0 < Dp; k0 += BK) {
        __syncthreads();                               // es free (and xx/ees of the previous tile consumed)
        stash();
        if (k0 == 0 && tid < BN) ees[tid] = (c0 + tid < K) ? __ldg(ee + c0 + tid) : INFINITY;
        __syncthreads();
        if (k0 + BK < Dp) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(&zs[(k0 + kk) * LDZ + ty * TM]);
          const float4 b0 = *reinterpret_cast<const float4*>(&es[kk * LDE + tx * 4]);
          const float4 b1 = *reinterpret_cast<const float4*>(&es[kk * LDE + 64 + tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
      }
      // epilogue: distances of this 64x128 block, running argmin in registers
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float xr = xx[ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int cl = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
          const int c = c0 + cl;
          const float d = __fsub_rn(__fadd_rn(xr, ees[cl]), __fmul_rn(2.0f, acc[i][j]));
          if (c < K && cand_better(d, c, best[i], bidx[i])) { best[i] = d; bidx[i] = c; }
        }
      }
    }
    // combine the 16 column-threads of each row (they sit in one half-warp)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best[i], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
        if (cand_better(od, oi, best[i], bidx[i])) { best[i] = od; bidx[i] = oi; }
      }
      const int r = ty * TM + i;
      if (tx == 0 && r < rows) {
        if (keys) {
          atomicMin(keys + n0 + r, cand_key(best[i], bidx[i]));
        } else {
          const long long n = row_list ? (long long)row_list[n0 + r] : n0 + r;
          idx_out[n] = bidx[i];
          if (best_out) best_out[n] = best[i];
        }
      }
    }
  }
}

size_t assign_simt_smem_bytes(int D) {
  using namespace simt;
  const int Dp = (D + BK - 1) / BK * BK;
  return ((size_t)Dp * LDZ + (size_t)BK * LDE + BM + BN) * sizeof(float);
}

// row_list == nullptr: all rows of the view.  Otherwise *row_count rows listed in row_list.
// keys: key_cap 64-bit merge keys (or null): lists of up to key_cap rows are swept split over the codes.
int launch_assign_simt_capped(const ZView& z, const float* E, const float* ee, int K, int D,
                              int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                              long long max_rows, cudaStream_t stream, unsigned long long* keys, long long key_cap) {
  using namespace simt;
  const size_t smem = assign_simt_smem_bytes(D);
  VQ_CHECK_ARG(smem <= 227 * 1024, VQB200_ESHAPE, "vq_assign(SIMT): D=%d needs %zu B of shared memory (max 232448)", D, smem);
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (smem > 48 * 1024 && smem > configured.load()) {
    VQ_CUDA(cudaFuncSetAttribute(vq_assign_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.store(smem);
  }
  const int Dp = (D + BK - 1) / BK * BK;
  const long long tiles = (max_rows + BM - 1) / BM;
  // split the codebook sweep when the caller provides the merge keys (short work lists, see the header comment)
  int splits = 1;
  if (keys && row_list && !best && key_cap > 0) {
    while (splits < 8 && (K / (splits * 2)) >= BN && (K % (splits * 2 * BN)) == 0) splits *= 2;
  }
  key_cap = min(key_cap, max_rows);
  if (splits > 1) VQ_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t)key_cap * sizeof(unsigned long long), stream));
  else { keys = nullptr; key_cap = 0; }
  const int grid = (int)max(1LL, min(tiles * splits, (long long)sm_count() * 2));
  vq_assign_simt_kernel<<<grid, NT, smem, stream>>>(z, E, ee, K, D, Dp, idx, best, row_list, row_count, splits, keys, key_cap);
  VQ_LAUNCH_CHECK("vq_assign_simt_kernel");
  if (splits > 1) {
    assign_keys_finalize_kernel<<<(int)max(1LL, min((key_cap + 255) / 256, (long long)sm_count())), 256, 0, stream>>>(
        keys, row_list, row_count, idx, key_cap);
    VQ_LAUNCH_CHECK("assign_keys_finalize_kernel");
  }
  return VQB200_OK;
}

int launch_assign_simt(const ZView& z, const float* E, const float* ee, int K, int D,
                       int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                       long long max_rows, cudaStream_t stream, unsigned long long* keys) {
  return launch_assign_simt_capped(z, E, ee, K, D, idx, best, row_list, row_count, max_rows, stream, keys, max_rows);
}

}  // namespace vqb200
