// vqb200 K3: EMA statistics (scatter-add) + EMA finalize, codebook-derived state, histogram,
// standard-VQ codebook gradient.  Replaces models/vqvae.py:44-50 and :35 of the reference.
#include "common.cuh"
#include <cuda_fp16.h>
#include "codebook.cuh"

namespace vqb200 {

int try_accumulate_tile(const ZView& z, const int32_t* idx, const float* E, int K, float* dw, float* cnt, int mode,
                        cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// codebook_prepare: ee[k] = sum_c E[k,c]^2, bf16 tile image, info.
// one warp per code row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
codebook_prepare_kernel(const float* __restrict__ E, int K, int D, float* __restrict__ ee,
                        unsigned char* __restrict__ image, float* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Kp = (int)img_kp(K), Dp = (int)img_dp(D);
  float* ee_img = image ? reinterpret_cast<float*>(image + img_tiles_bytes(K, D)) : nullptr;
  for (int k = warp; k < Kp; k += nwarps) {
    float s = 0.f;
    bool bad = false;
    for (int c = lane; c < Dp; c += 32) {
      float v = (k < K && c < D) ? __ldg(E + (size_t)k * D + c) : 0.f;
      s = fmaf(v, v, s);
      bad |= !(fabsf(v) <= 3.0e38f);
      if (image) img_store(image, k, c, Dp, v);
    }
    s = warp_sum(s);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (k < K) { ee[k] = s; info_update(info, s, bad); }
      if (ee_img) ee_img[k] = (k < K) ? -0.5f * s : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------
// ema_accumulate: cnt[k] += 1, dw[k,:] += x_n (mode 0) or (E[k,:] - x_n) (mode 1) for k = idx[n].
// CTA stages 64 rows in shared memory (coalesced for every layout), then each warp pushes rows with
// 16-byte vector reductions (red.global.add.v4.f32): one L2 atomic transaction per 4 floats.
// Counts go through a shared-memory histogram when K is small.
// ------------------------------------------------------------------------------------------
constexpr int ACC_BM = 64;
constexpr int ACC_NT = 256;
constexpr int ACC_HIST_MAX = 8192;

__global__ void __launch_bounds__(ACC_NT)
ema_accumulate_kernel(ZView z, const int32_t* __restrict__ idx, const float* __restrict__ E,
                      int K, int D, float* __restrict__ dw, float* __restrict__ cnt, int mode, int use_hist) {
  extern __shared__ __align__(16) float smem[];
  const int LD = D + 4;
  float* tile = smem;                                    // [ACC_BM][LD]
  int* hist = reinterpret_cast<int*>(tile + ACC_BM * LD);    // [K] if use_hist
  __shared__ int s_idx[ACC_BM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (use_hist) for (int k = tid; k < K; k += ACC_NT) hist[k] = 0;
  const long long ntiles = (z.N + ACC_BM - 1) / ACC_BM;
  const bool vec = ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) &&
                   (mode == 0 || (reinterpret_cast<uintptr_t>(E) & 15) == 0);
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long n0 = t * ACC_BM;
    const int rows = (int)min((long long)ACC_BM, z.N - n0);
    __syncthreads();
    load_rows(z, n0, rows, D, tid, ACC_NT, [&](int r, int k, float v) { tile[r * LD + k] = v; });
    if (tid < rows) {
      int k = __ldg(idx + n0 + tid);
      s_idx[tid] = k;
      if ((unsigned)k < (unsigned)K) {
        if (use_hist) atomicAdd(hist + k, 1); else atomicAdd(cnt + k, 1.0f);
      }
    }
    __syncthreads();
    if (vec) {
      const int d4 = D >> 2;
      int lpr = 1; while (lpr < d4 && lpr < 32) lpr <<= 1;       // lanes per row (power of two <= 32)
      const int rpw = 32 / lpr;                                  // rows per warp pass
      const int sub = lane / lpr, l = lane % lpr;
      for (int r = warp * rpw + sub; r < rows; r += (ACC_NT / 32) * rpw) {
        const int k = s_idx[r];
        if ((unsigned)k >= (unsigned)K) continue;
        for (int q = l; q < d4; q += lpr) {
          float4 v = *reinterpret_cast<const float4*>(tile + r * LD + 4 * q);
          if (mode == 1) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + q);
            v = make_float4(e.x - v.x, e.y - v.y, e.z - v.z, e.w - v.w);
          }
          red_add_v4(dw + (size_t)k * D + 4 * q, v.x, v.y, v.z, v.w);
        }
      }
    } else {
      for (int i = tid; i < rows * D; i += ACC_NT) {
        const int r = i / D, c = i - r * D;
        const int k = s_idx[r];
        if ((unsigned)k >= (unsigned)K) continue;
        float v = tile[r * LD + c];
        if (mode == 1) v = __ldg(E + (size_t)k * D + c) - v;
        atomicAdd(dw + (size_t)k * D + c, v);
      }
    }
  }
  if (use_hist) {
    __syncthreads();
    for (int k = tid; k < K; k += ACC_NT) { int h = hist[k]; if (h) atomicAdd(cnt + k, (float)h); }
  }
}

// histogram of indices only (eval / non-EMA metrics)
__global__ void __launch_bounds__(256)
histogram_kernel(const int32_t* __restrict__ idx, long long N, int K, float* __restrict__ cnt, int use_hist) {
  extern __shared__ int hist[];
  if (use_hist) { for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0; __syncthreads(); }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    int k = __ldg(idx + i);
    if ((unsigned)k < (unsigned)K) { if (use_hist) atomicAdd(hist + k, 1); else atomicAdd(cnt + k, 1.0f); }
  }
  if (use_hist) {
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) { int h = hist[k]; if (h) atomicAdd(cnt + k, (float)h); }
  }
}

// ------------------------------------------------------------------------------------------
// ema_finalize, step 1 (one CTA): cs <- decay*cs + (1-decay)*cnt ; n = sum(cs) ;
// scratch[k] = (cs_k + eps)/(n + K*eps)*n ; scratch[K] = n          models/vqvae.py:46,48-49
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ema_finalize_cs_kernel(const float* __restrict__ cnt, float* __restrict__ cs, int K,
                       float decay, float one_minus_decay, float eps, float k_eps,
                       float* __restrict__ scratch, float* __restrict__ info) {
  __shared__ double red[32];
  __shared__ float s_n;
  double part = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = fmaf(cnt[k], one_minus_decay, __fmul_rn(cs[k], decay));
    cs[k] = v;
    part += (double)v;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) s_n = (float)v;
  }
  __syncthreads();
  const float n = s_n;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float c = cs[k];
    scratch[k] = __fmul_rn(__fdiv_rn(__fadd_rn(c, eps), __fadd_rn(n, k_eps)), n);
  }
  if (threadIdx.x == 0) {
    scratch[K] = n;
    if (info) { info[0] = 0.f; info[1] = 0.f; info[2] = INFO2_RESET; info[3] = 0.f; }
  }
}

// step 2 (grid): w <- decay*w + (1-decay)*dw ; E <- w / cluster ; refresh ee / image / info
__global__ void __launch_bounds__(256)
ema_finalize_w_kernel(const float* __restrict__ dw, float* __restrict__ w, float* __restrict__ E,
                      int K, int D, float decay, float one_minus_decay, const float* __restrict__ cluster,
                      float* __restrict__ ee, unsigned char* __restrict__ image, float* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Kp = (int)img_kp(K), Dp = (int)img_dp(D);
  float* ee_img = image ? reinterpret_cast<float*>(image + img_tiles_bytes(K, D)) : nullptr;
  for (int k = warp; k < Kp; k += nwarps) {
    float s = 0.f;
    bool bad = false;
    const float cl = (k < K) ? cluster[k] : 1.f;
    for (int c = lane; c < Dp; c += 32) {
      float e = 0.f;
      if (k < K && c < D) {
        const size_t o = (size_t)k * D + c;
        const float wv = fmaf(dw[o], one_minus_decay, __fmul_rn(w[o], decay));
        w[o] = wv;
        e = __fdiv_rn(wv, cl);
        E[o] = e;
        bad |= !(fabsf(e) <= 3.0e38f);
      }
      s = fmaf(e, e, s);
      if (image) img_store(image, k, c, Dp, e);
    }
    s = warp_sum(s);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (k < K) { if (ee) ee[k] = s; info_update(info, s, bad); }
      if (ee_img) ee_img[k] = (k < K) ? -0.5f * s : -INFINITY;
    }
  }
}

__global__ void __launch_bounds__(256)
backward_codebook_kernel(const float* __restrict__ dw1, long long n, const float* __restrict__ g_loss,
                         float coef, float* __restrict__ gE) {
  const float s = (g_loss ? __ldg(g_loss) : 1.0f) * coef;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gE[i] = s * dw1[i];
}

__global__ void info_reset_kernel(float* info) { if (threadIdx.x < 4) info[threadIdx.x] = (threadIdx.x == 2) ? INFO2_RESET : 0.f; }

// ------------------------------------------------------------------------------------------
// fp16 filter image (codebook.cuh): one CTA per 128-code tile.  Runs after E / ee have been refreshed.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
codebook_image_f16_kernel(const float* __restrict__ E, int K, unsigned char* __restrict__ image, float* __restrict__ info) {
  constexpr int D = IMG_TILE_DIMS;
  __shared__ float s_red[8];
  __shared__ float s_scale;
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k0 = j * IMG_TILE_CODES;
  // 128 codes x 16 float4: thread t owns float4 q = t & 15 of rows (t >> 4) + 16 i
  float4 v[8];
  float m = 0.f;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (tid >> 4) + 16 * i, k = k0 + r;
    v[i] = (k < K) ? __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + (tid & 15)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float a = fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
    bad |= !(a <= 3.0e38f);
    m = fmaxf(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = 0.f;
    for (int w = 0; w < 8; ++w) mm = fmaxf(mm, s_red[w]);
    // s_j = 2^(10 - floor(log2 mm)); tiles whose maximum is below 2^-100 (or not finite) cannot be represented
    const int eb = (int)((__float_as_uint(mm) >> 23) & 255u);
    float scale = 1.0f, inv = 1.0f;
    if (mm != 0.f) {
      if (eb < 27 || eb == 255) { bad = true; }
      else { scale = __uint_as_float((unsigned)(264 - eb) << 23); inv = __uint_as_float((unsigned)(eb - 10) << 23); }
    }
    s_scale = scale;
    float* meta = reinterpret_cast<float*>(image + img_f16_meta_offset(K, D)) + (size_t)j * F16_META_FLOATS;
    meta[IMG_TILE_CODES + 0] = inv; meta[IMG_TILE_CODES + 1] = 0.f; meta[IMG_TILE_CODES + 2] = 0.f; meta[IMG_TILE_CODES + 3] = 0.f;
  }
  bad = __syncthreads_or(bad);
  const float sc = s_scale;
  unsigned char* tile = image + img_f16_offset(K, D) + (size_t)j * F16_TILE_BYTES;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (tid >> 4) + 16 * i, q = tid & 15;      // float4 q covers dims 4q..4q+3: half of 16-byte chunk q >> 1
    const __half2 h0 = __floats2half2_rn(v[i].x * sc, v[i].y * sc);
    const __half2 h1 = __floats2half2_rn(v[i].z * sc, v[i].w * sc);
    uint2 w;
    w.x = *reinterpret_cast<const unsigned*>(&h0);
    w.y = *reinterpret_cast<const unsigned*>(&h1);
    *reinterpret_cast<uint2*>(tile + r * 128 + (((q >> 1) ^ (r & 7)) << 4) + (q & 1) * 8) = w;
  }
  {
    // interleaved fp32 copy for the exact re-rank: thread (r = tid >> 4, q = tid & 15) holds dims 4q..4q+3 of rows r + 16 i
    float4* e4 = reinterpret_cast<float4*>(image + img_e4_offset(K, D)) + (size_t)j * (IMG_TILE_CODES * 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (tid >> 4) + 16 * i, q = tid & 15;
      e4[(size_t)(r >> 2) * 64 + q * 4 + (r & 3)] = v[i];
    }
  }
  float nmin = INFO2_RESET;
  if (tid < IMG_TILE_CODES) {
    const int k = k0 + tid;
    float* meta = reinterpret_cast<float*>(image + img_f16_meta_offset(K, D)) + (size_t)j * F16_META_FLOATS;
    // -|E_k|^2/2 as the finalize / prepare kernel has just written it behind the split tiles (-inf for padding codes)
    const float nh = reinterpret_cast<const float*>(image + img_tiles_bytes(K, D))[k];
    // padding codes: a large FINITE negative (the filter packs a group id into the low mantissa bits of its scores,
    // which would turn -inf into a NaN)
    meta[tid] = (k < K) ? nh : -3.0e38f;
    if (k < K && nh == nh) nmin = sqrtf(-2.0f * nh);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmin = fminf(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
  if (info && lane == 0 && warp < 4) atomicMin(reinterpret_cast<int*>(info + 2), __float_as_int(nmin));
  if (info && tid == 0 && bad) atomicExch(reinterpret_cast<int*>(info + 1), __float_as_int(1.0f));
}

// kernels are loaded lazily and a first load may synchronise the context: callers that are about to spin on a peer
// (peer.cu) load this one up front
int preload_image_f16() {
  cudaFuncAttributes fa;
  VQ_CUDA(cudaFuncGetAttributes(&fa, codebook_image_f16_kernel));
  return VQB200_OK;
}

int launch_image_f16(const float* E, long long K, long long D, void* image, float* info, cudaStream_t stream) {
  if (!image || !img_has_f16(D)) return VQB200_OK;
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(E) & 15) == 0, VQB200_EALIGN, "codebook image: E must be 16-byte aligned");
  codebook_image_f16_kernel<<<(int)(img_kp(K) / IMG_TILE_CODES), 256, 0, stream>>>(E, (int)K, (unsigned char*)image, info);
  VQ_LAUNCH_CHECK("codebook_image_f16_kernel");
  return VQB200_OK;
}

// ------------------------------------------------------------------------------------------
// Codebook health (opt-in, NOT in the reference): codes whose usage fell below a threshold are
// re-seeded from input rows.  The reference suffers 35-90 % dead codes (README.md:353-355) and starts EMA codebooks
// from ema_w ~ N(0,1) with zero counts (models/vqvae.py:24-26); nothing here runs unless the caller asks for it.
// Row choice is a pure function of (seed, k): splitmix64(seed + k * golden) mod N.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
codebook_revive_kernel(ZView z, const float* __restrict__ usage, float threshold, unsigned long long seed,
                       float* __restrict__ E, float* __restrict__ ema_cluster_size, float* __restrict__ ema_w,
                       int K, int D, int32_t* __restrict__ revived) {
  // one warp per code
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int k = warp; k < K; k += nwarps) {
    if (!(usage[k] < threshold)) continue;                  // NaN usage is left alone
    const long long n = (long long)(splitmix64(seed + (unsigned long long)k * 0x9E3779B97F4A7C15ull) % (unsigned long long)z.N);
    const float* src = z.p + z.row_base(n);
    for (int c = lane; c < D; c += 32) {
      const float v = __ldg(src + (long long)c * z.sC);
      E[(size_t)k * D + c] = v;
      if (ema_w) ema_w[(size_t)k * D + c] = v;
    }
    if (lane == 0) {
      if (ema_cluster_size) ema_cluster_size[k] = 1.0f;     // E == ema_w / ema_cluster_size stays consistent
      atomicAdd(revived, 1);
    }
  }
}

}  // namespace vqb200

using namespace vqb200;

extern "C" {

size_t vqb200_codebook_image_bytes(int64_t K, int64_t D) { return img_total_bytes(K, D); }

int vqb200_codebook_prepare(const float* E, int64_t K, int64_t D, float* ee, void* image, float* info,
                            vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(E && ee, VQB200_EINVAL, "codebook_prepare: null pointer");
  VQ_CHECK_ARG(K > 0 && D > 0 && K < (1LL << 30) && D < (1 << 20), VQB200_ESHAPE, "codebook_prepare: bad K=%lld D=%lld", (long long)K, (long long)D);
  VQ_CHECK_ARG(!image || (reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "codebook_prepare: image must be 1024-byte aligned");
  if (info) { info_reset_kernel<<<1, 32, 0, stream>>>(info); VQ_LAUNCH_CHECK("info_reset_kernel"); }
  const long long Kp = img_kp(K);
  const int grid = grid_for(Kp, 8, sm_count() * 8);
  codebook_prepare_kernel<<<grid, 256, 0, stream>>>(E, (int)K, (int)D, ee, (unsigned char*)image, info);
  VQ_LAUNCH_CHECK("codebook_prepare_kernel");
  return launch_image_f16(E, K, D, image, info, stream);
}

int vqb200_ema_accumulate(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                          const int32_t* idx, const float* E, int64_t K, float* stats, int mode,
                          vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(stats && ((z && idx) || B * T == 0), VQB200_EINVAL, "ema_accumulate: null pointer");
  VQ_CHECK_ARG(mode == 0 || (mode == 1 && E), VQB200_EINVAL, "ema_accumulate: mode %d needs E", mode);
  VQ_CHECK_ARG(B >= 0 && C > 0 && T > 0 && K > 0, VQB200_ESHAPE, "ema_accumulate: bad shape");
  const int D = (int)C;
  float* dw = stats;
  float* cnt = stats + (size_t)K * D;
  VQ_CUDA(cudaMemsetAsync(stats, 0, (size_t)K * (D + 1) * sizeof(float), stream));
  if (B * T == 0) return VQB200_OK;
  const ZView zv = make_zview(z, B, C, T, sB, sC, sT);
  {
    const int rc = try_accumulate_tile(zv, idx, E, (int)K, dw, cnt, mode, stream);
    if (rc != 0) return rc == 1 ? VQB200_OK : rc;
  }
  const int use_hist = (K <= ACC_HIST_MAX) ? 1 : 0;
  const size_t smem = (size_t)ACC_BM * (D + 4) * sizeof(float) + (use_hist ? (size_t)K * sizeof(int) : 0);
  VQ_CHECK_ARG(smem <= 227 * 1024, VQB200_ESHAPE, "ema_accumulate: D=%d K=%lld needs %zu B of shared memory", D, (long long)K, smem);
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (smem > 48 * 1024 && smem > configured.load()) {
    VQ_CUDA(cudaFuncSetAttribute(ema_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.store(smem);
  }
  const long long tiles = (zv.N + ACC_BM - 1) / ACC_BM;
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
  const int grid = (int)max(1LL, min(tiles, (long long)sm_count() * per_sm));
  ema_accumulate_kernel<<<grid, ACC_NT, smem, stream>>>(zv, idx, E, (int)K, D, dw, cnt, mode, use_hist);
  VQ_LAUNCH_CHECK("ema_accumulate_kernel");
  return VQB200_OK;
}

int vqb200_ema_finalize(const float* stats, float* ema_cluster_size, float* ema_w, float* E,
                        int64_t K, int64_t D, double decay, double eps, float* ee, void* image, float* info,
                        float* scratch, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(stats && ema_cluster_size && ema_w && E && scratch, VQB200_EINVAL, "ema_finalize: null pointer");
  VQ_CHECK_ARG(K > 0 && D > 0, VQB200_ESHAPE, "ema_finalize: bad K/D");
  VQ_CHECK_ARG(!image || (reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "ema_finalize: image must be 1024-byte aligned");
  const float* dw = stats;
  const float* cnt = stats + (size_t)K * D;
  const float fd = (float)decay, fo = (float)(1.0 - decay), fe = (float)eps, fke = (float)((double)K * eps);
  ema_finalize_cs_kernel<<<1, 1024, 0, stream>>>(cnt, ema_cluster_size, (int)K, fd, fo, fe, fke, scratch, info);
  VQ_LAUNCH_CHECK("ema_finalize_cs_kernel");
  const int grid = grid_for(img_kp(K), 8, sm_count() * 8);
  ema_finalize_w_kernel<<<grid, 256, 0, stream>>>(dw, ema_w, E, (int)K, (int)D, fd, fo, scratch, ee,
                                                  (unsigned char*)image, info);
  VQ_LAUNCH_CHECK("ema_finalize_w_kernel");
  return launch_image_f16(E, K, D, image, info, stream);
}

int vqb200_codebook_revive(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                           const float* usage, float threshold, uint64_t seed, float* E, float* ema_cluster_size,
                           float* ema_w, int64_t K, int32_t* revived, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(usage && E && revived, VQB200_EINVAL, "codebook_revive: null pointer");
  VQ_CHECK_ARG(B >= 0 && C > 0 && T > 0 && K > 0, VQB200_ESHAPE, "codebook_revive: bad shape");
  VQ_CUDA(cudaMemsetAsync(revived, 0, sizeof(int32_t), stream));
  if (B * T == 0) return VQB200_OK;
  VQ_CHECK_ARG(z != nullptr, VQB200_EINVAL, "codebook_revive: null input");
  const ZView zv = make_zview(z, B, C, T, sB, sC, sT);
  const int grid = (int)max(1LL, min((long long)(K + 7) / 8, (long long)sm_count() * 4));
  codebook_revive_kernel<<<grid, 256, 0, stream>>>(zv, usage, threshold, (unsigned long long)seed, E, ema_cluster_size,
                                                   ema_w, (int)K, (int)C, revived);
  VQ_LAUNCH_CHECK("codebook_revive_kernel");
  return VQB200_OK;
}

int vqb200_vq_histogram(const int32_t* idx, int64_t N, int64_t K, float* cnt, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(cnt && (idx || N == 0), VQB200_EINVAL, "vq_histogram: null pointer");
  VQ_CHECK_ARG(N >= 0 && K > 0, VQB200_ESHAPE, "vq_histogram: bad shape");
  VQ_CUDA(cudaMemsetAsync(cnt, 0, (size_t)K * sizeof(float), stream));
  if (N == 0) return VQB200_OK;
  const int use_hist = (K <= 12288) ? 1 : 0;
  const int grid = grid_for(N, 256 * 8, sm_count() * 4);
  histogram_kernel<<<grid, 256, use_hist ? (size_t)K * sizeof(int) : 0, stream>>>(idx, N, (int)K, cnt, use_hist);
  VQ_LAUNCH_CHECK("histogram_kernel");
  return VQB200_OK;
}

int vqb200_vq_backward_codebook(const float* stats, int64_t K, int64_t D, const float* g_loss, float coef,
                                float* gE, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(stats && gE, VQB200_EINVAL, "vq_backward_codebook: null pointer");
  const long long n = (long long)K * D;
  backward_codebook_kernel<<<grid_for(n, 256 * 4, sm_count() * 8), 256, 0, stream>>>(stats, n, g_loss, coef, gE);
  VQ_LAUNCH_CHECK("backward_codebook_kernel");
  return VQB200_OK;
}

}  // extern "C"
