// vqb200 -- EMA finalize fused with its all-reduce over NVLink peer memory (the one exchange step of
// the path: models/vqvae.py:44-50 under data parallelism).
//
// The multi-kernel data-parallel step is  K3a (local sums) -> all-reduce(stats) -> K3b (decay / normalise).  Here the
// all-reduce disappears into K3b: every rank accumulates its statistics straight into a slot of a *symmetric* buffer
// that all ranks of the node have mapped (CUDA IPC, NVLink / NVSwitch peer access); the finalize kernels
//   1. publish "my slot of epoch e is complete" with one release-store per peer into the peers' flag words and wait
//      for the same from everybody (one CTA, `world` threads: a one-shot NVLink barrier, no NCCL launch),
//   2. read every peer's [cnt | dw] slice directly over NVLink (coalesced L2-only loads) and add them in RANK ORDER,
//      so all ranks compute bit-identical sums -- and therefore bit-identical codebooks, with no broadcast --
//   3. apply decay, Laplace smoothing and the normalisation, refresh |E|^2 / the bf16 tile image, as ema.cu does.
// Slots are double-buffered by epoch parity: a rank overwrites slot e%2 only after it has passed barrier e+1, and a
// peer signals e+1 only after (stream order) its reads of epoch e have completed.
// A peer that never arrives (dead rank) trips a 120 s device-side timeout that traps instead of hanging the GPU.
#include <string.h>
#include "codebook.cuh"

namespace vqb200 {
namespace peer {

constexpr int MAX_PEERS = VQB200_MAX_PEERS;
constexpr unsigned long long TIMEOUT_NS = 120ull * 1000ull * 1000ull * 1000ull;   // rank skew at start-up can be many seconds

struct Table {
  const float* stats[MAX_PEERS];   // the SAME slot on every rank (index = rank), mapped into this process
  unsigned* flags[MAX_PEERS];      // flags[p][r]: last epoch rank r has published to rank p
  int world, rank;
  unsigned epoch;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// called by threads [0, world) of ONE CTA; all threads of the CTA must follow with __syncthreads()
__device__ __forceinline__ void publish_and_wait(const Table& t, int tid) {
  if (tid < t.world) {
    __threadfence_system();                                   // this rank's slot (earlier kernels) is visible system-wide
    unsigned* dst = t.flags[tid] + t.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(dst), "r"(t.epoch) : "memory");
    const unsigned* src = t.flags[t.rank] + tid;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if ((int)(v - t.epoch) >= 0) break;
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > TIMEOUT_NS) {
        printf("vqb200: rank %d waited 120 s for the EMA statistics of rank %d (epoch %u, saw %u)\n", t.rank, tid, t.epoch, v);
        __trap();
      }
    }
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(32) barrier_kernel(const Table t) {
  publish_and_wait(t, threadIdx.x);
}

// step 1 (one CTA): barrier, cnt = sum over ranks (rank order), cs / n / cluster as ema_finalize_cs_kernel
__global__ void __launch_bounds__(1024)
finalize_cs_kernel(const Table t, float* __restrict__ cnt_out, float* __restrict__ cs, int K, int D,
                   float decay, float one_minus_decay, float eps, float k_eps,
                   float* __restrict__ scratch, float* __restrict__ info) {
  __shared__ double red[32];
  __shared__ float s_n;
  publish_and_wait(t, threadIdx.x);
  __syncthreads();
  const size_t cnt_off = (size_t)K * D;
  double part = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float c = __ldcg(t.stats[0] + cnt_off + k);
    for (int p = 1; p < t.world; ++p) c = __fadd_rn(c, __ldcg(t.stats[p] + cnt_off + k));
    if (cnt_out) cnt_out[k] = c;
    const float v = fmaf(c, one_minus_decay, __fmul_rn(cs[k], decay));
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
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    scratch[k] = __fmul_rn(__fdiv_rn(__fadd_rn(cs[k], eps), __fadd_rn(n, k_eps)), n);
  if (threadIdx.x == 0) {
    scratch[K] = n;
    if (info) { info[0] = 0.f; info[1] = 0.f; info[2] = INFO2_RESET; info[3] = 0.f; }
  }
}

// step 2 (grid): dw = sum over ranks (rank order, read over NVLink); w / E / ee / image / info as ema_finalize_w_kernel
__global__ void __launch_bounds__(256)
finalize_w_kernel(const Table t, float* __restrict__ w, float* __restrict__ E, int K, int D,
                  float decay, float one_minus_decay, const float* __restrict__ cluster,
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
        float dw = __ldcg(t.stats[0] + o);
        for (int p = 1; p < t.world; ++p) dw = __fadd_rn(dw, __ldcg(t.stats[p] + o));
        const float wv = fmaf(dw, one_minus_decay, __fmul_rn(w[o], decay));
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

static int fill_table(Table& t, const float* const* peer_stats, uint32_t* const* peer_flags, int rank, int world,
                      uint32_t epoch, bool need_stats) {
  VQ_CHECK_ARG(peer_flags && (peer_stats || !need_stats), VQB200_EINVAL, "peer: null pointer table");
  VQ_CHECK_ARG(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, VQB200_ESHAPE,
               "peer: bad rank %d / world %d (at most %d ranks of one node)", rank, world, MAX_PEERS);
  for (int p = 0; p < MAX_PEERS; ++p) { t.stats[p] = nullptr; t.flags[p] = nullptr; }
  for (int p = 0; p < world; ++p) {
    VQ_CHECK_ARG(peer_flags[p] && (!need_stats || peer_stats[p]), VQB200_EINVAL, "peer: rank %d has no mapped buffer", p);
    t.flags[p] = reinterpret_cast<unsigned*>(peer_flags[p]);
    if (need_stats) {
      VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(peer_stats[p]) & 15) == 0, VQB200_EALIGN, "peer: slot of rank %d not 16-byte aligned", p);
      t.stats[p] = peer_stats[p];
    }
  }
  t.world = world; t.rank = rank; t.epoch = epoch;
  return VQB200_OK;
}

}  // namespace peer
}  // namespace vqb200

using namespace vqb200;

extern "C" {

int vqb200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle) {
  VQ_CHECK_ARG(dev_ptr && handle && bytes > 0, VQB200_EINVAL, "peer_alloc: null pointer / zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == VQB200_PEER_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  VQ_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "peer_alloc: cudaMemset / cudaIpcGetMemHandle"); }
  memcpy(handle, &h, sizeof(h));
  *dev_ptr = p;
  return VQB200_OK;
}

int vqb200_peer_open(const unsigned char* handle, void** dev_ptr) {
  VQ_CHECK_ARG(dev_ptr && handle, VQB200_EINVAL, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  VQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return VQB200_OK;
}

int vqb200_peer_close(void* dev_ptr) {
  if (dev_ptr) VQ_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return VQB200_OK;
}

int vqb200_peer_free(void* dev_ptr) {
  if (dev_ptr) VQ_CUDA(cudaFree(dev_ptr));
  return VQB200_OK;
}

int vqb200_peer_barrier(uint32_t* const* peer_flags, int32_t rank, int32_t world, uint32_t epoch,
                        vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  peer::Table t;
  const int rc = peer::fill_table(t, nullptr, peer_flags, rank, world, epoch, false);
  if (rc != VQB200_OK) return rc;
  {
    // CUDA loads kernels lazily, and loading may need a context-wide synchronisation: a first launch of the finalize
    // kernels queued behind a barrier that is still spinning would stall the host.  Load them here, once.
    static PerDevice loaded_;
    std::atomic<size_t>& loaded = loaded_.here();
    if (!loaded.load()) {
      cudaFuncAttributes fa;
      VQ_CUDA(cudaFuncGetAttributes(&fa, peer::finalize_cs_kernel));
      VQ_CUDA(cudaFuncGetAttributes(&fa, peer::finalize_w_kernel));
      VQ_CUDA(cudaFuncGetAttributes(&fa, peer::barrier_kernel));
      { const int rc2 = preload_image_f16(); if (rc2 != VQB200_OK) return rc2; }
      loaded.store(1);
    }
  }
  peer::barrier_kernel<<<1, 32, 0, stream>>>(t);
  VQ_LAUNCH_CHECK("peer::barrier_kernel");
  return VQB200_OK;
}

int vqb200_ema_finalize_peer(const float* const* peer_stats, uint32_t* const* peer_flags, int32_t rank, int32_t world,
                             uint32_t epoch, float* cnt_out, float* ema_cluster_size, float* ema_w, float* E,
                             int64_t K, int64_t D, double decay, double eps, float* ee, void* image, float* info,
                             float* scratch, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(ema_cluster_size && ema_w && E && scratch, VQB200_EINVAL, "ema_finalize_peer: null pointer");
  VQ_CHECK_ARG(K > 0 && D > 0, VQB200_ESHAPE, "ema_finalize_peer: bad K/D");
  VQ_CHECK_ARG(!image || (reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "ema_finalize_peer: image must be 1024-byte aligned");
  peer::Table t;
  const int rc = peer::fill_table(t, peer_stats, peer_flags, rank, world, epoch, true);
  if (rc != VQB200_OK) return rc;
  const float fd = (float)decay, fo = (float)(1.0 - decay), fe = (float)eps, fke = (float)((double)K * eps);
  peer::finalize_cs_kernel<<<1, 1024, 0, stream>>>(t, cnt_out, ema_cluster_size, (int)K, (int)D, fd, fo, fe, fke, scratch, info);
  VQ_LAUNCH_CHECK("peer::finalize_cs_kernel");
  const int grid = grid_for(img_kp(K), 8, sm_count() * 8);
  peer::finalize_w_kernel<<<grid, 256, 0, stream>>>(t, ema_w, E, (int)K, (int)D, fd, fo, scratch, ee,
                                                    (unsigned char*)image, info);
  VQ_LAUNCH_CHECK("peer::finalize_w_kernel");
  return launch_image_f16(E, K, D, image, info, stream);
}

}  // extern "C"
