// vqb200 K1 (tensor-core variant for D = 128 / 256 / 512): fused distance + argmin on tcgen05 / TMEM / bulk-TMA.
//
// Same exactness scheme as assign_tc.cu (split-bf16 filter + proven margin + exact fallback list), for the
// wider embedding dims of BASELINE cfg5.  A row no longer fits a register file for an in-place conversion, so the
// input is first split into a bf16 "row image" by z_image_kernel (per 128 rows x 64 dims one 16 KiB x_hi tile +
// one 16 KiB x_lo tile, K-major SWIZZLE_128B -- the mirror of the codebook image), after which BOTH operands of
// the GEMM arrive by 1-D bulk-TMA copies and the kernel is a pure tcgen05 pipeline:
//   warp 0     producer: whole A operand of a 128-row tile (KB x 32 KiB, one copy) + 32 KiB codebook tiles
//              (code tile j, dim block kb) through a 3-stage ring + 512 B of -|E|^2/2 per code tile
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer: per code tile 12*KB MMAs (M=128, N=128, K=16)
//   warps 2-5  epilogue: tcgen05.ld, running top-2 with packed index (tc_common.cuh), margin test
// At D >= 128 the MMAs (3 split products) outweigh the TMEM drain.  D = 512: the A operand of a row tile (256 KiB)
// no longer fits, so its 64-dim blocks are streamed through the ring next to the codebook tiles (STREAM_A).
#include "common.cuh"
#include "codebook.cuh"
#include "tc_common.cuh"

namespace vqb200 {

int launch_assign_simt(const ZView& z, const float* E, const float* ee, int K, int D,
                       int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                       long long max_rows, cudaStream_t stream, unsigned long long* keys = nullptr);

namespace tcg {
using namespace tcc;

constexpr int NHS = 5;                               // -|E|^2/2 ring (reuse distance >= ceil(NST/KB) + 2 code tiles)
constexpr int NTHREADS = 320;                        // producer, MMA, 8 epilogue warps
constexpr int A_BLOCK = 2 * TILE_M * 128;            // 32768: [x_hi | x_lo] for one 64-dim block
constexpr int SMEM_NH = NHS * BN * 4;
constexpr int SMEM_BAR = 256;
constexpr int SMEM_XCH = TILE_M * 16;                // per-row (g1, g2, gi) hand-over between the two column halves

// ---- z -> split-bf16 row image -------------------------------------------------------------------------
// image tile (row tile rt, dim block kb) at byte offset (rt*KB + kb) * 32768: 16 KiB hi then 16 KiB lo,
// row r = 128 bytes, 16-byte chunk j at position j ^ (r & 7).  Rows >= N are zero.
__global__ void __launch_bounds__(256)
z_image_kernel(ZView z, int D, int KB, int sub_rows, unsigned char* __restrict__ image, float* __restrict__ xnorm2,
               long long n_row_tiles) {
  extern __shared__ __align__(16) float tile[];       // [sub_rows][D + 4]
  const int LD = D + 4;
  const int tid = threadIdx.x;
  const int subs = TILE_M / sub_rows;                 // 128-row image tiles are filled in `subs` passes
  for (long long it = blockIdx.x; it < n_row_tiles * subs; it += gridDim.x) {
    const long long rt = it / subs;
    const int r_base = (int)(it - rt * subs) * sub_rows;       // first row of this pass inside the image tile
    const long long n0 = rt * TILE_M + r_base;
    const int rows = (int)max(0LL, min((long long)sub_rows, z.N - n0));
    __syncthreads();
    if (rows > 0) load_rows(z, n0, rows, D, tid, 256, [&](int r, int k, float v) { tile[r * LD + k] = v; });
    __syncthreads();
    if (tid < sub_rows) {                                  // exact fp32 |x|^2 per row (feeds the error bound)
      float acc = 0.f;
      if (tid < rows) for (int k = 0; k < D; ++k) { const float v = tile[tid * LD + ((k + tid) % D)]; acc = fmaf(v, v, acc); }
      xnorm2[n0 + tid] = acc;
    }
    // one thread per (row, 8-dim chunk)
    const int chunks = D >> 3;
    for (int i = tid; i < sub_rows * chunks; i += 256) {
      const int rl = i / chunks, c = i - rl * chunks;     // chunk c covers dims 8c..8c+7
      const int r = r_base + rl;                          // row inside the 128-row image tile
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = (rl < rows) ? tile[rl * LD + 8 * c + 2 * e] : 0.f;
        const float b = (rl < rows) ? tile[rl * LD + 8 * c + 2 * e + 1] : 0.f;
        hw[e] = pack_bf16x2(a, b);
        lw[e] = pack_bf16x2(a - __uint_as_float(hw[e] << 16), b - __uint_as_float(hw[e] & 0xFFFF0000u));
      }
      const int kb = c >> 3, j = c & 7;
      unsigned char* base = image + ((size_t)rt * KB + kb) * A_BLOCK + (size_t)r * 128 + ((j ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(base) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(base + TILE_M * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
}

struct Params {
  const unsigned char* zimg;      // split-bf16 row image of z
  const unsigned char* image;     // codebook tile image
  const float* neg_half_ee;
  const float* info;
  const float* xnorm2;            // |x_n|^2 per row (padded to whole tiles)
  int K, NT, D;
  long long N, ntiles;            // rows, 128-row tiles
  int32_t* idx;
  int32_t* list;
  int32_t* list_count;
  int* err;
};

template <int KB, int NST, bool EPI8, bool STREAM_A>
__global__ void __launch_bounds__(NTHREADS, 1)
vq_assign_tc_gen_kernel(const Params p) {
  constexpr int STAGE_BYTES = IMG_TILE_BYTES + (STREAM_A ? A_BLOCK : 0);   // [A block (streamed) |] codebook tile
  constexpr int SMEM_B = NST * STAGE_BYTES;
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int SMEM_A = STREAM_A ? 0 : KB * A_BLOCK;
  unsigned char* sA = smem;                       // [KB][hi 16K | lo 16K]
  unsigned char* sB = smem + SMEM_A;              // [NST][32768]
  float* sN = reinterpret_cast<float*>(sB + SMEM_B);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SMEM_B + SMEM_NH);
  uint64_t* full = bars;                 // [NST]
  uint64_t* empty = full + NST;          // [NST]
  uint64_t* tfull = empty + NST;         // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* afull = tempty + 2;          // [1]
  uint64_t* aempty = afull + 1;          // [1]
  uint64_t* nhfull = aempty + 1;         // [NHS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nhfull + NHS);
  float4* sX = reinterpret_cast<float4*>(sB + SMEM_B + SMEM_NH + SMEM_BAR);     // [128] top-2 exchange

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int s = 0; s < NHS; ++s) mbar_init(smem_u32(nhfull + s), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), EPI8 ? 8 : 4); }
    mbar_init(smem_u32(afull), 1); mbar_init(smem_u32(aempty), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int NT = p.NT;
  if ((smem_u32(smem) & 1023u) != 0u) { if (tid == 0 && p.err) atomicExch(p.err, 99); __trap(); }

  if (warp == 0) {
    if (lane == 0) {
      unsigned it = 0, nt = 0, tile_i = 0;
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
        if (!STREAM_A) {
          mbar_wait(smem_u32(aempty), (tile_i & 1) ^ 1, p.err, 7);           // MMAs of the previous tile left A
          mbar_expect_tx(smem_u32(afull), SMEM_A);
          bulk_g2s(smem_u32(sA), p.zimg + (size_t)tile * (KB * A_BLOCK), SMEM_A, smem_u32(afull));
        }
        for (int j = 0; j < NT; ++j, ++nt) {
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const unsigned s = it % NST, ph = (it / NST) & 1;
            mbar_wait(smem_u32(empty + s), ph ^ 1, p.err, 1);
            mbar_expect_tx(smem_u32(full + s), STAGE_BYTES);
            if (STREAM_A)
              bulk_g2s(smem_u32(sB + (size_t)s * STAGE_BYTES), p.zimg + ((size_t)tile * KB + kb) * A_BLOCK, A_BLOCK,
                       smem_u32(full + s));
            bulk_g2s(smem_u32(sB + (size_t)s * STAGE_BYTES + (STREAM_A ? A_BLOCK : 0)),
                     p.image + ((size_t)j * KB + kb) * IMG_TILE_BYTES, IMG_TILE_BYTES, smem_u32(full + s));
            if (kb == 0) {          // slot nt % NHS is free: see the reuse-distance argument in assign_tc.cu
              mbar_expect_tx(smem_u32(nhfull + nt % NHS), BN * 4);
              bulk_g2s(smem_u32(sN + (size_t)(nt % NHS) * BN), p.neg_half_ee + (size_t)j * BN, BN * 4,
                       smem_u32(nhfull + nt % NHS));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      unsigned it = 0, nt = 0, tile_i = 0;
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
        if (!STREAM_A) mbar_wait(smem_u32(afull), tile_i & 1, p.err, 3);
        for (int j = 0; j < NT; ++j, ++nt) {
          const unsigned as = nt & 1;
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const unsigned s = it % NST;
            mbar_wait(smem_u32(full + s), (it / NST) & 1, p.err, 2);
            if (kb == 0) mbar_wait(smem_u32(tempty + as), ((nt >> 1) & 1) ^ 1, p.err, 4);
            tc_fence_after();
            const uint32_t a_hi = STREAM_A ? smem_u32(sB + (size_t)s * STAGE_BYTES) : smem_u32(sA + (size_t)kb * A_BLOCK);
            const uint32_t a_lo = a_hi + TILE_M * 128;
            const uint32_t b_hi = smem_u32(sB + (size_t)s * STAGE_BYTES + (STREAM_A ? A_BLOCK : 0)), b_lo = b_hi + IMG_HALF_BYTES;
#pragma unroll
            for (int sp = 0; sp < 3; ++sp) {            // x_hi.E_hi + x_lo.E_hi + x_hi.E_lo
              const uint32_t a = (sp == 1) ? a_lo : a_hi;
              const uint32_t b = (sp == 2) ? b_lo : b_hi;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), IDESC, (kb | sp | k) ? 1u : 0u);
            }
            umma_commit(smem_u32(empty + s));
          }
          umma_commit(smem_u32(tfull + as));
          if (!STREAM_A && j == NT - 1) umma_commit(smem_u32(aempty));
        }
      }
    }
  } else if (EPI8 || warp < 6) {
    // EPI8: two warps per TMEM lane quarter -- warp pair (w, w+4) shares rows and splits the 128 columns of a
    // code tile; otherwise one warp per quarter scans all 128 columns
    const int q = warp & 3;
    const int half = EPI8 ? ((warp - 2) >> 2) : 0;
    const int row = q * 32 + lane;
    const float emax = p.info[0];
    const bool cb_bad = p.info[1] != 0.f;
    const uint32_t mask = 0xFFFFFF80u;
    unsigned nt = 0, tile_i = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
      const long long n0 = tile * TILE_M;
      const int rows = (int)min((long long)TILE_M, p.N - n0);
      const float xx = __ldg(p.xnorm2 + n0 + row);         // exact fp32 |x|^2 from z_image_kernel
      float g1 = -INFINITY, g2 = -INFINITY; int gi = 0;
      for (int j = 0; j < NT; ++j, ++nt) {
        const unsigned as = nt & 1;
        mbar_wait(smem_u32(tfull + as), (nt >> 1) & 1, p.err, 6);
        mbar_wait(smem_u32(nhfull + nt % NHS), (nt / NHS) & 1, p.err, 9);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * 64;
        const float4* nh = reinterpret_cast<const float4*>(sN + (size_t)(nt % NHS) * BN + half * 64);
        float t1 = -INFINITY, t2 = -INFINITY;
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_ld_wait();
        tmem_ld32(taddr + 32, vb);
        top2_chunk(va, nh, half * 64, mask, t1, t2);
        tmem_ld_wait();
        if (!EPI8) tmem_ld32(taddr + 64, va);
        top2_chunk(vb, nh + 8, half * 64 + 32, mask, t1, t2);
        if (!EPI8) {
          tmem_ld_wait();
          tmem_ld32(taddr + 96, vb);
          top2_chunk(va, nh + 16, 64, mask, t1, t2);
          tmem_ld_wait();
          top2_chunk(vb, nh + 24, 96, mask, t1, t2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_after(smem_u32(tempty + as), smem_u32(tmem_slot + 1), t1, t2);
        const int ti = j * BN + (int)(__float_as_uint(t1) & 127u);
        if (t1 > g1) { g2 = fmaxf(g1, t2); g1 = t1; gi = ti; }
        else { g2 = fmaxf(g2, t1); }
      }
      // merge the two column halves of every row (half 1 hands its top-2 to half 0)
      if (EPI8) {
        if (half == 1) sX[row] = make_float4(g1, g2, __int_as_float(gi), 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
          const float4 o = sX[row];
          if (o.x > g1) { g2 = fmaxf(g1, o.y); g1 = o.x; gi = __float_as_int(o.z); }
          else { g2 = fmaxf(g2, o.x); }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");        // sX free for the next tile
      }
      if (half == 0 && row < rows) {
        const long long n = n0 + row;
        const float mag = sqrtf(xx) * 1.0001f * emax;
        const float thr = 2.5f * score_error_bound(mag, emax, p.D);
        const bool proven = !cb_bad && (g1 - g2 > thr) && (fabsf(g1) < 1e37f) && (mag < 1e37f) && (gi < p.K);
        p.idx[n] = proven ? gi : 0;
        if (!proven) {
          const int pos = atomicAdd(p.list_count, 1);
          p.list[pos] = (int32_t)n;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int KB, int NST, bool EPI8, bool STREAM_A>
static int launch_kb(const Params& p, cudaStream_t stream) {
  constexpr int smem = (STREAM_A ? 0 : KB * A_BLOCK) + NST * (IMG_TILE_BYTES + (STREAM_A ? A_BLOCK : 0)) + SMEM_NH + SMEM_BAR +
                       (EPI8 ? SMEM_XCH : 0);
  static_assert(smem <= 232448, "shared memory budget");
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (!configured.load()) {
    VQ_CUDA(cudaFuncSetAttribute(vq_assign_tc_gen_kernel<KB, NST, EPI8, STREAM_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.store(1);
  }
  const int grid = (int)max(1LL, min(p.ntiles, (long long)sm_count()));
  vq_assign_tc_gen_kernel<KB, NST, EPI8, STREAM_A><<<grid, NTHREADS, smem, stream>>>(p);
  VQ_LAUNCH_CHECK("vq_assign_tc_gen_kernel");
  return VQB200_OK;
}

}  // namespace tcg

bool assign_tc_gen_eligible(const ZView& z, int K, int D) {
  return (D == 128 || D == 256 || D == 512) && z.C == D && K >= 1 && z.N >= 1;
}

// workspace: 256 B header (list_count, err) + row list (N int32, padded) + split-bf16 row image
static size_t gen_list_bytes(long long N) { return ((size_t)N * sizeof(int32_t) + 1023) & ~(size_t)1023; }
size_t assign_tc_gen_workspace_bytes(long long N, int D) {
  const long long tiles = (N + tcc::TILE_M - 1) / tcc::TILE_M;
  return 1024 + gen_list_bytes(N) + (size_t)tiles * tcc::TILE_M * sizeof(float) + (size_t)tiles * (D / 64) * tcg::A_BLOCK + 1024;
}

int launch_assign_tc_gen(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                         int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
  using namespace tcg;
  VQ_CHECK_ARG(workspace_bytes >= assign_tc_gen_workspace_bytes(z.N, D), VQB200_EWORKSPACE, "vq_assign(TC): workspace too small");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "vq_assign(TC): image must be 1024-byte aligned");
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  int32_t* hdr = reinterpret_cast<int32_t*>(base);
  VQ_CUDA(cudaMemsetAsync(hdr, 0, 256, stream));
  const int KB = D / 64;
  Params p;
  const long long tiles = (z.N + TILE_M - 1) / TILE_M;
  float* xn = reinterpret_cast<float*>(base + 1024 + gen_list_bytes(z.N));
  p.xnorm2 = xn;
  p.zimg = base + 1024 + gen_list_bytes(z.N) + (size_t)tiles * TILE_M * sizeof(float);     // stays 1024-byte aligned
  p.image = reinterpret_cast<const unsigned char*>(image);
  p.neg_half_ee = reinterpret_cast<const float*>(p.image + img_tiles_bytes(K, D));
  p.info = info;
  p.K = K; p.D = D;
  p.NT = (int)(img_kp(K) / IMG_TILE_CODES);
  p.N = z.N;
  p.ntiles = (z.N + TILE_M - 1) / TILE_M;
  p.idx = idx;
  p.list = reinterpret_cast<int32_t*>(base + 1024);
  p.list_count = hdr;
  p.err = hdr + 1;
  // 1. split the input into the bf16 row image
  {
    const int sub_rows = (D <= 256) ? TILE_M : TILE_M / 2;          // D = 512: two 64-row passes per image tile
    const size_t smem = (size_t)sub_rows * (D + 4) * sizeof(float);
    static PerDevice configured_;
    std::atomic<size_t>& configured = configured_.here();
    if (smem > 48 * 1024 && smem > configured.load()) {
      VQ_CUDA(cudaFuncSetAttribute(z_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured.store(smem);
    }
    const long long work = p.ntiles * (TILE_M / sub_rows);
    const int grid = (int)max(1LL, min(work, (long long)sm_count() * (smem > 100 * 1024 ? 1 : 2)));
    z_image_kernel<<<grid, 256, smem, stream>>>(z, D, KB, sub_rows, const_cast<unsigned char*>(p.zimg), xn, p.ntiles);
    VQ_LAUNCH_CHECK("z_image_kernel");
  }
  // 2. tcgen05 filter
  // ring depth / epilogue width = what 227 KiB of shared memory allow
  int rc = (KB == 2) ? launch_kb<2, 4, true, false>(p, stream)
         : (KB == 4) ? launch_kb<4, 3, false, false>(p, stream)
                     : launch_kb<8, 3, false, true>(p, stream);
  if (rc != VQB200_OK) return rc;
  // 3. exact re-do of the rows the filter could not prove
  return launch_assign_simt(z, E, ee, K, D, idx, best, p.list, p.list_count, z.N, stream);
}

}  // namespace vqb200
