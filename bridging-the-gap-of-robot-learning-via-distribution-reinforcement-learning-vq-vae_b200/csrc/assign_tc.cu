// vqb200 K1 (tensor-core variant): fused distance + argmin on tcgen05 / TMEM / bulk-TMA (sm_100a).
//
// Replaces models/vqvae.py:30-38 of the reference for D == 64 (the hidden size of every BASELINE
// config) at any K.  Never materialises the N x K matrix.
//
// Exactness scheme ("split-bf16 filter + proven margin"):
//   x and E are each split into two bf16 terms (x = x_hi + x_lo, E = E_hi + E_lo, remainder 2^-18);
//   the tensor cores accumulate x_hi.E_hi + x_lo.E_hi + x_hi.E_lo in fp32 (bf16 x bf16 products are
//   exact in fp32), i.e. x.E to ~2^-16 relative.  The epilogue forms the score
//   s_k = x.E_k - |E_k|^2/2  (arg max s == arg min of the reference distance), keeps the best two
//   scores per row, and accepts the best code only when it leads the runner-up by more than a
//   rigorous bound on the total error of both scores (split remainder + fp32 accumulation +
//   index packing).  Rows that cannot be proven (near-ties, non-finite data) are appended to a work
//   list and re-done by the exact fp32 CUDA-core kernel (assign_simt.cu) -- typically < 1 % of rows.
//
// Structure (one persistent CTA per SM, 448 threads, every hand-off through mbarriers):
//   warp 0       bulk-TMA producer: (a) prefetches the NEXT tile's raw fp32 rows (one contiguous slab
//                per 128-row group) into a ping-pong buffer, (b) streams 32 KiB codebook tiles
//                (E_hi | E_lo, pre-swizzled image written by ema_finalize / codebook_prepare) plus
//                their 512 B of -|E|^2/2 through a 3-stage ring
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=128, K=16, kind::f16)
//   warps 2-9    epilogue: two groups of 4 warps, one 128-row tile each; tcgen05.ld the fp32 scores
//                (one row per thread, software-pipelined) and run the running top-2 with the code
//                index packed into the low mantissa bits (3.5 ALU ops per score)
//   warps 10-13  converter: turns the prefetched raw rows IN PLACE into the swizzled K-major split-bf16
//                A operands of the next tile while the epilogue warps are still busy with this one
//   TMEM: 4 accumulators of 128 columns (2 row tiles x 2 stages) = all 512 columns, so the MMAs of
//   code tile j+1 overlap the epilogue of code tile j.
#include <stdlib.h>
#include "common.cuh"
#include "codebook.cuh"
#include "tc_common.cuh"

namespace vqb200 {

int launch_assign_simt(const ZView& z, const float* E, const float* ee, int K, int D,
                       int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                       long long max_rows, cudaStream_t stream, unsigned long long* keys = nullptr);

namespace tc {
using namespace tcc;

constexpr int RT = 2;                       // row tiles per CTA
constexpr int D = 64;
constexpr int NST = 3;                      // codebook ring stages
constexpr int NHS = 5;                      // -|E|^2/2 ring slots (reuse distance NST+2, see producer)
constexpr int A_HALF = TILE_M * 128;        // 16384 B: one of {hi, lo} for one row tile
constexpr int BUF_BYTES = 2 * A_HALF;       // 32768: raw fp32 rows, converted in place to [hi | lo]
constexpr int SMEM_BUF = RT * 2 * BUF_BYTES;    // 131072: ping-pong per row tile
constexpr int SMEM_B = NST * IMG_TILE_BYTES;    // 98304
constexpr int SMEM_NH = NHS * BN * 4;       // 2560
constexpr int SMEM_BAR = 256;
constexpr int SMEM_TOTAL = SMEM_BUF + SMEM_B + SMEM_NH + SMEM_BAR;   // 232192 <= 232448
enum StageMode : int { STG_DIRECT = 0, STG_ROWS = 1, STG_BCT = 2 };
constexpr int NTHREADS = 448;

__device__ __forceinline__ void converter_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Params {
  ZView z;
  const unsigned char* image;     // codebook tile image (hi|lo tiles) ...
  const float* neg_half_ee;       // ... followed by -|E_k|^2/2 (padded with -inf)
  const float* info;              // {max |E_k|, nonfinite flag}
  int K, NT;                      // codes, number of 128-code tiles
  int stage_mode;                 // how raw z reaches shared memory (StageMode)
  int R;                          // rows per 4-warp group: 128, or the largest multiple of T <= 128 (whole samples)
  long long ntiles;               // CTA tiles of 2*R rows
  int32_t* idx;
  int32_t* list;                  // rows that need the exact kernel
  int32_t* list_count;
  int* err;
  // fused residual update (RVQ stages >= 1): the staged rows are r_prev; the converter forms
  // r = r_prev - st(r_prev, prev_E[prev_idx]) (models/vqvae.py:94-98), stores it to r_out and quantizes THAT
  const int32_t* prev_idx;
  const float* prev_E;
  int prev_K;
  float* r_out;                   // same layout as z (contiguous); null = plain assignment
  int dbg;                        // development knobs (VQB200_TC_DEBUG): 1 = skip epilogue math, 2 = one k-block
};

// Byte range of the raw fp32 input that covers rows [n0, n0+rows) (staged modes only).
struct StagePlan { const float* src; uint32_t bytes; long long b_lo; };
__device__ __forceinline__ StagePlan stage_plan(const Params& p, long long n0, int rows) {
  StagePlan sp;
  sp.bytes = (uint32_t)rows * (D * 4);
  if (p.stage_mode == STG_ROWS) {
    sp.src = p.z.p + n0 * D; sp.b_lo = 0;
  } else {                         // STG_BCT: groups start on sample boundaries and hold whole samples
    sp.b_lo = n0 / p.z.T;
    sp.src = p.z.p + sp.b_lo * (D * p.z.T);
  }
  return sp;
}

__global__ void __launch_bounds__(NTHREADS, 1)
vq_assign_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBuf = smem;                     // [RT][2][32768]  raw rows -> A operands (in place)
  unsigned char* sB = smem + SMEM_BUF;            // [NST][32768]    codebook tiles
  float* sN = reinterpret_cast<float*>(sB + SMEM_B);          // [NHS][128]  -|E_k|^2/2 of in-flight code tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SMEM_B + SMEM_NH);
  uint64_t* full = bars;                 // [NST]        codebook tile landed
  uint64_t* empty = full + NST;          // [NST]        codebook tile consumed by the MMAs
  uint64_t* tfull = empty + NST;         // [2][RT]      accumulator ready
  uint64_t* tempty = tfull + 2 * RT;     // [2][RT]      accumulator drained
  uint64_t* rawfull = tempty + 2 * RT;   // [RT][2]      raw rows landed in the ping-pong buffer
  uint64_t* afull = rawfull + 2 * RT;    // [RT][2]      A operands written
  uint64_t* aempty = afull + 2 * RT;     // [RT][2]      A operands no longer read by the tensor core
  uint64_t* nhfull = aempty + 2 * RT;    // [NHS]        -|E|^2/2 slot landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nhfull + NHS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int s = 0; s < NHS; ++s) mbar_init(smem_u32(nhfull + s), 1);
    for (int i = 0; i < 2 * RT; ++i) {
      mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 4);
      mbar_init(smem_u32(rawfull + i), 1); mbar_init(smem_u32(afull + i), 4); mbar_init(smem_u32(aempty + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int NT = p.NT;
  const int R = p.R;
  const long long tile_rows = (long long)RT * R;
  if ((smem_u32(smem) & 1023u) != 0u) { if (tid == 0 && p.err) atomicExch(p.err, 99); __trap(); }
  const bool staged = p.stage_mode != STG_DIRECT;

  if (warp == 0) {
    // ================= bulk-TMA producer: raw z slabs (one tile ahead) + codebook tiles =================
    if (lane == 0) {
      auto issue_raw = [&](long long tile, unsigned tile_i) {
        if (!staged || tile >= p.ntiles) return;
        const unsigned pp = tile_i & 1, u = tile_i >> 1;
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
          const long long n0 = tile * tile_rows + (long long)rt * R;
          const int rows = (int)max(0LL, min((long long)R, p.z.N - n0));
          if (rows > 0) {
            const StagePlan sp = stage_plan(p, n0, rows);
            mbar_wait(smem_u32(aempty + rt * 2 + pp), (u & 1) ^ 1, p.err, 7);     // MMAs of tile_i-2 left the buffer
            mbar_expect_tx(smem_u32(rawfull + rt * 2 + pp), sp.bytes);
            bulk_g2s(smem_u32(sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES), sp.src, sp.bytes, smem_u32(rawfull + rt * 2 + pp));
          }
        }
      };
      unsigned it = 0, tile_i = 0;
      issue_raw(blockIdx.x, 0);
      const int j_raw = min(NST, NT - 1);         // by then the MMAs of the previous tile are done
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
        for (int j = 0; j < NT; ++j, ++it) {
          const unsigned s = it % NST, ph = (it / NST) & 1;
          mbar_wait(smem_u32(empty + s), ph ^ 1, p.err, 1);
          if (j == j_raw) issue_raw(tile + gridDim.x, tile_i + 1);
          mbar_expect_tx(smem_u32(full + s), IMG_TILE_BYTES);
          bulk_g2s(smem_u32(sB + (size_t)s * IMG_TILE_BYTES), p.image + (size_t)j * IMG_TILE_BYTES, IMG_TILE_BYTES,
                   smem_u32(full + s));
          // the epilogue of code tile `it` reads slot it % NHS until MMA(it+2) may start; this copy is issued
          // after MMA(it+NHS-NST) = MMA(it+2) has completed, so the slot is free AND its barrier cannot run a
          // phase ahead of the epilogue's parity wait (NHS = NST + 2)
          mbar_expect_tx(smem_u32(nhfull + it % NHS), BN * 4);
          bulk_g2s(smem_u32(sN + (size_t)(it % NHS) * BN), p.neg_half_ee + (size_t)j * BN, BN * 4,
                   smem_u32(nhfull + it % NHS));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      unsigned it = 0, tile_i = 0;
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
        const unsigned pp = tile_i & 1, u = tile_i >> 1;
        for (int j = 0; j < NT; ++j, ++it) {
          const unsigned s = it % NST, as = it & 1;
          mbar_wait(smem_u32(full + s), (it / NST) & 1, p.err, 2);
          tc_fence_after();
          const uint32_t b_hi = smem_u32(sB + (size_t)s * IMG_TILE_BYTES), b_lo = b_hi + IMG_HALF_BYTES;
#pragma unroll
          for (int rt = 0; rt < RT; ++rt) {
            if (j == 0) mbar_wait(smem_u32(afull + rt * 2 + pp), u & 1, p.err, 3);
            mbar_wait(smem_u32(tempty + as * RT + rt), ((it >> 1) & 1) ^ 1, p.err, 4);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES), a_lo = a_hi + A_HALF;
            const uint32_t d_tmem = tmem_base + (uint32_t)((as * RT + rt) * BN);
            const int nkb = (p.dbg & 2) ? 1 : 3;
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {            // x_hi.E_hi + x_lo.E_hi + x_hi.E_lo
              if (kb >= nkb) break;
              const uint32_t a = (kb == 1) ? a_lo : a_hi;
              const uint32_t b = (kb == 2) ? b_lo : b_hi;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), IDESC, (kb | k) ? 1u : 0u);
            }
            umma_commit(smem_u32(tfull + as * RT + rt));
            if (j == NT - 1) umma_commit(smem_u32(aempty + rt * 2 + pp));
          }
          umma_commit(smem_u32(empty + s));
        }
      }
    }
  } else if (warp >= 10) {
    // ================= converter: raw fp32 rows -> swizzled split-bf16 A operands, in place =================
    const int row = (warp - 10) * 32 + lane;      // 0..127: this thread's row inside the group
    unsigned tile_i = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
      const unsigned pp = tile_i & 1, u = tile_i >> 1;
#pragma unroll 1
      for (int rt = 0; rt < RT; ++rt) {
        const long long n0 = tile * tile_rows + (long long)rt * R;
        const int rows = (int)max(0LL, min((long long)R, p.z.N - n0));
        unsigned char* buf = sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES;
        const float* raw = reinterpret_cast<const float*>(buf);
        float v[D];
        const bool fuse = p.r_out != nullptr;       // staged modes only (checked by the launcher)
        int kp = 0;                                 // previous stage's code of this row, fetched before the wait
        if (fuse && row < rows) kp = min(max(__ldg(p.prev_idx + n0 + row), 0), p.prev_K - 1);
        if (staged && rows > 0) mbar_wait(smem_u32(rawfull + rt * 2 + pp), u & 1, p.err, 8);
        else mbar_wait(smem_u32(aempty + rt * 2 + pp), (u & 1) ^ 1, p.err, 5);     // nobody fills it for us: wait until free
        if (row < rows) {
          const long long n = n0 + row;
          if (p.stage_mode == STG_ROWS) {
            const float4* src = reinterpret_cast<const float4*>(raw + row * D);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
              const float4 f = src[c];
              v[4 * c] = f.x; v[4 * c + 1] = f.y; v[4 * c + 2] = f.z; v[4 * c + 3] = f.w;
            }
          } else if (p.stage_mode == STG_BCT) {
            const int T = (int)p.z.T;
            const long long b = n / T; const int t = (int)(n - b * T);
            const float* src = raw + (b - n0 / T) * (D * T) + t;
#pragma unroll
            for (int k = 0; k < D; ++k) v[k] = src[k * T];
          } else {
            const float* src = p.z.p + p.z.row_base(n);
#pragma unroll
            for (int k = 0; k < D; ++k) v[k] = __ldg(src + (long long)k * p.z.sC);
          }
          if (fuse) {                              // same three roundings as the stand-alone residual kernel
            const float4* q4 = reinterpret_cast<const float4*>(p.prev_E + (size_t)kp * D);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
              const float4 q = __ldg(q4 + c);
              const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = v[4 * c + e];
                v[4 * c + e] = __fsub_rn(x, __fadd_rn(x, __fsub_rn(qv[e], x)));
              }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < D; ++k) v[k] = 0.f;
        }
        converter_sync();                          // every raw read of this buffer is done
        if (fuse) {
          // write the new residual back in the raw layout and stream the slab to r_out before converting in place
          if (row < rows) {
            float* rawW = reinterpret_cast<float*>(buf);
            if (p.stage_mode == STG_ROWS) {
              float4* dst = reinterpret_cast<float4*>(rawW + row * D);
#pragma unroll
              for (int c = 0; c < D / 4; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            } else {
              const int T = (int)p.z.T;
              const long long n = n0 + row;
              const long long b = n / T; const int t = (int)(n - b * T);
              float* dst = rawW + (b - n0 / T) * (D * T) + t;
#pragma unroll
              for (int k = 0; k < D; ++k) dst[k * T] = v[k];
            }
          }
          fence_proxy_async();
          converter_sync();
          if (warp == 10 && lane == 0 && rows > 0) {
            const StagePlan sp = stage_plan(p, n0, rows);
            bulk_s2g(p.r_out + (sp.src - p.z.p), smem_u32(buf), sp.bytes);
            bulk_commit();
            bulk_wait_read<0>();                   // the store has read the buffer: safe to overwrite it
          }
          converter_sync();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = v[j * 8 + 2 * e], b = v[j * 8 + 2 * e + 1];
            hw[e] = pack_bf16x2(a, b);
            lw[e] = pack_bf16x2(a - __uint_as_float(hw[e] << 16), b - __uint_as_float(hw[e] & 0xFFFF0000u));
          }
          const int off = row * 128 + ((j ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(buf + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(buf + A_HALF + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
        converter_sync();
        if (lane == 0) mbar_arrive(smem_u32(afull + rt * 2 + pp));
      }
    }
  } else {
    // ================= epilogue groups (4 warps = 128 rows each) =================
    const int rt = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // accumulator lane == row inside the row tile
    const float emax = p.info[0];
    const bool cb_bad = p.info[1] != 0.f;
    const uint32_t mask = 0xFFFFFF80u;
    unsigned it = 0, tile_i = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
      const unsigned pp = tile_i & 1, u = tile_i >> 1;
      const long long n0 = tile * tile_rows + (long long)rt * R;
      const int rows = (int)max(0LL, min((long long)R, p.z.N - n0));
      // |x|^2 of this thread's row from the split operands (only feeds the error bound)
      mbar_wait(smem_u32(afull + rt * 2 + pp), u & 1, p.err, 10);
      const unsigned char* a_hi = sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES;
      float xx = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int off = row * 128 + ((j ^ (row & 7)) << 4);
        const uint4 h = *reinterpret_cast<const uint4*>(a_hi + off);
        const uint4 l = *reinterpret_cast<const uint4*>(a_hi + A_HALF + off);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
          const float v1 = __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(lw[e] & 0xFFFF0000u);
          xx = fmaf(v0, v0, xx); xx = fmaf(v1, v1, xx);
        }
      }
      // ---- running top-2 of s_k = x.E_k - |E_k|^2/2 over all code tiles ----
      float g1 = -INFINITY, g2 = -INFINITY; int gi = 0;
      for (int j = 0; j < NT; ++j, ++it) {
        const unsigned as = it & 1;
        mbar_wait(smem_u32(tfull + as * RT + rt), (it >> 1) & 1, p.err, 6);
        mbar_wait(smem_u32(nhfull + it % NHS), (it / NHS) & 1, p.err, 9);  // acquire the bulk-copied -|E|^2/2 slot
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * RT + rt) * BN);
        const float4* nh = reinterpret_cast<const float4*>(sN + (size_t)(it % NHS) * BN);
        float t1 = -INFINITY, t2 = -INFINITY;
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t (&cur)[32] = (c & 1) ? vb : va;
          uint32_t (&nxt)[32] = (c & 1) ? va : vb;
          tmem_ld_wait();
          if (c + 1 < BN / 32) tmem_ld32(taddr + (c + 1) * 32, nxt);     // overlaps with the math below
          if (p.dbg & 1) { t1 = fmaxf(t1, __uint_as_float(cur[0] ^ cur[13] ^ cur[31])); continue; }
          top2_chunk(cur, nh + c * 8, c * 32, mask, t1, t2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_after(smem_u32(tempty + as * RT + rt), smem_u32(tmem_slot + 1), t1, t2);
        // merge the tile-local top-2 into the row's running top-2
        const int ti = j * BN + (int)(__float_as_uint(t1) & 127u);
        if (t1 > g1) { g2 = fmaxf(g1, t2); g1 = t1; gi = ti; }
        else { g2 = fmaxf(g2, t1); }
      }
      if (row < rows) {
        const long long n = n0 + row;
        // rigorous bound on the error of one score: split remainder (3*2^-18) + fp32 accumulation over
        // 192 products + rounding of the -|E|^2/2 add + 7 index bits packed into the mantissa
        const float xn = sqrtf(xx) * 1.0001f;
        const float mag = xn * emax;
        const float thr = 2.5f * score_error_bound(mag, emax, D);
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

bool assign_tc_eligible(const ZView& z, int K, int D) {
  return D == tc::D && z.C == tc::D && K >= 1 && z.N >= 1;
}

// workspace: [0] int32 list_count, [1] int32 error word, [64 ..) int32 row list (N entries); for N <= SPLIT_MAX_ROWS
// additionally N 64-bit merge keys (8-byte aligned) so that the exact kernel can split short work lists over codes
constexpr long long SPLIT_MAX_ROWS = 262144;
static size_t tc_keys_offset(long long N) { return (256 + (size_t)(N > 0 ? N : 0) * sizeof(int32_t) + 7) & ~(size_t)7; }
size_t assign_tc_workspace_bytes(long long N) {
  const size_t base = 256 + (size_t)(N > 0 ? N : 0) * sizeof(int32_t);
  return (N > 0 && N <= SPLIT_MAX_ROWS) ? tc_keys_offset(N) + (size_t)N * sizeof(unsigned long long) : base;
}

bool assign_tc_can_fuse_residual(const ZView& z, const float* r_out) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(z.p) | reinterpret_cast<uintptr_t>(r_out)) & 15) == 0;
  if (!aligned || z.C != tc::D) return false;
  if (z.mode == Z_ROW) return z.T == 1 && z.sB == tc::D;       // [N,64] rows == contiguous [N,64,1]
  return z.mode == Z_BCT && z.T <= tcc::TILE_M;
}

int launch_assign_tc(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                     int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream, const int32_t* prev_idx, const float* prev_E, int prev_K, float* r_out) {
  using namespace tc;
  constexpr int TILE_M = tcc::TILE_M;
  VQ_CHECK_ARG(workspace_bytes >= assign_tc_workspace_bytes(z.N), VQB200_EWORKSPACE, "vq_assign(TC): workspace too small");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "vq_assign(TC): image must be 1024-byte aligned");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VQB200_EALIGN, "vq_assign(TC): workspace must be 16-byte aligned");
  static PerDevice configured_;
  std::atomic<size_t>& configured = configured_.here();
  if (!configured.load()) {
    VQ_CUDA(cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured.store(1);
  }
  int32_t* wsi = reinterpret_cast<int32_t*>(workspace);
  VQ_CUDA(cudaMemsetAsync(wsi, 0, 256, stream));
  Params p;
  p.z = z;
  p.image = reinterpret_cast<const unsigned char*>(image);
  p.neg_half_ee = reinterpret_cast<const float*>(p.image + img_tiles_bytes(K, D));
  p.info = info;
  p.K = K;
  p.NT = (int)(img_kp(K) / IMG_TILE_CODES);
  p.idx = idx;
  p.prev_idx = prev_idx; p.prev_E = prev_E; p.prev_K = prev_K; p.r_out = r_out;
  p.list = wsi + 64;
  p.list_count = wsi;
  p.err = wsi + 1;
  static const int tc_debug = [] { const char* d = getenv("VQB200_TC_DEBUG"); return d ? atoi(d) : 0; }();
  p.dbg = tc_debug;            // development knobs used for the measurements in DESIGN.md (0 in production)
  // how the raw fp32 rows reach shared memory: one bulk-TMA copy per 128-row tile when the rows of a
  // tile form one contiguous, 16-byte aligned byte range that fits the staging buffer
  p.stage_mode = STG_DIRECT;
  p.R = TILE_M;
  const bool aligned = (reinterpret_cast<uintptr_t>(z.p) & 15) == 0;
  if (aligned && z.mode == Z_ROW && ((z.T == 1 && z.sB == D) || (z.sT == D && z.sB == z.T * D))) {
    p.stage_mode = STG_ROWS;
  } else if (aligned && z.mode == Z_BCT && z.T <= TILE_M) {
    p.stage_mode = STG_BCT;                                     // whole samples per group: R = floor(128/T)*T
    p.R = (int)((TILE_M / z.T) * z.T);
  }
  p.ntiles = (z.N + (long long)RT * p.R - 1) / ((long long)RT * p.R);
  const int grid = (int)max(1LL, min(p.ntiles, (long long)sm_count()));
  vq_assign_tc_kernel<<<grid, NTHREADS, SMEM_TOTAL, stream>>>(p);
  VQ_LAUNCH_CHECK("vq_assign_tc_kernel");
  // exact re-do of the rows the filter could not prove (count lives on the device; no host sync)
  unsigned long long* keys = (z.N <= SPLIT_MAX_ROWS)
      ? reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + tc_keys_offset(z.N)) : nullptr;
  if (r_out) {
    VQ_CHECK_ARG(p.stage_mode != STG_DIRECT, VQB200_EUNSUPPORTED, "vq_assign(TC): fused residual needs a contiguous layout");
    ZView zr = z; zr.p = r_out;                    // the rows that were quantized are the NEW residual
    return launch_assign_simt(zr, E, ee, K, D, idx, best, p.list, p.list_count, z.N, stream, keys);
  }
  return launch_assign_simt(z, E, ee, K, D, idx, best, p.list, p.list_count, z.N, stream, keys);
}

}  // namespace vqb200
