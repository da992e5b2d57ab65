// vqb200 K1 (tensor-core variant, D == 64): fused distance + argmin on tcgen05 / TMEM / bulk-TMA (sm_100a).
//
// Replaces models/vqvae.py:30-38 of the reference for D == 64 (the hidden size of every BASELINE
// config) at any K.  Never materialises the N x K matrix.
//
// Exactness scheme ("one-pass fp16 group filter + exact re-rank"):
//   * filter: ONE tensor-core pass per score.  Rows are scaled by a per-row power of two, code tiles by a
//     per-tile power of two (both exact) and rounded to fp16 (11 significant bits); tcgen05 accumulates the
//     64 products in fp32.  The epilogue forms  S_k = x.E_k - |E_k|^2/2  (arg max S == arg min of the
//     reference distance) with one FFMA per score, reduces every group of 4 consecutive codes to its
//     maximum with 3-input max (0.5 ALU op per score), packs the 5-bit group id into the low mantissa
//     bits of that maximum and keeps the best two packed group maxima of the code tile (1.375 ALU ops
//     per score in total; the previous split-bf16 kernel needed 3 MMA passes and 3.5).  Per row the
//     tile results are merged into the best two groups plus the value of the third best.
//   * proof: e = rigorous bound on the error of one filter score (fp16 rounding of both operands, fp32
//     accumulation, FFMA rounding, id packing); thr = 2e + the reference's own fp32 rounding of (A + B) - 2M.
//     Every code tile reports its best group (value + id) and the value of its second best group, which bounds all
//     the others of that tile.  Per row (RowTrack) the tiles' best groups are kept as a sorted list g1..g4 (ids for
//     three) and the two largest "rest of a tile" bounds L, L2 (jL = tile of L).  With theta = g1 - thr
//     (row_decide): rest of every tile <= theta -> the arg min is provably inside the first k = 1..3 list groups
//     (kind k, needs g_{k+1} <= theta); only tile jL has a rest above theta -> inside tile jL or list groups 1..3
//     (kind 4, "wide"); anything else -> exact kernel through the work list (kind 0, 0.014 % of the rows at cfg3).
//     Codes whose norm exceeds 3|x| + 2 min_k|E_k| can never win for a row (neither exactly nor in the filter), so
//     the bound uses min(max_k|E_k|, 3|x| + 2 min_k|E_k|): dead codes of size 1e5 (the reference's EMA init,
//     models/vqvae.py:24-26) do not blow it up.
//   * finish, all in exact fp32 with the arithmetic of the CUDA-core kernel (assign_simt.cu: sequential fmaf chains,
//     d = (|x|^2 + |E|^2) - 2 x.E, ties to the lowest index -- bit-identical to the exact path):
//       - resident kernel: a kind-1 row is first resolved inside the filter kernel with its own fp16 operands
//         (resolve_group: packed-fp16 products, the best of the 4 codes must lead by thr + 4e-3 |x| R); 93 % of the
//         rows are final there and never touched again.  The rest go to a compact list;
//       - vq_rerank_finish_kernel: 4 lanes per listed row evaluate its 4 / 8 / 12 candidate codes from the
//         group-interleaved fp32 copy of the codebook (codebook.cuh, E4); one warp per wide row scans its code tile;
//       - streaming kernel (K > 1024): vq_rerank_kernel re-ranks every row (tiles of whole samples staged in smem).
//     (An in-kernel exact re-rank by the converter warps was measured first: correct, but 128 threads cannot keep
//     1.25 KB of L2 gathers per row in flight -- 7.7 ms instead of 2 ms per 10 M x 1024 launch.)
//
// Two kernels share the converter / epilogue code:
//   * vq_assign_f16_res_kernel (K <= 1024): the whole fp16 codebook (<= 128 KiB) is loaded into shared memory ONCE
//     per CTA, and row tiles become independent jobs (two in flight, each with its own pair of TMEM accumulators and
//     its own epilogue group, staggered; a ring of 3-4 row buffers).  Measured reason: with the codebook streamed
//     through a ring, every CTA re-reads it from L2 for every 256 rows, and 148 SMs streaming the same few hundred
//     KiB top out at ~16.5 B/clk/SM (4.6 TB/s chip-wide): 1050 of the 2150 cycles per code-tile pair were that stream
//     (knock-outs: no MMA, no TMEM load, no math still took 1050; tcgen05.ld itself sustains ~800 B/clk/SM).
//   * vq_assign_f16_kernel (any K): codebook tiles streamed through a 4-stage ring, two row tiles share every tile.
// Structure of both (one persistent CTA per SM, 448 / 608 threads, every hand-off through mbarriers; the resident kernel
// has a second MMA-issuing thread (warp 14): one thread needed ~1100 cycles per 128 x 128 unit, more than the tensor
// core):
//   warp 0       bulk-TMA producer: (a) prefetches the NEXT tile's raw fp32 rows (one contiguous slab
//                per 128-row group) into a ping-pong buffer, (b) streams 16 KiB fp16 codebook tiles
//                (pre-swizzled image written by ema_finalize / codebook_prepare) plus their 528 B of
//                {-|E|^2/2, 1/scale} through a 4-stage ring
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=128, K=16, kind::f16, fp16 in)
//   warps 2-9    epilogue: two groups of 4 warps, one 128-row tile each; tcgen05.ld the fp32 accumulators
//                (one row per thread, software-pipelined) and run the group filter
//   warps 10-13  converter: turns the prefetched raw rows IN PLACE into the swizzled K-major fp16 A operand
//                of the next tile (and, for RVQ stages >= 1, applies the residual update on the way); the resident
//                kernel has 8 converter warps (10-13, 15-18), two threads per row (convert_tile2)
//   TMEM: 4 accumulators of 128 columns (2 row tiles x 2 stages) = all 512 columns, so the MMAs of
//   code tile j+1 overlap the epilogue of code tile j.
#include <stdlib.h>
#include <limits.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "codebook.cuh"
#include "tc_common.cuh"

namespace vqb200 {

int launch_assign_simt(const ZView& z, const float* E, const float* ee, int K, int D,
                       int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                       long long max_rows, cudaStream_t stream, unsigned long long* keys = nullptr);
int launch_assign_simt_capped(const ZView& z, const float* E, const float* ee, int K, int D,
                              int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                              long long max_rows, cudaStream_t stream, unsigned long long* keys, long long key_cap);

// VQB200_K1_DEBUG (compile time): knock-out knobs and clock64 stamps inside the hot loops (tools/f16_stamps.py and the
// knock-out table in DESIGN.md were measured with it); the production build keeps only the launch-level knobs.
#ifdef VQB200_K1_DEBUG
#define K1DBG(x) (x)
#else
#define K1DBG(x) 0
#endif

namespace f16 {
using namespace tcc;

constexpr int RT = 2;                       // row tiles in flight per CTA
constexpr int D = 64;
constexpr int NST = 4;                      // streaming kernel: codebook ring stages
constexpr int NHS = NST + 2;                // streaming kernel: meta ring slots (reuse distance NST+2, see producer)
constexpr int A_BYTES = TILE_M * 128;       // 16384 B: fp16 A operand of one row tile
constexpr int BUF_BYTES = 2 * A_BYTES;      // 32768: raw fp32 rows, converted in place to [A | row info]
constexpr int SMEM_BUF = RT * 2 * BUF_BYTES;    // 131072: ping-pong per row tile
constexpr int SMEM_B = NST * F16_TILE_BYTES;    // 65536
constexpr int SMEM_META = NHS * F16_META_BYTES; // 3168
constexpr int SMEM_BAR = 512;
constexpr int SMEM_TOTAL = SMEM_BUF + SMEM_B + SMEM_META + SMEM_BAR;
constexpr int SMEM_MAX = 232448;
static_assert(SMEM_TOTAL <= SMEM_MAX, "shared memory budget");
constexpr int RES_MAX_NT = 8;               // resident kernel: at most 8 code tiles (K <= 1024) = 128 KiB of fp16 codebook
constexpr int RES_MAX_BUF = 4;
enum StageMode : int { STG_DIRECT = 0, STG_ROWS = 1, STG_BCT = 2 };
constexpr int NTHREADS = 448;
constexpr int NTHREADS_RES = 608;        // resident kernel: + second MMA issuer (warp 14) + 4 more converter warps (15-18)
// instruction descriptor: D=f32, A=B=f16, both K-major, N=128, M=128
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
constexpr int KIND_SHIFT = 28;              // idx[n] = best group | kind << 28 until the re-rank has run
constexpr int KIND_WIDE = 4;

__device__ __forceinline__ void converter_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Params {
  ZView z;
  const unsigned char* tiles;     // fp16 codebook tiles ...
  const float* meta;              // ... and their {-|E_k|^2/2 x 128, 1/scale, pad} records
  const float* info;              // {max |E_k|, bad flag, min |E_k|}
  int K, NT;                      // codes, number of 128-code tiles
  int stage_mode;                 // how raw z reaches shared memory (StageMode)
  int R;                          // rows per row tile: <= 128, a multiple of T for the [B,C,T] layout (whole samples)
  long long ntiles;               // streaming kernel: CTA tiles of 2*R rows; resident kernel: row tiles of R rows
  int nbuf, buf_bytes;            // resident kernel: row-buffer ring
  int32_t* idx;                   // out: best group | kind << 28 (kind = number of candidate groups, 4 = wide, 0 = work list)
  int32_t* cand2;                 // out: second / third candidate group of rows with kind >= 2 / >= 3
  int32_t* cand3;
  int32_t* list;                  // rows that need the exact kernel
  int32_t* list_count;
  int32_t* rr_list;               // resident kernel: rows that still need the exact re-rank (others are final)
  int32_t* rr_count;
  int2* wide;                     // {row, code tile} of rows whose candidates are a whole code tile + (cand1..3)
  int32_t* wide_count;
  int wide_cap;
  int wide_buckets;               // resident kernel at bandwidth-bound sizes: one list of wide rows PER CODE TILE (`wide` is then
                                  // int32 [NT][wide_cap], counts at wide_count[0..NT)), so that the re-rank can keep the tile in
                                  // shared memory (vq_rerank_wide_tile_kernel)
  int32_t* stat;                  // [0..1] rows with 2 / 3 candidate groups (only counted when dbg & 4)
  int* err;
  // fused residual update (RVQ stages >= 1): the staged rows are r_prev; the converter forms
  // r = r_prev - st(r_prev, prev_E[prev_idx]) (models/vqvae.py:94-98), stores it to r_out and quantizes THAT
  const int32_t* prev_idx;
  const float* prev_E;
  int prev_K;
  float* r_out;                   // same layout as z (contiguous); null = plain assignment
  int dbg;                        // development knobs (VQB200_TC_DEBUG): 1 = skip epilogue math, 2 = skip TMEM loads,
                                  // 4 = count multi-group rows, 8 = filter only, 64 = no MMAs, 128 = one MMA per tile
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

// ------------------------------------------------------------------------------------------
// converter: one row tile, raw fp32 rows (already in `buf` for the staged modes) -> fp16 A operand + row info, in place
// ------------------------------------------------------------------------------------------
template <class Wait>
__device__ __forceinline__ void convert_tile(const Params& p, unsigned char* buf, long long n0, int rows, int row, int kp,
                                             Wait wait_for_rows) {
  const float* raw = reinterpret_cast<const float*>(buf);
  const bool fuse = p.r_out != nullptr;
  float v[D];
  const float4* __restrict__ q4 = reinterpret_cast<const float4*>(p.prev_E + (size_t)kp * D);
  wait_for_rows();
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
    if (fuse) {
      // r = x - (x + (q - x)): the same three roundings as the stand-alone residual kernel.  The new residual goes
      // straight from the registers to r_out (a warp's 32 consecutive rows cover whole samples, so the 4-byte stores of
      // one component merge into full sectors in L2); staging it in shared memory for a bulk store cost two more
      // converter barriers and the store's read latency per row tile and made the converter the slowest role
      // (12 k cycles per job instead of 3.4 k: stages >= 1 took 4.05 ms instead of 3.0 ms per 10 M x 1024).
      float* __restrict__ dst = p.r_out + ((p.stage_mode == STG_ROWS) ? n * D : (n / p.z.T) * (long long)(D * p.z.T) + (n % p.z.T));
      const int so = (p.stage_mode == STG_ROWS) ? 1 : (int)p.z.T;
      // the codeword in two batches of 8 loads (all of a batch in flight together: the stores below may not be
      // reordered with loads, so a load per 4 components would cost 16 L2 round trips per row tile)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 q[D / 8];
#pragma unroll
        for (int c = 0; c < D / 8; ++c) q[c] = __ldg(q4 + h * (D / 8) + c);
#pragma unroll
        for (int c8 = 0; c8 < D / 8; ++c8) {
          const int c = h * (D / 8) + c8;
          const float qv[4] = {q[c8].x, q[c8].y, q[c8].z, q[c8].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x = v[4 * c + e];
            v[4 * c + e] = __fsub_rn(x, __fadd_rn(x, __fsub_rn(qv[e], x)));
          }
          if (p.stage_mode == STG_ROWS) {
            *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[(4 * c + e) * so] = v[4 * c + e];
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = 0.f;
  }
  converter_sync();                          // every raw read of this buffer is done
  // per-row power-of-two scale: the largest |component| lands in [2^10, 2^11) (fp16 keeps 11 bits of every
  // component down to 2^-24 of the row maximum); |x|^2 in exact fp32 for the error bound
  float m = 0.f, xx = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) { m = fmaxf(m, fabsf(v[k])); xx = fmaf(v[k], v[k], xx); }
  float sx = 1.0f, inv = 1.0f;
  {
    const int eb = (int)((__float_as_uint(m) >> 23) & 255u);
    if (m != 0.f) {
      if (eb < 27 || eb == 255 || !(xx < 3.0e38f)) inv = -1.0f;       // tiny / non-finite row: exact kernel
      else { sx = __uint_as_float((unsigned)(264 - eb) << 23); inv = __uint_as_float((unsigned)(eb - 10) << 23); }
    }
    if (!(m == m) || !(xx == xx)) inv = -1.0f;                         // NaN components (fmaxf drops them)
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t hw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 h = __floats2half2_rn(v[j * 8 + 2 * e] * sx, v[j * 8 + 2 * e + 1] * sx);
      hw[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const int off = row * 128 + ((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(buf + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  }
  *reinterpret_cast<float2*>(buf + A_BYTES + row * 8) = make_float2(inv, xx);
  fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
  converter_sync();
}


// Resident kernel: the same conversion with TWO threads per row (8 converter warps).  Thread (row, h) owns the 16-byte
// chunks j = h, h + 2, h + 4, h + 6 of the row's A operand, i.e. dims 8j .. 8j+7: half the registers per thread (no
// spills with the residual update fused in), the codeword arrives in ONE batch of 8 loads per thread instead of two, and
// twice as many loads / stores are in flight.  The two partner threads are adjacent lanes (shuffles for max and |x|^2);
// their shared-memory reads are 80 floats apart in the [B,C,T] layout (bank + 16: no conflict between partners).
__device__ __forceinline__ void converter_sync2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
template <class Wait>
__device__ __forceinline__ void convert_tile2(const Params& p, unsigned char* buf, long long n0, int rows, int row, int h,
                                              int kp, Wait wait_for_rows) {
  const float* raw = reinterpret_cast<const float*>(buf);
  const bool fuse = p.r_out != nullptr;
  float v[D / 2];                               // v[8 * i + e] = dim 8 * (h + 2 i) + e
  const float4* __restrict__ q4 = reinterpret_cast<const float4*>(p.prev_E + (size_t)kp * D);
  wait_for_rows();
  if (row < rows) {
    const long long n = n0 + row;
    if (p.stage_mode == STG_ROWS) {
      const float4* src = reinterpret_cast<const float4*>(raw + row * D);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = h + 2 * i;
        const float4 f0 = src[2 * j], f1 = src[2 * j + 1];
        v[8 * i + 0] = f0.x; v[8 * i + 1] = f0.y; v[8 * i + 2] = f0.z; v[8 * i + 3] = f0.w;
        v[8 * i + 4] = f1.x; v[8 * i + 5] = f1.y; v[8 * i + 6] = f1.z; v[8 * i + 7] = f1.w;
      }
    } else if (p.stage_mode == STG_BCT) {
      const int T = (int)p.z.T;
      const long long b = n / T; const int t = (int)(n - b * T);
      const float* src = raw + (b - n0 / T) * (D * T) + t;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[8 * i + e] = src[(8 * (h + 2 * i) + e) * T];
    } else {
      const float* src = p.z.p + p.z.row_base(n);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[8 * i + e] = __ldg(src + (long long)(8 * (h + 2 * i) + e) * p.z.sC);
    }
    if (fuse) {
      // r = x - (x + (q - x)): the same three roundings as the stand-alone residual kernel; straight to r_out
      float* __restrict__ dst = p.r_out + ((p.stage_mode == STG_ROWS) ? n * D : (n / p.z.T) * (long long)(D * p.z.T) + (n % p.z.T));
      const int so = (p.stage_mode == STG_ROWS) ? 1 : (int)p.z.T;
      float4 q[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { q[2 * i] = __ldg(q4 + 2 * (h + 2 * i)); q[2 * i + 1] = __ldg(q4 + 2 * (h + 2 * i) + 1); }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float qv[8] = {q[2 * i].x, q[2 * i].y, q[2 * i].z, q[2 * i].w, q[2 * i + 1].x, q[2 * i + 1].y, q[2 * i + 1].z, q[2 * i + 1].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = v[8 * i + e];
          v[8 * i + e] = __fsub_rn(x, __fadd_rn(x, __fsub_rn(qv[e], x)));
        }
        const int d0 = 8 * (h + 2 * i);
        if (p.stage_mode == STG_ROWS) {
          *reinterpret_cast<float4*>(dst + d0) = make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]);
          *reinterpret_cast<float4*>(dst + d0 + 4) = make_float4(v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) dst[(d0 + e) * so] = v[8 * i + e];
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < D / 2; ++k) v[k] = 0.f;
  }
  converter_sync2();                         // every raw read of this buffer is done
  float m = 0.f, xx = 0.f;
#pragma unroll
  for (int k = 0; k < D / 2; ++k) { m = fmaxf(m, fabsf(v[k])); xx = fmaf(v[k], v[k], xx); }
  // NaN components: fmaxf drops them, the sum keeps them
  const bool nan_here = !(xx == xx);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  xx += __shfl_xor_sync(0xffffffffu, xx, 1);
  const bool bad = __shfl_xor_sync(0xffffffffu, (int)nan_here, 1) != 0 || nan_here;
  float sx = 1.0f, inv = 1.0f;
  {
    const int eb = (int)((__float_as_uint(m) >> 23) & 255u);
    if (m != 0.f) {
      if (eb < 27 || eb == 255 || !(xx < 3.0e38f)) inv = -1.0f;       // tiny / non-finite row: exact kernel
      else { sx = __uint_as_float((unsigned)(264 - eb) << 23); inv = __uint_as_float((unsigned)(eb - 10) << 23); }
    }
    if (bad || !(m == m) || !(xx == xx)) inv = -1.0f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = h + 2 * i;
    uint32_t hw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hh = __floats2half2_rn(v[8 * i + 2 * e] * sx, v[8 * i + 2 * e + 1] * sx);
      hw[e] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    const int off = row * 128 + ((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(buf + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  }
  if (h == 0) *reinterpret_cast<float2*>(buf + A_BYTES + row * 8) = make_float2(inv, xx);
  fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
  converter_sync2();
}

// ------------------------------------------------------------------------------------------
// epilogue pieces
// ------------------------------------------------------------------------------------------
// Best two packed group maxima of one 32-column chunk (8 groups of 4 codes).  S = acc * c + nh.
__device__ __forceinline__ void group_chunk(const uint32_t (&cur)[32], const float4* nh, float c, int gbase,
                                            float& t1, float& t2) {
  const uint32_t mask = 0xFFFFFFE0u;
  float pk[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 h = nh[g];
    const float s0 = fmaf(__uint_as_float(cur[4 * g + 0]), c, h.x);
    const float s1 = fmaf(__uint_as_float(cur[4 * g + 1]), c, h.y);
    const float s2 = fmaf(__uint_as_float(cur[4 * g + 2]), c, h.z);
    const float s3 = fmaf(__uint_as_float(cur[4 * g + 3]), c, h.w);
    pk[g] = pack_col(fmaxf(fmax3(s0, s1, s2), s3), (uint32_t)(gbase + g), mask);
  }
#pragma unroll
  for (int g = 0; g < 8; g += 2) {
    const float hi = fmaxf(pk[g], pk[g + 1]);
    const float lo = fminf(pk[g], pk[g + 1]);
    const float tn = fmaxf(t1, hi);
    const float m = fminf(t1, hi);
    t2 = fmax3(t2, lo, m);
    t1 = tn;
  }
}

// development (VQB200_TC_DEBUG & 2048): the same filter on the raw accumulators (what folding scale and -|E|^2/2 into the
// MMA would leave in the epilogue)
__device__ __forceinline__ void group_chunk_raw(const uint32_t (&cur)[32], int gbase, float& t1, float& t2) {
  const uint32_t mask = 0xFFFFFFE0u;
  float pk[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float s0 = __uint_as_float(cur[4 * g + 0]), s1 = __uint_as_float(cur[4 * g + 1]);
    const float s2 = __uint_as_float(cur[4 * g + 2]), s3 = __uint_as_float(cur[4 * g + 3]);
    pk[g] = pack_col(fmaxf(fmax3(s0, s1, s2), s3), (uint32_t)(gbase + g), mask);
  }
#pragma unroll
  for (int g = 0; g < 8; g += 2) {
    const float hi = fmaxf(pk[g], pk[g + 1]);
    const float lo = fminf(pk[g], pk[g + 1]);
    const float tn = fmaxf(t1, hi);
    const float m = fminf(t1, hi);
    t2 = fmax3(t2, lo, m);
    t1 = tn;
  }
}

// one 128-code accumulator -> best two packed group maxima of the tile (software-pipelined tcgen05.ld)
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, const float* meta, float inv_sx, int dbg, float& t1, float& t2) {
  const float4* nh = reinterpret_cast<const float4*>(meta);
  const float c = inv_sx * meta[BN];          // 1 / (row scale * tile scale)
  t1 = -INFINITY; t2 = -INFINITY;
  if (K1DBG(dbg & 2)) return;                 // development: no TMEM reads at all
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll
  for (int ch = 0; ch < BN / 32; ++ch) {
    uint32_t (&cur)[32] = (ch & 1) ? vb : va;
    uint32_t (&nxt)[32] = (ch & 1) ? va : vb;
    tmem_ld_wait();
    if (ch + 1 < BN / 32) tmem_ld32(taddr + (ch + 1) * 32, nxt);     // overlaps with the math below
    if (K1DBG(dbg & 1)) { t1 = fmaxf(t1, __uint_as_float(cur[0] ^ cur[13] ^ cur[31])); continue; }
    if (K1DBG(dbg & 2048)) { group_chunk_raw(cur, ch * 8, t1, t2); continue; }
    group_chunk(cur, nh + ch * 8, c, ch * 8, t1, t2);
  }
}

// Per-row running state over the code tiles.  Every tile contributes its best group (with id) and the value of its
// second best group, which bounds every other group of the tile.  (g1..g4) is the sorted list of the tiles' best
// groups (ids for the best three); L / L2 are the largest and second largest per-tile bound on "the rest of the
// tile", jL the tile of L.
struct RowTrack {
  float g1, g2, g3, g4, L, L2;
  int i1, i2, i3, jL;
  __device__ __forceinline__ void init() {
    g1 = g2 = g3 = g4 = L = L2 = -INFINITY; i1 = i2 = i3 = jL = 0;
  }
  __device__ __forceinline__ void insert(float v, int id) {
    if (v > g1) { g4 = g3; g3 = g2; i3 = i2; g2 = g1; i2 = i1; g1 = v; i1 = id; }
    else if (v > g2) { g4 = g3; g3 = g2; i3 = i2; g2 = v; i2 = id; }
    else if (v > g3) { g4 = g3; g3 = v; i3 = id; }
    else g4 = fmaxf(g4, v);
  }
  __device__ __forceinline__ void merge(float t1, float t2, int j) {
    insert(t1, j * 32 + (int)(__float_as_uint(t1) & 31u));
    if (t2 > L) { L2 = L; L = t2; jL = j; } else L2 = fmaxf(L2, t2);
  }
};

// Verdict of one row (see the header comment).  theta = g1 - thr: everything at or below it is provably not the arg min.
//   rest of every tile <= theta           : candidates = the first k list entries, k = 1..3 (needs g_{k+1} <= theta)
//   only tile jL has a rest above theta   : candidates = every group of tile jL + list entries 1..3 (needs g4 <= theta)
//   otherwise                             : exact kernel (kind 0)
__device__ __forceinline__ uint32_t row_decide(const RowTrack& tr, float inv_sx, float xx, float emax, float nmin,
                                               bool cb_bad, float& thr, float& mag) {
  const float xn = sqrtf(xx) * 1.0001f;
  const float Rr = fminf(emax, fmaf(3.0f, xn, 2.0f * nmin));      // norm above which a code cannot win for this row
  mag = xn * Rr;
  // bound on the error of one filter score: fp16 rounding of x and E (2 * 2^-11), fp32 accumulation of 64
  // products, the FFMA and the 5 id bits (relative to |S| <= mag + R^2/2), fp16 underflow inside a tile
  const float e = 1.0e-3f * mag + 4.2e-6f * (mag + 0.5f * Rr * Rr) + 1.2e-10f * xn * emax;
  thr = 2.0f * e + 2.4e-7f * (xx + Rr * Rr);                      // + the reference's own rounding of (A + B) - 2M
  const bool ok = !cb_bad && (inv_sx > 0.f) && (fabsf(tr.g1) < 1e37f) && (mag < 1e37f) && (thr < 1e37f);
  const float theta = tr.g1 - thr;
  uint32_t kind = 0;
  if (ok) {
    if (tr.L <= theta) kind = (tr.g2 <= theta) ? 1u : (tr.g3 <= theta) ? 2u : (tr.g4 <= theta) ? 3u : 0u;
    else if (tr.L2 <= theta && tr.g4 <= theta) kind = KIND_WIDE;
  }
  return kind;
}

// Outputs of one row.  final_code >= 0: the filter has already proven the single winner (resident kernel), nothing is
// left to do for the row.  Called by whole warps (the re-rank list is appended with one atomic per warp).
__device__ __forceinline__ void row_emit(const Params& p, const RowTrack& tr, bool valid, uint32_t kind, int final_code,
                                         long long n, int lane) {
  if (valid && kind == KIND_WIDE) {
    if (p.wide_buckets) {
      const int pos = atomicAdd(p.wide_count + tr.jL, 1);
      if (pos < p.wide_cap) reinterpret_cast<int32_t*>(p.wide)[(size_t)tr.jL * p.wide_cap + pos] = (int32_t)n;
      else kind = 0;
    } else {
      const int pos = atomicAdd(p.wide_count, 1);
      if (pos < p.wide_cap) p.wide[pos] = make_int2((int)n, tr.jL);
      else kind = 0;
    }
  }
  if (valid) {
    if (final_code >= 0) {
      p.idx[n] = final_code;
    } else {
      p.idx[n] = (int32_t)((uint32_t)tr.i1 | (kind << KIND_SHIFT));
      if (kind >= 2) p.cand2[n] = tr.i2;
      if (kind >= 3) p.cand3[n] = tr.i3;
      if (kind == 0) {
        const int pos = atomicAdd(p.list_count, 1);
        p.list[pos] = (int32_t)n;
      }
    }
    if ((p.dbg & 4) && final_code < 0 && kind >= 1 && kind <= 3) atomicAdd(p.stat + (kind == 1 ? 0 : 1), 1);
  }
  if (p.rr_list) {
    const bool need = valid && final_code < 0 && kind >= 1 && kind <= 3;
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(p.rr_count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (need) p.rr_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)n;
    }
  }
}

// Resident kernel: resolve the 4 codes of the single candidate group with the filter's own operands (fp16 row of the A
// operand, fp16 codebook tile in shared memory).  Products and sums of 4 are formed in packed fp16 (HFMA2, the row
// pre-scaled by 2^-12 so that nothing overflows), the 16 partial sums per code are added in fp32: at most
// 2^-9 |x||E| of extra error, which thr_r carries.  If the best code leads the other three by more than thr_r it is
// provably the exact arg min and the row needs no re-rank at all (~93 % of the rows).
// all four codes in flight (and two partial sums per code).  With the bank conflicts of the loads gone: 2.77 ms per
// 10 M x 1024 (one code at a time 2.82, two 2.87); before that fix two were best (3.04 vs 3.07 / 3.09)
#ifndef RESOLVE_UNROLL
#define RESOLVE_UNROLL 4
#endif
constexpr int RESOLVE_UNROLL_N = RESOLVE_UNROLL;
__device__ __forceinline__ int resolve_group(const unsigned char* abuf, int row, const unsigned char* sCB, const float* sM,
                                             int grp, float inv_sx, float thr_r) {
  __half2 xh[D / 2];
  const __half2 down = __float2half2_rn(0.000244140625f);       // 2^-12
  // Every lane walks the 8 16-byte chunks of a row in its own order, chunk k = sg ^ j at step j.  The code rows a warp
  // reads in one step are rows 4 g + c of 32 different groups g: (row & 7) takes only TWO values, so with the natural
  // order all 32 lanes hit two 16-byte bank groups (16-way conflict: the 40 loads per row were ~60 % of this function's
  // time).  sg = (row & 7) times x in GF(8): both sg and sg ^ (row & 7) are permutations of 0..7 over 8 consecutive
  // rows, so the lanes of a quarter-warp spread over all bank groups for the code rows (at most 2-way) AND for their
  // own rows (conflict-free).  The sum over chunks is order-independent within the error bound.
  // (The code rows are 4 g + c: bit 2 of the row index is the parity of the group; folding it into the lane's order makes
  // the code-row slot sg ^ j ^ c, the same permutation for every lane of the warp whatever its group.)
  const int r7 = row & 7;
  const int sg = ((r7 << 1) & 7) ^ ((r7 & 4) ? 3 : 0) ^ ((grp & 1) << 2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 h = *reinterpret_cast<const uint4*>(abuf + row * 128 + ((((sg ^ j)) ^ r7) << 4));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) xh[4 * j + e] = __hmul2(*reinterpret_cast<const __half2*>(&hw[e]), down);
  }
  const int jt = grp >> 5, r0 = (grp & 31) * 4;
  const float* meta = sM + (size_t)jt * F16_META_FLOATS;
  const float cc = inv_sx * meta[BN] * 4096.0f;
  float s1 = -INFINITY, s2 = -INFINITY; int c1 = 0;
#pragma unroll RESOLVE_UNROLL_N
  for (int c = 0; c < 4; ++c) {
    const int r = r0 + c;
    const unsigned char* er = sCB + (size_t)jt * F16_TILE_BYTES + r * 128;
    float acc = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 h = *reinterpret_cast<const uint4*>(er + (((sg ^ j) ^ (r & 7)) << 4));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      __half2 a2 = __hmul2(xh[4 * j], *reinterpret_cast<const __half2*>(&hw[0]));
#pragma unroll
      for (int e = 1; e < 4; ++e) a2 = __hfma2(xh[4 * j + e], *reinterpret_cast<const __half2*>(&hw[e]), a2);
      const float2 f = __half22float2(a2);
      if (j & 1) { acc1 += f.x; acc1 += f.y; } else { acc += f.x; acc += f.y; }
    }
    acc += acc1;
    const float sc = fmaf(acc, cc, meta[r]);
    if (sc > s1) { s2 = s1; s1 = sc; c1 = c; } else s2 = fmaxf(s2, sc);
  }
  return (s1 - s2 > thr_r) ? grp * 4 + c1 : -1;
}

// ------------------------------------------------------------------------------------------
// streaming kernel (any K): codebook tiles through a ring, two row tiles share every tile
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
vq_assign_f16_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBuf = smem;                     // [RT][2][32768]  raw rows -> A operand + row info (in place)
  unsigned char* sB = smem + SMEM_BUF;            // [NST][16384]    codebook tiles
  float* sM = reinterpret_cast<float*>(sB + SMEM_B);              // [NHS][132]  meta of in-flight code tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SMEM_B + SMEM_META);
  uint64_t* full = bars;                 // [NST]        codebook tile landed
  uint64_t* empty = full + NST;          // [NST]        codebook tile consumed by the MMAs
  uint64_t* tfull = empty + NST;         // [2][RT]      accumulator ready
  uint64_t* tempty = tfull + 2 * RT;     // [2][RT]      accumulator drained
  uint64_t* rawfull = tempty + 2 * RT;   // [RT][2]      raw rows landed in the ping-pong buffer
  uint64_t* afull = rawfull + 2 * RT;    // [RT][2]      A operand + row info written
  uint64_t* aempty = afull + 2 * RT;     // [RT][2]      A operand no longer read by the tensor core
  uint64_t* mfull = aempty + 2 * RT;     // [NHS]        meta slot landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mfull + NHS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int s = 0; s < NHS; ++s) mbar_init(smem_u32(mfull + s), 1);
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
      int j_raw = min(NST, NT - 1);               // by then the MMAs of the previous tile are done
      if ((p.dbg >> 12) & 15) j_raw = min((p.dbg >> 12) & 15, NT - 1);      // development: move the raw prefetch
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
        for (int j = 0; j < NT; ++j, ++it) {
          const unsigned s = it % NST, ph = (it / NST) & 1;
          mbar_wait(smem_u32(empty + s), ph ^ 1, p.err, 1);
          if (j == j_raw) issue_raw(tile + gridDim.x, tile_i + 1);
          mbar_expect_tx(smem_u32(full + s), F16_TILE_BYTES);
          bulk_g2s(smem_u32(sB + (size_t)s * F16_TILE_BYTES), p.tiles + (size_t)j * F16_TILE_BYTES, F16_TILE_BYTES,
                   smem_u32(full + s));
          // the epilogue of code tile `it` reads slot it % NHS until MMA(it+2) may start; this copy is issued
          // after MMA(it+NHS-NST) = MMA(it+2) has completed, so the slot is free AND its barrier cannot run a
          // phase ahead of the epilogue's parity wait (NHS = NST + 2)
          mbar_expect_tx(smem_u32(mfull + it % NHS), F16_META_BYTES);
          bulk_g2s(smem_u32(sM + (size_t)(it % NHS) * F16_META_FLOATS), p.meta + (size_t)j * F16_META_FLOATS,
                   F16_META_BYTES, smem_u32(mfull + it % NHS));
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
          const uint32_t b = smem_u32(sB + (size_t)s * F16_TILE_BYTES);
#pragma unroll
          for (int rt = 0; rt < RT; ++rt) {
            if (j == 0) mbar_wait(smem_u32(afull + rt * 2 + pp), u & 1, p.err, 3);
            mbar_wait(smem_u32(tempty + as * RT + rt), ((it >> 1) & 1) ^ 1, p.err, 4);
            tc_fence_after();
            const uint32_t a = smem_u32(sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES);
            const uint32_t d_tmem = tmem_base + (uint32_t)((as * RT + rt) * BN);
            const int nk = K1DBG(p.dbg & 64) ? 0 : (K1DBG(p.dbg & 128) ? 1 : 4);     // development: fewer / no MMAs
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nk) umma_bf16(d_tmem, umma_desc(a + k * 32), umma_desc(b + k * 32), IDESC_F16, k ? 1u : 0u);
            umma_commit(smem_u32(tfull + as * RT + rt));
            if (j == NT - 1) umma_commit(smem_u32(aempty + rt * 2 + pp));
          }
          umma_commit(smem_u32(empty + s));
        }
      }
    }
  } else if (warp >= 10) {
    // ================= converter: raw fp32 rows -> swizzled fp16 A operand, in place =================
    const int row = (warp - 10) * 32 + lane;      // 0..127: this thread's row inside the group
    const bool fuse = p.r_out != nullptr;         // staged modes only (checked by the launcher)
    unsigned tile_i = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
      const unsigned pp = tile_i & 1, u = tile_i >> 1;
#pragma unroll 1
      for (int rt = 0; rt < RT; ++rt) {
        const long long n0 = tile * tile_rows + (long long)rt * R;
        const int rows = (int)max(0LL, min((long long)R, p.z.N - n0));
        unsigned char* buf = sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES;
        int kp = 0;                                 // previous stage's code of this row, fetched before the wait
        if (fuse && row < rows) kp = min(max(__ldg(p.prev_idx + n0 + row), 0), p.prev_K - 1);
        convert_tile(p, buf, n0, rows, row, kp, [&]() {
          if (staged && rows > 0) mbar_wait(smem_u32(rawfull + rt * 2 + pp), u & 1, p.err, 8);
          else mbar_wait(smem_u32(aempty + rt * 2 + pp), (u & 1) ^ 1, p.err, 5);   // nobody fills it for us: wait until free
        });
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
    const float nmin = p.info[2];
    unsigned it = 0, tile_i = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tile_i) {
      const unsigned pp = tile_i & 1, u = tile_i >> 1;
      const long long n0 = tile * tile_rows + (long long)rt * R;
      const int rows = (int)max(0LL, min((long long)R, p.z.N - n0));
      mbar_wait(smem_u32(afull + rt * 2 + pp), u & 1, p.err, 10);
      const float2 ri = *reinterpret_cast<const float2*>(sBuf + (size_t)(rt * 2 + pp) * BUF_BYTES + A_BYTES + row * 8);
      RowTrack tr; tr.init();
      for (int j = 0; j < NT; ++j, ++it) {
        const unsigned as = it & 1;
        mbar_wait(smem_u32(tfull + as * RT + rt), (it >> 1) & 1, p.err, 6);
        mbar_wait(smem_u32(mfull + it % NHS), (it / NHS) & 1, p.err, 9);   // acquire the bulk-copied meta slot
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * RT + rt) * BN);
        float t1, t2;
        epilogue_tile(taddr, sM + (size_t)(it % NHS) * F16_META_FLOATS, ri.x, p.dbg, t1, t2);
        // The arrive below hands the accumulator back to the MMA issuer and (through MMA(it+2) -> empty -> producer) lets
        // this tile's meta slot be refilled.  ptxas hoists a bare arrive above the last chunk's math; with the row-major
        // layout (the converter's bank-conflicted loads slow the epilogue warps down) that was measured to mis-assign
        // ~0.04 % of the rows, run to run different.  arrive_after() orders the arrive behind a store of t1/t2.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_after(smem_u32(tempty + as * RT + rt), smem_u32(tmem_slot + 1), t1, t2);
        tr.merge(t1, t2, j);
      }
      float thr, mag;
      const uint32_t kind = (row < rows) ? row_decide(tr, ri.x, ri.y, emax, nmin, cb_bad, thr, mag) : 0u;
      row_emit(p, tr, row < rows, kind, -1, n0 + row, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// resident kernel (NT <= 8): whole codebook in shared memory, row tiles are independent staggered jobs
//   job r of this CTA = row tile blockIdx.x + r * gridDim.x, row buffer r % nbuf, TMEM slot / epilogue group r & 1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS_RES, 1)
vq_assign_f16_res_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int NT = p.NT;
  unsigned char* sCB = smem;                                          // [NT][16384]  codebook tiles (resident)
  unsigned char* sBuf = smem + (size_t)NT * F16_TILE_BYTES;           // [nbuf][buf_bytes] raw rows -> A + row info
  float* sM = reinterpret_cast<float*>(sBuf + (size_t)p.nbuf * p.buf_bytes);     // [NT][132] meta (resident)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sM) + (((size_t)NT * F16_META_BYTES + 15) & ~(size_t)15));
  uint64_t* cbfull = bars;               // [1]          codebook + meta landed
  uint64_t* tfull = cbfull + 1;          // [2 slots][2] accumulator ready
  uint64_t* tempty = tfull + 4;          // [2 slots][2] accumulator drained
  uint64_t* rawfull = tempty + 4;        // [nbuf]       raw rows landed
  uint64_t* afull = rawfull + RES_MAX_BUF;   // [nbuf]   A operand + row info written
  uint64_t* aempty = afull + RES_MAX_BUF;    // [nbuf]   A operand no longer read by the tensor core
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + RES_MAX_BUF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbuf = p.nbuf;

  if (tid == 0) {
    mbar_init(smem_u32(cbfull), 1);
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 4); }
    for (int i = 0; i < RES_MAX_BUF; ++i) {
      // a row buffer is free again when the tensor core has read its last A operand AND the job's 4 epilogue warps
      // are done with it (they re-read their fp16 row to resolve the winning group at the end of the job)
      mbar_init(smem_u32(rawfull + i), 1); mbar_init(smem_u32(afull + i), 8); mbar_init(smem_u32(aempty + i), 5);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int R = p.R;
  if ((smem_u32(smem) & 1023u) != 0u) { if (tid == 0 && p.err) atomicExch(p.err, 99); __trap(); }
  const bool staged = p.stage_mode != STG_DIRECT;
  // development (VQB200_TC_DEBUG & 512): clock64 stamps of CTA 0 into the cand3 array: [role][event][4]
  long long* stamps = (K1DBG(p.dbg & 512) && blockIdx.x == 0) ? reinterpret_cast<long long*>(p.cand3) : nullptr;
  const int njobs = (p.ntiles > blockIdx.x) ? (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  auto job_n0 = [&](int r) { return ((long long)blockIdx.x + (long long)r * gridDim.x) * R; };
  auto job_rows = [&](int r) { return (int)max(0LL, min((long long)R, p.z.N - job_n0(r))); };

  if (warp == 0) {
    // ================= bulk-TMA producer: the codebook once, then one raw slab per job =================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(cbfull), (uint32_t)NT * (F16_TILE_BYTES + F16_META_BYTES));
      for (int j = 0; j < NT; ++j)
        bulk_g2s(smem_u32(sCB + (size_t)j * F16_TILE_BYTES), p.tiles + (size_t)j * F16_TILE_BYTES, F16_TILE_BYTES, smem_u32(cbfull));
      bulk_g2s(smem_u32(sM), p.meta, (uint32_t)NT * F16_META_BYTES, smem_u32(cbfull));
      if (staged) {
        for (int r = 0; r < njobs; ++r) {
          const int b = r % nbuf;
          const unsigned u = (unsigned)(r / nbuf);
          if (u > 0) mbar_wait(smem_u32(aempty + b), (u - 1) & 1, p.err, 7);       // MMAs of job r - nbuf left the buffer
          const StagePlan sp = stage_plan(p, job_n0(r), job_rows(r));
          mbar_expect_tx(smem_u32(rawfull + b), sp.bytes);
          bulk_g2s(smem_u32(sBuf + (size_t)b * p.buf_bytes), sp.src, sp.bytes, smem_u32(rawfull + b));
          if (stamps && r < 64) stamps[4 * 1024 * 4 + r * 4] = clock64();
        }
      }
    }
  } else if (warp == 1 || warp == 14) {
    // ================= MMA issuers: one thread per job slot (slot 0: warp 1, slot 1: warp 14) =================
    // A single thread needs ~1100 cycles per unit (4 MMAs + barrier probes + commits, all serial latency), more than the
    // 2 x 271 tensor cycles it has to feed; one issuing thread per slot halves that.
    if (lane == 0) {
      const int g = (warp == 1) ? 0 : 1;
      mbar_wait(smem_u32(cbfull), 0, p.err, 2);
      unsigned cnt = 0;
      int b = g % nbuf;
      unsigned u = (unsigned)(g / nbuf);
      const uint64_t bdesc0 = umma_desc(smem_u32(sCB));
      for (int r = g; r < njobs; r += 2) {
        mbar_wait(smem_u32(afull + b), u & 1, p.err, 3);
        const uint64_t adesc = umma_desc(smem_u32(sBuf + (size_t)b * p.buf_bytes));
        for (int j = 0; j < NT; ++j, ++cnt) {
          const unsigned st = cnt & 1;
          mbar_wait(smem_u32(tempty + g * 2 + st), ((cnt >> 1) & 1) ^ 1, p.err, 4);
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + (uint64_t)(j * (F16_TILE_BYTES >> 4));     // start-address field counts 16-byte units
          const uint32_t d_tmem = tmem_base + (uint32_t)((g * 2 + st) * BN);
          if (!K1DBG(p.dbg & 64)) {
            umma_bf16(d_tmem, adesc, bdesc, IDESC_F16, 0u);
            if (!K1DBG(p.dbg & 128)) {
              umma_bf16(d_tmem, adesc + 2, bdesc + 2, IDESC_F16, 1u);
              umma_bf16(d_tmem, adesc + 4, bdesc + 4, IDESC_F16, 1u);
              umma_bf16(d_tmem, adesc + 6, bdesc + 6, IDESC_F16, 1u);
            }
          }
          umma_commit(smem_u32(tfull + g * 2 + st));
          if (stamps && cnt < 512) {
            long long* e = stamps + (g * 512 + cnt) * 4;
            e[0] = clock64(); e[1] = 0; e[2] = 0; e[3] = r * 16 + j;
          }
        }
        umma_commit(smem_u32(aempty + b));
        b += 2; if (b >= nbuf) { b -= nbuf; ++u; }
      }
    }
  } else if (warp >= 10 && warp != 14) {
    // ================= converter (8 warps, two threads per row): jobs in order =================
    const int ct = (warp < 14 ? warp - 10 : warp - 11) * 32 + lane;       // 0 .. 255
    const int row = ct >> 1, h = ct & 1;
    const bool fuse = p.r_out != nullptr;
    for (int r = 0; r < njobs; ++r) {
      const int b = r % nbuf;
      const unsigned u = (unsigned)(r / nbuf);
      const long long n0 = job_n0(r);
      const int rows = job_rows(r);
      unsigned char* buf = sBuf + (size_t)b * p.buf_bytes;
      int kp = 0;
      if (fuse && row < rows) kp = min(max(__ldg(p.prev_idx + n0 + row), 0), p.prev_K - 1);
      const long long c0 = stamps ? clock64() : 0;
      convert_tile2(p, buf, n0, rows, row, h, kp, [&]() {
        if (staged) mbar_wait(smem_u32(rawfull + b), u & 1, p.err, 8);
        else if (u > 0) mbar_wait(smem_u32(aempty + b), (u - 1) & 1, p.err, 5);    // nobody fills it for us: wait until free
      });
      if (lane == 0) mbar_arrive(smem_u32(afull + b));
      if (stamps && warp == 10 && lane == 0 && r < 256) {
        long long* e = stamps + 3 * 1024 * 4 + r * 4;
        e[0] = c0; e[1] = c0; e[2] = clock64(); e[3] = r;
      }
    }
  } else {
    // ================= epilogue groups: group g takes jobs g, g + 2, ... =================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float emax = p.info[0];
    const bool cb_bad = p.info[1] != 0.f;
    const float nmin = p.info[2];
    mbar_wait(smem_u32(cbfull), 0, p.err, 9);     // meta resident (acquire of the bulk copies)
    unsigned cnt = 0;
    for (int r = g; r < njobs; r += 2) {
      const int b = r % nbuf;
      const unsigned u = (unsigned)(r / nbuf);
      const long long n0 = job_n0(r);
      const int rows = job_rows(r);
      mbar_wait(smem_u32(afull + b), u & 1, p.err, 10);
      const float2 ri = *reinterpret_cast<const float2*>(sBuf + (size_t)b * p.buf_bytes + A_BYTES + row * 8);
      RowTrack tr; tr.init();
      for (int j = 0; j < NT; ++j, ++cnt) {
        const unsigned st = cnt & 1;
        const long long c0 = stamps ? clock64() : 0;
        mbar_wait(smem_u32(tfull + g * 2 + st), (cnt >> 1) & 1, p.err, 6);
        tc_fence_after();
        const long long c1 = stamps ? clock64() : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * 2 + st) * BN);
        float t1, t2;
        epilogue_tile(taddr, sM + (size_t)j * F16_META_FLOATS, ri.x, p.dbg, t1, t2);
        tc_fence_before();
        __syncwarp();
        // as in the streaming kernel: the accumulator is handed back only after its values were used
        if (lane == 0) arrive_after(smem_u32(tempty + g * 2 + st), smem_u32(tmem_slot + 1), t1, t2);
        tr.merge(t1, t2, j);
        if (stamps && q == 0 && lane == 0 && cnt < 1024) {
          long long* e = stamps + (1 + g) * 1024 * 4 + cnt * 4;
          e[0] = c0; e[1] = c1; e[2] = clock64(); e[3] = r * 16 + j;
        }
      }
      float thr, mag;
      const uint32_t kind = (row < rows) ? row_decide(tr, ri.x, ri.y, emax, nmin, cb_bad, thr, mag) : 0u;
      int final_code = -1;
      if (kind == 1 && !K1DBG(p.dbg & 1024))
        final_code = resolve_group(sBuf + (size_t)b * p.buf_bytes, row, sCB, sM, tr.i1, ri.x, fmaf(4.0e-3f, mag, thr));
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(aempty + b));       // the A operand and the row info are no longer needed
      row_emit(p, tr, row < rows, kind, final_code, n0 + row, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// Exact re-rank of the filter's candidates (bandwidth class: reads every row once + 1 KiB of codebook per group).
// 4 lanes per row = the 4 codes of a candidate group; every (row, code) distance is the sequential fmaf chain of
// assign_simt.cu, the 4 lanes merge with cand_better.  Contiguous layouts stage a tile of whole samples (one byte
// range) in shared memory with coalesced 16-byte loads; arbitrary views read their rows through the strides.
// ------------------------------------------------------------------------------------------
constexpr int RR_ROWS = 64;
constexpr int RR_THREADS = RR_ROWS * 4;

__device__ __forceinline__ void load_row(const ZView& z, long long n, float (&x)[D]) {
  const float* src = z.p + z.row_base(n);
#pragma unroll
  for (int k = 0; k < D; ++k) x[k] = __ldg(src + (long long)k * z.sC);
}
// exact distance of row x (|x|^2 = xx) to `code`, merged into (best, bidx)
__device__ __forceinline__ void exact_code(const float (&x)[D], float xx, const float* __restrict__ E,
                                           const float* __restrict__ ee, int code, float& best, int& bidx) {
  const float4* e4 = reinterpret_cast<const float4*>(E + (size_t)code * D);
  const float e2 = __ldg(ee + code);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 e = __ldg(e4 + q);
    acc = fmaf(x[4 * q + 0], e.x, acc); acc = fmaf(x[4 * q + 1], e.y, acc);
    acc = fmaf(x[4 * q + 2], e.z, acc); acc = fmaf(x[4 * q + 3], e.w, acc);
  }
  const float d = __fsub_rn(__fadd_rn(xx, e2), __fmul_rn(2.0f, acc));
  if (cand_better(d, code, best, bidx)) { best = d; bidx = code; }
}

// one code of an interleaved group (codebook.cuh: float4 4*q + c of the group = dims 4q..4q+3 of code 4*grp + c)
__device__ __forceinline__ void exact_code_e4(const float (&x)[D], float xx, const float4* __restrict__ e4,
                                              const float* __restrict__ ee, int grp, int c, float& best, int& bidx) {
  const float4* g4 = e4 + (size_t)grp * 64 + c;
  const int code = grp * 4 + c;
  const float e2 = __ldg(ee + code);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 e = __ldg(g4 + 4 * q);
    acc = fmaf(x[4 * q + 0], e.x, acc); acc = fmaf(x[4 * q + 1], e.y, acc);
    acc = fmaf(x[4 * q + 2], e.z, acc); acc = fmaf(x[4 * q + 3], e.w, acc);
  }
  const float d = __fsub_rn(__fadd_rn(xx, e2), __fmul_rn(2.0f, acc));
  if (cand_better(d, code, best, bidx)) { best = d; bidx = code; }
}

template <bool TILED>
__global__ void __launch_bounds__(RR_THREADS)
vq_rerank_kernel(ZView z, int rows_per_tile, const float4* __restrict__ E4, const float* __restrict__ ee, int K,
                 int32_t* __restrict__ idx, const int32_t* __restrict__ cand2, const int32_t* __restrict__ cand3) {
  __shared__ __align__(16) float X[TILED ? RR_ROWS * D : 4];
  const int tid = threadIdx.x, r = tid >> 2, c = tid & 3;
  const int T = (int)z.T;
  const long long ntiles = (z.N + rows_per_tile - 1) / rows_per_tile;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long n0 = t * rows_per_tile;
    const int rows = (int)min((long long)rows_per_tile, z.N - n0);
    const long long n = n0 + r;
    if (TILED) {
      __syncthreads();
      const float4* s4 = reinterpret_cast<const float4*>(z.p + n0 * D);     // whole samples: one contiguous range
      float4* d4 = reinterpret_cast<float4*>(X);
      for (int i = tid; i < rows * (D / 4); i += RR_THREADS) d4[i] = __ldg(s4 + i);
      __syncthreads();
    }
    const uint32_t first = (r < rows) ? (uint32_t)idx[n] : 0u;
    int kind = (int)(first >> KIND_SHIFT);
    if (kind > 3) kind = 0;                      // wide rows belong to vq_rerank_wide_kernel
    const int grp0 = (int)(first & ((1u << KIND_SHIFT) - 1));
    const int grp1 = (kind >= 2) ? __ldg(cand2 + n) : 0;
    const int grp2 = (kind == 3) ? __ldg(cand3 + n) : 0;
    float x[D];
    float xx = 0.f;
    if (kind) {
      if (TILED) {
        const float* src = X + (r / T) * (D * T) + (r % T);
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = src[k * T];
      } else {
        load_row(z, n, x);
      }
#pragma unroll
      for (int k = 0; k < D; ++k) xx = fmaf(x[k], x[k], xx);
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) x[k] = 0.f;
    }
    float best = INFINITY; int bidx = INT_MAX;
    const int kmax = __reduce_max_sync(0xffffffffu, kind);
    for (int g = 0; g < kmax; ++g) {
      const int grp = (g == 0) ? grp0 : (g == 1) ? grp1 : grp2;
      if (g < kind && grp * 4 + c < K) exact_code_e4(x, xx, E4, ee, grp, c, best, bidx);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (cand_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if (kind && c == 0) idx[n] = bidx;
  }
}

// one code of an interleaved group, the row read from shared memory (16-byte broadcast loads)
__device__ __forceinline__ void exact_code_e4s(const float4* __restrict__ xs, float xx, const float4* __restrict__ e4,
                                               const float* __restrict__ ee, int grp, int c, float& best, int& bidx) {
  const float4* g4 = e4 + (size_t)grp * 64 + c;
  const int code = grp * 4 + c;
  const float e2 = __ldg(ee + code);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 e = __ldg(g4 + 4 * q);
    const float4 x = xs[q];
    acc = fmaf(x.x, e.x, acc); acc = fmaf(x.y, e.y, acc); acc = fmaf(x.z, e.z, acc); acc = fmaf(x.w, e.w, acc);
  }
  const float d = __fsub_rn(__fadd_rn(xx, e2), __fmul_rn(2.0f, acc));
  if (cand_better(d, code, best, bidx)) { best = d; bidx = code; }
}
__device__ __forceinline__ float row_sq(const float4* xs) {
  float xx = 0.f;
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 x = xs[q];
    xx = fmaf(x.x, x.x, xx); xx = fmaf(x.y, x.y, xx); xx = fmaf(x.z, x.z, xx); xx = fmaf(x.w, x.w, xx);
  }
  return xx;
}

// List mode (resident kernel): only the rows the filter could not finish itself; 4 lanes per listed row, the row staged
// in shared memory by its 4 lanes (16 components each).
constexpr int RL_LD = D + 4;
__device__ __forceinline__ void rerank_list_body(float* X, int bid, int nblocks, const ZView& z, const float4* __restrict__ E4,
                                                 const float* __restrict__ ee, int K, int32_t* __restrict__ idx,
                                                 const int32_t* __restrict__ cand2, const int32_t* __restrict__ cand3,
                                                 const int32_t* __restrict__ rr_list, const int32_t* __restrict__ rr_count) {
  const int c = threadIdx.x & 3, slot = threadIdx.x >> 2;
  const int total = *rr_count;
  const int per_pass = nblocks * 64;
  float* xrow = X + slot * RL_LD;
  for (int base = 0; base < total; base += per_pass) {            // warp-uniform trip count (shuffles below)
    const int e = base + bid * 64 + slot;
    const bool live = e < total;
    const long long n = live ? rr_list[e] : 0;
    const uint32_t first = live ? (uint32_t)idx[n] : 0u;
    int kind = (int)(first >> KIND_SHIFT);
    if (kind > 3) kind = 0;
    const int grp0 = (int)(first & ((1u << KIND_SHIFT) - 1));
    const int grp1 = (kind >= 2) ? __ldg(cand2 + n) : 0;
    const int grp2 = (kind == 3) ? __ldg(cand3 + n) : 0;
    __syncwarp();
    if (kind) {
      const float* src = z.p + z.row_base(n) + (long long)(16 * c) * z.sC;
#pragma unroll
      for (int k = 0; k < 16; ++k) xrow[16 * c + k] = __ldg(src + (long long)k * z.sC);
    }
    __syncwarp();
    const float4* xs = reinterpret_cast<const float4*>(xrow);
    const float xx = kind ? row_sq(xs) : 0.f;
    float best = INFINITY; int bidx = INT_MAX;
    const int kmax = __reduce_max_sync(0xffffffffu, kind);
    for (int g = 0; g < kmax; ++g) {
      const int grp = (g == 0) ? grp0 : (g == 1) ? grp1 : grp2;
      if (g < kind && grp * 4 + c < K) exact_code_e4s(xs, xx, E4, ee, grp, c, best, bidx);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (cand_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if (kind && c == 0) idx[n] = bidx;
  }
}

// Wide rows: one warp per row (staged in shared memory); 4 passes of 8 groups x 4 codes over the 32 groups of the row's
// code tile, a fifth pass over the row's candidate groups 1..3 (duplicates are harmless), then the warp merges.
__device__ __forceinline__ void rerank_wide_body(float* X, int bid, int nblocks, const ZView& z, const float4* __restrict__ E4,
                                                 const float* __restrict__ ee, int K, int32_t* __restrict__ idx,
                                                 const int32_t* __restrict__ cand2, const int32_t* __restrict__ cand3,
                                                 const int2* __restrict__ wide, const int32_t* __restrict__ wide_count, int wide_cap) {
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int warp = (bid * (int)blockDim.x + (int)threadIdx.x) >> 5, nwarps = (nblocks * (int)blockDim.x) >> 5;
  float* xrow = X + (threadIdx.x >> 5) * D;
  const int total = min(*wide_count, wide_cap);
  for (int w = warp; w < total; w += nwarps) {
    const int2 e = wide[w];
    const long long n = e.x;
    __syncwarp();
    {
      const float* src = z.p + z.row_base(n);
      xrow[lane] = __ldg(src + (long long)lane * z.sC);
      xrow[lane + 32] = __ldg(src + (long long)(lane + 32) * z.sC);
    }
    __syncwarp();
    const float4* xs = reinterpret_cast<const float4*>(xrow);
    const float xx = row_sq(xs);
    float best = INFINITY; int bidx = INT_MAX;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      const int grp = e.y * 32 + pass * 8 + g;
      if (grp * 4 + c < K) exact_code_e4s(xs, xx, E4, ee, grp, c, best, bidx);
    }
    if (g < 3) {
      const int grp = (g == 0) ? (int)((uint32_t)idx[n] & ((1u << KIND_SHIFT) - 1)) : (g == 1) ? __ldg(cand2 + n) : __ldg(cand3 + n);
      if (grp * 4 + c < K) exact_code_e4s(xs, xx, E4, ee, grp, c, best, bidx);
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (cand_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if (lane == 0) idx[n] = bidx;
  }
}

// Wide rows of the resident filter at bandwidth-bound sizes, bucketed by code tile: a CTA keeps ONE tile of the interleaved
// fp32 copy (128 codes x 64 dims = 32 KiB, groups padded to 1088 B so that the 8 groups x 4 codes a warp reads per step
// fall into 8 different 16-byte bank groups) in shared memory and runs its warps over that tile's rows.  The list-based
// kernel re-read the 32 KiB per row through L1 / L2 (3.5 GB per 10 M x 1024 call, 0.28 ms); same arithmetic, same bits.
constexpr int WT_LD = 68;                      // float4 per group in shared memory (64 + 4 of padding)
__global__ void __launch_bounds__(256)
vq_rerank_wide_tile_kernel(ZView z, const float4* __restrict__ E4, const float* __restrict__ ee, int K, int NT,
                           int32_t* __restrict__ idx, const int32_t* __restrict__ cand2, const int32_t* __restrict__ cand3,
                           const int32_t* __restrict__ wide_rows, const int32_t* __restrict__ wide_counts, int wide_cap) {
  __shared__ __align__(16) float4 sE[32 * WT_LD];
  __shared__ float s_ee[128];
  __shared__ __align__(16) float X[8 * D];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3;
  const int tile = (int)blockIdx.x % NT, sub = (int)blockIdx.x / NT, nsub = (int)gridDim.x / NT;
  if (sub >= nsub) return;                                     // grid not a multiple of NT
  const int count = min(wide_counts[tile], wide_cap);
  if (sub * 8 >= count) return;
  for (int i = tid; i < 32 * 64; i += 256) sE[(i >> 6) * WT_LD + (i & 63)] = __ldg(E4 + (size_t)tile * (32 * 64) + i);
  if (tid < 128) s_ee[tid] = (tile * 128 + tid < K) ? __ldg(ee + tile * 128 + tid) : 0.f;
  __syncthreads();
  float* xrow = X + warp * D;
  const float4* xs = reinterpret_cast<const float4*>(xrow);
  // software pipeline over the warp's rows: the next row (64 scattered 4-byte reads = one DRAM round trip) and the
  // row's other candidate groups are requested before the current row's 140 codes are evaluated
  const int w0 = sub * 8 + warp, wstep = nsub * 8;
  long long n_next = (w0 < count) ? wide_rows[(size_t)tile * wide_cap + w0] : 0;
  float p0 = 0.f, p1 = 0.f;
  if (w0 < count) {
    const float* src = z.p + z.row_base(n_next);
    p0 = __ldg(src + (long long)lane * z.sC);
    p1 = __ldg(src + (long long)(lane + 32) * z.sC);
  }
  for (int w = w0; w < count; w += wstep) {
    const long long n = n_next;
    __syncwarp();
    xrow[lane] = p0; xrow[lane + 32] = p1;
    __syncwarp();
    int cg = -1;                                                   // lanes of groups 0..2: the row's candidate group g
    if (g < 3) cg = (g == 0) ? (int)((uint32_t)idx[n] & ((1u << KIND_SHIFT) - 1)) : (g == 1) ? __ldg(cand2 + n) : __ldg(cand3 + n);
    if (w + wstep < count) {
      n_next = wide_rows[(size_t)tile * wide_cap + w + wstep];
      const float* src = z.p + z.row_base(n_next);
      p0 = __ldg(src + (long long)lane * z.sC);
      p1 = __ldg(src + (long long)(lane + 32) * z.sC);
    }
    const float xx = row_sq(xs);
    float best = INFINITY; int bidx = INT_MAX;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      const int gl = pass * 8 + g;                             // group inside the tile
      const int code = (tile * 32 + gl) * 4 + c;
      if (code < K) {
        const float4* e4 = sE + gl * WT_LD + c;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
          const float4 e = e4[4 * q];
          const float4 x = xs[q];
          acc = fmaf(x.x, e.x, acc); acc = fmaf(x.y, e.y, acc); acc = fmaf(x.z, e.z, acc); acc = fmaf(x.w, e.w, acc);
        }
        const float d = __fsub_rn(__fadd_rn(xx, s_ee[gl * 4 + c]), __fmul_rn(2.0f, acc));
        if (cand_better(d, code, best, bidx)) { best = d; bidx = code; }
      }
    }
    if (g < 3 && cg * 4 + c < K) exact_code_e4s(xs, xx, E4, ee, cg, c, best, bidx);
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (cand_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if (lane == 0) idx[n] = bidx;
  }
}

// One launch for both: `lblocks` CTAs (the even ones) work on the re-rank list (bound by the 32-byte sectors of the scattered
// row reads), the others on the wide rows (bound by L1 / L2 reads of the code tile) -- disjoint rows, different
// bottlenecks, so they overlap instead of queueing (0.24 + 0.28 ms per 10 M x 1024 as two launches).  lblocks == 0:
// no list (streaming filter, its re-rank is vq_rerank_kernel).
__global__ void __launch_bounds__(256)
vq_rerank_finish_kernel(ZView z, const float4* __restrict__ E4, const float* __restrict__ ee, int K,
                        int32_t* __restrict__ idx, const int32_t* __restrict__ cand2, const int32_t* __restrict__ cand3,
                        const int32_t* __restrict__ rr_list, const int32_t* __restrict__ rr_count, int lblocks,
                        const int2* __restrict__ wide, const int32_t* __restrict__ wide_count, int wide_cap) {
  __shared__ __align__(16) float X[64 * RL_LD];
  const int bid = (int)blockIdx.x;
  if (lblocks == 0) {
    rerank_wide_body(X, bid, (int)gridDim.x, z, E4, ee, K, idx, cand2, cand3, wide, wide_count, wide_cap);
  } else if (2 * lblocks == (int)gridDim.x) {       // both kinds, interleaved so that they are resident together
    if (bid & 1) rerank_wide_body(X, bid >> 1, lblocks, z, E4, ee, K, idx, cand2, cand3, wide, wide_count, wide_cap);
    else rerank_list_body(X, bid >> 1, lblocks, z, E4, ee, K, idx, cand2, cand3, rr_list, rr_count);
  } else {                                          // list only
    rerank_list_body(X, bid, lblocks, z, E4, ee, K, idx, cand2, cand3, rr_list, rr_count);
  }
}

}  // namespace f16

bool assign_f16_eligible(const ZView& z, int K, int D) {
  return D == f16::D && z.C == f16::D && K >= 1 && z.N >= 1;
}

// workspace: 64 int32 header ([0] list_count, [1] error word, [2..3] rows with 2 / 3 candidate groups (debug),
// [4] wide_count, [5] re-rank list count), then int32 arrays: row list (N), second and third candidate group, re-rank
// list (N each), wide records
// (RES_MAX_NT * wide_cap: one {row, tile} list, or one row list per code tile); for N <= SPLIT_MAX_ROWS additionally N 64-bit merge keys (8-byte aligned) so that the exact kernel
// can split short work lists over codes
constexpr long long F16_SPLIT_MAX_ROWS = 262144;
static long long f16_wide_cap(long long N) { return N / 8 + 64; }
static size_t f16_n2(long long N) { return ((size_t)(N > 0 ? N : 0) + 1) & ~(size_t)1; }      // per-row arrays keep 8-byte alignment
static size_t f16_ints(long long N) { return 64 + 4 * f16_n2(N) + (size_t)f16::RES_MAX_NT * (size_t)f16_wide_cap(N > 0 ? N : 0); }
static size_t f16_keys_offset(long long N) { return (f16_ints(N) * sizeof(int32_t) + 7) & ~(size_t)7; }
// merge keys of the exact kernel: one per LISTED row, at most F16_SPLIT_MAX_ROWS of them (longer lists are swept unsplit)
size_t assign_f16_workspace_bytes(long long N) {
  return (N > 0) ? f16_keys_offset(N) + (size_t)min(N, F16_SPLIT_MAX_ROWS) * sizeof(unsigned long long)
                 : f16_ints(N) * sizeof(int32_t);
}

bool assign_f16_can_fuse_residual(const ZView& z, const float* r_out) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(z.p) | reinterpret_cast<uintptr_t>(r_out)) & 15) == 0;
  if (!aligned || z.C != f16::D) return false;
  if (z.mode == Z_ROW) return z.T == 1 && z.sB == f16::D;       // [N,64] rows == contiguous [N,64,1]
  return z.mode == Z_BCT && z.T <= tcc::TILE_M;
}

int launch_assign_f16(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                      int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream, const int32_t* prev_idx, const float* prev_E, int prev_K, float* r_out) {
  using namespace f16;
  constexpr int TILE_M = tcc::TILE_M;
  VQ_CHECK_ARG(workspace_bytes >= assign_f16_workspace_bytes(z.N), VQB200_EWORKSPACE, "vq_assign(TC): workspace too small");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 1023) == 0, VQB200_EALIGN, "vq_assign(TC): image must be 1024-byte aligned");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VQB200_EALIGN, "vq_assign(TC): workspace must be 16-byte aligned");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(E) & 15) == 0, VQB200_EALIGN, "vq_assign(TC): codebook must be 16-byte aligned");
  int32_t* wsi = reinterpret_cast<int32_t*>(workspace);
  VQ_CUDA(cudaMemsetAsync(wsi, 0, 256, stream));
  Params p;
  p.z = z;
  const unsigned char* img = reinterpret_cast<const unsigned char*>(image);
  p.tiles = img + img_f16_offset(K, D);
  p.meta = reinterpret_cast<const float*>(img + img_f16_meta_offset(K, D));
  p.info = info;
  p.K = K;
  p.NT = (int)(img_kp(K) / IMG_TILE_CODES);
  p.idx = idx;
  p.prev_idx = prev_idx; p.prev_E = prev_E; p.prev_K = prev_K; p.r_out = r_out;
  p.list = wsi + 64;
  p.cand2 = wsi + 64 + f16_n2(z.N);
  p.cand3 = wsi + 64 + 2 * f16_n2(z.N);
  p.rr_list = nullptr;
  p.rr_count = wsi + 5;
  p.wide = reinterpret_cast<int2*>(wsi + 64 + 4 * f16_n2(z.N));
  p.wide_cap = (int)f16_wide_cap(z.N);
  p.wide_count = wsi + 4;
  p.wide_buckets = 0;
  p.list_count = wsi;
  p.err = wsi + 1;
  p.stat = wsi + 2;
  static const int tc_debug = [] { const char* d = getenv("VQB200_TC_DEBUG"); return d ? atoi(d) : 0; }();
  p.dbg = tc_debug;            // development knobs used for the measurements in DESIGN.md (0 in production)
  // how the raw fp32 rows reach shared memory: one bulk-TMA copy per row tile when the rows of a
  // tile form one contiguous, 16-byte aligned byte range that fits the staging buffer
  const bool resident = p.NT <= RES_MAX_NT && !(p.dbg & 256);
  const int Rcap = (resident && p.NT > 6) ? 120 : TILE_M;       // 128 KiB of codebook leave room for 3 x 30 KiB row buffers
  p.stage_mode = STG_DIRECT;
  p.R = Rcap;
  const bool aligned = (reinterpret_cast<uintptr_t>(z.p) & 15) == 0;
  if (aligned && z.mode == Z_ROW && ((z.T == 1 && z.sB == D) || (z.sT == D && z.sB == z.T * D))) {
    p.stage_mode = STG_ROWS;
  } else if (aligned && z.mode == Z_BCT && z.T <= Rcap) {
    p.stage_mode = STG_BCT;                                     // whole samples per row tile: R = floor(cap/T)*T
    p.R = (int)((Rcap / z.T) * z.T);
  }
  p.nbuf = 0; p.buf_bytes = 0;
  if (resident) {
    p.buf_bytes = ((p.R * D * 4 + 1023) & ~1023);
    if (p.buf_bytes < A_BYTES + TILE_M * 8) p.buf_bytes = A_BYTES + 1024;
    const size_t fixed = (size_t)p.NT * F16_TILE_BYTES + (((size_t)p.NT * F16_META_BYTES + 15) & ~(size_t)15) + SMEM_BAR;
    p.nbuf = (int)min((size_t)RES_MAX_BUF, (SMEM_MAX - fixed) / p.buf_bytes);
    VQ_CHECK_ARG(p.nbuf >= 3, VQB200_EUNSUPPORTED, "vq_assign(TC): resident layout does not fit (NT=%d, R=%d)", p.NT, p.R);
    const size_t smem = fixed + (size_t)p.nbuf * p.buf_bytes;
    static PerDevice configured_;
    std::atomic<size_t>& configured = configured_.here();
    if (!configured.load()) {
      VQ_CUDA(cudaFuncSetAttribute(vq_assign_f16_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
      configured.store(1);
    }
    p.ntiles = (z.N + p.R - 1) / p.R;
    p.rr_list = wsi + 64 + 3 * f16_n2(z.N);          // the kernel finishes most rows itself and lists the rest
    if (z.N > F16_SPLIT_MAX_ROWS && !(p.dbg & 16384)) { p.wide_buckets = 1; p.wide_count = wsi + 8; }
    const int grid = (int)max(1LL, min((p.ntiles + 1) / 2, (long long)sm_count()));
    vq_assign_f16_res_kernel<<<grid, NTHREADS_RES, smem, stream>>>(p);
    VQ_LAUNCH_CHECK("vq_assign_f16_res_kernel");
  } else {
    static PerDevice configured_;
    std::atomic<size_t>& configured = configured_.here();
    if (!configured.load()) {
      VQ_CUDA(cudaFuncSetAttribute(vq_assign_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
      configured.store(1);
    }
    p.ntiles = (z.N + (long long)RT * p.R - 1) / ((long long)RT * p.R);
    const int grid = (int)max(1LL, min(p.ntiles, (long long)sm_count()));
    vq_assign_f16_kernel<<<grid, NTHREADS, SMEM_TOTAL, stream>>>(p);
    VQ_LAUNCH_CHECK("vq_assign_f16_kernel");
  }
  if (p.dbg & 8) return VQB200_OK;             // development: time the filter alone
  ZView zq = z; if (r_out) zq.p = r_out;       // the rows that were quantized are the NEW residual when fused
  // exact re-rank of the 4 / 8 / 12 candidates of every proven row that is not final yet
  {
    const float4* E4 = reinterpret_cast<const float4*>(img + img_e4_offset(K, D));
    int lblocks = 0;
    const bool with_wide = !(p.dbg & 16);
    if (p.rr_list) {
      lblocks = (int)max(1LL, min((z.N + 63) / 64, (long long)sm_count() * 8));
    } else {
      // contiguous layouts with T <= 64: tiles of whole samples staged in shared memory
      const bool tiled = p.stage_mode != STG_DIRECT && z.T <= RR_ROWS && (z.mode == Z_BCT || z.T == 1);
      const int rpt = tiled ? (int)((RR_ROWS / z.T) * z.T) : RR_ROWS;
      const long long rtiles = (z.N + rpt - 1) / rpt;
      const int rgrid = (int)max(1LL, min(rtiles, (long long)sm_count() * 8));
      if (tiled) vq_rerank_kernel<true><<<rgrid, RR_THREADS, 0, stream>>>(zq, rpt, E4, ee, K, idx, p.cand2, p.cand3);
      else vq_rerank_kernel<false><<<rgrid, RR_THREADS, 0, stream>>>(zq, rpt, E4, ee, K, idx, p.cand2, p.cand3);
      VQ_LAUNCH_CHECK("vq_rerank_kernel");
    }
    // launch-bound sizes: one launch for both kinds (interleaved CTAs, -18 us per cfg1 step); at bandwidth-bound sizes
    // two launches of full width are faster (3.79 vs 3.90 ms per 10 M x 1024 assignment on the same GPU)
    const bool merged = lblocks > 0 && with_wide && z.N <= F16_SPLIT_MAX_ROWS;
    const int wfull = (int)max(1LL, min((long long)(p.wide_cap + 7) / 8, (long long)sm_count() * 8));
    if (merged) {
      const int half = (int)max(1LL, min((z.N + 63) / 64, (long long)sm_count() * 4));
      vq_rerank_finish_kernel<<<2 * half, 256, 0, stream>>>(zq, E4, ee, K, idx, p.cand2, p.cand3, p.rr_list, p.rr_count,
                                                            half, p.wide, p.wide_count, p.wide_cap);
      VQ_LAUNCH_CHECK("vq_rerank_finish_kernel");
    } else {
      if (lblocks > 0) {
        vq_rerank_finish_kernel<<<lblocks, 256, 0, stream>>>(zq, E4, ee, K, idx, p.cand2, p.cand3, p.rr_list, p.rr_count,
                                                             lblocks, p.wide, p.wide_count, p.wide_cap);
        VQ_LAUNCH_CHECK("vq_rerank_finish_kernel(list)");
      }
      if (with_wide && p.wide_buckets) {
        const int per_tile = max(1, (sm_count() * 6) / p.NT);
        vq_rerank_wide_tile_kernel<<<p.NT * per_tile, 256, 0, stream>>>(zq, E4, ee, K, p.NT, idx, p.cand2, p.cand3,
                                                                         reinterpret_cast<const int32_t*>(p.wide), p.wide_count, p.wide_cap);
        VQ_LAUNCH_CHECK("vq_rerank_wide_tile_kernel");
      } else if (with_wide) {
        vq_rerank_finish_kernel<<<wfull, 256, 0, stream>>>(zq, E4, ee, K, idx, p.cand2, p.cand3, p.rr_list, p.rr_count,
                                                           0, p.wide, p.wide_count, p.wide_cap);
        VQ_LAUNCH_CHECK("vq_rerank_finish_kernel(wide)");
      }
    }
  }
  // exact re-do of the rows the filter could not prove (count lives on the device; no host sync)
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + f16_keys_offset(z.N));
  if (r_out) VQ_CHECK_ARG(p.stage_mode != STG_DIRECT, VQB200_EUNSUPPORTED, "vq_assign(TC): fused residual needs a contiguous layout");
  if (p.dbg & 32) return VQB200_OK;
  return launch_assign_simt_capped(zq, E, ee, K, D, idx, best, p.list, p.list_count, z.N, stream, keys,
                                   min(z.N, F16_SPLIT_MAX_ROWS));
}

}  // namespace vqb200
