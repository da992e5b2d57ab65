// vqb200 -- codebook-derived state shared by codebook_prepare and ema_finalize.
//
// The tcgen05 assignment kernel consumes the codebook as a bf16 "tile image".  Every fp32 entry is
// split into two bf16 terms E = E_hi + E_lo (+ 2^-18 relative remainder); for every block of 128
// codes (nt) and every block of 64 dims (kb) the image holds one 16 KiB E_hi tile followed by one
// 16 KiB E_lo tile, K-major, laid out exactly as UMMA's SWIZZLE_128B shared-memory layout expects
// (row r = 128 bytes, 16-byte chunk j stored at chunk position j ^ (r & 7)), so that ONE 1-D bulk-TMA
// copy lands ready-to-multiply B operands.  Tiles are ordered nt-major, kb-minor.  After the last
// tile come Kp = roundup(K,128) fp32 values  -|E_k|^2 / 2  (-inf for the padding codes): the kernel's
// score is  s_k = x.E_k - |E_k|^2/2  (arg max s == arg min distance).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace vqb200 {

constexpr int IMG_TILE_CODES = 128;
constexpr int IMG_TILE_DIMS = 64;
constexpr int IMG_HALF_BYTES = IMG_TILE_CODES * IMG_TILE_DIMS * 2;   // 16384: one of {hi, lo}
constexpr int IMG_TILE_BYTES = 2 * IMG_HALF_BYTES;                   // 32768: hi tile + lo tile

__host__ __device__ inline long long img_kp(long long K) { return (K + IMG_TILE_CODES - 1) / IMG_TILE_CODES * IMG_TILE_CODES; }
__host__ __device__ inline long long img_dp(long long D) { return (D + IMG_TILE_DIMS - 1) / IMG_TILE_DIMS * IMG_TILE_DIMS; }
__host__ __device__ inline size_t img_tiles_bytes(long long K, long long D) {
  return (size_t)(img_kp(K) / IMG_TILE_CODES) * (size_t)(img_dp(D) / IMG_TILE_DIMS) * IMG_TILE_BYTES;
}
__host__ __device__ inline size_t img_total_bytes(long long K, long long D) {
  return img_tiles_bytes(K, D) + (size_t)img_kp(K) * sizeof(float);
}

// byte offset of the hi term of element (code k, dim c) inside the image; the lo term sits
// IMG_HALF_BYTES further.
__device__ __forceinline__ size_t img_elem_offset(int k, int c, int Dp) {
  const int nt = k / IMG_TILE_CODES, r = k % IMG_TILE_CODES;
  const int kb = c / IMG_TILE_DIMS, cc = c % IMG_TILE_DIMS;
  const int chunk = cc >> 3, within = cc & 7;
  const size_t tile = (size_t)nt * (Dp / IMG_TILE_DIMS) + kb;
  return tile * IMG_TILE_BYTES + (size_t)r * 128 + (size_t)((chunk ^ (r & 7)) << 4) + within * 2;
}

__device__ __forceinline__ void img_store(unsigned char* image, int k, int c, int Dp, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  unsigned char* p = image + img_elem_offset(k, c, Dp);
  *reinterpret_cast<__nv_bfloat16*>(p) = hi;
  *reinterpret_cast<__nv_bfloat16*>(p + IMG_HALF_BYTES) = lo;
}

// info[0] = max_k |E_k| (stored through an int atomicMax, valid for non-negative floats),
// info[1] = 1 if any codebook entry is non-finite.
__device__ __forceinline__ void info_update(float* info, float ee_k, bool nonfinite) {
  if (!info) return;
  if (nonfinite || !(ee_k == ee_k) || isinf(ee_k)) { atomicExch(reinterpret_cast<int*>(info + 1), __float_as_int(1.0f)); return; }
  atomicMax(reinterpret_cast<int*>(info), __float_as_int(sqrtf(ee_k)));
}

}  // namespace vqb200
