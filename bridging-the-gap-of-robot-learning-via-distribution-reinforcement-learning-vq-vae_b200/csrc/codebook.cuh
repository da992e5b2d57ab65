// vqb200 -- codebook-derived state shared by codebook_prepare and ema_finalize.
//
// The tcgen05 assignment kernel consumes the codebook as a bf16 "tile image": for every block of
// 256 codes (nt) and every block of 64 dims (kb) one 32 KiB tile, K-major, laid out exactly as
// UMMA's SWIZZLE_128B shared-memory layout expects (row r = 128 bytes, 16-byte chunk j stored at
// chunk position j ^ (r & 7)), so that a single 1-D bulk-TMA copy lands a ready-to-multiply B
// operand.  Tiles are ordered nt-major, kb-minor; after the last tile come Kp = roundup(K,256)
// fp32 |E_k|^2 values (+inf for the padding codes).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace vqb200 {

constexpr int IMG_TILE_CODES = 256;
constexpr int IMG_TILE_DIMS = 64;
constexpr int IMG_TILE_BYTES = IMG_TILE_CODES * IMG_TILE_DIMS * 2;   // 32768

__host__ __device__ inline long long img_kp(long long K) { return (K + IMG_TILE_CODES - 1) / IMG_TILE_CODES * IMG_TILE_CODES; }
__host__ __device__ inline long long img_dp(long long D) { return (D + IMG_TILE_DIMS - 1) / IMG_TILE_DIMS * IMG_TILE_DIMS; }
__host__ __device__ inline size_t img_tiles_bytes(long long K, long long D) {
  return (size_t)(img_kp(K) / IMG_TILE_CODES) * (size_t)(img_dp(D) / IMG_TILE_DIMS) * IMG_TILE_BYTES;
}
__host__ __device__ inline size_t img_total_bytes(long long K, long long D) {
  return img_tiles_bytes(K, D) + (size_t)img_kp(K) * sizeof(float);
}

// byte offset of element (code k, dim c) inside the image
__device__ __forceinline__ size_t img_elem_offset(int k, int c, int Dp) {
  const int nt = k / IMG_TILE_CODES, r = k % IMG_TILE_CODES;
  const int kb = c / IMG_TILE_DIMS, cc = c % IMG_TILE_DIMS;
  const int chunk = cc >> 3, within = cc & 7;
  const size_t tile = (size_t)nt * (Dp / IMG_TILE_DIMS) + kb;
  return tile * IMG_TILE_BYTES + (size_t)r * 128 + (size_t)((chunk ^ (r & 7)) << 4) + within * 2;
}

// info[0] = max_k |E_k| (stored through an int atomicMax, valid for non-negative floats),
// info[1] = 1 if any codebook entry is non-finite.
__device__ __forceinline__ void info_update(float* info, float ee_k, bool nonfinite) {
  if (!info) return;
  if (nonfinite || !(ee_k == ee_k) || isinf(ee_k)) { atomicExch(reinterpret_cast<int*>(info + 1), __float_as_int(1.0f)); return; }
  atomicMax(reinterpret_cast<int*>(info), __float_as_int(sqrtf(ee_k)));
}

}  // namespace vqb200
