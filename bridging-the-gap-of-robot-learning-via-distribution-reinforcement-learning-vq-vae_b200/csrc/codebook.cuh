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
__host__ __device__ inline size_t img_split_bytes(long long K, long long D) {     // bf16 split tiles + -|E|^2/2
  return img_tiles_bytes(K, D) + (size_t)img_kp(K) * sizeof(float);
}

// ---- fp16 filter image (D == 64 only; consumed by assign_f16.cu) ----------------------------------
// Appended to the split image at the next 1024-byte boundary: for every block of 128 codes one 16 KiB
// tile of fp16(E * s_j) in the same SWIZZLE_128B K-major layout (s_j = a power of two chosen per tile so
// that the tile's largest |element| lands in [2^10, 2^11): exact scaling, full fp16 relative precision for
// elements down to 2^-24 of the tile maximum), followed by one 528-byte meta record per tile:
// 128 floats -|E_k|^2/2 (-inf for padding codes) and {1/s_j, 0, 0, 0}.
constexpr int F16_TILE_BYTES = IMG_TILE_CODES * IMG_TILE_DIMS * 2;   // 16384
constexpr int F16_META_FLOATS = IMG_TILE_CODES + 4;                  // 132
constexpr int F16_META_BYTES = F16_META_FLOATS * 4;                  // 528
__host__ __device__ inline bool img_has_f16(long long D) { return D == IMG_TILE_DIMS; }
__host__ __device__ inline size_t img_f16_offset(long long K, long long D) { return (img_split_bytes(K, D) + 1023) & ~(size_t)1023; }
__host__ __device__ inline size_t img_f16_meta_offset(long long K, long long D) {
  return img_f16_offset(K, D) + (size_t)(img_kp(K) / IMG_TILE_CODES) * F16_TILE_BYTES;
}
// After the meta records: the fp32 codebook again, interleaved per group of 4 consecutive codes for the exact re-rank
// (group g = 64 float4: float4 4*q + c holds dims 4q..4q+3 of code 4g + c), so that the 4 lanes that evaluate a
// candidate group read 64 contiguous bytes per load instead of 4 separate cache lines.
__host__ __device__ inline size_t img_e4_offset(long long K, long long D) {
  return (img_f16_meta_offset(K, D) + (size_t)(img_kp(K) / IMG_TILE_CODES) * F16_META_BYTES + 1023) & ~(size_t)1023;
}
__host__ __device__ inline size_t img_total_bytes(long long K, long long D) {
  if (!img_has_f16(D)) return img_split_bytes(K, D);
  return img_e4_offset(K, D) + (size_t)img_kp(K) * IMG_TILE_DIMS * sizeof(float);
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
// info[1] = 1 if any codebook entry is non-finite (or too small for the fp16 image),
// info[2] = min_k |E_k| (int atomicMin; reset to +inf; written by the fp16 image kernel).
constexpr float INFO2_RESET = __builtin_huge_valf();
__device__ __forceinline__ void info_update(float* info, float ee_k, bool nonfinite) {
  if (!info) return;
  if (nonfinite || !(ee_k == ee_k) || isinf(ee_k)) { atomicExch(reinterpret_cast<int*>(info + 1), __float_as_int(1.0f)); return; }
  atomicMax(reinterpret_cast<int*>(info), __float_as_int(sqrtf(ee_k)));
}

// builds the fp16 filter image + info[2] from the refreshed E and split image (ema.cu); no-op unless D == 64 and image != NULL
int launch_image_f16(const float* E, long long K, long long D, void* image, float* info, cudaStream_t stream);
int preload_image_f16();

}  // namespace vqb200
