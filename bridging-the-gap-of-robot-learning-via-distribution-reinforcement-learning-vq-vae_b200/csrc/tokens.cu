// vqb200 token export / decode-only path.
//
// The reference never materialises tokens: scripts/deployment/export_motion.py:25-83 runs encoder -> quantizer ->
// decoder per sliding window with B = 1.  Here the quantizer's device-side results become a compact token stream and
// can be turned back into the quantized latent without the encoder:
//   pack    per vector n: S codebook indices (code_bits each, the RVQ stages / the LFQ bit pattern) followed by d
//           signed FSQ digits (digit_bits each, two's complement; digit = round(z_e), saturated -- FSQ rounding is
//           unbounded in the reference, models/vqvae.py:127-131, so the mixed-radix index of :135 is NOT invertible
//           and the digits themselves are stored), little-endian bit stream, token size rounded up to whole bytes
//   unpack  the inverse
//   decode  quantized[b,c,t] = (W_out digits + b_out)[c] + ((0 + E_0[i_0][c]) + E_1[i_1][c]) + ...
//           = FSQ.project_out (:133) + the ResidualVQ sum (:94-98) + HybridVQ's z_fsq + z_vq (:229), written as a
//           contiguous [B,C,T] tensor ready for the decoder
#include "common.cuh"

namespace vqb200 {

constexpr int TOK_MAX_BITS = 256;          // per token
constexpr int TOK_MAX_S = 8;
constexpr int TOK_MAX_D = 16;

struct TokLayout { int S, code_bits, d, digit_bits, bytes; };

__host__ __device__ inline int tok_bits(const TokLayout& L) { return L.S * L.code_bits + L.d * L.digit_bits; }

struct BitWriter {
  unsigned long long w[TOK_MAX_BITS / 64];
  int pos;
  __device__ BitWriter() : pos(0) {
#pragma unroll
    for (int i = 0; i < TOK_MAX_BITS / 64; ++i) w[i] = 0ull;
  }
  __device__ void put(unsigned long long v, int bits) {          // bits <= 32
    const int word = pos >> 6, sh = pos & 63;
#pragma unroll
    for (int i = 0; i < TOK_MAX_BITS / 64; ++i) {
      if (i == word) w[i] |= v << sh;
      if (i == word + 1 && sh + bits > 64) w[i] |= v >> (64 - sh);
    }
    pos += bits;
  }
};

struct BitReader {
  unsigned long long w[TOK_MAX_BITS / 64];
  int pos;
  __device__ unsigned long long get(int bits) {                  // bits <= 32
    const int word = pos >> 6, sh = pos & 63;
    unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll
    for (int i = 0; i < TOK_MAX_BITS / 64; ++i) {
      if (i == word) lo = w[i];
      if (i == word + 1) hi = w[i];
    }
    unsigned long long v = lo >> sh;
    if (sh + bits > 64) v |= hi << (64 - sh);
    pos += bits;
    return v & ((1ull << bits) - 1ull);
  }
};

__global__ void __launch_bounds__(256)
tokens_pack_kernel(const int32_t* __restrict__ codes, const float* __restrict__ z_e, long long B, int T, TokLayout L,
                   unsigned char* __restrict__ out, int* __restrict__ overflow) {
  const long long N = B * T;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    BitWriter bw;
    bool ovf = false;
    for (int s = 0; s < L.S; ++s) {
      const long long v = codes[(long long)s * N + n];
      if (v < 0 || (L.code_bits < 32 && v >= (1LL << L.code_bits))) ovf = true;
      bw.put((unsigned long long)v & ((1ull << L.code_bits) - 1ull), L.code_bits);
    }
    const long long b = n / T; const int t = (int)(n - b * T);
    const long long lim = 1LL << (L.digit_bits - 1);
    for (int j = 0; j < L.d; ++j) {
      const float zh = rintf(z_e[(b * L.d + j) * T + t]);           // the value of z + (round(z) - z), :130-131
      long long q = (zh >= 9.0e18f || zh <= -9.0e18f || zh != zh) ? (ovf = true, 0LL) : (long long)zh;
      if (q >= lim) { q = lim - 1; ovf = true; }
      if (q < -lim) { q = -lim; ovf = true; }
      bw.put((unsigned long long)q & ((1ull << L.digit_bits) - 1ull), L.digit_bits);
    }
    unsigned char* dst = out + n * L.bytes;
    for (int i = 0; i < L.bytes; ++i) dst[i] = (unsigned char)(bw.w[i >> 3] >> ((i & 7) * 8));
    if (ovf) atomicExch(overflow, 1);
  }
}

__global__ void __launch_bounds__(256)
tokens_unpack_kernel(const unsigned char* __restrict__ in, long long B, int T, TokLayout L,
                     int32_t* __restrict__ codes, float* __restrict__ digits) {
  const long long N = B * T;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    BitReader br;
    br.pos = 0;
#pragma unroll
    for (int i = 0; i < TOK_MAX_BITS / 64; ++i) br.w[i] = 0ull;
    const unsigned char* src = in + n * L.bytes;
    for (int i = 0; i < L.bytes; ++i) {
      const unsigned long long v = (unsigned long long)src[i] << ((i & 7) * 8);
#pragma unroll
      for (int k = 0; k < TOK_MAX_BITS / 64; ++k) if (k == (i >> 3)) br.w[k] |= v;
    }
    for (int s = 0; s < L.S; ++s) codes[(long long)s * N + n] = (int32_t)br.get(L.code_bits);
    const long long b = n / T; const int t = (int)(n - b * T);
    for (int j = 0; j < L.d; ++j) {
      long long q = (long long)br.get(L.digit_bits);
      if (q & (1LL << (L.digit_bits - 1))) q -= (1LL << L.digit_bits);   // sign-extend
      digits[(b * L.d + j) * T + t] = (float)q;
    }
  }
}

struct DecodeParams {
  const int32_t* codes;                 // [S,N] or null
  const float* E[TOK_MAX_S]; int K[TOK_MAX_S];
  int S;
  const float* digits;                  // [B,d,T] or null
  const float* W_out; const float* b_out; int d;
  long long B; int C, T;
  float* out;
};

// one thread per output element, memory order of the contiguous [B,C,T] result
__global__ void __launch_bounds__(256)
tokens_decode_kernel(const DecodeParams p) {
  const long long CT = (long long)p.C * p.T, total = p.B * CT, N = p.B * p.T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / CT;
    const int rem = (int)(i - b * CT);
    const int c = rem / p.T, t = rem - c * p.T;
    const long long n = b * p.T + t;
    float vq = 0.f;                                                // ((0 + q_0) + q_1) + ...   :94-98
    for (int s = 0; s < p.S; ++s) {
      int k = __ldg(p.codes + (long long)s * N + n);
      k = min(max(k, 0), p.K[s] - 1);
      vq = __fadd_rn(vq, __ldg(p.E[s] + (size_t)k * p.C + c));
    }
    float o = vq;
    if (p.digits) {
      float f = __ldg(p.b_out + c);                                // project_out(z_hard)  :133
      for (int j = 0; j < p.d; ++j) f = fmaf(__ldg(p.W_out + c * p.d + j), __ldg(p.digits + (b * p.d + j) * p.T + t), f);
      o = p.S > 0 ? __fadd_rn(f, vq) : f;                          // z_fsq + z_vq  :229
    }
    p.out[i] = o;
  }
}

static int tok_layout(int64_t S, int64_t code_bits, int64_t d, int64_t digit_bits, TokLayout& L) {
  VQ_CHECK_ARG(S >= 0 && S <= TOK_MAX_S && d >= 0 && d <= TOK_MAX_D && S + d > 0, VQB200_ESHAPE,
               "tokens: unsupported field counts S=%lld d=%lld", (long long)S, (long long)d);
  VQ_CHECK_ARG((S == 0 || (code_bits >= 1 && code_bits <= 32)) && (d == 0 || (digit_bits >= 2 && digit_bits <= 32)),
               VQB200_ESHAPE, "tokens: field widths must be 1..32 (codes) / 2..32 (digits) bits");
  L.S = (int)S; L.code_bits = (int)code_bits; L.d = (int)d; L.digit_bits = (int)digit_bits;
  const int bits = tok_bits(L);
  VQ_CHECK_ARG(bits <= TOK_MAX_BITS, VQB200_ESHAPE, "tokens: %d bits per token exceed the maximum of %d", bits, TOK_MAX_BITS);
  L.bytes = (bits + 7) / 8;
  return VQB200_OK;
}

}  // namespace vqb200

using namespace vqb200;

extern "C" {

int64_t vqb200_token_bytes(int64_t S, int64_t code_bits, int64_t d, int64_t digit_bits) {
  TokLayout L;
  const int rc = tok_layout(S, code_bits, d, digit_bits, L);
  return rc == VQB200_OK ? L.bytes : rc;
}

int vqb200_tokens_pack(const int32_t* codes, int64_t S, int64_t code_bits, const float* z_e, int64_t d,
                       int64_t digit_bits, int64_t B, int64_t T, uint8_t* tokens, int32_t* overflow,
                       vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TokLayout L;
  const int rc = tok_layout(S, code_bits, d, digit_bits, L);
  if (rc != VQB200_OK) return rc;
  VQ_CHECK_ARG(B >= 0 && T > 0 && overflow, VQB200_EINVAL, "tokens_pack: bad arguments");
  VQ_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int32_t), stream));
  if (B == 0) return VQB200_OK;
  VQ_CHECK_ARG(tokens && (S == 0 || codes) && (d == 0 || z_e), VQB200_EINVAL, "tokens_pack: null pointer");
  const long long N = B * T;
  const int grid = (int)max(1LL, min((N + 255) / 256, (long long)sm_count() * 16));
  tokens_pack_kernel<<<grid, 256, 0, stream>>>(codes, z_e, B, (int)T, L, tokens, overflow);
  VQ_LAUNCH_CHECK("tokens_pack_kernel");
  return VQB200_OK;
}

int vqb200_tokens_unpack(const uint8_t* tokens, int64_t S, int64_t code_bits, int64_t d, int64_t digit_bits,
                         int64_t B, int64_t T, int32_t* codes, float* digits, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TokLayout L;
  const int rc = tok_layout(S, code_bits, d, digit_bits, L);
  if (rc != VQB200_OK) return rc;
  VQ_CHECK_ARG(B >= 0 && T > 0, VQB200_EINVAL, "tokens_unpack: bad arguments");
  if (B == 0) return VQB200_OK;
  VQ_CHECK_ARG(tokens && (S == 0 || codes) && (d == 0 || digits), VQB200_EINVAL, "tokens_unpack: null pointer");
  const long long N = B * T;
  const int grid = (int)max(1LL, min((N + 255) / 256, (long long)sm_count() * 16));
  tokens_unpack_kernel<<<grid, 256, 0, stream>>>(tokens, B, (int)T, L, codes, digits);
  VQ_LAUNCH_CHECK("tokens_unpack_kernel");
  return VQB200_OK;
}

int vqb200_tokens_decode(const int32_t* codes, int64_t S, const float* const* E, const int64_t* K,
                         const float* digits, int64_t d, const float* W_out, const float* b_out,
                         int64_t B, int64_t C, int64_t T, float* out, vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG(S >= 0 && S <= TOK_MAX_S && d >= 0 && d <= TOK_MAX_D && S + d > 0, VQB200_ESHAPE,
               "tokens_decode: unsupported field counts S=%lld d=%lld", (long long)S, (long long)d);
  VQ_CHECK_ARG(B >= 0 && C > 0 && T > 0, VQB200_ESHAPE, "tokens_decode: bad shape");
  if (B == 0) return VQB200_OK;
  VQ_CHECK_ARG(out && (S == 0 || (codes && E && K)) && (d == 0 || (digits && W_out && b_out)), VQB200_EINVAL,
               "tokens_decode: null pointer");
  DecodeParams p = {};
  p.codes = codes; p.S = (int)S;
  for (int s = 0; s < (int)S; ++s) {
    VQ_CHECK_ARG(E[s] && K[s] > 0, VQB200_EINVAL, "tokens_decode: null codebook %d", s);
    p.E[s] = E[s]; p.K[s] = (int)K[s];
  }
  p.digits = d > 0 ? digits : nullptr; p.W_out = W_out; p.b_out = b_out; p.d = (int)d;
  p.B = B; p.C = (int)C; p.T = (int)T; p.out = out;
  const long long total = B * C * T;
  const int grid = (int)max(1LL, min((total + 255) / 256, (long long)sm_count() * 16));
  tokens_decode_kernel<<<grid, 256, 0, stream>>>(p);
  VQ_LAUNCH_CHECK("tokens_decode_kernel");
  return VQB200_OK;
}

}  // extern "C"
