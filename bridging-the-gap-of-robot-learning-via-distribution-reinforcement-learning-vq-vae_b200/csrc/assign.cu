// vqb200 K1 dispatch: picks the tcgen05 filter+rerank kernel (assign_tc.cu) when the shape is
// eligible and falls back to the exact CUDA-core kernel (assign_simt.cu) otherwise.
#include "common.cuh"
#include "codebook.cuh"

namespace vqb200 {
int launch_assign_simt(const ZView& z, const float* E, const float* ee, int K, int D,
                       int32_t* idx, float* best, const int32_t* row_list, const int32_t* row_count,
                       long long max_rows, cudaStream_t stream, unsigned long long* keys = nullptr);
bool assign_tc_eligible(const ZView& z, int K, int D);
int launch_assign_tc(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                     int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream, const int32_t* prev_idx = nullptr, const float* prev_E = nullptr,
                     int prev_K = 0, float* r_out = nullptr);
bool assign_tc_can_fuse_residual(const ZView& z, const float* r_out);
size_t assign_tc_workspace_bytes(long long N);
bool assign_tc_gen_eligible(const ZView& z, int K, int D);
int launch_assign_tc_gen(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                         int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
size_t assign_tc_gen_workspace_bytes(long long N, int D);
bool assign_f16_eligible(const ZView& z, int K, int D);
int launch_assign_f16(const ZView& z, const float* E, const float* ee, const void* image, const float* info,
                      int K, int D, int32_t* idx, float* best, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream, const int32_t* prev_idx = nullptr, const float* prev_E = nullptr,
                      int prev_K = 0, float* r_out = nullptr);
bool assign_f16_can_fuse_residual(const ZView& z, const float* r_out);
size_t assign_f16_workspace_bytes(long long N);
}  // namespace vqb200

using namespace vqb200;

extern "C" {

size_t vqb200_assign_workspace_bytes(int64_t N, int64_t D) {
  if (D == 128 || D == 256 || D == 512) return assign_tc_gen_workspace_bytes(N, (int)D);
  const size_t a = assign_tc_workspace_bytes(N), b = assign_f16_workspace_bytes(N);
  return a > b ? a : b;
}

int vqb200_vq_assign(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                     const float* E, const float* ee, const void* image, const float* info, int64_t K,
                     int32_t* idx, float* best, void* workspace, size_t workspace_bytes, int algo,
                     vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG((z && E && ee && idx) || B * T == 0, VQB200_EINVAL, "vq_assign: null pointer");
  VQ_CHECK_ARG(B >= 0 && C > 0 && T > 0 && K > 0 && K < (1LL << 30) && C < (1 << 16), VQB200_ESHAPE,
               "vq_assign: bad shape B=%lld C=%lld T=%lld K=%lld", (long long)B, (long long)C, (long long)T, (long long)K);
  VQ_CHECK_ARG(B * T < (1LL << 31), VQB200_ESHAPE, "vq_assign: N=%lld exceeds int32 row ids", (long long)(B * T));
  if (B * T == 0) return VQB200_OK;
  const ZView zv = make_zview(z, B, C, T, sB, sC, sT);
  const int D = (int)C;
  bool use_tc = false;
  const bool gen = assign_tc_gen_eligible(zv, (int)K, D);          // D = 128 / 256: bf16 row-image variant
  const bool eligible = gen || assign_tc_eligible(zv, (int)K, D);  // D = 64: in-place conversion variant
  const size_t need = vqb200_assign_workspace_bytes(zv.N, D);
  const bool split = (algo == VQB200_ASSIGN_TC_SPLIT);        // development: the round-1 split-bf16 kernel (D == 64)
  if (split) algo = VQB200_ASSIGN_TC;
  if (algo == VQB200_ASSIGN_TC) {
    VQ_CHECK_ARG(image && info && workspace, VQB200_EINVAL, "vq_assign(TC): image, info and workspace are required");
    VQ_CHECK_ARG(!best, VQB200_EUNSUPPORTED, "vq_assign(TC): the winning distance is only produced by the SIMT algorithm");
    VQ_CHECK_ARG(eligible, VQB200_EUNSUPPORTED, "vq_assign(TC): shape K=%lld D=%d not eligible (D must be 64, 128, 256 or 512)", (long long)K, D);
    use_tc = true;
  } else if (algo == VQB200_ASSIGN_AUTO) {
    use_tc = image && info && workspace && !best && eligible && workspace_bytes >= need && zv.N >= 2048;
  } else {
    VQ_CHECK_ARG(algo == VQB200_ASSIGN_SIMT, VQB200_EINVAL, "vq_assign: unknown algo %d", algo);
  }
  if (use_tc && gen) return launch_assign_tc_gen(zv, E, ee, image, info, (int)K, D, idx, best, workspace, workspace_bytes, stream);
  // measured on B200 (4 M rows): K = 512: 1.03 ms split-bf16 vs 1.29 ms fp16 filter (per-row resolve / re-rank extras
  // dominate a short codebook sweep); K = 1024: 1.77 vs 1.60; K = 4096: 6.5 vs 5.2
  if (use_tc && !split && algo == VQB200_ASSIGN_AUTO && K <= 512) return launch_assign_tc(zv, E, ee, image, info, (int)K, D, idx, best, workspace, workspace_bytes, stream);
  if (use_tc && split) return launch_assign_tc(zv, E, ee, image, info, (int)K, D, idx, best, workspace, workspace_bytes, stream);
  if (use_tc) return launch_assign_f16(zv, E, ee, image, info, (int)K, D, idx, best, workspace, workspace_bytes, stream);
  return launch_assign_simt(zv, E, ee, (int)K, D, idx, best, nullptr, nullptr, zv.N, stream);
}

// RVQ stage s >= 1 in one call (models/vqvae.py:94-98 then :30-38):
//   r_out = r_in - st,  st = r_in + (E_prev[idx_prev] - r_in)        (what vq_gather_st writes as `residual`)
//   idx   = argmin_k d(r_out, E_k)
// For the D == 64 tensor-core path with a contiguous layout the residual update runs inside the assignment kernel
// (the rows pass through shared memory anyway), which saves one full read of r_in; otherwise the two stand-alone
// kernels run back to back.  r_out is a contiguous [B,C,T] tensor and may alias r_in only when r_in is contiguous.
int vqb200_vq_assign_residual(const float* r_in, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                              const float* E_prev, const int32_t* idx_prev, int64_t K_prev, float* r_out,
                              const float* E, const float* ee, const void* image, const float* info, int64_t K,
                              int32_t* idx, void* workspace, size_t workspace_bytes, int algo,
                              vqb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_CHECK_ARG((r_in && E_prev && idx_prev && r_out && E && ee && idx) || B * T == 0, VQB200_EINVAL,
               "vq_assign_residual: null pointer");
  VQ_CHECK_ARG(B >= 0 && C > 0 && T > 0 && K > 0 && K_prev > 0 && K < (1LL << 30) && K_prev < (1LL << 30) && C < (1 << 16),
               VQB200_ESHAPE, "vq_assign_residual: bad shape B=%lld C=%lld T=%lld K=%lld", (long long)B, (long long)C,
               (long long)T, (long long)K);
  VQ_CHECK_ARG(B * T < (1LL << 31), VQB200_ESHAPE, "vq_assign_residual: N=%lld exceeds int32 row ids", (long long)(B * T));
  if (B * T == 0) return VQB200_OK;
  const ZView zv = make_zview(r_in, B, C, T, sB, sC, sT);
  const int D = (int)C;
  const bool split = (algo == VQB200_ASSIGN_TC_SPLIT);
  if (split) algo = VQB200_ASSIGN_TC;
  const bool tc_ok = (algo == VQB200_ASSIGN_AUTO || algo == VQB200_ASSIGN_TC) && image && info && workspace &&
                     assign_f16_eligible(zv, (int)K, D) && workspace_bytes >= vqb200_assign_workspace_bytes(zv.N, D) &&
                     (zv.N >= 2048 || algo == VQB200_ASSIGN_TC) && assign_f16_can_fuse_residual(zv, r_out);
  if (tc_ok && (split || (algo == VQB200_ASSIGN_AUTO && K <= 512)))
    return launch_assign_tc(zv, E, ee, image, info, (int)K, D, idx, nullptr, workspace, workspace_bytes, stream,
                            idx_prev, E_prev, (int)K_prev, r_out);
  if (tc_ok)
    return launch_assign_f16(zv, E, ee, image, info, (int)K, D, idx, nullptr, workspace, workspace_bytes, stream,
                            idx_prev, E_prev, (int)K_prev, r_out);
  const int rc = vqb200_vq_gather_st(r_in, B, C, T, sB, sC, sT, E_prev, idx_prev, K_prev, nullptr, r_out, nullptr, 0,
                                     nullptr, stream_);
  if (rc != VQB200_OK) return rc;
  return vqb200_vq_assign(r_out, B, C, T, C * T, T, 1, E, ee, image, info, K, idx, nullptr, workspace, workspace_bytes,
                          algo, stream_);
}

}  // extern "C"
