/*
 * vqb200 -- C ABI of the B200-native vector-quantization engine.
 *
 * Drop-in boundary for the quantizer layer of the reference VQ-VAE retargeter.  The reference
 * has no FFI of its own: its boundary is the Python nn.Module contract of
 *   models/vqvae.py:10-259   (VectorQuantizer / ResidualVQ / FSQ / LFQ / HybridVQ / IdentityVQ)
 * consumed at models/vqvae.py:540-560,588,605.  Each entry point below replaces the ATen ops the
 * cited reference lines issue; the Python glue (package `vqb200`, drop-in `models/vqvae.py`)
 * binds them with ctypes and re-creates the module contract on top.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller (torch);
 *     the library never allocates or frees persistent device memory;
 *   - z tensors are fp32 [B, C, T] addressed by ELEMENT strides (sB, sC, sT); vector n = b*T + t
 *     has component k at z[b*sB + k*sC + t*sT]  (covers the contiguous channel-major layout and
 *     the permuted T'=1 view the transformer encoder emits, models/vqvae.py:458-463);
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t); no host synchronisation,
 *     no host read-back: every call is CUDA-graph capturable;
 *   - return value: 0 = success, < 0 = VQB200_E* argument error, > 0 = cudaError_t;
 *     vqb200_last_error_string() describes the last failure on the calling thread;
 *   - there is no CPU fallback.
 */
#ifndef VQB200_H_
#define VQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB200_ABI_VERSION 1

#define VQB200_OK            0
#define VQB200_EINVAL       -1   /* null pointer / non-positive size                        */
#define VQB200_ESHAPE       -2   /* unsupported shape (e.g. C != D, D too large for smem)   */
#define VQB200_EALIGN       -3   /* pointer not aligned as documented                       */
#define VQB200_EUNSUPPORTED -4   /* configuration not built (e.g. tensor-core path, D%64)   */
#define VQB200_EWORKSPACE   -5   /* workspace too small                                     */

typedef void* vqb200_stream_t;   /* cudaStream_t */

/* assignment algorithm selector for vqb200_vq_assign */
#define VQB200_ASSIGN_AUTO  0    /* tcgen05 filter + exact rerank when eligible, else SIMT  */
#define VQB200_ASSIGN_SIMT  1    /* exact fp32 CUDA-core path                               */
#define VQB200_ASSIGN_TC    2    /* force the tcgen05 path (error if not eligible)          */
#define VQB200_ASSIGN_TC_SPLIT 3 /* development: round-1 three-pass split-bf16 kernel (D = 64) */

int         vqb200_abi_version(void);
const char* vqb200_last_error_string(void);
/* number of kernels this library has launched from the calling process (bench.py's gpu_launches) */
int64_t     vqb200_launch_count(void);

/* ---- codebook-derived state ---------------------------------------------------------------
 * |E_k|^2 (replaces torch.sum(weight**2, dim=1), models/vqvae.py:35) and the bf16 tile image the
 * tcgen05 assignment kernel streams with bulk-TMA.  `image` may be NULL (SIMT only).
 * image size in bytes: vqb200_codebook_image_bytes(K, D).  info[0..3] (device, 4 floats):
 * {max_k |E_k|, nonfinite flag, reserved, reserved}. */
size_t vqb200_codebook_image_bytes(int64_t K, int64_t D);
int vqb200_codebook_prepare(const float* E, int64_t K, int64_t D,
                            float* ee, void* image, float* info, vqb200_stream_t stream);

/* ---- K1: fused distance + argmin ------------------- models/vqvae.py:30-38 (rows a2-a4) ----
 * idx[n] = first index of min_k fl(fl(|x_n|^2 + |E_k|^2) - 2 x_n.E_k); NaN distance wins.
 * Never materialises the N x K matrix.  idx: int32 [B*T].  best (optional, may be NULL): the
 * winning fp32 distance per row (SIMT algorithm only).  workspace: vqb200_assign_workspace_bytes(N, D)
 * bytes (may be NULL for the SIMT algorithm).  The tcgen05 path covers D = 64 (raw rows converted in shared
 * memory) and D = 128 / 256 (input first split into a bf16 row image inside the workspace). */
size_t vqb200_assign_workspace_bytes(int64_t N, int64_t D);
int vqb200_vq_assign(const float* z, int64_t B, int64_t C, int64_t T,
                     int64_t sB, int64_t sC, int64_t sT,
                     const float* E, const float* ee, const void* image, const float* info,
                     int64_t K, int32_t* idx, float* best,
                     void* workspace, size_t workspace_bytes, int algo, vqb200_stream_t stream);

/* ---- K3a: EMA statistics ------------------------------ models/vqvae.py:44-45 (row a5) -----
 * stats = [dw (K*D) | cnt (K)] fp32, zeroed by this call then filled:
 *   cnt[k] = #{n: idx_n = k}  (== encodings.sum(0)),  dw[k,:] = sum_{idx_n=k} x_n (== one_hot^T @ x).
 * mode 0: as above.  mode 1 (standard-VQ codebook gradient, row a11): dw[k,:] = sum (E[k,:] - x_n).
 * This buffer is what data-parallel ranks all-reduce. */
int vqb200_ema_accumulate(const float* z, int64_t B, int64_t C, int64_t T,
                          int64_t sB, int64_t sC, int64_t sT,
                          const int32_t* idx, const float* E, int64_t K,
                          float* stats, int mode, vqb200_stream_t stream);

/* ---- K3b: EMA finalize -------------------------------- models/vqvae.py:46-50 (row a5) -----
 * cs <- decay*cs + (1-decay)*cnt; w <- decay*w + (1-decay)*dw; n = sum(cs);
 * E <- w / ((cs+eps)/(n+K*eps)*n)  in place; refreshes ee / image / info like codebook_prepare.
 * decay / eps are the Python doubles of the reference (0.99, 1e-5); the kernels use
 * (float)decay, (float)(1-decay), (float)eps and (float)(K*eps) exactly as torch does when it
 * mixes Python scalars with fp32 tensors.  scratch: >= (K+8) floats. */
int vqb200_ema_finalize(const float* stats, float* ema_cluster_size, float* ema_w, float* E,
                        int64_t K, int64_t D, double decay, double eps,
                        float* ee, void* image, float* info, float* scratch,
                        vqb200_stream_t stream);

/* ---- K3b fused with its all-reduce over NVLink peer memory --------- SURVEY.md §8e ---------
 * Data-parallel ranks of ONE node (one process per GPU) replace
 *     ema_accumulate -> ncclAllReduce(stats) -> ema_finalize
 * by accumulating straight into a slot of a symmetric buffer every rank has mapped, then calling
 * vqb200_ema_finalize_peer: a one-shot flag barrier over peer memory, direct NVLink reads of every
 * rank's [dw | cnt] slot summed in RANK ORDER (bit-identical result on all ranks, no broadcast),
 * and the same decay / Laplace / normalise arithmetic as vqb200_ema_finalize (models/vqvae.py:46-50).
 *
 * vqb200_peer_alloc / _open / _close / _free are the only calls of this library that own device
 * memory: the buffer must come from cudaMalloc (not a framework's sub-allocator) to be exportable
 * with CUDA IPC.  `handle` is a 64-byte cudaIpcMemHandle_t the caller ships to its peers
 * (torch.distributed.all_gather_object in the glue).  Suggested layout (what the glue uses):
 *     [ flags: VQB200_MAX_PEERS x uint32, padded to 256 B | slot 0 | slot 1 ]
 * peer_stats[p] / peer_flags[p] (HOST arrays of `world` device pointers): rank p's slot of the
 * current epoch / rank p's flag words, as mapped into THIS process (own buffer for p == rank).
 * `epoch` increases by one per call on every rank (slot = epoch & 1); flags start at 0, so the
 * first epoch is 1.  cnt_out (K floats, may be NULL) receives the reduced counts for
 * vqb200_vq_metrics.  A peer that does not arrive within 120 s traps the kernel (loud, no hang). */
#define VQB200_MAX_PEERS 16
#define VQB200_PEER_HANDLE_BYTES 64
int vqb200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle);
int vqb200_peer_open(const unsigned char* handle, void** dev_ptr);
int vqb200_peer_close(void* dev_ptr);
int vqb200_peer_free(void* dev_ptr);
int vqb200_peer_barrier(uint32_t* const* peer_flags, int32_t rank, int32_t world, uint32_t epoch,
                        vqb200_stream_t stream);
int vqb200_ema_finalize_peer(const float* const* peer_stats, uint32_t* const* peer_flags,
                             int32_t rank, int32_t world, uint32_t epoch, float* cnt_out,
                             float* ema_cluster_size, float* ema_w, float* E, int64_t K, int64_t D,
                             double decay, double eps, float* ee, void* image, float* info,
                             float* scratch, vqb200_stream_t stream);

/* ---- code histogram only (eval / non-EMA) --------------- models/vqvae.py:66,71 (row a9) --- */
int vqb200_vq_histogram(const int32_t* idx, int64_t N, int64_t K, float* cnt, vqb200_stream_t stream);

/* ---- K2: gather + straight-through + loss partial ----- models/vqvae.py:52-63,76 (a6-a8,a10)
 * out[b,c,t] = x + (E[idx][c] - x)  written contiguous [B,C,T]; *sse += sum (E[idx][c]-x)^2
 * (fp64 accumulator, zeroed by this call).
 * residual (optional): residual[b,c,t] = x - out[b,c,t]  (ResidualVQ, models/vqvae.py:96).
 * accum (optional):    accum[b,c,t]  = (accum_init ? accum : 0) + out  (models/vqvae.py:97). */
int vqb200_vq_gather_st(const float* z, int64_t B, int64_t C, int64_t T,
                        int64_t sB, int64_t sC, int64_t sT,
                        const float* E, const int32_t* idx, int64_t K,
                        float* out, float* residual, float* accum, int accum_init,
                        double* sse, vqb200_stream_t stream);

/* ---- codebook health (opt-in, not in the reference) ---------------------- SURVEY.md §8f rank 4 ----
 * Every code k with usage[k] < threshold (usage = ema_cluster_size, or the histogram of the last step) is re-seeded
 * from input row n_k = splitmix64(seed + k * 0x9E3779B97F4A7C15) mod N:  E[k] = z[n_k]; if given, ema_w[k] = z[n_k] and
 * ema_cluster_size[k] = 1.  *revived = number of codes replaced.  The caller must refresh |E|^2 / the tile image
 * (vqb200_codebook_prepare) afterwards.  Never called by the drop-in modules unless the user asks. */
int vqb200_codebook_revive(const float* z, int64_t B, int64_t C, int64_t T, int64_t sB, int64_t sC, int64_t sT,
                           const float* usage, float threshold, uint64_t seed,
                           float* E, float* ema_cluster_size, float* ema_w, int64_t K,
                           int32_t* revived, vqb200_stream_t stream);

/* ---- K1 + residual update ------------------- models/vqvae.py:94-98 (r = r - q) then :30-38 --
 * One call per RVQ stage s >= 1:  r_out = r_in - st  with  st = r_in + (E_prev[idx_prev] - r_in)  (bit-identical
 * to the `residual` output of vqb200_vq_gather_st), then idx = argmin_k d(r_out, E_k) as vqb200_vq_assign.
 * For D == 64 on a contiguous [B,C,T] tensor the update runs inside the tensor-core assignment kernel (the rows
 * pass through shared memory anyway; saves one full read of r_in per stage); otherwise the two stand-alone
 * kernels run back to back.  r_out: contiguous [B,C,T]. */
int vqb200_vq_assign_residual(const float* r_in, int64_t B, int64_t C, int64_t T,
                              int64_t sB, int64_t sC, int64_t sT,
                              const float* E_prev, const int32_t* idx_prev, int64_t K_prev, float* r_out,
                              const float* E, const float* ee, const void* image, const float* info, int64_t K,
                              int32_t* idx, void* workspace, size_t workspace_bytes, int algo,
                              vqb200_stream_t stream);

/* ---- RVQ output chain ---------------------------------- models/vqvae.py:94-98, all stages at once --
 * Recomputes r_0 = z; st_s = r_s + (E_s[idx_s] - r_s); out = ((0 + st_0) + st_1) + ...; r_{s+1} = r_s - st_s
 * from z, the S index arrays and the S (already updated) codebooks -- bit-identical to running the stages one
 * after the other, because every codebook is updated once per step -- and the S loss sums
 * sse[s] = sum (E_s[idx_s] - r_s)^2.  E / idx / K are HOST arrays of S device pointers / sizes.
 * out: contiguous [B,C,T].  scratch (B*C*T floats) is only needed for non-contiguous views with S > 1. */
int vqb200_rvq_output_chain(const float* z, int64_t B, int64_t C, int64_t T,
                            int64_t sB, int64_t sC, int64_t sT,
                            int32_t S, const float* const* E, const int32_t* const* idx, const int64_t* K,
                            float* out, double* sse, float* scratch, vqb200_stream_t stream);

/* ---- K4: single-launch ResidualVQ for launch-bound shapes --- models/vqvae.py:87-108 (whole loop) ----
 * One thread-block cluster (16 CTAs) runs ALL stages: exact fp32 assignment, EMA statistics, EMA finalize,
 * gather / residual / running sum, loss + metrics, ordered by cluster barriers; the residual never leaves shared
 * memory.  Eligible (vqb200_rvq_small_eligible): D == 64, N = B*T <= 4096, K_s <= 4096, S <= 8; single process
 * only (no inter-GPU all-reduce inside the launch).  E / ema_cluster_size / ema_w / K are HOST arrays of S device
 * pointers / sizes; codebooks and EMA buffers are updated in place when training && use_ema.
 * workspace: vqb200_rvq_small_workspace_floats(S, K) floats (16-byte aligned); sse: S doubles;
 * idx: int32 [S, N]; out: contiguous [B,C,T]; m3: [S+1,3] = {loss, perplexity, dcr} per stage;
 * row S = {sum of the stage losses, mean perplexity, mean dcr} (what ResidualVQ.forward returns, :104-108).
 * Derived codebook state (|E|^2, tile image) is NOT refreshed: call vqb200_codebook_prepare before the next
 * vqb200_vq_assign on these codebooks. */
int    vqb200_rvq_small_eligible(int64_t N, int64_t D, int32_t S, const int64_t* K);
size_t vqb200_rvq_small_workspace_floats(int32_t S, const int64_t* K);
int vqb200_rvq_small_forward(const float* z, int64_t B, int64_t C, int64_t T,
                             int64_t sB, int64_t sC, int64_t sT,
                             int32_t S, float* const* E, float* const* ema_cluster_size, float* const* ema_w,
                             const int64_t* K, double decay, double eps, float commitment_cost, int use_ema,
                             int training, float* workspace, double* sse, int32_t* idx, float* out, float* m3,
                             vqb200_stream_t stream);

/* ---- K4 under data parallelism: the single-launch ResidualVQ with the exchange inside ---------
 * vqb200_rvq_small_forward for the EMA training step of data-parallel ranks of one node: the
 * whole-GPU (cooperative) kernel accumulates each stage's statistics straight into this rank's
 * peer slot (vqb200_peer_alloc above), and its per-stage grid barrier also spans the ranks --
 * CTA 0 publishes epoch0 + s into every peer's flag word over NVLink and waits for theirs -- after
 * which every CTA reads all ranks' [dw | cnt] over peer memory and sums them in RANK ORDER.  One
 * launch per step and rank, S NVLink barriers inside it, bit-identical codebooks on all ranks.
 * peer_stats[p]: rank p's slot of THIS call (vqb200_rvq_small_stats_floats(S, K) floats, the
 * same slot index on every rank; alternate two slots between consecutive calls); peer_flags as
 * for vqb200_ema_finalize_peer; the call consumes the epochs epoch0 .. epoch0 + S - 1;
 * n_total = vectors of all ranks (perplexity / dcr are global, the loss is this rank's).
 * Every rank must take this path for the same step: vqb200_rvq_small_peer_eligible(N, D, S, K)
 * (1 = the whole-GPU variant applies on the current device) must agree across ranks, i.e. shards
 * of equal size.  world == 1 degenerates to vqb200_rvq_small_forward. */
int vqb200_rvq_small_peer_eligible(int64_t N, int64_t D, int32_t S, const int64_t* K);
size_t vqb200_rvq_small_stats_floats(int32_t S, const int64_t* K);
int vqb200_rvq_small_forward_peer(const float* z, int64_t B, int64_t C, int64_t T,
                                  int64_t sB, int64_t sC, int64_t sT, int32_t S,
                                  float* const* E, float* const* ema_cluster_size, float* const* ema_w,
                                  const int64_t* K, double decay, double eps, float commitment_cost,
                                  float* workspace, double* sse, int32_t* idx, float* out, float* m3,
                                  const float* const* peer_stats, uint32_t* const* peer_flags,
                                  int32_t rank, int32_t world, uint32_t epoch0, int64_t n_total,
                                  vqb200_stream_t stream);

/* ---- loss + metrics as device scalars ----------------- models/vqvae.py:55-61,66-74 (a7,a9) -
 * out3 = {loss, perplexity, dcr}.  loss = c*mse (EMA) or mse + c*mse (standard), mse = sse/numel;
 * perplexity = exp(-sum p log(p+1e-10)), p = cnt/N; dcr = 1 - #{cnt>0}/K. */
int vqb200_vq_metrics(const float* cnt, int64_t K, int64_t N, const double* sse, int64_t numel,
                      float commitment_cost, int use_ema, float* out3, vqb200_stream_t stream);

/* ---- K2b: input gradient ------------------------------------------------- (row a11) --------
 * gz[b,c,t] = g[b,c,t] + g_loss[0]*coef*(x - E[idx][c]),  coef = c*2/(N*D).  g is addressed with
 * its own element strides; gz is written contiguous [B,C,T].  g may be NULL (treated as 0);
 * g_loss is a device scalar (NULL = 1). */
int vqb200_vq_backward_input(const float* g, int64_t gsB, int64_t gsC, int64_t gsT,
                             const float* z, int64_t B, int64_t C, int64_t T,
                             int64_t sB, int64_t sC, int64_t sT,
                             const float* E, const int32_t* idx, int64_t K,
                             const float* g_loss, float coef, float* gz, vqb200_stream_t stream);

/* ---- K3b': standard-VQ codebook gradient from mode-1 stats --------------- (row a11) --------
 * gE[k,:] = g_loss[0]*coef*dw1[k,:],  coef = 2/(N*D),  dw1 from vqb200_ema_accumulate(mode=1). */
int vqb200_vq_backward_codebook(const float* stats, int64_t K, int64_t D, const float* g_loss,
                                float coef, float* gE, vqb200_stream_t stream);

/* ---- K5: FSQ elementwise stage -------------------- models/vqvae.py:127-147,152-154 (a13) --
 * z_e: contiguous [B,d,T] (post project_in).  z_hard = z + (rint(z) - z) (half-to-even, unbounded);
 * idx[b,t] = (int64) trunc(sum_i fl(z_hard_i * basis_i)) evaluated left to right in fp32.
 * Unique-code count without host sync: workspace (vqb200_unique_workspace_bytes(), zeroed by this
 * call) holds a bitmap window plus a hash set for out-of-window codes.
 * out2 = {perplexity = float(#unique), dcr = float(1 - #unique/codebook_size)} (device). */
size_t vqb200_unique_workspace_bytes(void);
int vqb200_fsq_forward(const float* z_e, int64_t B, int64_t d, int64_t T, const int32_t* basis,
                       int64_t codebook_size, float* z_hard, int64_t* idx,
                       void* workspace, float* out2, vqb200_stream_t stream);

/* ---- K5: LFQ elementwise stage ----------------------- models/vqvae.py:171-191 (row a14) ----
 * z_q = z_e + (sign - z_e) with sign = (z_e > 0 ? +1 : -1); idx[b,t] = sum_i (z_q_i > 0) << i;
 * out3 = {loss = -mean(H_b(sigmoid(z_e)))*w, perplexity = #unique, dcr = 1 - #unique/2^d}. */
int vqb200_lfq_forward(const float* z_e, int64_t B, int64_t d, int64_t T, float entropy_loss_weight,
                       float* z_q, int64_t* idx, void* workspace, float* out3,
                       vqb200_stream_t stream);
/* dL/dz_e = g_zq + g_loss[0] * (-w/M) * dH/dp * p(1-p)   (closed form of autograd, row a14) */
int vqb200_lfq_backward(const float* z_e, const float* g_zq, const float* g_loss, int64_t numel,
                        float entropy_loss_weight, float* g_ze, vqb200_stream_t stream);

/* ---- K5f: FSQ / LFQ with their 1x1 projections fused in ------------ SURVEY.md §8f rank 1 ----
 * One pass over z [B,64,T] (contiguous, 16-byte aligned, T <= 128) for the whole module forward:
 *   FSQ models/vqvae.py:126-154: z_e = W_in z + b_in; z_hard = z_e + (round(z_e) - z_e); idx, metrics as
 *       vqb200_fsq_forward; out = W_out z_hard + b_out.
 *   LFQ models/vqvae.py:170-194: same with the sign instead of the rounding and the entropy loss (out3[0]).
 * W_in = project_in.weight [d,64,1], W_out = project_out.weight [64,d,1], d <= 16.  z_e [B,d,T] is written for the
 * backward pass (and is what the index is bit-exact against).  vqb200_proj_fused_backward is the autograd of both:
 *   g_z = W_in^T g_ze, g_ze = W_out^T g_out (+ LFQ entropy term, scaled by g_loss[0]);
 *   grads = [dW_in (d*64) | db_in (d) | dW_out (64*d) | db_out (64)]  (zeroed inside the call). */
int vqb200_proj_fused_eligible(int64_t B, int64_t D, int64_t d, int64_t T);
size_t vqb200_proj_fused_grad_floats(int64_t D, int64_t d);
int vqb200_fsq_fused_forward(const float* z, int64_t B, int64_t D, int64_t T,
                             const float* W_in, const float* b_in, const float* W_out, const float* b_out,
                             int64_t d, const int32_t* basis, int64_t codebook_size,
                             float* out, float* z_e, int64_t* idx, void* workspace, float* out2,
                             vqb200_stream_t stream);
int vqb200_lfq_fused_forward(const float* z, int64_t B, int64_t D, int64_t T,
                             const float* W_in, const float* b_in, const float* W_out, const float* b_out,
                             int64_t d, float entropy_loss_weight,
                             float* out, float* z_e, int64_t* idx, void* workspace, float* out3,
                             vqb200_stream_t stream);
int vqb200_proj_fused_backward(int is_lfq, const float* g_out, const float* z, const float* z_e,
                               int64_t B, int64_t D, int64_t T, const float* W_in, const float* W_out, int64_t d,
                               const float* g_loss, float entropy_loss_weight, float* g_z, float* grads,
                               vqb200_stream_t stream);

/* ---- token export / decode-only path -------------------------------- SURVEY.md §8f rank 2 ----
 * The reference never materialises tokens (scripts/deployment/export_motion.py:25-83 re-runs encoder -> quantizer ->
 * decoder per window).  A token = S codebook indices of `code_bits` bits (RVQ stages, or the LFQ bit pattern) followed
 * by d signed FSQ digits of `digit_bits` bits (digit = round(z_e), two's complement, saturated with *overflow = 1;
 * FSQ rounding is unbounded, models/vqvae.py:127-131, so the mixed-radix index of :135 is not invertible and the
 * digits are stored instead), little-endian bit stream, rounded up to whole bytes (<= 256 bits).
 *   codes  int32 [S, B*T];  z_e / digits  fp32 [B, d, T];  tokens  uint8 [B*T, vqb200_token_bytes(...)].
 * vqb200_tokens_decode rebuilds the quantized latent without the encoder:
 *   out[b,c,t] = (W_out digits + b_out)[c]  +  ((0 + E_0[i_0][c]) + E_1[i_1][c]) + ...     (:133, :94-98, :229)
 * E / K are HOST arrays of S device pointers / sizes. */
int64_t vqb200_token_bytes(int64_t S, int64_t code_bits, int64_t d, int64_t digit_bits);
int vqb200_tokens_pack(const int32_t* codes, int64_t S, int64_t code_bits, const float* z_e, int64_t d,
                       int64_t digit_bits, int64_t B, int64_t T, uint8_t* tokens, int32_t* overflow,
                       vqb200_stream_t stream);
int vqb200_tokens_unpack(const uint8_t* tokens, int64_t S, int64_t code_bits, int64_t d, int64_t digit_bits,
                         int64_t B, int64_t T, int32_t* codes, float* digits, vqb200_stream_t stream);
int vqb200_tokens_decode(const int32_t* codes, int64_t S, const float* const* E, const int64_t* K,
                         const float* digits, int64_t d, const float* W_out, const float* b_out,
                         int64_t B, int64_t C, int64_t T, float* out, vqb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* VQB200_H_ */
