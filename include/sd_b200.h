/*
 * sd_b200.h -- C ABI of the B200-native BrainEncoder + CLIPLoss hot path.
 *
 * The reference (SeanNobel/speech-decoding) is pure Python/PyTorch and has no
 * FFI of its own; the interface each entry point replaces is therefore a
 * PyTorch call site in the reference, cited as file:line under
 * /root/reference/.  The Python drop-in (speech-decoding_b200/speech_decoding)
 * binds these with ctypes.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a CUDA stream as void*.
 *   - every function returns 0 on success, non-zero on failure;
 *     sd_last_error() returns a thread-local message for the last failure.
 *   - all launches are asynchronous on `stream`; nothing synchronises.
 *   - the library never allocates device memory that outlives a call; the
 *     caller (PyTorch's caching allocator) owns every buffer.
 *   - re-entrant: forward runs on the main thread, backward on autograd's
 *     worker thread.
 *
 * Internal activation layout ("BTC"): (B, T, Cp) channels-last, Cp = C rounded
 * up to a multiple of 8, pad channels hold zeros.  dtype is SD_F32 or SD_BF16.
 * Packed weights: forward layout wf (G, taps, Np, Kp) and data-gradient layout
 * wd (G, taps, Kp, Np) with the taps reversed; both K-major for the MMA.
 */
#ifndef SD_B200_H
#define SD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_F32 0
#define SD_BF16 1
#define SD_TF32 2 /* conv / wgrad only: fp32 storage, tcgen05 kind::tf32 math (BASELINE.json configs[1] "fp32/TF32").
                     With the *_lo operand planes given (sd_tf32_split) every product is hi*hi + hi*lo + lo*hi:
                     3xTF32, fp32-class accuracy on the tensor cores.  Elementwise entry points take SD_F32. */

#define SD_ACT_NONE 0
#define SD_ACT_GELU 1 /* out = gelu(p); p optionally saved to `preact`           (models.py:194-195) */
#define SD_ACT_GLU 2  /* out[:, c] = p[:, c] * sigmoid(p[:, D2 + c]); p saved     (models.py:164)     */

#define SD_OUT_BTC 0     /* (B, T, Np) in `dtype`                                                      */
#define SD_OUT_NCT_F32 1 /* (B, N, T) fp32 contiguous: the module-facing layout of Z (models.py:195)   */

#define SD_IMPL_AUTO 0
#define SD_IMPL_SIMT 1 /* fp32-accumulate CUDA-core kernels (any dtype)                                */
#define SD_IMPL_TC 2   /* tcgen05/TMEM/TMA kernels (bf16)                                              */
#define SD_IMPL_TC_1CTA 3 /* SD_IMPL_TC without CTA-pair (cta_group::2) tiles: tests compare the variants   */
#define SD_IMPL_TC_WS 4   /* SD_IMPL_TC with weight-stationary CTA-pair tiles (experimental, not the default)  */

const char* sd_last_error(void);
int sd_abi_version(void);
/* sm count and compute capability of the current device */
int sd_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* force a kernel family for conv/wgrad (tests compare TC against SIMT); default SD_IMPL_AUTO */
int sd_set_impl(int impl);
/* cap the SMs the persistent tensor-core kernels occupy (0 = all).  Data-parallel runs leave a few SMs to the
 * concurrent NCCL kernels: a persistent grid that does not fit next to them is serialised into two waves. */
int sd_set_sm_limit(int n);

/* ---- layout conversion ------------------------------------------------------------------------ */
/* X (B,C,T) fp32 -> (B,T,Cp) dtype, zero padded.  Replaces the implicit layout of
 * einsum("oi,bit->bot") operands, models.py:65. */
int sd_nct_to_btc(const float* x, void* out, int B, int C, int T, int Cp, int dtype, void* stream);
/* the same from a bf16 X (sensor windows shipped from the host in bf16: the bf16 mode rounds X to bf16 here anyway, so the
 * result is bit-identical to shipping fp32 and half the bytes cross PCIe) */
int sd_nct_to_btc_bf16in(const void* x, void* out, int B, int C, int T, int Cp, int dtype, void* stream);
/* (B,T,Cp) dtype -> (B,C,T) fp32 (module-facing outputs / gradients of standalone sub-modules) */
int sd_btc_to_nct(const void* in, float* out, int B, int C, int T, int Cp, int dtype, void* stream);

/* One weight (N, K, taps) fp32 in PyTorch Conv1d layout (models.py:97-109,128-150,188-189)
 * -> wf (taps, Np, Kp) and wd (taps, Kp, Np) with reversed taps, zero padded, in `dtype`.
 * Either destination may be NULL. */
int sd_pack_weight(const float* w, void* wf, void* wd, int N, int K, int taps, int Np, int Kp, int dtype,
                   void* stream);
/* batched form: `table` is a device array of n entries of sd_pack_entry;
 * max_tiles = max over entries of ceil(Np/32)*ceil(Kp/32) (grid extent; smaller entries exit early) */
typedef struct {
  const float* w;
  void* wf;
  void* wd;
  int N, K, taps, Np, Kp, dtype;
} sd_pack_entry;
int sd_pack_weights(const sd_pack_entry* table, int n, int max_tiles, void* stream);

/* ---- SpatialAttention (models.py:45-65) + SpatialDropout (models.py:77-86) -------------------- */
/* a = Re(z)·cos + Im(z)·sin (models.py:49-53); w = softmax(a, -1) (:58); masked w·mask is what the
 * channel mix uses (dropping input channels == zeroing weight columns).
 *   z_ri (D1,K2,2) interleaved re/im; cos,sin (K2,C); mask (C) or NULL (eval)
 *   w_soft (D1,C) fp32 saved for backward; w_packed (1,D1p,Cp) `dtype` for sd_conv_fwd. */
#define SD_SA_MPARTS 32 /* frequency partitions of the logits kernel; scratch = SD_SA_MPARTS*D1*C floats */
int sd_sa_weights_fwd(const float* z_ri, const float* cos_t, const float* sin_t, const float* mask,
                      float* w_soft, void* w_packed, float* scratch, int D1, int K2, int C, int D1p, int Cp,
                      int dtype, void* stream);
/* dwm (D1,C) fp32 = gradient w.r.t. the masked mixing weights; dz_ri (D1,K2,2) = z.grad
 * (autograd convention dL/dRe + i dL/dIm).  cos_T / sin_T are the TRANSPOSED
 * tables (C,K2) (the buffers of models.py:39-40 transposed once by the caller) for coalesced access. */
int sd_sa_weights_bwd(const float* dwm, const float* w_soft, const float* mask, const float* cos_T,
                      const float* sin_T, float* dz_ri, int D1, int K2, int C, void* stream);

/* ---- implicit-GEMM Conv1d: forward and data-gradient ------------------------------------------ */
/* out[b,t,n] = act( A(n) ( bias[n] + res[b,t,n] + sum_j sum_k in[b, t+(j-(taps-1)/2)*dil, k] * w[g(b),j,n,k] ) )
 * with A(n)(v) = affine[n]*v + affine[Np+n] when `affine` is given (eval-mode BatchNorm folded into the epilogue,
 * SURVEY 8f rank 4), identity otherwise;
 * zero outside [0,T) ("same" padding, models.py:128-150).  With taps=1 this is the 1x1 convs
 * (models.py:97,188-189), the channel mix (models.py:65) and, with widx, the per-subject layer
 * (models.py:98-116, a grouped GEMM indexed by subject id).  dgrad is the same op on wd. */
typedef struct {
  const void* in;    /* (B,T,Kp) dtype                                      */
  const void* w;     /* (G,taps,Np,Kp) dtype                                */
  const float* bias; /* (N) fp32 or NULL                                    */
  const void* res;   /* (B,T,Np) dtype or NULL: residual add (models.py:156,160) */
  const int* widx;   /* (B) int32 group per sample, or NULL (G = 1)         */
  void* out;         /* see out_mode; with SD_ACT_GLU: (B,T,Op), Op = roundup8(N/2) */
  void* preact;      /* (B,T,Np) dtype or NULL: pre-activation, saved for backward */
  double* stats;     /* (2,Np) or NULL: += per-channel sum and sum of squares of the stored
                        pre-BN output (nn.BatchNorm1d batch statistics, models.py:135,143) */
  float* rownorm2;   /* (B) or NULL: += sum over (n,t) of out^2 per sample (CLIP norm, loss.py:65) */
  int B, T, K, Kp, N, Np, taps, dil, G;
  int act, out_mode, dtype;
  const float* affine; /* (2,Np) fp32 per-channel scale and shift applied before `act`, or NULL.  Inference only:
                          eval-mode BatchNorm1d + GELU (models.py:158,161) fused into the conv that feeds it
                          (tensor-core path, SD_ACT_GELU with a BTC output) */
  const void* in_lo;   /* SD_TF32 only, or NULL: low planes of `in` / `w` (same shapes) from sd_tf32_split; both or   */
  const void* w_lo;    /* neither.  Given: 3xTF32 (in, w must then be the HIGH planes).                                */
  const void* bnr_y;   /* BatchNorm backward fused into a data-gradient conv (bf16 tensor-core path), or NULL.  The conv result */
  const float* bnr_ss; /* is du, the gradient w.r.t. u = gelu(bn(y)) (models.py:158,161).  With bnr_y = that y (B,T,Np) and  */
                       /* bnr_ss = (2,Np) the BatchNorm's scale / shift, the epilogue stores g = du * gelu'(scale*y + shift)   */
                       /* to `out` instead of du and adds sum g, sum g*y over (b,t) to `stats` (2,Np) -- the reduce pass of   */
                       /* sd_bn_gelu_bwd_reduce; sd_bn_bwd_apply then runs with g_ready = 1.  Needs act NONE, BTC out, no bias. */
} sd_conv_args;
int sd_conv_fwd(const sd_conv_args* a, void* stream);

/* ---- Conv1d weight/bias gradient ----------------------------------------------------------------
 * dw[g, n, k, j] += sum_{b in group g} sum_t dout[b,t,n] * in[b, t+(j-(taps-1)/2)*dil, k]
 * written with element strides (gs, sn, sk, sj) so it lands in PyTorch's (N,K,taps) layout;
 * dbias[n] += sum_{b,t} dout[b,t,n].  Caller zeroes dw/dbias.  Groups: samples sorted by subject. */
typedef struct {
  const void* dout;         /* (B,T,Np) dtype */
  const void* in;           /* (B,T,Kp) dtype */
  float* dw;
  float* dbias;             /* (N) or NULL */
  const int* sample_order;  /* (B) int32: sample indices sorted by group, or NULL (identity) */
  const int* group_offsets; /* (G+1) int32 offsets into sample_order, or NULL (G = 1) */
  int B, T, K, Kp, N, Np, taps, dil, G;
  int64_t gs, sn, sk, sj;
  int dtype;
  void* workspace;          /* optional scratch for split-K partials (tensor-core path); NULL => atomics */
  int64_t workspace_bytes;
  const void* dout_lo;      /* SD_TF32 only, or NULL: low planes of dout / in (3xTF32, see sd_conv_args) */
  const void* in_lo;
} sd_wgrad_args;
int sd_conv_wgrad(const sd_wgrad_args* a, void* stream);

/* x (n fp32, n % 4 == 0) -> hi = x rounded to the nearest TF32 value (low 13 mantissa bits zero), lo = x - hi rounded
 * to TF32 (|x - hi - lo| <= 2^-23 |x|).  Operand planes of the 3xTF32 conv / wgrad: what cuDNN's TF32 convolution of the reference
 * (models.py:128-150 on a GPU) loses in one rounding is carried by the lo plane. */
int sd_tf32_split(const float* x, float* hi, float* lo, int64_t n, void* stream);

/* ---- BatchNorm1d + GELU (models.py:158,161) ----------------------------------------------------- */
/* column sums over a (rows, Cp) BTC tensor: stats[0:Cp] += sum, stats[Cp:2Cp] += sum of squares */
int sd_colstats(const void* x, double* stats, int64_t rows, int Cp, int dtype, void* stream);
/* out[c] += sum over the rows of x[:, c] for c < C, accumulated in fp64 (`scratch`: 2*Cp doubles, zeroed here): the conv
 * bias gradient sum_{b,t} dout[b,t,n] of the 3xTF32 mode, which keeps this cancellation-prone sum off the tensor core */
int sd_colsum_add(const void* x, float* out, double* scratch, int64_t rows, int C, int Cp, int dtype, void* stream);
/* training: mean/var from stats, running-stat update (momentum, unbiased var), num_batches_tracked += 1.
 * eval: use running stats.  ss (4,Cp) fp32 = scale, shift, mean, invstd. */
int sd_bn_finalize(const double* stats, int C, int Cp, int64_t n, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                   float eps, int training, float* ss, void* stream);
/* u = gelu(y*scale + shift) */
int sd_bn_gelu_fwd(const void* y, const float* ss, void* u, int64_t rows, int Cp, int dtype, void* stream);
/* with g = du * gelu'(y*scale+shift) (not stored): red (2,Cp) += sum g, sum g*y
 * (sum g*xhat = invstd*(sum g*y - mean*sum g) is formed in fp64 by sd_bn_bwd_apply) */
int sd_bn_gelu_bwd_reduce(void* du_g, const void* y, const float* ss, double* red, int64_t rows, int Cp,
                          int dtype, void* stream);
/* the same, and g (rounded to `dtype`) is written in place over du: sd_bn_bwd_apply_g then finishes without a second
 * evaluation of the GELU derivative (both passes are instruction-issue bound: 23 + 26 -> 24 + 13 instructions per element) */
int sd_bn_gelu_bwd_reduce_g(void* du_g, const void* y, const float* ss, double* red, int64_t rows, int Cp,
                            int dtype, void* stream);
/* dy = scale*(g - sum_g/n - xhat*sum_gx/n), g recomputed from du and y, written in place over du;
 * also dgamma = sum_gx, dbeta = sum_g (C).
 * n = n_stat = number of rows the statistics cover (rows * world size under SyncBN).
 * dgamma/dbeta are written as red * dparam_scale (1/world under SyncBN, where red is already the global sum
 * and the parameter gradients are summed across ranks once more afterwards).
 * eval-mode BN (training=0): dy = scale*g. */
int sd_bn_bwd_apply(void* g_dy, const void* y, const float* ss, const double* red, float* dgamma,
                    float* dbeta, int64_t rows, int64_t n_stat, float dparam_scale, int C, int Cp, int training,
                    int dtype, void* stream);
/* the same when g_dy already holds g (written by a conv with bnr_y): no GELU derivative is recomputed */
int sd_bn_bwd_apply_g(void* g_dy, const void* y, const float* ss, const double* red, float* dgamma,
                      float* dbeta, int64_t rows, int64_t n_stat, float dparam_scale, int C, int Cp, int training,
                      int dtype, void* stream);

/* ---- GLU (models.py:164) and GELU backward ------------------------------------------------------- */
int sd_glu_fwd(const void* y2, void* out, int64_t rows, int D2, int Np, int Op, int dtype, void* stream);
int sd_glu_bwd(const void* dout, const void* y2, void* dy2, int64_t rows, int D2, int Np, int Op, int dtype,
               void* stream);
/* dp = du * gelu'(p), in place over du; (rows,Cp) BTC */
int sd_gelu_bwd(void* du_dp, const void* p, int64_t rows, int Cp, int dtype, void* stream);
/* dZ (B,N,T) fp32 NCT, p (B,T,Np) -> dp (B,T,Np) = dZ^T * gelu'(p) */
int sd_gelu_bwd_nct(const float* dz, const void* p, void* dp, int B, int N, int T, int Np, int dtype,
                    void* stream);

/* ---- CLIPLoss (loss.py:38-84) --------------------------------------------------------------------- */
/* nrm2[i] = sum_d x[i,d]^2  (x (M,D) fp32) */
int sd_rownorm2(const float* x, float* nrm2, int M, int64_t D, void* stream);
/* dots[i,j] += sum_d x[i,d] * z[j,d]   (x (M,D), z (N,D) fp32; dots (M,N) fp32, caller zeroes):
 * the similarity GEMM torch.matmul(x, y.T), loss.py:68, on un-normalised rows. */
int sd_clip_dots(const float* x, const float* z, float* dots, int M, int N, int64_t D, void* stream);
/* Phase 1: logits[i,j] = exp(temp) * dots[i,j] / (|x_i||z_j|)  (loss.py:64-71);
 *   row_stat (M,2) = (max_j, sum_j exp(l - max)) over the local columns,
 *   col_lse (N)    = logsumexp_i (all M rows are local to every rank).
 * Multi-GPU: x holds the global batch (M rows), z the local shard (N columns); the caller
 * all-reduces row_stat across ranks between phase 1 and phase 2. */
int sd_clip_phase1(const float* dots, const float* xn2, const float* zn2, const float* temp, float* logits,
                   float* row_stat, float* col_lse, int M, int N, void* stream);
/* Phase 2: with row_lse (M) global,
 *   G = d loss / d logits = scale/2 * (softmax_rows + softmax_cols - 2*I), I at (diag0 + j, j)
 *   coef[i,j] = exp(temp) * G[i,j] / (|x_i||z_j|);  cz[j] = (sum_i G*logits)[j] / |z_j|^2
 *   partial[0] += this rank's share of the loss, partial[1] += sum G*logits (= d loss / d temp)
 *   scale = 1/M_global for reduction="mean", 1 for "sum" (loss.py:32,79).
 *   coef_t (N, roundup4(M)) or NULL: the same coefficients transposed (A operand of sd_clip_dz_tc). */
int sd_clip_phase2(const float* logits, const float* row_lse, const float* col_lse, const float* xn2,
                   const float* zn2, const float* temp, float scale, int diag0, float* coef, float* coef_t, float* cz,
                   float* partial, int M, int N, void* stream);
/* dz[j,d] = gscale * (sum_i coef[i,j] * x[i,d] - cz[j] * z[j,d])   (appendix A.5);
 * gscale: device pointer to the upstream gradient of the scalar loss, or NULL (= 1) */
int sd_clip_dz(const float* coef, const float* cz, const float* x, const float* z, float* dz, const float* gscale,
               int M, int N, int64_t D, void* stream);

/* Tensor-core (tcgen05 kind::tf32, operands read as fp32 by TMA) forms of the two streaming GEMMs, used by
 * the bf16 mode.  Need D % 4 == 0.  workspace: sd_clip_dots_workspace_bytes() bytes (split-K partials). */
int64_t sd_clip_dots_workspace_bytes(int M, int N, int64_t D);
int sd_clip_dots_tc(const float* x, const float* z, float* dots, void* workspace, int M, int N, int64_t D,
                    void* stream);
int sd_clip_dz_tc(const float* coef_t, const float* cz, const float* x, const float* z, float* dz,
                  const float* gscale, int M, int N, int64_t D, void* stream);

/* bf16 transport of the speech rows for data-parallel training (sd_b200/dist.py; the reference is single-device,
 * loss.py:60-71 sees the whole batch): every rank rounds its rows to bf16 once, the all-gather moves and the two
 * GEMMs re-read half the bytes, and the GEMMs run as tcgen05 kind::f16.  Need D % 8 == 0.
 *   sd_cast_rows_bf16:   y (M,D) bf16 = round(x); nrm2[i] = |y_i|^2 (norms of the ROUNDED rows)
 *   sd_clip_coef_t_bf16: coef_t (N, Mp) bf16 = coef^T, Mp = roundup8(M), pad columns zero
 *   sd_clip_dots_tc_bf16 / sd_clip_dz_tc_bf16: as the tf32 forms with x, z (dots) and coef_t, x (dz) in bf16;
 *   the projection row z and dz stay fp32. */
int sd_cast_rows_bf16(const float* x, void* y, float* nrm2, int M, int64_t D, void* stream);
/* squared norms of rows that already are bf16 (M, D), D % 8 == 0: speech embeddings shipped from the host in bf16 */
int sd_rownorm2_bf16(const void* x, float* nrm2, int M, int64_t D, void* stream);
/* gathered (world, M, 2) per-rank row statistics (max_j, sum_j exp(l - max)) over each rank's columns -> row_lse (M) over all
 * columns (loss.py:79 with global-batch negatives); world = 1 finishes the single-GPU statistics */
int sd_clip_merge_row_stats(const float* gathered, int world, int M, float* row_lse, void* stream);
int sd_clip_coef_t_bf16(const float* coef, void* coef_t, int M, int N, int Mp, void* stream);
int sd_clip_dots_tc_bf16(const void* x, const void* z, float* dots, void* workspace, int M, int N, int64_t D,
                         void* stream);
int sd_clip_dz_tc_bf16(const void* coef_t, const float* cz, const void* x, const float* z, float* dz,
                       const float* gscale, int M, int N, int64_t D, void* stream);

/* ---- optimizer ------------------------------------------------------------------ */
/* torch.optim.Adam over brain_encoder.parameters() + loss_func.parameters() (train.py:161-163,201-203) as ONE launch
 * over a device table of the parameters that received a gradient (absent subjects are skipped like torch does).
 * Per entry: step_size = lr / (1 - beta1^step), bias_correction2_sqrt = sqrt(1 - beta2^step) of THAT parameter's
 * step count; complex parameters are passed as 2n interleaved floats.  fp32, torch's operation order. */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float step_size;
  float bias_correction2_sqrt;
} sd_adam_entry;
int sd_adam_step(const sd_adam_entry* table, int n_entries, int blocks_per_entry, float beta1, float beta2, float eps,
                 float weight_decay, void* stream);

/* ---- batch preprocessing ---------------------------------------------------- */
/* Gwilliams2022Collator.forward (dataclass/gwilliams2022.py:653-661): for every (sample, channel) row of T samples
 *   y = x - mean(x[:baseline_len])                          baseline_correction_single, utils/preproc_utils.py:128-142
 *   out = clamp((y - median(y)) / IQR(y), +-clamp_lim)      scaleAndClamp / sklearn RobustScaler, utils/preproc_utils.py:69-90
 * with sklearn's float64 arithmetic (numpy linear-interpolated quartiles, IQR < 10*eps -> 1) and a float32 result.
 * x, out: (rows, T) fp32 contiguous, rows = B*C; T <= 2048; NaN-free input (the reference's nan-aware statistics are
 * not reproduced); in place (out == x) is allowed. */
int sd_collate_preproc(const float* x, float* out, int64_t rows, int T, int baseline_len, float clamp_lim, int clamp,
                       void* stream);

/* ---- peer memory for the data-parallel exchange (the reference is single-device, train.py:31) ---- */
/* The speech rows of every rank must reach every other rank before the CLIP GEMMs (loss.py:64-68 over the global batch).
 * Each rank owns a receive buffer allocated by sd_peer_alloc (plain cudaMalloc: it has a CUDA IPC handle -- the one
 * allocation of this library that outlives a call, owned by the Python PeerGather object and released by sd_peer_free),
 * peers map it with sd_ipc_open_handle and push their rows with sd_memcpy_async: a copy-engine transfer over NVLink that
 * occupies no SM.  Handles are cudaIpcMemHandle_t blobs of sd_ipc_handle_bytes() bytes exchanged by the host. */
int sd_peer_alloc(void** ptr, int64_t bytes);
int sd_peer_free(void* ptr);
int sd_ipc_handle_bytes(void);
int sd_ipc_get_handle(void* ptr, void* handle_out);
int sd_ipc_open_handle(const void* handle, void** ptr_out);
int sd_ipc_close_handle(void* ptr);
int sd_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
/* arrival fence of the copy-engine gather: returns (in stream order) once flags[0..world) all equal `expected`; every peer
 * writes its flag after its rows.  Traps after 10 s instead of hanging. */
int sd_peer_wait_flags(const int* flags, int world, int expected, void* stream);
/* Small all-gather through peer memory in ONE kernel: this rank's `bytes` (multiple of 4, <= cap_bytes) are stored into its
 * row of every peer's mailbox over NVLink, the epoch is published, the peers' epochs are awaited, and the world rows are
 * copied to `gathered` (world x bytes).  Mailbox (sd_peer_alloc, zero-filled, IPC-mapped by every peer):
 * [2][world][cap_bytes] payload followed by [2][world] int32 flags; peers_dev = device array of the world mailbox pointers
 * as mapped into THIS process; epoch = 1, 2, 3, ... in lock-step on all ranks.  Replaces the latency-bound NCCL
 * all-reduces of the data-parallel step (SyncBN statistics, CLIP row statistics, loss partials). */
int sd_peer_exchange(const void* src, int64_t bytes, void* const* peers_dev, int rank, int world, int64_t cap_bytes, int epoch,
                     void* gathered, void* stream);
/* out[i] = sum over q < world of in[q][i], in rank order (fp32 or fp64): the reduction of a gathered exchange */
int sd_sum_rows(const void* in, void* out, int world, int n, int is_f64, void* stream);
/* dst (device) <- src, up to 4 MB, by a KERNEL.  src may be pinned host memory (device-addressable under unified
 * addressing): the per-step integer tables and the Adam table of the CUDA-graph step (train.py:187-203 as one graph) are
 * fetched this way so that they never queue on the host-to-device copy engine behind the bulk transfer of the next batch */
int sd_copy_small(void* dst, const void* src, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_H */
