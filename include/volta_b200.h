/*
 * volta_b200 C ABI - the drop-in boundary of the B200-native VOLTA encoder path.
 *
 * The reference (IntuitionMachines/volta) has exactly one native boundary on
 * this path: the pybind module `fused_layer_norm_cuda` with
 *   forward_affine(input, normalized_shape, gamma, beta, eps) -> (out, mean, invvar)
 *   backward_affine(dout, mean, invvar, input, shape, gamma, beta, eps) -> (dx, dgamma, dbeta)
 * (apex/csrc/layer_norm_cuda.cpp:139-160, 205-240).  Everything else on the
 * path is ATen calls issued from volta/encoders.py / volta/embeddings.py.
 * This header replaces both: `vb_layernorm_*` is the LN boundary, the other
 * entry points are the fused ops that stand in for the ATen call sites cited
 * on each declaration.  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *  - plain C: raw device pointers, int32/int64 extents, no torch types.
 *  - the caller owns every buffer; nothing is allocated or retained here
 *    (the only process state is a driver entry point looked up once).
 *  - every call launches on the `stream` argument only (a cudaStream_t passed
 *    as void*), never synchronises the device, and is re-entrant (PyTorch's
 *    autograd thread calls the backward entry points).
 *  - return value: 0 ok; negative = argument error (VB_ERR_*); positive =
 *    cudaError_t.  `vb_last_error()` returns a thread-local message.
 *  - dtype codes: VB_F32 = 0, VB_BF16 = 1.  Matrices are row-major with an
 *    explicit leading dimension in elements.
 */
#ifndef VOLTA_B200_H_
#define VOLTA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_F32 0
#define VB_BF16 1

#define VB_ERR_ARG (-1)
#define VB_ERR_UNSUPPORTED (-2)
#define VB_ERR_DEVICE (-3)
#define VB_ERR_WORKSPACE (-4)

/* ---- library ---------------------------------------------------------- */
int vb_version(void);                 /* ABI version (major*1000+minor); 1002: vb_gemm_args grew B2 / bias2 / split_M */
const char* vb_last_error(void);      /* thread-local, never NULL */
/* 0 if the current device is sm_100 (B200); VB_ERR_DEVICE otherwise. */
int vb_device_check(void);
int vb_sm_count(void);
/* Device-resident seed base for CUDA-graph replay.  After vb_set_seed_pointer(p) (p: device uint64, or NULL to
 * switch the feature off) every launch that draws a dropout mask - vb_gemm BIAS_DROP_RES, vb_layernorm_bwd dx_drop,
 * vb_attention_fwd/bwd, vb_eltwise DROPOUT - uses `seed + *p`, read on the device when the kernel runs.  A captured
 * graph therefore draws new masks on every replay once the host updates *p; forward and backward captured against
 * the same p regenerate identical masks.  Process-wide; set it around capture, not concurrently with other launches. */
int vb_set_seed_pointer(const void* dev_ptr);
/* Caps the number of SMs the persistent tcgen05 GEMM occupies (0 = all).  Used by the data-parallel wrapper during the
 * backward so NCCL's all-reduce blocks find free SMs instead of displacing GEMM blocks (tiles are assigned statically). */
int vb_set_gemm_sm_limit(int32_t sms);

/* ---- GEMM -------------------------------------------------------------
 * D[M,N] = epilogue( sum_k opA[m,k] * opB[n,k] )
 *
 * Replaces nn.Linear forward (volta/encoders.py:204-206,215-217,376,390,463,
 * 477,518,533; volta/embeddings.py:134,323,419; heads :646,663,686-696,727,
 * 747) and its autograd (dgrad / wgrad).
 *
 * a_kmajor = 1: A is [M,K] row-major (lda >= K).  a_kmajor = 0: A is stored
 *   [K,M] row-major (lda >= M), i.e. the reduction index is the slow one
 *   (wgrad: A = dY[rows, N] read as dY^T).
 * b_kmajor = 1: B is [N,K] row-major (nn.Linear weight layout).
 * b_kmajor = 0: B is [K,N] row-major (dgrad: B = W[out,in]; wgrad: B = X).
 *
 * backend: VB_GEMM_AUTO picks the tcgen05/TMEM/TMA kernel for bf16 operands
 *   whose layout TMA accepts (16-byte aligned bases and row pitches) and the
 *   SIMT kernel otherwise (fp32 validation mode, K=5 box projection, ...).
 *   VB_GEMM_TC / VB_GEMM_SIMT force one and fail if it cannot run.
 */
#define VB_GEMM_AUTO 0
#define VB_GEMM_TC 1
#define VB_GEMM_SIMT 2

#define VB_EPI_NONE 0          /* D = acc                                    */
#define VB_EPI_BIAS 1          /* D = acc + bias[n]                          */
#define VB_EPI_BIAS_GELU 2     /* x = acc + bias; D = gelu_erf(x) (encoders.py:136); aux_out = gelu'(x) for the backward */
#define VB_EPI_BIAS_RELU 3     /* D = relu(acc + bias)          (poolers, encoders.py:596-637)   */
#define VB_EPI_BIAS_DROP_RES 4 /* D = dropout(acc + bias) + aux_in   (encoders.py:410-423, 552-565) */
#define VB_EPI_MUL 5           /* D = acc * aux_in   (activation backward: aux_in = saved gelu'(x)) */
#define VB_EPI_ADD 6           /* D = acc + aux_in                           */
#define VB_EPI_DRELU 7         /* D = acc * (aux_in > 0)                     */
#define VB_EPI_ACCUM 8         /* D += acc  (fp32 D, atomic; split-K allowed) */
#define VB_EPI_BIAS_ADD 9      /* D = acc + bias[n] + aux_in                 */

typedef struct vb_gemm_args {
  const void* A; const void* B; void* D;
  int64_t M, N, K;
  int64_t lda, ldb, ldd;
  int32_t ab_dtype;      /* VB_F32 | VB_BF16 (A and B) */
  int32_t d_dtype;       /* VB_F32 | VB_BF16 */
  int32_t a_kmajor, b_kmajor;
  int32_t epilogue;      /* VB_EPI_* */
  int32_t backend;       /* VB_GEMM_* */
  const float* bias;     /* [N] fp32 or NULL */
  const void* aux_in;    /* [M, ld_aux] in d_dtype (residual / pre-activation) */
  void* aux_out;         /* [M, ld_aux] in d_dtype (activation derivative for BIAS_GELU) */
  int64_t ld_aux;
  float alpha;           /* scales acc before the epilogue (1.0 normally) */
  float dropout_p;       /* BIAS_DROP_RES only; 0 disables */
  uint64_t seed, offset; /* counter-hash RNG stream for the dropout mask: element index = m*N+n */
  int32_t split_k;       /* ACCUM only: 0 = auto */
  int32_t reserved;
  /* Row-split launch (tcgen05 back-end only): rows [0, split_M) use B / bias, rows [split_M, M) use B2 / bias2 (same
   * shape, pitch and layout as B).  One launch then serves the text rows and the vision rows of a two-stream sub-layer
   * whose hidden states share one buffer (different nn.Linear weights per stream, volta/encoders.py:204-217).
   * split_M must be a multiple of 256; 0 / B2 == NULL = ordinary GEMM. */
  const void* B2;
  const float* bias2;
  int64_t split_M;
  /* Device-side extents (NULL = none): int32 values in device memory, written by an earlier kernel of the stream
   * (vb_compact_plan).  Only rows [0, *m_limit) of A / D and reduction indices [0, *k_limit) hold data: output tiles
   * and k-blocks entirely beyond them are skipped, so the cost follows the device-side count while the launch
   * (grid, pointers, CUDA-graph node) stays static.  The tcgen05 back-end works in whole 256-row tiles / 64-deep
   * k-blocks: the caller keeps the operands zero from the limit up to the next multiple of 256. */
  const int32_t* m_limit;
  const int32_t* k_limit;
  /* Two row segments with their own counts: rows [0, limit_seg_M) are limited by *m_limit, rows [limit_seg_M, M) by
   * *m_limit2 (counted from limit_seg_M) - the packed text rows and the packed vision rows of a joint buffer.
   * limit_seg_M = 0: one segment.  Must be a multiple of 256 for the tcgen05 back-end. */
  const int32_t* m_limit2;
  int64_t limit_seg_M;
  /* Two reduction segments (VB_EPI_ACCUM on the tcgen05 back-end: the weight gradients of the text and the vision
   * Linear of a two-stream sub-layer whose rows share one buffer, in ONE launch): reduction indices [0, k_seg)
   * accumulate into D, indices [k_seg, K) into D2 (same shape and pitch as D; D2 == D for shared weights).  k_limit
   * bounds the first segment, k_limit2 the second (counted from k_seg).  k_seg must be a multiple of 64;
   * k_seg = 0 / D2 == NULL: one segment. */
  void* D2;
  const int32_t* k_limit2;
  int64_t k_seg;
  /* Host-side estimate of the rows a device-limited launch (m_limit / m_limit2) will really produce, 0 = unknown.  It only
   * steers the tile choice (how many tiles of which width fill the SMs in how many rounds); results do not depend on it. */
  int64_t m_hint;
} vb_gemm_args;

int vb_gemm(const vb_gemm_args* args, void* stream);
/* Measurement helper (bench.py's roofline leg): launches the same GEMM `repeats` times back to back between two CUDA
 * events on `stream` and returns the average device time per launch in milliseconds.  Synchronises the stream. */
int vb_gemm_bench(const vb_gemm_args* args, void* stream, int32_t repeats, float* ms_per_launch);
/* Which backend vb_gemm would use for these arguments (VB_GEMM_TC / VB_GEMM_SIMT). */
int vb_gemm_backend(const vb_gemm_args* args);

/* ---- LayerNorm --------------------------------------------------------
 * Replaces apex fused_layer_norm_cuda.forward_affine / backward_affine
 * (apex/csrc/layer_norm_cuda.cpp:139-160, 205-240; kernels
 * layer_norm_cuda_kernel.cu:279-322, 403-637) and the Python BertLayerNorm
 * (volta/encoders.py:48-61): biased variance, eps inside the sqrt.
 *
 * forward : y = gamma * (x - mean) * rstd + beta ; mean/rstd saved in fp32.
 *   out_scale multiplies y (LXMERT averages two LNs, embeddings.py:169);
 *   if `accum` != 0, y += ... instead of y = ....
 * backward: dx = rstd * (g - mean_H(g) - xhat * mean_H(g * xhat)),  g = dy*gamma*out_scale
 *   dgamma/dbeta are ACCUMULATED (fp32 atomics) - the caller zeroes them.
 *   Optional fused extras for the residual blocks:
 *     dx_drop  = dx * dropmask / (1-p)  (mask regenerated from seed/offset, index = row*H+col)
 *     dbias   += column sums of (dx_drop if given else dx)
 *
 * seqs (optional, NULL = every row is computed): which rows are TRAILING PADDING of their sample.  Rows [0, rows_a)
 *   are samples of L_a rows each with valid length len_a[sample] (int32, device; from vb_mask_lengths), rows
 *   [rows_a, rows) samples of L_b rows with len_b (NULL = none) - the text rows and the vision rows of a joint buffer.
 *   A row at or beyond its sample's length (length 0 = sample computed in full) is written as zeros (y / dx / dx_drop;
 *   mean = rstd = 0) and neither read nor accumulated: the reference computes such rows but nothing reads them
 *   (SURVEY.md 8a' item 3) and their gradient is exactly zero, so every valid row and every parameter gradient is
 *   unchanged.  Ignored when accum != 0.
 */
typedef struct vb_row_seqs {
  const int32_t* len_a; int32_t L_a; int64_t rows_a;
  const int32_t* len_b; int32_t L_b;
} vb_row_seqs;

/* gamma2 / beta2 / split_row (NULL / 0 = one LayerNorm): rows [split_row, rows) belong to a SECOND LayerNorm module -
 * the vision rows of a joint text+vision buffer (volta/encoders.py:392 v_LayerNorm) - normalised by the same launch with
 * their own affine parameters; the backward accumulates their dgamma2 / dbeta2 / dbias2 separately. */
int vb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int32_t H, float eps, float out_scale, int32_t accum, const vb_row_seqs* seqs,
                     const float* gamma2, const float* beta2, int64_t split_row, int32_t dtype, void* stream);
int vb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                     void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t H, float out_scale,
                     void* dx_drop, float dropout_p, uint64_t seed, uint64_t offset, float* dbias,
                     const vb_row_seqs* seqs, const float* gamma2, float* dgamma2, float* dbeta2, float* dbias2,
                     int64_t split_row, int32_t dtype, void* stream);

/* ---- gated attention --------------------------------------------------
 * Replaces BertGatedSelfAttention.forward after the Q/K/V projections
 * (volta/encoders.py:258-340): scores = QK^T/sqrt(d) + (1-mask)*-10000 for
 * each enabled pair, JOINT softmax over the concatenated key axis (text keys
 * first, then vision keys), dropout on the probabilities, context = P.V with
 * heads merged back to [rows, heads*d].
 *
 * Q/K/V are read in place from the packed projection outputs:
 *   qkv_t: [B*Lt, ld_t] with Q at column 0, K at column A, V at column 2A
 *   (A = heads*d); qkv_v likewise.  pair_mask bit0=tt bit1=tv bit2=vt bit3=vv.
 * mask_t/mask_v are the 0/1 int32 padding masks [B,Lt] / [B,Lv].
 * lse_t/lse_v: [B, heads, L] fp32 row log-sum-exp saved for the backward,
 * which recomputes P from them (no probability tensor is stored).
 * dropout: counter-hash mask indexed by ((b*heads+h)*Lq_total + q)*Lk_total + k over the
 * concatenated sequence [text; vision].
 */
#define VB_ATTN_AUTO 0
#define VB_ATTN_MMA 1
#define VB_ATTN_SIMT 2
#define VB_ATTN_TC 3    /* forward on the tcgen05 / TMEM / TMA kernel (bf16, d 64|128, <= 128 rows per sample); backward = VB_ATTN_MMA */

typedef struct vb_attn_args {
  const void* qkv_t; const void* qkv_v;
  int64_t ld_t, ld_v;
  const int32_t* mask_t; const int32_t* mask_v;
  void* ctx_t; void* ctx_v;           /* [B*L, heads*d], ld = ld_ctx_* */
  int64_t ld_ctx_t, ld_ctx_v;
  float* lse_t; float* lse_v;
  /* backward only */
  const void* dctx_t; const void* dctx_v;
  void* dqkv_t; void* dqkv_v;         /* same layout as qkv_* */
  int32_t B, Lt, Lv, heads, d;
  int32_t pair_mask;
  int32_t dtype;
  float scale;                        /* 1/sqrt(d) */
  float dropout_p_t, dropout_p_v;     /* text-query rows / vision-query rows */
  uint64_t seed, offset;
  int32_t backend;                    /* VB_ATTN_AUTO | VB_ATTN_MMA (warp-level mma.sync) | VB_ATTN_SIMT | VB_ATTN_TC */
  int32_t reserved;
  /* Optional [B] int32 per stream (NULL = none), from vb_mask_lengths: positions at or beyond len[b] are padding.  The
   * tensor-core back-end then skips whole 16-row tiles of padding: as keys their softmax weight is exactly 0 (additive
   * -10000 underflows), as queries nothing reads them (their context / dQ rows are written as zeros, the reference
   * leaves finite garbage there).  Masked positions below the length are handled by the additive mask as usual. */
  const int32_t* len_t; const int32_t* len_v;
  /* Optional [B] int32 per stream (needs len_*; tensor-core back-end only): the qkv / ctx / dctx / dqkv buffers of that
   * stream are PACKED - sample b owns rows [cu[b], cu[b] + len[b]) instead of [b*L, (b+1)*L).  Masks and lse keep
   * their [B, L] / [B, heads, L] layout. */
  const int32_t* cu_t; const int32_t* cu_v;
} vb_attn_args;

int vb_attention_fwd(const vb_attn_args* args, void* stream);
/* Attention probabilities for inspection (config.visualization + output_all_attention_masks, volta/encoders.py:342-356):
 * probs_t fp32 [B, heads, Lt, Lt+Lv], probs_v fp32 [B, heads, Lv, Lt+Lv] (text keys first; zero columns for a key stream
 * the row does not attend), joint softmax, no dropout; ctx / lse of args are ignored.  Padded layout only. */
int vb_attention_probs(const vb_attn_args* args, float* probs_t, float* probs_v, void* stream);
int vb_attention_bwd(const vb_attn_args* args, void* stream);
/* len[b] = 1 + index of the last position with mask[b, j] != 0, or 0 if the sample has none. */
int vb_mask_lengths(const int32_t* mask, int32_t B, int32_t L, int32_t* len, void* stream);

/* ---- embeddings -------------------------------------------------------
 * vb_embed_gather_fwd: out[r,:] = (accum? out[r,:] : 0) + sum_j table_j[idx_j[r], :]  for up to 4
 *   fp32 tables (word + position + token-type, volta/embeddings.py:55-65).  idx_j == NULL with
 *   table_j != NULL means a constant row `const_row_j` (VisualBERT's visual position / type rows,
 *   embeddings.py:352-386).  idx are int64 (loader dtype).  Output in `dtype`, stride ld_out.
 * vb_embed_gather_bwd: scatter-add of d_out rows into the fp32 table grads (atomic).  Rows whose index equals
 *   skip_idx[j] get no gradient (nn.Embedding(padding_idx=0) of the word table, embeddings.py:46); -1 = none.
 * vb_colsum_by_key: out[key[r], c] += a[r, c] for a handful of keys (token-type table gradient: 2-3 rows that
 *   every token hits - done as per-block column sums instead of contended atomics).
 * vb_box_project: out[r,:] (+)= loc[r,0:nl] . W[:,0:nl]^T + b   - the K = num_locs (4/5)
 *   Linear of embeddings.py:135,141 done as FMAs; bwd accumulates dW, db.
 */
int vb_embed_gather_fwd(const float* const* tables, const int64_t* const* idx, const int64_t* const_rows,
                        int32_t n_tables, void* out, int64_t ld_out, int64_t rows, int32_t H, int32_t accum,
                        int32_t dtype, void* stream);
int vb_embed_gather_bwd(float* const* dtables, const int64_t* const* idx, const int64_t* const_rows,
                        const int64_t* skip_idx, int32_t n_tables, const void* dout, int64_t ld_dout, int64_t rows,
                        int32_t H, int32_t dtype, void* stream);
int vb_colsum_by_key(const void* a, int32_t dtype, int64_t lda, int64_t rows, int64_t cols, const int64_t* keys,
                     int32_t nkeys, float* out, void* stream);
int vb_box_project_fwd(const float* loc, int32_t nl, const float* W, const float* b, void* out, int64_t ld_out,
                       int64_t rows, int32_t H, int32_t accum, int32_t dtype, void* stream);
int vb_box_project_bwd(const float* loc, int32_t nl, const void* dout, int64_t ld_dout, float* dW, float* db,
                       int64_t rows, int32_t H, int32_t dtype, void* stream);

/* ---- losses -----------------------------------------------------------
 * vb_xent_fwd: CrossEntropyLoss(ignore_index=-1), mean over non-ignored rows
 *   (volta/encoders.py:1030, 1095-1109).  logits [rows, ld] in `dtype`, C classes, labels int64.
 *   Writes per-row lse (fp32), loss_sum and count (fp32 scalars, ACCUMULATED - zero them first).
 * vb_xent_bwd: dlogits = (softmax - onehot) * (*dloss) / max(count,1) on non-ignored rows, 0 elsewhere.
 *   row_limit (both; NULL = none): device-side int32 row count - rows at or beyond *row_limit are not touched
 *   (the compacted masked-LM rows, see vb_compact_plan).
 * vb_kl_fwd / vb_kl_bwd: the `kl_1601` criterion (volta/losses.py:16-22):
 *   sum_{label==1} KL(target || softmax(pred)) / max(#label==1, 1) * weight; target fp32 [rows, C].
 *   `row_map`: pred row r reads logits row row_map_base(r) = (r / per) * per_ld + r % per + skip, so the
 *   global-feature row is skipped in place (encoders.py:1085) without a copy.
 */
int vb_xent_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C, float* lse,
                float* loss_sum, float* count, const int32_t* row_limit, int32_t dtype, void* stream);
int vb_xent_bwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C, const float* lse,
                const float* count, const float* dloss, void* dlogits, const int32_t* row_limit, int32_t dtype,
                void* stream);
int vb_kl_fwd(const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip, const int64_t* label,
              const float* target, int64_t rows, int32_t C, float weight, float* lse, float* loss_sum,
              float* count, int32_t dtype, void* stream);
int vb_kl_bwd(const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip, const int64_t* label,
              const float* target, int64_t rows, int32_t C, float weight, const float* lse, const float* count,
              const float* dloss, void* dlogits, int32_t dtype, void* stream);
/* loss = loss_sum / max(count, 1) for `n` (sum,count) pairs -> out[n] */
/* The remaining masked-region criteria of volta/losses.py over the same row mapping as vb_kl_*:
 *   VB_REGION_XENT  : xent_1600 / xent_400 (confidence-weighted, losses.py:87-107) and xent_1601 (conf = NULL, :125-133);
 *                     tgt_idx int64 [rows], conf fp32 [rows] or NULL; lse [rows] is written by fwd and read by bwd.
 *   VB_REGION_MSE / VB_REGION_HUBER : mse_2048 (:25-33) / huber_2048 (:110-122) against tgt_vec fp32 [rows, ld_t];
 *                     pass weight / C: the shared denominator max(#label==1, 1) then gives the reference's mean.
 * loss_sum / count are accumulated; add_count != 0 for exactly one criterion per step (they share the denominator). */
#define VB_REGION_XENT 0
#define VB_REGION_MSE 1
#define VB_REGION_HUBER 2
int vb_region_loss_fwd(int32_t kind, const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip,
                       const int64_t* label, const int64_t* tgt_idx, const float* conf, const float* tgt_vec, int64_t ld_t,
                       int64_t rows, int32_t C, float weight, float* lse, float* loss_sum, float* count, int32_t add_count,
                       int32_t dtype, void* stream);
int vb_region_loss_bwd(int32_t kind, const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip,
                       const int64_t* label, const int64_t* tgt_idx, const float* conf, const float* tgt_vec, int64_t ld_t,
                       int64_t rows, int32_t C, float weight, const float* lse, const float* count, const float* dloss,
                       void* dlogits, int32_t dtype, void* stream);
int vb_loss_finalize(const float* loss_sum, const float* count, float* out, int32_t n, void* stream);
/* ---- task criteria (volta/task_utils.py:24-27 LossMap, used at :103-137 and :237-282) -----------
 * vb_bce_logits_fwd : nn.BCEWithLogitsLoss over logits [rows, C] (bf16 / fp32, row stride ld) and fp32 targets:
 *                     row_loss[r] = sum_c max(x,0) - x t + log(1 + exp(-|x|));  *loss_out = scale * sum_r row_loss[r]
 *                     (scale = 1 / (rows * C) is reduction="mean"; deterministic: no atomics).
 * vb_bce_logits_bwd : dlogits = (sigmoid(x) - t) * (*dloss) * scale.
 * vb_argmax_pick    : arg[r] = argmax_c logits[r, c] (first of equal maxima), pick[r] = labels[r, arg[r]] (labels NULL:
 *                     the maximum itself) - `torch.max(logits, 1)` + the one-hot product of compute_score_with_logits
 *                     (:429-434), `preds` / `select_idx` of the VL-logit / V-logit branches (:113-121).  arg or pick may be NULL.
 * The CrossEntropyLoss entry of LossMap is vb_xent_fwd / vb_xent_bwd (negative labels ignored, mean over the rest). */
int vb_bce_logits_fwd(const void* logits, int64_t ld, const float* target, int64_t ld_t, int64_t rows, int32_t C,
                      float* row_loss, float scale, float* loss_out, int32_t dtype, void* stream);
int vb_bce_logits_bwd(const void* logits, int64_t ld, const float* target, int64_t ld_t, int64_t rows, int32_t C,
                      const float* dloss, float scale, void* dlogits, int64_t ld_d, int32_t dtype, void* stream);
int vb_argmax_pick(const void* logits, int64_t ld, const float* labels, int64_t ld_l, int64_t rows, int32_t C, int64_t* arg,
                   float* pick, int32_t dtype, void* stream);

/* ---- element-wise / reductions (glue that the reference issues as ATen ops) */
#define VB_ELT_COPY 0      /* out = a                     (also dtype conversion)       */
#define VB_ELT_ADD 1       /* out = a + b                                                */
#define VB_ELT_MUL 2       /* out = a * b                 (pooled_t * pooled_v, :770)    */
#define VB_ELT_RELU 3      /* out = relu(a)                                              */
#define VB_ELT_DRELU 4     /* out = a * (b > 0)                                          */
#define VB_ELT_GELU 5
#define VB_ELT_DGELU 6     /* out = a * gelu'(b)                                         */
#define VB_ELT_SCALE 7     /* out = a * alpha                                            */
#define VB_ELT_DROPOUT 8   /* out = a * mask / (1-p)      (counter hash, index = element) */
#define VB_ELT_AXPY 9      /* out = out + alpha * a                                      */
#define VB_ELT_SWISH 10    /* out = a * sigmoid(a)        (ACT2FN["swish"], volta/encoders.py:153-157) */
#define VB_ELT_DSWISH 11   /* out = swish'(a)             (saved next to the activation, multiplied in by the backward GEMM) */
/* 2-D strided element-wise op over [rows, cols]; a/b/out may have different dtypes. */
int vb_eltwise(int32_t op, const void* a, int32_t a_dtype, int64_t lda, const void* b, int32_t b_dtype,
               int64_t ldb, void* out, int32_t out_dtype, int64_t ldo, int64_t rows, int64_t cols,
               float alpha, float p, uint64_t seed, uint64_t offset, void* stream);
/* out[c] += sum_r a[r,c]   (fp32 out, atomic accumulate; bias gradients).  row_limit: optional device-side int32
 * row count (rows at or beyond *row_limit are not read), NULL = all `rows`. */
int vb_colsum(const void* a, int32_t dtype, int64_t lda, int64_t rows, int64_t cols, float* out,
              const int32_t* row_limit, void* stream);

/* ---- row compaction with a device-side count (masked-LM head) ----------
 * The reference computes vocabulary logits for all B*T positions (volta/encoders.py:686-696, 772) and
 * CrossEntropyLoss(ignore_index=-1) (encoders.py:1030, 1095-1099) discards the ~90 % that carry no label.  Packing the
 * labelled rows first and giving the count to the kernels as a DEVICE value (vb_gemm_args::m_limit / k_limit,
 * row_limit above) gives the same loss and gradients for a tenth of the decoder work, with static launches.
 * vb_compact_plan: one pass over labels[rows] (int64; label >= 0 = keep, order preserved):
 *   limits[0] = count, limits[1] = min(rows, count rounded up to `pad`);  src_index[i] = source row of compact row i
 *   (-1 beyond count);  pos[r] = compact row of source row r or -1;  labels_c[i] = label of compact row i (-1 beyond).
 * vb_compact_gather: dst[i,:] = src[src_index[i],:] for i < limits[0], zeros for limits[0] <= i < limits[1], rows
 *   beyond limits[1] are left untouched (never read by the limited kernels).
 * vb_compact_scatter: dst[r,:] = pos[r] >= 0 ? src_c[pos[r],:] : 0 for every r < rows (the gather's adjoint). */
int vb_compact_plan(const int64_t* labels, int64_t rows, int32_t pad, int32_t* limits, int32_t* src_index, int32_t* pos,
                    int64_t* labels_c, void* stream);
int vb_compact_gather(const void* src, int64_t ld_src, const int32_t* src_index, const int32_t* limits, void* dst,
                      int64_t ld_dst, int64_t rows, int32_t cols, int32_t dtype, void* stream);
int vb_compact_scatter(const void* src_c, int64_t ld_src, const int32_t* pos, void* dst, int64_t ld_dst, int64_t rows,
                       int32_t cols, int32_t dtype, void* stream);
/* Zero rows [row_offset + *row_begin, row_offset + *row_end) of buf (device-side bounds, at most max_rows rows): the
 * tail between a packed row count and its 256-row padding, which the tile-granular GEMMs read.  An optional second
 * range (row_begin2 / row_end2 / row_offset2, NULL = none) of the same buffer is cleared by the same launch. */
int vb_zero_rows(void* buf, int64_t ld, int32_t cols, const int32_t* row_begin, const int32_t* row_end, int64_t row_offset,
                 const int32_t* row_begin2, const int32_t* row_end2, int64_t row_offset2, int32_t max_rows, int32_t dtype,
                 void* stream);
/* out[r,:] = a[idx[r],:] (gather rows; idx int64) and its scatter-add adjoint */
int vb_gather_rows(const void* a, int64_t lda, const int64_t* idx, void* out, int64_t ldo, int64_t rows,
                   int64_t cols, int32_t dtype, void* stream);
int vb_scatter_rows_add(const void* a, int64_t lda, const int64_t* idx, void* out, int64_t ldo, int64_t rows,
                        int64_t cols, int32_t dtype, void* stream);
/* fp32 [rows, cols] -> bf16, n_mats matrices described by a device table; one launch for all weights. */
typedef struct vb_cast_item { const float* src; void* dst; int64_t n; } vb_cast_item;
int vb_multi_cast_bf16(const vb_cast_item* items_dev, int32_t n_items, int64_t max_n, void* stream);
/* VL-BERT sin/cos box embedding (volta/embeddings.py:102-124): boxes fp32 [rows, nl] -> out [rows, 8*dim] */
int vb_box_sincos(const float* boxes, int32_t nl, int32_t dim, void* out, int64_t ldo, int64_t rows,
                  int32_t dtype, void* stream);
/* VL-BERT input preparation (volta/embeddings.py:240-292):
 *  vb_vlbert_prep : per region row r: masked[r] = all(feat[r,:] == 0); out[r, 0:F] = masked ? mask_emb : feat[r,:]
 *                   (in `dtype`, stride ld_out); kind[r] = (r % K == K-1) ? end_kind : (masked && mvrc ? 1 : 0).
 *  vb_vlbert_index: text_end[b] = #non-zero ids; the reference shifts text position p by K for EVERY sample as
 *                   soon as p >= min_b text_end[b] (in-place add on an expanded view, embeddings.py:285-287);
 *                   tpos[b,t], opos[b,k] = text_end[b] (+1 for the last region), last_row[b] = b*K + K-1.
 *  vb_batched_colsum: out[b, c] = sum_r a[b, r, c]  (fp32 out). */
int vb_vlbert_prep(const float* feat, const float* mask_emb, void* out, int64_t ld_out, int64_t* masked, int64_t* kind,
                   int64_t rows, int32_t F, int32_t K, int32_t mvrc, int32_t end_kind, int32_t dtype, void* stream);
int vb_vlbert_index(const int64_t* ids, int32_t B, int32_t T, int32_t K, int64_t* tpos, int64_t* opos, int64_t* text_end,
                    int64_t* last_row, void* stream);
int vb_batched_colsum(const void* a, int32_t dtype, int64_t batch, int64_t rows, int64_t cols, float* out, void* stream);
/* ---- optimizer step over the flat buffers (the step right after the path) ---------------------------
 * vb_sumsq      : *out += sum x[i]^2   (global gradient norm for torch.nn.utils.clip_grad_norm_, train_concap.py:305-307)
 * vb_adamw_flat : pytorch_transformers.AdamW (train_concap.py:227; correct_bias folded into the per-group step size) over
 *                 n fp32 parameters p with gradients g and moments m, v; also writes the bf16 operand copy p16 (or NULL).
 *                 chunk_group[i] = hyper-parameter group of elements [64 i, 64 i + 64), 0xFFFF = not optimised;
 *                 group_hp[g] = {lr, weight_decay, step_size = lr*sqrt(1-b2^t)/(1-b1^t), active != 0} (fp32 x 4).
 *                 sumsq != NULL and max_norm > 0: gradients are scaled by min(1, max_norm / (sqrt(*sumsq) + 1e-6)).
 *                 beta1 / beta2 are doubles: the kernel uses (float)beta and (float)(1.0 - beta), the two fp32 scalars the
 *                 reference's Python expressions `mul_(beta)` / `add_(1 - beta, grad)` produce. */
int vb_sumsq(const float* x, int64_t n, float* out, void* stream);
int vb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const uint16_t* chunk_group,
                  const float* group_hp, double beta1, double beta2, float eps, const float* sumsq, float max_norm, void* stream);
/* vb_radam_flat : the reference's own RAdam (volta/optimization.py:9-93; train_task.py:227-228 `--optim RAdam`), same
 *                 buffers and chunk table:  m, v as above;  p -= (wd lr) p;  then  p -= step_size m / (sqrt(v) + eps)  when the
 *                 variance is rectifiable (N_sma >= 5, :66-81)  or  p -= step_size m  (:83, 91).
 *                 group_hp[g] = {weight_decay * lr, N_sma >= 5 ? 1 : 0, step_size, active != 0}. */
int vb_radam_flat(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const uint16_t* chunk_group,
                  const float* group_hp, double beta1, double beta2, float eps, const float* sumsq, float max_norm, void* stream);
/* ---- input pipeline: region batch assembly on the device ------------------------------------
 * vb_region_batch replaces the per-batch numpy work of the reference's Conceptual-Captions loader between the detector
 * output and the model call: box area + normalisation (volta/datasets/concept_cap_dataset.py:450-461), zeroing of the
 * regions picked for prediction and exclusion of the regions overlapping them from the mean (IoU > 0.4, :31-68,
 * :661-690), padding mask / labels (:562-568), and the mean-pooled global feature with its box and mask
 * (ConceptCapLoaderTrain.__iter__, :237-264).  The task datasets' reader (_image_features_reader.py:102-113) builds the
 * same layout but divides by the BOX count: identical when no region is picked and every sample has R boxes.
 *   raw_feat [B,R,F] in_dtype (fp32 as stored, or bf16 to halve the upload); raw_box [B,R,4] fp32 PIXELS x1,y1,x2,y2;
 *   image_hw [B,2] fp32 (h, w); num_boxes [B]; sel [B,R]: 0 = untouched, 1 = predicted + features zeroed, 2 = predicted,
 *   features kept (the host's random draws, :669-686).  Rows at or beyond num_boxes[b] are never read (treated as 0).
 *   global_pos: 0 none, 1 first, 2 last (config.add_global_imgfeat);  V = R + (global_pos != 0).
 *   Outputs: feat [B,V,F] out_dtype; loc [B,V,num_locs] fp32 (num_locs 4 or 5); mask [B,V] int64; label [B,R] int64
 *   (1 / -1); count [B] int32 = divisor of the global feature (regions not excluded, padded rows included, min 1).
 * fp32 outputs are bit-identical to the reference's numpy arithmetic (same operation order, round-to-nearest). */
int vb_region_batch(const void* raw_feat, int32_t in_dtype, const float* raw_box, const float* image_hw,
                    const int32_t* num_boxes, const uint8_t* sel, int32_t B, int32_t R, int32_t F, int32_t num_locs,
                    int32_t global_pos, void* feat, int32_t out_dtype, float* loc, int64_t* mask, int64_t* label,
                    int32_t* count, void* stream);
/* ---- retrieval ranking on the device -------------------------------------------------------
 * vb_retrieval_ranks replaces the per-caption / per-image numpy argsorts of eval_retrieval.py over its score and target
 * matrices (scores [C,I] fp32, pitch ld; targets [C,I] uint8, 1 = matching pair, pitch ld_t).  The place of an element
 * in a descending sort is the number of elements that beat it; equal scores keep index order (numpy's default argsort
 * leaves that case unspecified).
 *   axis 0 (image retrieval, eval_retrieval.py:201-208): rank[c] = place of the FIRST matching image of caption c in
 *          its row, I if it has none (:166);  topk (or NULL) [C,k] int32 = first k images of the row's descending order
 *          (:222, k = 20), -1 beyond I.
 *   axis 1 (text retrieval, :248-256): rank[i] = best place of a matching caption in column i, C if there is none. */
int vb_retrieval_ranks(const float* scores, int64_t ld, const uint8_t* targets, int64_t ld_t, int32_t C, int32_t I,
                       int32_t axis, int32_t* rank, int32_t* topk, int32_t k, void* stream);
/* ---- gradient all-reduce over NVLink / NVSwitch (replaces the per-bucket NCCL call of
 * apex/apex/parallel/distributed.py:449-454: dist.all_reduce(flat_bucket) + mul_(1/world)) --------------------------
 * In-place all-reduce (sum, then * scale) of elements [offset, offset + count) of a SYMMETRIC fp32 buffer: the same
 * allocation on every rank, peer-mapped.  buffer_ptrs_dev / signal_pad_ptrs_dev: device arrays [world] of the buffer /
 * signal-pad base addresses on every rank as seen from this rank (torch.distributed._symmetric_memory supplies them);
 * multicast_ptr: the NVSwitch multicast address of the buffer, or 0.  With a multicast address every rank reduces its
 * 1/world slice inside the switch (multimem.ld_reduce) and broadcasts it (multimem.st); without one it sums its slice
 * over peer loads and stores the result to every peer.  The kernel takes no shared memory and 128 threads per block, so
 * its blocks share SMs with the persistent GEMM kernels of the backward.  EVERY rank must issue the same sequence of
 * calls (same offset, count, blocks); bytes [4096, vb_allreduce_pad_bytes()) of the signal pads belong to this kernel
 * and must be zero at the first call.  offset and count are multiples of 4; blocks <= 304 (0 = one per SM). */
int vb_allreduce_f32(const uint64_t* buffer_ptrs_dev, uint64_t multicast_ptr, const uint64_t* signal_pad_ptrs_dev,
                     int32_t rank, int32_t world, int64_t offset, int64_t count, float scale, int32_t blocks, void* stream);
int vb_allreduce_pad_bytes(void);
/* L2-flush helper for benchmarks: writes `bytes` to buf. */
int vb_fill(void* buf, int64_t bytes, int32_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VOLTA_B200_H_ */
