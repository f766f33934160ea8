/*
 * topicgcn.h — C-ABI of the B200-native TopicGCN graph-convolution hot path.
 *
 * The reference (anargh-t/Graph-Convolutional-Networks-for-Text-Classification) has no FFI layer of its
 * own: its hot path is two `torch.spmm` calls per layer (reference layer.py:102, layer.py:106), the
 * elementwise tail of `GCN.forward` (layer.py:181-188) and the masked cross-entropy at the call site
 * (trainer.py:358-361).  This header is the boundary a maintainer would bind instead (ctypes stub in
 * INTEGRATION.md).  Every entry point states the reference line(s) it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in the signatures (`stream` is a
 *     cudaStream_t passed as void*; NULL = legacy default stream);
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - dense matrices are row-major fp32 with an explicit leading dimension (in elements);
 *   - CSR uses int32 `rowptr[n_rows+1]`, int32 `colidx[nnz]`, fp32 `vals[nnz]`;
 *   - return value: 0 = TG_OK, otherwise a tg_status code; tg_last_error() gives the message;
 *   - hot-path calls (tg_spmm*, tg_gc*, tg_dense*, tg_colsum*, tg_masked_ce*) never allocate, never
 *     synchronise the host and never use floating-point atomics: results are bitwise reproducible;
 *   - plan-time calls (tg_csr_from_coo, tg_csr_transpose, tg_plan_create) may allocate and synchronise.
 */
#ifndef TOPICGCN_H_
#define TOPICGCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_VERSION 100 /* 0.1.0 */

typedef enum tg_status {
    TG_OK = 0,
    TG_ERR_INVALID_ARG = 1,
    TG_ERR_CUDA = 2,
    TG_ERR_UNSUPPORTED = 3,
    TG_ERR_WORKSPACE = 4,
    TG_ERR_OVERFLOW = 5
} tg_status;

/* flags written by tg_csr_from_coo */
#define TG_COO_WAS_SORTED 1u     /* input was already row-major sorted and duplicate free */
#define TG_COO_HAD_DUPLICATES 2u /* duplicates were summed (torch coalesce() semantics) */

int tg_version(void);
const char* tg_last_error(void);
const char* tg_status_string(int status);

/* ------------------------------------------------------------------------------------------------
 * Boundary input: torch.sparse COO  ->  device CSR           (plan time, once per adjacency)
 * Replaces the per-call `coalesce()` + COO->CSR conversion that `torch.spmm(adj, support)` performs on
 * every forward/backward (reference layer.py:106 via ATen s_addmm_out_sparse_dense_cuda); input layout is
 * what utils.sparse_mx_to_torch_sparse_tensor produces (reference utils.py:196-203): int64 indices,
 * fp32 values, possibly flagged uncoalesced.
 *   rows/cols   [nnz] int64 device
 *   rowptr      [n_rows+1] int32 out; colidx/vals_out [nnz] out (only the first *nnz_out_host are valid)
 *   nnz_out_host host int64 out: number of stored entries after duplicate merging
 *   flags_host  host uint32 out: TG_COO_* bits
 * Duplicates (same row, col) are summed in input order (stable), like coalesce().
 * ---------------------------------------------------------------------------------------------- */
int tg_csr_from_coo(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz,
                    int64_t n_rows, int64_t n_cols, int32_t* rowptr, int32_t* colidx, float* vals_out,
                    int64_t* nnz_out_host, uint32_t* flags_host, void* stream);

/* CSR -> CSR of the transpose (used once for the backward operand, reference autograd `sparse.t().mm`,
 * SURVEY §3.3).  is_symmetric_host (optional) receives 1 when the transpose is bit-identical. */
int tg_csr_transpose(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                     int64_t n_cols, int64_t nnz, int32_t* t_rowptr, int32_t* t_colidx, float* t_vals,
                     int32_t* is_symmetric_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SpMM plan: row classification for the doc/topic skew (short rows -> one lane group per row, hub rows
 * with more than `hub_threshold` stored entries -> split into `segment_nnz` segments reduced in fixed
 * order).  Opaque; owns a few small device tables.
 */
typedef struct tg_plan tg_plan;
/*
 * When the matrix is square and its hub set is compact (<= 512 hub rows holding >= 1/8 of the entries: the
 * document-topic-topic graphs), the plan also carries the "column-chunk streaming" layout (tg_stream.cu): a
 * chunk-major copy of the hub rows' entries and an (index, value) interleaved copy of all entries.  The plan
 * therefore SNAPSHOTS the values: rebuild it when `vals` change.  vals may be NULL (no streaming layout); colidx may
 * be NULL too (then the split-row segments of the gather kernel run in storage order instead of column order).
 * Environment knobs read at creation: TG_STREAM=0 disables it, TG_STREAM_CHUNK=nodes per chunk (multiple of 32, default 128).
 * ---------------------------------------------------------------------------------------------- */
int tg_plan_create(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                   int64_t n_cols, int64_t nnz, int32_t hub_threshold /*<=0: default*/,
                   int32_t segment_nnz /*<=0: default*/, tg_plan** plan_out, void* stream);
void tg_plan_destroy(tg_plan* plan);
/* info: [0]=n_hub_rows [1]=n_segments [2]=hub_nnz [3]=max_row_nnz [4]=hub_threshold [5]=segment_nnz
 *       [6]=bits 0-1: role-specialised streaming kernels available (square graph, hub rows <= 1280),
 *           bits 2-3: rectangular sub-plan (1 = resident-table product X*W, 2 = all-hub product X^T*dS)
 *       [7]=nodes per hub chunk | hub slot groups << 16 | float4 chunks per lane of the document role << 24
 *           | lanes per hub slot sub-group << 32 */
int tg_plan_info(const tg_plan* plan, int64_t info_host[8]);
/* bytes of scratch a tg_spmm* / tg_gc* call with `n_feat` columns needs (partials of split hub rows) */
size_t tg_plan_workspace_bytes(const tg_plan* plan, int32_t n_feat);
/* number of kernels a tg_spmm* / tg_gc* call on this plan launches for a dense operand B (device pointer, only its
 * alignment is inspected) with leading dimension ldb and `n_feat` columns; philox: 1 = the call draws a Philox dropout
 * mask, 2 = the call is tg_gc2_loss_fwd_f32 (row-wise loss epilogue), 0 = neither;
 * out_vec4_ok: outputs / bias are 16-byte aligned with leading dimensions that are multiples of 4.  Instrumentation
 * only (the launch counter of bench.py); mirrors the kernel selection of the compute entry points exactly. */
int tg_plan_spmm_launches(const tg_plan* plan, const float* B, int64_t ldb, int32_t n_feat, int32_t philox,
                          int32_t out_vec4_ok);

/* ------------------------------------------------------------------------------------------------
 * Y[n_rows x F] = A_csr * B[n_cols x F]  (+ bias)            replaces layer.py:106 (+ :109-110)
 * Same call with the transposed CSR (or the same CSR when symmetric) is the backward dS = A^T dZ.
 * bias may be NULL.  out_scale: optional DEVICE scalar multiplied into the result (the upstream gradient of a scalar
 * loss in the backward pass, so that no separate scaling pass over dZ is needed); NULL = 1.
 * ---------------------------------------------------------------------------------------------- */
int tg_spmm_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                const float* B, int64_t ldb, float* Y, int64_t ldy, int32_t n_feat, const float* bias,
                const float* out_scale, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layer-1 fused forward:  H1 = dropout(relu(A*S + b1), p)   replaces layer.py:106,109-110,182,185
 *   keep_mask: optional explicit [n_rows x F] uint8 keep mask (1 keep / 0 drop, ld = F), the parity mode
 *              that reproduces torch's `bernoulli_(1-p)` mask bit for bit; NULL -> counter-based
 *              Philox4x32-10 keyed on (seed, offset, row, col), see tg_dropout_keep_mask.
 *   offset_dev: optional DEVICE uint64 added to `offset` when the Philox mask is used: a train step captured in a
 *              CUDA graph bumps it on the device, so every replay draws a fresh mask (NULL = none).
 *   p == 0 or training == 0 -> no dropout (eval mode, layer.py:185 `train=self.training`).
 *   raw_row_begin: rows >= raw_row_begin are stored as plain sums A*S without bias/relu/dropout (negative = none).
 *              The document-sharded multi-GPU mode uses it for the replicated topic rows, whose partial sums are
 *              all-reduced across ranks before the epilogue is applied to them (a second call on those rows).
 * ---------------------------------------------------------------------------------------------- */
int tg_gc1_fwd_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   const float* S, int64_t lds, const float* bias, float* H1, int64_t ldh, int32_t n_feat,
                   float p, int32_t training, const uint8_t* keep_mask, uint64_t seed, uint64_t offset,
                   const uint64_t* offset_dev, int64_t raw_row_begin, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Materialise the Philox keep mask the fused kernel uses (tests / debugging). */
int tg_dropout_keep_mask(uint8_t* keep_mask, int64_t n_rows, int32_t n_feat, float p, uint64_t seed,
                         uint64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layer-2 fused forward + loss:  Z2 = A*S2 + b2 ; loss = mean_{train rows} CE(Z2, y)
 * replaces layer.py:106,109-110 + trainer.py:358-359 and the first three backward nodes
 * (nll_loss_backward, _log_softmax_backward_data, index backward).
 *   row_label  [n_rows] int32, label of a train row, -1 for rows outside the train list
 *   inv_count  1 / (number of train rows)
 *   logits     optional out [n_rows x C]; dZ2 optional out [n_rows x C] = (softmax - onehot) * inv_count
 *              on train rows, 0 elsewhere;  row_loss out [n_rows] (0 on non-train rows)
 * The scalar loss is tg_reduce_sum_f32(row_loss) (fixed-order tree, deterministic).
 * ---------------------------------------------------------------------------------------------- */
int tg_gc2_loss_fwd_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx,
                        const float* vals, const float* S2, int64_t lds, const float* bias,
                        const int32_t* row_label, float inv_count, float* logits, int64_t ldl, float* dZ2,
                        int64_t ldd, float* row_loss, int32_t n_class, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Stand-alone masked cross-entropy on existing logits (same math as the fused epilogue above);
 * replaces trainer.py:358-359 when the caller keeps GCN.forward -> logits. */
int tg_masked_ce_f32(const float* logits, int64_t ldl, const int32_t* row_label, float inv_count,
                     float* dZ, int64_t ldd, float* row_loss, int64_t n_rows, int32_t n_class,
                     void* stream);

/* out[0] = sum(x[0..n)) in a fixed order; scratch must hold tg_reduce_scratch_floats(n) floats. */
int64_t tg_reduce_scratch_floats(int64_t n);
int tg_reduce_sum_f32(const float* x, int64_t n, float* scratch, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Plain fp32 product for a DENSE layer-1 feature matrix                replaces layer.py:102 when infeatn is dense
 *   trans_a = 0:  C[m x n] = A[m x k]   * B[k x n]        (support = X W1)
 *   trans_a = 1:  C[m x n] = A[k x m]^T * B[k x n]        (dW1 = X^T dS1; the long axis k is split and the partial
 *                 products are summed in a fixed order: deterministic, no float atomics)
 * scratch: tg_gemm_scratch_floats(trans_a, m, n, k) floats (0 for trans_a = 0).
 * ---------------------------------------------------------------------------------------------- */
int64_t tg_gemm_scratch_floats(int32_t trans_a, int64_t m, int64_t n, int64_t k);
int tg_gemm_f32(int32_t trans_a, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m,
                int64_t n, int64_t k, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Skinny dense products around the hidden layer                    replaces layer.py:102 (layer 2)
 *   tg_dense_nn_f32 :  C[n x c] = A[n x h] * W[h x c]                      (S2 = H1 * W2)
 *   tg_hidden_bwd_f32: fused backward of  S2 = H1*W2, dropout, relu, +b1   
 *        dH1 = dS2 * W2^T ;  dZ1 = dH1 * [H1 > 0] * scale   (H1 > 0 <=> kept by dropout and relu active)
 *        dW2 = H1^T * dS2 ;  db1 = colsum(dZ1)
 *      scale = 1/(1-p) in training mode, 1 otherwise.  `partials` is scratch of
 *      tg_hidden_bwd_scratch_floats(n, h, c) floats; dW2 [h x c] and db1 [h] are overwritten.
 *   tg_colsum_f32   :  out[c] = sum_rows X[n x c]                           (db2)
 * ---------------------------------------------------------------------------------------------- */
int tg_dense_nn_f32(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                    int64_t n, int32_t h, int32_t c, void* stream);
int64_t tg_hidden_bwd_scratch_floats(int64_t n, int32_t h, int32_t c);
int tg_hidden_bwd_f32(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2,
                      int64_t ldw, float scale, float* dZ1, int64_t ldz, float* dW2, float* db1,
                      float* partials, int64_t n, int32_t h, int32_t c, void* stream);
/* the same with a row limit for the two reductions: rows >= n_count get their dZ1 but do not count into dW2 / db1 (the
 * replicated topic rows of a document-sharded graph count on one rank only, yet every rank needs their dZ1) */
int tg_hidden_bwd_rows_f32(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2,
                           int64_t ldw, float scale, float* dZ1, int64_t ldz, float* dW2, float* db1,
                           float* partials, int64_t n, int32_t h, int32_t c, int64_t n_count, void* stream);
int64_t tg_colsum_scratch_floats(int64_t n, int32_t c);
int tg_colsum_f32(const float* X, int64_t ldx, int64_t n, int32_t c, float* scratch, float* out,
                  void* stream);

/* dZ = dH * [H > 0] * scale (elementwise backward of relu+dropout when the caller supplies dH1;
 * used by the stand-alone GraphConvolution path), replaces threshold_backward + mask mul. */
int tg_relu_dropout_bwd_f32(const float* H, int64_t ldh, const float* dH, int64_t lddh, float scale,
                            float* dZ, int64_t ldz, int64_t n, int32_t f, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation metrics on the device: the per-class counts behind utils.accuracy and
 * utils.macro_f1 (reference utils.py:25-109, called from trainer.py:385-389) in ONE kernel instead of
 * 3 * nclass + 1 `.item()` round trips.
 *   tg_class_counts_i32 : for every row with row_label >= 0: pred = argmax(logits[row, 0..c)) (ties: lowest
 *                         class);  counts[0*c + k] true positives, [1*c + k] false positives, [2*c + k]
 *                         false negatives of class k.  counts [3*c] int32 is zeroed by the call.
 * ---------------------------------------------------------------------------------------------- */
int tg_class_counts_i32(const float* logits, int64_t ldl, const int32_t* row_label, int64_t n, int32_t c,
                        int32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer for the featureless first layer.
 *   tg_adam_f32 : one Adam step on n elements, in place — torch.optim.Adam semantics without amsgrad
 *                 (reference trainer.py:307 `th.optim.Adam(model.parameters(), lr=0.02)`, step at :362):
 *                     g   = grad + weight_decay * param
 *                     m   = m + (1 - beta1) * (g - m) ;  v = beta2 * v + (1 - beta2) * g * g
 *                     param -= (lr / (1 - beta1^step)) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 *                 `step` is the 1-based step count; the hyper-parameters are doubles like torch's (1 - beta and the
 *                 bias corrections are formed in double, then rounded to fp32).  With X = I the weight of layer 1 is [N x hidden]
 *                 (1 GB at 1 M nodes): the update is one pass, 16 B read + 12 B written per element.
 * ---------------------------------------------------------------------------------------------- */
int tg_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOPICGCN_H_ */
