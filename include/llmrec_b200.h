/*
 * llmrec_b200 -- C ABI of the B200 (sm_100a) kernels behind the LLMRec training-and-eval hot path.
 *
 * Boundary contract: the reference has no FFI layer -- its boundary is the
 * Python API (MM_Model.forward, Trainer.bpr_loss/prune_loss, AdamW.step, test_torch).  These entry
 * points are what a ctypes/cffi binding of that API binds: plain device pointers, sizes and a
 * cudaStream_t.  No torch types.  Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all matrices are row-major fp32 with an explicit leading dimension (elements);
 *   - index arrays are int32 (CSR rowptr/col; nnz < 2^31) unless stated;
 *   - nothing is allocated, nothing synchronises the host; work is enqueued on `stream`;
 *   - return 0 on success, non-zero on error; llmrec_last_error() gives the message
 *     (the reference's only error convention is Python exceptions / sys.exit on NaN, main.py:287-289).
 * Reference citations are file:line into HKUDS/LLMRec @ 169f3614.
 */
#ifndef LLMREC_B200_H
#define LLMREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* llmrec_stream_t; /* cudaStream_t */

#define LLMREC_ABI_VERSION 2
#define LLMREC_MAX_SEG 16

int llmrec_abi_version(void);
const char* llmrec_last_error(void);
/* 1 when the running device is sm_100 (B200); kernels refuse to launch otherwise. */
int llmrec_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * Propagation SpMM.  Replaces torch.sparse.mm / torch.mm(COO, dense) at Models.py:57-61,152-183
 * (20 calls per forward) and their autograd transposes.
 *
 *   Y_s[r,:] (+)= epi( rs[r] * sum_{e in row r} v[e] * cs[col[e]] * X_s[col[e],:] )     s = 0..nseg-1
 *
 * One launch propagates `nseg` dense operands that share the sparsity pattern (the image / text /
 * 5 attribute / profile / ID operands of one graph direction), reading the index stream once.
 * vals, row_scale, col_scale may each be NULL (= 1).  The Trainer's graphs are binary patterns with
 * rs = (deg+1e-8)^-1/2 (main.py:114-126), so vals == NULL there.
 * seg flags: bit0 = row softmax over the d columns after scaling (Models.py:174-175); the optional
 * addend Z is added after the epilogue (used by the backward chain: g_out = g_direct + A^T g).
 * Work decomposition: `tiling` (required for the vectorised kernels; NULL falls back to the scalar
 * warp-per-row kernel) lists nnz-bounded tiles so that low-degree rows are batched and power-law rows are
 * split over several warps and reduced by a deterministic second pass.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  const float* X;   /* gathered operand  [n_cols x d], leading dimension ldx */
  float* Y;         /* output            [n_rows x d], ldy */
  const float* Z;   /* optional addend   [n_rows x d], ldz: Y = epi(...) + Z  (Z == Y accumulates in place) */
  int64_t ldx;
  int64_t ldy;
  int64_t ldz;
  int32_t flags;    /* LLMREC_SPMM_SOFTMAX */
  int32_t _pad;
} llmrec_spmm_seg;

typedef struct {
  const int32_t* tiles;       /* [n_tiles][8] = {first row, #complete rows (0 = one piece of a long row), e0, e1} followed by
                                 16 one-byte row-end offsets relative to e0 (pieces: {split index, first piece, #pieces, 0} instead);
                                 32-byte aligned; pieces of long rows are numbered first (llmrec_spmm_plan_tiles builds this) */
  const int32_t* split_row;   /* [n_split] rows that were cut into pieces */
  const int32_t* split_first; /* [n_split+1] first piece (tile id) of each split row */
  float* scratch;             /* [n_split_tiles * min(nseg,16) * d] partial sums of the pieces */
  int32_t n_tiles, n_split, n_split_tiles, _pad;
  int32_t* split_tickets;     /* optional int32[n_split * 64], zeroed once: with it the LAST piece of a long row to finish adds the row's
                                 pieces (in piece order) and runs the epilogue inside the same launch; NULL = second-pass kernel */
  const uint32_t* src_mask;   /* optional bitmask over SOURCE rows (= pattern columns): a clear bit promises that row of X is all zero;
                                 its fetch is skipped (identical sums).  NULL = every row is live. */
} llmrec_spmm_tiling;

/* Host-side planner (runs once per graph): tiles of <= tile_nnz (8..248) non-zeros, each a run of <= max_rows (<= 15)
 * complete rows or one piece of a longer row.  Call with tiles_out == NULL to get the sizes in
 * counts_out = {n_tiles, n_split, n_split_tiles}, allocate, call again.  All pointers are HOST memory. */
int llmrec_spmm_plan_tiles(const int32_t* rowptr_host, int32_t n_rows, int32_t tile_nnz, int32_t max_rows,
                           int32_t* tiles_out, int32_t* split_row_out, int32_t* split_first_out, int32_t* counts_out);

#define LLMREC_SPMM_SOFTMAX 1

int llmrec_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* vals,
                        const float* row_scale, const float* col_scale,
                        int32_t n_rows, int32_t n_cols, int32_t d,
                        const llmrec_spmm_seg* segs_host, int32_t nseg,
                        const llmrec_spmm_tiling* tiling_host, llmrec_stream_t stream);

/* Row-LIST form of the same product: only rows[0 .. *n_rows_dev) (device list, device-side length, capped at max_rows; entries < 0 skipped)
 * are computed and written; all other rows of Y are left untouched.  One segment, d in {32, 64, 128}.  For the products of a training
 * step that are provably consumed on a small row subset (the last propagation layer reaches the loss only through the batch's
 * neighbourhood -- dist.py "demand" mode).  Persistent grid, no host synchronisation. */
int llmrec_spmm_rows_f32(const int32_t* rowptr, const int32_t* col, const float* vals, const float* row_scale, const float* col_scale,
                         int32_t d, const llmrec_spmm_seg* seg_host, const int32_t* rows, const int32_t* n_rows_dev, int32_t max_rows,
                         const uint32_t* src_mask, int32_t cta_per_row /* 1: one CTA per listed row -- short lists of possibly very long rows */,
                         llmrec_stream_t stream);
int llmrec_row_softmax_bwd_rows_f32(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx,
                                    const int32_t* rows, const int32_t* n_rows_dev, int32_t max_rows, int32_t d, llmrec_stream_t stream);
/* Device-side row sets: mask |= {col[e] : e in rows list[.] of the CSR}, mask |= {ids}, and mask -> (unordered) id list with *count += #bits. */
int llmrec_mark_neighbors(const int32_t* rowptr, const int32_t* col, const int32_t* list, int32_t n_list, uint32_t* mask, llmrec_stream_t stream);
int llmrec_mark_ids(const int32_t* ids, int32_t n, uint32_t* mask, llmrec_stream_t stream);
int llmrec_compact_mask(const uint32_t* mask, int32_t n_bits, int32_t* list_out, int32_t* count, llmrec_stream_t stream);
int llmrec_zero_rows_f32(float* Y, int64_t ldy, const int32_t* idx, int32_t n, int32_t d, llmrec_stream_t stream);
int llmrec_assign_rows_f32(const float* G, int64_t ldg, const int32_t* idx, int32_t n, int32_t d, float* Y, int64_t ldy, llmrec_stream_t stream);
/* Dense AdamW over one [n_rows x width] table whose gradient is row-sparse: g is read only where row_mask has the row's bit set. */
int llmrec_adamw_step_rows_f32(float* p, const float* g, float* m, float* v, int64_t n_rows, int32_t width, const uint32_t* row_mask,
                               const double* state, float lr, float beta1, float beta2, float eps, float weight_decay, llmrec_stream_t stream);

/* Row softmax Y = softmax(X, dim=-1) and its backward dX = S*(dS - sum(dS*S)) (Models.py:174-175). */
int llmrec_row_softmax_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n, int32_t d, llmrec_stream_t stream);
int llmrec_row_softmax_bwd_f32(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx,
                               int64_t n, int32_t d, llmrec_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Side-feature projection  Y = X W^T + b   (nn.Linear at Models.py:145-150; 8 per forward) and its
 * weight gradient dW = dY^T X, db = colsum(dY) (autograd of the same; features are constants so
 * there is no dX).  X:[n x k] ldx, W:[d x k] (nn.Linear layout), Y:[n x d] ldy.
 * mode: 0 = tcgen05 3xTF32 (fp32-accurate, default), 1 = tcgen05 1xTF32, 2 = exact fp32 SIMT.
 * --------------------------------------------------------------------------------------------- */
int llmrec_proj_fwd_f32(const float* X, int64_t ldx, const float* W, const float* bias,
                        float* Y, int64_t ldy, int64_t n, int32_t k, int32_t d, int32_t mode,
                        float* wsplit /* 2*d*k floats, mode 0 only */, llmrec_stream_t stream);
int llmrec_proj_wgrad_f32(const float* X, int64_t ldx, const float* dY, int64_t lddy,
                          float* dW, float* db, int64_t n, int32_t k, int32_t d, int32_t accumulate,
                          int32_t mode, float* scratch, int64_t scratch_elems, llmrec_stream_t stream);
/* scratch elements needed by llmrec_proj_wgrad_f32 for this shape/mode */
int64_t llmrec_proj_wgrad_scratch(int64_t n, int32_t k, int32_t d, int32_t mode);

/* Grouped forms: all projections of one step (image, text, user, 5 attribute tables) in ONE persistent
 * launch.  Problems that share W (the 5 attribute tables use item_trans) share `wsplit`; wgrad problems
 * that share dW/db list them in order with accumulate = 1 after the first. */
typedef struct {
  const float* X; const float* W; const float* bias; float* Y; float* wsplit;
  int64_t ldx, ldy, n;
  int32_t k, _reserved;       /* 0 */
} llmrec_proj_fwd_problem;
typedef struct {
  const float* X; const float* dY; float* dW; float* db;
  int64_t ldx, lddy, n;
  int32_t k, accumulate;      /* LLMREC_WGRAD_ACCUMULATE */
} llmrec_proj_wgrad_problem;
#define LLMREC_WGRAD_ACCUMULATE 1
int llmrec_proj_fwd_group_f32(const llmrec_proj_fwd_problem* probs_host, int32_t n_prob, int32_t d, int32_t mode,
                              llmrec_stream_t stream);
int llmrec_proj_wgrad_group_f32(const llmrec_proj_wgrad_problem* probs_host, int32_t n_prob, int32_t d, int32_t mode,
                                float* scratch /* zero-initialised ONCE by the caller: its FIRST word is a ticket the kernels leave at zero */,
                                int64_t scratch_elems, llmrec_stream_t stream);
/* Stream note: when any problem has db != NULL, the bias column sums (they read dY only) are enqueued on a library-owned side
 * stream forked from `stream` with an event and joined back into it before the call returns, so they overlap the persistent
 * weight-gradient kernel; inside a stream capture this becomes a parallel graph branch.  One side stream + two events per device,
 * created on first use (the only objects the library ever creates); LLMREC_BRANCHES=0 keeps everything on `stream`. */
int64_t llmrec_proj_wgrad_group_scratch(const llmrec_proj_wgrad_problem* probs_host, int32_t n_prob, int32_t d, int32_t mode);

/* ---------------------------------------------------------------------------------------------
 * Fusion (Models.py:185-197):  out = mean(layer_0..layer_{L}) + sum_t coef[t] * x_t / max(||x_t||_2, 1e-12)
 * and its backward: d layer_l = g/(L+1) (written once to d_layer, may be NULL);
 * d x_t (+)= coef[t] * (g - y_t (y_t . g)) / max(||x_t||,1e-12),  y_t = x_t/max(||x_t||,1e-12).
 * Pointer tables are HOST arrays (copied into the launch parameters).
 * --------------------------------------------------------------------------------------------- */
/* rows: optional int32 device list of row ids to process (NULL = rows 0..n-1; n = list length otherwise).
 * llmrec_fuse_fwd_f32 with rows != NULL and n < 0: COMPACT form over |n| list entries -- the layer tables are read at rows[b], the
 * side operands and `out` at the compact position b (side features projected on the batch's rows only, hoisted mode). */
int llmrec_fuse_fwd_f32(const float* const* layers_host, const int64_t* ld_layers_host, int32_t n_layers,
                        const float* const* sides_host, const int64_t* ld_sides_host, const float* coef_host,
                        int32_t n_sides, float* out, int64_t ldo, const int32_t* rows, int64_t n, int32_t d,
                        llmrec_stream_t stream);
int llmrec_fuse_bwd_f32(const float* g, int64_t ldg, int32_t n_layers, float* d_layer, int64_t lddl,
                        const float* const* sides_host, const int64_t* ld_sides_host, const float* coef_host,
                        float* const* d_sides_host, const int64_t* ld_dsides_host, int32_t n_sides,
                        int32_t accumulate, const int32_t* rows, int64_t n, int32_t d, llmrec_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * BPR + prune heads (main.py:330-342 bpr_loss, :158-165 prune_loss, :232-254 the 8 heads).
 * For head h with user matrix XU_h and item matrix XI_h (row-major, ld):
 *   x_b   = <XU[u_b], XI[p_b]> - <XU[u_b], XI[n_b]>;  maxi_b = logsigmoid(x_b + 1e-8)
 *   keep  = the int((1-drop_rate)*B) smallest maxi (ties -> lower b);  mf_h = -mean(maxi[keep])
 *   emb_h = regs0/batch_size * sum_{t in u,p,n} 1/(2*sum||row_t||^2 + 1e-8)
 * loss += w_mf[h]*mf_h + w_emb[h]*emb_h.  Gradients w.r.t. the gathered rows are scatter-added
 * (atomicAdd) into GU_h / GI_h (same shapes as XU_h / XI_h; NULL = skip that head's grads).
 * out_host-visible results live in `out` (device, 4 floats per head: mf, emb, kept, _) .
 * idx are int32 device arrays of length B.  `n_keep` = int((1-drop_rate)*B) computed by the caller
 * in double arithmetic like the reference (main.py:161-162).
 * `meta` (may be NULL): DEVICE int32[2] = {live B', n_keep}.  When given, B is only the CAPACITY the launch is sized
 * for (index arrays and `work` hold B entries) and the kernels read the live length and n_keep from `meta` -- one
 * captured CUDA graph then serves every batch length <= B (the augmented-edge filter makes B' vary per step).
 * Two launches: score + per-head radix select of the kept set (ties -> lower position) + loss; scatter of row gradients.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  const float* XU; const float* XI;
  float* GU; float* GI;
  int64_t ldxu, ldxi, ldgu, ldgi;
  float w_mf, w_emb;
} llmrec_bpr_head;

int llmrec_bpr_heads_f32(const llmrec_bpr_head* heads_host, int32_t n_heads,
                         const int32_t* users, const int32_t* pos, const int32_t* neg, int32_t B,
                         int32_t n_keep, const int32_t* meta, float regs0_over_bs, int32_t d,
                         float* out /* [n_heads*4] */, float* loss_accum /* [1], += */,
                         float* work /* llmrec_bpr_work_elems() floats, zeroed once */, llmrec_stream_t stream);
int64_t llmrec_bpr_work_elems(int32_t n_heads, int32_t B);

/* First touch of every gradient buffer of a step, one launch instead of a memset per buffer: region r is written
 * G_r[n x width] = X_r ? c_r * X_r : 0, and *loss = sum_r 0.5 * c_r * sum(X_r^2) (OVERWRITTEN: this is the first term of
 * the step's loss) -- feat_reg_loss_calculation (main.py:151-156) and its gradient for the regions with X, plain zeroing
 * for the buffers the BPR heads scatter-add into.  <= 16 regions (128-bit accesses when width / ld % 4 == 0 and aligned).
 * scratch: llmrec_grad_init_scratch() floats, zeroed once by the caller (the kernel leaves its ticket at zero). */
typedef struct {
  float* G; const float* X;
  int64_t ldg, ldx, n;
  int32_t width; float c;
} llmrec_grad_region;
int llmrec_grad_init_f32(const llmrec_grad_region* regions_host, int32_t n_regions, float* loss, float* scratch, llmrec_stream_t stream);
int64_t llmrec_grad_init_scratch(void);

/* feat_reg_loss_calculation (main.py:151-156): loss += c * 0.5*sum(X^2) ; G = (accumulate? G:0) + c*X.
 * G may be NULL (loss only). */
int llmrec_sqnorm_grad_f32(const float* X, int64_t ldx, float* G, int64_t ldg, int64_t n, int32_t d,
                           float c, int32_t accumulate, float* loss_accum, float* partial /* >= 1024 floats */,
                           llmrec_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense AdamW over a list of tensors (torch.optim.AdamW defaults at main.py:100-104,278:
 * decoupled weight decay, bias correction).  `state` is a device block of 4 doubles
 * {step, lr/bc1, sqrt(bc2), _}; llmrec_adamw_advance increments step and recomputes the scalars
 * on device (CUDA-graph friendly).  Tensor tables are HOST arrays.
 * --------------------------------------------------------------------------------------------- */
int llmrec_adamw_advance(double* state, double lr, double beta1, double beta2, llmrec_stream_t stream);
int llmrec_adamw_step_f32(float* const* p_host, const float* const* g_host, float* const* m_host, float* const* v_host,
                          const int64_t* numel_host, int32_t n_tensors, const double* state,
                          float lr, float beta1, float beta2, float eps, float weight_decay,
                          llmrec_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Full-catalog scoring + top-K (utility/batch_test.py:149-152 scores, :21-36,100-102 ranking).
 *   score[b,i] = <U[users[b]], I[i]>;  train items of the user are excluded; top-K by score,
 *   ties -> lowest item id (heapq.nlargest over ascending candidates).
 * mask CSR: mask_rowptr int32[n_users_total+1], mask_col int32 with every row SORTED ASCENDING (the fused
 * selection walks it with a merge pointer).  users = int32[b].
 * out_idx int32 [b x K], out_val fp32 [b x K] (may be NULL; exact fp32 scores).  K <= 64.
 * mode: 0 = tcgen05 3xTF32 scoring with the select fused into the epilogue + exact fp32 rescoring of the
 *           K+16.. candidates (d in {32,64,96,128}); 2 = exact fp32 SIMT.  Both return the same lists unless two
 *           scores closer than the TF32x3 rounding (~1e-5 relative) straddle rank K+16.
 * --------------------------------------------------------------------------------------------- */
int llmrec_score_topk_f32(const float* U, int64_t ldu, const float* I, int64_t ldi,
                          const int32_t* users, int32_t n_batch, int32_t n_items, int32_t d,
                          const int32_t* mask_rowptr, const int32_t* mask_col,
                          int32_t K, int32_t* out_idx, float* out_val, int32_t mode,
                          float* scratch, int64_t scratch_elems, llmrec_stream_t stream);
int64_t llmrec_score_topk_scratch(int32_t n_batch, int32_t n_items, int32_t d, int32_t K, int32_t mode);

/* hits[b,j] = 1 if out_idx[b,j] in truth row of users[b] (test_set membership, batch_test.py:30-34). */
int llmrec_topk_hits(const int32_t* idx, int32_t n_batch, int32_t K, const int32_t* users,
                     const int32_t* truth_rowptr, const int32_t* truth_col, uint8_t* hits,
                     llmrec_stream_t stream);

/* test_flag == 'full' (utility/batch_test.py:38-68, utility/metrics.py:95-100): per-user ROC-AUC of the exact fp32 scores over the
 * candidates (all items minus the user's mask row), positives = the user's truth row; 0 when either class is empty (the
 * reference swallows sklearn's ValueError).  mask / truth rows sorted ascending.  out_auc fp32[n_batch]. */
int llmrec_user_auc_f32(const float* U, int64_t ldu, const float* I, int64_t ldi, const int32_t* users, int32_t n_batch, int32_t n_items, int32_t d,
                        const int32_t* mask_rowptr, const int32_t* mask_col, const int32_t* truth_rowptr, const int32_t* truth_col,
                        float* out_auc, llmrec_stream_t stream);

/* Host-side (CPU, no GPU needed) BPR item sampler, bit-identical to Data.sample()'s numpy draws
 * (utility/load_data.py:166-187): hand over numpy's legacy MT19937 state (np.random.get_state()), get the
 * positives / rejection-sampled negatives for `users` and the advanced state back.  All pointers HOST. */
int llmrec_host_sample_items(uint32_t* mt_key /* [624] */, int32_t* mt_pos, const int32_t* users, int32_t n_users_in_batch,
                             const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                             int32_t* pos_out, int32_t* neg_out);

/* One whole training batch on the host, bit-identical to Data.sample() followed by the augmented-edge step of
 * main.py:213-224, written straight into a (pinned) [3 x ld] int32 staging buffer (rows: users, pos, neg):
 *   users    random.sample(exist_users, batch) -- or `batch` random.choice draws when batch > n_exist (load_data.py:158-161)
 *            -- from CPython's MT19937 stream (random.getstate(): key[624] + pos)
 *   pos/neg  the np.random.randint draws of llmrec_host_sample_items from numpy's legacy global stream
 *   aug      random.sample(users, n_aug) over the batch list; (u, aug_pos[u], aug_neg[u]) appended when both ids < aug_limit
 *            (the column count of train_mat, main.py:84-85,221);
 *            INT32_MIN in the tables = uid missing from augmented_sample_dict (KeyError upstream -> return 4)
 * `*_pool_branch` = which branch of CPython's sample() applies (n <= setsize), decided by the caller in Python arithmetic.
 * stamp: int32[n_exist] zero-initialised once, `epoch` > 0 and different on every call; pool: int32[max(n_exist if
 * users_pool_branch, 0) and >= batch + n_aug].  *n_out = batch + kept augmented edges.  All pointers HOST. */
int llmrec_host_sample_batch(uint32_t* py_key, int32_t* py_pos, uint32_t* np_key, int32_t* np_pos,
                             const int32_t* exist_users, int32_t n_exist, int32_t batch, int32_t users_pool_branch,
                             const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                             int32_t n_aug, int32_t aug_pool_branch, const int32_t* aug_pos, const int32_t* aug_neg,
                             int32_t n_aug_table, int32_t aug_limit, int32_t* stamp, int32_t epoch, int32_t* pool,
                             int32_t* out, int64_t ld, int32_t* n_out);

/* Device-side batch sampler (utility/load_data.py:157-195 + main.py:216-224 on the GPU): ONE kernel fills the [4 x cap]
 * int32 index buffer of a training step -- rows users / pos / neg and the meta row {B', n_keep} looked up in meta_table[2*B' ..] -- from
 * DEVICE copies of exist_users, the train CSR (rows SORTED ascending) and the augmented-edge tables (ids < 0 or >= aug_limit are dropped,
 * as upstream's filter does; INT32_MIN = uid missing).  state = device uint64[2] {seed, step}; the kernel advances `step`, so the launch
 * can live inside a captured CUDA graph.  NOT bit-compatible with the reference's host RNG streams (that is llmrec_host_sample_batch, the
 * default): same distributions, counter-based generator.  key_scratch: uint32[max(n_exist, batch)]. */
int llmrec_device_sample_batch(const int32_t* exist_users, int32_t n_exist, int32_t batch,
                               const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                               int32_t n_aug, const int32_t* aug_pos, const int32_t* aug_neg, int32_t n_aug_table, int32_t aug_limit,
                               const int32_t* meta_table, int32_t cap, uint64_t* state, int32_t* out, uint32_t* key_scratch,
                               llmrec_stream_t stream);

/* Row helpers of the sharded (multi-GPU) path: epilogue of an item-side propagation applied AFTER the cross-rank
 * sum of per-rank partials, and gather / scatter-add of batch rows by index (idx < 0 = row not owned: zeros / skipped). */
int llmrec_row_scale_softmax_f32(const float* X, int64_t ldx, const float* scale, float* Y, int64_t ldy, int64_t n, int32_t d,
                                 int32_t softmax, llmrec_stream_t stream);
int llmrec_gather_rows_f32(const float* X, int64_t ldx, const int32_t* idx, int32_t n, int32_t d, float* out, int64_t ldo,
                           llmrec_stream_t stream);
int llmrec_scatter_add_rows_f32(const float* G, int64_t ldg, const int32_t* idx, int32_t n, int32_t d, float* Y, int64_t ldy,
                                llmrec_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Hoisted side-feature mode (Models.py:145-167 with dropout p = 0 and the mask branch off):
 * iu.ui.(X W^T + 1 b^T) = (iu.ui.X) W^T + (iu.ui.1) b^T, so the propagated TABLES are precomputed once and a step projects
 * only the gathered rows of its batch.  The three helpers below are what that needs besides gather / projection / wgrad:
 *   rank1_add      Y[r, c] += scale[r * lds] * bias[c]          (the (iu.ui.1) b^T term on the compact rows)
 *   scaled_colsum  out[c] (+)= sum_terms sum_r scale[r * lds] * G[r * ldg + c]   (bias gradient; scale == NULL means 1)
 *   feat_reg_gram  feat_reg (main.py:151-156) over ALL rows through the k x k Gram matrix of a propagated table X~ and
 *                  h = X~^T s, n2 = |s|^2:  loss += c/2 (tr(W G W^T) + 2 b^T W h + n2 |b|^2);  dW += c (W G + b h^T);
 *                  db += c (W h + n2 b).  Two launches: W G by the exact-fp32 SIMT GEMM, then one pass over [d x k].
 *                  scratch: llmrec_feat_reg_gram_scratch(d, k) floats, zeroed once (ticket re-zeroed by the kernel).
 * --------------------------------------------------------------------------------------------- */
typedef struct { float* Y; const float* scale; const float* bias; int64_t ldy, lds, n; int32_t width, _pad; } llmrec_rank1_block;
typedef struct { const float* G; const float* scale; int64_t ldg, lds, n; } llmrec_colsum_term;
int llmrec_rank1_add_f32(const llmrec_rank1_block* blocks_host, int32_t n_blocks, llmrec_stream_t stream);
int llmrec_scaled_colsum_f32(const llmrec_colsum_term* terms_host, int32_t n_terms, int32_t width, float* out, int32_t accumulate,
                             float* scratch /* llmrec_scaled_colsum_scratch(width) floats, zeroed once */, llmrec_stream_t stream);
int64_t llmrec_scaled_colsum_scratch(int32_t width);
int llmrec_feat_reg_gram_f32(const float* W, const float* bias, const float* G, const float* h, float n2, int32_t d, int32_t k, float c,
                             float* dW, float* db, float* loss_accum, float* scratch, llmrec_stream_t stream);
int64_t llmrec_feat_reg_gram_scratch(int32_t d, int32_t k);

/* small utilities used by the host mirror */
int llmrec_fill_f32(float* p, int64_t n, float v, llmrec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LLMREC_B200_H */
