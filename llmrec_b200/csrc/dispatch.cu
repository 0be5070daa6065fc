// C-ABI entry points that pick between the tcgen05 kernels and the exact SIMT kernels by `mode`.
#include "common.cuh"

namespace llmrec {
int proj_fwd_simt(const float*, int64_t, const float*, const float*, float*, int64_t, int64_t, int, int, cudaStream_t);
int proj_wgrad_simt(const float*, int64_t, const float*, int64_t, float*, float*, int64_t, int, int, int, cudaStream_t);
int score_topk_simt(const float*, int64_t, const float*, int64_t, const int*, int, int, int, const int*, const int*, int, int*, float*, float*, int64_t, cudaStream_t);
}  // namespace llmrec
using namespace llmrec;

extern "C" int llmrec_proj_fwd_f32(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy,
                                   int64_t n, int32_t k, int32_t d, int32_t mode, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (n <= 0) return 0;
  (void)mode;
  return proj_fwd_simt(X, ldx, W, bias, Y, ldy, n, k, d, as_stream(stream));
}
extern "C" int64_t llmrec_proj_wgrad_scratch(int64_t n, int32_t k, int32_t d, int32_t mode) { (void)n; (void)k; (void)d; (void)mode; return 0; }
extern "C" int llmrec_proj_wgrad_f32(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dW, float* db,
                                     int64_t n, int32_t k, int32_t d, int32_t accumulate, int32_t mode,
                                     float* scratch, int64_t scratch_elems, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  (void)mode; (void)scratch; (void)scratch_elems;
  return proj_wgrad_simt(X, ldx, dY, lddy, dW, db, n, k, d, accumulate, as_stream(stream));
}
extern "C" int64_t llmrec_score_topk_scratch(int32_t n_batch, int32_t n_items, int32_t d, int32_t K, int32_t mode) {
  (void)d; (void)K; (void)mode;
  int64_t want = (int64_t)n_batch * n_items;
  int64_t cap = (int64_t)1 << 28;  // 1 GiB of fp32 scores at most; the kernel loops over user sub-blocks
  if (want > cap) want = (cap / n_items) * n_items;
  if (want < n_items) want = n_items;
  return want;
}
extern "C" int llmrec_score_topk_f32(const float* U, int64_t ldu, const float* I, int64_t ldi, const int32_t* users, int32_t n_batch,
                                     int32_t n_items, int32_t d, const int32_t* mask_rowptr, const int32_t* mask_col, int32_t K,
                                     int32_t* out_idx, float* out_val, int32_t mode, float* scratch, int64_t scratch_elems,
                                     llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(K >= 1 && K <= 64 && K <= n_items, "score_topk: K=%d unsupported (1..64, <= n_items)", K);
  if (n_batch <= 0) return 0;
  (void)mode;
  return score_topk_simt(U, ldu, I, ldi, users, n_batch, n_items, d, mask_rowptr, mask_col, K, out_idx, out_val, scratch, scratch_elems, as_stream(stream));
}
