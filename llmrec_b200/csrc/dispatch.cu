// C-ABI entry points that pick between the tcgen05 kernels and the exact SIMT kernels by `mode`
// (0 = tcgen05 3xTF32, 1 = tcgen05 TF32, 2 = fp32 SIMT) and by shape support.
#include "common.cuh"

namespace llmrec {
int proj_fwd_simt(const float*, int64_t, const float*, const float*, float*, int64_t, int64_t, int, int, cudaStream_t);
int proj_wgrad_simt(const float*, int64_t, const float*, int64_t, float*, float*, int64_t, int, int, int, cudaStream_t);
int score_topk_simt(const float*, int64_t, const float*, int64_t, const int*, int, int, int, const int*, const int*, int, int*, float*, float*, int64_t, cudaStream_t);
bool proj_tc_supported(int d, int64_t ldx, const void* X, int k, bool wgrad);
int proj_fwd_tc_group(const llmrec_proj_fwd_problem*, int, int, int, cudaStream_t);
int proj_wgrad_tc_group(const llmrec_proj_wgrad_problem*, int, int, int, float*, int64_t, cudaStream_t);
int64_t proj_wgrad_tc_scratch(const llmrec_proj_wgrad_problem*, int, int);
bool score_tc_supported(int d, int K, long long ldu, long long ldi, const void* U, const void* I);
long long score_tc_scratch(int n_batch, int n_items, int d, int K);
int score_topk_tc(const float*, long long, const float*, long long, const int*, int, int, int, const int*, const int*, int, int*, float*, float*, long long, cudaStream_t);
}  // namespace llmrec
using namespace llmrec;

static bool fwd_tc_ok(const llmrec_proj_fwd_problem* pr, int n, int d, int mode) {
  if (mode == 2 || n > 8) return false;
  for (int p = 0; p < n; ++p)
    if (!proj_tc_supported(d, pr[p].ldx, pr[p].X, pr[p].k, false) || pr[p].ldy % 4 != 0 || !aligned16(pr[p].Y) || !aligned16(pr[p].W) ||
        (pr[p].bias && !aligned16(pr[p].bias)) || (mode == 0 && !pr[p].wsplit))
      return false;
  return true;
}
static bool wg_tc_ok(const llmrec_proj_wgrad_problem* pr, int n, int d, int mode) {
  if (mode == 2 || n > 8) return false;
  for (int p = 0; p < n; ++p)
    if (!proj_tc_supported(d, pr[p].ldx, pr[p].X, pr[p].k, true) || pr[p].lddy % 4 != 0 || !aligned16(pr[p].dY)) return false;
  return true;
}

extern "C" int llmrec_proj_fwd_group_f32(const llmrec_proj_fwd_problem* pr, int32_t n_prob, int32_t d, int32_t mode, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(n_prob >= 1 && d >= 1, "proj_fwd_group: bad sizes");
  cudaStream_t st = as_stream(stream);
  for (int p0 = 0; p0 < n_prob; p0 += 8) {
    int np = n_prob - p0 < 8 ? n_prob - p0 : 8;
    if (fwd_tc_ok(pr + p0, np, d, mode)) {
      int rc = proj_fwd_tc_group(pr + p0, np, d, mode, st);
      if (rc) return rc;
    } else {
      for (int p = p0; p < p0 + np; ++p) {
        if (pr[p].n <= 0) continue;
        int rc = proj_fwd_simt(pr[p].X, pr[p].ldx, pr[p].W, pr[p].bias, pr[p].Y, pr[p].ldy, pr[p].n, pr[p].k, d, st);
        if (rc) return rc;
      }
    }
  }
  return 0;
}
extern "C" int llmrec_proj_fwd_f32(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy,
                                   int64_t n, int32_t k, int32_t d, int32_t mode, float* wsplit, llmrec_stream_t stream) {
  llmrec_proj_fwd_problem p{X, W, bias, Y, wsplit, ldx, ldy, n, k, 0};
  if (n <= 0) return 0;
  return llmrec_proj_fwd_group_f32(&p, 1, d, mode, stream);
}

extern "C" int64_t llmrec_proj_wgrad_group_scratch(const llmrec_proj_wgrad_problem* pr, int32_t n_prob, int32_t d, int32_t mode) {
  int64_t need = 0;
  for (int p0 = 0; p0 < n_prob; p0 += 8) {
    int np = n_prob - p0 < 8 ? n_prob - p0 : 8;
    if (wg_tc_ok(pr + p0, np, d, mode)) { int64_t s = proj_wgrad_tc_scratch(pr + p0, np, d); need = s > need ? s : need; }
  }
  return need;
}
extern "C" int llmrec_proj_wgrad_group_f32(const llmrec_proj_wgrad_problem* pr, int32_t n_prob, int32_t d, int32_t mode,
                                           float* scratch, int64_t scratch_elems, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(n_prob >= 1 && d >= 1, "proj_wgrad_group: bad sizes");
  cudaStream_t st = as_stream(stream);
  for (int p0 = 0; p0 < n_prob; p0 += 8) {
    int np = n_prob - p0 < 8 ? n_prob - p0 : 8;
    if (wg_tc_ok(pr + p0, np, d, mode)) {
      int rc = proj_wgrad_tc_group(pr + p0, np, d, mode, scratch, scratch_elems, st);
      if (rc) return rc;
    } else {
      for (int p = p0; p < p0 + np; ++p) {
        int rc = proj_wgrad_simt(pr[p].X, pr[p].ldx, pr[p].dY, pr[p].lddy, pr[p].dW, pr[p].db, pr[p].n, pr[p].k, d, pr[p].accumulate & LLMREC_WGRAD_ACCUMULATE, st);
        if (rc) return rc;
      }
    }
  }
  return 0;
}
extern "C" int64_t llmrec_proj_wgrad_scratch(int64_t n, int32_t k, int32_t d, int32_t mode) {
  llmrec_proj_wgrad_problem p{nullptr, nullptr, nullptr, nullptr, 4, 4, n, k, 0};
  // alignment of real pointers is checked at call time; size the scratch for the tensor-core path
  if (mode == 2 || d % 32 != 0 || d > 256 || k % 4 != 0) return 0;
  return proj_wgrad_tc_scratch(&p, 1, d);
}
extern "C" int llmrec_proj_wgrad_f32(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dW, float* db,
                                     int64_t n, int32_t k, int32_t d, int32_t accumulate, int32_t mode,
                                     float* scratch, int64_t scratch_elems, llmrec_stream_t stream) {
  llmrec_proj_wgrad_problem p{X, dY, dW, db, ldx, lddy, n, k, accumulate};
  return llmrec_proj_wgrad_group_f32(&p, 1, d, mode, scratch, scratch_elems, stream);
}

extern "C" int64_t llmrec_score_topk_scratch(int32_t n_batch, int32_t n_items, int32_t d, int32_t K, int32_t mode) {
  if (mode != 2 && (d == 32 || d == 64 || d == 96 || d == 128) && K <= 64) return score_tc_scratch(n_batch, n_items, d, K);
  int64_t want = (int64_t)n_batch * n_items;
  int64_t cap = (int64_t)1 << 28;  // 1 GiB of fp32 scores at most; the SIMT kernel loops over user sub-blocks
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
  if (mode != 2 && score_tc_supported(d, K, ldu, ldi, U, I))
    return score_topk_tc(U, ldu, I, ldi, users, n_batch, n_items, d, mask_rowptr, mask_col, K, out_idx, out_val, scratch, scratch_elems, as_stream(stream));
  return score_topk_simt(U, ldu, I, ldi, users, n_batch, n_items, d, mask_rowptr, mask_col, K, out_idx, out_val, scratch, scratch_elems, as_stream(stream));
}
