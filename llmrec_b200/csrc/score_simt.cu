// Exact-fp32 full-catalog scoring + top-K (utility/batch_test.py:149-152 and :21-36,100-102) and the
// hit-vector lookup (:30-34).  mode 2 of llmrec_score_topk_f32: sequential-FMA fp32 scores, exact
// selection with ties -> lowest item id.  It is the on-device checker of the tcgen05 kernel and the
// path for K/d the tensor-core kernel does not cover.
#include "common.cuh"

namespace llmrec {

// scores for a block of users into scratch[b][n_items]; train items -> -inf
__global__ void __launch_bounds__(256) score_rows_kernel(const float* __restrict__ U, int64_t ldu, const float* __restrict__ I, int64_t ldi,
                                                         const int* __restrict__ users, int n_items, int d, float* __restrict__ S) {
  extern __shared__ float us[];  // d
  const int b = blockIdx.y;
  const float* u = U + (int64_t)users[b] * ldu;
  for (int j = threadIdx.x; j < d; j += blockDim.x) us[j] = u[j];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const float* it = I + (int64_t)i * ldi;
  float a = 0.f;
  for (int j = 0; j < d; ++j) a = fmaf(us[j], it[j], a);
  S[(int64_t)b * n_items + i] = a;
}
__global__ void mask_rows_kernel(const int* __restrict__ users, const int* __restrict__ rowptr, const int* __restrict__ col, int n_items, float* __restrict__ S) {
  const int b = blockIdx.x;
  const int u = users[b];
  for (int e = rowptr[u] + threadIdx.x; e < rowptr[u + 1]; e += blockDim.x) {
    int c = col[e];
    if (c >= 0 && c < n_items) S[(int64_t)b * n_items + c] = -INFINITY;
  }
}
// K rounds of block arg-max with (score desc, id asc) order
__global__ void __launch_bounds__(256) select_topk_kernel(float* __restrict__ S, int n_items, int K, int* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ float bv[8]; __shared__ int bi[8];
  __shared__ int win;
  const int b = blockIdx.x;
  float* s = S + (int64_t)b * n_items;
  for (int r = 0; r < K; ++r) {
    float best = -INFINITY; int besti = 0x7fffffff;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
      float v = s[i];
      if (v > best || (v == best && i < besti && v != -INFINITY)) { best = v; besti = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) if (bv[w] > best || (bv[w] == best && bi[w] < besti)) { best = bv[w]; besti = bi[w]; }
      bool ok = besti != 0x7fffffff && best != -INFINITY;
      out_idx[(int64_t)b * K + r] = ok ? besti : -1;
      if (out_val) out_val[(int64_t)b * K + r] = ok ? best : -INFINITY;
      if (ok) s[besti] = -INFINITY;
      win = besti;
    }
    __syncthreads();
  }
}

__global__ void topk_hits_kernel(const int* __restrict__ idx, int n_batch, int K, const int* __restrict__ users,
                                 const int* __restrict__ rowptr, const int* __restrict__ col, uint8_t* __restrict__ hits) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_batch * K) return;
  const int b = (int)(t / K);
  const int item = idx[t];
  const int u = users[b];
  uint8_t h = 0;
  if (item >= 0) for (int e = rowptr[u]; e < rowptr[u + 1]; ++e) if (col[e] == item) { h = 1; break; }
  hits[t] = h;
}

int score_topk_simt(const float* U, int64_t ldu, const float* I, int64_t ldi, const int* users, int n_batch, int n_items, int d,
                    const int* mask_rowptr, const int* mask_col, int K, int* out_idx, float* out_val,
                    float* scratch, int64_t scratch_elems, cudaStream_t st) {
  LLMREC_CHECK_ARG(scratch && scratch_elems >= (int64_t)n_items, "score_topk(simt): scratch too small");
  int64_t per = scratch_elems / n_items;
  if (per > 65535) per = 65535;
  for (int b0 = 0; b0 < n_batch; b0 += (int)per) {
    int nb = (int)((per < (int64_t)(n_batch - b0)) ? per : (int64_t)(n_batch - b0));
    dim3 grid((n_items + 255) / 256, nb);
    score_rows_kernel<<<grid, 256, d * sizeof(float), st>>>(U, ldu, I, ldi, users + b0, n_items, d, scratch);
    LLMREC_CHECK_LAUNCH("score_rows");
    if (mask_rowptr) { mask_rows_kernel<<<nb, 128, 0, st>>>(users + b0, mask_rowptr, mask_col, n_items, scratch); LLMREC_CHECK_LAUNCH("mask_rows"); }
    select_topk_kernel<<<nb, 256, 0, st>>>(scratch, n_items, K, out_idx + (int64_t)b0 * K, out_val ? out_val + (int64_t)b0 * K : nullptr);
    LLMREC_CHECK_LAUNCH("select_topk");
  }
  return 0;
}
}  // namespace llmrec

extern "C" int llmrec_topk_hits(const int32_t* idx, int32_t n_batch, int32_t K, const int32_t* users,
                                const int32_t* truth_rowptr, const int32_t* truth_col, uint8_t* hits, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  int64_t total = (int64_t)n_batch * K;
  if (total <= 0) return 0;
  llmrec::topk_hits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, llmrec::as_stream(stream)>>>(idx, n_batch, K, users, truth_rowptr, truth_col, hits);
  LLMREC_CHECK_LAUNCH("topk_hits");
  return 0;
}
