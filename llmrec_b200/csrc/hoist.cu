// Helpers of the hoisted side-feature mode (Models.py:145-167 with dropout p = 0, mask off).
//
// The propagated side features are linear in the raw feature tables:  iu.ui.(X W^T + 1 b^T) = (iu.ui.X) W^T + (iu.ui.1) b^T.
// With the propagated TABLES (ui.X, iu.ui.X, ...) and the propagated ones-vectors precomputed once, a training step needs
// the side features only on the <= 3 B' rows of its batch: gather those rows of the tables, run the grouped tcgen05
// projection on the compact block, and add  scale[r] * bias  (rank-1 term, llmrec_rank1_add_f32).  Weight gradients come
// from the same compact rows (tcgen05 wgrad) plus the scaled column sums for the biases (llmrec_scaled_colsum_f32);
// feat_reg (main.py:151-156) and its gradient over ALL rows reduce to a k x k Gram matrix per modality
// (llmrec_feat_reg_gram_f32):  |X~ W^T + s b^T|_F^2 = tr(W G W^T) + 2 b^T W h + |s|^2 |b|^2,  G = X~^T X~,  h = X~^T s.
#include "common.cuh"

namespace llmrec {

constexpr int kMaxBlocks = 32;

struct Rank1Params { llmrec_rank1_block blk[kMaxBlocks]; int n; };

// Y[r, c] += scale[r * lds] * bias[c]
__global__ void __launch_bounds__(256) rank1_add_kernel(const Rank1Params p) {
  const llmrec_rank1_block b = p.blk[blockIdx.y];
  const int64_t total = b.n * (int64_t)b.width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / b.width; const int c = (int)(i - r * b.width);
    b.Y[r * b.ldy + c] = fmaf(__ldg(b.scale + r * b.lds), __ldg(b.bias + c), b.Y[r * b.ldy + c]);
  }
}

struct ColsumParams { llmrec_colsum_term term[kMaxBlocks]; int n_terms; int width; float* out; int accumulate; float* partial; unsigned* ticket; };
constexpr int kColsumSlices = 128;

// out[c] (+)= sum over terms, rows of scale[r] * G[r, c]; grid (ceil(width/32), kColsumSlices): warp w of a CTA takes rows w, w+8.. of its slice,
// partials are combined in a fixed order by the last CTA of each column group (deterministic)
__global__ void __launch_bounds__(256) scaled_colsum_kernel(const ColsumParams p) {
  __shared__ float red[8][32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  for (int t = 0; t < p.n_terms; ++t) {
    const llmrec_colsum_term tm = p.term[t];
    const int64_t st = (int64_t)kColsumSlices * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + w;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;                       // independent chains: the loads of 4 rows are in flight together
    if (c < p.width) {
      for (; r + 3 * st < tm.n; r += 4 * st) {
        const float s0 = tm.scale ? __ldg(tm.scale + r * tm.lds) : 1.0f, s1 = tm.scale ? __ldg(tm.scale + (r + st) * tm.lds) : 1.0f;
        const float s2 = tm.scale ? __ldg(tm.scale + (r + 2 * st) * tm.lds) : 1.0f, s3 = tm.scale ? __ldg(tm.scale + (r + 3 * st) * tm.lds) : 1.0f;
        a0 = fmaf(s0, tm.G[r * tm.ldg + c], a0); a1 = fmaf(s1, tm.G[(r + st) * tm.ldg + c], a1);
        a2 = fmaf(s2, tm.G[(r + 2 * st) * tm.ldg + c], a2); a3 = fmaf(s3, tm.G[(r + 3 * st) * tm.ldg + c], a3);
      }
      for (; r < tm.n; r += st) a0 = fmaf(tm.scale ? __ldg(tm.scale + r * tm.lds) : 1.0f, tm.G[r * tm.ldg + c], a0);
    }
    acc += (a0 + a1) + (a2 + a3);
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    if (c < p.width) p.partial[((size_t)blockIdx.x * kColsumSlices + blockIdx.y) * 32 + lane] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(p.ticket + blockIdx.x, 1u) == kColsumSlices - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (w == 0 && c < p.width) {
    float t = 0.f;
    for (int s = 0; s < kColsumSlices; ++s) t += __ldcg(p.partial + ((size_t)blockIdx.x * kColsumSlices + s) * 32 + lane);
    p.out[c] = p.accumulate ? p.out[c] + t : t;
  }
  if (threadIdx.x == 0) p.ticket[blockIdx.x] = 0u;
}

// First half of feat_reg_gram: WG = W G for W[d x k], G[k x k] symmetric, as split-K partial tiles (75 MFLOP at k = 768: a 12-CTA GEMM takes
// 100 us, 96 CTAs take ~5).  grid (k/64 column tiles, kGramSplit K-chunks, d/64 row tiles); part[ks][i][j] holds the chunk's partial sum;
// the finish kernel adds the chunks in a fixed order (deterministic).
constexpr int kGramSplit = 8;
__global__ void __launch_bounds__(256) gram_wg_kernel(const float* __restrict__ W, const float* __restrict__ G, int d, int k, float* __restrict__ part) {
  __shared__ float Ws[16][64 + 4];   // [kk][row i]
  __shared__ float Gs[16][64 + 4];   // [kk][col j]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j0 = blockIdx.x * 64, i0 = blockIdx.z * 64;
  const int chunk = (k + kGramSplit - 1) / kGramSplit;
  const int l_beg = blockIdx.y * chunk, l_end = min(k, l_beg + chunk);
  float acc[4][4] = {};
  for (int l0 = l_beg; l0 < l_end; l0 += 16) {
    for (int t = threadIdx.x; t < 64 * 16; t += 256) {
      const int r = t >> 4, kk = t & 15;                 // W tile: consecutive threads walk k (contiguous in W)
      Ws[kk][r] = (i0 + r < d && l0 + kk < l_end) ? W[(size_t)(i0 + r) * k + l0 + kk] : 0.f;
      const int kk2 = t >> 6, c = t & 63;                // G tile: consecutive threads walk columns (contiguous in G)
      Gs[kk2][c] = (j0 + c < k && l0 + kk2 < l_end) ? __ldg(G + (size_t)(l0 + kk2) * k + j0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = Ws[kk][ty * 4 + q]; b[q] = Gs[kk][tx * 4 + q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(a[q], b[r], acc[q][r]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.y * d * k;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ty * 4 + q;
    if (i >= d) continue;
#pragma unroll
    for (int r = 0; r < 4; ++r) { const int j = j0 + tx * 4 + r; if (j < k) out[(size_t)i * k + j] = acc[q][r]; }
  }
}

// Second half of feat_reg_gram: WG[i,:] = sum over the kGramSplit partial tiles of gram_wg_kernel (fixed order).
// One CTA per row i of W[d x k]:  dW[i,:] += c (WG[i,:] + b_i h^T);  db_i += c (W[i,:].h + n2 b_i);
// loss += c/2 (WG[i,:].W[i,:] + 2 b_i W[i,:].h + n2 b_i^2) summed over i in a fixed order by the last CTA.
__global__ void __launch_bounds__(256) feat_reg_finish_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ WG,
                                                              const float* __restrict__ h, float n2, int d, int k, float c,
                                                              float* dW, float* db, float* loss, float* partial, unsigned* ticket) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const int i = blockIdx.x;
  const float bi = b ? b[i] : 0.f;
  float quad = 0.f, wh = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float w = W[(size_t)i * k + j];
    float a = 0.f;
#pragma unroll
    for (int ks = 0; ks < kGramSplit; ++ks) a += WG[((size_t)ks * d + i) * k + j];
    const float hj = h ? __ldg(h + j) : 0.f;
    quad = fmaf(a, w, quad);
    wh = fmaf(w, hj, wh);
    dW[(size_t)i * k + j] += c * (a + bi * hj);
  }
  auto bsum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
    return t;
  };
  quad = bsum(quad);
  wh = bsum(wh);
  if (threadIdx.x == 0) {
    if (db) db[i] += c * (wh + n2 * bi);
    partial[i] = 0.5f * c * (quad + 2.f * bi * wh + n2 * bi * bi);
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int q = 0; q < d; ++q) t += __ldcg(partial + q);
    if (loss) *loss += t;
    *ticket = 0u;
  }
}
}  // namespace llmrec

using namespace llmrec;

extern "C" int llmrec_rank1_add_f32(const llmrec_rank1_block* blocks, int32_t n_blocks, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(n_blocks >= 0 && n_blocks <= kMaxBlocks, "rank1_add: n_blocks=%d out of range", n_blocks);
  if (n_blocks == 0) return 0;
  Rank1Params p{};
  int64_t mx = 0;
  for (int i = 0; i < n_blocks; ++i) { p.blk[i] = blocks[i]; mx = max(mx, blocks[i].n * (int64_t)blocks[i].width); }
  p.n = n_blocks;
  if (mx <= 0) return 0;
  int64_t bx64 = (mx + 255) / 256;
  int bx = bx64 > 148 * 4 ? 148 * 4 : (int)bx64;
  rank1_add_kernel<<<dim3(bx, n_blocks), 256, 0, as_stream(stream)>>>(p);
  LLMREC_CHECK_LAUNCH("rank1_add");
  return 0;
}

extern "C" int64_t llmrec_scaled_colsum_scratch(int32_t width) { return (int64_t)((width + 31) / 32) * (kColsumSlices * 32 + 1) + 4; }

extern "C" int llmrec_scaled_colsum_f32(const llmrec_colsum_term* terms, int32_t n_terms, int32_t width, float* out, int32_t accumulate,
                                        float* scratch, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(n_terms >= 1 && n_terms <= kMaxBlocks && width >= 1, "scaled_colsum: n_terms=%d width=%d out of range", n_terms, width);
  ColsumParams p{};
  for (int i = 0; i < n_terms; ++i) p.term[i] = terms[i];
  const int groups = (width + 31) / 32;
  p.n_terms = n_terms; p.width = width; p.out = out; p.accumulate = accumulate;
  p.partial = scratch; p.ticket = reinterpret_cast<unsigned*>(scratch + (size_t)groups * kColsumSlices * 32);
  scaled_colsum_kernel<<<dim3(groups, kColsumSlices), 256, 0, as_stream(stream)>>>(p);
  LLMREC_CHECK_LAUNCH("scaled_colsum");
  return 0;
}

extern "C" int64_t llmrec_feat_reg_gram_scratch(int32_t d, int32_t k) { return (int64_t)kGramSplit * d * k + d + 4; }

extern "C" int llmrec_feat_reg_gram_f32(const float* W, const float* bias, const float* G, const float* h, float n2, int32_t d, int32_t k, float c,
                                        float* dW, float* db, float* loss_accum, float* scratch, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(d >= 1 && k >= 1 && scratch, "feat_reg_gram: d=%d k=%d out of range", d, k);
  cudaStream_t st = as_stream(stream);
  float* WG = scratch + d + 4;                                    // [kGramSplit][d x k] partial tiles after the partial/ticket block
  gram_wg_kernel<<<dim3((k + 63) / 64, kGramSplit, (d + 63) / 64), 256, 0, st>>>(W, G, d, k, WG);
  LLMREC_CHECK_LAUNCH("gram_wg");
  feat_reg_finish_kernel<<<d, 256, 0, st>>>(W, bias, WG, h, n2, d, k, c, dW, db, loss_accum, scratch, reinterpret_cast<unsigned*>(scratch + d));
  LLMREC_CHECK_LAUNCH("feat_reg_finish");
  return 0;
}
