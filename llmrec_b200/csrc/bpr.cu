// BPR + prune heads (main.py:330-342 bpr_loss, :158-165 prune_loss, :232-254) and
// feat_reg (main.py:151-156): forward values AND row gradients, no host round trip.
//
// The reference does, per head: 3 row gathers, mul+sum, logsigmoid, a D2H copy + CPU argsort + H2D
// (x8 per step), mean of the kept 29 %.  Here: kernel A scores every (head, triplet) with one warp,
// kernel B ranks the B' values of a head inside one CTA (exact order statistics, ties -> lower
// position) and emits per-triplet gradient coefficients, kernel C scatter-adds the row gradients.
#include "common.cuh"

namespace llmrec {

constexpr int kMaxHeads = 16;

struct BprParams {
  llmrec_bpr_head head[kMaxHeads];
  int n_heads; const int* users; const int* pos; const int* neg; int B; int n_keep; float c_emb; int d;
  float* out; float* loss; float* work; unsigned* counter;
};
// work layout per head h (stride 7*B + 8): x[B], maxi[B], su[B], sp[B], sn[B], gcoef[B], keep[B], eu, ep, en
__device__ __forceinline__ float* work_of(const BprParams& p, int h) { return p.work + (size_t)h * (7 * (size_t)p.B + 8); }

__device__ __forceinline__ float logsigmoidf(float z) {  // min(z,0) - log1p(exp(-|z|))
  return fminf(z, 0.f) - log1pf(expf(-fabsf(z)));
}

__global__ void __launch_bounds__(256) bpr_score_kernel(const BprParams p) {
  const int h = blockIdx.y, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const llmrec_bpr_head hd = p.head[h];
  const float* u = hd.XU + (int64_t)p.users[b] * hd.ldxu;
  const float* pi = hd.XI + (int64_t)p.pos[b] * hd.ldxi;
  const float* ni = hd.XI + (int64_t)p.neg[b] * hd.ldxi;
  float dp = 0.f, dn = 0.f, su = 0.f, sp = 0.f, sn = 0.f;
  for (int j = lane; j < p.d; j += 32) {
    float a = u[j], q = pi[j], r = ni[j];
    dp = fmaf(a, q, dp); dn = fmaf(a, r, dn);
    su = fmaf(a, a, su); sp = fmaf(q, q, sp); sn = fmaf(r, r, sn);
  }
  dp = warp_sum(dp); dn = warp_sum(dn); su = warp_sum(su); sp = warp_sum(sp); sn = warp_sum(sn);
  if (lane == 0) {
    float* w = work_of(p, h);
    float x = dp - dn + 1e-8f;
    w[b] = x;
    w[p.B + b] = logsigmoidf(x);
    w[2 * p.B + b] = su; w[3 * p.B + b] = sp; w[4 * p.B + b] = sn;
  }
}

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < nw; ++i) t += red[i];
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// exact order statistics: rank_i = #{j : maxi_j < maxi_i or (== and j < i)}; 4 threads per element split the j range.
// grid (ceil(B/64), heads) -- the O(B^2) compare work is spread over ~150 CTAs instead of one per head.
__global__ void __launch_bounds__(256) bpr_rank_kernel(const BprParams p) {
  extern __shared__ float sv[];  // B values of this head
  const int h = blockIdx.y;
  float* w = work_of(p, h);
  const float* maxi = w + p.B;
  for (int i = threadIdx.x; i < p.B; i += blockDim.x) sv[i] = maxi[i];
  __syncthreads();
  const int i = blockIdx.x * 64 + (threadIdx.x >> 2), part = threadIdx.x & 3;
  int rank = 0;
  if (i < p.B) {
    const float v = sv[i];
    for (int j = part; j < p.B; j += 4) {
      const float o = sv[j];
      rank += (o < v) || (o == v && j < i);
    }
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  if (i < p.B && part == 0) {
    const bool keep = rank < p.n_keep;
    // d(-mean(maxi[keep]))/dx = -(1/n_keep) * sigmoid(-x); torch's log_sigmoid_backward form
    const float x = w[i];
    const float z = expf(-fabsf(x));
    const float dls = (x < 0.f) ? (1.f - z / (1.f + z)) : (z / (1.f + z));
    w[5 * p.B + i] = keep ? (-p.head[h].w_mf * dls / (float)p.n_keep) : 0.f;
    w[6 * p.B + i] = keep ? 1.f : 0.f;
  }
}

// one CTA per head: kept-mean (gcoef != 0 marks the kept set), regulariser sums, head outputs; last CTA adds the loss
__global__ void __launch_bounds__(1024) bpr_select_kernel(const BprParams p) {
  __shared__ float red[32];
  __shared__ bool last;
  const int h = blockIdx.x;
  float* w = work_of(p, h);
  float kept_sum = 0.f, su = 0.f, sp = 0.f, sn = 0.f;
  for (int i = threadIdx.x; i < p.B; i += blockDim.x) {
    if (w[6 * p.B + i] != 0.f) kept_sum += w[p.B + i];
    su += w[2 * p.B + i]; sp += w[3 * p.B + i]; sn += w[4 * p.B + i];
  }
  kept_sum = block_sum(kept_sum, red);
  su = block_sum(su, red); sp = block_sum(sp, red); sn = block_sum(sn, red);
  if (threadIdx.x == 0) {
    float mf = p.n_keep > 0 ? -(kept_sum / (float)p.n_keep) : 0.f / 0.f;  // mean of empty = nan like torch
    float du = 2.f * su + 1e-8f, dq = 2.f * sp + 1e-8f, dn = 2.f * sn + 1e-8f;
    float emb = p.c_emb * (1.f / du + 1.f / dq + 1.f / dn);
    p.out[h * 4 + 0] = mf; p.out[h * 4 + 1] = emb; p.out[h * 4 + 2] = (float)p.n_keep; p.out[h * 4 + 3] = 0.f;
    // d emb / d row = w_emb * c * (-4 row / (2S+eps)^2)
    const float we = p.head[h].w_emb * p.c_emb;
    w[7 * p.B + 0] = -4.f * we / (du * du);
    w[7 * p.B + 1] = -4.f * we / (dq * dq);
    w[7 * p.B + 2] = -4.f * we / (dn * dn);
    __threadfence();
    unsigned done = atomicAdd(p.counter, 1u);
    last = (done == (unsigned)p.n_heads - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {  // fixed-order head sum -> deterministic loss
    __threadfence();
    float tot = 0.f;
    for (int k = 0; k < p.n_heads; ++k) {
      volatile float* o = p.out + k * 4;
      tot += p.head[k].w_mf * o[0] + p.head[k].w_emb * o[1];
    }
    if (p.loss) *p.loss += tot;
    *p.counter = 0u;
  }
}

__global__ void __launch_bounds__(256) bpr_grad_kernel(const BprParams p) {
  const int h = blockIdx.y, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const llmrec_bpr_head hd = p.head[h];
  if (!hd.GU && !hd.GI) return;
  const float* w = work_of(p, h);
  const float g = w[5 * p.B + b];
  const float eu = w[7 * p.B + 0], ep = w[7 * p.B + 1], en = w[7 * p.B + 2];
  if (g == 0.f && hd.w_emb == 0.f) return;
  const int iu = p.users[b], ip = p.pos[b], in_ = p.neg[b];
  const float* u = hd.XU + (int64_t)iu * hd.ldxu;
  const float* pi = hd.XI + (int64_t)ip * hd.ldxi;
  const float* ni = hd.XI + (int64_t)in_ * hd.ldxi;
  for (int j = lane; j < p.d; j += 32) {
    float a = u[j], q = pi[j], r = ni[j];
    if (hd.GU) atomicAdd(hd.GU + (int64_t)iu * hd.ldgu + j, g * (q - r) + eu * a);
    if (hd.GI) {
      atomicAdd(hd.GI + (int64_t)ip * hd.ldgi + j, g * a + ep * q);
      atomicAdd(hd.GI + (int64_t)in_ * hd.ldgi + j, -g * a + en * r);
    }
  }
}

// feat_reg: loss += c * 0.5 * sum(X^2);  G = (acc ? G : 0) + c * X
__global__ void __launch_bounds__(256) sqnorm_grad_kernel(const float* X, int64_t ldx, float* G, int64_t ldg, int64_t n, int d,
                                                          float c, int accumulate, float* partial) {
  __shared__ float red[32];
  const int64_t total = n * d;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / d; int j = (int)(i - r * d);
    float v = X[r * ldx + j];
    s = fmaf(v, v, s);
    if (G) { float* gp = G + r * ldg + j; *gp = accumulate ? (*gp + c * v) : c * v; }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void sqnorm_final_kernel(const float* partial, int nb, float c, float* loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0 && loss) *loss += c * 0.5f * s;
}
}  // namespace llmrec

using namespace llmrec;

extern "C" int64_t llmrec_bpr_work_elems(int32_t n_heads, int32_t B) { return (int64_t)n_heads * (7 * (int64_t)B + 8) + 4; }

extern "C" int llmrec_bpr_heads_f32(const llmrec_bpr_head* heads, int32_t n_heads,
                                    const int32_t* users, const int32_t* pos, const int32_t* neg, int32_t B,
                                    int32_t n_keep, float regs0_over_bs, int32_t d,
                                    float* out, float* loss_accum, float* work, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(n_heads >= 1 && n_heads <= kMaxHeads, "bpr: n_heads=%d out of range", n_heads);
  LLMREC_CHECK_ARG(B >= 1 && (size_t)B * 4 <= 200 * 1024, "bpr: batch %d unsupported (max 51200 triplets per head)", B);
  BprParams p{};
  for (int h = 0; h < n_heads; ++h) p.head[h] = heads[h];
  p.n_heads = n_heads; p.users = users; p.pos = pos; p.neg = neg; p.B = B; p.n_keep = n_keep; p.c_emb = regs0_over_bs; p.d = d;
  p.out = out; p.loss = loss_accum; p.work = work;
  // the last 4 floats of `work` hold the head counter (zeroed by the caller once; the kernel re-zeroes it)
  p.counter = reinterpret_cast<unsigned*>(work + (size_t)n_heads * (7 * (size_t)B + 8));
  cudaStream_t st = as_stream(stream);
  dim3 grid((B + 7) / 8, n_heads);
  bpr_score_kernel<<<grid, 256, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("bpr_score");
  size_t smem = (size_t)B * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(bpr_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  bpr_rank_kernel<<<dim3((B + 63) / 64, n_heads), 256, smem, st>>>(p);
  LLMREC_CHECK_LAUNCH("bpr_rank");
  bpr_select_kernel<<<n_heads, 1024, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("bpr_select");
  bpr_grad_kernel<<<grid, 256, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("bpr_grad");
  return 0;
}

extern "C" int llmrec_sqnorm_grad_f32(const float* X, int64_t ldx, float* G, int64_t ldg, int64_t n, int32_t d,
                                      float c, int32_t accumulate, float* loss_accum, float* partial /* >= 1024 floats */,
                                      llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (n <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  int64_t total = n * d;
  int nb = (int)((total + 256 * 8 - 1) / (256 * 8));
  if (nb > 1024) nb = 1024;
  if (nb < 1) nb = 1;
  sqnorm_grad_kernel<<<nb, 256, 0, st>>>(X, ldx, G, ldg, n, d, c, accumulate, partial);
  LLMREC_CHECK_LAUNCH("sqnorm_grad");
  sqnorm_final_kernel<<<1, 256, 0, st>>>(partial, nb, c, loss_accum);
  LLMREC_CHECK_LAUNCH("sqnorm_final");
  return 0;
}
