// Dense fused AdamW (torch.optim.AdamW defaults; main.py:100-104,278): one launch updates every
// parameter tensor (p, m, v read-modify-write, g read: 28 B/param), decoupled weight decay first,
// then bias-corrected step -- the op order of torch's _single_tensor_adamw / foreach path.
#include "common.cuh"

namespace llmrec {
constexpr int kMaxTensors = 16;
struct AdamParams {
  float* p[kMaxTensors]; const float* g[kMaxTensors]; float* m[kMaxTensors]; float* v[kMaxTensors];
  int64_t numel[kMaxTensors]; int n;
  const double* state; float lr, b1, b2, eps, wd;
};

__global__ void adamw_advance_kernel(double* state, double lr, double b1, double b2) {
  double t = state[0] + 1.0;
  state[0] = t;
  state[1] = lr / (1.0 - pow(b1, t));      // step_size
  state[2] = sqrt(1.0 - pow(b2, t));       // bias_correction2_sqrt
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float decay, float b1, float b2, float eps, float step, float bc2s) {
  p *= decay;
  m = m + (1.f - b1) * (g - m);            // lerp(m, g, 1-b1)
  v = v * b2 + (1.f - b2) * g * g;
  float denom = sqrtf(v) / bc2s + eps;
  p = p - step * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
  const int t = blockIdx.y;
  const int64_t n = a.numel[t];
  float* p = a.p[t]; const float* g = a.g[t]; float* m = a.m[t]; float* v = a.v[t];
  const float step = (float)a.state[1], bc2s = (float)a.state[2];
  const float decay = 1.f - a.lr * a.wd;
  const int64_t n4 = n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    adam1(pp.x, gg.x, mm.x, vv.x, decay, a.b1, a.b2, a.eps, step, bc2s);
    adam1(pp.y, gg.y, mm.y, vv.y, decay, a.b1, a.b2, a.eps, step, bc2s);
    adam1(pp.z, gg.z, mm.z, vv.z, decay, a.b1, a.b2, a.eps, step, bc2s);
    adam1(pp.w, gg.w, mm.w, vv.w, decay, a.b1, a.b2, a.eps, step, bc2s);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
    adam1(p[i], g[i], m[i], v[i], decay, a.b1, a.b2, a.eps, step, bc2s);
}
// One [n_rows x w] table whose gradient is ROW-SPARSE: g is read only for rows whose bit is set in `row_mask`, every other row takes the
// g = 0 update (weight decay and moment decay still apply: the reference's AdamW is dense, main.py:100-104).  24 B/param instead of 28,
// and no dense gradient buffer has to be zeroed and written per step (the user table of the large synthetic graph: 5 GB each).
__global__ void __launch_bounds__(256) adamw_rows_kernel(float* p, const float* g, float* m, float* v, int64_t n_rows, int w4, const unsigned* row_mask,
                                                         const double* state, float lr, float b1, float b2, float eps, float wd) {
  const float step = (float)state[1], bc2s = (float)state[2];
  const float decay = 1.f - lr * wd;
  const int64_t total = n_rows * (int64_t)w4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / w4;
    const bool on = (__ldg(row_mask + (r >> 5)) >> (r & 31)) & 1u;
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = on ? reinterpret_cast<const float4*>(g)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam1(pp.x, gg.x, mm.x, vv.x, decay, b1, b2, eps, step, bc2s);
    adam1(pp.y, gg.y, mm.y, vv.y, decay, b1, b2, eps, step, bc2s);
    adam1(pp.z, gg.z, mm.z, vv.z, decay, b1, b2, eps, step, bc2s);
    adam1(pp.w, gg.w, mm.w, vv.w, decay, b1, b2, eps, step, bc2s);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
}
}  // namespace llmrec

using namespace llmrec;

extern "C" int llmrec_adamw_step_rows_f32(float* p, const float* g, float* m, float* v, int64_t n_rows, int32_t width, const uint32_t* row_mask,
                                          const double* state, float lr, float beta1, float beta2, float eps, float weight_decay, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(width >= 4 && width % 4 == 0 && row_mask && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adamw_rows: width %% 4 == 0, a row mask and 16-byte aligned tensors required");
  if (n_rows <= 0) return 0;
  int64_t bx = (n_rows * (width / 4) + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  adamw_rows_kernel<<<(unsigned)bx, 256, 0, as_stream(stream)>>>(p, g, m, v, n_rows, width / 4, row_mask, state, lr, beta1, beta2, eps, weight_decay);
  LLMREC_CHECK_LAUNCH("adamw_rows");
  return 0;
}


extern "C" int llmrec_adamw_advance(double* state, double lr, double beta1, double beta2, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  adamw_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(state, lr, beta1, beta2);
  LLMREC_CHECK_LAUNCH("adamw_advance");
  return 0;
}

extern "C" int llmrec_adamw_step_f32(float* const* p, const float* const* g, float* const* m, float* const* v,
                                     const int64_t* numel, int32_t n_tensors, const double* state,
                                     float lr, float beta1, float beta2, float eps, float weight_decay,
                                     llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  cudaStream_t st = as_stream(stream);
  for (int t0 = 0; t0 < n_tensors; t0 += kMaxTensors) {
    AdamParams a{};
    a.n = min(n_tensors - t0, kMaxTensors);
    int64_t mx = 0;
    for (int t = 0; t < a.n; ++t) {
      a.p[t] = p[t0 + t]; a.g[t] = g[t0 + t]; a.m[t] = m[t0 + t]; a.v[t] = v[t0 + t]; a.numel[t] = numel[t0 + t];
      LLMREC_CHECK_ARG(aligned16(a.p[t]) && aligned16(a.g[t]) && aligned16(a.m[t]) && aligned16(a.v[t]), "adamw: tensor %d not 16-byte aligned", t0 + t);
      mx = max(mx, a.numel[t]);
    }
    a.state = state; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay;
    int64_t bx = (mx / 4 + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 148 * 8) bx = 148 * 8;
    adamw_kernel<<<dim3((unsigned)bx, a.n), 256, 0, st>>>(a);
    LLMREC_CHECK_LAUNCH("adamw");
  }
  return 0;
}
