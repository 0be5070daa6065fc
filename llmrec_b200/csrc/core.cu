// Error state, device check, tiny utilities.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace llmrec {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool device_ok() {
  static int cached = -1;
  if (cached >= 0) return cached == 1;
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    cudaGetLastError();
    cached = 0;
    return false;
  }
  cached = (p.major == 10) ? 1 : 0;
  return cached == 1;
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
}  // namespace llmrec

extern "C" {
int llmrec_abi_version(void) { return LLMREC_ABI_VERSION; }
const char* llmrec_last_error(void) { return llmrec::g_err; }
int llmrec_device_ok(void) { return llmrec::device_ok() ? 1 : 0; }

int llmrec_fill_f32(float* p, int64_t n, float v, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (n <= 0) return 0;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 16) blocks = 148 * 16;
  llmrec::fill_kernel<<<blocks, 256, 0, llmrec::as_stream(stream)>>>(p, n, v);
  LLMREC_CHECK_LAUNCH("fill");
  return 0;
}
}
