// Ceiling of the projection kernels' fetch pattern, without any math: a persistent grid streams a feature table X[n x k]
// (fp32) through an S-stage shared-memory ring with TMA, one elected consumer thread just recycles the slots.
//
//   layout rows    2D map over row-major X, box [R rows x 32 floats]: R pieces of 128 B at a pitch of 4k bytes  (what
//                  proj_fwd reads today)
//   layout panels  2D map over the 32-column panel form ([k/32][n][32]): every box is one contiguous run of R*128 B
//   layout bulk    cp.async.bulk (no tensor map) of the same contiguous runs
// Output: one line per configuration with the achieved GB/s (CUDA events, 20 repetitions after 3 warm-ups).
//
//   tools/build_tma_stream.sh && ./tools/tma_stream [n rows = 17366] [tables = 5] [k = 1536]
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"

using namespace llmrec;
using namespace llmrec::tc;

struct StreamParams {
  CUtensorMap tm;
  const float* base;
  long long n;
  int k, rows, stages, layout;   // layout: 0 rows, 1 panels, 2 bulk (panel form)
  int tiles;                     // row tiles = ceil(n / rows)
  int rot;                       // 1: CTA b starts its k loop at block b mod (k/32)
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ StreamParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t box_bytes = (uint32_t)P.rows * 128u;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * box_bytes);
  uint64_t* empty = full + P.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
    prefetch_tmap(&P.tm);
  }
  __syncthreads();
  const int kb_n = P.k / 32;
  if (P.layout >= 3) {
    // proj_wgrad's pattern: work item = (2048-row chunk, 128-feature tile); per stage FOUR [32 rows x 32 floats] boxes
    // (layout 3: row-major X -> 32 x 512 B at the row pitch; layout 4: panel form -> four contiguous 4 KiB runs)
    const int ft_n = P.k / 128, chunks = (int)((P.n + 2047) / 2048);
    if (threadIdx.x == 0) {
      PipeState st(P.stages);
      for (int item = blockIdx.x; item < ft_n * chunks; item += gridDim.x) {
        const int chunk = item / ft_n, ft = item - chunk * ft_n;
        const long long r0 = (long long)chunk * 2048, r1 = r0 + 2048 < P.n ? r0 + 2048 : P.n;
        for (long long r = r0; r < r1; r += 32) {
          mbar_wait(&empty[st.stage], st.phase ^ 1);
          uint8_t* dst = smem + (size_t)st.stage * box_bytes;
          mbar_arrive_expect_tx(&full[st.stage], 16384u);
          for (int a = 0; a < 4; ++a) {
            if (P.layout == 4) tma_load_2d(dst + a * 4096, &P.tm, &full[st.stage], 0, (int)((long long)(ft * 4 + a) * P.n + r));
            else tma_load_2d(dst + a * 4096, &P.tm, &full[st.stage], ft * 128 + a * 32, (int)r);
          }
          st.advance();
        }
      }
    } else if (threadIdx.x == 32) {
      PipeState st(P.stages);
      for (int item = blockIdx.x; item < ft_n * chunks; item += gridDim.x) {
        const int chunk = item / ft_n;
        const long long r0 = (long long)chunk * 2048, r1 = r0 + 2048 < P.n ? r0 + 2048 : P.n;
        for (long long r = r0; r < r1; r += 32) {
          mbar_wait(&full[st.stage], st.phase);
          mbar_arrive(&empty[st.stage]);
          st.advance();
        }
      }
    }
    return;
  }
  if (threadIdx.x == 0) {
    PipeState st(P.stages);
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const long long r0 = (long long)tile * P.rows;
      for (int kb0 = 0; kb0 < kb_n; ++kb0) {
        const int kb = P.rot ? (kb0 + (int)blockIdx.x) % kb_n : kb0;
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        void* dst = smem + (size_t)st.stage * box_bytes;
        if (P.layout == 2) {
          long long rows = P.n - r0 < P.rows ? P.n - r0 : P.rows;
          mbar_arrive_expect_tx(&full[st.stage], (uint32_t)rows * 128u);
          bulk_load_1d(dst, P.base + ((long long)kb * P.n + r0) * 32, (uint32_t)rows * 128u, &full[st.stage]);
        } else {
          mbar_arrive_expect_tx(&full[st.stage], box_bytes);
          if (P.layout == 1) tma_load_2d(dst, &P.tm, &full[st.stage], 0, (int)(kb * P.n + r0));
          else tma_load_2d(dst, &P.tm, &full[st.stage], kb * 32, (int)r0);
        }
        st.advance();
      }
    }
  } else if (threadIdx.x == 32) {
    PipeState st(P.stages);
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x)
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&full[st.stage], st.phase);
        mbar_arrive(&empty[st.stage]);
        st.advance();
      }
  }
}

__global__ void fill_kernel(float* p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (float)(i & 1023) * 1e-3f;
}
__global__ void read_kernel(const float4* __restrict__ p, long long n4, float* out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) { float4 v = __ldg(p + i); acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456f) *out = acc;
}

static float time_ms(cudaStream_t st, int reps, const std::function<void()>& f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b, st);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 17366;
  const int tables = argc > 2 ? atoi(argv[2]) : 5;        // tables of k = 1536 streamed back to back (the 5 attribute tables)
  const int k = argc > 3 ? atoi(argv[3]) : 1536;
  cudaStream_t st;
  cudaStreamCreate(&st);
  const long long elems = n * k;
  std::vector<float*> X(tables);
  for (int t = 0; t < tables; ++t) {
    if (cudaMalloc(&X[t], elems * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    fill_kernel<<<148 * 8, 256, 0, st>>>(X[t], elems);
  }
  float* sink; cudaMalloc(&sink, 4);
  cudaStreamSynchronize(st);
  const double bytes = (double)elems * 4 * tables;
  {
    float ms = time_ms(st, 20, [&] { for (int t = 0; t < tables; ++t) read_kernel<<<148 * 16, 512, 0, st>>>(reinterpret_cast<const float4*>(X[t]), elems / 4, sink); });
    printf("ldg.128 read                         %8.3f ms  %8.1f GB/s\n", ms, bytes / ms * 1e-6);
  }
  const char* names[3] = {"rows  ", "panels", "bulk  "};
  for (int layout = 0; layout < 3; ++layout)
    for (int rows : {64, 128, 256})
      for (int stages : {4, 6, 8, 12}) {
        if ((size_t)stages * rows * 128 > 200 * 1024) continue;
        std::vector<StreamParams> P(tables);
        bool ok = true;
        for (int t = 0; t < tables; ++t) {
          StreamParams& p = P[t];
          p.base = X[t]; p.n = n; p.k = k; p.rows = rows; p.stages = stages; p.layout = layout;
          p.tiles = (int)((n + rows - 1) / rows);
          if (layout == 0) ok = ok && make_tmap_2d_f32(&p.tm, X[t], (uint64_t)k, (uint64_t)n, (uint64_t)k * 4, 32, (uint32_t)rows);
          else ok = ok && make_tmap_2d_f32(&p.tm, X[t], 32, (uint64_t)(k / 32) * (uint64_t)n, 128, 32, (uint32_t)rows);
        }
        if (!ok) { printf("tensor map failed: %s\n", llmrec_last_error()); return 1; }
        const size_t smem = (size_t)stages * rows * 128 + 1024 + 256;
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int ctas : {148, 296}) {
          if (ctas == 296 && smem > 100 * 1024) continue;
          float ms = time_ms(st, 20, [&] { for (int t = 0; t < tables; ++t) stream_kernel<<<ctas, 64, smem, st>>>(P[t]); });
          cudaError_t e = cudaStreamSynchronize(st);
          if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
          printf("%s box %3d rows  stages %2d  ctas %3d  %8.3f ms  %8.1f GB/s\n", names[layout], rows, stages, ctas, ms, bytes / ms * 1e-6);
        }
      }
  // k rotation (LLMREC_PROJ_KROT): same boxes, but concurrent CTAs sit at different feature columns
  for (int layout = 0; layout < 2; ++layout)
    for (int stages : {6, 12}) {
      std::vector<StreamParams> P(tables);
      bool ok = true;
      for (int t = 0; t < tables; ++t) {
        StreamParams& p = P[t];
        p.base = X[t]; p.n = n; p.k = k; p.rows = 128; p.stages = stages; p.layout = layout; p.rot = 1;
        p.tiles = (int)((n + 127) / 128);
        if (layout == 0) ok = ok && make_tmap_2d_f32(&p.tm, X[t], (uint64_t)k, (uint64_t)n, (uint64_t)k * 4, 32, 128);
        else ok = ok && make_tmap_2d_f32(&p.tm, X[t], 32, (uint64_t)(k / 32) * (uint64_t)n, 128, 32, 128);
      }
      if (!ok) { printf("tensor map failed: %s\n", llmrec_last_error()); return 1; }
      const size_t smem = (size_t)stages * 16384 + 1024 + 256;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float ms = time_ms(st, 20, [&] { for (int t = 0; t < tables; ++t) stream_kernel<<<148, 64, smem, st>>>(P[t]); });
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      printf("%s box 128 rows  stages %2d  ctas 148  k-rotated  %8.3f ms  %8.1f GB/s\n", names[layout], stages, ms, bytes / ms * 1e-6);
    }
  // proj_wgrad's fetch pattern (stage = four 32 x 32 boxes = 16 KiB; rows of a tail block past n are zero-filled, so the
  // expected byte count stays 16 KiB)
  for (int layout = 3; layout <= 4; ++layout)
    for (int stages : {4, 6, 8, 12}) {
      std::vector<StreamParams> P(tables);
      bool ok = true;
      for (int t = 0; t < tables; ++t) {
        StreamParams& p = P[t];
        p.base = X[t]; p.n = n; p.k = k; p.rows = 128; p.stages = stages; p.layout = layout; p.tiles = 0;
        if (layout == 3) ok = ok && make_tmap_2d_f32(&p.tm, X[t], (uint64_t)k, (uint64_t)n, (uint64_t)k * 4, 32, 32, true);
        else ok = ok && make_tmap_2d_f32(&p.tm, X[t], 32, (uint64_t)(k / 32) * (uint64_t)n, 128, 32, 32, true);
      }
      if (!ok) { printf("tensor map failed: %s\n", llmrec_last_error()); return 1; }
      const size_t smem = (size_t)stages * 16384 + 1024 + 256;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      for (int ctas : {148, 296}) {
        if (ctas == 296 && smem > 100 * 1024) continue;
        float ms = time_ms(st, 20, [&] { for (int t = 0; t < tables; ++t) stream_kernel<<<ctas, 64, smem, st>>>(P[t]); });
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        printf("%s 4 x [32 x 32] boxes  stages %2d  ctas %3d  %8.3f ms  %8.1f GB/s\n", layout == 3 ? "wgrad rows  " : "wgrad panels", stages, ctas, ms, bytes / ms * 1e-6);
      }
    }
  return 0;
}
