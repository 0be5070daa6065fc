// Side-feature projection on the 5th-gen tensor cores (tcgen05, TMEM accumulators, TMA-fed).
//
//   forward  Y[n x d]  = X[n x k] W[d x k]^T + b          (nn.Linear, Models.py:145-150)
//   wgrad    dW[d x k] = dY[n x d]^T X[n x k], db = colsum(dY)   (its autograd; X is a constant feature table)
//
// Both are HBM-bound (32-64 flop/B at d = 64-128): the job of the kernel is to stream X once at HBM
// speed.  All 8 projections of a step run as ONE persistent grouped launch (148 CTAs, static tile
// round-robin over a problem table) -- a single projection has only ~136 row tiles.
//
// Precision: fp32 operands are split  x = hi + lo  (hi = top 19 bits, exactly TF32-representable) by a
// transform warp-group while the tile sits in shared memory, and three kind::tf32 MMAs accumulate
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator ("3xTF32", error ~2^-21 relative per product, i.e.
// fp32-class).  mode 1 skips the split (plain TF32, ~2^-11).
//
// Warp roles (384 threads, 1 CTA/SM):  w0 TMA producer | w1 MMA issuer | w2 TMEM allocator | w4-7 split
// transform | w8-11 epilogue (TMEM -> registers -> global).  Pipelines: smem ring full/xform/empty
// mbarriers, double-buffered TMEM accumulator with tmem_full/tmem_empty mbarriers.
//
// Layouts: forward operands are K-major 128-byte-swizzled tiles ([rows][32 fp32]); the wgrad contraction
// runs over ROWS, so its operands are MN-major tiles ([32-column atom][32 rows][32 fp32], 128B swizzle with
// 32-byte atoms = UMMA layout SWIZZLE_128B_BASE32B / TMA SWIZZLE_128B_ATOM_32B, the one layout tcgen05 takes for
// MN-major tf32) fetched as 32x32 TMA boxes -- no transpose pass over X or dY is ever materialised.
#include <mutex>
#include <stdlib.h>
#include <vector>
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "proj_tc.cuh"

namespace llmrec {

using namespace tc;


// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

bool make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer, bool swizzle32) {
  struct Entry { TmapKey k; CUtensorMap m; };
  static std::vector<Entry> cache;
  static std::mutex mu;
  TmapKey key{base, inner, outer, row_stride_bytes, box_inner, box_outer, swizzle32 ? 1u : 0u, 0u};
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : cache)
    if (memcmp(&e.k, &key, sizeof(key)) == 0) { *out = e.m; return true; }
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner, box_outer); return false; }
  if (cache.size() > 256) cache.clear();
  cache.push_back({key, *out});
  return true;
}

bool make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                      uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return false; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1, stride2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (rank 3) failed (%d) dims=%llu,%llu,%llu strides=%llu,%llu box=%u,%u,%u", (int)r,
                                     (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)stride1,
                                     (unsigned long long)stride2, b0, b1, b2); return false; }
  return true;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(384, 1) proj_fwd_tc_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int d = P.d, stages = P.stages;
  const uint32_t b_bytes = (uint32_t)d * 128u;
  const uint32_t stage_bytes = (kTileA + b_bytes) * (SPLIT ? 2u : 1u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars; uint64_t* xform = bars + stages; uint64_t* empty = bars + 2 * stages;
  uint64_t* tfull = bars + 3 * stages; uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  auto sA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto sAlo = [&](int s) { return sA(s) + kTileA; };
  auto sB = [&](int s) { return sA(s) + kTileA * (SPLIT ? 2u : 1u); };
  auto sBlo = [&](int s) { return sB(s) + b_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_prob; ++p) { prefetch_tmap(&P.tmA[p]); prefetch_tmap(&P.tmW[p]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&xform[s], 4);   /* one arrival per transform warp: 128 per-thread arrivals on one mbarrier serialise (~1000 clk per stage) */ mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto locate = [&](int tile, int& p, int& mblk) {
    p = 0;
    while (p + 1 < P.n_prob && tile >= P.prob[p + 1].tile_start) ++p;
    mblk = tile - P.prob[p].tile_start;
  };

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    PipeState st(stages);
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const int kb_n = P.prob[p].kblocks;
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        mbar_arrive_expect_tx(&full[st.stage], kTileA + b_bytes * (SPLIT ? 2u : 1u));
        const int kbe = kb;
        tma_load_2d(sA(st.stage), &P.tmA[p], &full[st.stage], kbe * BK, mblk * BM);
        tma_load_2d(sB(st.stage), &P.tmW[p], &full[st.stage], kbe * BK, 0);
        if (SPLIT) tma_load_2d(sBlo(st.stage), &P.tmW[p], &full[st.stage], kbe * BK, d);
        st.advance();
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    PipeState st(stages);
    const uint32_t idesc = idesc_tf32(BM, d, 0, 0);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const int kb_n = P.prob[p].kblocks;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * d);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(SPLIT ? &xform[st.stage] : &full[st.stage], st.phase);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA(st.stage)), al0 = smem_u32(sAlo(st.stage));
        const uint32_t b0 = smem_u32(sB(st.stage)), bl0 = smem_u32(sBlo(st.stage));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // UMMA_K = 8 fp32 = 32 bytes inside the 128-byte swizzle row
          const uint64_t ad = smem_desc_sw128(a0 + kk * 32, 0, 1024);
          const uint64_t bd = smem_desc_sw128(b0 + kk * 32, 0, 1024);
          const uint32_t first = (kb | kk) != 0;
          if (SPLIT) {
            const uint64_t ald = smem_desc_sw128(al0 + kk * 32, 0, 1024);
            const uint64_t bld = smem_desc_sw128(bl0 + kk * 32, 0, 1024);
            umma_tf32(d_tmem, ald, bd, idesc, first);   // lo * hi
            umma_tf32(d_tmem, ad, bld, idesc, 1);       // hi * lo
            umma_tf32(d_tmem, ad, bd, idesc, 1);        // hi * hi
          } else {
            umma_tf32(d_tmem, ad, bd, idesc, first);
          }
        }
        umma_commit(&empty[st.stage]);
        st.advance();
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (SPLIT && warp >= 4 && warp < 8) {
    // ===== split transform: A -> A_hi (in place) + A_lo =====
    PipeState st(stages);
    const int tid = threadIdx.x - 128;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const int kb_n = P.prob[p].kblocks;
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&full[st.stage], st.phase);
        split_tile_inplace(reinterpret_cast<float4*>(sA(st.stage)), reinterpret_cast<float4*>(sAlo(st.stage)), kTileA / 16, tid, 128);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform[st.stage]);
        st.advance();
      }
    }
  } else if (warp >= 8) {
    // ===== epilogue: TMEM -> registers -> (+bias) -> global =====
    const int wq = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const FwdProblem pr = P.prob[p];
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = mblk * BM + wq * 32 + lane;
      const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * d);
      float* yrow = pr.Y + (long long)row * pr.ldy;
      int c0 = 0;
      for (; c0 + 32 <= d; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t0 + c0, r);
        tmem_ld_wait();
        if (row < pr.n) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = pr.bias ? __ldg(reinterpret_cast<const float4*>(pr.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            st4(yrow + c0 + j, make_float4(__uint_as_float(r[j]) + b.x, __uint_as_float(r[j + 1]) + b.y,
                                           __uint_as_float(r[j + 2]) + b.z, __uint_as_float(r[j + 3]) + b.w));
          }
        }
      }
      if (c0 < d) {  // d % 32 == 16
        uint32_t r[16];
        tmem_ld_32x16(t0 + c0, r);
        tmem_ld_wait();
        if (row < pr.n) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = pr.bias ? __ldg(reinterpret_cast<const float4*>(pr.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            st4(yrow + c0 + j, make_float4(__uint_as_float(r[j]) + b.x, __uint_as_float(r[j + 1]) + b.y,
                                           __uint_as_float(r[j + 2]) + b.z, __uint_as_float(r[j + 3]) + b.w));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

// W -> [hi ; lo]  ([2d x k], hi exactly TF32-representable); every distinct weight matrix of a grouped launch in ONE launch (blockIdx.y)
struct WsplitParams { const float* W[kMaxProb]; float* out[kMaxProb]; long long n[kMaxProb]; };
__global__ void wsplit_kernel(const WsplitParams P) {
  const float* __restrict__ W = P.W[blockIdx.y];
  float* __restrict__ out = P.out[blockIdx.y];
  const long long n = P.n[blockIdx.y];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = W[i]; const float h = tf32_hi(v); out[i] = h; out[n + i] = v - h;
  }
}

static uint32_t pow2_cols(int c) { uint32_t r = 32; while ((int)r < c) r <<= 1; return r; }

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(384, 1) proj_wgrad_tc_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int d = P.d, stages = P.stages;
  const uint32_t b_bytes = (uint32_t)d * 128u;  // [d/32 atoms][32 rows][128 B]
  const uint32_t stage_bytes = (kTileA + b_bytes) * (SPLIT ? 2u : 1u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars; uint64_t* xform = bars + stages; uint64_t* empty = bars + 2 * stages;
  uint64_t* tfull = bars + 3 * stages; uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  auto sA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto sAlo = [&](int s) { return sA(s) + kTileA; };
  auto sB = [&](int s) { return sA(s) + kTileA * (SPLIT ? 2u : 1u); };
  auto sBlo = [&](int s) { return sB(s) + b_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_prob; ++p) { prefetch_tmap(&P.tmX[p]); prefetch_tmap(&P.tmG[p]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&xform[s], 4);   /* one arrival per transform warp: 128 per-thread arrivals on one mbarrier serialise (~1000 clk per stage) */ mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (problem, feature tile, row range)
  auto locate = [&](int item, int& p, int& ft, int& r0, int& kb_n) {
    p = 0;
    while (p + 1 < P.n_prob && item >= P.prob[p + 1].item_start) ++p;
    const WgProblem pr = P.prob[p];
    const int local = item - pr.item_start;
    const int chunk = local / pr.ft_tiles;
    ft = local - chunk * pr.ft_tiles;
    r0 = chunk * pr.rows_per_chunk;
    const int r1 = min(pr.n, r0 + pr.rows_per_chunk);
    kb_n = (r1 - r0 + BK - 1) / BK;
  };

  if (warp == 0 && lane == 0) {
    PipeState st(stages);
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        mbar_arrive_expect_tx(&full[st.stage], kTileA + b_bytes);
        const int r = r0 + kb * BK;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          tma_load_2d(sA(st.stage) + a * 4096, &P.tmX[p], &full[st.stage], ft * BM + a * 32, r);
        }
        for (int b = 0; b < d / 32; ++b) tma_load_2d(sB(st.stage) + b * 4096, &P.tmG[p], &full[st.stage], b * 32, r);
        st.advance();
      }
    }
  } else if (warp == 1 && lane == 0) {
    PipeState st(stages);
    const uint32_t idesc = idesc_tf32(BM, d, 1, 1);
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * d);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(SPLIT ? &xform[st.stage] : &full[st.stage], st.phase);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA(st.stage)), al0 = smem_u32(sAlo(st.stage));
        const uint32_t b0 = smem_u32(sB(st.stage)), bl0 = smem_u32(sBlo(st.stage));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // UMMA_K = 8 rows = two 4-row x 128 B atoms along K (SBO 512), MN atoms 4096 B apart (LBO)
          const uint64_t ad = smem_desc_sw128(a0 + kk * 1024, 4096, 512, 1);
          const uint64_t bd = smem_desc_sw128(b0 + kk * 1024, 4096, 512, 1);
          const uint32_t first = (kb | kk) != 0;
          if (SPLIT) {
            const uint64_t ald = smem_desc_sw128(al0 + kk * 1024, 4096, 512, 1);
            const uint64_t bld = smem_desc_sw128(bl0 + kk * 1024, 4096, 512, 1);
            umma_tf32(d_tmem, ald, bd, idesc, first);
            umma_tf32(d_tmem, ad, bld, idesc, 1);
            umma_tf32(d_tmem, ad, bd, idesc, 1);
          } else {
            umma_tf32(d_tmem, ad, bd, idesc, first);
          }
        }
        umma_commit(&empty[st.stage]);
        st.advance();
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (SPLIT && warp >= 4 && warp < 8) {
    PipeState st(stages);
    const int tid = threadIdx.x - 128;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&full[st.stage], st.phase);
        split_tile_inplace(reinterpret_cast<float4*>(sA(st.stage)), reinterpret_cast<float4*>(sAlo(st.stage)), kTileA / 16, tid, 128);
        split_tile_inplace(reinterpret_cast<float4*>(sB(st.stage)), reinterpret_cast<float4*>(sBlo(st.stage)), (int)(b_bytes / 16), tid, 128);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform[st.stage]);
        st.advance();
      }
    }
  } else if (warp >= 8) {
    const int wq = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * d);
      float* out = P.partial + ((long long)item * BM + wq * 32 + lane) * d;
      for (int c0 = 0; c0 < d; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t0 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          st4(out + c0 + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

// dW[dc][f] = sum over (problems sharing this dW, in list order) x (row chunks, in order) of partial[item][f % 128][dc]
// One launch for every output.  Block = (output feature tile, group of 4 features): 64 float4 elements x 4 source
// slices; each thread sums every 4th (problem, chunk) source with independent loads in flight, the 4 slices are
// combined in a fixed order -> deterministic and latency-tolerant (the naive per-element loop was latency-bound).
struct ReduceOut { float* dW; int k, n_src, accumulate, blk_start; int src[kMaxProb]; };
struct ReduceParams { ReduceOut out[kMaxProb]; int n_out; int d; const float* partial; WgProblem prob[kMaxProb]; };
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const ReduceParams R) {
  __shared__ float4 part[4][64];
  int o = 0;
  while (o + 1 < R.n_out && (int)blockIdx.x >= R.out[o + 1].blk_start) ++o;
  const ReduceOut ro = R.out[o];
  const int d = R.d, f4_per_feat = d >> 2;                 // d % 32 == 0 on this path
  const int feats_per_blk = 64 / f4_per_feat;              // 4 at d = 64, 2 at d = 128, 1 at d = 256
  const int local = blockIdx.x - ro.blk_start;
  const int groups_per_ft = BM / feats_per_blk;
  const int ft = local / groups_per_ft, fg = local - ft * groups_per_ft;
  const int e = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int f = fg * feats_per_blk + e / f4_per_feat;      // feature inside the 128-feature tile
  const int dc4 = e - (e / f4_per_feat) * f4_per_feat;     // float4 column
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = 0;
  for (int j = 0; j < ro.n_src; ++j) {
    const WgProblem pr = R.prob[ro.src[j]];
#pragma unroll 4
    for (int c = 0; c < pr.chunks; ++c, ++idx) {
      if ((idx & 3) != sl) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(R.partial + ((long long)(pr.item_start + c * pr.ft_tiles + ft) * BM + f) * d) + dc4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[sl][e] = s;
  __syncthreads();
  if (sl == 0) {
    float4 a = part[0][e], b = part[1][e], c = part[2][e], g = part[3][e];
    float4 t = make_float4((a.x + b.x) + (c.x + g.x), (a.y + b.y) + (c.y + g.y), (a.z + b.z) + (c.z + g.z), (a.w + b.w) + (c.w + g.w));
    const int gf = ft * BM + f;
    if (gf < ro.k) {
      float* p = ro.dW + (long long)(dc4 * 4) * ro.k + gf;
      if (ro.accumulate) { t.x += p[0]; t.y += p[ro.k]; t.z += p[2LL * ro.k]; t.w += p[3LL * ro.k]; }
      p[0] = t.x; p[ro.k] = t.y; p[2LL * ro.k] = t.z; p[3LL * ro.k] = t.w;
    }
  }
}

// db[dc] (+)= sum_r dY[r][dc] : kColsumSlices row-slices per problem -> partial; the LAST block to finish (ticket) combines them in a fixed
// order (deterministic; problems sharing one db -- the 5 attribute matrices behind item_trans -- accumulate in problem order).  One launch.
struct ColsumParams { const float* dY[kMaxProb]; long long ld[kMaxProb]; long long n[kMaxProb]; float* db[kMaxProb]; int acc[kMaxProb]; int d; int n_prob; float* partial; unsigned* ticket; };
constexpr int kColsumSlices = 128;
__global__ void __launch_bounds__(256) colsum_kernel(const ColsumParams P) {
  __shared__ float red[256];
  __shared__ bool s_last;
  const int p = blockIdx.y, b = blockIdx.x, d = P.d;
  const int groups = 256 / d > 0 ? 256 / d : 1;  // d <= 256
  const int g = threadIdx.x / d, c = threadIdx.x - g * d;
  float s = 0.f;
  if (g < groups && P.db[p]) {
#pragma unroll 4
    for (long long r = (long long)b * groups + g; r < P.n[p]; r += (long long)kColsumSlices * groups) s += __ldg(P.dY[p] + r * P.ld[p] + c);
  }
  red[threadIdx.x] = (g < groups) ? s : 0.f;
  __syncthreads();
  if (threadIdx.x < d) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += red[gg * d + threadIdx.x];
    P.partial[((long long)p * kColsumSlices + b) * d + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(P.ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int q = 0; q < P.n_prob; ++q) {          // problems in order: shared db accumulate deterministically
    if (!P.db[q]) continue;                     // uniform
    float t = 0.f;
    if (g < groups)
      for (int bb = g; bb < kColsumSlices; bb += groups) t += __ldcg(P.partial + ((long long)q * kColsumSlices + bb) * d + c);
    __syncthreads();
    red[threadIdx.x] = (g < groups) ? t : 0.f;
    __syncthreads();
    if (threadIdx.x < d) {
      float u = 0.f;
      for (int gg = 0; gg < groups; ++gg) u += red[gg * d + threadIdx.x];
      float* o = P.db[q] + threadIdx.x;
      *o = P.acc[q] ? (*o + u) : u;
    }
  }
  if (threadIdx.x == 0) *P.ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// host API (grouped)
// ------------------------------------------------------------------------------------------------
static int stages_for(int d, bool split, uint32_t* smem_bytes) {
  const uint32_t stage = (kTileA + (uint32_t)d * 128u) * (split ? 2u : 1u);
  int s = (int)((200u * 1024u) / stage);
  if (s > 8) s = 8;
  if (s < 2) s = 2;
  *smem_bytes = (uint32_t)s * stage + 1024 /*align slack*/ + 256 /*barriers*/;
  return s;
}

bool proj_tc_supported(int d, int64_t ldx, const void* X, int k, bool wgrad) {
  return d % (wgrad ? 32 : 16) == 0 && d >= 16 && d <= 256 && ldx % 4 == 0 && aligned16(X) && k >= 1 && k % 4 == 0;
}

int proj_fwd_tc_group(const llmrec_proj_fwd_problem* pr, int n_prob, int d, int mode, cudaStream_t st) {
  const bool split = (mode == 0);
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.n_prob = n_prob; P.d = d;
  int tiles = 0;
  WsplitParams WS;
  memset(&WS, 0, sizeof(WS));
  int n_ws = 0;
  long long ws_max = 0;
  for (int p = 0; p < n_prob; ++p) {
    const float* wsrc = split ? pr[p].wsplit : pr[p].W;
    LLMREC_CHECK_ARG(!split || pr[p].wsplit, "proj_fwd: 3xTF32 mode needs a wsplit buffer of 2*d*k floats");
    bool fresh = true;
    for (int q = 0; q < p; ++q) fresh = fresh && !(pr[q].W == pr[p].W && pr[q].wsplit == pr[p].wsplit);
    if (split && fresh) {
      WS.W[n_ws] = pr[p].W; WS.out[n_ws] = pr[p].wsplit; WS.n[n_ws] = (long long)d * pr[p].k;
      ws_max = WS.n[n_ws] > ws_max ? WS.n[n_ws] : ws_max;
      ++n_ws;
    }
    if (!make_tmap_2d_f32(&P.tmA[p], pr[p].X, (uint64_t)pr[p].k, (uint64_t)pr[p].n, (uint64_t)pr[p].ldx * 4, BK, BM)) return 4;
    if (!make_tmap_2d_f32(&P.tmW[p], wsrc, (uint64_t)pr[p].k, (uint64_t)(split ? 2 * d : d), (uint64_t)pr[p].k * 4, BK, (uint32_t)d)) return 4;
    P.prob[p].n = (int)pr[p].n; P.prob[p].k = pr[p].k; P.prob[p].kblocks = (pr[p].k + BK - 1) / BK;
    P.prob[p].tile_start = tiles; P.prob[p].ldy = pr[p].ldy; P.prob[p].Y = pr[p].Y; P.prob[p].bias = pr[p].bias;
    tiles += (int)((pr[p].n + BM - 1) / BM);
  }
  P.total_tiles = tiles;
  if (n_ws > 0) {
    wsplit_kernel<<<dim3((unsigned)((ws_max + 255) / 256), n_ws), 256, 0, st>>>(WS);
    LLMREC_CHECK_LAUNCH("wsplit");
  }
  static const int dbg = getenv("LLMREC_PROJ_DBG") ? atoi(getenv("LLMREC_PROJ_DBG")) : 0;   // TIMING experiments (wrong results): 1 no W loads, 2 no MMAs, 4 no transform
  P.dbg = dbg;
  uint32_t smem;
  P.stages = stages_for(d, split, &smem);
  P.tmem_cols = (int)pow2_cols(2 * d);
  int grid = tiles < 148 ? tiles : 148;
  if (grid <= 0) return 0;
  static const bool force_v1 = getenv("LLMREC_PROJ_V1") != nullptr;
  if (split && d <= 128 && !force_v1) return proj_fwd_ts_launch(P, grid, st);   // A operand from tensor memory (proj_tc2.cu)
  if (split) {
    cudaFuncSetAttribute(proj_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    proj_fwd_tc_kernel<true><<<grid, 384, smem, st>>>(P);
  } else {
    cudaFuncSetAttribute(proj_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    proj_fwd_tc_kernel<false><<<grid, 384, smem, st>>>(P);
  }
  LLMREC_CHECK_LAUNCH("proj_fwd_tc");
  return 0;
}

static int wg_rows_per_chunk(int64_t n) {
  int r = 2048;                       // rows per work item (1024 / 4096 measured: no better)
  while (r > 256 && n / r < 4) r >>= 1;
  return r;
}

int64_t proj_wgrad_tc_scratch(const llmrec_proj_wgrad_problem* pr, int n_prob, int d) {
  int64_t items = 0;
  for (int p = 0; p < n_prob; ++p) {
    int rpc = wg_rows_per_chunk(pr[p].n);
    items += (int64_t)((pr[p].k + BM - 1) / BM) * ((pr[p].n + rpc - 1) / rpc);
  }
  return items * BM * d + (int64_t)n_prob * kColsumSlices * d + 4;   // + the colsum ticket (must start at zero; the kernel re-zeroes it)
}

int proj_wgrad_tc_group(const llmrec_proj_wgrad_problem* pr, int n_prob, int d, int mode, float* scratch, int64_t scratch_elems, cudaStream_t st) {
  const bool split = (mode == 0);
  WgParams P;
  memset(&P, 0, sizeof(P));
  P.n_prob = n_prob; P.d = d;
  int items = 0;
  ColsumParams C;
  memset(&C, 0, sizeof(C));
  C.d = d;
  for (int p = 0; p < n_prob; ++p) {
    if (!make_tmap_2d_f32(&P.tmX[p], pr[p].X, (uint64_t)pr[p].k, (uint64_t)pr[p].n, (uint64_t)pr[p].ldx * 4, 32, BK, true)) return 4;
    if (!make_tmap_2d_f32(&P.tmG[p], pr[p].dY, (uint64_t)d, (uint64_t)pr[p].n, (uint64_t)pr[p].lddy * 4, 32, BK, true)) return 4;
    WgProblem& w = P.prob[p];
    w.n = (int)pr[p].n; w.k = pr[p].k; w.ft_tiles = (pr[p].k + BM - 1) / BM;
    w.rows_per_chunk = wg_rows_per_chunk(pr[p].n);
    w.chunks = (int)((pr[p].n + w.rows_per_chunk - 1) / w.rows_per_chunk);
    w.item_start = items;
    items += w.ft_tiles * w.chunks;
    C.dY[p] = pr[p].dY; C.ld[p] = pr[p].lddy; C.n[p] = pr[p].n; C.db[p] = pr[p].db; C.acc[p] = pr[p].accumulate & LLMREC_WGRAD_ACCUMULATE;
  }
  P.total_items = items;
  const int64_t need = (int64_t)items * BM * d + (int64_t)n_prob * kColsumSlices * d + 4;
  LLMREC_CHECK_ARG(scratch && scratch_elems >= need, "proj_wgrad: scratch too small (%lld < %lld)", (long long)scratch_elems, (long long)need);
  P.partial = scratch + 4;                                   // word 0 of the scratch is the colsum ticket (fixed position for every problem set)
  C.n_prob = n_prob; C.partial = scratch + 4 + (int64_t)items * BM * d;
  C.ticket = reinterpret_cast<unsigned*>(scratch);
  uint32_t smem;
  P.stages = stages_for(d, split, &smem);
  P.tmem_cols = (int)pow2_cols(2 * d);
  int grid = items < 148 ? items : 148;
  if (grid <= 0) return 0;
  // The bias gradients depend on dY only: colsum runs as a BRANCH beside the persistent weight-gradient kernel (1 CTA of 512 threads per SM
  // leaves room for its 256-thread blocks) -- fork/join through events, so inside a stream capture it becomes a parallel graph branch.
  // The side stream and the two events are per device, created on first use (never during the call that is being captured in practice:
  // callers run one eager step first); LLMREC_BRANCHES=0 keeps everything on `st`.
  bool any_db = false;
  for (int p = 0; p < n_prob; ++p) any_db = any_db || pr[p].db != nullptr;
  static const bool branches = !(getenv("LLMREC_BRANCHES") && atoi(getenv("LLMREC_BRANCHES")) == 0);
  struct Branch { cudaStream_t side; cudaEvent_t fork, join; };
  static Branch per_device[64] = {};                      // one side stream + event pair per device the process drives
  int dev = 0;
  LLMREC_CHECK_CUDA(cudaGetDevice(&dev));
  LLMREC_CHECK_ARG(dev >= 0 && dev < 64, "proj_wgrad: device ordinal %d out of range", dev);
  cudaStream_t& side = per_device[dev].side;
  cudaEvent_t& ev_fork = per_device[dev].fork;
  cudaEvent_t& ev_join = per_device[dev].join;
  bool forked = false;
  if (any_db) {
    cudaStream_t cs = st;
    if (branches) {
      if (!side) {
        LLMREC_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        LLMREC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        LLMREC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
      LLMREC_CHECK_CUDA(cudaEventRecord(ev_fork, st));
      LLMREC_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      cs = side; forked = true;
    }
    colsum_kernel<<<dim3(kColsumSlices, n_prob), 256, 0, cs>>>(C);
    LLMREC_CHECK_LAUNCH("colsum");
    if (forked) LLMREC_CHECK_CUDA(cudaEventRecord(ev_join, side));
  }
  static const bool force_v1 = getenv("LLMREC_PROJ_V1") != nullptr;
  if (split && d <= 128 && !force_v1) {
    int rc = proj_wgrad_ts_launch(P, grid, st);
    if (rc) return rc;
  } else if (split) {
    cudaFuncSetAttribute(proj_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    proj_wgrad_tc_kernel<true><<<grid, 384, smem, st>>>(P);
  } else {
    cudaFuncSetAttribute(proj_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    proj_wgrad_tc_kernel<false><<<grid, 384, smem, st>>>(P);
  }
  LLMREC_CHECK_LAUNCH("proj_wgrad_tc");
  ReduceParams R;
  memset(&R, 0, sizeof(R));
  R.d = d; R.partial = scratch + 4;
  int blocks = 0;
  for (int p = 0; p < n_prob; ++p) {
    R.prob[p] = P.prob[p];
    int o = -1;
    for (int q = 0; q < R.n_out; ++q) if (R.out[q].dW == pr[p].dW) o = q;
    if (o < 0) {
      o = R.n_out++;
      R.out[o].dW = pr[p].dW; R.out[o].k = pr[p].k; R.out[o].accumulate = pr[p].accumulate & LLMREC_WGRAD_ACCUMULATE; R.out[o].n_src = 0;
    }
    LLMREC_CHECK_ARG(R.out[o].k == pr[p].k, "proj_wgrad: problems sharing dW must share k");
    R.out[o].src[R.out[o].n_src++] = p;
  }
  const int feats_per_blk = 64 / (d / 4);
  for (int o = 0; o < R.n_out; ++o) { R.out[o].blk_start = blocks; blocks += ((R.out[o].k + BM - 1) / BM) * (BM / feats_per_blk); }
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(R);
  LLMREC_CHECK_LAUNCH("wgrad_reduce");
  if (forked) LLMREC_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  return 0;
}

}  // namespace llmrec
