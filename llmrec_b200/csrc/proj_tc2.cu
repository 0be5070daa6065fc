// Projection kernels, second generation: the split A operand lives in TENSOR MEMORY.
//
// proj_tc.cu (v1) keeps A_hi / A_lo in shared memory; ncu showed it bound by shared-memory bandwidth, not HBM
// (profiles/r1_proj_tc_ncu_full.txt: 152 KiB of smem traffic per 16 KiB k-block -- TMA fill, in-place split, and
// three UMMA operand passes over A).  Here the transform warps read each fp32 A tile from shared memory ONCE, split it
// into TF32 hi/lo in registers and store both halves to a TMEM ring with tcgen05.st; the MMAs take A from TMEM
// (tcgen05.mma [d], [a_tmem], b_desc) and only the small B operand (W / dY, 8-16 KiB per k-block) is read from shared
// memory.  Per k-block: 16 KiB TMA fill + 16 KiB transform read + B traffic, i.e. the kernel is back on the HBM roofline.
//
//   forward  Y[n x d] = X W^T + b       A = X tile [128 rows x 32 k]  K-major in smem -> TMEM lane = row, column = k
//   wgrad    dW^T[k x d] = X^T dY       A = X tile [32 rows x 128 feats] (MN-major atoms) -> TMEM lane = feature, column = row:
//                                       the transposition the contraction needs happens in the smem -> TMEM copy, for free.
// Warp roles (512 threads): w0 TMA | w1 MMA | w2 TMEM alloc | w4-7 + w12-15 two transform groups alternating k-blocks |
// w8-11 epilogue; barriers, tile scheduling and epilogues are those of proj_tc.cu.  Supports d <= 128 (TMEM: 2 accumulators of
// d columns + 4 ring slots of 64 columns); larger d falls back to v1.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "proj_tc.cuh"

namespace llmrec {
using namespace tc;

constexpr int kTsMaxStages = 6;   // smem ring == TMEM A ring: 6 slots at d <= 64 (2d + 6*64 <= 512 columns, 6 x 32 KiB smem), else 4
constexpr int kSlotCols = 64;     // hi 32 + lo 32 columns per ring slot

__global__ void __launch_bounds__(512, 1) proj_fwd_ts_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int d = P.d;
  const int stages = P.stages;
  const uint32_t b_bytes = (uint32_t)d * 128u;
  const uint32_t stage_bytes = kTileA + 2u * b_bytes;           // A fp32 | W_hi | W_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars; uint64_t* xform = bars + stages; uint64_t* empty = bars + 2 * stages;
  uint64_t* tfull = bars + 3 * stages; uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  auto sA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto sB = [&](int s) { return sA(s) + kTileA; };
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
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_col0 = 2u * (uint32_t)d;  // [0, 2d) accumulators | [2d, 2d + 4*64) A ring

  auto locate = [&](int tile, int& p, int& mblk) {
    p = 0;
    while (p + 1 < P.n_prob && tile >= P.prob[p + 1].tile_start) ++p;
    mblk = tile - P.prob[p].tile_start;
  };

  if (warp == 0 && lane == 0) {
    PipeState st(stages);
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const int kb_n = P.prob[p].kblocks;
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        mbar_arrive_expect_tx(&full[st.stage], (P.dbg & 1) ? kTileA : stage_bytes);
        tma_load_2d(sA(st.stage), &P.tmA[p], &full[st.stage], kb * BK, mblk * BM);
        if (!(P.dbg & 1)) {
          tma_load_2d(sB(st.stage), &P.tmW[p], &full[st.stage], kb * BK, 0);
          tma_load_2d(sBlo(st.stage), &P.tmW[p], &full[st.stage], kb * BK, d);
        }
        st.advance();
      }
    }
  } else if (warp == 1 && lane == 0) {
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
        mbar_wait(&xform[st.stage], st.phase);
        tc_fence_after();
        const uint32_t b0 = smem_u32(sB(st.stage)), bl0 = smem_u32(sBlo(st.stage));
        const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)(st.stage * kSlotCols), a_lo = a_hi + 32;
        if (!(P.dbg & 2)) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bd = smem_desc_sw128(b0 + kk * 32, 0, 1024);
          const uint64_t bld = smem_desc_sw128(bl0 + kk * 32, 0, 1024);
          umma_tf32_ts(d_tmem, a_lo + kk * 8, bd, idesc, (kb | kk) != 0);   // lo * hi
          umma_tf32_ts(d_tmem, a_hi + kk * 8, bld, idesc, 1);                // hi * lo
          umma_tf32_ts(d_tmem, a_hi + kk * 8, bd, idesc, 1);                 // hi * hi
        }
        }
        umma_commit(&empty[st.stage]);
        st.advance();
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ===== transform: smem A row -> registers -> split -> TMEM (lane = row of the tile) =====
    // two warp-groups (warps 4-7: even k-blocks, warps 12-15: odd k-blocks) so that the LDS -> split -> tcgen05.st
    // latency of one k-block overlaps the next one
    const int grp = warp >= 12 ? 1 : 0;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + a_col0;
    uint32_t it = 0;   // running k-block counter over all tiles of this CTA
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int p, mblk; locate(tile, p, mblk);
      const int kb_n = P.prob[p].kblocks;
      for (int kb = 0; kb < kb_n; ++kb, ++it) {
        if ((int)(it & 1u) != grp) continue;
        const int stage = (int)(it % stages);
        const uint32_t phase = (it / stages) & 1u;
        mbar_wait(&full[stage], phase);
        if (P.dbg & 4) { __syncwarp(); if (lane == 0) mbar_arrive(&xform[stage]); continue; }   // timing experiment: no transform work
        const uint8_t* rowp = sA(stage) + (size_t)row * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // logical 16-byte chunk c sits at physical chunk c ^ (row & 7) (SWIZZLE_128B)
          const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (row & 7)) << 4));
          const float h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
          hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1); hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
          lo[4 * c] = __float_as_uint(v.x - h0); lo[4 * c + 1] = __float_as_uint(v.y - h1);
          lo[4 * c + 2] = __float_as_uint(v.z - h2); lo[4 * c + 3] = __float_as_uint(v.w - h3);
        }
        tmem_st_32x32(lane_base + (uint32_t)(stage * kSlotCols), hi);
        tmem_st_32x32(lane_base + (uint32_t)(stage * kSlotCols + 32), lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform[stage]);
      }
    }
  } else if (warp >= 8 && warp < 12) {
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
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512u); }
}

__global__ void __launch_bounds__(512, 1) proj_wgrad_ts_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int d = P.d;
  const int stages = P.stages;
  const uint32_t b_bytes = (uint32_t)d * 128u;                  // [d/32 atoms][32 rows][128 B]
  const uint32_t stage_bytes = kTileA + 2u * b_bytes;           // X tile fp32 | dY_hi (in place) | dY_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars; uint64_t* xform = bars + stages; uint64_t* empty = bars + 2 * stages;
  uint64_t* tfull = bars + 3 * stages; uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  auto sA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto sB = [&](int s) { return sA(s) + kTileA; };
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
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_col0 = 2u * (uint32_t)d;

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
    // ===== TMA producer (issuing the six boxes of a stage from several lanes in parallel was measured: no gain) =====
    PipeState st(stages);
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        mbar_arrive_expect_tx(&full[st.stage], kTileA + b_bytes);
        const int r = r0 + kb * BK;
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_2d(sA(st.stage) + a * 4096, &P.tmX[p], &full[st.stage], ft * BM + a * 32, r);
        for (int b = 0; b < d / 32; ++b) tma_load_2d(sB(st.stage) + b * 4096, &P.tmG[p], &full[st.stage], b * 32, r);
        st.advance();
      }
    }
  } else if (warp == 1 && lane == 0) {
    PipeState st(stages);
    const uint32_t idesc = idesc_tf32(BM, d, 0, 1);   // A: TMEM (lane = M, column = K); B: MN-major in smem
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * d);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&xform[st.stage], st.phase);
        tc_fence_after();
        const uint32_t b0 = smem_u32(sB(st.stage)), bl0 = smem_u32(sBlo(st.stage));
        const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)(st.stage * kSlotCols), a_lo = a_hi + 32;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 8 rows per MMA: TMEM columns kk*8.., smem B advances 8 rows = 1024 B
          const uint64_t bd = smem_desc_sw128(b0 + kk * 1024, 4096, 512, 1);
          const uint64_t bld = smem_desc_sw128(bl0 + kk * 1024, 4096, 512, 1);
          umma_tf32_ts(d_tmem, a_lo + kk * 8, bd, idesc, (kb | kk) != 0);
          umma_tf32_ts(d_tmem, a_hi + kk * 8, bld, idesc, 1);
          umma_tf32_ts(d_tmem, a_hi + kk * 8, bd, idesc, 1);
        }
        umma_commit(&empty[st.stage]);
        st.advance();
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ===== transform: X tile column (one FEATURE over 32 rows) -> TMEM lane; dY tile split in place =====
    // two warp-groups alternate k-blocks (see the forward kernel)
    const int grp = warp >= 12 ? 1 : 0;
    const int wq = warp & 3;                          // feature atom: features [32*wq, 32*wq + 32)
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + a_col0;
    const int tid = (warp & 3) * 32 + lane;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int p, ft, r0, kb_n; locate(item, p, ft, r0, kb_n);
      for (int kb = 0; kb < kb_n; ++kb, ++it) {
        if ((int)(it & 1u) != grp) continue;
        const int stage = (int)(it % stages);
        const uint32_t phase = (it / stages) & 1u;
        mbar_wait(&full[stage], phase);
        // element (row r, feature e = lane) of atom wq: byte r*128 + (((e >> 3) ^ (r & 3)) << 5) + (e & 7)*4   (SWIZZLE_128B_ATOM_32B)
        const uint8_t* atom = sA(stage) + (size_t)wq * 4096;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float v = *reinterpret_cast<const float*>(atom + r * 128 + ((((lane >> 3) ^ (r & 3)) << 5) | ((lane & 7) << 2)));
          const float h = tf32_hi(v);
          hi[r] = __float_as_uint(h); lo[r] = __float_as_uint(v - h);
        }
        tmem_st_32x32(lane_base + (uint32_t)(stage * kSlotCols), hi);
        tmem_st_32x32(lane_base + (uint32_t)(stage * kSlotCols + 32), lo);
        split_tile_inplace(reinterpret_cast<float4*>(sB(stage)), reinterpret_cast<float4*>(sBlo(stage)), (int)(b_bytes / 16), tid, 128);
        fence_proxy_async_smem();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform[stage]);
      }
    }
  } else if (warp >= 8 && warp < 12) {
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
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512u); }
}

// The ring must cover HBM latency + transform + MMA time of a stage (~2900 clk): 4 stages of 16 KiB sustain only ~3.5 TB/s.
static int ts_stages(int d) { return d <= 64 ? kTsMaxStages : 4; }
static uint32_t ts_smem_bytes(int d) { return (uint32_t)ts_stages(d) * (kTileA + 2u * (uint32_t)d * 128u) + 1024 + 256; }

int proj_fwd_ts_launch(const FwdParams& P0, int grid, cudaStream_t st) {
  FwdParams P = P0;
  P.stages = ts_stages(P.d);
  const uint32_t smem = ts_smem_bytes(P.d);
  cudaFuncSetAttribute(proj_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  proj_fwd_ts_kernel<<<grid, 512, smem, st>>>(P);
  LLMREC_CHECK_LAUNCH("proj_fwd_ts");
  return 0;
}
int proj_wgrad_ts_launch(const WgParams& P0, int grid, cudaStream_t st) {
  WgParams P = P0;
  P.stages = ts_stages(P.d);
  const uint32_t smem = ts_smem_bytes(P.d);
  cudaFuncSetAttribute(proj_wgrad_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  proj_wgrad_ts_kernel<<<grid, 512, smem, st>>>(P);
  LLMREC_CHECK_LAUNCH("proj_wgrad_ts");
  return 0;
}

}  // namespace llmrec
