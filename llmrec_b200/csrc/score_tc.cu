// Full-catalog scoring + top-K on the 5th-gen tensor cores (utility/batch_test.py:149-152 scores, :21-36,100-102 ranking).
//
//   score[b, i] = <U[users[b]], I[i]>   for every item i;  items of the user's train row are excluded;
//   top-K by (score desc, item id asc)  ==  heapq.nlargest over ascending candidates.
//
// The reference materialises the [2048 x n_items] score block, copies it to the host and ranks each user in Python.
// Here the score matrix never exists in memory:
//   * CTA = (tile of 128 users, slice of the catalog).  The 128 user rows are gathered ONCE, split into TF32 hi/lo and
//     parked in TENSOR MEMORY (tcgen05.st) as the A operand; item rows (pre-split hi/lo copies of I) stream through a
//     TMA-fed shared-memory ring as the B operand; three kind::tf32 MMAs per K step (lo*hi + hi*lo + hi*hi, fp32
//     accumulate) fill a double-buffered [128 users x 128 items] TMEM accumulator.
//   * the epilogue warps read each accumulator tile with tcgen05.ld (one user row per thread, items in ascending id) and
//     fuse the selection: train-item masking by a merge pointer into the user's sorted train row, threshold test against the
//     thread's current K'-th best, replace-min insertion into a per-thread candidate list in shared memory
//     (K' = K + 16..32 slack).  Ties keep the lower item id (strict > against the minimum; eviction of the largest id
//     among equal minima).
//   * a small exact pass (rescore_topk_kernel) recomputes the K' candidates of every catalog slice in sequential fp32 FMA
//     order -- the same arithmetic as the SIMT reference kernel -- and emits the final top-K by (score desc, id asc), so
//     the 3xTF32 rounding of the selection pass (~1e-5 relative) cannot change the result unless it misranks by more
//     than the slack.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace llmrec {
using namespace tc;

constexpr int SBM = 128;  // users per tile (UMMA M, TMEM lanes)
constexpr int SBN = 128;  // items per accumulator tile (UMMA N)
constexpr int SBK = 32;   // fp32 per 128-byte swizzle row

struct ScoreParams {
  CUtensorMap tmIhi, tmIlo;  // [n_items x d] hi / lo copies of I, box {32, 128}, SWIZZLE_128B
  const float* U; long long ldu;
  const int* users; int n_batch, n_items, d;
  const int* mask_rowptr; const int* mask_col;  // train rows, columns sorted ascending
  int Kc, splits, tiles_per_split, stages, tmem_cols;
  int* cand_idx; float* cand_val;  // [n_batch][splits][Kc]
};

__global__ void __launch_bounds__(384, 1) score_topk_tc_kernel(const __grid_constant__ ScoreParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int d = P.d, stages = P.stages, Kc = P.Kc;
  constexpr uint32_t kTileB = SBN * SBK * 4;  // 16 KiB per hi or lo k-block
  const uint32_t stage_bytes = 2 * kTileB;
  uint8_t* ring = smem;
  float* lv = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);  // heap values [Kc][128]
  int* li = reinterpret_cast<int*>(lv + (size_t)Kc * SBM);                     // heap ids    [Kc][128]
  float* scratch32 = reinterpret_cast<float*>(li + (size_t)Kc * SBM);          // [32][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch32 + 32 * SBM);
  uint64_t* full = bars; uint64_t* empty = bars + stages; uint64_t* tfull = bars + 2 * stages; uint64_t* tempty = tfull + 2;
  uint64_t* a_ready = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int utile = blockIdx.x, split = blockIdx.y;
  const int kb_n = d / SBK;
  const int tile0 = split * P.tiles_per_split;
  const int n_item_tiles = (P.n_items + SBN - 1) / SBN;
  const int tile1 = min(n_item_tiles, tile0 + P.tiles_per_split);

  if (warp == 0 && lane == 0) { prefetch_tmap(&P.tmIhi); prefetch_tmap(&P.tmIlo); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(a_ready, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: [0, 2*SBN) two accumulators | [256, 256+d) A_hi | [256+d, 256+2d) A_lo
  const uint32_t a_hi_col = 2 * SBN, a_lo_col = 2 * SBN + d;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: item k-blocks (hi, lo) =====
    PipeState st(stages);
    for (int t = tile0; t < tile1; ++t) {
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&empty[st.stage], st.phase ^ 1);
        mbar_arrive_expect_tx(&full[st.stage], stage_bytes);
        tma_load_2d(ring + (size_t)st.stage * stage_bytes, &P.tmIhi, &full[st.stage], kb * SBK, t * SBN);
        tma_load_2d(ring + (size_t)st.stage * stage_bytes + kTileB, &P.tmIlo, &full[st.stage], kb * SBK, t * SBN);
        st.advance();
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: A from TMEM, B from the ring =====
    PipeState st(stages);
    const uint32_t idesc = idesc_tf32(SBM, SBN, 0, 0);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = tile0; t < tile1; ++t) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SBN);
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(&full[st.stage], st.phase);
        tc_fence_after();
        const uint32_t bh = smem_u32(ring + (size_t)st.stage * stage_bytes), bl = bh + kTileB;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ah = tmem_base + a_hi_col + (uint32_t)(kb * SBK + kk * 8);
          const uint32_t al = tmem_base + a_lo_col + (uint32_t)(kb * SBK + kk * 8);
          const uint64_t bhd = smem_desc_sw128(bh + kk * 32, 0, 1024);
          const uint64_t bld = smem_desc_sw128(bl + kk * 32, 0, 1024);
          umma_tf32_ts(d_tmem, al, bhd, idesc, (kb | kk) != 0);
          umma_tf32_ts(d_tmem, ah, bld, idesc, 1);
          umma_tf32_ts(d_tmem, ah, bhd, idesc, 1);
        }
        umma_commit(&empty[st.stage]);
        st.advance();
      }
      umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== A loader: gather the user row, split hi/lo, park in TMEM (lane = user) =====
    const int wq = warp & 3;
    const int b = utile * SBM + wq * 32 + lane;
    const float* urow = (b < P.n_batch) ? P.U + (long long)P.users[b] * P.ldu : nullptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    for (int kb = 0; kb < kb_n; ++kb) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = urow ? ldg4(urow + kb * SBK + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        float h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
        hi[j] = __float_as_uint(h0); hi[j + 1] = __float_as_uint(h1); hi[j + 2] = __float_as_uint(h2); hi[j + 3] = __float_as_uint(h3);
        lo[j] = __float_as_uint(v.x - h0); lo[j + 1] = __float_as_uint(v.y - h1); lo[j + 2] = __float_as_uint(v.z - h2); lo[j + 3] = __float_as_uint(v.w - h3);
      }
      tmem_st_32x32(lane_base + a_hi_col + kb * SBK, hi);
      tmem_st_32x32(lane_base + a_lo_col + kb * SBK, lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
  } else if (warp >= 8) {
    // ===== epilogue: fused masking + top-K' selection, one user row per thread =====
    const int wq = warp & 3;
    const int t_in = wq * 32 + lane;  // row inside the tile == TMEM lane
    const int b = utile * SBM + t_in;
    const bool live = b < P.n_batch;
    int mp = 0, mend = 0, next_masked = 0x7fffffff;
    if (live && P.mask_rowptr) {
      const int u = P.users[b];
      mp = P.mask_rowptr[u]; mend = P.mask_rowptr[u + 1];
      next_masked = mp < mend ? __ldg(P.mask_col + mp) : 0x7fffffff;
    }
    // Per-thread binary MIN-heap of the K' best (score, id) seen so far, in shared memory with layout [k][thread]
    // (bank = thread for every k: conflict-free however the heap paths of the 32 lanes diverge).  Heap order: a is
    // "smaller" than b when a.score < b.score, or equal scores and a.id > b.id, so the root is the entry to evict and a
    // new item enters only when strictly better than the root -> ties keep the lower item id.
    float* hv = lv + t_in;
    int* hi_ = li + t_in;
    float* scr = scratch32 + t_in;   // [32][128] staging of one 32-column group
    auto HV = [&](int k) -> float& { return hv[(size_t)k * SBM]; };
    auto HI = [&](int k) -> int& { return hi_[(size_t)k * SBM]; };
    auto lower = [](float av, int ai, float bv, int bi) { return av < bv || (av == bv && ai > bi); };
    int count = 0;
    float thr = -INFINITY;   // root score once the heap is full
    int thr_id = -1;
    auto sift_down = [&](int pos, float v, int id, int n) {
      for (;;) {
        int c = 2 * pos + 1;
        if (c >= n) break;
        float cv = HV(c); int ci = HI(c);
        if (c + 1 < n) {
          const float rv = HV(c + 1); const int ri = HI(c + 1);
          if (lower(rv, ri, cv, ci)) { ++c; cv = rv; ci = ri; }
        }
        if (!lower(cv, ci, v, id)) break;
        HV(pos) = cv; HI(pos) = ci;
        pos = c;
      }
      HV(pos) = v; HI(pos) = id;
    };
    // train-item test on the slow path only: merge pointer into the user's SORTED train row.  Offered items arrive in
    // ascending id, so the pointer only moves forward; the usual case is one register compare, no memory access.
    auto is_masked = [&](int item) -> bool {
      while (next_masked < item) { ++mp; next_masked = mp < mend ? __ldg(P.mask_col + mp) : 0x7fffffff; }
      return next_masked == item;
    };
    auto offer = [&](float s, int item) {
      if (item >= P.n_items || is_masked(item)) return;
      if (count < Kc) {
        HV(count) = s; HI(count) = item;
        if (++count == Kc) {
          for (int p = Kc / 2 - 1; p >= 0; --p) sift_down(p, HV(p), HI(p), Kc);   // heapify once
          thr = HV(0); thr_id = HI(0);
        }
      } else if (s > thr) {   // (equal score, larger id) never displaces: lowest id wins ties
        sift_down(0, s, item, Kc);
        thr = HV(0); thr_id = HI(0);
      }
    };
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = tile0; t < tile1; ++t) {
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * SBN);
#pragma unroll 1
      for (int c0 = 0; c0 < SBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t0 + c0, r);
        tmem_ld_wait();
        // fast path: one compare per score builds the mask of columns that beat the current threshold
        unsigned hit = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(r[j]) > thr ? 1u : 0u) << j;
        if (!live) hit = 0;
        const int item0 = t * SBN + c0;
        if (hit) {
          // slow path: park the 32 scores in this thread's scratch column and walk the set bits in ascending item id with
          // ONE copy of the heap code (32 inlined copies overflowed the instruction cache)
#pragma unroll
          for (int j = 0; j < 32; ++j) scr[(size_t)j * SBM] = __uint_as_float(r[j]);
          unsigned h = hit;
          while (h) {
            const int j = __ffs(h) - 1;
            h &= h - 1;
            const float s = scr[(size_t)j * SBM];
            if (count < Kc || s > thr) offer(s, item0 + j);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    (void)thr_id;
    if (live) {
      int* oi = P.cand_idx + ((long long)b * P.splits + split) * Kc;
      float* ov = P.cand_val + ((long long)b * P.splits + split) * Kc;
      for (int k = 0; k < Kc; ++k) {
        const bool has = k < count;
        oi[k] = has ? HI(k) : -1;
        ov[k] = has ? HV(k) : -INFINITY;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

// I -> hi / lo copies (hi exactly TF32-representable)
__global__ void split_hi_lo_kernel(const float* __restrict__ X, long long ldx, long long n, int d, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const long long r = i / d; const int c = (int)(i - r * d);
  const float v = X[r * ldx + c]; const float h = tf32_hi(v);
  hi[i] = h; lo[i] = v - h;
}

// exact fp32 rescoring of the candidates of one user + final (score desc, id asc) top-K: ONE WARP per user
constexpr int kMaxCandPerLane = 20;   // 32 x 20 = 640 candidates (<= 6 slices x 96)
__global__ void __launch_bounds__(256) rescore_topk_kernel(const float* __restrict__ U, long long ldu, const float* __restrict__ I, long long ldi,
                                                           const int* __restrict__ users, int n_batch, int d, const int* __restrict__ cand_idx,
                                                           int n_cand, int K, int* __restrict__ out_idx, float* __restrict__ out_val) {
  extern __shared__ float sm[];  // 8 warps x d
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + w;
  if (b >= n_batch) return;
  float* us = sm + (size_t)w * d;
  const float* u = U + (long long)users[b] * ldu;
  for (int j = lane; j < d; j += 32) us[j] = u[j];
  __syncwarp();
  const int* ci = cand_idx + (long long)b * n_cand;
  float sc[kMaxCandPerLane]; int id[kMaxCandPerLane];
#pragma unroll
  for (int q = 0; q < kMaxCandPerLane; ++q) {
    const int c = q * 32 + lane;
    int item = c < n_cand ? ci[c] : -1;
    float a = -INFINITY;
    if (item >= 0) {
      const float* it = I + (long long)item * ldi;
      a = 0.f;
      for (int j = 0; j < d; ++j) a = fmaf(us[j], it[j], a);   // same order as score_rows_kernel (score_simt.cu)
    }
    sc[q] = a; id[q] = item;
  }
  for (int r = 0; r < K; ++r) {
    float best = -INFINITY; int besti = 0x7fffffff, bestq = -1;
#pragma unroll
    for (int q = 0; q < kMaxCandPerLane; ++q) {
      if (id[q] >= 0 && sc[q] != -INFINITY && (sc[q] > best || (sc[q] == best && id[q] < besti))) { best = sc[q]; besti = id[q]; bestq = q; }
    }
    float wb = best; int wi = besti;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wb, o); const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov > wb || (ov == wb && oi < wi)) { wb = ov; wi = oi; }
    }
    const bool ok = wi != 0x7fffffff;
    if (lane == 0) {
      out_idx[(long long)b * K + r] = ok ? wi : -1;
      if (out_val) out_val[(long long)b * K + r] = ok ? wb : -INFINITY;
    }
    if (ok && bestq >= 0 && besti == wi) {   // the owning lane retires the winner (ids are unique per user)
#pragma unroll
      for (int q = 0; q < kMaxCandPerLane; ++q) if (q == bestq) id[q] = -1;
    }
  }
}

bool score_tc_supported(int d, int K, long long ldu, long long ldi, const void* U, const void* I) {
  return (d == 32 || d == 64 || d == 96 || d == 128) && K <= 64 && ldu % 4 == 0 && ldi % 4 == 0 && aligned16(U) && aligned16(I);
}

static int score_kc(int K) { int kc = ((K + 16 + 7) / 8) * 8; return kc > 96 ? 96 : kc; }

static void score_plan(int n_batch, int n_items, int K, int* splits, int* tiles_per_split) {
  const int utiles = (n_batch + SBM - 1) / SBM;
  const int itiles = (n_items + SBN - 1) / SBN;
  int s = utiles >= 74 ? 1 : (148 + utiles - 1) / utiles;   // cut the catalog only when user tiles cannot fill the SMs
  const int max_s = itiles / 32 > 0 ? itiles / 32 : 1;   // keep >= 4096 items per slice
  if (s > max_s) s = max_s;
  if (s > 6) s = 6;       // rescore_topk_kernel holds at most 32 x 20 candidates per user
  if (s < 1) s = 1;
  *tiles_per_split = (itiles + s - 1) / s;
  *splits = (itiles + *tiles_per_split - 1) / *tiles_per_split;
  (void)K;
}

long long score_tc_scratch(int n_batch, int n_items, int d, int K) {
  int splits, tps;
  score_plan(n_batch, n_items, K, &splits, &tps);
  return 2LL * n_items * d + 2LL * n_batch * splits * score_kc(K) + 64;
}

int score_topk_tc(const float* U, long long ldu, const float* I, long long ldi, const int* users, int n_batch, int n_items, int d,
                  const int* mask_rowptr, const int* mask_col, int K, int* out_idx, float* out_val, float* scratch, long long scratch_elems,
                  cudaStream_t st) {
  LLMREC_CHECK_ARG(scratch && scratch_elems >= score_tc_scratch(n_batch, n_items, d, K), "score_topk: scratch too small");
  ScoreParams P;
  memset(&P, 0, sizeof(P));
  score_plan(n_batch, n_items, K, &P.splits, &P.tiles_per_split);
  P.Kc = score_kc(K);
  float* Ihi = scratch; float* Ilo = scratch + (long long)n_items * d;
  float* cval = Ilo + (long long)n_items * d;
  int* cidx = reinterpret_cast<int*>(cval + (long long)n_batch * P.splits * P.Kc);
  {
    const long long n = (long long)n_items * d;
    split_hi_lo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(I, ldi, n_items, d, Ihi, Ilo);
    LLMREC_CHECK_LAUNCH("split_hi_lo");
  }
  if (!make_tmap_2d_f32(&P.tmIhi, Ihi, (uint64_t)d, (uint64_t)n_items, (uint64_t)d * 4, SBK, SBN)) return 4;
  if (!make_tmap_2d_f32(&P.tmIlo, Ilo, (uint64_t)d, (uint64_t)n_items, (uint64_t)d * 4, SBK, SBN)) return 4;
  P.U = U; P.ldu = ldu; P.users = users; P.n_batch = n_batch; P.n_items = n_items; P.d = d;
  P.mask_rowptr = mask_rowptr; P.mask_col = mask_col; P.cand_idx = cidx; P.cand_val = cval;
  P.tmem_cols = 512;
  const size_t list_bytes = (size_t)P.Kc * SBM * 8 + 32 * SBM * 4;
  int stages = (int)((225 * 1024 - list_bytes - 512) / (2 * 16384));
  if (stages > 6) stages = 6;
  LLMREC_CHECK_ARG(stages >= 2, "score_topk: not enough shared memory for the pipeline");
  P.stages = stages;
  const size_t smem = (size_t)stages * 2 * 16384 + list_bytes + 512 + 1024;
  cudaFuncSetAttribute(score_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((n_batch + SBM - 1) / SBM, P.splits);
  score_topk_tc_kernel<<<grid, 384, smem, st>>>(P);
  LLMREC_CHECK_LAUNCH("score_topk_tc");
  const int n_cand = P.splits * P.Kc;
  LLMREC_CHECK_ARG(n_cand <= 32 * kMaxCandPerLane, "score_topk: %d candidates per user exceed the rescoring capacity", n_cand);
  rescore_topk_kernel<<<(n_batch + 7) / 8, 256, (size_t)8 * d * sizeof(float), st>>>(U, ldu, I, ldi, users, n_batch, d, cidx, n_cand, K, out_idx, out_val);
  LLMREC_CHECK_LAUNCH("rescore_topk");
  return 0;
}

}  // namespace llmrec
