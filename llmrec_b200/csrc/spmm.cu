// Propagation SpMM over a CSR pattern with row/column scaling (Models.py:57-61,152-183).
//
//   Y_s[r,:] = epi( rs[r] * sum_{e in row r} v[e] * cs[col[e]] * X_s[col[e],:] ) + Z_s[r,:]
//
// HBM/L2-bound gather -- index and byte work, no tensor cores.  Design (sm_100a):
//   * CSR-row-TILED and nnz-flattened: the host plans tiles of <= tile_nnz non-zeros (llmrec_spmm_plan_tiles):
//     a tile is either a run of up to 15 consecutive complete rows, or one piece of a long row.  One
//     lane-group (8/16/32 lanes, by operand width) owns a tile: ONE coalesced read of the tile's row
//     pointers, ONE coalesced read of its (col, v, cs) stream, then the gathers of ALL its rows are issued
//     back to back, U at a time -- the dependent-load chain per row (rowptr -> col -> X) that makes
//     warp-per-row kernels latency-bound on low-degree graphs is paid once per tile, and power-law rows
//     no longer serialise on one warp.  Long-row pieces write raw partial sums; a second pass adds them
//     in fixed order (deterministic, no atomics).
//   * every dense operand sharing the pattern ("segment") is gathered in the same pass, so the index
//     stream is read once for the image/text/attribute/profile/ID operands: the reference's 20 forward
//     SpMMs are 4 launches.  Wide concatenations are cut into column windows over blockIdx.y so that
//     each lane keeps at most CH 16-byte chunks (more warps in flight, fewer registers).
//   * neighbour rows come in with 128-bit read-only loads, U x CH in flight per lane; row scale,
//     optional row-softmax over a segment's d columns (lane-group shuffle reduction), optional addend
//     and the store are fused at the row boundary.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace llmrec {

constexpr int kTicketWindows = 64;   // column windows (blockIdx.y) a launch can have: 16 segments x 128 floats / 32 lanes
struct SpmmParams {
  const int* rowptr; const int* col; const float* vals; const float* rs; const float* cs;
  int n_rows; int d; int nseg; int f4_per_seg; int total_f4; int any_softmax;
  const int4* tiles; int n_tiles; int n_split_tiles; float* scratch;
  const int* split_row; const int* split_first; int n_split;
  int* split_tickets;         // optional int32[n_split * kTicketWindows]: the LAST piece of a long row to finish adds the pieces up (no second launch)
  const unsigned* src_mask;   // optional bitmask over SOURCE rows (columns of the pattern): clear bit = row known to be zero, never fetched
  const int* rows; const int* n_rows_dev; int max_rows;   // row-list form (spmm_rows_kernel)
  llmrec_spmm_seg seg[LLMREC_MAX_SEG];
};
__device__ __forceinline__ bool src_active(const unsigned* m, int c) { return (__ldg(m + (c >> 5)) >> (c & 31)) & 1u; }

template <int CH>
struct LaneChunks {
  const float* xb[CH];
  float* yb[CH];
  const float* zb[CH];
  int64_t ldx[CH], ldy[CH], ldz[CH];
  int flags[CH];
  bool on[CH];
};

template <int LPR, int CH>
__device__ __forceinline__ void setup_chunks(const SpmmParams& p, int q0, int lane_in, LaneChunks<CH>& lc) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int q = q0 + c * LPR + lane_in;
    bool on = q < p.total_f4;
    int s = on ? q / p.f4_per_seg : 0;
    int off = on ? (q - s * p.f4_per_seg) * 4 : 0;
    lc.on[c] = on;
    lc.xb[c] = p.seg[s].X + off;
    lc.yb[c] = p.seg[s].Y + off;
    lc.zb[c] = p.seg[s].Z ? p.seg[s].Z + off : nullptr;
    lc.ldx[c] = p.seg[s].ldx;
    lc.ldy[c] = p.seg[s].ldy;
    lc.ldz[c] = p.seg[s].ldz;
    lc.flags[c] = on ? p.seg[s].flags : 0;
  }
}

// scale + (softmax) + (addend) + store for one output row; only the lanes of `gmask` take part
template <int CH>
__device__ __forceinline__ void finish_row(const SpmmParams& p, const LaneChunks<CH>& lc, float4 (&acc)[CH], int row, unsigned gmask) {
  const float s = p.rs ? __ldg(p.rs + row) : 1.0f;
  const int G = p.f4_per_seg;  // lanes per segment (power of two <= group size whenever a softmax flag is set)
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float4 a = acc[c];
    a.x *= s; a.y *= s; a.z *= s; a.w *= s;
    if (p.any_softmax) {  // launch-uniform: every lane of the group runs the shuffles
      float m = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
      for (int o = G >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(gmask, m, o));
      float4 e = make_float4(expf(a.x - m), expf(a.y - m), expf(a.z - m), expf(a.w - m));
      float t = (e.x + e.y) + (e.z + e.w);
      for (int o = G >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(gmask, t, o);
      if (lc.flags[c] & LLMREC_SPMM_SOFTMAX) {
        float inv = 1.0f / t;
        a = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
      }
    }
    if (lc.on[c]) {
      if (lc.zb[c]) {
        float4 o = *reinterpret_cast<const float4*>(lc.zb[c] + (int64_t)row * lc.ldz[c]);
        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      }
      st4(lc.yb[c] + (int64_t)row * lc.ldy[c], a);
    }
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int LPR, int CH, int U, int MINB = 4>
__global__ void __launch_bounds__(256, MINB) spmm_tile_kernel(const SpmmParams p) {
  constexpr int RPW = 32 / LPR;  // tiles per warp
  const int lane = threadIdx.x & 31;
  const int lane_in = lane % LPR;
  const int sub = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int tile = ((blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5)) * RPW + sub;
  if (tile >= p.n_tiles) return;  // whole lane-group leaves together

  LaneChunks<CH> lc;
  setup_chunks<LPR, CH>(p, blockIdx.y * (LPR * CH), lane_in, lc);

  // 32-byte tile descriptor: {first row, #complete rows (0 = piece of a long row), e0, e1} + 16 one-byte row-end
  // deltas (rowptr[row0+i+1] - e0): no dependent rowptr read, the row boundaries travel with the descriptor.
  const int4 t = __ldg(p.tiles + 2 * tile);
  const uint4 dl = __ldg(reinterpret_cast<const uint4*>(p.tiles + 2 * tile + 1));
  const int row0 = t.x, nrows = t.y, e0 = t.z, e1 = t.w;
  const bool piece = nrows == 0;
  auto row_end_of = [&](int i) -> int {  // end offset of local row i-1 == rowptr[row0 + i]
    const int b = i - 1;
    const unsigned wsel = (b >> 2) == 0 ? dl.x : ((b >> 2) == 1 ? dl.y : ((b >> 2) == 2 ? dl.z : dl.w));
    return e0 + (int)((wsel >> ((b & 3) * 8)) & 0xffu);
  };
  int cur = 0;
  int row_end = piece ? e1 : row_end_of(1);

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int len = e1 - e0;
  for (int k0 = 0; k0 < len; k0 += LPR) {
    const int e = e0 + k0 + lane_in;
    int cidx = 0;
    float w = 0.f;
    if (e < e1) {
      cidx = __ldg(p.col + e);
      w = p.vals ? __ldg(p.vals + e) : 1.0f;
      if (p.cs) w *= __ldg(p.cs + cidx);
      if (p.src_mask && !src_active(p.src_mask, cidx)) w = 0.f;   // inactive source row: contributes exactly zero, skip its fetch
    }
    const int cnt = min(LPR, len - k0);
    for (int j0 = 0; j0 < cnt; j0 += U) {
      float4 x[U][CH];
      float wj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = (j0 + u) & (LPR - 1);
        const int cj = __shfl_sync(gmask, cidx, jj, LPR);
        wj[u] = __shfl_sync(gmask, w, jj, LPR);
        const bool live = j0 + u < cnt && (!p.src_mask || wj[u] != 0.f);
#pragma unroll
        for (int c = 0; c < CH; ++c)
          x[u][c] = (live && lc.on[c]) ? ldg4(lc.xb[c] + (int64_t)cj * lc.ldx[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u < cnt) {
          const int ge = e0 + k0 + j0 + u;
          while (!piece && ge >= row_end) {  // row boundary (group-uniform); also steps over empty rows
            finish_row<CH>(p, lc, acc, row0 + cur, gmask);
            ++cur;
            row_end = row_end_of(cur + 1);
          }
#pragma unroll
          for (int c = 0; c < CH; ++c) fma4(acc[c], wj[u], x[u][c]);
        }
      }
    }
  }
  if (piece) {
    // raw partial sums of one piece of a long row (pieces are numbered first: slot == tile id)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int q = blockIdx.y * (LPR * CH) + c * LPR + lane_in;
      if (q < p.total_f4) st4(p.scratch + ((int64_t)tile * p.total_f4 + q) * 4, acc[c]);
    }
    if (p.split_tickets) {
      // the descriptor's spare words carry {split index, first piece, #pieces}: whoever finishes LAST adds the pieces in piece order
      // (the same order as the second-pass kernel: bit-identical sums) and runs the row epilogue -- no spmm_finish launch
      const int sidx = (int)dl.x, first = (int)dl.y, npieces = (int)dl.z;
      __threadfence();
      __syncwarp(gmask);
      int ticket = 0;
      int* cnt = p.split_tickets + (size_t)sidx * kTicketWindows + blockIdx.y;
      if (lane_in == 0) ticket = atomicAdd(cnt, 1);
      ticket = __shfl_sync(gmask, ticket, 0, LPR);
      if (ticket == npieces - 1) {
        __threadfence();
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t2 = first; t2 < first + npieces; ++t2) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int q = blockIdx.y * (LPR * CH) + c * LPR + lane_in;
            if (q < p.total_f4) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(p.scratch + ((int64_t)t2 * p.total_f4 + q) * 4));
              acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
            }
          }
        }
        finish_row<CH>(p, lc, acc, row0, gmask);
        if (lane_in == 0) *cnt = 0;
      }
    }
  } else {
    for (; cur < nrows; ++cur) finish_row<CH>(p, lc, acc, row0 + cur, gmask);  // last row and trailing empty rows
  }
}

// TMA-STAGED form of the tile kernel (d = 128, one segment): the neighbour rows of a tile do not travel through registers -- each
// gathered row X[col[e], :] (512 B, contiguous) is ONE bulk async copy (cp.async.bulk.shared.global, SASS UBLKCP) into a per-warp ring in
// shared memory, 8 rows per stage and kBulkStages stages per warp, each stage guarded by an mbarrier armed with the bytes it expects;
// the warp then reads the staged rows conflict-free (lane = 16-byte chunk) and accumulates with the same row-boundary / epilogue logic
// as the register kernel.  Per SM up to 12 warps x 3 stages x 4 KiB = 144 KiB of gathers are in flight without holding a register,
// which is what a DRAM-latency-bound random gather wants.  Same tile descriptors, same sums in the same order (bit-identical results).
constexpr bool kSpmmBulkAuto = true;    // A/B at the synthetic scale (profiles/r2_spmm_ab.txt): 12 % faster than register gathers when the gathered table is far
                                        // larger than L2 (5.1 GB user table, DRAM-bound), 10 % slower on the 512 MB item table (46 % L2 hits) -> size threshold below
constexpr int kBulkStages = 3;
constexpr int kBulkRows = 8;
constexpr int kBulkWarps = 8;   // 8 warps x 3 stages x 4 KiB = 96 KiB per CTA, 2 CTAs per SM
__device__ __forceinline__ void bulk_row_load(void* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(kBulkWarps * 32, 2) spmm_bulk_kernel(const SpmmParams p) {
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  constexpr int ROWB = 512;                                        // d = 128 fp32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = bulk_smem + (size_t)warp * kBulkStages * kBulkRows * ROWB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bulk_smem + (size_t)kBulkWarps * kBulkStages * kBulkRows * ROWB) + warp * kBulkStages;
  if (lane == 0) {
    for (int s = 0; s < kBulkStages; ++s) tc::mbar_init(&bars[s], 1);
    tc::fence_barrier_init();
  }
  __syncwarp();
  const int tile = blockIdx.x * kBulkWarps + warp;
  if (tile >= p.n_tiles) return;
  LaneChunks<1> lc;
  setup_chunks<32, 1>(p, 0, lane, lc);
  const int4 t = __ldg(p.tiles + 2 * tile);
  const uint4 dl = __ldg(reinterpret_cast<const uint4*>(p.tiles + 2 * tile + 1));
  const int row0 = t.x, nrows = t.y, e0 = t.z, e1 = t.w;
  const bool piece = nrows == 0;
  auto row_end_of = [&](int i) -> int {
    const int b = i - 1;
    const unsigned wsel = (b >> 2) == 0 ? dl.x : ((b >> 2) == 1 ? dl.y : ((b >> 2) == 2 ? dl.z : dl.w));
    return e0 + (int)((wsel >> ((b & 3) * 8)) & 0xffu);
  };
  int cur = 0;
  int row_end = piece ? e1 : row_end_of(1);
  float4 acc[1] = {make_float4(0.f, 0.f, 0.f, 0.f)};
  const int len = e1 - e0;
  const int n_groups = (len + kBulkRows - 1) / kBulkRows;
  const float* X = p.seg[0].X;
  const int64_t ldx = p.seg[0].ldx;
  // column / weight stream in 32-entry chunks: chunk c lives in (cidx[c & 1], w[c & 1]); the issue side runs < 32 entries ahead
  int cidx0 = 0, cidx1 = 0, cidx2 = 0;
  float wv0 = 0.f, wv1 = 0.f, wv2 = 0.f;
  auto load_chunk = [&](int c) {                                   // chunk c -> register slot c % 3; issued ONE chunk ahead of its first use
    if (c * 32 >= len) return;
    const int e = e0 + c * 32 + lane;
    int ci = 0; float w = 0.f;
    if (e < e1) {
      ci = __ldg(p.col + e);
      w = p.vals ? __ldg(p.vals + e) : 1.0f;
      if (p.cs) w *= __ldg(p.cs + ci);
    }
    const int sl = c % 3;
    if (sl == 0) { cidx0 = ci; wv0 = w; } else if (sl == 1) { cidx1 = ci; wv1 = w; } else { cidx2 = ci; wv2 = w; }
  };
  auto cidx_of = [&](int c) { const int sl = c % 3; return sl == 0 ? cidx0 : (sl == 1 ? cidx1 : cidx2); };
  auto w_of = [&](int c) { const int sl = c % 3; return sl == 0 ? wv0 : (sl == 1 ? wv1 : wv2); };
  load_chunk(0);
  auto issue = [&](int g) {                                        // stage g % S <- rows of edges [8g, 8g + 8)
    if ((g & 3) == 0) load_chunk((g >> 2) + 1);                    // prefetch the NEXT chunk: its latency hides behind 4 groups of copies
    const int s = g % kBulkStages;
    const int k0 = g * kBulkRows;
    const int nvalid = min(kBulkRows, len - k0);
    const int cj = __shfl_sync(0xffffffffu, cidx_of(g >> 2), (k0 + (lane & 7)) & 31);
    if (lane == 0) tc::mbar_arrive_expect_tx(&bars[s], (uint32_t)nvalid * ROWB);
    __syncwarp();
    if (lane < nvalid) bulk_row_load(ring + ((size_t)s * kBulkRows + lane) * ROWB, X + (int64_t)cj * ldx, ROWB, &bars[s]);
  };
  const int pre = min(n_groups, kBulkStages - 1);
  for (int g = 0; g < pre; ++g) issue(g);
  for (int g = 0; g < n_groups; ++g) {
    if (g + kBulkStages - 1 < n_groups) issue(g + kBulkStages - 1);
    const int s = g % kBulkStages;
    tc::mbar_wait(&bars[s], (uint32_t)((g / kBulkStages) & 1));
    const int k0 = g * kBulkRows;
    const int nvalid = min(kBulkRows, len - k0);
    const float wmine = w_of(g >> 2);
#pragma unroll
    for (int j = 0; j < kBulkRows; ++j) {
      if (j < nvalid) {
        const float wj = __shfl_sync(0xffffffffu, wmine, (k0 + j) & 31);
        const int ge = e0 + k0 + j;
        while (!piece && ge >= row_end) {
          finish_row<1>(p, lc, acc, row0 + cur, 0xffffffffu);
          ++cur;
          row_end = row_end_of(cur + 1);
        }
        const float4 x = *reinterpret_cast<const float4*>(ring + ((size_t)s * kBulkRows + j) * ROWB + lane * 16);
        fma4(acc[0], wj, x);
      }
    }
    __syncwarp();                                                  // every lane has read stage s before it is refilled
  }
  if (piece) {
    if (lane < p.total_f4) st4(p.scratch + ((int64_t)tile * p.total_f4 + lane) * 4, acc[0]);
  } else {
    for (; cur < nrows; ++cur) finish_row<1>(p, lc, acc, row0 + cur, 0xffffffffu);
  }
}

// Row-LIST form: only the rows named in a device list are computed (and written); every other output row is left untouched.
// Used where a training step provably needs a small subset of a product's rows (dist.py "demand" mode: the last propagation layer
// is consumed on the batch's neighbourhood only).  Persistent grid, one lane-group per listed row, the list length is read from
// device memory (no host sync).  One segment; d/4 lanes per row (d in {32, 64, 128}).
template <int LPR, int U>
__global__ void __launch_bounds__(256) spmm_rows_kernel(const SpmmParams p) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, lane_in = lane % LPR, sub = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  int n = __ldg(p.n_rows_dev);
  n = n < p.max_rows ? n : p.max_rows;
  const llmrec_spmm_seg sg = p.seg[0];
  const int off = lane_in * 4;
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * RPW;
  for (long long idx = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub; idx < n; idx += stride) {
    const int row = __ldg(p.rows + idx);
    if (row < 0) continue;
    const int e0 = __ldg(p.rowptr + row), e1 = __ldg(p.rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = e0; k0 < e1; k0 += LPR) {
      const int e = k0 + lane_in;
      int cidx = 0; float w = 0.f;
      if (e < e1) {
        cidx = __ldg(p.col + e);
        w = p.vals ? __ldg(p.vals + e) : 1.0f;
        if (p.cs) w *= __ldg(p.cs + cidx);
        if (p.src_mask && !src_active(p.src_mask, cidx)) w = 0.f;
      }
      const int cnt = min(LPR, e1 - k0);
      for (int j0 = 0; j0 < cnt; j0 += U) {
        float4 x[U]; float wj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = (j0 + u) & (LPR - 1);
          const int cj = __shfl_sync(gmask, cidx, jj, LPR);
          wj[u] = __shfl_sync(gmask, w, jj, LPR);
          const bool live = j0 + u < cnt && (!p.src_mask || wj[u] != 0.f);
          x[u] = live ? ldg4(sg.X + (int64_t)cj * sg.ldx + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) if (j0 + u < cnt) fma4(acc, wj[u], x[u]);
      }
    }
    const float sc = p.rs ? __ldg(p.rs + row) : 1.0f;
    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
    if (sg.flags & LLMREC_SPMM_SOFTMAX) {
      float m = fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w));
      for (int o = LPR >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(gmask, m, o));
      float4 ev = make_float4(expf(acc.x - m), expf(acc.y - m), expf(acc.z - m), expf(acc.w - m));
      float t = (ev.x + ev.y) + (ev.z + ev.w);
      for (int o = LPR >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(gmask, t, o);
      const float inv = 1.0f / t;
      acc = make_float4(ev.x * inv, ev.y * inv, ev.z * inv, ev.w * inv);
    }
    if (sg.Z) {
      const float4 z = *reinterpret_cast<const float4*>(sg.Z + (int64_t)row * sg.ldz + off);
      acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
    }
    st4(sg.Y + (int64_t)row * sg.ldy + off, acc);
  }
}

// Row-list form for SHORT lists of possibly very LONG rows (the 2B' batch items of a step: a hub item has 1e5 neighbours): one CTA of 8
// warps per listed row, warp w takes the row's 32-edge chunks w, w+8, ..; the 8 partial rows are added in warp order through shared
// memory (deterministic), then the same epilogue.  d = 128 (lane = 16-byte chunk), one segment.
__global__ void __launch_bounds__(256) spmm_rows_cta_kernel(const SpmmParams p) {
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int n = __ldg(p.n_rows_dev);
  n = n < p.max_rows ? n : p.max_rows;
  const llmrec_spmm_seg sg = p.seg[0];
  const int f4 = p.d >> 2;                                   // 8, 16 or 32 chunks per row; lanes >= f4 idle in the gathers
  for (int idx = blockIdx.x; idx < n; idx += gridDim.x) {
    const int row = __ldg(p.rows + idx);
    if (row < 0) continue;
    const int e0 = __ldg(p.rowptr + row), e1 = __ldg(p.rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = e0 + w * 32; k0 < e1; k0 += 8 * 32) {
      const int e = k0 + lane;
      int cidx = 0; float wt = 0.f;
      if (e < e1) {
        cidx = __ldg(p.col + e);
        wt = p.vals ? __ldg(p.vals + e) : 1.0f;
        if (p.cs) wt *= __ldg(p.cs + cidx);
        if (p.src_mask && !src_active(p.src_mask, cidx)) wt = 0.f;
      }
      const int cnt = min(32, e1 - k0);
      for (int j0 = 0; j0 < cnt; j0 += 4) {
        float4 x[4]; float wj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cj = __shfl_sync(0xffffffffu, cidx, (j0 + u) & 31);
          wj[u] = __shfl_sync(0xffffffffu, wt, (j0 + u) & 31);
          const bool live = j0 + u < cnt && lane < f4 && (!p.src_mask || wj[u] != 0.f);
          x[u] = live ? ldg4(sg.X + (int64_t)cj * sg.ldx + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j0 + u < cnt) fma4(acc, wj[u], x[u]);
      }
    }
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0) {
      float4 a = part[0][lane];
#pragma unroll
      for (int q = 1; q < 8; ++q) { const float4 b = part[q][lane]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      const float sc = p.rs ? __ldg(p.rs + row) : 1.0f;
      a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
      if (sg.flags & LLMREC_SPMM_SOFTMAX) {
        float m = lane < f4 ? fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)) : -INFINITY;
        m = warp_max(m);
        float4 ev = make_float4(expf(a.x - m), expf(a.y - m), expf(a.z - m), expf(a.w - m));
        float t = lane < f4 ? (ev.x + ev.y) + (ev.z + ev.w) : 0.f;
        t = warp_sum(t);
        const float inv = 1.0f / t;
        a = make_float4(ev.x * inv, ev.y * inv, ev.z * inv, ev.w * inv);
      }
      if (lane < f4) {
        if (sg.Z) { const float4 z = *reinterpret_cast<const float4*>(sg.Z + (int64_t)row * sg.ldz + lane * 4); a.x += z.x; a.y += z.y; a.z += z.z; a.w += z.w; }
        st4(sg.Y + (int64_t)row * sg.ldy + lane * 4, a);
      }
    }
    __syncthreads();
  }
}

// second pass for long rows: ordered sum of the piece partials, then the fused epilogue (one warp per row and window)
template <int CH>
__global__ void __launch_bounds__(256) spmm_finish_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
  if (w >= p.n_split) return;
  LaneChunks<CH> lc;
  setup_chunks<32, CH>(p, blockIdx.y * (32 * CH), lane, lc);
  const int row = p.split_row[w];
  const int t0 = p.split_first[w], t1 = p.split_first[w + 1];
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int q = blockIdx.y * (32 * CH) + c * 32 + lane;
      if (q < p.total_f4) {
        float4 v = *reinterpret_cast<const float4*>(p.scratch + ((int64_t)t * p.total_f4 + q) * 4);
        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
      }
    }
  }
  finish_row<CH>(p, lc, acc, row, 0xffffffffu);
}

// generic scalar path: any d (not a multiple of 4, or unaligned operands); one warp per row
__global__ void spmm_scalar_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
  if (row >= p.n_rows) return;
  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  const float s = p.rs ? p.rs[row] : 1.0f;
  for (int sgi = 0; sgi < p.nseg; ++sgi) {
    const llmrec_spmm_seg sg = p.seg[sgi];
    float m = -INFINITY;
    for (int j = lane; j < p.d; j += 32) {
      float a = 0.f;
      for (int e = beg; e < end; ++e) {
        int c = p.col[e];
        float w = p.vals ? p.vals[e] : 1.0f;
        if (p.cs) w *= p.cs[c];
        a = fmaf(w, sg.X[(int64_t)c * sg.ldx + j], a);
      }
      a *= s;
      float* y = sg.Y + (int64_t)row * sg.ldy + j;
      if (sg.flags & LLMREC_SPMM_SOFTMAX) {
        m = fmaxf(m, a);
        *y = a;
      } else {
        *y = sg.Z ? (sg.Z[(int64_t)row * sg.ldz + j] + a) : a;
      }
    }
    if (sg.flags & LLMREC_SPMM_SOFTMAX) {
      m = warp_max(m);
      float t = 0.f;
      for (int j = lane; j < p.d; j += 32) {
        float* y = sg.Y + (int64_t)row * sg.ldy + j;
        float e = expf(*y - m);
        *y = e;
        t += e;
      }
      t = warp_sum(t);
      float inv = 1.0f / t;
      for (int j = lane; j < p.d; j += 32) {
        float* y = sg.Y + (int64_t)row * sg.ldy + j;
        *y = *y * inv + (sg.Z ? sg.Z[(int64_t)row * sg.ldz + j] : 0.f);
      }
    }
  }
}

template <int LPR, int CH, int U, int MINB = 4>
static int launch_spmm(const SpmmParams& p, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int warps = (p.n_tiles + RPW - 1) / RPW;
  const int blocks = (warps + 7) / 8;
  const int windows = (p.total_f4 + LPR * CH - 1) / (LPR * CH);
  if (blocks > 0) spmm_tile_kernel<LPR, CH, U, MINB><<<dim3(blocks, windows), 256, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("spmm_tile");
  if (p.n_split > 0 && !p.split_tickets) {
    const int fw = (p.total_f4 + 32 * CH - 1) / (32 * CH);
    spmm_finish_kernel<CH><<<dim3((p.n_split + 7) / 8, fw), 256, 0, st>>>(p);
    LLMREC_CHECK_LAUNCH("spmm_finish");
  }
  return 0;
}

}  // namespace llmrec

using namespace llmrec;

// Host-side tile planner (native graph-builder step).  Tiles are 8 x int32: {row0, nrows, e0, e1} + 16 row-end byte deltas; pieces of
// long rows (nrows == 0) come first.  Call with tiles_out == NULL to size the outputs:
// counts_out = {n_tiles, n_split, n_split_tiles}.
extern "C" int llmrec_spmm_plan_tiles(const int32_t* rowptr_host, int32_t n_rows, int32_t tile_nnz, int32_t max_rows,
                                      int32_t* tiles_out, int32_t* split_row_out, int32_t* split_first_out, int32_t* counts_out) {
  LLMREC_CHECK_ARG(tile_nnz >= 8 && tile_nnz <= 248 && max_rows >= 1 && max_rows <= 15, "spmm_plan: tile_nnz=%d (8..248) max_rows=%d (1..15) out of range", tile_nnz, max_rows);
  int64_t n_split = 0, n_pieces = 0, n_groups = 0;
  // pass 1: pieces
  for (int32_t r = 0; r < n_rows; ++r) {
    const int32_t deg = rowptr_host[r + 1] - rowptr_host[r];
    if (deg > tile_nnz) {
      if (tiles_out) {
        split_row_out[n_split] = r;
        split_first_out[n_split] = (int32_t)n_pieces;
        for (int32_t b = rowptr_host[r]; b < rowptr_host[r + 1]; b += tile_nnz) {
          int32_t* t = tiles_out + 8 * n_pieces++;
          t[0] = r; t[1] = 0; t[2] = b; t[3] = b + tile_nnz < rowptr_host[r + 1] ? b + tile_nnz : rowptr_host[r + 1];
          t[4] = (int32_t)n_split;                                            // split index (ticket slot)
          t[5] = split_first_out[n_split];                                    // first piece (tile id) of this row
          t[6] = (rowptr_host[r + 1] - rowptr_host[r] + tile_nnz - 1) / tile_nnz;   // pieces of this row
          t[7] = 0;
        }
      } else {
        n_pieces += (deg + tile_nnz - 1) / tile_nnz;
      }
      ++n_split;
    }
  }
  if (tiles_out) split_first_out[n_split] = (int32_t)n_pieces;
  // pass 2: groups of consecutive complete rows
  int32_t r = 0;
  while (r < n_rows) {
    const int32_t deg0 = rowptr_host[r + 1] - rowptr_host[r];
    if (deg0 > tile_nnz) { ++r; continue; }
    int32_t r1 = r + 1;
    while (r1 < n_rows && r1 - r < max_rows) {
      const int32_t dn = rowptr_host[r1 + 1] - rowptr_host[r1];
      if (dn > tile_nnz || rowptr_host[r1 + 1] - rowptr_host[r] > tile_nnz) break;
      ++r1;
    }
    if (tiles_out) {
      int32_t* t = tiles_out + 8 * (n_pieces + n_groups);
      t[0] = r; t[1] = r1 - r; t[2] = rowptr_host[r]; t[3] = rowptr_host[r1];
      uint8_t* dl = reinterpret_cast<uint8_t*>(t + 4);
      for (int i = 0; i < 16; ++i) dl[i] = (uint8_t)(i < r1 - r ? rowptr_host[r + i + 1] - rowptr_host[r] : 0);
    }
    ++n_groups;
    r = r1;
  }
  LLMREC_CHECK_ARG(n_pieces + n_groups < 0x7fffffff, "spmm_plan: too many tiles");
  counts_out[0] = (int32_t)(n_pieces + n_groups);
  counts_out[1] = (int32_t)n_split;
  counts_out[2] = (int32_t)n_pieces;
  return 0;
}

extern "C" int llmrec_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* vals,
                                   const float* row_scale, const float* col_scale,
                                   int32_t n_rows, int32_t n_cols, int32_t d,
                                   const llmrec_spmm_seg* segs, int32_t nseg,
                                   const llmrec_spmm_tiling* tiling, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  (void)n_cols;
  LLMREC_CHECK_ARG(nseg >= 1 && d >= 1 && n_rows >= 0, "spmm: bad sizes nseg=%d d=%d n_rows=%d", nseg, d, n_rows);
  if (n_rows == 0) return 0;
  cudaStream_t st = as_stream(stream);

  bool vec_ok = (d % 4 == 0) && tiling != nullptr && tiling->n_tiles > 0;
  bool any_softmax = false;
  for (int s = 0; s < nseg; ++s) {
    vec_ok = vec_ok && aligned16(segs[s].X) && aligned16(segs[s].Y) && segs[s].ldx % 4 == 0 && segs[s].ldy % 4 == 0 &&
             (!segs[s].Z || (aligned16(segs[s].Z) && segs[s].ldz % 4 == 0));
    any_softmax = any_softmax || (segs[s].flags & LLMREC_SPMM_SOFTMAX);
  }
  const int f4 = d / 4;
  const bool pow2 = (f4 & (f4 - 1)) == 0;
  if (any_softmax && !(pow2 && f4 <= 32)) vec_ok = false;  // fused softmax needs a power-of-two lane group

  LLMREC_CHECK_ARG(vec_ok || !(tiling && tiling->src_mask), "spmm: the source-row mask needs the vectorised kernels (d %% 4 == 0, aligned operands, a tile plan)");
  if (!vec_ok) {
    for (int s0 = 0; s0 < nseg; s0 += LLMREC_MAX_SEG) {
      SpmmParams p{};
      p.rowptr = rowptr; p.col = col; p.vals = vals; p.rs = row_scale; p.cs = col_scale;
      p.n_rows = n_rows; p.d = d; p.nseg = min(nseg - s0, LLMREC_MAX_SEG);
      for (int s = 0; s < p.nseg; ++s) p.seg[s] = segs[s0 + s];
      spmm_scalar_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(p);
      LLMREC_CHECK_LAUNCH("spmm_scalar");
    }
    return 0;
  }

  for (int s0 = 0; s0 < nseg; s0 += LLMREC_MAX_SEG) {
    SpmmParams p{};
    p.rowptr = rowptr; p.col = col; p.vals = vals; p.rs = row_scale; p.cs = col_scale;
    p.n_rows = n_rows; p.d = d; p.nseg = min(nseg - s0, LLMREC_MAX_SEG);
    p.f4_per_seg = f4; p.total_f4 = f4 * p.nseg; p.any_softmax = any_softmax ? 1 : 0;
    for (int s = 0; s < p.nseg; ++s) p.seg[s] = segs[s0 + s];
    p.tiles = reinterpret_cast<const int4*>(tiling->tiles); p.n_tiles = tiling->n_tiles;
    p.n_split_tiles = tiling->n_split_tiles; p.scratch = tiling->scratch;
    p.split_row = tiling->split_row; p.split_first = tiling->split_first; p.n_split = tiling->n_split;
    p.src_mask = tiling->src_mask;
    p.split_tickets = ((p.total_f4 + 7) / 8 <= kTicketWindows) ? tiling->split_tickets : nullptr;
    LLMREC_CHECK_ARG(p.n_split == 0 || p.scratch != nullptr, "spmm: long-row pieces need the scratch buffer");
    int rc = 0;
    static const int bulk_mode = getenv("LLMREC_SPMM_BULK") ? atoi(getenv("LLMREC_SPMM_BULK")) : -1;   // -1 auto, 0 off, 1 on
    const bool bulk_ok = p.nseg == 1 && d == 128 && !p.src_mask && segs[s0].ldx == 128;
    if (bulk_ok && (bulk_mode == 1 || (bulk_mode == -1 && kSpmmBulkAuto && (int64_t)n_cols * 512 >= ((int64_t)576 << 20)))) {
      p.split_tickets = nullptr;                                     // the staged kernel keeps the second-pass reduction of long rows
      const size_t smem = (size_t)kBulkWarps * kBulkStages * kBulkRows * 512 + kBulkWarps * kBulkStages * sizeof(uint64_t);
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(spmm_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
      spmm_bulk_kernel<<<(p.n_tiles + kBulkWarps - 1) / kBulkWarps, kBulkWarps * 32, smem, st>>>(p);
      LLMREC_CHECK_LAUNCH("spmm_bulk");
      if (p.n_split > 0) { spmm_finish_kernel<1><<<dim3((p.n_split + 7) / 8, 1), 256, 0, st>>>(p); LLMREC_CHECK_LAUNCH("spmm_finish"); }
      continue;
    }
    if (p.total_f4 <= 8) rc = launch_spmm<8, 1, 4>(p, st);
    else if (p.total_f4 <= 16) rc = launch_spmm<16, 1, 4>(p, st);
    else rc = launch_spmm<32, 1, 4, 4>(p, st);  // 32 lanes x one 16-byte chunk, 4 gathers in flight per lane, 32 warps/SM (8 in flight / 48-64 warps per SM
                                                // measured 15-140 % slower at the synthetic scale); wider concatenations run as 32-chunk column windows over blockIdx.y
    if (rc) return rc;
  }
  return 0;
}

extern "C" int llmrec_spmm_rows_f32(const int32_t* rowptr, const int32_t* col, const float* vals, const float* row_scale, const float* col_scale,
                                    int32_t d, const llmrec_spmm_seg* seg, const int32_t* rows, const int32_t* n_rows_dev, int32_t max_rows,
                                    const uint32_t* src_mask, int32_t cta_per_row, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  LLMREC_CHECK_ARG(seg && rows && n_rows_dev && max_rows >= 0, "spmm_rows: row list, its device-side length and a segment are required");
  LLMREC_CHECK_ARG((d == 32 || d == 64 || d == 128) && aligned16(seg->X) && aligned16(seg->Y) && seg->ldx % 4 == 0 && seg->ldy % 4 == 0 &&
                   (!seg->Z || (aligned16(seg->Z) && seg->ldz % 4 == 0)), "spmm_rows: d in {32, 64, 128} and 16-byte aligned operands required (d=%d)", d);
  if (max_rows == 0) return 0;
  SpmmParams p{};
  p.rowptr = rowptr; p.col = col; p.vals = vals; p.rs = row_scale; p.cs = col_scale; p.d = d; p.nseg = 1; p.seg[0] = *seg;
  p.src_mask = src_mask; p.rows = rows; p.n_rows_dev = n_rows_dev; p.max_rows = max_rows;
  const int lpr = d / 4, rpw = 32 / lpr;
  long long want = ((long long)max_rows + 8LL * rpw - 1) / (8LL * rpw);
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);          // persistent: 8 CTAs of 8 warps per SM
  cudaStream_t st = as_stream(stream);
  if (cta_per_row) {
    spmm_rows_cta_kernel<<<max_rows < 148 * 8 ? max_rows : 148 * 8, 256, 0, st>>>(p);
    LLMREC_CHECK_LAUNCH("spmm_rows_cta");
    return 0;
  }
  if (lpr == 32) spmm_rows_kernel<32, 4><<<blocks, 256, 0, st>>>(p);
  else if (lpr == 16) spmm_rows_kernel<16, 4><<<blocks, 256, 0, st>>>(p);
  else spmm_rows_kernel<8, 4><<<blocks, 256, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("spmm_rows");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// stand-alone row softmax / softmax backward (Models.py:174-175 and its autograd)
// ---------------------------------------------------------------------------------------------
namespace llmrec {
__global__ void row_softmax_kernel(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)(blockDim.x >> 5)) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* x = X + row * ldx;
  float* y = Y + row * ldy;
  float m = -INFINITY;
  for (int j = lane; j < d; j += 32) m = fmaxf(m, x[j]);
  m = warp_max(m);
  float t = 0.f;
  for (int j = lane; j < d; j += 32) t += expf(x[j] - m);
  t = warp_sum(t);
  float inv = 1.0f / t;
  for (int j = lane; j < d; j += 32) y[j] = expf(x[j] - m) * inv;
}
__global__ void row_softmax_bwd_rows_kernel(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx,
                                            const int* rows, const int* n_rows_dev, int max_rows, int d) {
  const int lane = threadIdx.x & 31;
  int n = __ldg(n_rows_dev);
  n = n < max_rows ? n : max_rows;
  for (long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < n; idx += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int64_t row = rows[idx];
    if (row < 0) continue;
    const float* s = S + row * lds;
    const float* g = dS + row * ldds;
    float* o = dX + row * lddx;
    float t = 0.f;
    for (int j = lane; j < d; j += 32) t = fmaf(g[j], s[j], t);
    t = warp_sum(t);
    for (int j = lane; j < d; j += 32) o[j] = s[j] * (g[j] - t);
  }
}
__global__ void row_softmax_bwd_kernel(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)(blockDim.x >> 5)) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* s = S + row * lds;
  const float* g = dS + row * ldds;
  float* o = dX + row * lddx;
  float t = 0.f;
  for (int j = lane; j < d; j += 32) t = fmaf(g[j], s[j], t);
  t = warp_sum(t);
  for (int j = lane; j < d; j += 32) o[j] = s[j] * (g[j] - t);
}
}  // namespace llmrec

extern "C" int llmrec_row_softmax_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n, int32_t d, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (n <= 0) return 0;
  row_softmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(X, ldx, Y, ldy, n, d);
  LLMREC_CHECK_LAUNCH("row_softmax");
  return 0;
}
extern "C" int llmrec_row_softmax_bwd_f32(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx,
                                          int64_t n, int32_t d, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (n <= 0) return 0;
  row_softmax_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(S, lds, dS, ldds, dX, lddx, n, d);
  LLMREC_CHECK_LAUNCH("row_softmax_bwd");
  return 0;
}
extern "C" int llmrec_row_softmax_bwd_rows_f32(const float* S, int64_t lds, const float* dS, int64_t ldds, float* dX, int64_t lddx,
                                               const int32_t* rows, const int32_t* n_rows_dev, int32_t max_rows, int32_t d, llmrec_stream_t stream) {
  LLMREC_REQUIRE_DEVICE();
  if (max_rows <= 0) return 0;
  long long want = ((long long)max_rows + 7) / 8;
  row_softmax_bwd_rows_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, as_stream(stream)>>>(S, lds, dS, ldds, dX, lddx, rows, n_rows_dev, max_rows, d);
  LLMREC_CHECK_LAUNCH("row_softmax_bwd_rows");
  return 0;
}
