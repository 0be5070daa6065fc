// Propagation SpMM over a CSR pattern with row/column scaling (Models.py:57-61,152-183).
//
//   Y_s[r,:] (+)= epi( rs[r] * sum_{e in row r} v[e] * cs[col[e]] * X_s[col[e],:] )
//
// HBM-bound gather.  Design (sm_100a, no tensor cores -- this is index/byte work):
//   * CSR-row-tiled: one warp (or sub-warp for narrow operands) owns a (row, [beg,end)) tile; rows
//     longer than tile_nnz are split over several warps and reduced by a deterministic second pass.
//   * the index stream (col, v, cs) is read once, coalesced, 32 entries per warp step and broadcast
//     with warp shuffles; every dense operand sharing the pattern ("segment") is gathered in the same
//     pass, so the 20 reference SpMMs per forward collapse to 4 launches.
//   * each lane owns up to CH 16-byte column chunks of the concatenated segment row and keeps the
//     partial sums in registers; neighbour rows are fetched with 128-bit read-only loads, 4 x CH in
//     flight per lane (unrolled), which is what hides the gather latency.
//   * scale, optional row-softmax over a segment's d columns (sub-warp shuffle reduction) and
//     optional accumulate are fused into the store.
#include "common.cuh"

namespace llmrec {

struct SpmmParams {
  const int* rowptr; const int* col; const float* vals; const float* rs; const float* cs;
  int n_rows; int d; int nseg; int f4_per_seg; int total_f4;
  // tiling (n_tiles == 0 -> one work item per row)
  const int* tile_row; const int* tile_beg; int n_tiles; int tile_nnz; int n_split_tiles; float* scratch;
  const int* split_row; const int* split_first; int n_split;
  llmrec_spmm_seg seg[LLMREC_MAX_SEG];
};

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
__device__ __forceinline__ void setup_chunks(const SpmmParams& p, int lane_in, LaneChunks<CH>& lc) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int q = c * LPR + lane_in;
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

// scale + (softmax) + (accumulate) + store for one output row
template <int LPR, int CH>
__device__ __forceinline__ void finish_row(const SpmmParams& p, const LaneChunks<CH>& lc, float4 (&acc)[CH], int row, bool valid) {
  float s = (p.rs != nullptr && valid) ? p.rs[row] : 1.0f;
  const int G = p.f4_per_seg;  // lanes per segment (power of two when any softmax flag is set)
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float4 a = acc[c];
    a.x *= s; a.y *= s; a.z *= s; a.w *= s;
    // softmax is warp-uniform per launch group decision: all lanes run the shuffles
    bool any_sm = __any_sync(0xffffffffu, (lc.flags[c] & LLMREC_SPMM_SOFTMAX) != 0);
    if (any_sm) {
      float m = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
      for (int o = G >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float4 e = make_float4(expf(a.x - m), expf(a.y - m), expf(a.z - m), expf(a.w - m));
      float t = (e.x + e.y) + (e.z + e.w);
      for (int o = G >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lc.flags[c] & LLMREC_SPMM_SOFTMAX) {
        float inv = 1.0f / t;
        a = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
      }
    }
    if (valid && lc.on[c]) {
      float* y = lc.yb[c] + (int64_t)row * lc.ldy[c];
      if (lc.zb[c]) {
        float4 o = *reinterpret_cast<const float4*>(lc.zb[c] + (int64_t)row * lc.ldz[c]);
        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      }
      st4(y, a);
    }
  }
}

template <int LPR, int CH>
__global__ void __launch_bounds__(256) spmm_kernel(const SpmmParams p) {
  constexpr int RPW = 32 / LPR;  // work items per warp
  const int lane = threadIdx.x & 31;
  const int lane_in = lane % LPR;
  const int sub = lane / LPR;
  const int warp_global = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
  const int n_items = p.n_tiles > 0 ? p.n_tiles : p.n_rows;
  const int item = warp_global * RPW + sub;
  if (warp_global * RPW >= n_items) return;  // whole warp idle
  const bool valid = item < n_items;

  LaneChunks<CH> lc;
  setup_chunks<LPR, CH>(p, lane_in, lc);

  int row = 0, beg = 0, end = 0;
  bool whole = true;
  if (valid) {
    if (p.n_tiles > 0) {
      row = p.tile_row[item];
      beg = p.tile_beg[item];
      int rend = p.rowptr[row + 1];
      end = min(beg + p.tile_nnz, rend);
      whole = (beg == p.rowptr[row]) && (end == rend);
    } else {
      row = item;
      beg = p.rowptr[row];
      end = p.rowptr[row + 1];
    }
  }
  int len = end - beg;
  int maxlen = len;
  if (RPW > 1) {
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  }

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int k0 = 0; k0 < maxlen; k0 += LPR) {
    int e = beg + k0 + lane_in;
    int cidx = 0;
    float w = 0.f;
    if (e < end) {
      cidx = __ldg(p.col + e);
      w = p.vals ? __ldg(p.vals + e) : 1.0f;
      if (p.cs) w *= __ldg(p.cs + cidx);
    }
    int cnt = min(LPR, maxlen - k0);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      int cj = __shfl_sync(0xffffffffu, cidx, j, LPR);
      float wj = __shfl_sync(0xffffffffu, w, j, LPR);
      if (k0 + j < len) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if (lc.on[c]) fma4(acc[c], wj, ldg4(lc.xb[c] + (int64_t)cj * lc.ldx[c]));
        }
      }
    }
  }

  if (whole) {
    finish_row<LPR, CH>(p, lc, acc, row, valid);
  } else {
    // partial tile of a split row: raw sums to scratch (tiles of split rows are numbered first)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int q = c * LPR + lane_in;
      if (q < p.total_f4) st4(p.scratch + ((int64_t)item * p.total_f4 + q) * 4, acc[c]);
    }
    // keep warp-collective softmax shuffles of other sub-groups safe: nothing to do, finish_row
    // is only entered by sub-groups with whole rows; with RPW>1 splitting is disabled on the host.
  }
}

// second pass for split rows: ordered sum of the tile partials, then the fused epilogue
template <int CH>
__global__ void __launch_bounds__(256) spmm_finish_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
  if (w >= p.n_split) return;
  LaneChunks<CH> lc;
  setup_chunks<32, CH>(p, lane, lc);
  const int row = p.split_row[w];
  const int t0 = p.split_first[w], t1 = p.split_first[w + 1];
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int q = c * 32 + lane;
      if (q < p.total_f4) {
        float4 v = *reinterpret_cast<const float4*>(p.scratch + ((int64_t)t * p.total_f4 + q) * 4);
        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
      }
    }
  }
  finish_row<32, CH>(p, lc, acc, row, true);
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
    // pass 1: values (kept in Y), track max for softmax
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

template <int LPR, int CH>
static int launch_spmm(const SpmmParams& p, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int n_items = p.n_tiles > 0 ? p.n_tiles : p.n_rows;
  const int warps = (n_items + RPW - 1) / RPW;
  const int blocks = (warps + 7) / 8;
  if (blocks > 0) spmm_kernel<LPR, CH><<<blocks, 256, 0, st>>>(p);
  LLMREC_CHECK_LAUNCH("spmm");
  if (p.n_split > 0) {
    if constexpr (LPR == 32) {
      spmm_finish_kernel<CH><<<(p.n_split + 7) / 8, 256, 0, st>>>(p);
      LLMREC_CHECK_LAUNCH("spmm_finish");
    }
  }
  return 0;
}

}  // namespace llmrec

using namespace llmrec;

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

  bool vec_ok = (d % 4 == 0);
  bool any_softmax = false;
  for (int s = 0; s < nseg; ++s) {
    vec_ok = vec_ok && aligned16(segs[s].X) && aligned16(segs[s].Y) && segs[s].ldx % 4 == 0 && segs[s].ldy % 4 == 0 &&
             (!segs[s].Z || (aligned16(segs[s].Z) && segs[s].ldz % 4 == 0));
    any_softmax = any_softmax || (segs[s].flags & LLMREC_SPMM_SOFTMAX);
  }
  const int f4 = d / 4;
  const bool pow2 = (f4 & (f4 - 1)) == 0;
  if (any_softmax && !(pow2 && f4 <= 32)) vec_ok = false;  // fused softmax needs a power-of-two lane group

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

  // group segments so one launch covers at most 8 chunks per lane (256 float4) and MAX_SEG segments
  int max_seg = 256 / f4;
  if (max_seg < 1) max_seg = 1;
  if (max_seg > LLMREC_MAX_SEG) max_seg = LLMREC_MAX_SEG;
  LLMREC_CHECK_ARG(f4 <= 256, "spmm: d=%d too wide (max 1024)", d);
  for (int s0 = 0; s0 < nseg; s0 += max_seg) {
    SpmmParams p{};
    p.rowptr = rowptr; p.col = col; p.vals = vals; p.rs = row_scale; p.cs = col_scale;
    p.n_rows = n_rows; p.d = d; p.nseg = min(nseg - s0, max_seg);
    p.f4_per_seg = f4; p.total_f4 = f4 * p.nseg;
    for (int s = 0; s < p.nseg; ++s) p.seg[s] = segs[s0 + s];
    const bool narrow = p.total_f4 <= 16 && pow2;
    if (tiling && tiling->n_tiles > 0 && !narrow) {
      p.tile_row = tiling->tile_row; p.tile_beg = tiling->tile_beg; p.n_tiles = tiling->n_tiles;
      p.tile_nnz = tiling->tile_nnz; p.n_split_tiles = tiling->n_split_tiles; p.scratch = tiling->scratch;
      p.split_row = tiling->split_row; p.split_first = tiling->split_first; p.n_split = tiling->n_split;
      LLMREC_CHECK_ARG(p.n_split == 0 || p.scratch != nullptr, "spmm: tiling with split rows needs scratch");
    }
    int rc = 0;
    if (narrow && p.total_f4 <= 8) rc = launch_spmm<8, 1>(p, st);
    else if (narrow) rc = launch_spmm<16, 1>(p, st);
    else {
      int ch = (p.total_f4 + 31) / 32;
      switch (ch) {
        case 1: rc = launch_spmm<32, 1>(p, st); break;
        case 2: rc = launch_spmm<32, 2>(p, st); break;
        case 3: rc = launch_spmm<32, 3>(p, st); break;
        case 4: rc = launch_spmm<32, 4>(p, st); break;
        case 5: rc = launch_spmm<32, 5>(p, st); break;
        case 6: rc = launch_spmm<32, 6>(p, st); break;
        case 7: rc = launch_spmm<32, 7>(p, st); break;
        default: rc = launch_spmm<32, 8>(p, st); break;
      }
    }
    if (rc) return rc;
  }
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
