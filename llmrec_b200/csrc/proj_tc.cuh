// Parameter blocks shared by the projection kernels (proj_tc.cu: operands from shared memory; proj_tc2.cu: A operand
// from tensor memory).
#pragma once
#include "tc_common.cuh"

namespace llmrec {

constexpr int kMaxProb = 8;
constexpr int BM = 128;  // tile rows (fwd) / tile features (wgrad) = UMMA M
constexpr int BK = 32;   // fp32 per 128-byte swizzle row
constexpr uint32_t kTileA = BM * BK * 4;  // 16 KiB

struct FwdProblem { int n, k, kblocks, tile_start; long long ldy; float* Y; const float* bias; int panel; /* 0 = row-major X, else rows per column panel (n rounded up to 128) */ };
struct FwdParams {
  CUtensorMap tmA[kMaxProb];
  CUtensorMap tmW[kMaxProb];  // [2d x k] (hi rows then lo rows) when SPLIT, [d x k] otherwise
  FwdProblem prob[kMaxProb];
  int n_prob, total_tiles, d, stages, tmem_cols;
  int nw, nt; // v3 forward (proj_fwd_ts3_kernel): slots of the W ring and of the TMEM A ring (`stages` = slots of the smem A ring)
  int wbox;   // experiment (LLMREC_PROJ_WBOX): W_hi and W_lo of a k-block arrive as ONE [2d x 32] TMA box (they are adjacent rows of the split matrix and adjacent in the stage)
  int skipw;  // TIMING experiment only (LLMREC_PROJ_SKIPW, results are wrong): the W tiles are not fetched, isolating the L2->SM cost of re-reading W per k-block
  int krot;   // experiment (LLMREC_PROJ_KROT): CTA b starts its k loop at block b mod kblocks, so concurrent CTAs read different columns
};


struct WgProblem { int n, k, ft_tiles, chunks, rows_per_chunk, item_start, panel; /* 0 = row-major X, else rows per column panel */
                   int x3d, g3d; /* experiments: the X tile / the dY tile of a stage arrives as ONE rank-3 TMA box (tmX / tmG are rank-3 maps then) */ };
struct WgParams {
  CUtensorMap tmX[kMaxProb];
  CUtensorMap tmG[kMaxProb];
  WgProblem prob[kMaxProb];
  int n_prob, total_items, d, stages, tmem_cols;
  int ng, nt; // v3 wgrad (proj_wgrad_ts3_kernel): slots of the dY ring and of the TMEM A ring (`stages` = slots of the smem X ring)
  float* partial;  // [total_items][128][d]
};


int proj_fwd_ts_launch(const FwdParams& P, int grid, cudaStream_t st);
int proj_fwd_ts3_launch(const FwdParams& P, int grid, cudaStream_t st);
int proj_wgrad_ts_launch(const WgParams& P, int grid, cudaStream_t st);
int proj_wgrad_ts3_launch(const WgParams& P, int grid, cudaStream_t st);

}  // namespace llmrec
