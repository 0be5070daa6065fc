// Parameter blocks shared by the projection kernels (proj_tc.cu: operands from shared memory; proj_tc2.cu: A operand
// from tensor memory).
#pragma once
#include "tc_common.cuh"

namespace llmrec {

constexpr int kMaxProb = 8;
constexpr int BM = 128;  // tile rows (fwd) / tile features (wgrad) = UMMA M
constexpr int BK = 32;   // fp32 per 128-byte swizzle row
constexpr uint32_t kTileA = BM * BK * 4;  // 16 KiB

struct FwdProblem { int n, k, kblocks, tile_start; long long ldy; float* Y; const float* bias; };
struct FwdParams {
  CUtensorMap tmA[kMaxProb];
  CUtensorMap tmW[kMaxProb];  // [2d x k] (hi rows then lo rows) when SPLIT, [d x k] otherwise
  FwdProblem prob[kMaxProb];
  int n_prob, total_tiles, d, stages, tmem_cols;
  int dbg;    // TIMING experiments only (LLMREC_PROJ_DBG, results are wrong): bit 1 no W loads, 2 no MMAs, 4 no transform -- DESIGN.md 5 quotes the numbers
};


struct WgProblem { int n, k, ft_tiles, chunks, rows_per_chunk, item_start; };
struct WgParams {
  CUtensorMap tmX[kMaxProb];
  CUtensorMap tmG[kMaxProb];
  WgProblem prob[kMaxProb];
  int n_prob, total_items, d, stages, tmem_cols;
  float* partial;  // [total_items][128][d]
};


int proj_fwd_ts_launch(const FwdParams& P, int grid, cudaStream_t st);
int proj_wgrad_ts_launch(const WgParams& P, int grid, cudaStream_t st);

}  // namespace llmrec
