// Shared definitions of the persistent, warp-specialised tcgen05 GEMM:  D[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
// The kernel itself is gemm3_tc_kernel (gemm3_tc.cuh, 2-CTA clusters, 16 epilogue warps); its cluster helpers live in
// gemm2_tc.cuh.  The round-1 single-CTA and 8-epilogue-warp kernels were removed after the A/B measurements in
// profiles/r01_ncu_summary.md and profiles/r02_bench_ab.md.
//
// One kernel serves every dense contraction of the Segmenter forward path:
//   * Linear layers (feature projection, QKV, out-proj, FFN1, FFN2)       -> A is a plain [M,K] matrix
//   * conv1..conv6 of the feature encoder as implicit GEMM                 -> A is an overlapping-row
//     view of the channels-last activation (row stride = conv_stride*C_in, row length = k*C_in)
// Split precision ("3-pass"): the K loop is run over (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) into the same
// accumulator, which restores ~fp32 accuracy from fp16 tensor-core operands (precision).
#pragma once

#include "common.cuh"

namespace syl {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_N = 256;
constexpr int GEMM_BLOCK_K = 64;   // 64 fp16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARP0 = 4;  // warps 0..3: TMA producer, MMA issuer, TMEM allocator, spare
constexpr uint32_t GEMM_TMEM_COLS = 512;                        // 2 accumulator stages x 256 columns

struct GemmParams {
  // problem shape
  int rows_per_batch;   // valid output rows per batch item (M when batches == 1)
  int batches;
  int N;                // output columns (multiple of 256)
  int kb_per_pass;      // number of 64-wide K blocks in one pass
  int n_pass;           // 1, or 3 for split precision
  // epilogue
  const float* bias;        // [N] or null
  const int* valid_rows;    // [batches] or null: rows >= valid_rows[b] are written as zero
  int skip_invalid_tiles;   // trimmed mode: 256-row tiles that start at or beyond valid_rows[b] are not computed at all
  int out_f32;              // write fp32 through map o_f32
  int out_hi;               // write fp16 hi through map o_hi
  int out_lo;               // write fp16 lo through map o_lo (requires out_hi)
  int act;                  // 0 none, 1 GELU
  float col_scale;          // tiles whose first column is < col_scale_limit are multiplied by col_scale
  int col_scale_limit;      // (multiple of 256)
#ifdef SYL_DIAG
  long long* trace;         // timeline probe (tools/gemm_trace.py): CTA 0 writes clock64 stamps, slots in gemm3_tc.cuh
  int epi_skip;             // SYL_GEMM_EPI_SKIP (timing experiments, wrong results): 1 = no staging / stores, 2 = no epilogue work at all
#endif
};

// Timeline probe of the GEMM kernel, diagnostic build only: CTA 0 (leader of cluster 0) stamps clock64 into
// p.trace[slot].  Slots: 0 globaltimer at entry, 1 entry, 2 prologue done, 3 first TMA issued, 8 + 3 t + {0, 1, 2} MMA
// thread, tile t: accumulator free / first operands landed / last MMA committed, 64 + 2 t + {0, 1} epilogue warp 4, tile t:
// accumulator complete / last store committed, 100 stores complete, 101 exit, 102 globaltimer at exit.
#ifdef SYL_DIAG
#define GEMM_TRACE(p, slot)                                                                  \
  do {                                                                                       \
    if ((p).trace != nullptr && blockIdx.x == 0 && (slot) < 128) (p).trace[(slot)] = clock64(); \
  } while (0)
#define GEMM_TRACE_NS(p, slot)                                                               \
  do {                                                                                       \
    if ((p).trace != nullptr && blockIdx.x == 0) {                                           \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      (p).trace[(slot)] = (long long)t_;                                                     \
    }                                                                                        \
  } while (0)
#else
#define GEMM_TRACE(p, slot) do { } while (0)
#define GEMM_TRACE_NS(p, slot) do { } while (0)
#endif

}  // namespace syl
