// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
//
//   warp 0      : TMA producer (one elected lane) - A/B tiles, 128B-swizzled, K-major, mbarrier ring
//   warp 1      : MMA issuer (one elected lane)   - tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                                                   two accumulator stages so the epilogue overlaps the next tile
//   warp 2      : TMEM allocator / deallocator
//   warps 4..11 : epilogue - tcgen05.ld (software pipelined), bias / scale / GELU / row mask in registers,
//                 swizzled staging in shared memory, TMA stores (fp32 and/or fp16 hi(+lo)); no thread ever
//                 issues a global store, so every HBM/L2 write is a full-line bulk transfer
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
constexpr int GEMM_THREADS = 384;  // 12 warps
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BLOCK_N * GEMM_BLOCK_K * 2;   // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_EPI_STAGE_BYTES = 4096;                      // per epilogue warp: 32 rows x 128 B
constexpr int GEMM_SMEM_BIAS = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES;
constexpr int GEMM_SMEM_BAR = GEMM_SMEM_BIAS + 2 * GEMM_BLOCK_N * 4;
constexpr int GEMM_SMEM_TOTAL = GEMM_SMEM_BAR + 256;            // 231 680 B of the 232 448 B a CTA may own; no slack:
                                                                // the kernel has no static smem, so the dynamic base is 1024B aligned
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
  int out_f32;              // write fp32 through map o_f32
  int out_hi;               // write fp16 hi through map o_hi
  int out_lo;               // write fp16 lo through map o_lo (requires out_hi)
  int act;                  // 0 none, 1 GELU
  float col_scale;          // tiles whose first column is < col_scale_limit are multiplied by col_scale
  int col_scale_limit;      // (multiple of 256)
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
               const __grid_constant__ CUtensorMap o_f32, const __grid_constant__ CUtensorMap o_hi,
               const __grid_constant__ CUtensorMap o_lo, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned stage buffers
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + GEMM_STAGES * GEMM_A_BYTES;
  uint8_t* smem_epi = smem + GEMM_STAGES * GEMM_STAGE_BYTES;
  float* smem_bias = reinterpret_cast<float*>(smem + GEMM_SMEM_BIAS);   // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_BAR);
  uint64_t* full_bar = bars;                          // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * GEMM_STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * GEMM_STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int tiles_m_per_batch = (p.rows_per_batch + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles_n = p.N / GEMM_BLOCK_N;
  const int num_tiles = p.batches * tiles_m_per_batch * tiles_n;
  const int kb_total = p.kb_per_pass * p.n_pass;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&a_hi);
    tma_prefetch_desc(&b_hi);
    if (p.n_pass > 1) {
      tma_prefetch_desc(&a_lo);
      tma_prefetch_desc(&b_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < GEMM_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<GEMM_TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % tiles_n;
        const int m_tile = tile / tiles_n;
        const int batch = m_tile / tiles_m_per_batch;
        const int row0 = (m_tile % tiles_m_per_batch) * GEMM_BLOCK_M;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pass = kb / p.kb_per_pass;
          const int kk = kb - pass * p.kb_per_pass;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
          tma_load_3d(smem_a + stage * GEMM_A_BYTES, (pass == 1) ? &a_lo : &a_hi, &full_bar[stage],
                      kk * GEMM_BLOCK_K, row0, batch);
          tma_load_2d(smem_b + stage * GEMM_B_BYTES, (pass == 2) ? &b_lo : &b_hi, &full_bar[stage],
                      kk * GEMM_BLOCK_K, n_tile * GEMM_BLOCK_N);
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BLOCK_M, GEMM_BLOCK_N, 0, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * GEMM_BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t adesc = make_desc_k_sw128(smem_u32(smem_a + stage * GEMM_A_BYTES));
          const uint64_t bdesc = make_desc_k_sw128(smem_u32(smem_b + stage * GEMM_B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advancing K by 16 fp16 = 32 bytes inside the 128B swizzle row: +2 in 16-byte units
            umma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= GEMM_EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - GEMM_EPI_WARP0;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;                // which 128-column half of the tile
    const int lane = (int)lane_id();
    const int epi_tid = threadIdx.x - GEMM_EPI_WARP0 * 32;   // 0..255
    uint8_t* stage_buf = smem_epi + ew * GEMM_EPI_STAGE_BYTES;
    // swizzled staging addresses of this lane's row: 128-byte rows (fp32) and 64-byte rows (fp16)
    uint8_t* row128 = stage_buf + lane * 128;
    uint8_t* row64_hi = stage_buf + lane * 64;
    uint8_t* row64_lo = stage_buf + 2048 + lane * 64;
    const int sw128 = lane & 7;
    const int sw64 = (lane >> 1) & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % tiles_n;
      const int m_tile = tile / tiles_n;
      const int batch = m_tile / tiles_m_per_batch;
      const int warp_row0 = (m_tile % tiles_m_per_batch) * GEMM_BLOCK_M + quarter * 32;
      const int row_in_batch = warp_row0 + lane;
      const bool warp_ok = warp_row0 < p.rows_per_batch;
      const bool zero_row = p.valid_rows != nullptr && row_in_batch >= __ldg(p.valid_rows + batch);
      const float scale = (n_tile * GEMM_BLOCK_N < p.col_scale_limit) ? p.col_scale : 1.0f;
      // stage this tile's bias in shared memory (one column per epilogue thread, double buffered by tile parity)
      float* sbias = smem_bias + (it & 1) * GEMM_BLOCK_N;
      sbias[epi_tid] = p.bias ? __ldg(p.bias + n_tile * GEMM_BLOCK_N + epi_tid) : 0.0f;
      named_bar_sync(1, GEMM_EPI_WARPS * 32);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * GEMM_BLOCK_N + half * 128);
      uint32_t r[2][32];
      tmem_ld_32x32b_x32(taddr0, r[0]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c + 1 < 4) tmem_ld_32x32b_x32(taddr0 + (c + 1) * 32, r[(c + 1) & 1]);   // prefetch next chunk
        const int col0 = n_tile * GEMM_BLOCK_N + half * 128 + c * 32;
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(sbias + half * 128 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = b4[i];
          v[4 * i + 0] = (__uint_as_float(r[c & 1][4 * i + 0]) + bb.x) * scale;
          v[4 * i + 1] = (__uint_as_float(r[c & 1][4 * i + 1]) + bb.y) * scale;
          v[4 * i + 2] = (__uint_as_float(r[c & 1][4 * i + 2]) + bb.z) * scale;
          v[4 * i + 3] = (__uint_as_float(r[c & 1][4 * i + 3]) + bb.w) * scale;
        }
        if (p.act == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) gelu_fast2(v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1]);
        }
        if (zero_row) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.0f;
        }
        if (warp_ok) {
          if (p.out_f32) {
            if (lane == 0) tma_store_wait_read();   // previous bulk store has finished reading the staging buffer
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(row128 + ((i ^ sw128) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&o_f32, stage_buf, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
          if (p.out_hi) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(row64_hi + ((i ^ sw64) << 4)) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            if (p.out_lo) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<uint4*>(row64_lo + ((i ^ sw64) << 4)) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&o_hi, stage_buf, col0, warp_row0, batch);
              if (p.out_lo) tma_store_3d(&o_lo, stage_buf + 2048, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
        }
        if (c + 1 < 4) tmem_ld_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all();   // bulk stores must be complete before the CTA exits
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<GEMM_TMEM_COLS>(tmem_base);
  }
}

}  // namespace syl
