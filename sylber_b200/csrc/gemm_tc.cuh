// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
//
//   warp 0      : TMA producer (one elected lane) - A/B tiles, 128B-swizzled, K-major, 4-stage mbarrier ring
//   warp 1      : MMA issuer (one elected lane)   - tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                                                   two accumulator stages so the epilogue overlaps the next tile
//   warp 2      : TMEM allocator / deallocator
//   warps 4..11 : epilogue - tcgen05.ld, bias / scale / erf-GELU / residual / row mask, fp32 and/or
//                 fp16 hi(+lo) stores
//
// One kernel serves every dense contraction of the Segmenter forward path:
//   * Linear layers (feature projection, QKV, out-proj, FFN1, FFN2)       -> A is a plain [M,K] matrix
//   * conv1..conv6 of the feature encoder as implicit GEMM                 -> A is an overlapping-row
//     view of the channels-last activation (row stride = conv_stride*C_in, row length = k*C_in)
//   * the grouped positional conv (k=128, 16 groups of 48)                 -> K loop walks the 128 taps,
//     each tap shifts the A row window by one frame; TMA zero-fills rows outside [0,T)
// Split precision ("3-pass"): the K loop is run over (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) into the same
// accumulator, which restores ~fp32 accuracy from fp16 tensor-core operands (precision).
#pragma once

#include "common.cuh"

namespace syl {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 fp16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 384;  // 12 warps
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_WARPS = 8;

struct GemmParams {
  // problem shape
  int rows_per_batch;   // valid output rows per batch item (M when batches == 1)
  int batches;
  int N;                // output columns (multiple of BLOCK_N)
  int kb_per_pass;      // number of 64-wide K blocks in one pass
  int n_pass;           // 1, or 3 for split precision
  // A addressing: k-block kk of a pass -> tap = kk / kb_per_tap, kc = kk % kb_per_tap
  //   A column  = a_col_per_ntile * n_tile + kc * 64
  //   A row     = tile_row0 + row_offset + tap * tap_row_step
  int kb_per_tap;
  int tap_row_step;
  int row_offset;
  int a_col_per_ntile;
  // epilogue
  const float* bias;        // [N] or null
  const float* residual;    // [M_total, ldo] fp32 or null (added after activation)
  const int* valid_rows;    // [batches] or null: rows >= valid_rows[b] are written as zero
  float* out_f32;           // [M_total, ldo] or null
  __half* out_hi;           // [M_total, ldo] or null
  __half* out_lo;           // [M_total, ldo] or null (requires out_hi)
  int ldo;
  int act;                  // 0 none, 1 erf-GELU
  float col_scale;          // columns < col_scale_limit are multiplied by col_scale after the bias
  int col_scale_limit;
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : 6;
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemColsPerStage = (BLOCK_N <= 32) ? 32 : (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;
  static constexpr int kTmemCols = 2 * kTmemColsPerStage;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
               const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int kStages = S::kStages;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N <= 256, "UMMA N constraint for M=128");
  static_assert(S::kBBytes % 1024 == 0, "B stage must keep 1024B alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int tiles_m_per_batch = (p.rows_per_batch + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles_n = p.N / BLOCK_N;
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
    for (int i = 0; i < kStages; ++i) {
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
    tmem_alloc<S::kTmemCols>(tmem_ptr);
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
        const int row0 = (m_tile % tiles_m_per_batch) * GEMM_BLOCK_M + p.row_offset;
        const int a_col0 = p.a_col_per_ntile * n_tile;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pass = kb / p.kb_per_pass;
          const int kk = kb - pass * p.kb_per_pass;
          const int tap = kk / p.kb_per_tap;
          const int kc = kk - tap * p.kb_per_tap;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_3d(smem_a + stage * S::kABytes, (pass == 1) ? &a_lo : &a_hi, &full_bar[stage],
                      a_col0 + kc * GEMM_BLOCK_K, row0 + tap * p.tap_row_step, batch);
          tma_load_2d(smem_b + stage * S::kBBytes, (pass == 2) ? &b_lo : &b_hi, &full_bar[stage],
                      kk * GEMM_BLOCK_K, n_tile * BLOCK_N);
          if (++stage == kStages) {
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
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BLOCK_M, BLOCK_N, 0, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * S::kTmemColsPerStage;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t adesc = make_desc_k_sw128(smem_u32(smem_a + stage * S::kABytes));
          const uint64_t bdesc = make_desc_k_sw128(smem_u32(smem_b + stage * S::kBBytes));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advancing K by 16 fp16 = 32 bytes inside the 128B swizzle row: +2 in 16-byte units
            umma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (++stage == kStages) {
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
    const int half = ew >> 2;                // which half of the tile's columns
    constexpr int kColsPerHalf = (BLOCK_N >= 64) ? BLOCK_N / 2 : BLOCK_N;
    constexpr int kChunk = (kColsPerHalf % 32 == 0) ? 32 : 16;
    constexpr int kChunks = kColsPerHalf / kChunk;
    const bool active = (BLOCK_N >= 64) || (half == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % tiles_n;
      const int m_tile = tile / tiles_n;
      const int batch = m_tile / tiles_m_per_batch;
      const int row_in_batch = (m_tile % tiles_m_per_batch) * GEMM_BLOCK_M + quarter * 32 + (int)lane_id();
      const bool row_ok = row_in_batch < p.rows_per_batch;
      const size_t grow = (size_t)batch * p.rows_per_batch + row_in_batch;
      const bool zero_row = p.valid_rows != nullptr && row_in_batch >= p.valid_rows[batch];

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      if (active) {
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n_tile * BLOCK_N + half * kColsPerHalf + c * kChunk;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                 (uint32_t)(acc * S::kTmemColsPerStage + half * kColsPerHalf + c * kChunk);
          float v[kChunk];
          if constexpr (kChunk == 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          } else {
            uint32_t r[16];
            tmem_ld_32x32b_x16(taddr, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i) {
              float x = v[i];
              if (p.bias) x += __ldg(p.bias + col0 + i);
              if (col0 + i < p.col_scale_limit) x *= p.col_scale;
              if (p.act == 1) x = gelu_erf(x);
              v[i] = x;
            }
            if (p.residual) {
              const float4* rp = reinterpret_cast<const float4*>(p.residual + grow * p.ldo + col0);
#pragma unroll
              for (int i = 0; i < kChunk / 4; ++i) {
                float4 t = __ldg(rp + i);
                v[4 * i + 0] += t.x;
                v[4 * i + 1] += t.y;
                v[4 * i + 2] += t.z;
                v[4 * i + 3] += t.w;
              }
            }
            if (zero_row) {
#pragma unroll
              for (int i = 0; i < kChunk; ++i) v[i] = 0.0f;
            }
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(p.out_f32 + grow * p.ldo + col0);
#pragma unroll
              for (int i = 0; i < kChunk / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.out_hi) {
              uint32_t hi[kChunk / 2], lo[kChunk / 2];
#pragma unroll
              for (int i = 0; i < kChunk / 2; ++i) {
                __half h0, l0, h1, l1;
                split_f16(v[2 * i], h0, l0);
                split_f16(v[2 * i + 1], h1, l1);
                hi[i] = pack_h2(h0, h1);
                lo[i] = pack_h2(l0, l1);
              }
              uint4* hp = reinterpret_cast<uint4*>(p.out_hi + grow * p.ldo + col0);
#pragma unroll
              for (int i = 0; i < kChunk / 8; ++i) hp[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              if (p.out_lo) {
                uint4* lp = reinterpret_cast<uint4*>(p.out_lo + grow * p.ldo + col0);
#pragma unroll
                for (int i = 0; i < kChunk / 8; ++i) lp[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

}  // namespace syl
