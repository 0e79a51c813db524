// Persistent, warp-specialised tcgen05 GEMM on 2-CTA clusters (plumbing: gemm2_tc.cuh) with SIXTEEN epilogue warps
// (640 threads): four epilogue warps per scheduler instead of the two of its round-1 predecessor.
//   warp 0      : TMA producer (one elected lane) - A/B tiles, 128B-swizzled, K-major, 6-stage mbarrier ring
//   warp 1      : MMA issuer (leader CTA, one elected lane) - tcgen05.mma.cta_group::2 kind::f16, fp32 accumulators
//                 in TMEM, two accumulator stages so the epilogue overlaps the next tile
//   warp 2      : TMEM allocator / deallocator
//   warps 4..19 : epilogue - tcgen05.ld (software pipelined), bias / scale / GELU / row mask in registers, staging in
//                 shared memory, TMA stores (fp32 and/or fp16 hi(+lo)); no thread ever issues a global store  With K = 768 (12 k-blocks, ~6 100 tensor cycles per
// tile) the 8-warp epilogue - two warps per scheduler walking 128 columns each through dependent MUFU / convert /
// store chains - took longer than the main loop: QKV ran at 77 %, FFN1 (GELU) at 61-72 % and the out-projection at
// 50 % tensor-pipe activity (profiles/r02_ncu_summary.md).  Each warp now owns a [32 rows x 64 columns] block and
// walks it in 16-column steps so that its state fits the 96 registers per thread of a 640-thread CTA.
// Output staging stays 2 KB per warp: fp32 as {16 col, 32 row} boxes (64-byte rows, SWIZZLE_64B), fp16 hi-only as
// {32, 32} boxes filled by two steps, fp16 hi + lo as unswizzled {16, 32} boxes (1 KB each).
#pragma once

#include "gemm2_tc.cuh"

namespace syl {

constexpr int GEMM3_EPI_WARPS = 16;
constexpr int GEMM3_THREADS = (GEMM_EPI_WARP0 + GEMM3_EPI_WARPS) * 32;   // 640
constexpr int GEMM3_EPI_STAGE_BYTES = 2048;
static_assert(GEMM3_EPI_WARPS * GEMM3_EPI_STAGE_BYTES == GEMM2_EPI_BYTES, "staging area");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM3_THREADS, 1)
gemm3_tc_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
               const __grid_constant__ CUtensorMap o_f32, const __grid_constant__ CUtensorMap o_hi,
               const __grid_constant__ CUtensorMap o_lo, const GemmParams p) {
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    GEMM_TRACE_NS(p, 0);
    GEMM_TRACE(p, 1);
  }
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned stage buffers
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + GEMM2_STAGES * GEMM2_A_BYTES;
  uint8_t* smem_epi = smem + GEMM2_SMEM_EPI;   // 16 warps x 2 KB
  float* smem_bias = reinterpret_cast<float*>(smem + GEMM2_SMEM_BIAS);   // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM2_SMEM_BAR);
  uint64_t* full_bar = bars;                           // [STAGES]  (used in the leader CTA)
  uint64_t* empty_bar = bars + GEMM2_STAGES;           // [STAGES]  per CTA
  uint64_t* tmem_full = bars + 2 * GEMM2_STAGES;       // [2]       per CTA
  uint64_t* tmem_empty = bars + 2 * GEMM2_STAGES + 2;  // [2]       (used in the leader CTA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GEMM2_STAGES + 4);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int warp = threadIdx.x >> 5;
  const int tiles_m_per_batch = (p.rows_per_batch + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);   // 256-row cluster tiles
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
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
    for (int i = 0; i < GEMM2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);            // the leader's arrive.expect_tx; both CTAs' TMA bytes complete_tx on it
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * GEMM3_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta<GEMM_TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before_sync();
  cluster_sync_all();                        // barriers of both CTAs are initialised before anyone touches them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) GEMM_TRACE(p, 2);
  griddep_wait();                            // the predecessor kernel's output (our A operand, valid_rows) is complete
  // Trimmed mode (p.skip_invalid_tiles): M tiles that start in an utterance's padding are not computed.  The tile list
  // every role walks is COMPACTED - cluster c takes the c-th, (c + C)-th, ... computed tile - so the persistent grid
  // stays balanced however the valid lengths are distributed; m_prefix[b] = computed M tiles of utterances < b.
  // (read after griddep_wait: valid_rows is written by a predecessor kernel)
  int* m_prefix = reinterpret_cast<int*>(smem + GEMM2_SMEM_PREFIX);
  const bool compact = p.skip_invalid_tiles != 0;
  int total_tiles = num_tiles;
  if (compact) {
    if (threadIdx.x == 0) {
      int acc_t = 0;
      for (int bb = 0; bb < p.batches; ++bb) {
        m_prefix[bb] = acc_t;
        acc_t += min(tiles_m_per_batch, (max(__ldg(p.valid_rows + bb), 0) + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M));
      }
      m_prefix[p.batches] = acc_t;
    }
    __syncthreads();
    total_tiles = m_prefix[p.batches] * tiles_n;
  }
  // tile index -> (utterance, M tile inside it); `cur` is the caller's monotonically advancing utterance cursor
  auto decode_m = [&](int m_idx, int& cur, int& batch, int& m_in_batch) {
    if (compact) {
      while (m_idx >= m_prefix[cur + 1]) ++cur;
      batch = cur;
      m_in_batch = m_idx - m_prefix[cur];
    } else {
      batch = m_idx / tiles_m_per_batch;
      m_in_batch = m_idx % tiles_m_per_batch;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int cur = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int n_tile = tile % tiles_n;
        int batch, m_in_batch;
        decode_m(tile / tiles_n, cur, batch, m_in_batch);
        const int row0 = m_in_batch * 2 * GEMM_BLOCK_M + (int)cta_rank * GEMM_BLOCK_M;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pass = kb / p.kb_per_pass;
          const int kk = kb - pass * p.kb_per_pass;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);   // bytes of BOTH CTAs' loads
          // split precision: the two correction passes (A_lo B_hi, A_hi B_lo) come FIRST and A_hi B_hi last.  The tensor
          // core adds each K = 16 step into the fp32 accumulator with truncation, i.e. a bias of ~half an ulp of the running
          // sum per step (measured: conv1, 288 steps, 9.2e-6 relative = 288 x 3e-8, profiles/r03_precision.md); while the
          // correction passes run the sum is 2^-11 of its final size, so only the last pass's steps cost a full ulp.
          const bool a_is_lo = p.n_pass == 3 && pass == 0, b_is_lo = p.n_pass == 3 && pass == 1;
          tma_load_3d_2sm(smem_a + stage * GEMM2_A_BYTES, a_is_lo ? &a_lo : &a_hi, full_leader,
                          kk * GEMM_BLOCK_K, row0, batch);
          tma_load_2d_2sm(smem_b + stage * GEMM2_B_BYTES, b_is_lo ? &b_lo : &b_hi, full_leader,
                          kk * GEMM_BLOCK_K, n_tile * GEMM_BLOCK_N + (int)cta_rank * 128);
          if (tile == cluster_id && kb == 0) GEMM_TRACE(p, 3);
          if (++stage == GEMM2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(2 * GEMM_BLOCK_M, GEMM_BLOCK_N, 0, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        GEMM_TRACE(p, 8 + 3 * ((tile - cluster_id) / num_clusters));
        const uint32_t tmem_d = tmem_base + acc * GEMM_BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (kb == 0) GEMM_TRACE(p, 9 + 3 * ((tile - cluster_id) / num_clusters));
          const uint64_t adesc = make_desc_k_sw128(smem_u32(smem_a + stage * GEMM2_A_BYTES));
          const uint64_t bdesc = make_desc_k_sw128(smem_u32(smem_b + stage * GEMM2_B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advancing K by 16 fp16 = 32 bytes inside the 128B swizzle row: +2 in 16-byte units
            umma_f16_ss_2cta(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_2cta(&empty_bar[stage]);  // frees this smem stage in both CTAs once the MMAs have read it
          if (++stage == GEMM2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tmem_full[acc]);  // accumulator complete -> epilogue warps of both CTAs
        GEMM_TRACE(p, 10 + 3 * ((tile - cluster_id) / num_clusters));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= GEMM_EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue
    // 16 epilogue warps: warp (quarter, cb) owns TMEM lanes [32 quarter, +32) x columns [64 cb, +64) of the tile and
    // walks them in four steps of 16 columns (double buffered tcgen05.ld.x16), so that the whole epilogue state fits
    // the 96 registers a 640-thread CTA leaves per thread.
    const int ew = warp - GEMM_EPI_WARP0;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int cb = ew >> 2;                  // which 64-column block of the tile
    const int lane = (int)lane_id();
    const int epi_tid = threadIdx.x - GEMM_EPI_WARP0 * 32;   // 0..511
    uint8_t* stage_buf = smem_epi + ew * GEMM3_EPI_STAGE_BYTES;      // 2 KB
    const int sw64 = (lane >> 1) & 3;        // 64-byte rows, SWIZZLE_64B: 16-byte chunk index ^ ((row >> 1) & 3)
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    // the staging buffer may be overwritten once the previous bulk store has finished reading it
    auto acquire = [&]() {
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    };
    int cur = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int n_tile = tile % tiles_n;
      int batch, m_in_batch;
      decode_m(tile / tiles_n, cur, batch, m_in_batch);
      const int warp_row0 = m_in_batch * 2 * GEMM_BLOCK_M + (int)cta_rank * GEMM_BLOCK_M + quarter * 32;
      const int row_in_batch = warp_row0 + lane;
      const bool warp_ok = warp_row0 < p.rows_per_batch;
      // The cluster's LAST tile: its epilogue is the exposed tail of the launch (tools/gemm_trace.py: ~4 200 cycles, most of
      // them the four serial "store has finished reading the staging buffer" waits), and the main loop's operand ring is
      // idle by then - every step gets its own 2 KB of it (16 warps x 4 steps = 128 KB of the 192 KB) and never waits.
      const bool last_tile = tile + num_clusters >= total_tiles;
      const bool zero_row = p.valid_rows != nullptr && row_in_batch >= __ldg(p.valid_rows + batch);
      const float scale = (n_tile * GEMM_BLOCK_N < p.col_scale_limit) ? p.col_scale : 1.0f;
      // stage this tile's bias (pre-multiplied by the column scale, a power of two) in shared memory, double
      // buffered by tile parity
      float* sbias = smem_bias + (it & 1) * GEMM_BLOCK_N;
      if (epi_tid < GEMM_BLOCK_N) sbias[epi_tid] = p.bias ? __ldg(p.bias + n_tile * GEMM_BLOCK_N + epi_tid) * scale : 0.0f;
      named_bar_sync(1, GEMM3_EPI_WARPS * 32);
      const f32x2 scale2 = pack2(scale, scale);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      if (ew == 0 && lane == 0) GEMM_TRACE(p, 64 + 2 * it);
#ifdef SYL_DIAG
      if (p.epi_skip & 2) {      // timing experiment: the accumulator is released unread
        tc_fence_before_sync();
        __syncwarp();
        if (ew == 0 && lane == 0) GEMM_TRACE(p, 65 + 2 * it);
        if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        ++it;
        continue;
      }
#endif
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * GEMM_BLOCK_N + cb * 64);
      uint32_t r[2][16];
      tmem_ld_32x32b_x16(taddr0, r[0]);
      tmem_ld_wait();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s + 1 < 4) tmem_ld_32x32b_x16(taddr0 + (s + 1) * 16, r[(s + 1) & 1]);   // prefetch the next 16 columns
        // staging buffer of this step (hi-only output: steps 2 k and 2 k + 1 fill one buffer)
        uint8_t* const sbuf = last_tile ? smem + (ew * 4 + s) * GEMM3_EPI_STAGE_BYTES : stage_buf;
        uint8_t* const sbuf2 = last_tile ? smem + (ew * 4 + (s & ~1)) * GEMM3_EPI_STAGE_BYTES : stage_buf;
        const uint32_t sbuf_u32 = smem_u32(sbuf), sbuf2_u32 = smem_u32(sbuf2);
        const int col0 = n_tile * GEMM_BLOCK_N + cb * 64 + s * 16;
        float v[16];
        const float4* b4 = reinterpret_cast<const float4*>(sbias + cb * 64 + s * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = b4[i];
          // (acc + bias) * scale == acc * scale + bias * scale exactly: scale is a power of two
          unpack2(fma2(pack2(__uint_as_float(r[s & 1][4 * i + 0]), __uint_as_float(r[s & 1][4 * i + 1])), scale2, pack2(bb.x, bb.y)),
                  v[4 * i + 0], v[4 * i + 1]);
          unpack2(fma2(pack2(__uint_as_float(r[s & 1][4 * i + 2]), __uint_as_float(r[s & 1][4 * i + 3])), scale2, pack2(bb.z, bb.w)),
                  v[4 * i + 2], v[4 * i + 3]);
        }
        if (p.act == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) gelu_fast2(v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1]);
        }
        if (zero_row) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
#ifdef SYL_DIAG
        if (p.epi_skip & 1) {    // timing experiment: TMEM is read and the epilogue math done, nothing is staged or stored
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("" ::"f"(v[i]));
        } else
#endif
        if (warp_ok) {
          if (p.out_f32) {
            // 16 fp32 columns = 64-byte rows, 2 KB: box {16, 32}, SWIZZLE_64B
            if (!last_tile) acquire();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_shared_v4(sbuf_u32 + lane * 64 + ((i ^ sw64) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                           __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
            fence_proxy_async_smem();
            __syncwarp();
#ifdef SYL_DIAG
            if (lane == 0 && !(p.epi_skip & 4)) {  // 4 = timing experiment: staged in shared memory, never stored
#else
            if (lane == 0) {
#endif
              tma_store_3d(&o_f32, sbuf, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
          if (p.out_hi && !p.out_lo && !p.out_f32) {
            // hi only: two steps fill 64-byte rows of 32 fp16 columns (2 KB), one bulk store {32, 32}, SWIZZLE_64B
            uint32_t hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) hi[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
            if ((s & 1) == 0 && !last_tile) acquire();
            const int c0 = (s & 1) * 2;                        // 16-byte chunk of the 64-byte row
            st_shared_v4(sbuf2_u32 + lane * 64 + (((c0 + 0) ^ sw64) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(sbuf2_u32 + lane * 64 + (((c0 + 1) ^ sw64) << 4), hi[4], hi[5], hi[6], hi[7]);
            if ((s & 1) == 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&o_hi, sbuf2, col0 - 16, warp_row0, batch);
                tma_store_commit();
              }
            }
          } else if (p.out_hi) {
            // hi (+ lo) next to another destination: 16 columns = 32-byte rows, 1 KB (+ 1 KB), unswizzled boxes {16, 32}
            uint32_t hi[8], lo[8];
            if (p.out_lo) {
#pragma unroll
              for (int i = 0; i < 8; ++i) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) hi[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
            }
            // (next to an fp32 destination the fp32 store of this step has just been issued from sbuf: wait for it)
            if (!last_tile || p.out_f32) acquire();
            st_shared_v4(sbuf_u32 + lane * 32, hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(sbuf_u32 + lane * 32 + 16, hi[4], hi[5], hi[6], hi[7]);
            if (p.out_lo) {
              st_shared_v4(sbuf_u32 + 1024 + lane * 32, lo[0], lo[1], lo[2], lo[3]);
              st_shared_v4(sbuf_u32 + 1024 + lane * 32 + 16, lo[4], lo[5], lo[6], lo[7]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&o_hi, sbuf, col0, warp_row0, batch);
              if (p.out_lo) tma_store_3d(&o_lo, sbuf + 1024, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
        }
        if (s + 1 < 4) tmem_ld_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (ew == 0 && lane == 0) GEMM_TRACE(p, 65 + 2 * it);
      if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
      ++it;                                   // bias staging parity: counts processed tiles only
    }
    if (lane == 0) tma_store_wait_all();   // bulk stores must be complete before the CTA exits
    if (ew == 0 && lane == 0) GEMM_TRACE(p, 100);
  }

  tc_fence_before_sync();
  cluster_sync_all();                        // neither CTA may exit (or free TMEM) while its peer still uses it
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2cta<GEMM_TMEM_COLS>(tmem_base);
  }
  if (threadIdx.x == 0) {
    GEMM_TRACE(p, 101);
    GEMM_TRACE_NS(p, 102);
  }
}


}  // namespace syl
