// 2-CTA variant of gemm_tc_kernel: thread-block clusters of two CTAs on one TPC issue tcgen05.mma.cta_group::2
// (UMMA M=256, N=256, K=16).  Each CTA stages only ITS 128 rows of A and ITS 128 of the 256 B rows per k-block
// (32 KB instead of 48 KB), the tensor cores of the pair exchange the B halves on chip, and each CTA's TMEM holds
// the 128x256 fp32 accumulator of its own rows.  Why: with cta_group::1 a 128x256x64 k-block needs 48 KB of
// TMA writes plus 48 KB of operand reads per 512 tensor-core cycles, i.e. 192 B/cycle against the 128 B/cycle a
// shared memory can move, so the 1-CTA kernel tops out near 930 cycles per k-block (measured); the pair needs
// 64 KB + 64 KB per 2 SMs = 125 B/cycle per SM.  The smaller stage also buys 6 pipeline stages instead of 4.
//
// Protocol differences from the 1-CTA kernel (same roles, same epilogue):
//   * full barriers live in the leader CTA (rank 0): both producers' TMA loads complete_tx on it (the leader's
//     expect_tx covers the bytes of both CTAs, the peer producer does not arrive at all); only the leader's MMA
//     thread issues MMAs;
//   * tcgen05.commit multicasts to both CTAs: smem-empty and accumulator-full barriers are per CTA;
//   * accumulator-empty is the leader's barrier, the peer's epilogue warps arrive remotely.
#pragma once

#include "gemm_tc.cuh"

namespace syl {

constexpr int GEMM2_A_BYTES = 128 * GEMM_BLOCK_K * 2;   // 16 KB: this CTA's 128 rows of the 256-row M tile
constexpr int GEMM2_B_BYTES = 128 * GEMM_BLOCK_K * 2;   // 16 KB: this CTA's 128 of the 256 N rows
constexpr int GEMM2_STAGE_BYTES = GEMM2_A_BYTES + GEMM2_B_BYTES;

// A double-buffered epilogue staging (two 4 KB buffers per warp, 5 smem stages, wait_group.read 1) was measured in
// round 2 and brought nothing (QKV 0.507 vs 0.474 ms): the epilogue is not waiting for its bulk stores.
constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_SMEM_EPI = GEMM2_STAGES * GEMM2_STAGE_BYTES;
constexpr int GEMM2_SMEM_BIAS = GEMM2_SMEM_EPI + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES;
constexpr int GEMM2_SMEM_BAR = GEMM2_SMEM_BIAS + 2 * GEMM_BLOCK_N * 4;
constexpr int GEMM2_SMEM_TOTAL = GEMM2_SMEM_BAR + 256;
static_assert(GEMM2_SMEM_TOTAL <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // default semantics (release at CTA scope): a .cluster-scope release compiles to MEMBAR.ALL.GPU + ERRBAR, which
  // stalled the issuing thread for thousands of cycles per arrive (profiles/r01_gemm2_fence.md)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to a barrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit to the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t tmem_base) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
               const __grid_constant__ CUtensorMap o_f32, const __grid_constant__ CUtensorMap o_hi,
               const __grid_constant__ CUtensorMap o_lo, const GemmParams p) {
  griddep_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned stage buffers
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + GEMM2_STAGES * GEMM2_A_BYTES;
  uint8_t* smem_epi = smem + GEMM2_SMEM_EPI;
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
      mbar_init(&tmem_empty[i], 2 * GEMM_EPI_WARPS);
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
  griddep_wait();                            // the predecessor kernel's output (our A operand, valid_rows) is complete

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int n_tile = tile % tiles_n;
        const int m_tile = tile / tiles_n;
        const int batch = m_tile / tiles_m_per_batch;
        const int row0 = (m_tile % tiles_m_per_batch) * 2 * GEMM_BLOCK_M + (int)cta_rank * GEMM_BLOCK_M;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pass = kb / p.kb_per_pass;
          const int kk = kb - pass * p.kb_per_pass;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);   // bytes of BOTH CTAs' loads
          tma_load_3d_2sm(smem_a + stage * GEMM2_A_BYTES, (pass == 1) ? &a_lo : &a_hi, full_leader,
                          kk * GEMM_BLOCK_K, row0, batch);
          tma_load_2d_2sm(smem_b + stage * GEMM2_B_BYTES, (pass == 2) ? &b_lo : &b_hi, full_leader,
                          kk * GEMM_BLOCK_K, n_tile * GEMM_BLOCK_N + (int)cta_rank * 128);
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
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * GEMM_BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
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
    uint8_t* stage_base = smem_epi + ew * GEMM_EPI_STAGE_BYTES;
    const int sw128 = lane & 7;
    const int sw64 = (lane >> 1) & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    // staging buffer for the next bulk store, free to be overwritten when this returns
    auto acquire = [&]() -> uint8_t* {
      if (lane == 0) tma_store_wait_read();   // the previous bulk store has finished reading it
      __syncwarp();
      return stage_base;
    };
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int n_tile = tile % tiles_n;
      const int m_tile = tile / tiles_n;
      const int batch = m_tile / tiles_m_per_batch;
      const int warp_row0 = (m_tile % tiles_m_per_batch) * 2 * GEMM_BLOCK_M + (int)cta_rank * GEMM_BLOCK_M + quarter * 32;
      const int row_in_batch = warp_row0 + lane;
      const bool warp_ok = warp_row0 < p.rows_per_batch;
      const bool zero_row = p.valid_rows != nullptr && row_in_batch >= __ldg(p.valid_rows + batch);
      const float scale = (n_tile * GEMM_BLOCK_N < p.col_scale_limit) ? p.col_scale : 1.0f;
      // stage this tile's bias (pre-multiplied by the column scale, a power of two) in shared memory: one column per
      // epilogue thread, double buffered by tile parity
      float* sbias = smem_bias + (it & 1) * GEMM_BLOCK_N;
      sbias[epi_tid] = p.bias ? __ldg(p.bias + n_tile * GEMM_BLOCK_N + epi_tid) * scale : 0.0f;
      named_bar_sync(1, GEMM_EPI_WARPS * 32);
      const f32x2 scale2 = pack2(scale, scale);

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
          // (acc + bias) * scale == acc * scale + bias * scale exactly: scale is a power of two
          unpack2(fma2(pack2(__uint_as_float(r[c & 1][4 * i + 0]), __uint_as_float(r[c & 1][4 * i + 1])), scale2, pack2(bb.x, bb.y)),
                  v[4 * i + 0], v[4 * i + 1]);
          unpack2(fma2(pack2(__uint_as_float(r[c & 1][4 * i + 2]), __uint_as_float(r[c & 1][4 * i + 3])), scale2, pack2(bb.z, bb.w)),
                  v[4 * i + 2], v[4 * i + 3]);
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
            uint8_t* buf = acquire();
            uint8_t* row128 = buf + lane * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(row128 + ((i ^ sw128) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&o_f32, buf, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
          if (p.out_hi) {
            uint32_t hi[16], lo[16];
            if (p.out_lo) {
#pragma unroll
              for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) hi[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
            }
            uint8_t* buf = acquire();
            uint8_t* row64_hi = buf + lane * 64;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(row64_hi + ((i ^ sw64) << 4)) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            if (p.out_lo) {
              uint8_t* row64_lo = buf + 2048 + lane * 64;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<uint4*>(row64_lo + ((i ^ sw64) << 4)) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&o_hi, buf, col0, warp_row0, batch);
              if (p.out_lo) tma_store_3d(&o_lo, buf + 2048, col0, warp_row0, batch);
              tma_store_commit();
            }
          }
        }
        if (c + 1 < 4) tmem_ld_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all();   // bulk stores must be complete before the CTA exits
  }

  tc_fence_before_sync();
  cluster_sync_all();                        // neither CTA may exit (or free TMEM) while its peer still uses it
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2cta<GEMM_TMEM_COLS>(tmem_base);
  }
}


}  // namespace syl
