// Cluster-of-two (tcgen05.mma.cta_group::2) plumbing of the GEMM kernel in gemm3_tc.cuh: thread-block clusters of two
// CTAs on one TPC issue UMMA M=256, N=256, K=16.  Each CTA stages only ITS 128 rows of A and ITS 128 of the 256 B rows
// per k-block (32 KB instead of 48 KB), the tensor cores of the pair exchange the B halves on chip, and each CTA's
// TMEM holds the 128x256 fp32 accumulator of its own rows.  Why: with cta_group::1 a 128x256x64 k-block needs 48 KB of
// TMA writes plus 48 KB of operand reads per 512 tensor-core cycles, i.e. 192 B/cycle against the 128 B/cycle a
// shared memory can move, so a 1-CTA kernel tops out near 930 cycles per k-block (measured in round 1); the pair needs
// 64 KB + 64 KB per 2 SMs = 125 B/cycle per SM.  The smaller stage also buys 6 pipeline stages instead of 4.
//
// Protocol:
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
constexpr int GEMM2_EPI_BYTES = 32768;                          // epilogue staging area: 16 warps x 2 KB
constexpr int GEMM2_SMEM_EPI = GEMM2_STAGES * GEMM2_STAGE_BYTES;
constexpr int GEMM2_SMEM_BIAS = GEMM2_SMEM_EPI + GEMM2_EPI_BYTES;
constexpr int GEMM2_SMEM_BAR = GEMM2_SMEM_BIAS + 2 * GEMM_BLOCK_N * 4;
constexpr int GEMM2_SMEM_PREFIX = GEMM2_SMEM_BAR + 256;         // trimmed mode: prefix sums of computed M tiles per utterance
constexpr int GEMM2_MAX_TRIM_BATCHES = 127;
constexpr int GEMM2_SMEM_TOTAL = GEMM2_SMEM_PREFIX + (GEMM2_MAX_TRIM_BATCHES + 1) * 4;
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

}  // namespace syl
