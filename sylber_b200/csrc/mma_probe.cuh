// Micro-benchmark (diagnostic entry point syl_mma_probe): cycles per tcgen05.mma when one thread issues a long
// back-to-back stream of M=128 x N x K=16 fp16 MMAs on operands that stay in shared memory.  Answers "what is the
// dispatch floor of a small MMA?", which decides how attention and the positional conv should size their MMAs.
#pragma once

#include "common.cuh"

namespace syl {

template <int N>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 1 && elect_one()) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0, 0, 0);
    const uint64_t ad = make_desc_k_sw128(smem_u32(smem));
    const uint64_t bd = make_desc_k_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + (uint32_t)((i & 1) * 256), ad + 2 * k, bd + 2 * k, idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles_out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace syl
