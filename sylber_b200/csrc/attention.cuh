// Non-causal multi-head attention with a key-padding mask, head_dim 64, for sm_100a (tcgen05 + TMA).
// Reference arithmetic: transformers HubertAttention.forward (modeling_hubert.py:296-345) calling
// softmax(Q K^T / sqrt(64) + key_mask) V; reached from sylber/model/sylber.py:122.
//
// Persistent kernel, one CTA per SM.  A work item is (utterance, head, PAIR of 128-query tiles); the two query
// tiles share every K/V block that streams in and are processed by two independent softmax warpgroups, so the
// tensor core always has the other tile's MMAs to run while one warpgroup is in its exponentials:
//   warp 0       : TMA producer - Q pair per item, K_j / V_j tiles through two 3-stage mbarrier rings
//   warp 1       : MMA issuer   - in units of 64 keys: S_x(u) = Q_x K_u^T (128x64) and PV_x(u) = P_x(u) V_u (128x64),
//                                 both double buffered in TMEM so the softmax never waits for the tensor core
//   warp 2       : TMEM allocator
//   warps 4..7   : softmax warpgroup of query tile 0        warps 8..11 : softmax warpgroup of query tile 1
//                  one query row per thread: tcgen05.ld S, (mask,) running max / sum in fp32, P -> fp16 into
//                  128B-swizzled smem (A operand of the PV MMA), PV folded into fp32 registers with the running
//                  rescale, final O / l staged through smem and written with TMA stores
// Q arrives pre-scaled by 1/sqrt(64) (exact power of two, folded into the QKV GEMM epilogue).
// The score matrix never leaves the SM: HBM traffic is Q,K,V in and O out, 4*T*768*2 bytes per utterance per
// layer; K/V re-reads by the other query-tile pairs of the same head are L2 hits.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;        // keys per K/V tile (one TMA load each)
constexpr int ATT_UNIT = 64;        // keys per MMA / softmax unit (half a K/V tile)
constexpr int ATT_THREADS = 384;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: any [128 x 64] fp16 tile
constexpr int ATT_SMEM_Q = 0;                                             // 2 tiles
constexpr int ATT_SMEM_K = ATT_SMEM_Q + 2 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_V = ATT_SMEM_K + ATT_KV_STAGES * ATT_TILE_BYTES;
constexpr int ATT_SMEM_P = ATT_SMEM_V + ATT_KV_STAGES * ATT_TILE_BYTES;   // 2 query tiles x 2 key halves
constexpr int ATT_SMEM_BAR = ATT_SMEM_P + 4 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_TOTAL = ATT_SMEM_BAR + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;      // 2 query tiles x 2 buffers x 64 columns
constexpr uint32_t ATT_TMEM_O = 256;    // 2 query tiles x 64 columns: the output accumulator

struct AttnParams {
  int T;                  // frames per utterance (rows per batch item in qkv)
  int batches;
  int heads;
  int model_dim;          // heads * 64
  const int* kv_len;      // [B] number of valid keys per utterance (== T when nothing is padded), or null
  int out_lo;             // also write the fp16 lo part through map o_lo
  int debug;              // timing experiments only (SYL_ATTN_DEBUG): 2 skip exp, 4 skip P store
};

template <bool kMask>
__device__ __forceinline__ void attn_row_max(const uint32_t (&r)[64], int base, int n_valid, float& m) {
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    float s = __uint_as_float(r[i]);
    if (kMask) s = (base + i < n_valid) ? s : -INFINITY;
    m = fmaxf(m, s);
  }
}

template <bool kMask>
__device__ __forceinline__ void attn_row_exp(const uint32_t (&r)[64], int base, int n_valid, float m_scaled, float& l,
                                             uint32_t (&packed)[32]) {
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float p0, p1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(r[2 * i]), kLog2e, -m_scaled)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(r[2 * i + 1]), kLog2e, -m_scaled)));
    if (kMask) {
      p0 = (base + 2 * i < n_valid) ? p0 : 0.0f;
      p1 = (base + 2 * i + 1 < n_valid) ? p1 : 0.0f;
    }
    l += p0 + p1;
    packed[i] = pack_f16x2_sat(p0, p1);
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap o_hi,
                 const __grid_constant__ CUtensorMap o_lo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars;             // [1]
  uint64_t* q_empty = bars + 1;        // [1]
  uint64_t* k_full = bars + 2;         // [3]
  uint64_t* k_empty = bars + 5;        // [3]
  uint64_t* v_full = bars + 8;         // [3]
  uint64_t* v_empty = bars + 11;       // [3]
  uint64_t* s_full = bars + 14;        // [2 query tiles][2 buffers]
  uint64_t* p_full = bars + 18;        // [2][2]
  uint64_t* o_full = bars + 22;        // [2][2]: PV unit completions of the query tile, alternating by unit parity
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5;
  const int q_tiles = (p.T + ATT_BQ - 1) / ATT_BQ;
  const int n_pairs = (q_tiles + 1) / 2;
  const int num_items = p.batches * p.heads * n_pairs;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&qkv_map);
    tma_prefetch_desc(&o_hi);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<ATT_TMEM_COLS>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  auto item_coords = [&](int item, int& b, int& h, int& q0) {
    const int pair = item % n_pairs;
    const int bh = item / n_pairs;
    h = bh % p.heads;
    b = bh / p.heads;
    q0 = pair * 2 * ATT_BQ;
  };
  auto item_kv_len = [&](int b) { return p.kv_len ? max(1, min(__ldg(p.kv_len + b), p.T)) : p.T; };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t kv_it = 0, item_it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++item_it) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        const int n_blocks = (item_kv_len(b) + ATT_BKV - 1) / ATT_BKV;
        mbar_wait(q_empty, (item_it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        tma_load_3d(smem + ATT_SMEM_Q, &qkv_map, q_full, h * ATT_D, q0, b);
        tma_load_3d(smem + ATT_SMEM_Q + ATT_TILE_BYTES, &qkv_map, q_full, h * ATT_D, q0 + ATT_BQ, b);
        for (int j = 0; j < n_blocks; ++j, ++kv_it) {
          const int st = kv_it % ATT_KV_STAGES;
          const uint32_t ph = (kv_it / ATT_KV_STAGES) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
          tma_load_3d(smem + ATT_SMEM_K + st * ATT_TILE_BYTES, &qkv_map, &k_full[st], p.model_dim + h * ATT_D,
                      j * ATT_BKV, b);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
          tma_load_3d(smem + ATT_SMEM_V + st * ATT_TILE_BYTES, &qkv_map, &v_full[st], 2 * p.model_dim + h * ATT_D,
                      j * ATT_BKV, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // Work is issued in UNITS of 64 keys: S_x(u) is 128x64 (double buffered in TMEM per query tile), PV_x(u) is
    // 128x64 over K = 64 keys (double buffered too), so neither softmax warpgroup ever waits for the tensor core.
    // These MMAs are small (32 tensor-core cycles each), so the scalar bookkeeping of this one thread is what
    // bounds the kernel (measured: 100 SASS instructions per 4 MMAs cost more than the softmax).  Hence: every
    // descriptor is a 32-bit base plus a multiply-add, ring positions are running counters (no division), and the
    // 64-bit descriptors are assembled right at the MMA.
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_UNIT, 0, 0, 0);   // Q (K-major) x K (K-major), N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 0, 1);      // P (K-major) x V (MN-major)
      // low words of the shared-memory matrix descriptors (address >> 4, LBO field); high words are constants
      constexpr uint32_t kDescHiK = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61) >> 32);
      const uint32_t q_lo = ((smem_u32(smem + ATT_SMEM_Q) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t k_lo = ((smem_u32(smem + ATT_SMEM_K) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t p_lo = ((smem_u32(smem + ATT_SMEM_P) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t v_lo = ((smem_u32(smem + ATT_SMEM_V) & 0x3FFFF) >> 4) | ((uint32_t)(ATT_TILE_BYTES >> 4) << 16);
      auto desc = [](uint32_t lo) { return ((uint64_t)kDescHiK << 32) | lo; };
      // S_x(unit) = Q_x * K(stage, half)^T  -> TMEM S buffer `buf`
      auto issue_s = [&](int x, uint32_t stage, uint32_t half, uint32_t buf) {
        const uint64_t qd = desc(q_lo + x * 1024);
        const uint64_t kd = desc(k_lo + stage * 1024 + half * 512);
        const uint32_t d = tmem_base + ATT_TMEM_S + x * 128 + buf * 64;
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_f16_ss(d, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[x * 2 + buf]);
      };
      // O_x (+)= P_x(buf) * V(stage, half): the output accumulates in TMEM over all units of the item
      auto issue_pv = [&](int x, uint32_t stage, uint32_t half, uint32_t buf, uint32_t first) {
        const uint64_t pd = desc(p_lo + x * 2048 + buf * 1024);
        const uint64_t vd = desc(v_lo + stage * 1024 + half * 512);
        const uint32_t d = tmem_base + ATT_TMEM_O + x * 64;
#pragma unroll
        for (int kk = 0; kk < ATT_UNIT / 16; ++kk) umma_f16_ss(d, pd + 2 * kk, vd + 128 * kk, idesc_o, (kk != 0) | (first == 0));
        umma_commit(&o_full[x * 2 + buf]);   // two barriers alternate so a waiter may lag two units without aliasing
      };
      // ring cursors: (stage, phase) of the K block the next S unit reads and of the V block the next PV unit reads
      uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
      uint32_t g = 0;            // global unit counter -> TMEM / P buffer (g & 1) and barrier parity ((g >> 1) & 1)
      uint32_t item_par = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_par ^= 1) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        const int U = (item_kv_len(b) + ATT_UNIT - 1) / ATT_UNIT;
        // S units are issued two ahead of the PV units; `su` counts issued S units of this item
        int su = 0;
        mbar_wait(q_full, item_par);
        mbar_wait(&k_full[ks], kph);
        tc_fence_after_sync();
        // prologue: units 0 and 1 (both live in K block 0)
        issue_s(0, ks, 0, g & 1);
        issue_s(1, ks, 0, g & 1);
        su = 1;
        if (U > 1) {
          issue_s(0, ks, 1, (g + 1) & 1);
          issue_s(1, ks, 1, (g + 1) & 1);
          su = 2;
        }
        umma_commit(&k_empty[ks]);                 // K block 0 fully issued
        if (++ks == ATT_KV_STAGES) { ks = 0; kph ^= 1; }
        if (U <= 2) umma_commit(q_empty);          // ... and those were the last S of this item
        for (int u = 0; u < U; ++u, ++g) {
          const uint32_t half = u & 1;
          const uint32_t buf = g & 1;
          const uint32_t par = (g >> 1) & 1;
          if (half == 0) mbar_wait(&v_full[vs], vph);
          const bool more = su < U;                // unit `su` = u + 2 reuses the S buffer that unit u frees
          const uint32_t s_half = su & 1;
          if (more && s_half == 0) mbar_wait(&k_full[ks], kph);
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&p_full[x * 2 + buf], par);  // P_x(u) is in smem, S_x(u) has been consumed
            tc_fence_after_sync();
            issue_pv(x, vs, half, buf, u == 0);
            if (more) issue_s(x, ks, s_half, buf);
          }
          if (half == 1 || u == U - 1) {           // both units of this V block (or the item's last unit) are issued
            umma_commit(&v_empty[vs]);
            if (++vs == ATT_KV_STAGES) { vs = 0; vph ^= 1; }
          }
          if (more) {
            ++su;
            if (s_half == 1 || su == U) {          // K block fully issued
              umma_commit(&k_empty[ks]);
              if (++ks == ATT_KV_STAGES) { ks = 0; kph ^= 1; }
            }
            if (su == U) umma_commit(q_empty);     // last S of this item
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warpgroups
    const int x = (warp - 4) >> 2;            // query tile of the pair
    const int quarter = warp & 3;
    const int lane = (int)lane_id();
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_base = tmem_base + lane_addr + ATT_TMEM_S + x * 128;
    const uint32_t o_addr = tmem_base + lane_addr + ATT_TMEM_O + x * 64;
    uint8_t* pbuf = smem + ATT_SMEM_P + x * 2 * ATT_TILE_BYTES;
    const int row_off = (row >> 3) * 1024 + (row & 7) * 128;   // this row inside a [128 x 64] swizzled tile
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    // TMEM reads, not the tensor core or the exponentials, bound this kernel at head_dim 64 (a 128x64 fp32 tile
    // takes ~512 cycles to read; measured floor with all softmax math removed: 45 of 78 us).  So the output
    // accumulates in TMEM across units (tcgen05.mma accumulate) and is read ONCE per item, instead of folding every
    // PV unit into registers; the running maximum is only refreshed - and O rescaled in TMEM - when a unit's
    // maximum exceeds the reference by more than 2^8 (probabilities stay <= 256, exact in fp16/fp32), which in
    // practice happens in the first units of an item only.  The freed registers double-buffer the score row, so
    // the TMEM read of unit u+1 overlaps the exponentials of unit u.
    constexpr float kRescaleThreshold = 8.0f;
    // Prefetching S(u+1) into registers while unit u computes needs S(u+1) to exist already, but the tensor core
    // only issues it after P(u-1) - measured slower (102 vs 78 us) because every unit then waits for a fresh MMA.
    constexpr bool kPrefetch = false;
    uint32_t g0 = 0;     // global unit counter at the start of the item
    uint32_t ow = 0;     // number of PV completions of this query tile already waited for
    auto wait_pv_upto = [&](uint32_t g_target) {       // PV units complete in order; wait for every phase once
      // unit n completes on barrier (n & 1) as that barrier's phase (n >> 1).  This thread never lags the tensor
      // core by more than two units (PV(u) needs this warpgroup's P(u)), i.e. by one phase per barrier, so the
      // parity test cannot alias.
      while ((int32_t)(g_target - ow) >= 0) {
        mbar_wait(&o_full[x * 2 + (ow & 1)], (ow >> 1) & 1);
        ++ow;
      }
    };
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      const int kv_len = item_kv_len(b);
      const int U = (kv_len + ATT_UNIT - 1) / ATT_UNIT;
      float m_ref = -INFINITY, l_run = 0.0f;

      auto load_s = [&](uint32_t g, uint32_t (&dst)[64]) {           // asynchronous: caller issues tmem_ld_wait()
        mbar_wait(&s_full[x * 2 + (g & 1)], (g >> 1) & 1);
        tc_fence_after_sync();
        tmem_ld_32x32b_x64(s_base + (g & 1) * 64, dst);
      };
      auto process = [&](int u, uint32_t (&cur)[64], uint32_t (&nxt)[64]) {
        const uint32_t g = g0 + u;
        const int n_valid = kv_len - u * ATT_UNIT;      // keys of this unit below the mask (may exceed 64)
        const bool masked = n_valid < ATT_UNIT;
        if (!kPrefetch) load_s(g, cur);
        tmem_ld_wait();                                  // `cur` has landed
        if (kPrefetch && u + 1 < U) load_s(g + 1, nxt);  // prefetch the next unit's scores
        float m_blk = -INFINITY;
        if (masked) attn_row_max<true>(cur, 0, n_valid, m_blk);
        else attn_row_max<false>(cur, 0, n_valid, m_blk);
        const bool need = (m_blk - m_ref) * kLog2e > kRescaleThreshold;   // also true for the first unit (m_ref = -inf)
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, m_blk);
          float scale;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(scale) : "f"((m_ref - m_new) * kLog2e));   // 0 when m_ref = -inf
          if (u > 0) {                                   // rescale the accumulator in TMEM (rare)
            if (kPrefetch && u + 1 < U) tmem_ld_wait();  // keep the prefetch out of the registers reused below
            wait_pv_upto(g - 1);                         // every PV issued so far has completed
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t r[16];
              tmem_ld_32x32b_x16(o_addr + c * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * scale);
              tmem_st_32x32b_x16(o_addr + c * 16, r);
            }
            tmem_st_wait();
          }
          l_run *= scale;
          m_ref = m_new;
        }
        if (u >= 2) wait_pv_upto(g - 2);                 // PV(u-2) has finished reading the P buffer reused now
        const float m_scaled = m_ref * kLog2e;
        float l_blk = 0.0f;
        uint8_t* trow = pbuf + (g & 1) * ATT_TILE_BYTES + row_off;
        // 8 keys (one 16-byte chunk of the swizzled P row) at a time, so the packed values never pile up in registers
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k0 = q * 8 + 2 * i;
            float p0, p1;
            if (p.debug & 2) {
              p0 = __uint_as_float(cur[k0]);
              p1 = __uint_as_float(cur[k0 + 1]);
            } else {
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(cur[k0]), kLog2e, -m_scaled)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(cur[k0 + 1]), kLog2e, -m_scaled)));
            }
            if (masked) {
              p0 = (k0 < n_valid) ? p0 : 0.0f;
              p1 = (k0 + 1 < n_valid) ? p1 : 0.0f;
            }
            l_blk += p0 + p1;
            pk[i] = pack_f16x2_sat(p0, p1);
          }
          if (!(p.debug & 4)) *reinterpret_cast<uint4*>(trow + ((q ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l_run += l_blk;
        fence_proxy_async_smem();      // generic-proxy P writes -> visible to the tensor core's async proxy
        tc_fence_before_sync();        // S loads / O stores are complete before the MMA warp touches the buffers
        mbar_arrive(&p_full[x * 2 + (g & 1)]);
      };

      uint32_t s_a[64], s_b[64];
      if (kPrefetch) load_s(g0, s_a);
      for (int u = 0; u < U; u += 2) {
        process(u, s_a, s_b);
        if (u + 1 < U) process(u + 1, s_b, s_a);
      }
      // the accumulator is complete once the last PV unit has finished
      wait_pv_upto(g0 + U - 1);
      tc_fence_after_sync();
      float o[ATT_D];
      {
        uint32_t r[64];
        tmem_ld_32x32b_x64(o_addr, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = __uint_as_float(r[i]);
      }
      tc_fence_before_sync();
      g0 += U;
      // normalise, stage through this warp's 4 KB slice of the (now idle) P tiles, TMA store
      const float inv_l = 1.0f / l_run;
      const int warp_row0 = q0 + x * ATT_BQ + quarter * 32;
      if (warp_row0 < p.T) {
        uint8_t* st_hi = pbuf + quarter * 4096;
        uint8_t* st_lo = pbuf + ATT_TILE_BYTES + quarter * 4096;
        uint8_t* my_hi = st_hi + lane * 128;
        uint8_t* my_lo = st_lo + lane * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_pair(o[8 * i + 2 * q] * inv_l, o[8 * i + 2 * q + 1] * inv_l, hi[q], lo[q]);
          *reinterpret_cast<uint4*>(my_hi + ((i ^ (lane & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (p.out_lo) *reinterpret_cast<uint4*>(my_lo + ((i ^ (lane & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&o_hi, st_hi, h * ATT_D, warp_row0, b);
          if (p.out_lo) tma_store_3d(&o_lo, st_lo, h * ATT_D, warp_row0, b);
          tma_store_commit();
          tma_store_wait_read();     // the next item's P writes reuse this smem
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace syl
