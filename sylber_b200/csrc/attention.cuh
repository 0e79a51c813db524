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
constexpr uint32_t ATT_TMEM_O = 256;    // 2 query tiles x 2 buffers x 64 columns

struct AttnParams {
  int T;                  // frames per utterance (rows per batch item in qkv)
  int batches;
  int heads;
  int model_dim;          // heads * 64
  const int* kv_len;      // [B] number of valid keys per utterance (== T when nothing is padded), or null
  int out_lo;             // also write the fp16 lo part through map o_lo
};

template <bool kMask>
__device__ __forceinline__ void attn_row_max(const uint32_t (&r)[32], int base, int n_valid, float& m) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float s = __uint_as_float(r[i]);
    if (kMask) s = (base + i < n_valid) ? s : -INFINITY;
    m = fmaxf(m, s);
  }
}

template <bool kMask>
__device__ __forceinline__ void attn_row_exp(const uint32_t (&r)[32], int base, int n_valid, float m_scaled, float& l,
                                             uint32_t (&packed)[16]) {
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
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
  uint64_t* o_full = bars + 22;        // [2][2]
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
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_UNIT, 0, 0, 0);   // Q (K-major) x K (K-major), N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 0, 1);      // P (K-major) x V (MN-major)
      uint32_t blk0 = 0, item_it = 0, g0 = 0;   // global K/V block and unit counters at the start of the item
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++item_it) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        const int kv_len = item_kv_len(b);
        const int U = (kv_len + ATT_UNIT - 1) / ATT_UNIT;
        const int NB = (U + 1) / 2;
        auto issue_s = [&](int x, int u) {
          const uint32_t g = g0 + u;
          const int st = (blk0 + (u >> 1)) % ATT_KV_STAGES;
          const uint64_t qdesc = make_desc_k_sw128(smem_u32(smem + ATT_SMEM_Q + x * ATT_TILE_BYTES));
          const uint64_t kdesc = make_desc_k_sw128(smem_u32(smem + ATT_SMEM_K + st * ATT_TILE_BYTES + (u & 1) * (ATT_UNIT * 128)));
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_f16_ss(tmem_base + ATT_TMEM_S + x * 128 + (g & 1) * 64, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[x * 2 + (g & 1)]);
        };
        auto issue_pv = [&](int x, int u) {
          const uint32_t g = g0 + u;
          const int st = (blk0 + (u >> 1)) % ATT_KV_STAGES;
          const uint32_t pbase = smem_u32(smem + ATT_SMEM_P + x * 2 * ATT_TILE_BYTES + (g & 1) * ATT_TILE_BYTES);
          const uint32_t vbase = smem_u32(smem + ATT_SMEM_V + st * ATT_TILE_BYTES + (u & 1) * (ATT_UNIT * 128));
#pragma unroll
          for (int kk = 0; kk < ATT_UNIT / 16; ++kk) {
            const uint64_t pdesc = make_desc_k_sw128(pbase) + 2 * kk;
            const uint64_t vdesc = make_desc_mn_sw128(vbase + kk * 16 * 128, ATT_TILE_BYTES);
            umma_f16_ss(tmem_base + ATT_TMEM_O + x * 128 + (g & 1) * 64, pdesc, vdesc, idesc_o, kk != 0);
          }
          umma_commit(&o_full[x * 2 + (g & 1)]);
        };
        auto k_wait = [&](int blk) { mbar_wait(&k_full[(blk0 + blk) % ATT_KV_STAGES], ((blk0 + blk) / ATT_KV_STAGES) & 1); };
        auto k_release = [&](int blk) { umma_commit(&k_empty[(blk0 + blk) % ATT_KV_STAGES]); };

        mbar_wait(q_full, item_it & 1);
        k_wait(0);
        tc_fence_after_sync();
        for (int u = 0; u < 2 && u < U; ++u) {
          issue_s(0, u);
          issue_s(1, u);
        }
        k_release(0);                       // K block 0 holds units 0 and 1, both issued
        if (U <= 2) umma_commit(q_empty);   // ... and they were the last S of this item
        for (int u = 0; u < U; ++u) {
          const uint32_t g = g0 + u;
          if ((u & 1) == 0) mbar_wait(&v_full[(blk0 + (u >> 1)) % ATT_KV_STAGES], ((blk0 + (u >> 1)) / ATT_KV_STAGES) & 1);
          const int un = u + 2;                       // the unit whose S reuses the buffer freed by unit u
          if (un < U && (un & 1) == 0) k_wait(un >> 1);
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&p_full[x * 2 + (g & 1)], (g >> 1) & 1);   // P_x(u) is in smem, S_x(u) has been consumed
            tc_fence_after_sync();
            issue_pv(x, u);
            if (un < U) issue_s(x, un);
          }
          if ((u & 1) == 1 || u == U - 1) umma_commit(&v_empty[(blk0 + (u >> 1)) % ATT_KV_STAGES]);
          if (un < U) {
            if ((un & 1) == 1 || un == U - 1) k_release(un >> 1);   // both units of that K block have been issued
            if (un == U - 1) umma_commit(q_empty);                  // last S of this item
          }
        }
        blk0 += NB;
        g0 += U;
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
    const uint32_t o_base = tmem_base + lane_addr + ATT_TMEM_O + x * 128;
    uint8_t* pbuf = smem + ATT_SMEM_P + x * 2 * ATT_TILE_BYTES;
    const int row_off = (row >> 3) * 1024 + (row & 7) * 128;   // this row inside a [128 x 64] swizzled tile
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    uint32_t g0 = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      const int kv_len = item_kv_len(b);
      const int U = (kv_len + ATT_UNIT - 1) / ATT_UNIT;
      float o[ATT_D];
#pragma unroll
      for (int i = 0; i < ATT_D; ++i) o[i] = 0.0f;
      float m_run = -INFINITY, l_run = 0.0f, a_prev1 = 0.0f, a_prev2 = 0.0f;

      auto fold = [&](uint32_t g, float alpha) {   // O = O * alpha + PV(g)
        mbar_wait(&o_full[x * 2 + (g & 1)], (g >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_base + (g & 1) * 64 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(r[i]));
        }
      };

      for (int u = 0; u < U; ++u) {
        const uint32_t g = g0 + u;
        const int n_valid = kv_len - u * ATT_UNIT;      // keys of this unit below the mask (may exceed 64)
        const bool masked = n_valid < ATT_UNIT;
        mbar_wait(&s_full[x * 2 + (g & 1)], (g >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t s_addr = s_base + (g & 1) * 64;
        // pass A: row maximum of the unit
        float m_blk = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, r);
          tmem_ld_wait();
          if (masked) attn_row_max<true>(r, c * 32, n_valid, m_blk);
          else attn_row_max<false>(r, c * 32, n_valid, m_blk);
        }
        const float m_new = fmaxf(m_run, m_blk);
        float alpha;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"((m_run - m_new) * kLog2e));   // 0 on the first unit
        const float m_scaled = m_new * kLog2e;
        // fold PV(u-2): frees the O buffer and the P buffer this unit is about to reuse; it was issued a whole
        // unit ago, so this wait is normally already satisfied
        if (u >= 2) fold(g - 2, a_prev2);
        // pass B: probabilities -> fp16 -> swizzled smem, row sum in fp32
        float l_blk = 0.0f;
        uint8_t* trow = pbuf + (g & 1) * ATT_TILE_BYTES + row_off;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, r);
          tmem_ld_wait();
          uint32_t packed[16];
          if (masked) attn_row_exp<true>(r, c * 32, n_valid, m_scaled, l_blk, packed);
          else attn_row_exp<false>(r, c * 32, n_valid, m_scaled, l_blk, packed);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(trow + (((c * 4 + q) ^ sw) << 4)) =
                make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        a_prev2 = a_prev1;
        a_prev1 = alpha;
        fence_proxy_async_smem();      // generic-proxy P writes -> visible to the tensor core's async proxy
        tc_fence_before_sync();        // S / PV loads are complete before the MMA warp may overwrite the buffers
        mbar_arrive(&p_full[x * 2 + (g & 1)]);
      }
      // the last two PV units of the item
      if (U >= 2) fold(g0 + U - 2, a_prev2);
      fold(g0 + U - 1, a_prev1);
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
