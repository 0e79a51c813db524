// Non-causal multi-head attention with a key-padding mask, head_dim 64, for sm_100a (tcgen05 + TMA).
// Reference arithmetic: transformers HubertAttention.forward (modeling_hubert.py:296-345) calling
// softmax(Q K^T / sqrt(64) + key_mask) V; reached from sylber/model/sylber.py:122.
//
// Persistent kernel, one CTA per SM.  A work item is (utterance, head, FOUR 128-query tiles = 512 queries); the
// tiles share every K/V block that streams in and each has its own softmax warpgroup.  Round-1 measurements
// (profiles/r01_attention_experiments.md) showed that with one query row per thread the kernel is bound by the
// exposed per-unit dependency chain of the softmax warps, not by the tensor core, the SFU, TMEM bandwidth or the
// MMA-issuing thread; four warpgroups put four softmax warps on every scheduler to hide that chain.
//   warp 0        : TMA producer - 4 Q tiles per item, K_j / V_j tiles (128 keys) through two 2-stage rings
//   warp 1        : MMA issuer   - in units of 64 keys: S_x(u) = Q_x K_u^T (128x64 fp32 in TMEM) and
//                                  O_x += P_x(u) V_u (128x64 fp32, accumulated in TMEM over the whole item)
//   warp 2        : TMEM allocator
//   warps 4..19   : softmax warpgroup x = (warp-4)/4, one query row per thread: tcgen05.ld S, (mask,) fp32 max,
//                   exp2, P -> fp16 into 128B-swizzled smem (A operand of the PV MMA); the running maximum is
//                   refreshed lazily (O rescaled in TMEM only when a unit's max exceeds the reference by > 2^8);
//                   at the end of the item O / l is staged through smem and written with TMA stores
// Ordering that keeps S, P and O single buffered per tile: the tensor core executes this CTA's MMAs in issue
// order, and PV_x(u-1) is always issued before S_x(u), so once a warpgroup sees S_x(u) complete it also knows
// that PV_x(u-1) has finished reading P_x and updating O_x.
// Q arrives pre-scaled by 1/sqrt(64) (exact power of two, folded into the QKV GEMM epilogue).
// The score matrix never leaves the SM: HBM traffic is Q,K,V in and O out, 4*T*768*2 bytes per utterance per
// layer; K/V re-reads by other items of the same head are L2 hits.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_QT = 4;           // query tiles per work item / softmax warpgroups per CTA
constexpr int ATT_BKV = 128;        // keys per K/V tile (one TMA load each)
constexpr int ATT_UNIT = 64;        // keys per MMA / softmax unit (half a K/V tile)
constexpr int ATT_THREADS = 128 + ATT_QT * 128;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: any [128 x 64] fp16 tile
constexpr int ATT_SMEM_Q = 0;                                             // 4 tiles
constexpr int ATT_SMEM_K = ATT_SMEM_Q + ATT_QT * ATT_TILE_BYTES;
constexpr int ATT_SMEM_V = ATT_SMEM_K + ATT_KV_STAGES * ATT_TILE_BYTES;
constexpr int ATT_SMEM_P = ATT_SMEM_V + ATT_KV_STAGES * ATT_TILE_BYTES;   // one P tile per query tile
constexpr int ATT_SMEM_BAR = ATT_SMEM_P + ATT_QT * ATT_TILE_BYTES;
constexpr int ATT_SMEM_TOTAL = ATT_SMEM_BAR + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;      // 4 query tiles x 64 columns
constexpr uint32_t ATT_TMEM_O = 256;    // 4 query tiles x 64 columns: the output accumulators

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
__device__ __forceinline__ void attn_row_max(const uint32_t (&r)[64], int n_valid, float& m) {
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    float s = __uint_as_float(r[i]);
    if (kMask) s = (i < n_valid) ? s : -INFINITY;
    m = fmaxf(m, s);
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
  uint64_t* k_full = bars + 2;         // [2]
  uint64_t* k_empty = bars + 4;        // [2]
  uint64_t* v_full = bars + 6;         // [2]
  uint64_t* v_empty = bars + 8;        // [2]
  uint64_t* s_full = bars + 10;        // [4] per query tile, one phase per unit
  uint64_t* p_full = bars + 14;        // [4] per query tile, one phase per unit (128 arrivals)
  uint64_t* o_done = bars + 18;        // [4] per query tile, one phase per item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int q_tiles = (p.T + ATT_BQ - 1) / ATT_BQ;
  const int n_groups = (q_tiles + ATT_QT - 1) / ATT_QT;
  const int num_items = p.batches * p.heads * n_groups;

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
    for (int i = 0; i < ATT_QT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<ATT_TMEM_COLS>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  auto item_coords = [&](int item, int& b, int& h, int& q0) {
    const int grp = item % n_groups;
    const int bh = item / n_groups;
    h = bh % p.heads;
    b = bh / p.heads;
    q0 = grp * ATT_QT * ATT_BQ;
  };
  auto item_kv_len = [&](int b) { return p.kv_len ? max(1, min(__ldg(p.kv_len + b), p.T)) : p.T; };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t ks = 0, kph = 0, item_par = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_par ^= 1) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        const int n_blocks = (item_kv_len(b) + ATT_BKV - 1) / ATT_BKV;
        mbar_wait(q_empty, item_par ^ 1);
        mbar_arrive_expect_tx(q_full, ATT_QT * ATT_TILE_BYTES);
#pragma unroll
        for (int x = 0; x < ATT_QT; ++x)
          tma_load_3d(smem + ATT_SMEM_Q + x * ATT_TILE_BYTES, &qkv_map, q_full, h * ATT_D, q0 + x * ATT_BQ, b);
        for (int j = 0; j < n_blocks; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], ATT_TILE_BYTES);
          tma_load_3d(smem + ATT_SMEM_K + ks * ATT_TILE_BYTES, &qkv_map, &k_full[ks], p.model_dim + h * ATT_D,
                      j * ATT_BKV, b);
          mbar_wait(&v_empty[ks], kph ^ 1);
          mbar_arrive_expect_tx(&v_full[ks], ATT_TILE_BYTES);
          tma_load_3d(smem + ATT_SMEM_V + ks * ATT_TILE_BYTES, &qkv_map, &v_full[ks], 2 * p.model_dim + h * ATT_D,
                      j * ATT_BKV, b);
          if (++ks == ATT_KV_STAGES) { ks = 0; kph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_UNIT, 0, 0, 0);   // Q (K-major) x K (K-major), N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 0, 1);      // P (K-major) x V (MN-major)
      // low words of the shared-memory matrix descriptors (address >> 4, LBO field); the high word is a constant
      constexpr uint32_t kDescHi = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61) >> 32);
      const uint32_t q_lo = ((smem_u32(smem + ATT_SMEM_Q) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t k_lo = ((smem_u32(smem + ATT_SMEM_K) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t p_lo = ((smem_u32(smem + ATT_SMEM_P) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t v_lo = ((smem_u32(smem + ATT_SMEM_V) & 0x3FFFF) >> 4) | ((uint32_t)(ATT_TILE_BYTES >> 4) << 16);
      auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
      uint32_t ks = 0, kph = 0, vs = 0, vph = 0;   // ring cursors of the K block S reads / the V block PV reads
      uint32_t g = 0;                              // global unit counter: barrier parity of s_full / p_full
      uint32_t item_par = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_par ^= 1) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        const int U = (item_kv_len(b) + ATT_UNIT - 1) / ATT_UNIT;
        mbar_wait(q_full, item_par);
        // iteration u issues, for every query tile, PV(u-1) (once its P is ready) and then S(u)
        for (int u = 0; u <= U; ++u) {
          const bool has_s = u < U, has_pv = u > 0;
          if (has_s && (u & 1) == 0) mbar_wait(&k_full[ks], kph);
          if (has_pv && ((u - 1) & 1) == 0) mbar_wait(&v_full[vs], vph);
          if (u == 0) tc_fence_after_sync();
          // serve the query tiles in the order their P becomes ready (polling instead of waiting on tile 0 first):
          // a blocking in-order wait would march the four warpgroups in lock step and serialise softmax and MMA
          uint32_t pending = (1u << ATT_QT) - 1;
          while (pending) {
#pragma unroll
            for (int x = 0; x < ATT_QT; ++x) {
              if (!(pending & (1u << x))) continue;
              if (has_pv) {
                if (!mbar_try_wait(&p_full[x], (g - 1) & 1)) continue;   // P_x(u-1) not in smem yet
                tc_fence_after_sync();
                const uint64_t pd = desc(p_lo + x * 1024);
                const uint64_t vd = desc(v_lo + vs * 1024 + ((u - 1) & 1) * 512);
                const uint32_t d = tmem_base + ATT_TMEM_O + x * 64;
#pragma unroll
                for (int kk = 0; kk < ATT_UNIT / 16; ++kk)
                  umma_f16_ss(d, pd + 2 * kk, vd + 128 * kk, idesc_o, (kk != 0) | (u > 1));
                if (u == U) umma_commit(&o_done[x]);   // last PV of the item: the accumulator is final
              }
              if (has_s) {
                const uint64_t qd = desc(q_lo + x * 1024);
                const uint64_t kd = desc(k_lo + ks * 1024 + (u & 1) * 512);
                const uint32_t d = tmem_base + ATT_TMEM_S + x * 64;
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_f16_ss(d, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[x]);
              }
              pending &= ~(1u << x);
            }
          }
          if (has_pv && (((u - 1) & 1) == 1 || u == U)) {   // both units of the V block (or the last unit) issued
            umma_commit(&v_empty[vs]);
            if (++vs == ATT_KV_STAGES) { vs = 0; vph ^= 1; }
          }
          if (has_s) {
            if ((u & 1) == 1 || u == U - 1) {               // K block fully issued
              umma_commit(&k_empty[ks]);
              if (++ks == ATT_KV_STAGES) { ks = 0; kph ^= 1; }
            }
            if (u == U - 1) umma_commit(q_empty);           // last S of this item: Q may be overwritten
            ++g;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warpgroups
    const int x = (warp - 4) >> 2;            // query tile of the item
    const int quarter = warp & 3;
    const int lane = (int)lane_id();
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + ATT_TMEM_S + x * 64;
    const uint32_t o_addr = tmem_base + lane_addr + ATT_TMEM_O + x * 64;
    uint8_t* pbuf = smem + ATT_SMEM_P + x * ATT_TILE_BYTES;
    uint8_t* trow = pbuf + (row >> 3) * 1024 + (row & 7) * 128;   // this row inside the [128 x 64] swizzled P tile
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;    // refresh the reference max when a unit exceeds it by more than 2^8
    uint32_t g = 0, item_par = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_par ^= 1) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      const int kv_len = item_kv_len(b);
      const int U = (kv_len + ATT_UNIT - 1) / ATT_UNIT;
      float m_ref = -INFINITY, l_run = 0.0f;
      for (int u = 0; u < U; ++u, ++g) {
        const int n_valid = kv_len - u * ATT_UNIT;      // keys of this unit below the mask (may exceed 64)
        const bool masked = n_valid < ATT_UNIT;
        mbar_wait(&s_full[x], g & 1);                   // S(u) done - and, by issue order, PV(u-1) as well
        tc_fence_after_sync();
        uint32_t s[64];
        tmem_ld_32x32b_x64(s_addr, s);
        tmem_ld_wait();
        float m_blk = -INFINITY;
        if (masked) attn_row_max<true>(s, n_valid, m_blk);
        else attn_row_max<false>(s, n_valid, m_blk);
        const bool need = (m_blk - m_ref) * kLog2e > kRescaleThreshold;   // also true for the first unit (m_ref = -inf)
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, m_blk);
          float scale;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(scale) : "f"((m_ref - m_new) * kLog2e));   // 0 when m_ref = -inf
          if (u > 0) {                                   // rescale the accumulator in TMEM (rare)
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
        const float m_scaled = m_ref * kLog2e;
        float l_blk = 0.0f;
        // 8 keys (one 16-byte chunk of the swizzled P row) at a time, so the packed values never pile up in registers
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k0 = q * 8 + 2 * i;
            float p0, p1;
            if (p.debug & 2) {
              p0 = __uint_as_float(s[k0]);
              p1 = __uint_as_float(s[k0 + 1]);
            } else {
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(s[k0]), kLog2e, -m_scaled)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(s[k0 + 1]), kLog2e, -m_scaled)));
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
        mbar_arrive(&p_full[x]);
      }
      // the accumulator is complete once the last PV unit has finished
      mbar_wait(&o_done[x], item_par);
      tc_fence_after_sync();
      const float inv_l = 1.0f / l_run;
      const int warp_row0 = q0 + x * ATT_BQ + quarter * 32;
      uint8_t* st_buf = pbuf + quarter * 4096;          // this warp's 32 rows of the (now idle) P tile
      uint8_t* my_row = st_buf + lane * 128;
      uint32_t o[64];
      tmem_ld_32x32b_x64(o_addr, o);
      tmem_ld_wait();
      tc_fence_before_sync();
      if (warp_row0 < p.T) {
        // fp16 hi: stage, store; then (split mode) the lo residuals through the same 4 KB
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t hi[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            hi[q] = pack_f16x2_sat(__uint_as_float(o[8 * i + 2 * q]) * inv_l, __uint_as_float(o[8 * i + 2 * q + 1]) * inv_l);
          *reinterpret_cast<uint4*>(my_row + ((i ^ (lane & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&o_hi, st_buf, h * ATT_D, warp_row0, b);
          tma_store_commit();
          tma_store_wait_read();
        }
        __syncwarp();
        if (p.out_lo) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint32_t lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t hh;
              split_pair(__uint_as_float(o[8 * i + 2 * q]) * inv_l, __uint_as_float(o[8 * i + 2 * q + 1]) * inv_l, hh, lo[q]);
            }
            *reinterpret_cast<uint4*>(my_row + ((i ^ (lane & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&o_lo, st_buf, h * ATT_D, warp_row0, b);
            tma_store_commit();
            tma_store_wait_read();     // the next item's P writes reuse this smem
          }
          __syncwarp();
        }
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
