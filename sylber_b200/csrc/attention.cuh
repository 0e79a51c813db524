// Non-causal multi-head attention with a key-padding mask, head_dim 64, for sm_100a (tcgen05 + TMA).
// Reference arithmetic: transformers HubertAttention.forward (modeling_hubert.py:296-345) calling
// softmax(Q K^T / sqrt(64) + key_mask) V; reached from sylber/model/sylber.py:122.
//
// Persistent kernel, one CTA per SM.  A work item is (utterance, head, FOUR 128-query tiles = 512 queries); the
// tiles share every K/V block that streams in (two 2-stage rings of 128-key tiles) and each has its own softmax
// warpgroup.  S = Q K^T and O += P V run on tcgen05 in units of 64 keys: S 128x64 fp32 and the output accumulator
// 128x64 fp32 per tile live in TMEM (4 x 64 + 4 x 64 = all 512 columns), P goes through 128B-swizzled shared
// memory as the A operand of the PV MMA, V is consumed as an MN-major B operand straight from the QKV buffer.
// Q arrives pre-scaled by 1/sqrt(64) (exact power of two, folded into the QKV GEMM epilogue).
// The score matrix never leaves the SM: HBM traffic is Q,K,V in and O out, 4*T*768*2 bytes per utterance per
// layer; K/V re-reads by other items of the same head are L2 hits.
// The round-1 kernel (lock-step units, one MMA thread, 64-register S rows) is described in
// profiles/r01_attention_experiments.md; how it became attention7_kernel is in profiles/r02_attention.md.
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
  int trim;               // trimmed mode: query tiles at or beyond kv_len[b] (padded frames) are not computed
  const int* order;       // trimmed mode: utterance indices sorted by valid length, longest first (null = identity), so that
                          // the statically strided items of one round cost about the same on every CTA
  int debug;              // timing experiments only (SYL_ATTN_DEBUG): 2 skip exp, 4 skip P store
  long long* trace;       // timeline probe (tools/attn_trace.py): CTA 0 logs clock64 stamps, 7 writers x trace_cap
  int trace_cap;
};

// ------------------------------------------------------------------------------------------------------------------
// attention7_kernel: same tiles, TMEM and shared-memory layout as attention_kernel, restructured after timing its
// chains with an in-kernel timeline (tools/attn_trace.py, profiles/r02_attention.md):
//   * the four query tiles of an item advance INDEPENDENTLY: every tile is its own chain
//     S(u) -> softmax -> PV(u) ...; the chains share only the K/V ring (a block is released when the last tile has
//     issued its MMAs on it) and Q tiles have their own full/empty barriers.  In attention_kernel unit u of all four
//     tiles was served before unit u+1 of any, so tensor and MUFU phases alternated instead of overlapping;
//   * TWO MMA-issuing threads (warp 1: tiles 0,1; warp 3: tiles 2,3).  One thread serving four tiles was busy
//     ~770 cycles per tile step (450 issuing 8 MMAs + commits, 320 polling), i.e. the whole 3 400-cycle unit period;
//     four threads (one per tile) need a fifth control warp, which costs 16 registers per thread and was slower;
//   * S(u) is issued BEFORE PV(u-1): "P(u-1) is in shared memory" also means the warpgroup has finished reading
//     S(u-1), so S(u) completes while PV(u-1) still runs; a separate pv_done barrier tells the warpgroup when P and O
//     may be touched again (it waits for it just before its first P store);
//   * the softmax thread walks its 64-column S row in four 16-column chunks (double buffered tcgen05.ld) and
//     computes P against the RUNNING reference maximum while it finds the unit's true maximum in the same pass; only
//     when that exceeds the reference by more than 2^8 the accumulator is rescaled and the pass repeated (S is still
//     in TMEM).  ~70 live registers instead of a spilled 64-register row (640 threads leave 96 per thread);
//   * query tiles that lie entirely beyond T are skipped.
// Warps: 0 TMA producer, 1 and 3 MMA issuers, 2 TMEM allocator, 4..19 softmax warpgroups (one query row per thread).
// ------------------------------------------------------------------------------------------------------------------
constexpr int ATT7_BAR_Q_FULL = 0;     // [4]
constexpr int ATT7_BAR_Q_EMPTY = 4;    // [4]
constexpr int ATT7_BAR_K_FULL = 8;     // [2]
constexpr int ATT7_BAR_K_EMPTY = 10;   // [2]  two arrivals: one per MMA thread
constexpr int ATT7_BAR_V_FULL = 12;    // [2]
constexpr int ATT7_BAR_V_EMPTY = 14;   // [2]  two arrivals
constexpr int ATT7_BAR_S_FULL = 16;    // [4]
constexpr int ATT7_BAR_P_FULL = 20;    // [4]  128 arrivals
constexpr int ATT7_BAR_O_DONE = 24;    // [4]
constexpr int ATT7_BAR_PV_DONE = 28;   // [4]
constexpr int ATT7_BAR_COUNT = 32;
constexpr int ATT7_THREADS = ATT_THREADS;                       // 640
constexpr int ATT7_SMEM_TOTAL = ATT_SMEM_BAR + 512 + 1024;      // 512 B of barriers, 1 KB alignment slack
static_assert(ATT7_BAR_COUNT * 8 + 8 <= 512, "barrier block");
static_assert(ATT_QT == 4, "two MMA threads x two tiles");

// exp2 on the FMA pipe for a pair of arguments (Cody-Waite split + degree-4 minimax polynomial on [-0.5, 0.5],
// relative error 2.7e-6, far below the fp16 rounding of P).  Measured: no gain at 1 pair in 4, slower at 2 in 4 -
// the softmax warps are bound by issue/latency, not by the MUFU (45 % busy) - so the default is 0 pairs; the switch
// (SYL_ATTN_POLY) stays for the record.
__device__ __forceinline__ void exp2_poly2(float a0, float a1, float& p0, float& p1) {
  const f32x2 a = pack2(fmaxf(a0, -125.0f), fmaxf(a1, -125.0f));
  const f32x2 t = add2(a, pack2(12582912.0f, 12582912.0f));            // 1.5 * 2^23: integer part lands in the low mantissa bits
  const f32x2 n = add2(t, pack2(-12582912.0f, -12582912.0f));
  const f32x2 f = fma2(n, pack2(-1.0f, -1.0f), a);                    // fractional part in [-0.5, 0.5]
  f32x2 q = fma2(pack2(0.009570101276040077f, 0.009570101276040077f), f, pack2(0.05591785907745361f, 0.05591785907745361f));
  q = fma2(q, f, pack2(0.240247443318367f, 0.240247443318367f));
  q = fma2(q, f, pack2(0.6931217908859253f, 0.6931217908859253f));
  q = fma2(q, f, pack2(0.9999992847442627f, 0.9999992847442627f));
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(q, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));   // * 2^n through the exponent field
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// Row maximum of a unit (first unit of an item only: there is no reference maximum yet).
__device__ __forceinline__ float att7_row_max(uint32_t s_addr, int n_valid) {
  float mx[2] = {-INFINITY, -INFINITY};
  uint32_t r[2][16];
  tmem_ld_32x32b_x16(s_addr, r[0]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c + 1 < 4) tmem_ld_32x32b_x16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = (c * 16 + j < n_valid) ? __uint_as_float(r[c & 1][j]) : -INFINITY;
      mx[j & 1] = fmaxf(mx[j & 1], v);
    }
    if (c + 1 < 4) tmem_ld_wait();
  }
  return fmaxf(mx[0], mx[1]);
}

// One pass of a softmax thread over its 64-column S row, in four 16-column chunks (double buffered tcgen05.ld):
// the row maximum of the unit, P = exp2((s - m_ref) log2e) as fp16 into the swizzled P row (shared-memory address
// p_row) and its fp32 row sum.  kMasked applies the key-padding mask (keys >= n_valid contribute exp2(-inf) = 0).
// before_store() runs once, before the first P store.
template <bool kMasked, int kPolyPairs, typename BeforeStore>
__device__ __forceinline__ void att7_pass(uint32_t s_addr, uint32_t p_row, int sw, int n_valid, float m_ref, int debug,
                                          float& m_blk, float& l_blk, BeforeStore before_store) {
  constexpr float kLog2e = 1.4426950408889634f;
  const f32x2 l2e2 = pack2(kLog2e, kLog2e);
  const float neg_m = -m_ref * kLog2e;
  const f32x2 negm2 = pack2(neg_m, neg_m);
  f32x2 lsum = pack2(0.0f, 0.0f);
  float mx[2] = {-INFINITY, -INFINITY};
  uint32_t r[2][16];
  tmem_ld_32x32b_x16(s_addr, r[0]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c + 1 < 4) tmem_ld_32x32b_x16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {            // two 16-byte chunks of the swizzled P row per TMEM chunk
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = hq * 8 + 2 * i;           // column inside the TMEM chunk
        const int k0 = c * 16 + j;              // key inside the unit
        float s0 = __uint_as_float(r[c & 1][j]), s1 = __uint_as_float(r[c & 1][j + 1]);
        if (kMasked) {
          if (k0 >= n_valid) s0 = -INFINITY;
          if (k0 + 1 >= n_valid) s1 = -INFINITY;
        }
        mx[0] = fmaxf(mx[0], s0);
        mx[1] = fmaxf(mx[1], s1);
        float a0, a1, p0, p1;
        unpack2(fma2(pack2(s0, s1), l2e2, negm2), a0, a1);
        if (debug & 2) {
          p0 = a0;
          p1 = a1;
        } else if (i >= 4 - kPolyPairs) {       // this pair's exponentials on the FMA pipe
          exp2_poly2(a0, a1, p0, p1);
        } else {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(a1));
        }
        lsum = add2(lsum, pack2(p0, p1));
        pk[i] = pack_f16x2_sat(p0, p1);
      }
      if (c == 0 && hq == 0) before_store();    // PV(u-1) must have finished reading the P tile
      if (!(debug & 4)) st_shared_v4(p_row + (((c * 2 + hq) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
    if (c + 1 < 4) tmem_ld_wait();
  }
  m_blk = fmaxf(mx[0], mx[1]);
  float l0, l1;
  unpack2(lsum, l0, l1);
  l_blk = l0 + l1;
}

template <int kPolyPairs, bool kTrace>
__global__ void __launch_bounds__(ATT7_THREADS, 1)
attention7_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap o_hi,
                  const __grid_constant__ CUtensorMap o_lo, const AttnParams p) {
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars + ATT7_BAR_Q_FULL;
  uint64_t* q_empty = bars + ATT7_BAR_Q_EMPTY;
  uint64_t* k_full = bars + ATT7_BAR_K_FULL;
  uint64_t* k_empty = bars + ATT7_BAR_K_EMPTY;
  uint64_t* v_full = bars + ATT7_BAR_V_FULL;
  uint64_t* v_empty = bars + ATT7_BAR_V_EMPTY;
  uint64_t* s_full = bars + ATT7_BAR_S_FULL;
  uint64_t* p_full = bars + ATT7_BAR_P_FULL;
  uint64_t* o_done = bars + ATT7_BAR_O_DONE;
  uint64_t* pv_done = bars + ATT7_BAR_PV_DONE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + ATT7_BAR_COUNT);

  const int warp = threadIdx.x >> 5;
  const int q_tiles = (p.T + ATT_BQ - 1) / ATT_BQ;
  const int n_groups = (q_tiles + ATT_QT - 1) / ATT_QT;
  const int num_items = p.batches * p.heads * n_groups;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&qkv_map);
    tma_prefetch_desc(&o_hi);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < ATT_QT; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<ATT_TMEM_COLS>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();                  // QKV (and kv_len) written by the predecessor kernels are complete
  // timeline probe (kTrace builds only): writer 1 + x = lane 0 of warpgroup x's first warp, 5 + j = MMA thread j
  int trace_n = 0;
  auto trace = [&](int writer, int kind, int x, int u) {
    if (kTrace && p.trace != nullptr && blockIdx.x == 0 && trace_n < p.trace_cap)
      p.trace[(size_t)writer * p.trace_cap + trace_n++] =
          ((long long)kind << 56) | ((long long)x << 52) | ((long long)(u & 0xfff) << 40) | (clock64() & 0xffffffffffLL);
  };

  // item -> (utterance, head, first query row, number of query tiles that hold at least one row < T)
  auto item_kv_len = [&](int b) { return p.kv_len ? max(1, min(__ldg(p.kv_len + b), p.T)) : p.T; };
  // n_act <= 0 (trimmed mode only): the whole item lies in the utterance's padding and every role skips it
  auto item_coords = [&](int item, int& b, int& h, int& q0, int& n_act) {
    const int grp = item % n_groups;
    const int bh = item / n_groups;
    h = bh % p.heads;
    b = bh / p.heads;
    if (p.order != nullptr) b = __ldg(p.order + b);
    q0 = grp * ATT_QT * ATT_BQ;
    const int tiles_b = p.trim ? (item_kv_len(b) + ATT_BQ - 1) / ATT_BQ : q_tiles;
    n_act = min(ATT_QT, tiles_b - grp * ATT_QT);
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t ks = 0, kph = 0;
      uint32_t qpar = 0;                  // bit x: parity of tile x's next use of its Q buffer
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int b, h, q0, n_act;
        item_coords(item, b, h, q0, n_act);
        if (n_act <= 0) continue;
        const int n_blocks = (item_kv_len(b) + ATT_BKV - 1) / ATT_BKV;
#pragma unroll
        for (int x = 0; x < ATT_QT; ++x) {
          if (x < n_act) {
            mbar_wait(&q_empty[x], ((qpar >> x) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[x], ATT_TILE_BYTES);
            tma_load_3d(smem + ATT_SMEM_Q + x * ATT_TILE_BYTES, &qkv_map, &q_full[x], h * ATT_D, q0 + x * ATT_BQ, b);
            qpar ^= 1u << x;
          }
        }
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
  } else if (warp == 1 || warp == 3) {
    // ---------------------------------------------------------------- MMA issuers: warp 1 tiles 0,1; warp 3 tiles 2,3
    const int mt = warp >> 1;                  // 0 or 1
    const int xa = mt * 2;                     // first tile of this thread
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_UNIT, 0, 0, 0);   // Q (K-major) x K (K-major), N = 64 keys
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 0, 1);      // P (K-major) x V (MN-major)
      constexpr uint32_t kDescHi = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61) >> 32);
      const uint32_t q_lo = (((smem_u32(smem + ATT_SMEM_Q) & 0x3FFFF) >> 4) | (1u << 16)) + xa * 1024;
      const uint32_t k_lo = ((smem_u32(smem + ATT_SMEM_K) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t p_lo = (((smem_u32(smem + ATT_SMEM_P) & 0x3FFFF) >> 4) | (1u << 16)) + xa * 1024;
      const uint32_t v_lo = ((smem_u32(smem + ATT_SMEM_V) & 0x3FFFF) >> 4) | ((uint32_t)(ATT_TILE_BYTES >> 4) << 16);
      auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
      uint32_t gbase[2] = {0, 0};        // per tile: units completed in earlier items (parity of p_full)
      uint32_t ipar = 0;                 // bit j: parity of tile xa + j's q_full for the current item
      int su[2];                         // per tile: next step of the current item (0..U; U + 1 = finished)
      uint32_t blk0 = 0;                 // K/V blocks consumed by earlier items (ring slot = index & 1)
      uint32_t k_seen = 0, v_seen = 0;   // blocks [0, seen) are known to have landed in shared memory
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int b, h, q0, n_act;
        item_coords(item, b, h, q0, n_act);
        if (n_act <= 0) continue;
        const int U = (item_kv_len(b) + ATT_UNIT - 1) / ATT_UNIT;
        const int NB = (U + 1) >> 1;
        const int my_act = max(0, min(2, n_act - xa));
        if (my_act == 0) {
          // both tiles lie beyond T: this thread only takes part in releasing the K/V ring (after the block has
          // landed, so that the arrival is counted in the right phase)
          for (int j = 0; j < NB; ++j) {
            const uint32_t G = blk0 + (uint32_t)j;
            mbar_wait(&k_full[G & 1], (G >> 1) & 1);
            mbar_arrive(&k_empty[G & 1]);
            mbar_wait(&v_full[G & 1], (G >> 1) & 1);
            mbar_arrive(&v_empty[G & 1]);
          }
          blk0 += (uint32_t)NB;
          k_seen = v_seen = blk0;
          continue;
        }
        su[0] = 0;
        su[1] = (my_act > 1) ? 0 : U + 1;
        int remaining = my_act;
        uint32_t kcnt0 = 0, kcnt1 = 0, vcnt0 = 0, vcnt1 = 0;   // tiles of this thread that are done with ring slot 0 / 1
        auto k_ready = [&](uint32_t G) {
          if (G < k_seen) return true;
          if (!mbar_test_wait(&k_full[G & 1], (G >> 1) & 1)) return false;
          k_seen = G + 1;
          return true;
        };
        auto v_ready = [&](uint32_t G) {
          if (G < v_seen) return true;
          if (!mbar_test_wait(&v_full[G & 1], (G >> 1) & 1)) return false;
          v_seen = G + 1;
          return true;
        };
        while (remaining > 0) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = su[j];
            if (u > U) continue;
            const int x = xa + j;
            // step u: S(u) (u < U), then PV(u-1) (u > 0).  P(u-1) in shared memory also means that the warpgroup
            // has finished reading S(u-1), so S(u) goes first and completes while PV(u-1) is still running.
            if (u == 0) {
              if (!mbar_test_wait(&q_full[x], (ipar >> j) & 1)) continue;
            } else {
              if (!mbar_test_wait(&p_full[x], (gbase[j] + (uint32_t)(u - 1)) & 1)) continue;
              if (!v_ready(blk0 + (uint32_t)((u - 1) >> 1))) continue;
            }
            if (u < U && !k_ready(blk0 + (uint32_t)(u >> 1))) continue;
            tc_fence_after_sync();
            trace(5 + mt, 4, x, u);
            if (u < U) {
              const uint32_t G = blk0 + (uint32_t)(u >> 1);
              const uint64_t qd = desc(q_lo + j * 1024);
              const uint64_t kd = desc(k_lo + (G & 1) * 1024 + (u & 1) * 512);
              const uint32_t d = tmem_base + ATT_TMEM_S + x * 64;
#pragma unroll
              for (int k = 0; k < ATT_D / 16; ++k) umma_f16_ss(d, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
              umma_commit(&s_full[x]);
              if ((u & 1) == 1 || u == U - 1) {            // this tile is done with the K block
                uint32_t& cnt = (G & 1) ? kcnt1 : kcnt0;
                if (++cnt == (uint32_t)my_act) {           // ... and so is the thread's other tile
                  umma_commit(&k_empty[G & 1]);
                  cnt = 0;
                }
              }
              if (u == U - 1) umma_commit(&q_empty[x]);    // last S of the item: Q_x may be overwritten
            }
            if (u > 0) {
              const uint32_t G = blk0 + (uint32_t)((u - 1) >> 1);
              const uint64_t pd = desc(p_lo + j * 1024);
              const uint64_t vd = desc(v_lo + (G & 1) * 1024 + ((u - 1) & 1) * 512);
              const uint32_t d = tmem_base + ATT_TMEM_O + x * 64;
#pragma unroll
              for (int kk = 0; kk < ATT_UNIT / 16; ++kk)
                umma_f16_ss(d, pd + 2 * kk, vd + 128 * kk, idesc_o, (kk != 0) | (u > 1));
              if (u == U) umma_commit(&o_done[x]);         // last PV of the item: the accumulator is final
              else umma_commit(&pv_done[x]);               // P and O may be touched again
              if (((u - 1) & 1) == 1 || u == U) {
                uint32_t& cnt = (G & 1) ? vcnt1 : vcnt0;
                if (++cnt == (uint32_t)my_act) {
                  umma_commit(&v_empty[G & 1]);
                  cnt = 0;
                }
              }
            }
            trace(5 + mt, 5, x, u);
            su[j] = u + 1;
            if (u == U) --remaining;
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (j < my_act) gbase[j] += (uint32_t)U;
        ipar ^= (1u << my_act) - 1;
        blk0 += (uint32_t)NB;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warpgroups
    const int x = (warp - 4) >> 2;            // query tile of the item
    const int quarter = warp & 3;
    const int lane = (int)lane_id();
    const int row = quarter * 32 + lane;
    const bool tracer = kTrace && quarter == 0 && lane == 0;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + ATT_TMEM_S + x * 64;
    const uint32_t o_addr = tmem_base + lane_addr + ATT_TMEM_O + x * 64;
    uint8_t* pbuf = smem + ATT_SMEM_P + x * ATT_TILE_BYTES;
    const uint32_t p_row = smem_u32(pbuf + (row >> 3) * 1024 + (row & 7) * 128);   // this row inside the swizzled P tile
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;    // refresh the reference max when a unit exceeds it by more than 2^8
    uint32_t g = 0, item_par = 0, pvc = 0;       // pvc: pv_done phases consumed (one per unit except an item's last)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int b, h, q0, n_act;
      item_coords(item, b, h, q0, n_act);
      if (x >= n_act) continue;                 // this tile lies beyond T: no barrier phase is consumed for it
      const int kv_len = item_kv_len(b);
      const int U = (kv_len + ATT_UNIT - 1) / ATT_UNIT;
      float m_ref = -INFINITY, l_run = 0.0f;
      for (int u = 0; u < U; ++u, ++g) {
        const int n_valid = kv_len - u * ATT_UNIT;      // keys of this unit below the mask (may exceed 64)
        const bool masked = n_valid < ATT_UNIT;
        if (tracer) trace(1 + x, 1, x, u);
        mbar_wait(&s_full[x], g & 1);                   // S(u) is complete
        tc_fence_after_sync();
        if (tracer) trace(1 + x, 2, x, u);
        bool pv_pending = u > 0;                        // PV(u-1) was issued after S(u): its completion is a separate event
        auto wait_pv = [&]() {
          if (pv_pending) {
            mbar_wait(&pv_done[x], pvc & 1);
            tc_fence_after_sync();
            ++pvc;
            pv_pending = false;
          }
        };
        if (u == 0) m_ref = att7_row_max(s_addr, n_valid);   // first unit: its own maximum is the reference
        float m_blk, l_blk;
        for (;;) {
          if (!masked) att7_pass<false, kPolyPairs>(s_addr, p_row, sw, n_valid, m_ref, p.debug, m_blk, l_blk, wait_pv);
          else att7_pass<true, 0>(s_addr, p_row, sw, n_valid, m_ref, p.debug, m_blk, l_blk, wait_pv);
          const bool need = (m_blk - m_ref) * kLog2e > kRescaleThreshold;
          if (!__any_sync(0xffffffffu, need)) break;
          // rare: rescale the accumulator in TMEM and the running sum, then repeat the pass - S is still in TMEM
          const float m_new = fmaxf(m_ref, m_blk);
          float scale;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(scale) : "f"((m_ref - m_new) * kLog2e));
          wait_pv();                                     // PV(u-1) must have finished updating O
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
          l_run *= scale;
          m_ref = m_new;
        }
        l_run += l_blk;
        if (tracer) trace(1 + x, 3, x, u);
        fence_proxy_async_smem();      // generic-proxy P writes -> visible to the tensor core's async proxy
        tc_fence_before_sync();        // S loads / O stores are complete before the MMA warp touches the buffers
        mbar_arrive(&p_full[x]);
        if (tracer) trace(1 + x, 6, x, u);
      }
      // the accumulator is complete once the last PV unit has finished
      mbar_wait(&o_done[x], item_par);
      item_par ^= 1;
      if (tracer) trace(1 + x, 7, x, U);
      tc_fence_after_sync();
      const float inv_l = 1.0f / l_run;
      const int warp_row0 = q0 + x * ATT_BQ + quarter * 32;
      uint8_t* st_buf = pbuf + quarter * 4096;          // this warp's 32 rows of the (now idle) P tile
      const uint32_t my_row = smem_u32(st_buf + lane * 128);
      if (warp_row0 < p.T) {
        // O / l in four 16-column chunks: fp16 hi staged and stored; then (split mode) the lo residuals through
        // the same 4 KB
        for (int part = 0; part < (p.out_lo ? 2 : 1); ++part) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(o_addr + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              uint32_t w[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float v0 = __uint_as_float(r[hq * 8 + 2 * q]) * inv_l, v1 = __uint_as_float(r[hq * 8 + 2 * q + 1]) * inv_l;
                uint32_t hh, ll;
                split_pair(v0, v1, hh, ll);
                w[q] = part ? ll : hh;
              }
              st_shared_v4(my_row + (((c * 2 + hq) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(part ? &o_lo : &o_hi, st_buf, h * ATT_D, warp_row0, b);
            tma_store_commit();
            tma_store_wait_read();     // the lo pass / the next item's P writes reuse this smem
          }
          __syncwarp();
        }
      }
      tc_fence_before_sync();
      if (tracer) trace(1 + x, 8, x, U);
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
