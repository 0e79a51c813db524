// Non-causal multi-head attention with a key-padding mask, head_dim 64, for sm_100a (tcgen05 + TMA).
// Reference arithmetic: transformers HubertAttention.forward (modeling_hubert.py:296-345) calling
// softmax(Q K^T / sqrt(64) + key_mask) V; reached from sylber/model/sylber.py:122.
//
// One CTA owns a (batch, head, 128-query tile) and streams 128-key blocks:
//   warp 0     : TMA producer - Q once, then K_j / V_j tiles through two 2-stage rings
//   warp 1     : MMA issuer   - S_j = Q K_j^T (128x128, fp32 in TMEM, double buffered) and
//                               PV_j = P_j V_j (128x64, fp32 in TMEM, double buffered)
//   warp 2     : TMEM allocator
//   warps 4..7 : softmax      - one query row per thread: tcgen05.ld S_j, mask, online max / sum in fp32,
//                               P_j -> fp16 into 128B-swizzled smem (the A operand of the PV MMA), then fold
//                               PV_{j-1} into the fp32 output registers with the running rescale
// Q arrives pre-scaled by 1/sqrt(64) (exact power of two, folded into the QKV GEMM epilogue).
// The score matrix never leaves the SM: HBM traffic is Q,K,V in and O out, 4*T*768*2 bytes per utterance
// per layer, K/V re-reads by the other query tiles of the same head are L2 hits.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_THREADS = 256;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: any [128 x 64] fp16 tile
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_K = ATT_SMEM_Q + ATT_TILE_BYTES;
constexpr int ATT_SMEM_V = ATT_SMEM_K + ATT_KV_STAGES * ATT_TILE_BYTES;
constexpr int ATT_SMEM_P = ATT_SMEM_V + ATT_KV_STAGES * ATT_TILE_BYTES;   // 2 buffers x 2 tiles
constexpr int ATT_SMEM_BAR = ATT_SMEM_P + 4 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_TOTAL = ATT_SMEM_BAR + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;      // 2 x 128 columns
constexpr uint32_t ATT_TMEM_O = 256;    // 2 x 64 columns

struct AttnParams {
  int T;                  // frames per utterance (rows per batch item in qkv)
  int heads;
  int model_dim;          // heads * 64
  const int* kv_len;      // [B] number of valid keys per utterance (== T when nothing is padded)
  __half* out_hi;         // [B*T, model_dim]
  __half* out_lo;         // optional
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap qkv_map, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2]
  uint64_t* p_full = bars + 11;       // [2]
  uint64_t* o_full = bars + 13;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int kv_len = p.kv_len ? min(p.kv_len[batch], p.T) : p.T;
  const int n_blocks = (kv_len + ATT_BKV - 1) / ATT_BKV;

  if (warp == 0 && elect_one()) tma_prefetch_desc(&qkv_map);
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
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

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(smem + ATT_SMEM_Q, &qkv_map, q_full, head * ATT_D, q0, batch);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_3d(smem + ATT_SMEM_K + st * ATT_TILE_BYTES, &qkv_map, &k_full[st], p.model_dim + head * ATT_D,
                    j * ATT_BKV, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_3d(smem + ATT_SMEM_V + st * ATT_TILE_BYTES, &qkv_map, &v_full[st], 2 * p.model_dim + head * ATT_D,
                    j * ATT_BKV, batch);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, ATT_BKV, 0, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_f16(128, ATT_D, 0, 0, 1);     // P (K-major) x V (MN-major)
      const uint64_t qdesc = make_desc_k_sw128(smem_u32(smem + ATT_SMEM_Q));
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t kdesc = make_desc_k_sw128(smem_u32(smem + ATT_SMEM_K + st * ATT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16_ss(tmem_base + ATT_TMEM_S + st * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        if (j + 1 < n_blocks) issue_s(j + 1);
        mbar_wait(&p_full[st], ph);   // P_j is in smem (and S_j has been consumed)
        mbar_wait(&v_full[st], ph);
        tc_fence_after_sync();
        const uint32_t pbase = smem_u32(smem + ATT_SMEM_P + st * 2 * ATT_TILE_BYTES);
        const uint32_t vbase = smem_u32(smem + ATT_SMEM_V + st * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          const uint64_t pdesc = make_desc_k_sw128(pbase + (kk >> 2) * ATT_TILE_BYTES) + 2 * (kk & 3);
          const uint64_t vdesc = make_desc_mn_sw128(vbase + kk * 16 * 128, ATT_TILE_BYTES);
          umma_f16_ss(tmem_base + ATT_TMEM_O + st * 64, pdesc, vdesc, idesc_o, kk != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&o_full[st]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax + output
    const int quarter = warp & 3;
    const int row = quarter * 32 + (int)lane_id();
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    float o[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, alpha_prev = 0.0f;

    for (int j = 0; j < n_blocks; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int n_valid = kv_len - j * ATT_BKV;   // keys of this block below the mask (may exceed 128)
      mbar_wait(&s_full[st], ph);
      tc_fence_after_sync();
      const uint32_t s_addr = tmem_base + lane_addr + ATT_TMEM_S + st * 128;
      // pass A: row maximum of the block
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (c * 32 + i < n_valid) ? __uint_as_float(r[i]) : -INFINITY;
          m_blk = fmaxf(m_blk, s);
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = exp2f((m_run - m_new) * kLog2e);   // 0 on the first block
      const float m_scaled = m_new * kLog2e;
      // pass B: probabilities -> fp16 -> swizzled smem, row sum in fp32
      float l_blk = 0.0f;
      uint8_t* pbuf = smem + ATT_SMEM_P + st * 2 * ATT_TILE_BYTES;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, r);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k0 = c * 32 + 2 * i;
          const float p0 = (k0 < n_valid) ? exp2f(fmaf(__uint_as_float(r[2 * i]), kLog2e, -m_scaled)) : 0.0f;
          const float p1 = (k0 + 1 < n_valid) ? exp2f(fmaf(__uint_as_float(r[2 * i + 1]), kLog2e, -m_scaled)) : 0.0f;
          l_blk += p0 + p1;
          packed[i] = pack_h2(__float2half_rn(p0), __float2half_rn(p1));
        }
        // keys [c*32, c*32+32) live in K-major tile (c>>1), 16-byte chunks (c&1)*4 .. +3 of this row
        uint8_t* trow = pbuf + (c >> 1) * ATT_TILE_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(trow + chunk * 16) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      fence_proxy_async_smem();      // generic-proxy P writes -> visible to the tensor core's async proxy
      tc_fence_before_sync();        // S_j loads are complete before the MMA warp may overwrite the buffer
      mbar_arrive(&p_full[st]);

      if (j > 0) {                   // fold PV_{j-1}:  O = O * alpha_{j-1} + PV_{j-1}
        const int so = (j - 1) & 1;
        mbar_wait(&o_full[so], ((j - 1) >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + ATT_TMEM_O + so * 64 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(r[i]));
        }
      }
      alpha_prev = alpha;
    }
    {
      const int so = (n_blocks - 1) & 1;
      mbar_wait(&o_full[so], ((n_blocks - 1) >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + ATT_TMEM_O + so * 64 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(r[i]));
      }
    }
    const float inv_l = 1.0f / l_run;
    if (q0 + row < p.T) {
      const size_t off = ((size_t)batch * p.T + q0 + row) * p.model_dim + head * ATT_D;
#pragma unroll
      for (int i = 0; i < ATT_D / 8; ++i) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half h0, l0, h1, l1;
          split_f16(o[8 * i + 2 * q] * inv_l, h0, l0);
          split_f16(o[8 * i + 2 * q + 1] * inv_l, h1, l1);
          hi[q] = pack_h2(h0, h1);
          lo[q] = pack_h2(l0, l1);
        }
        *reinterpret_cast<uint4*>(p.out_hi + off + 8 * i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + off + 8 * i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace syl
