// Positional convolution of the HuBERT encoder on tcgen05 (sm_100a):
//     pos[b, t, g*48+co] = GELU( bias + sum_{j<128} sum_{ci<48} w[g*48+co, ci, j] * h[b, t + j - 64, g*48+ci] )
// Reference arithmetic: transformers HubertPositionalConvEmbedding.forward (modeling_hubert.py:83-92): a
// weight-normed Conv1d(768, 768, k=128, padding=64, groups=16), the last output frame dropped (:98-103), erf-GELU.
//
// A grouped conv with 48-channel groups has only N=48 per GEMM, so streaming a fresh [128 x 64] activation tile per
// tap (what a generic implicit GEMM does) is bound by L2->SMEM bandwidth, not by the tensor cores.  This kernel
// keeps the activation window resident instead: one CTA owns 256 consecutive frames of one (utterance, group) and
// loads the 384-frame window [t0-64, t0+320) x 48 channels ONCE (TMA zero-fills frames outside [0,T), which is the
// conv's zero padding).  Tap j of M-tile m is then just the UMMA shared-memory descriptor whose start address is
// advanced by (j + 128 m) rows of 128 bytes; only the 6 KB weight tile of each tap streams through an mbarrier ring.
//
//   warp 0 : TMA producer (window + weight ring)      warp 1 : MMA issuer (128 taps x 2 M-tiles x passes x 3 k-steps)
//   warp 2 : TMEM allocator                            warps 4..11 : epilogue (bias + GELU, TMA store of fp32)
//
// Tap pairs (kPair, the single-pass mode).  The issuing thread, not the tensor core, bounds this kernel: an N = 48 MMA
// is 24 cycles of math but ~90 cycles of issue + dispatch here (ncu: tensor pipe 31 % active).  So two taps share one
// MMA: the B operand is [W_j ; W_{j+1}] (N = 96, adjacent halves of a ring stage) against the window shifted by the
// EVEN tap j.  Columns 0..47 then hold D[r] = sum_{j even} W_j x[r + j], columns 48..95 hold
// E[r] = sum_{j even} W_{j+1} x[r + j], and since the odd taps need x[r + j + 1],  out[r] = D[r] + E[r + 1].
// The epilogue fetches E[r + 1] from the neighbouring row through shared memory; a CTA tile therefore yields 255
// output frames (row 255 has no E[256]) and tiles advance by 255 frames.  Half the MMAs, half the window re-reads.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int PC_TAPS = 128;
constexpr int PC_CG = 48;                  // channels per group
constexpr int PC_GROUPS = 16;
constexpr int PC_TILE_T = 256;             // frames per CTA tile (2 MMA M-tiles)
constexpr int PC_TILE_T_PAIR = 255;        // output frames per CTA tile in tap-pair mode
constexpr int PC_WIN_ROWS = 384;           // window rows: 64 left halo + 256 + 63 right halo, rounded up
constexpr int PC_WIN_BYTES = PC_WIN_ROWS * 128;      // 48 KB (64 fp16 columns per row, 48 used)
constexpr int PC_W_BYTES = PC_CG * 128;              // 6 KB per tap
constexpr int PC_W_STAGES = 8;
constexpr int PC_THREADS = 384;
constexpr int PC_SMEM_WIN_HI = 0;
constexpr int PC_SMEM_WIN_LO = PC_WIN_BYTES;
constexpr int PC_SMEM_W = 2 * PC_WIN_BYTES;                          // ring: [stage][hi|lo]
constexpr int PC_SMEM_EPI = PC_SMEM_W + PC_W_STAGES * 2 * PC_W_BYTES; // 8 warps x 2 KB
constexpr int PC_SMEM_BAR = PC_SMEM_EPI + 8 * 2048;
constexpr int PC_SMEM_TOTAL = PC_SMEM_BAR + 256 + 1024;
constexpr uint32_t PC_TMEM_COLS = 512;     // 2 accumulator stages x 2 M-tiles x 128 columns (96 used when split)

struct PosConvParams {
  int T;
  int batches;
  int n_pass;            // 1 or 3
  const float* bias;     // [768]
  int use_base_offset;   // descriptor base-offset field for row-shifted starts (see make_desc below)
};

// K-major SWIZZLE_128B descriptor whose start address is shifted by whole 128-byte rows, i.e. NOT aligned to the
// 1024-byte swizzle atom.  Measured on B200: the tensor core applies the swizzle XOR to absolute shared-memory
// address bits (as TMA does when it writes the window), so the plain descriptor is already correct and the
// matrix-base-offset field (bits 49..51) must stay 0; use_base_offset = 1 is kept only to reproduce that experiment.
__device__ __forceinline__ uint64_t make_desc_k_sw128_shifted(uint32_t smem_addr, int use_base_offset) {
  uint64_t d = make_desc_k_sw128(smem_addr);
  if (use_base_offset) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
  return d;
}

template <bool kPair>
__global__ void __launch_bounds__(PC_THREADS, 1)
posconv_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap w_hi, const __grid_constant__ CUtensorMap w_lo,
               const __grid_constant__ CUtensorMap o_map, const __grid_constant__ CUtensorMap o_map31,
               const PosConvParams p) {
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PC_SMEM_BAR);
  uint64_t* win_full = bars;                 // [1]
  uint64_t* win_empty = bars + 1;            // [1]
  uint64_t* w_full = bars + 2;               // [PC_W_STAGES]
  uint64_t* w_empty = bars + 2 + PC_W_STAGES;
  uint64_t* tmem_full = bars + 2 + 2 * PC_W_STAGES;    // [2]
  uint64_t* tmem_empty = bars + 4 + 2 * PC_W_STAGES;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6 + 2 * PC_W_STAGES);

  const int warp = threadIdx.x >> 5;
  constexpr int kTileT = kPair ? PC_TILE_T_PAIR : PC_TILE_T;
  const int t_tiles = (p.T + kTileT - 1) / kTileT;
  const int num_tiles = p.batches * t_tiles * PC_GROUPS;
  const bool split = !kPair && p.n_pass == 3;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&a_hi);
    tma_prefetch_desc(&w_hi);
    tma_prefetch_desc(&o_map);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(win_full, 1);
    mbar_init(win_empty, 1);
    for (int i = 0; i < PC_W_STAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<PC_TMEM_COLS>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, win_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = tile % PC_GROUPS;
        const int tt = (tile / PC_GROUPS) % t_tiles;
        const int b = tile / (PC_GROUPS * t_tiles);
        const int row0 = tt * kTileT - PC_TAPS / 2;
        mbar_wait(win_empty, win_phase ^ 1);
        mbar_arrive_expect_tx(win_full, split ? 2 * PC_WIN_BYTES : PC_WIN_BYTES);
        tma_load_3d(smem + PC_SMEM_WIN_HI, &a_hi, win_full, g * PC_CG, row0, b);
        tma_load_3d(smem + PC_SMEM_WIN_HI + PC_WIN_BYTES / 2, &a_hi, win_full, g * PC_CG, row0 + PC_WIN_ROWS / 2, b);
        if (split) {
          tma_load_3d(smem + PC_SMEM_WIN_LO, &a_lo, win_full, g * PC_CG, row0, b);
          tma_load_3d(smem + PC_SMEM_WIN_LO + PC_WIN_BYTES / 2, &a_lo, win_full, g * PC_CG, row0 + PC_WIN_ROWS / 2, b);
        }
        win_phase ^= 1;
        for (int tap = 0; tap < PC_TAPS; tap += (kPair ? 2 : 1)) {
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], (split || kPair) ? 2 * PC_W_BYTES : PC_W_BYTES);
          uint8_t* ws = smem + PC_SMEM_W + stage * 2 * PC_W_BYTES;
          tma_load_2d(ws, &w_hi, &w_full[stage], tap * 64, g * PC_CG);
          if (split) tma_load_2d(ws + PC_W_BYTES, &w_lo, &w_full[stage], tap * 64, g * PC_CG);
          if (kPair) tma_load_2d(ws + PC_W_BYTES, &w_hi, &w_full[stage], (tap + 1) * 64, g * PC_CG);   // [W_j ; W_{j+1}]
          if (++stage == PC_W_STAGES) {
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
      // Split precision needs A_hi*W_hi + A_lo*W_hi + A_hi*W_lo.  The single issuing thread, not the tensor core, is
      // the limit with MMAs this small (N=48: 24 cycles each), so W_hi and W_lo - adjacent in the ring stage - are
      // consumed as ONE N=96 operand by the A_hi pass: columns 0..47 accumulate A_hi*W_hi (+ A_lo*W_hi from the
      // second MMA), columns 48..95 accumulate A_hi*W_lo, and the epilogue adds the two halves.  12 MMAs per tap
      // instead of 18.
      constexpr uint32_t idesc = make_idesc_f16(128, PC_CG, 0, 0, 0);
      constexpr uint32_t idesc96 = make_idesc_f16(128, 2 * PC_CG, 0, 0, 0);
      const uint32_t win_hi = smem_u32(smem + PC_SMEM_WIN_HI);
      const uint32_t win_lo = smem_u32(smem + PC_SMEM_WIN_LO);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, win_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(win_full, win_phase);
        win_phase ^= 1;
        tc_fence_after_sync();
        for (int tap = 0; tap < PC_TAPS; tap += (kPair ? 2 : 1)) {
          mbar_wait(&w_full[stage], phase);
          tc_fence_after_sync();
          const uint32_t wb = smem_u32(smem + PC_SMEM_W + stage * 2 * PC_W_BYTES);
          const uint64_t bd_hi = make_desc_k_sw128(wb);      // [W_hi ; W_lo] (split) or [W_j ; W_{j+1}] (pair): 96 rows, else 48
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const uint32_t tmem_d = tmem_base + acc * 256 + m * 128;
            const uint32_t roff = (uint32_t)(tap + 128 * m) * 128;
            const uint64_t ad_hi = make_desc_k_sw128_shifted(win_hi + roff, p.use_base_offset);
            if (split) {
              const uint64_t ad_lo = make_desc_k_sw128_shifted(win_lo + roff, p.use_base_offset);
#pragma unroll
              for (int k = 0; k < 3; ++k) umma_f16_ss(tmem_d, ad_hi + 2 * k, bd_hi + 2 * k, idesc96, (tap | k) != 0);
#pragma unroll
              for (int k = 0; k < 3; ++k) umma_f16_ss(tmem_d, ad_lo + 2 * k, bd_hi + 2 * k, idesc, 1);
            } else if (kPair) {
#pragma unroll
              for (int k = 0; k < 3; ++k) umma_f16_ss(tmem_d, ad_hi + 2 * k, bd_hi + 2 * k, idesc96, (tap | k) != 0);
            } else {
#pragma unroll
              for (int k = 0; k < 3; ++k) umma_f16_ss(tmem_d, ad_hi + 2 * k, bd_hi + 2 * k, idesc, (tap | k) != 0);
            }
          }
          umma_commit(&w_empty[stage]);
          if (++stage == PC_W_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(win_empty);          // the window may be overwritten once every MMA of this tile has read it
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int m = ew >> 2;
    const int lane = (int)lane_id();
    uint8_t* stage_buf = smem + PC_SMEM_EPI + ew * 2048;
    uint8_t* row64 = stage_buf + lane * 64;
    const int sw64 = (lane >> 1) & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int g = tile % PC_GROUPS;
      const int tt = (tile / PC_GROUPS) % t_tiles;
      const int b = tile / (PC_GROUPS * t_tiles);
      const int warp_row0 = tt * kTileT + m * 128 + quarter * 32;
      const bool warp_ok = warp_row0 < p.T;
      // tap-pair mode: rows exchange their E halves through the (unused) lo-window region, one 16 KB slab per column
      // chunk; the barrier keeps this tile's writes behind the previous tile's reads
      const int R = m * 128 + quarter * 32 + lane;                 // row inside the CTA tile
      const uint32_t xbuf = smem_u32(smem + PC_SMEM_WIN_LO);
      if (kPair) named_bar_sync(1, 256);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + m * 128);
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        uint32_t r[16], r2[16];
        tmem_ld_32x32b_x16(taddr + c * 16, r);
        if (split || kPair) tmem_ld_32x32b_x16(taddr + PC_CG + c * 16, r2);     // the A_hi * W_lo half / the odd-tap half E
        tmem_ld_wait();
        if (kPair) {
          // E[R] -> shared memory (64-byte rows, 16-byte chunks XOR-swizzled by (row >> 1) & 3), then read E[R + 1]
          const uint32_t slab = xbuf + c * 16384;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_shared_v4(slab + R * 64 + ((i ^ ((R >> 1) & 3)) << 4), r2[4 * i], r2[4 * i + 1], r2[4 * i + 2], r2[4 * i + 3]);
          named_bar_sync(2, 256);
          const int Rn = min(R + 1, 255);                          // row 255 is never stored
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t a = slab + Rn * 64 + ((i ^ ((Rn >> 1) & 3)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r2[4 * i]), "=r"(r2[4 * i + 1]), "=r"(r2[4 * i + 2]), "=r"(r2[4 * i + 3]) : "r"(a));
          }
        }
        const int col0 = g * PC_CG + c * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float acc_v = __uint_as_float(r[i]);
          if (split || kPair) acc_v += __uint_as_float(r2[i]);
          v[i] = acc_v + __ldg(p.bias + col0 + i);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) gelu_fast2(v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1]);
        if (warp_ok) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(row64 + ((i ^ sw64) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // tap-pair mode: the tile's last warp stores 31 rows (row 255 belongs to the next tile)
            tma_store_3d((kPair && m == 1 && quarter == 3) ? &o_map31 : &o_map, stage_buf, col0, warp_row0, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<PC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace syl
