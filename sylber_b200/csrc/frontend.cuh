// Waveform front end: conv0 (1->512, k=10, s=5, no bias) + GroupNorm(512 groups) + erf-GELU, and the
// row-wise LayerNorm kernels.  Reference arithmetic: transformers HubertGroupNormConvLayer
// (modeling_hubert.py:154-175), HubertFeatureProjection (:225-231), HubertEncoder (:441-443),
// HubertEncoderLayer (:394-398); reached from sylber/model/sylber.py:122.
//
// These stages are HBM bound.  The GroupNorm statistics never touch the 2 GB conv0 activation: because
// conv0 is linear in the 10 taps, the per-channel sum and sum of squares over time follow from the
// 10x10 second-moment matrix of the strided waveform windows,
//     sum_t y_c[t]   = w_c . S          S[j]    = sum_t x[5t+j]
//     sum_t y_c[t]^2 = w_c^T R w_c      R[j,j'] = sum_t x[5t+j] x[5t+j']
// accumulated in fp64 (exact products of fp32 inputs), so the statistics pass reads only the waveform.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int C0_K = 10;
constexpr int C0_S = 5;
constexpr int C0_OUT = 512;
constexpr int C0_NMOM = C0_K + C0_K * (C0_K + 1) / 2;  // 10 sums + 55 upper-triangle second moments

// ----------------------------------------------------------------------------------------------
// conv0 moments: grid (chunks, B), block 128.  Each block writes its partial sums to part[b][chunk][65]; the
// coefficient kernel adds the chunks in index order, so the statistics are bit-reproducible run to run.
// ----------------------------------------------------------------------------------------------
constexpr int MOM_THREADS = 128;
constexpr int MOM_T_PER_BLOCK = 1024;

__global__ void __launch_bounds__(MOM_THREADS)
conv0_moments_kernel(const float* __restrict__ wav, int t_samp, int L0, double* __restrict__ part) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float xs[MOM_T_PER_BLOCK * C0_S + C0_K];
  __shared__ double red[MOM_THREADS / 32][C0_NMOM];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MOM_T_PER_BLOCK;
  const int nt = min(MOM_T_PER_BLOCK, L0 - t0);
  const float* w = wav + (size_t)b * t_samp + (size_t)t0 * C0_S;
  const int nsamp = nt * C0_S + (C0_K - C0_S);
  for (int i = threadIdx.x; i < nsamp; i += MOM_THREADS) xs[i] = w[i];
  __syncthreads();

  double acc[C0_NMOM];
#pragma unroll
  for (int i = 0; i < C0_NMOM; ++i) acc[i] = 0.0;
  for (int t = threadIdx.x; t < nt; t += MOM_THREADS) {
    double x[C0_K];
#pragma unroll
    for (int j = 0; j < C0_K; ++j) x[j] = (double)xs[t * C0_S + j];
    int idx = C0_K;
#pragma unroll
    for (int j = 0; j < C0_K; ++j) {
      acc[j] += x[j];
#pragma unroll
      for (int k = j; k < C0_K; ++k) acc[idx++] += x[j] * x[k];
    }
  }
#pragma unroll
  for (int i = 0; i < C0_NMOM; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < C0_NMOM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < MOM_THREADS / 32; ++w) v += red[w][threadIdx.x];
    part[((size_t)b * gridDim.x + blockIdx.x) * C0_NMOM + threadIdx.x] = v;
  }
}

// per (b, c): GroupNorm scale/shift so that  gn(y) = y * scale + shift   (eps 1e-5, biased variance)
__global__ void conv0_gn_coeff_kernel(const double* __restrict__ part, int chunks, const float* __restrict__ w0,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, int L0,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ double m[C0_NMOM];
  const int b = blockIdx.x;
  const int c = threadIdx.x;
  if (c < C0_NMOM) {
    double v = 0.0;
    for (int k = 0; k < chunks; ++k) v += part[((size_t)b * chunks + k) * C0_NMOM + c];
    m[c] = v;
  }
  __syncthreads();
  double w[C0_K];
#pragma unroll
  for (int j = 0; j < C0_K; ++j) w[j] = (double)w0[c * C0_K + j];
  double s1 = 0.0, s2 = 0.0;
  int idx = C0_K;
#pragma unroll
  for (int j = 0; j < C0_K; ++j) {
    s1 += w[j] * m[j];
#pragma unroll
    for (int k = j; k < C0_K; ++k) {
      const double r = m[idx++];
      s2 += (j == k ? 1.0 : 2.0) * w[j] * w[k] * r;
    }
  }
  const double mean = s1 / L0;
  double var = s2 / L0 - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + 1e-5);
  const double sc = (double)gamma[c] * rstd;
  scale[b * C0_OUT + c] = (float)sc;
  shift[b * C0_OUT + c] = (float)((double)beta[c] - mean * sc);
}

// ----------------------------------------------------------------------------------------------
// conv0 + GroupNorm + GELU with the 10-tap dot products on the tensor cores (warp-level mma.sync m16n8k16).
// The round-1 kernel (10 FFMA + ~13 other instructions per output element, one thread per 4 channels) was bound by
// instruction issue: 0.66 ms against 0.43 ms here (profiles/r02_bench_ab.md).  Output channels-last fp16 hi (+ lo).
// Here a warp computes [16 frames] x [8 channels] x [K = 16: taps 0-9, zero padded] per MMA; the fp32 waveform and
// weights enter as fp16 hi + lo and three MMAs (hi*hi, lo*hi, hi*lo) accumulate in fp32, which keeps 22 significant
// bits of both operands (the GEMM kernels' split scheme).  What remains per element is the GroupNorm FFMA2, the GELU
// and the fp16 conversion.  tcgen05 is not used here on purpose: K = 10 and one TMEM round trip per 512-channel row
// would cost more than the MMA saves, while mma.sync keeps the accumulator in the registers the epilogue needs.
//
// grid (ceil(L0 / 256), B), block 256: warp w owns channels [128 (w & 3), + 128) of half of the block's frames; the B
// fragments of the 64 channel octets come from a 32 KB table built at syl_finalize and read through L1.  Inside a group
// of four octets the MMA columns are permuted so that a lane ends up with 8 CONSECUTIVE channels per frame row, i.e.
// one 16-byte store:
//     MMA column n of octet jj in group grp  <->  channel 128 w + 32 grp + 8 (n >> 1) + 2 jj + (n & 1)
// ----------------------------------------------------------------------------------------------
constexpr int C0M_THREADS = 256;
constexpr int C0M_T = 256;       // frames per block = 16 frame tiles of 16: warps 0-3 take tiles 0-7, warps 4-7 tiles 8-15

__device__ __forceinline__ void mma_m16n8k16_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// B fragments ("col" operand) of all 64 channel octets, built once at syl_finalize: tab[octet][lane] =
// {hi(k = 2 tig, 2 tig + 1), hi(k = 2 tig + 8, + 9), lo(...), lo(...)} of MMA column n = lane >> 2, i.e. of channel
// 128 w + 32 grp + 8 (n >> 1) + 2 jj + (n & 1) with octet = 16 w + 4 grp + jj; taps >= 10 are zero.  32 KB, read
// through L1 by every block - keeping the fragments in registers cost 64 registers per thread and held the kernel
// at 4 warps per scheduler.
__global__ void conv0_bfrag_kernel(const float* __restrict__ w0, uint4* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 32) return;
  const int o = i >> 5, lane = i & 31, g = lane >> 2, tig = lane & 3;
  const int ch = (o >> 4) * 128 + ((o >> 2) & 3) * 32 + (g >> 1) * 8 + (o & 3) * 2 + (g & 1);
  const float* wc = w0 + ch * C0_K;
  const float w10 = (tig == 0) ? wc[8] : 0.0f, w11 = (tig == 0) ? wc[9] : 0.0f;
  uint4 v;
  split_pair(wc[2 * tig], wc[2 * tig + 1], v.x, v.z);
  split_pair(w10, w11, v.y, v.w);
  tab[i] = v;
}

// 4 blocks per SM (64 registers, 32 warps per SM).  A 5-block build (48 registers, 28 bytes of spills, 40 warps) was
// measured slower, 0.429 vs 0.410 ms (profiles/r03_variants_ab.md), and removed.
template <bool kLo>
__global__ void __launch_bounds__(C0M_THREADS, 4)
conv0_mma_kernel(const float* __restrict__ wav, int t_samp, int L0, const uint4* __restrict__ bfrag,
                 const float* __restrict__ scale, const float* __restrict__ shift, __half* __restrict__ out_hi,
                 __half* __restrict__ out_lo, const int32_t* __restrict__ needed_rows) {
  __shared__ float xs[C0M_T * C0_S + 16];
  __shared__ __align__(16) float s_sc[C0_OUT];
  __shared__ __align__(16) float s_sh[C0_OUT];
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * C0M_T;
  if (needed_rows != nullptr && t0 >= __ldg(needed_rows + b)) return;   // trimmed mode: no valid frame reads these rows
  const int nt = min(C0M_T, L0 - t0);
  const float* w = wav + (size_t)b * t_samp + (size_t)t0 * C0_S;
  const int nsamp = nt * C0_S + (C0_K - C0_S);
  for (int i = threadIdx.x; i < C0M_T * C0_S + 16; i += C0M_THREADS) xs[i] = (i < nsamp) ? w[i] : 0.0f;
  for (int i = threadIdx.x; i < C0_OUT; i += C0M_THREADS) {
    s_sc[i] = scale[b * C0_OUT + i];
    s_sh[i] = shift[b * C0_OUT + i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = warp & 3, half = warp >> 2;
  const int g = lane >> 2, tig = lane & 3;
  const uint4* btab = bfrag + (size_t)slab * 16 * 32 + lane;
  __syncthreads();

  const int ch0 = slab * 128 + tig * 8;         // + 32 grp: the 8 consecutive channels of this lane in group grp
  for (int ft = half * 8; ft < half * 8 + 8; ++ft) {
    if (ft * 16 >= nt) break;
    // A fragments (waveform windows, row-major 16 x 16): rows g and g + 8, k as for B
    uint32_t ah[4], al[4];
    {
      const float* x0 = xs + (ft * 16 + g) * C0_S;
      const float* x1 = x0 + 8 * C0_S;
      split_pair(x0[2 * tig], x0[2 * tig + 1], ah[0], al[0]);
      split_pair(x1[2 * tig], x1[2 * tig + 1], ah[1], al[1]);
      const float e00 = (tig == 0) ? x0[8] : 0.0f, e01 = (tig == 0) ? x0[9] : 0.0f;
      const float e10 = (tig == 0) ? x1[8] : 0.0f, e11 = (tig == 0) ? x1[9] : 0.0f;
      split_pair(e00, e01, ah[2], al[2]);
      split_pair(e10, e11, ah[3], al[3]);
    }
    const int r0 = ft * 16 + g, r1 = r0 + 8;
    const size_t o0 = ((size_t)b * L0 + t0 + r0) * C0_OUT + ch0;
    const size_t o1 = o0 + (size_t)8 * C0_OUT;
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      const float4 sc0 = *reinterpret_cast<const float4*>(s_sc + ch0 + grp * 32);
      const float4 sc1 = *reinterpret_cast<const float4*>(s_sc + ch0 + grp * 32 + 4);
      const float4 sh0 = *reinterpret_cast<const float4*>(s_sh + ch0 + grp * 32);
      const float4 sh1 = *reinterpret_cast<const float4*>(s_sh + ch0 + grp * 32 + 4);
      const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      uint32_t h0[4], h1[4], l0[4], l1[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint4 bf = __ldg(btab + (grp * 4 + jj) * 32);     // {hi0, hi1, lo0, lo1}
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        mma_m16n8k16_f16(c, al, bf.x, bf.y);     // small terms first
        mma_m16n8k16_f16(c, ah, bf.z, bf.w);
        mma_m16n8k16_f16(c, ah, bf.x, bf.y);
        // c[0], c[1]: row g, channels ch0 + 32 grp + 2 jj + {0, 1};  c[2], c[3]: row g + 8, same channels
        const f32x2 sc2 = pack2(scv[2 * jj], scv[2 * jj + 1]), sh2 = pack2(shv[2 * jj], shv[2 * jj + 1]);
        float y0, y1, y2, y3;
        unpack2(fma2(pack2(c[0], c[1]), sc2, sh2), y0, y1);
        unpack2(fma2(pack2(c[2], c[3]), sc2, sh2), y2, y3);
        gelu_fast2(y0, y1, y0, y1);
        gelu_fast2(y2, y3, y2, y3);
        if (kLo) {
          split_pair(y0, y1, h0[jj], l0[jj]);
          split_pair(y2, y3, h1[jj], l1[jj]);
        } else {
          h0[jj] = pack_f16x2_sat(y0, y1);
          h1[jj] = pack_f16x2_sat(y2, y3);
        }
      }
      if (r0 < nt) {
        *reinterpret_cast<uint4*>(out_hi + o0 + grp * 32) = make_uint4(h0[0], h0[1], h0[2], h0[3]);
        if (kLo) *reinterpret_cast<uint4*>(out_lo + o0 + grp * 32) = make_uint4(l0[0], l0[1], l0[2], l0[3]);
      }
      if (r1 < nt) {
        *reinterpret_cast<uint4*>(out_hi + o1 + grp * 32) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
        if (kLo) *reinterpret_cast<uint4*>(out_lo + o1 + grp * 32) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Row LayerNorm (eps 1e-5), one warp per row, D in {512, 768}:
//    y = LN(x (+ add)) * gamma + beta  ->  fp32 and / or fp16 hi (+ lo)
// The residual may come as fp32 (`add`) or as the fp16 pair (add_hi, add_lo) this kernel wrote itself one layer
// earlier: inside the encoder the residual stream lives ONLY as hi + lo (22 significant bits, the same 4 bytes as
// fp32), so a LayerNorm call moves 12 bytes per element (fp32 GEMM output + pair in, pair out) instead of 14.
// (Two rows per warp - one resident wave of warps, twice the loads in flight - was measured slower: 0.78 vs 0.64 ms.)
// ----------------------------------------------------------------------------------------------
// __launch_bounds__(256, 4): the 64-register allocation this asks for schedules all row loads up front; the kernel is
// latency bound, and 0.65 -> 0.575 ms per step against the 59-register default (40 / 32 registers with 6 / 8 blocks
// per SM spill and are slower: 0.65 / 0.75 ms; profiles/r02_bench_ab.md)
template <int D>
__global__ void __launch_bounds__(256, 4)
layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ add, const __half* add_hi, const __half* add_lo,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int rows, float* __restrict__ out_f32,
                      __half* out_hi, __half* out_lo, const int32_t* __restrict__ valid, int T) {
  constexpr int V = D / 128;  // float4 per lane
  griddep_launch_dependents();
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  if (valid != nullptr && (row % T) >= __ldg(valid + row / T)) {
    // trimmed mode: a padded frame.  Nothing upstream computed it; its outputs are defined as zero.
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const size_t o = (size_t)row * D + (size_t)(lane + 32 * i) * 4;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (out_hi) *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(0u, 0u);
      if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(0u, 0u);
    }
    return;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xr[lane + 32 * i];
  if (add) {
    const float4* ar = reinterpret_cast<const float4*>(add + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 a = ar[lane + 32 * i];
      v[i].x += a.x;
      v[i].y += a.y;
      v[i].z += a.z;
      v[i].w += a.w;
    }
  } else if (add_hi) {
    const uint2* hr = reinterpret_cast<const uint2*>(add_hi + (size_t)row * D);
    const uint2* lr = reinterpret_cast<const uint2*>(add_lo + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const uint2 hh = hr[lane + 32 * i], ll = lr[lane + 32 * i];
      const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hh.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&hh.y));
      const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&ll.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&ll.y));
      v[i].x += h0.x + l0.x;
      v[i].y += h0.y + l0.y;
      v[i].z += h1.x + l1.x;
      v[i].w += h1.y + l1.y;
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / D) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + bb.x;
    y.y = (v[i].y - mean) * rstd * g.y + bb.y;
    y.z = (v[i].z - mean) * rstd * g.z + bb.z;
    y.w = (v[i].w - mean) * rstd * g.w + bb.w;
    const size_t o = (size_t)row * D + (size_t)(lane + 32 * i) * 4;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = y;
    if (out_hi) {
      if (out_lo) {
        uint32_t h0, l0, h1, l1;
        split_pair(y.x, y.y, h0, l0);
        split_pair(y.z, y.w, h1, l1);
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(l0, l1);
      } else {
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(pack_f16x2_sat(y.x, y.y), pack_f16x2_sat(y.z, y.w));
      }
    }
  }
}

}  // namespace syl
