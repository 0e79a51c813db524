// The steps either side of the Segmenter forward path, on the device:
//   * prepare_pcm16: int16 PCM -> fp32, (w - mean) / std with the unbiased std, zero padding to the batch maximum
//     (sylber/model/sylber.py:83-87 file branch, :93-118 padding) - replaces the host-side preprocessing and halves
//     the host->device bytes;
//   * kmeans_assign: nearest-centroid lookup of segment features (sylber/model/quantizer.py:86-135 KMQuantizer, whose
//     vector_quantize_pytorch codebook does an exhaustive Euclidean search; optional token normalisation of :99).
// Both are bandwidth / latency bound helpers, not tensor-core work.
#pragma once

#include "common.cuh"

namespace syl {

constexpr int PCM_THREADS = 256;
constexpr int PCM_CHUNK = 8192;          // samples per block

__device__ __forceinline__ float pcm_to_float(int16_t v) { return (float)v * (1.0f / 32768.0f); }
__device__ __forceinline__ float pcm_to_float(float v) { return v; }

// partial (sum, sum of squares) of the samples (int16 / 32768, or fp32 as is) per (utterance, chunk), fp64 so that the
// statistics do not depend on the summation order beyond rounding of the final result
template <typename TIn>
__global__ void __launch_bounds__(PCM_THREADS)
pcm16_stats_kernel(const TIn* __restrict__ pcm, const int64_t* __restrict__ offsets, const int32_t* __restrict__ n_samples,
                   int chunks, double* __restrict__ part) {
  __shared__ double red[2][PCM_THREADS / 32];
  const int b = blockIdx.y;
  const int n = n_samples[b];
  const int64_t base = offsets[b];
  const int i0 = blockIdx.x * PCM_CHUNK;
  double s = 0.0, q = 0.0;
  for (int i = i0 + threadIdx.x; i < min(i0 + PCM_CHUNK, n); i += PCM_THREADS) {
    const double x = (double)pcm_to_float(pcm[base + i]);
    s += x;
    q += x * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane_id() == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < PCM_THREADS / 32; ++w) {
      ts += red[0][w];
      tq += red[1][w];
    }
    part[((size_t)b * chunks + blockIdx.x) * 2 + 0] = ts;
    part[((size_t)b * chunks + blockIdx.x) * 2 + 1] = tq;
  }
}

// out[b, i] = (x_i - mean) / std for i < n (normalize) or x_i (plain), 0 for n <= i < t_max
template <typename TIn>
__global__ void __launch_bounds__(PCM_THREADS)
pcm16_apply_kernel(const TIn* __restrict__ pcm, const int64_t* __restrict__ offsets, const int32_t* __restrict__ n_samples,
                   int chunks, const double* __restrict__ part, int normalize, int t_max, float* __restrict__ out) {
  __shared__ float s_mean, s_rstd;
  const int b = blockIdx.y;
  const int n = n_samples[b];
  const int64_t base = offsets[b];
  if (threadIdx.x == 0) {
    float mean = 0.0f, rstd = 1.0f;
    if (normalize) {
      double ts = 0.0, tq = 0.0;
      const int used = (n + PCM_CHUNK - 1) / PCM_CHUNK;
      for (int k = 0; k < used; ++k) {
        ts += part[((size_t)b * chunks + k) * 2 + 0];
        tq += part[((size_t)b * chunks + k) * 2 + 1];
      }
      const double m = ts / (double)n;
      const double var = (tq - (double)n * m * m) / (double)(n - 1);   // torch.std: unbiased
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var));
    }
    s_mean = mean;
    s_rstd = rstd;
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  const int i0 = blockIdx.x * PCM_CHUNK;
  for (int i = i0 + threadIdx.x; i < min(i0 + PCM_CHUNK, t_max); i += PCM_THREADS) {
    float v = 0.0f;
    if (i < n) v = (pcm_to_float(pcm[base + i]) - mean) * rstd;
    out[(size_t)b * t_max + i] = v;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Band-limited sinc resampling (torchaudio.transforms.Resample, sylber.py:85): one windowed-sinc FIR per output phase,
//   out[b, i * new_g + p] = sum_k h[p, k] * xpad[i * orig_g + k],   xpad = x with `width` zeros in front and zeros behind,
// h [new_g, K = 2 width + orig_g] from sylber_b200/resample.py.  n_out[b] = ceil(new_g * n_in[b] / orig_g); the rest of
// the row is zero filled.  grid (ceil(t_out_max / 256), B), block 256: one output sample per thread.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, const int32_t* __restrict__ n_in, int t_in_max, const float* __restrict__ h,
                int orig_g, int new_g, int width, float* __restrict__ out, int32_t* __restrict__ n_out, int t_out_max) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_in[b];
  const int nout = (int)(((long long)new_g * n + orig_g - 1) / orig_g);
  if (j == 0 && n_out) n_out[b] = nout;
  if (j >= t_out_max) return;
  float acc = 0.0f;
  if (j < nout) {
    const int i = j / new_g, p = j - i * new_g;
    const int K = 2 * width + orig_g;
    const float* hp = h + (size_t)p * K;
    const float* x = in + (size_t)b * t_in_max;
    const int m0 = i * orig_g - width;                  // sample index of tap 0
    const int k_lo = max(0, -m0), k_hi = min(K, n - m0);
    for (int k = k_lo; k < k_hi; ++k) acc = fmaf(__ldg(hp + k), x[m0 + k], acc);
  }
  out[(size_t)b * t_out_max + j] = acc;
}

// ----------------------------------------------------------------------------------------------------------------
// k-means assignment: idx[r] = argmin_k sum_d (x[r, d] - c[k, d])^2, D = 768, first minimum wins.
// grid ceil(n / 8): a block keeps 8 feature rows in shared memory; its 8 warps stride over the centroids, each lane
// holding 24 of the 768 coordinates; the 8 per-row partial distances are reduced with xor shuffles.
// ----------------------------------------------------------------------------------------------------------------
constexpr int KM_ROWS = 8;
constexpr int KM_WARPS = 8;
constexpr int KM_D = 768;

__global__ void __launch_bounds__(KM_WARPS * 32)
kmeans_assign_kernel(const float* __restrict__ feats, int n, const float* __restrict__ cent, int K, int normalize,
                     int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ float xs[KM_ROWS][KM_D];
  __shared__ float best_d[KM_WARPS][KM_ROWS];
  __shared__ int best_k[KM_WARPS][KM_ROWS];
  const int r0 = blockIdx.x * KM_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // load (and optionally normalise: x / sqrt(sum x^2 + 1e-8) * 6, quantizer.py:99) one row per warp
  {
    const int r = r0 + warp;
    float v[KM_D / 32];
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < KM_D / 32; ++i) {
      v[i] = (r < n) ? feats[(size_t)r * KM_D + lane + 32 * i] : 0.0f;
      ss += v[i] * v[i];
    }
    float sc = 1.0f;
    if (normalize) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sc = 6.0f / sqrtf(ss + 1e-8f);
    }
#pragma unroll
    for (int i = 0; i < KM_D / 32; ++i) xs[warp][lane + 32 * i] = v[i] * sc;
  }
  __syncthreads();
  float bd[KM_ROWS];
  int bk[KM_ROWS];
#pragma unroll
  for (int r = 0; r < KM_ROWS; ++r) {
    bd[r] = INFINITY;
    bk[r] = 0;
  }
  for (int k = warp; k < K; k += KM_WARPS) {
    float c[KM_D / 32];
#pragma unroll
    for (int i = 0; i < KM_D / 32; ++i) c[i] = __ldg(cent + (size_t)k * KM_D + lane + 32 * i);
#pragma unroll
    for (int r = 0; r < KM_ROWS; ++r) {
      float d = 0.0f;
#pragma unroll
      for (int i = 0; i < KM_D / 32; ++i) {
        const float e = xs[r][lane + 32 * i] - c[i];
        d = fmaf(e, e, d);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (d < bd[r]) {          // k ascends within a warp: strict < keeps the first minimum
        bd[r] = d;
        bk[r] = k;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < KM_ROWS; ++r) {
      best_d[warp][r] = bd[r];
      best_k[warp][r] = bk[r];
    }
  }
  __syncthreads();
  if (threadIdx.x < KM_ROWS && r0 + threadIdx.x < n) {
    const int r = threadIdx.x;
    float d = best_d[0][r];
    int k = best_k[0][r];
    for (int w = 1; w < KM_WARPS; ++w) {
      const float dw = best_d[w][r];
      const int kw = best_k[w][r];
      if (dw < d || (dw == d && kw < k)) {
        d = dw;
        k = kw;
      }
    }
    idx_out[r0 + r] = k;
    if (dist_out) dist_out[r0 + r] = d;
  }
}

}  // namespace syl
