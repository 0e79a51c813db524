// On-device syllable segmentation and segment-mean pooling.
// Reference: sylber/utils/segment_utils.py:68-131 (cossim, get_segment) and sylber/model/sylber.py:126-133.
//
// The reference is NumPy float32 on the host; segment indices must match it bit for bit, so every
// floating-point reduction here reproduces NumPy's evaluation ORDER (see oracle/segment_oracle.c):
//   * a 768-long float32 `.sum()` is NumPy's pairwise sum: 8 blocks of 96 elements, each block reduced by
//     8 interleaved sequential accumulators and a fixed 3-level tree, blocks combined by a 3-level tree.
//     A warp holds the 64 accumulators two per lane, so levels map to one in-lane add + 5 xor-shuffles.
//   * products are rounded before they are summed (no FMA contraction): __fmul_rn / __fadd_rn / __fdiv_rn.
//   * row means are sequential row-by-row accumulations followed by a division.
//   * `np.float32 ** .5` on a NumPy *scalar* is libm powf(x, .5f), not sqrtf: powf_half() below replays
//     glibc's powf (FMA build, as selected on every AVX2 x86-64 host) in fp64, see powf_tables.cuh.
// The scan over frames is serial inside a RUN (a maximal stretch of frames above the norm threshold: each decision
// depends on the running centroid), but runs are independent of each other: a masked-off frame resets the scan state
// (segment_utils.py:83-89), a mid-boundary only ever joins two segments of one run, and the refinement window stays
// inside those two segments (:110-128).  So the work item is a run, not an utterance: CTA (c, b) owns the runs of
// utterance b that START in frames [16 c, 16 c + 16) (SEG_CHUNK) and follows the last of them to its end; the segments of a CTA go
// to slots (first run start + k) of a frame-indexed table, which cannot collide with another CTA's because a segment has
// at least one frame; the last CTA of an utterance to finish compacts the table in frame order (segment_utils.py:130).
// One run spanning the whole utterance costs what the one-CTA-per-utterance kernel did; speech (pauses) and the synthetic
// bench states (15 % of the frames below the threshold) split into many.  Inside a run what bounds the scan is
// the LATENCY of one frame step, and the kernel is organised around that (round 3):
//   * the rows of the utterance stream into a shared-memory ring by bulk async copies (cp.async.bulk + mbarrier)
//     issued SEG_RING frames ahead, so no step waits on an L2 / HBM round trip;
//   * powf(|x_i|^2 + eps, .5) of every frame is computed by the parallel norm kernel;
//   * the merged centroid (curr * cnt + x) / (cnt + 1), its squared norm and the powf of that are computed
//     speculatively, concurrently with the cosine that decides whether the merge happens;
//   * the divisions by the small integer cnt + 1 use one reciprocal r = RN(1 / n) and Markstein's FMA correction
//     (q0 = RN(a r), rem = a - q0 n exact in one FMA, q = RN(q0 + rem r)): with a correctly rounded reciprocal this is the
//     correctly rounded quotient unless the significand of n is all ones (never for n <= 2^23) - bit-identical to NumPy's
//     division (tests/test_div_by_count.py); a single guard per step falls back to IEEE division if an operand is tiny.
#pragma once

#include "common.cuh"
#include "powf_tables.cuh"

namespace syl {

constexpr int SEG_D = 768;
constexpr int SEG_PER_LANE = SEG_D / 32;  // 24 features per lane = 2 accumulators x 12 terms

// element owned by (lane, accumulator a in {0,1}, term m in 0..11) under NumPy's pairwise order
__device__ __forceinline__ int seg_elem(int lane, int m) { return 96 * (lane >> 2) + 8 * m + 2 * (lane & 3); }

struct LaneVec {
  float v[SEG_PER_LANE];  // v[2*m + a] = x[seg_elem(lane, m) + a]
};

__device__ __forceinline__ void lane_load(LaneVec& d, const float* __restrict__ row, int lane) {
#pragma unroll
  for (int m = 0; m < 12; ++m) {
    const float2 t = *reinterpret_cast<const float2*>(row + seg_elem(lane, m));
    d.v[2 * m] = t.x;
    d.v[2 * m + 1] = t.y;
  }
}

// 16-byte aligned global -> shared bulk copy whose bytes complete on an mbarrier (no tensor map needed)
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared-memory image of one row: the 8 pairwise blocks of 96 floats padded to 104 so that the 8 lane groups, which
// read the same offset of different blocks, hit different banks
constexpr int SEG_BLK_PAD = 104;
constexpr int SEG_ROW_FLOATS = 8 * SEG_BLK_PAD;   // 832 floats = 3328 bytes
constexpr int SEG_RING = 12;                      // rows in flight (40 KB)

// finish NumPy's pairwise tree from the two in-lane accumulators; result identical in all lanes
__device__ __forceinline__ float pairwise_finish(float a0, float a1) {
  float s = __fadd_rn(a0, a1);                                   // r[2k] + r[2k+1]
#pragma unroll
  for (int o = 1; o <= 16; o <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  return __fadd_rn(0.0f, s);                                     // ndarray.sum() starts from the identity
}

// NumPy (x*y).sum(-1) over 768 float32
__device__ __forceinline__ float np_dot768(const LaneVec& x, const LaneVec& y) {
  float a0 = __fmul_rn(x.v[0], y.v[0]), a1 = __fmul_rn(x.v[1], y.v[1]);
#pragma unroll
  for (int m = 1; m < 12; ++m) {
    a0 = __fadd_rn(a0, __fmul_rn(x.v[2 * m], y.v[2 * m]));
    a1 = __fadd_rn(a1, __fmul_rn(x.v[2 * m + 1], y.v[2 * m + 1]));
  }
  return pairwise_finish(a0, a1);
}

// two dot products sharing one shuffle tree: (x*y).sum(), (x*x).sum()
__device__ __forceinline__ void np_dot768_pair(const LaneVec& x, const LaneVec& y, float& xy, float& xx) {
  float a0 = __fmul_rn(x.v[0], y.v[0]), a1 = __fmul_rn(x.v[1], y.v[1]);
  float b0 = __fmul_rn(x.v[0], x.v[0]), b1 = __fmul_rn(x.v[1], x.v[1]);
#pragma unroll
  for (int m = 1; m < 12; ++m) {
    a0 = __fadd_rn(a0, __fmul_rn(x.v[2 * m], y.v[2 * m]));
    a1 = __fadd_rn(a1, __fmul_rn(x.v[2 * m + 1], y.v[2 * m + 1]));
    b0 = __fadd_rn(b0, __fmul_rn(x.v[2 * m], x.v[2 * m]));
    b1 = __fadd_rn(b1, __fmul_rn(x.v[2 * m + 1], x.v[2 * m + 1]));
  }
  float s = __fadd_rn(a0, a1), q = __fadd_rn(b0, b1);
#pragma unroll
  for (int o = 1; o <= 16; o <<= 1) {
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, o));
  }
  xy = __fadd_rn(0.0f, s);
  xx = __fadd_rn(0.0f, q);
}

// NumPy pairwise sum of a contiguous float vector read by ONE thread (used for the short sweep sums)
__device__ float np_pairwise_serial(const float* a, int n) {
  if (n < 8) {
    float r = -0.0f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(np_pairwise_serial(a, n2), np_pairwise_serial(a + n2, n - n2));
}
__device__ __forceinline__ float np_sum_serial(const float* a, int n) { return __fadd_rn(0.0f, np_pairwise_serial(a, n)); }

// ----------------------------------------------------------------------------------------------
// per-frame squared norms + eps, one warp per frame:  nsq[i] = (x_i**2).sum() + 1e-8   (float32), and
// pw[i] = +-powf(nsq[i], .5f) - the denominator the scalar cossim of the scan uses for frame i (segment_utils.py:96),
// with the sign carrying the norm-threshold decision of segment_utils.py:76 (array path: sqrt): negative = masked off
// ----------------------------------------------------------------------------------------------
// It also resets what segment_kernel's CTAs share per utterance (scratch layout below): the end of slot i of the segment
// table to SEG_SLOT_UNUSED and, from frame 0, the count of finished CTAs.
constexpr int32_t SEG_SLOT_UNUSED = (int32_t)0x80000000;

__global__ void __launch_bounds__(256)
frame_sqnorm_kernel(const float* __restrict__ states, int rows, int T, float thr_norm, float* __restrict__ nsq, float* __restrict__ pw,
                    int32_t* __restrict__ scratch_all) {
  griddep_launch_dependents();
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = lane_id();
  LaneVec x;
  lane_load(x, states + (size_t)row * SEG_D, lane);
  const float s = np_dot768(x, x);
  if (lane == 0) {
    const float v = __fadd_rn(s, 1e-8f);
    nsq[row] = v;
    const float p = powf_half(v);        // > 0 for every finite v >= 1e-8; NaN stays NaN and compares "off" like NumPy's >=
    pw[row] = (__fsqrt_rn(v) >= thr_norm) ? p : -p;
    const int b = row / T, i = row - b * T;
    int32_t* scratch = scratch_all + (size_t)b * 6 * (T + 1);
    scratch[(T + 1) + i] = SEG_SLOT_UNUSED;      // seg_e[i]
    if (i == 0) scratch[T] = 0;                  // CTAs of this utterance that have finished (segment_kernel)
  }
}

// mean over rows [s, e) in NumPy order; every lane keeps its 24 features
__device__ __forceinline__ void lane_mean_rows(LaneVec& acc, const float* __restrict__ states, int s, int e, int lane) {
  if (e <= s) {
#pragma unroll
    for (int k = 0; k < SEG_PER_LANE; ++k) acc.v[k] = __int_as_float(0x7fc00000);
    return;
  }
  lane_load(acc, states + (size_t)s * SEG_D, lane);
  int r = s + 1;
  for (; r + 3 < e; r += 4) {            // four rows of loads in flight, added in row order
    LaneVec x0, x1, x2, x3;
    lane_load(x0, states + (size_t)r * SEG_D, lane);
    lane_load(x1, states + (size_t)(r + 1) * SEG_D, lane);
    lane_load(x2, states + (size_t)(r + 2) * SEG_D, lane);
    lane_load(x3, states + (size_t)(r + 3) * SEG_D, lane);
#pragma unroll
    for (int k = 0; k < SEG_PER_LANE; ++k)
      acc.v[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.v[k], x0.v[k]), x1.v[k]), x2.v[k]), x3.v[k]);
  }
  for (; r < e; ++r) {
    LaneVec x;
    lane_load(x, states + (size_t)r * SEG_D, lane);
#pragma unroll
    for (int k = 0; k < SEG_PER_LANE; ++k) acc.v[k] = __fadd_rn(acc.v[k], x.v[k]);
  }
  const float fn = (float)(e - s);
#pragma unroll
  for (int k = 0; k < SEG_PER_LANE; ++k) acc.v[k] = __fdiv_rn(acc.v[k], fn);
}

// ----------------------------------------------------------------------------------------------
// segmentation: grid (ceil(T / SEG_CHUNK), B) - CTA (c, b) owns the runs of utterance b that start in its chunk of frames.
//   states  [B, T, 768] fp32           pw [B, T]  from frame_sqnorm_kernel (sign = norm-threshold decision)
//   seg     [B, max_seg, 2] int32 out  seg_count [B] out
//   scratch [B, 6*(T+1)] int32/float workspace, six arrays of T + 1 per utterance:
//     seg_s, seg_e   slot table: a CTA whose first run starts at frame f0 keeps its k-th segment in slot f0 + k (its
//                    segments start at distinct frames >= f0, so slot f0 + k lies before the first frame of any later
//                    CTA's runs); seg_e < 0 = absorbed (-1 - end) or SEG_SLOT_UNUSED; seg_s[T] counts finished CTAs
//     mid_bd, mid_seg  the CTA's mid-boundaries, slots f0 + k likewise
//     sim_prev, sim_next  sweep cosines, indexed by frame (the window lies inside the CTA's own runs)
// ----------------------------------------------------------------------------------------------
constexpr int SEG_SCAN_THREADS = 96;     // warps 0, 1: the scan; warp 2: one thread that keeps the row ring full
#ifndef SYL_SEG_CHUNK
#define SYL_SEG_CHUNK 16
#endif
constexpr int SEG_CHUNK = SYL_SEG_CHUNK;  // frames whose run starts a CTA owns (one ballot, <= 32).  Measured stage time at
                                          // 32 x 10 s / 8 x 60 s: 32 -> 0.084 / 0.109 ms, 16 -> 0.074 / 0.106, 8 -> 0.074 / 0.115
static_assert(SEG_CHUNK >= 1 && SEG_CHUNK <= 32, "one lane per frame of the chunk");

__global__ void __launch_bounds__(SEG_SCAN_THREADS)
segment_kernel(const float* __restrict__ states_all, const float* __restrict__ pw_all, int T, float thr_merge,
               int32_t* __restrict__ seg_all, int32_t* __restrict__ seg_count, int max_seg, int32_t* __restrict__ scratch_all) {
  __shared__ __align__(128) float ring[SEG_RING][SEG_ROW_FLOATS];
  __shared__ __align__(8) uint64_t full_bar[SEG_RING];
  __shared__ __align__(8) uint64_t empty_bar[SEG_RING];
  __shared__ float2 part[2][2];          // [scoring-frame parity][warp]: partial sums of (curr*x) and (cand*cand)
  griddep_launch_dependents();
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * SEG_CHUNK;
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  const float* states = states_all + (size_t)b * T * SEG_D;
  const float* pw = pw_all + (size_t)b * T;
  int32_t* scratch = scratch_all + (size_t)b * 6 * (T + 1);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < SEG_RING; ++k) {
      mbar_init(&full_bar[k], 1);
      mbar_init(&empty_bar[k], 2);       // one arrival per scan warp
    }
    fence_barrier_init();
  }
  __syncthreads();
  griddep_wait();

  // ---- the frames this CTA scans: [f0, f1) = from the first run start inside its chunk to the end of the run that
  // starts last inside it (every warp evaluates this redundantly; a chunk without a run start has f0 = f1)
  int f0 = 0, f1 = 0;
  {
    const int i = c0 + lane;
    const bool start = lane < SEG_CHUNK && i < T && pw[i] > 0.0f && !(i > 0 && pw[i - 1] > 0.0f);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    if (m != 0u) {
      f0 = c0 + __ffs(m) - 1;
      f1 = T;
      for (int j0 = c0 + 32 - __clz(m); j0 < T; j0 += 32) {      // first masked-off frame behind the last run start
        const int idx = j0 + lane;
        const unsigned off = __ballot_sync(0xffffffffu, idx < T && !(pw[idx] > 0.0f));
        if (off != 0u) {
          f1 = j0 + __ffs(off) - 1;
          break;
        }
      }
    }
  }
  const int nfr = f1 - f0;
  int32_t* seg_s = scratch + f0;
  int32_t* seg_e = scratch + (T + 1) + f0;
  int32_t* mid_bd = scratch + 2 * (T + 1) + f0;
  int32_t* mid_seg = scratch + 3 * (T + 1) + f0;
  float* sim_prev_all = reinterpret_cast<float*>(scratch + 4 * (T + 1));
  float* sim_next_all = reinterpret_cast<float*>(scratch + 5 * (T + 1));

  if (w == 2) {
    // ---- producer: frame f0 + r -> ring slot r % SEG_RING as eight 384-byte bulk copies (one per pairwise
    // block, into the padded image), re-armed as soon as both scan warps have released the slot
    if (lane == 0) {
      for (int r = 0; r < nfr; ++r) {
        const int slot = r % SEG_RING;
        if (r >= SEG_RING) mbar_wait(&empty_bar[slot], (uint32_t)(r / SEG_RING - 1) & 1u);
        mbar_arrive_expect_tx(&full_bar[slot], SEG_D * sizeof(float));
        const float* src = states + (size_t)(f0 + r) * SEG_D;
#pragma unroll
        for (int k = 0; k < 8; ++k) bulk_copy_g2s(&ring[slot][SEG_BLK_PAD * k], src + 96 * k, 96 * sizeof(float), &full_bar[slot]);
      }
    }
    return;
  }

  int nseg = 0, nmid = 0;
  // ---- phase 1: greedy scan (segment_utils.py:79-108), TWO warps per utterance ----
  // Lane (g, a) of warp w owns accumulator a of pairwise block 4 w + g, i.e. the 12 elements 96 (4 w + g) + 8 m + a: the
  // in-lane chain is NumPy's sequential accumulation, xor-shuffles 1, 2, 4 are its tree inside a block, 8 and 16 the
  // tree over the warp's four blocks, and the two warps exchange one float2 through shared memory for the last level.
  // Both warps evaluate every decision redundantly (identical inputs, identical operations), so control flow stays
  // uniform.  The step is latency bound (one warp per scheduler, dependent chains), hence: selects instead of branches,
  // and powf of the merged centroid's norm deferred to the next scoring frame, where it overlaps that frame's dot product.
  {
    constexpr int E = 12;
    const int ring_off = SEG_BLK_PAD * (4 * w + (lane >> 3)) + (lane & 7);
    float curr[E];
#pragma unroll
    for (int m = 0; m < E; ++m) curr[m] = 0.0f;
    float p_curr = 1.0f;        // powf((curr**2).sum() + 1e-8, .5f), valid unless pow_pending
    float sq_pending = 1.0f;    // (curr**2).sum() + 1e-8 of a centroid merged in the previous scoring frame
    bool pow_pending = false;
    int cnt = 0, s = -1;
    int sc = 0;                 // scoring frames so far: the exchange buffer alternates per SCORING frame (frames in between
                                // take no barrier), so a buffer is rewritten only after the other warp passed the next barrier
    for (int r0 = 0; r0 < nfr; r0 += 32) {
      const float my_pw = pw[min(f0 + r0 + lane, T - 1)];
      const int n_here = min(32, nfr - r0);
      for (int j = 0; j < n_here; ++j) {
        const int r = r0 + j;
        const int i = f0 + r;
        const float pxs = __shfl_sync(0xffffffffu, my_pw, j);
        const bool on = pxs > 0.0f;
        const float px = fabsf(pxs);
        const int slot = r % SEG_RING;
        mbar_wait(&full_bar[slot], (uint32_t)(r / SEG_RING) & 1u);
        float x[E];
        if (on) {
          const float* row = &ring[slot][ring_off];
#pragma unroll
          for (int m = 0; m < E; ++m) x[m] = row[8 * m];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);    // every lane's loads of the slot are issued: the producer may refill it
        if (!on) {
          if (s > -1) {
            if (threadIdx.x == 0) { seg_s[nseg] = s; seg_e[nseg] = i; }
            ++nseg;
          }
          s = -1;
          cnt = 0;
          continue;
        }
        if (cnt == 0) {                                  // first frame of a run
#pragma unroll
          for (int m = 0; m < E; ++m) curr[m] = x[m];
          p_curr = px;
          pow_pending = false;
          cnt = 1;
          s = i;
          continue;
        }
        // denominator of the current centroid (independent of everything below until the division)
        const float p_deferred = powf_half(sq_pending);
        p_curr = pow_pending ? p_deferred : p_curr;
        // the decision's dot product ...
        float a0 = __fmul_rn(curr[0], x[0]);
#pragma unroll
        for (int m = 1; m < E; ++m) a0 = __fadd_rn(a0, __fmul_rn(curr[m], x[m]));
        // ... and, speculatively, the merged centroid with its squared norm (independent of the decision)
        const float fc = (float)cnt, fc1 = (float)(cnt + 1), rc1 = __frcp_rn(fc1);
        float num[E], cand[E];
        bool tiny = false;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          num[m] = __fadd_rn(__fmul_rn(curr[m], fc), x[m]);
          const float q0 = __fmul_rn(num[m], rc1);
          cand[m] = __fmaf_rn(__fmaf_rn(-q0, fc1, num[m]), rc1, q0);     // reciprocal + FMA correction, unguarded ...
          tiny |= !(fabsf(num[m]) > 1e-30f);
        }
        if (__any_sync(0xffffffffu, tiny)) {                              // ... which is taken once for all elements
#pragma unroll
          for (int m = 0; m < E; ++m) cand[m] = __fdiv_rn(num[m], fc1);
        }
        float b0 = __fmul_rn(cand[0], cand[0]);
#pragma unroll
        for (int m = 1; m < E; ++m) b0 = __fadd_rn(b0, __fmul_rn(cand[m], cand[m]));
#pragma unroll
        for (int o = 1; o <= 16; o <<= 1) {
          a0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, o));
          b0 = __fadd_rn(b0, __shfl_xor_sync(0xffffffffu, b0, o));
        }
        if (lane == 0) part[sc & 1][w] = make_float2(a0, b0);
        named_bar_sync(1, 64);                           // the two scan warps' partials are visible to each other
        const float2 other = part[sc & 1][w ^ 1];
        ++sc;
        const float xy = __fadd_rn(0.0f, __fadd_rn(a0, other.x));
        const float sq_cand = __fadd_rn(__fadd_rn(0.0f, __fadd_rn(b0, other.y)), 1e-8f);
        const float sim = __fdiv_rn(__fdiv_rn(xy, p_curr), px);   // scalar path: powf denominators
        const bool merge = sim >= thr_merge;
        cnt += 1;               // also after a split: reference quirk (segment_utils.py:103)
#pragma unroll
        for (int m = 0; m < E; ++m) curr[m] = merge ? cand[m] : x[m];
        pow_pending = merge;
        sq_pending = sq_cand;
        p_curr = merge ? p_curr : px;
        if (!merge) {
          if (threadIdx.x == 0) { seg_s[nseg] = s; seg_e[nseg] = i; mid_bd[nmid] = i; mid_seg[nmid] = nseg; }
          ++nseg;
          ++nmid;
          s = i;
        }
      }
    }
    if (s > -1) {                                        // closed by the masked-off frame f1, or by the end of the utterance
      if (threadIdx.x == 0) { seg_s[nseg] = s; seg_e[nseg] = f1; }
      ++nseg;
    }
  }
  if (w != 0) return;      // phases 2 and 3 run on warp 0 (its lane 0 wrote the segment arrays)
  __syncwarp();

  // ---- phase 2: boundary merge / refinement, in order (segment_utils.py:110-128) ----
  // dead flags are kept as negative ends to avoid another array: seg_e < 0 marks an absorbed segment
  for (int m = 0; m < nmid; ++m) {
    const int bd = mid_bd[m], a = mid_seg[m];
    if (a >= nseg - 1) continue;
    const int bsg = a + 1;
    const int as = seg_s[a], ae = seg_e[a], bs = seg_s[bsg], be = seg_e[bsg];
    LaneVec ca, cb;
    lane_mean_rows(ca, states, as, ae, lane);
    lane_mean_rows(cb, states, bs, be, lane);
    float ab, aa, bb, tmp;
    np_dot768_pair(ca, cb, ab, aa);
    np_dot768_pair(cb, cb, bb, tmp);
    aa = __fadd_rn(aa, 1e-8f);
    bb = __fadd_rn(bb, 1e-8f);
    const float simc = __fdiv_rn(__fdiv_rn(ab, powf_half(aa)), powf_half(bb));
    if (simc >= thr_merge) {
      __syncwarp();
      if (lane == 0) { seg_s[bsg] = as; seg_e[a] = -1 - ae; }
      __syncwarp();
      continue;
    }
    const int la = ae - as, lb = be - bs;
    const int lo = max(as, bd - max(1, la / 2));
    const int hi = min(be, bd + max(1, lb / 2));
    const int W = hi - lo;
    float* sim_prev = sim_prev_all + lo;
    float* sim_next = sim_next_all + lo;
    const float na = __fsqrt_rn(aa), nb = __fsqrt_rn(bb);   // 2-D path of cossim: sqrt
    for (int r = 0; r < W; ++r) {
      LaneVec x;
      lane_load(x, states + (size_t)(lo + r) * SEG_D, lane);
      float xa, xx, xb;
      np_dot768_pair(x, ca, xa, xx);
      xb = np_dot768(x, cb);
      const float nx = __fsqrt_rn(__fadd_rn(xx, 1e-8f));
      if (lane == 0) {
        sim_prev[r] = __fdiv_rn(__fdiv_rn(xa, nx), na);
        sim_next[r] = __fdiv_rn(__fdiv_rn(xb, nx), nb);
      }
    }
    __syncwarp();
    float best_v = 0.0f;
    int best_i = 0x7fffffff;
    for (int i = lane; i < W; i += 32) {
      const float v = __fadd_rn(np_sum_serial(sim_prev, i), np_sum_serial(sim_next + i, W - i));
      if (best_i == 0x7fffffff || v > best_v) { best_v = v; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (oi != 0x7fffffff && (best_i == 0x7fffffff || ov > best_v || (ov == best_v && oi < best_i))) {
        best_v = ov;
        best_i = oi;
      }
    }
    const int opt = lo + best_i;
    __syncwarp();
    if (lane == 0) { seg_e[a] = opt; seg_s[bsg] = opt; }
    __syncwarp();
  }

  // ---- phase 3: drop absorbed segments (segment_utils.py:130-131) ----
  // This CTA's slots are final.  The last CTA of the utterance to get here packs the slot table in frame order
  // (fence + counter: the pattern of the threadFenceReduction sample; the counter was zeroed by frame_sqnorm_kernel).
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence();
    last = atomicAdd(scratch + T, 1) == (int)gridDim.x - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  {
    const int32_t* all_s = scratch;
    const int32_t* all_e = scratch + (T + 1);
    int32_t* out = seg_all + (size_t)b * max_seg * 2;
    int n = 0;
    for (int i0 = 0; i0 < T; i0 += 128) {
      int32_t e[4], st[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {      // four independent pairs of loads in flight
        const int i = i0 + 32 * u + lane;
        e[u] = i < T ? __ldcg(all_e + i) : SEG_SLOT_UNUSED;
        st[u] = i < T ? __ldcg(all_s + i) : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool alive = e[u] >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        const int pos = n + __popc(m & ((1u << lane) - 1u));
        if (alive && pos < max_seg) {
          out[2 * pos] = st[u];
          out[2 * pos + 1] = e[u];
        }
        n += __popc(m);
      }
    }
    if (lane == 0) seg_count[b] = n;
  }
}

// ----------------------------------------------------------------------------------------------
// segment features (sylber.py:133): mean of the hidden states over each segment, NumPy order.
// grid (min(max_seg, SEG_POOL_GRID), B), block 192: each thread owns 4 consecutive features; a block walks the segments
// sidx, sidx + gridDim.x, ... (a grid of max_seg blocks per utterance spent most of its time launching blocks without a segment).
// ----------------------------------------------------------------------------------------------
constexpr int SEG_POOL_GRID = 128;

__global__ void __launch_bounds__(192)
segment_pool_kernel(const float* __restrict__ states_all, int T, const int32_t* __restrict__ seg_all,
                    const int32_t* __restrict__ seg_count, int max_seg, float* __restrict__ feat_all) {
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.y;
  const int n = min(seg_count[b], max_seg);
  const float* st = states_all + (size_t)b * T * SEG_D + threadIdx.x * 4;
  for (int sidx = blockIdx.x; sidx < n; sidx += gridDim.x) {
    const int s = seg_all[((size_t)b * max_seg + sidx) * 2], e = seg_all[((size_t)b * max_seg + sidx) * 2 + 1];
    float4 acc;
    if (e <= s) {
      const float nan = __int_as_float(0x7fc00000);
      acc = make_float4(nan, nan, nan, nan);
    } else {
      acc = *reinterpret_cast<const float4*>(st + (size_t)s * SEG_D);
      for (int r = s + 1; r < e; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(st + (size_t)r * SEG_D);
        acc.x = __fadd_rn(acc.x, x.x);
        acc.y = __fadd_rn(acc.y, x.y);
        acc.z = __fadd_rn(acc.z, x.z);
        acc.w = __fadd_rn(acc.w, x.w);
      }
      const float fn = (float)(e - s);
      acc.x = __fdiv_rn(acc.x, fn);
      acc.y = __fdiv_rn(acc.y, fn);
      acc.z = __fdiv_rn(acc.z, fn);
      acc.w = __fdiv_rn(acc.w, fn);
    }
    *reinterpret_cast<float4*>(feat_all + ((size_t)b * max_seg + sidx) * SEG_D + threadIdx.x * 4) = acc;
  }
}

}  // namespace syl
