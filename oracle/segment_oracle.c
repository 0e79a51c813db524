/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
 *
 * Plain-C restatement of the reference's syllable segmentation,
 *     sylber/utils/segment_utils.py:68-69   cossim
 *     sylber/utils/segment_utils.py:72-131  get_segment
 * and of the per-segment mean pooling at sylber/model/sylber.py:133.
 *
 * The reference runs on NumPy float32.  "Bit-exact segments" therefore means reproducing NumPy's
 * float32 evaluation order, which this file does explicitly:
 *   - ndarray.sum() over a contiguous float32 vector is NumPy's pairwise summation
 *     (blocks of <=128 elements with 8 interleaved partial sums, halves split at multiples of 8);
 *   - x.mean(0) over rows is a sequential row-by-row float32 accumulation followed by a division;
 *   - `arr ** .5` on an ndarray is sqrtf, but `np.float32_scalar ** .5` calls libm powf(x, 0.5f),
 *     which differs from sqrtf in ~0.06 % of inputs (measured exhaustively).  The
 *     scalar form is what cossim() hits when both arguments are 1-D (segment_utils.py:96,114-115),
 *     the array form when the first argument is 2-D (segment_utils.py:123-124);
 *   - Python-scalar thresholds compare in float32 (NumPy 2 weak-scalar promotion).
 * Pinned against the unmodified reference in tests/test_oracle_segment.py (reference imported from
 * /root/reference when present) and against tests/golden/segment_cases.npz.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).  -ffp-contract=off matters:
 * NumPy never fuses a multiply with the following add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SYL_D 768

/* NumPy's FLOAT_pairwise_sum for unit stride. */
static float np_pairwise_sum(const float *a, int64_t n) {
  if (n < 8) {
    float res = -0.0f;
    for (int64_t i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

/* ndarray.sum(): the reduction starts from the additive identity. */
static float np_sum(const float *a, int64_t n) { return 0.0f + np_pairwise_sum(a, n); }

static float dot_np(const float *x, const float *y, int d, float *tmp) {
  for (int k = 0; k < d; k++) tmp[k] = x[k] * y[k];
  return np_sum(tmp, d);
}

/* the `(v**2).sum(-1) + 1e-8` part of cossim */
static float sqnorm_eps(const float *x, int d, float *tmp) { return dot_np(x, x, d, tmp) + 1e-8f; }

/* mean over rows [s, e) of a (T, d) matrix, NumPy order: acc = row s; acc += row s+1; ...; acc / n */
static void mean_rows(const float *states, int d, int s, int e, float *out) {
  int n = e - s;
  if (n <= 0) {
    for (int k = 0; k < d; k++) out[k] = NAN; /* empty slice: NumPy's 0/0 */
    return;
  }
  memcpy(out, states + (size_t)s * d, sizeof(float) * d);
  for (int r = s + 1; r < e; r++) {
    const float *row = states + (size_t)r * d;
    for (int k = 0; k < d; k++) out[k] += row[k];
  }
  float fn = (float)n;
  for (int k = 0; k < d; k++) out[k] = out[k] / fn;
}

/* cossim of two 1-D vectors: NumPy scalar path, `** .5` is powf */
static float cossim_scalar(const float *x, const float *y, int d, float *tmp) {
  float xy = dot_np(x, y, d, tmp);
  float nx = powf(sqnorm_eps(x, d, tmp), 0.5f);
  float ny = powf(sqnorm_eps(y, d, tmp), 0.5f);
  return xy / nx / ny;
}

/* Segment one utterance.
 *   states   (T, d) float32, row major
 *   seg_out  capacity T rows of (start, end); returns the number of segments
 *   norms_in optional precomputed norms (the reference's `norms=` argument), else NULL
 */
int64_t syl_oracle_get_segment(const float *states, int64_t T64, int64_t d64, float norm_thr, float merge_thr,
                               const float *norms_in, int64_t *seg_out) {
  const int T = (int)T64, d = (int)d64;
  if (T == 0) return 0;
  float *tmp = (float *)malloc(sizeof(float) * d);
  float *curr = (float *)malloc(sizeof(float) * d);
  float *ca = (float *)malloc(sizeof(float) * d);
  float *cb = (float *)malloc(sizeof(float) * d);
  int *seg_s = (int *)malloc(sizeof(int) * (T + 1));
  int *seg_e = (int *)malloc(sizeof(int) * (T + 1));
  char *dead = (char *)calloc(T + 1, 1);
  int *mid_bd = (int *)malloc(sizeof(int) * (T + 1));
  int *mid_seg = (int *)malloc(sizeof(int) * (T + 1));
  float *sim_prev = (float *)malloc(sizeof(float) * (T + 1));
  float *sim_next = (float *)malloc(sizeof(float) * (T + 1));
  int nseg = 0, nmid = 0;

  /* phase 1: segment_utils.py:74-108 */
  int cnt = 0, s = -1;
  for (int i = 0; i < T; i++) {
    const float *x = states + (size_t)i * d;
    float norm = norms_in ? norms_in[i] : sqrtf(sqnorm_eps(x, d, tmp)); /* 2-D array path: sqrt */
    if (!(norm >= norm_thr)) {
      if (s > -1) {
        seg_s[nseg] = s;
        seg_e[nseg] = i;
        nseg++;
      }
      s = -1;
      cnt = 0;
    } else if (cnt == 0) {
      memcpy(curr, x, sizeof(float) * d);
      cnt = 1;
      s = i;
    } else {
      float sim = cossim_scalar(curr, x, d, tmp);
      if (sim >= merge_thr) {
        float fc = (float)cnt, fc1 = (float)(cnt + 1);
        for (int k = 0; k < d; k++) curr[k] = (curr[k] * fc + x[k]) / fc1;
        cnt += 1;
      } else {
        memcpy(curr, x, sizeof(float) * d);
        cnt += 1; /* the reference does not reset the count here (segment_utils.py:103) */
        seg_s[nseg] = s;
        seg_e[nseg] = i;
        nseg++;
        mid_bd[nmid] = i;
        mid_seg[nmid] = nseg - 1;
        nmid++;
        s = i;
      }
    }
  }
  if (s > -1) {
    seg_s[nseg] = s;
    seg_e[nseg] = T;
    nseg++;
  }

  /* phase 2: segment_utils.py:110-128 */
  for (int m = 0; m < nmid; m++) {
    int bd = mid_bd[m], a = mid_seg[m];
    if (a >= nseg - 1) continue;
    int b = a + 1;
    mean_rows(states, d, seg_s[a], seg_e[a], ca);
    mean_rows(states, d, seg_s[b], seg_e[b], cb);
    if (cossim_scalar(ca, cb, d, tmp) >= merge_thr) {
      seg_s[b] = seg_s[a];
      dead[a] = 1;
      continue;
    }
    int la = seg_e[a] - seg_s[a], lb = seg_e[b] - seg_s[b];
    int ha = la / 2 > 1 ? la / 2 : 1, hb = lb / 2 > 1 ? lb / 2 : 1;
    int lo = seg_s[a] > bd - ha ? seg_s[a] : bd - ha;
    int hi = seg_e[b] < bd + hb ? seg_e[b] : bd + hb;
    int W = hi - lo;
    /* 2-D array path of cossim: norms via sqrtf */
    float na = sqrtf(sqnorm_eps(ca, d, tmp)), nb = sqrtf(sqnorm_eps(cb, d, tmp));
    for (int r = 0; r < W; r++) {
      const float *x = states + (size_t)(lo + r) * d;
      float nx = sqrtf(sqnorm_eps(x, d, tmp));
      sim_prev[r] = dot_np(x, ca, d, tmp) / nx / na;
      sim_next[r] = dot_np(x, cb, d, tmp) / nx / nb;
    }
    int best = 0;
    float best_v = 0.0f;
    for (int i = 0; i < W; i++) {
      float v = np_sum(sim_prev, i) + np_sum(sim_next + i, W - i);
      if (i == 0 || v > best_v) { /* np.argmax: first maximum */
        best_v = v;
        best = i;
      }
    }
    int opt = lo + best;
    seg_e[a] = opt;
    seg_s[b] = opt;
  }

  /* phase 3: segment_utils.py:130-131 */
  int64_t n_out = 0;
  for (int i = 0; i < nseg; i++) {
    if (dead[i]) continue;
    seg_out[2 * n_out] = seg_s[i];
    seg_out[2 * n_out + 1] = seg_e[i];
    n_out++;
  }
  free(tmp); free(curr); free(ca); free(cb); free(seg_s); free(seg_e); free(dead);
  free(mid_bd); free(mid_seg); free(sim_prev); free(sim_next);
  return n_out;
}

/* sylber/model/sylber.py:133: per-segment mean of the hidden states, float32 */
void syl_oracle_segment_mean(const float *states, int64_t d64, const int64_t *seg, int64_t n_seg, float *out) {
  const int d = (int)d64;
  for (int64_t i = 0; i < n_seg; i++) mean_rows(states, d, (int)seg[2 * i], (int)seg[2 * i + 1], out + (size_t)i * d);
}

/* exposed so tests can pin the summation order against NumPy directly */
float syl_oracle_np_sum(const float *a, int64_t n) { return np_sum(a, n); }
float syl_oracle_powf_half(float x) { return powf(x, 0.5f); }
