// C ABI of libsylber_b200.so (declared in include/sylber_b200.h): weight packing, workspace planning,
// TMA descriptor construction and the launch sequence of the Segmenter forward path
// (sylber/model/sylber.py:122-133 -> transformers HubertModel.forward, get_segment, segment means).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sylber_b200.h"
#include "attention.cuh"
#include "frontend.cuh"
#include "frontdoor.cuh"
#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "gemm3_tc.cuh"
#include "posconv.cuh"
#ifdef SYL_DIAG
#include "mma_probe.cuh"
#endif
#include "segment.cuh"

using namespace syl;

namespace {

constexpr int kConvK[7] = {10, 3, 3, 3, 3, 2, 2};
constexpr int kConvS[7] = {5, 2, 2, 2, 2, 2, 2};
constexpr int kC = 512;        // conv channels
constexpr int kH = 768;        // hidden size
constexpr int kHeads = 12;
constexpr int kF = 3072;       // FFN size
constexpr int kPosK = 128;
constexpr int kPosG = 16;
constexpr int kPosCg = 48;     // channels per group

std::string g_create_error;

// device-time profiling by stage (CUDA events on the launch stream, enabled by syl_profile_enable)
enum Stage { ST_CONV0 = 0, ST_CONV, ST_LN, ST_PROJ, ST_POS, ST_QKV, ST_ATTN, ST_OUT, ST_FFN1, ST_FFN2, ST_SEG, ST_CONV1, ST_LN_ENC, ST_COUNT };
const char* const kStageNames[ST_COUNT] = {"conv0_gn_gelu", "conv2_6_gemm", "layernorm", "feature_proj_gemm", "pos_conv_gemm",
                                           "qkv_gemm", "attention", "out_proj_gemm", "ffn1_gemm", "ffn2_gemm", "segment_pool", "conv1_gemm", "layernorm_encoder"};
struct ProfRec {
  int stage;
  cudaEvent_t a, b;
};

// ------------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tiled tensor map of up to 3 dims; elem_bytes 2 (fp16) or 4 (fp32); swizzle_bytes 128, 64 or 0 (none); box = {box0, box1, 1}
bool make_tmap(CUtensorMap* map, const void* base, int elem_bytes, int rank, const uint64_t* dims,
               const uint64_t* strides_elems, uint32_t box0, uint32_t box1, int swizzle_bytes, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    *err = "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)";
    return false;
  }
  cuuint64_t gdim[3] = {1, 1, 1};
  cuuint64_t gstr[2] = {0, 0};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i];
  for (int i = 1; i < rank; ++i) gstr[i - 1] = strides_elems[i] * (uint64_t)elem_bytes;
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu strides %llu %llu box %u %u",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
             (unsigned long long)gstr[0], (unsigned long long)gstr[1], box0, box1);
    *err = buf;
    return false;
  }
  return true;
}

// fp16 operand map: 128B swizzle, box = {64, box1, 1}
bool make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   uint32_t box1, std::string* err) {
  return make_tmap(map, base, 2, rank, dims, strides_elems, 64, box1, 128, err);
}

// ------------------------------------------------------------------------------------------------
// small utility kernels (weight packing, dtype conversion)
// ------------------------------------------------------------------------------------------------
__global__ void split_f32_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                                 size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    __half h, l;
    split_f16(in[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

__global__ void join_f16_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out,
                                size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __half2float(hi[i]) + (lo ? __half2float(lo[i]) : 0.0f);
}

// conv weight [co][ci][k] fp32 -> GEMM B operand [co][j*C + ci] fp16 hi/lo
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int k, __half* __restrict__ hi,
                                   __half* __restrict__ lo) {
  const size_t n = (size_t)kC * kC * k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % kC;
    const int j = (i / kC) % k;
    const int co = i / ((size_t)kC * k);
    __half h, l;
    split_f16(w[((size_t)co * kC + ci) * k + j], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// weight-norm denominator of the positional conv: norm over (out, in) for each tap (dim=2)
__global__ void pos_tap_norm_kernel(const float* __restrict__ v, float* __restrict__ norm) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < kH * kPosCg; i += blockDim.x) {
    const double x = v[(size_t)i * kPosK + j];
    s += x * x;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[j] = (float)sqrt(red[0]);
}

// pos conv weight v [co][ci(48)][j(128)], g [128] -> B operand [co][j*64 + ci] (ci padded to 64 with zeros)
__global__ void pack_pos_w_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                  const float* __restrict__ norm, __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t n = (size_t)kH * kPosK * 64;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % 64;
    const int j = (i / 64) % kPosK;
    const int co = i / (64 * kPosK);
    float w = 0.0f;
    if (ci < kPosCg) w = v[((size_t)co * kPosCg + ci) * kPosK + j] * (g[j] / norm[j]);
    __half h, l;
    split_f16(w, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

__global__ void add_inplace_kernel(float* __restrict__ x, const float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] += y[i];
}

// valid[b] = frames of utterance b that come from real samples (modeling_hubert.py:675-700), clamped to [1, T];
// valid[(1 + i) * batch + b] = rows of conv layer i (0..6) those frames read - what the trimmed mode computes
// (backwards through the receptive fields: rows_i = (rows_{i+1} - 1) * stride_{i+1} + kernel_{i+1})
__global__ void valid_frames_kernel(const int32_t* __restrict__ n_samples, int batch, int T, int32_t* __restrict__ valid) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int k[7] = {10, 3, 3, 3, 3, 2, 2}, s[7] = {5, 2, 2, 2, 2, 2, 2};
  int n = n_samples ? n_samples[b] : 0x3fffffff;
  if (n_samples) {
    for (int i = 0; i < 7; ++i) n = (n >= k[i]) ? (n - k[i]) / s[i] + 1 : 0;
  }
  const int v = max(1, min(n, T));
  valid[b] = v;
  int need = v;
  valid[7 * batch + b] = need;                       // conv6 output rows = frames
  for (int i = 5; i >= 0; --i) {
    need = (need - 1) * s[i + 1] + k[i + 1];
    valid[(1 + i) * batch + b] = need;
  }
}

// trimmed mode: order[k] = index of the utterance with the k-th largest valid length (ties by index); one block, B <= 1024
__global__ void length_order_kernel(const int32_t* __restrict__ valid, int batch, int32_t* __restrict__ order) {
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const int v = valid[b];
    int rank = 0;
    for (int o = 0; o < batch; ++o) {
      const int vo = valid[o];
      rank += (vo > v) || (vo == v && o < b);
    }
    order[rank] = b;
  }
}

__global__ void fill_i32_kernel(int32_t* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// fp16 range check (diagnostic, off the hot path): counts elements that sit at the saturation value +-65504 or are not
// finite.  Every fp16 store of the forward saturates instead of overflowing (pack_f16x2_sat, split_pair), so a non-zero
// count means an activation left the fp16 range and the result of that forward is not trustworthy in this precision.
__global__ void count_saturated_kernel(const __half* __restrict__ x, size_t n, unsigned long long* __restrict__ count) {
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned short b = __half_as_ushort(x[i]) & 0x7fffu;
    local += (b >= 0x7bffu) ? 1u : 0u;
  }
  if (local) atomicAdd(count, local);
}

__global__ void powf_half_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = powf_half(x[i]);
}

// Launch with the programmatic-dependent-launch attribute (common.cuh: every kernel launched this way waits for its
// predecessor with griddepcontrol.wait after its own prologue).  Measured in round 2 (profiles/r02_bench_ab.md): the
// device-resident step does not change beyond run-to-run noise (4.78-5.08 vs 4.86-4.94 ms), and the end-to-end path,
// which keeps three prioritised sub-batch streams in flight, gets SLOWER (6.4 vs 5.8 ms): early-launched dependents
// of one stream hold scheduling slots the other streams need.  Hence opt-in: SYL_PDL=1.
// The product build reads NO environment variables: every experiment switch below exists only in the diagnostic
// build (-DSYL_DIAG, sylber_b200/_lib.py build_library(diag=True) -> libsylber_b200_diag.so).
#ifdef SYL_DIAG
int diag_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool pdl_enabled() {
  static const int v = diag_env("SYL_PDL", 0);
  return v != 0;
}
#else
constexpr bool pdl_enabled() { return false; }
#endif

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(size_t n, int block = 256) { return (int)std::min<size_t>((n + block - 1) / block, 148 * 16); }

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct PackedLinear {   // B operand [N][K] fp16 hi/lo + tensor maps + fp32 bias
  __half* hi = nullptr;
  __half* lo = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
  CUtensorMap map_hi, map_lo;       // box {64, 256}: whole B tile (1-CTA GEMM) / {64, 48} for the positional conv
  CUtensorMap map2_hi, map2_lo;     // box {64, 128}: one CTA's half of the B tile (2-CTA GEMM)
};

struct LayerW {
  PackedLinear qkv, out, ffn1, ffn2;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

struct GemmOp {
  CUtensorMap a_hi, a_lo;
  CUtensorMap o3_f32, o3_hi, o3_lo;     // fp32 {16, 32} SWIZZLE_64B; fp16 {32, 32} SWIZZLE_64B (hi only) or {16, 32} plain
  const PackedLinear* w = nullptr;
  GemmParams p;
};

struct PosOp {
  CUtensorMap a_hi, a_lo, o_map, o_map31;
  PosConvParams p;
};

struct WsLayout {
  size_t total = 0;
  size_t mom, gn_scale, gn_shift, valid, nsq, seg_scratch;
  size_t act_hi[6], act_lo[6], conv6;
  size_t ln_hi, ln_lo, h, h16_hi, h16_lo, pos, pre, qkv, ctx_hi, ctx_lo, mid_hi, mid_lo;
  int L[7];
  int T;
};

struct Plan {
  bool valid = false;
  uint64_t last_use = 0;   // LRU stamp
  int uses = 0;            // forwards run with this plan: a CUDA graph is only captured from the second one on
  int batch = 0, t_samp = 0;
  bool order_valid = false;   // trimmed mode: the length-order array of this plan's workspace has been written by a forward
  void* ws = nullptr;
  float* hidden = nullptr;
  WsLayout lay;
  GemmOp conv[6];
  GemmOp proj;
  PosOp pos;
  std::vector<GemmOp> qkv, out, ffn1, ffn2;   // per layer (A maps are shared, params differ in weights)
  CUtensorMap attn_map, ctx_hi_map, ctx_lo_map;
  // CUDA graphs of the whole forward, one per distinct argument set (pointers + thresholds)
  struct GraphEntry {
    const void* key[6];
    int max_seg, active_layers;
    float thr_norm, thr_merge;
    cudaGraphExec_t exec;
  };
  std::vector<GraphEntry> graphs;
};

}  // namespace

// Entry points run on the handle's device and leave the caller's current device as they found it.
struct DeviceGuard {
  int prev = -1, dev;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

struct syl_handle {
  std::mutex mu;      // plan cache, graph cache and profiling records are per-handle state: entry points serialise on it
  int device = 0;
  int n_layers = 9;
  int active_layers = -1;
  int mode = SYL_MODE_FAST;
  bool trim = false;      // SYL_TRIM_PADDING: padded frames are not computed (opt-in deviation, include/sylber_b200.h)
  bool finalized = false;
  std::string err;
  std::map<std::string, std::pair<float*, std::vector<int64_t>>> raw;
  std::vector<void*> owned;
  // packed weights
  float* conv0_w = nullptr;
  uint4* conv0_bfrag = nullptr;    // mma.sync B fragments of conv0 (frontend.cuh)
  float *gn_g = nullptr, *gn_b = nullptr;
  PackedLinear convw[6];
  float *fp_ln_g = nullptr, *fp_ln_b = nullptr;
  PackedLinear proj;
  PackedLinear pos;
  float *enc_ln_g = nullptr, *enc_ln_b = nullptr;
  std::vector<LayerW> layers;
  static constexpr int kPlans = 16;   // LRU cache: sub-batch streams and length buckets alternate between shapes / workspaces
  Plan plans[kPlans];
  int plan_cur = 0;
  uint64_t plan_clock = 0;
  int sm_count = 148;
  bool profile = false;
  bool use_graphs = true;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  size_t pool_used = 0;
};

namespace {

// does conv layer i (1..6) run split precision?
bool conv_split(int mode, int i) {
  static const int bit[7] = {0, SYL_SPLIT_CONV1, SYL_SPLIT_CONV2, SYL_SPLIT_CONV3, SYL_SPLIT_CONV4, SYL_SPLIT_CONV5, SYL_SPLIT_CONV6};
  return i >= 1 && i <= 6 && (mode & bit[i]) != 0;
}

int fail(syl_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

// Launch checks read cudaGetLastError(), which also returns whatever error ANOTHER user of the CUDA runtime on this thread
// left behind (PyTorch, a profiler, a query that answered "not ready"): a stale error would be blamed on our launch and
// fail a call that did nothing wrong.  Every entry point that launches therefore drops a pending non-sticky error first
// (SYL_ENTER); a sticky one - a device fault - is returned by every later runtime call anyway and still surfaces.
// launch_ok() keeps the error it consumed so that the message can name it.
#define SYL_ENTER() (void)cudaGetLastError()
thread_local cudaError_t g_launch_err = cudaSuccess;
inline bool launch_ok() {
  g_launch_err = cudaGetLastError();
  return g_launch_err == cudaSuccess;
}
inline const char* launch_err() { return cudaGetErrorString(g_launch_err); }

cudaEvent_t prof_event(syl_handle* h) {
  if (h->pool_used == h->pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->pool.push_back(e);
  }
  return h->pool[h->pool_used++];
}

struct StageTimer {
  syl_handle* h;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  StageTimer(syl_handle* h_, int stage, cudaStream_t st_) : h(h_), st(st_) {
    if (!h->profile) return;
    cudaEvent_t a = prof_event(h);
    b = prof_event(h);
    cudaEventRecord(a, st);
    h->recs.push_back({stage, a, b});
  }
  ~StageTimer() {
    if (b) cudaEventRecord(b, st);
  }
};

#define CUDA_TRY(h, expr)                                                                        \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(h, SYL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

template <typename T>
T* dev_alloc(syl_handle* h, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  h->owned.push_back(p);
  return reinterpret_cast<T*>(p);
}

const float* raw_ptr(syl_handle* h, const std::string& name, size_t expect_elems) {
  auto it = h->raw.find(name);
  if (it == h->raw.end()) {
    fail(h, SYL_E_STATE, "missing weight '%s'", name.c_str());
    return nullptr;
  }
  size_t n = 1;
  for (int64_t d : it->second.second) n *= (size_t)d;
  if (n != expect_elems) {
    fail(h, SYL_E_STATE, "weight '%s' has %zu elements, expected %zu", name.c_str(), n, expect_elems);
    return nullptr;
  }
  return it->second.first;
}

bool make_weight_maps(syl_handle* h, PackedLinear& w, int block_n) {
  uint64_t dims[2] = {(uint64_t)w.K, (uint64_t)w.N};
  uint64_t str[2] = {1, (uint64_t)w.K};
  return make_tmap_f16(&w.map_hi, w.hi, 2, dims, str, block_n, &h->err) &&
         make_tmap_f16(&w.map_lo, w.lo, 2, dims, str, block_n, &h->err) &&
         make_tmap_f16(&w.map2_hi, w.hi, 2, dims, str, std::min(block_n, 128), &h->err) &&
         make_tmap_f16(&w.map2_lo, w.lo, 2, dims, str, std::min(block_n, 128), &h->err);
}

// pack a torch Linear weight [N][K] (+ bias) ; several sources may be concatenated along N (QKV)
int pack_linear(syl_handle* h, PackedLinear& out, const std::vector<std::string>& wnames,
                const std::vector<std::string>& bnames, int n_each, int K) {
  const int N = n_each * (int)wnames.size();
  out.N = N;
  out.K = K;
  out.hi = dev_alloc<__half>(h, (size_t)N * K);
  out.lo = dev_alloc<__half>(h, (size_t)N * K);
  out.bias = dev_alloc<float>(h, N);
  if (!out.hi || !out.lo || !out.bias) return fail(h, SYL_E_CUDA, "cudaMalloc failed while packing weights");
  for (size_t i = 0; i < wnames.size(); ++i) {
    const float* w = raw_ptr(h, wnames[i], (size_t)n_each * K);
    const float* b = raw_ptr(h, bnames[i], (size_t)n_each);
    if (!w || !b) return SYL_E_STATE;
    const size_t n = (size_t)n_each * K;
    split_f32_kernel<<<grid_for(n), 256>>>(w, out.hi + i * n, out.lo + i * n, n);
    CUDA_TRY(h, cudaMemcpy(out.bias + i * n_each, b, n_each * sizeof(float), cudaMemcpyDeviceToDevice));
  }
  if (!make_weight_maps(h, out, 256)) return SYL_E_CUDA;
  return SYL_OK;
}

float* copy_vec(syl_handle* h, const std::string& name, size_t n) {
  const float* src = raw_ptr(h, name, n);
  if (!src) return nullptr;
  float* dst = dev_alloc<float>(h, n);
  if (!dst) {
    fail(h, SYL_E_CUDA, "cudaMalloc failed");
    return nullptr;
  }
  if (cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice) != cudaSuccess) {
    fail(h, SYL_E_CUDA, "cudaMemcpy failed for %s", name.c_str());
    return nullptr;
  }
  return dst;
}

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
void conv_lengths(int t_samp, int* L) {
  int n = t_samp;
  for (int i = 0; i < 7; ++i) {
    n = (n >= kConvK[i]) ? (n - kConvK[i]) / kConvS[i] + 1 : 0;
    L[i] = n;
  }
}

WsLayout make_layout(int batch, int t_samp) {
  WsLayout w;
  conv_lengths(t_samp, w.L);
  w.T = w.L[6];
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 1023) & ~size_t(1023);
    return o;
  };
  const size_t B = batch, M = B * w.T;
  w.mom = take(B * (size_t)((w.L[0] + MOM_T_PER_BLOCK - 1) / MOM_T_PER_BLOCK) * C0_NMOM * sizeof(double));
  w.gn_scale = take(B * kC * sizeof(float));
  w.gn_shift = take(B * kC * sizeof(float));
  w.valid = take(9 * B * sizeof(int32_t));   // frames, needed rows of conv0..conv6 (valid_frames_kernel), length order
  w.nsq = take(2 * M * sizeof(float));   // squared norms, then their scalar powf (segment.cuh)
  w.seg_scratch = take(B * 6 * (size_t)(w.T + 1) * sizeof(int32_t));
  for (int i = 0; i < 6; ++i) {
    w.act_hi[i] = take(B * w.L[i] * kC * sizeof(__half));
    w.act_lo[i] = take(B * w.L[i] * kC * sizeof(__half));
  }
  w.conv6 = take(M * kC * sizeof(float));
  w.ln_hi = take(M * kC * sizeof(__half));
  w.ln_lo = take(M * kC * sizeof(__half));
  w.h = take(M * kH * sizeof(float));
  w.h16_hi = take(M * kH * sizeof(__half));
  w.h16_lo = take(M * kH * sizeof(__half));
  w.pos = take(M * kH * sizeof(float));
  w.pre = take(M * kH * sizeof(float));
  w.qkv = take(M * 3 * kH * sizeof(__half));
  w.ctx_hi = take(M * kH * sizeof(__half));
  w.ctx_lo = take(M * kH * sizeof(__half));
  w.mid_hi = take(M * kF * sizeof(__half));
  w.mid_lo = take(M * kF * sizeof(__half));
  w.total = off;
  return w;
}

template <typename T>
T* at(void* ws, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(ws) + off);
}

GemmParams base_params() {
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.n_pass = 1;
  p.col_scale = 1.0f;
  p.col_scale_limit = 0;
  return p;
}

// A operand over a [batches][rows][K] fp16 tensor (plain matrices use batches = 1)
bool make_a_maps(syl_handle* h, GemmOp& op, const __half* hi, const __half* lo, int K, int rows, int batches,
                 uint64_t row_stride, uint64_t batch_stride) {
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, (uint64_t)batches};
  uint64_t str[3] = {1, row_stride, batch_stride};
  if (!make_tmap_f16(&op.a_hi, hi, 3, dims, str, 128, &h->err)) return false;
  if (!make_tmap_f16(&op.a_lo, lo ? lo : hi, 3, dims, str, 128, &h->err)) return false;
  op.p.rows_per_batch = rows;
  op.p.batches = batches;
  op.p.kb_per_pass = K / GEMM_BLOCK_K;
  return true;
}

// output maps; any of the three destinations may be null
bool make_o_maps(syl_handle* h, GemmOp& op, float* f32, __half* hi, __half* lo, int ld) {
  const int rows = op.p.rows_per_batch, nb = op.p.batches;
  memset(&op.o3_f32, 0, sizeof(CUtensorMap));
  memset(&op.o3_hi, 0, sizeof(CUtensorMap));
  memset(&op.o3_lo, 0, sizeof(CUtensorMap));
  {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows, (uint64_t)nb};
    uint64_t str[3] = {1, (uint64_t)ld, (uint64_t)rows * ld};
    const bool hi_wide = hi && !lo && !f32;
    if (f32 && !make_tmap(&op.o3_f32, f32, 4, 3, dims, str, 16, 32, 64, &h->err)) return false;
    if (hi && !make_tmap(&op.o3_hi, hi, 2, 3, dims, str, hi_wide ? 32 : 16, 32, hi_wide ? 64 : 0, &h->err)) return false;
    if (lo && !make_tmap(&op.o3_lo, lo, 2, 3, dims, str, 16, 32, 0, &h->err)) return false;
  }
  op.p.out_f32 = f32 != nullptr;
  op.p.out_hi = hi != nullptr;
  op.p.out_lo = lo != nullptr;
  return true;
}

#ifdef SYL_DIAG
long long* g_gemm_trace = nullptr;       // syl_gemm_set_trace: timeline probe of the next GEMM launches (tools/gemm_trace.py)
#endif

int launch_gemm3_raw(const GemmOp& op, const CUtensorMap& b_hi, const CUtensorMap& b_lo, cudaStream_t st, int sm_count) {
#ifdef SYL_DIAG
  GemmParams p = op.p;
  p.trace = g_gemm_trace;
  static const int epi_skip = diag_env("SYL_GEMM_EPI_SKIP", 0);
  p.epi_skip = epi_skip;
#else
  const GemmParams& p = op.p;
#endif
  const int tiles_m = p.batches * ((p.rows_per_batch + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M));
  const int tiles = tiles_m * (p.N / GEMM_BLOCK_N);
  const int clusters = std::min(tiles, sm_count / 2);
  if (clusters <= 0) return SYL_OK;
  launch_pdl(gemm3_tc_kernel, dim3(2 * clusters), dim3(GEMM3_THREADS), GEMM2_SMEM_TOTAL, st, op.a_hi, op.a_lo, b_hi, b_lo, op.o3_f32,
             op.o3_hi, op.o3_lo, p);
  return launch_ok() ? SYL_OK : SYL_E_CUDA;
}

int launch_gemm(syl_handle* h, const GemmOp& op, cudaStream_t st, int sm_count) {
  if (launch_gemm3_raw(op, op.w->map2_hi, op.w->map2_lo, st, sm_count) != SYL_OK)
    return fail(h, SYL_E_CUDA, "gemm launch failed: %s", launch_err());
  return SYL_OK;
}

// single-pass mode pairs taps into N = 96 MMAs (posconv.cuh); SYL_POSCONV_PAIR=0 keeps one tap per MMA (A/B timing)
#ifdef SYL_DIAG
bool posconv_pair_enabled() {
  static const int v = diag_env("SYL_POSCONV_PAIR", 1);
  return v != 0;
}
#else
constexpr bool posconv_pair_enabled() { return true; }
#endif

int launch_posconv(syl_handle* h, const PosOp& op, cudaStream_t st, int sm_count) {
  const bool pair = op.p.n_pass == 1 && posconv_pair_enabled();
  const int tile_t = pair ? PC_TILE_T_PAIR : PC_TILE_T;
  const int t_tiles = (op.p.T + tile_t - 1) / tile_t;
  const int tiles = op.p.batches * t_tiles * PC_GROUPS;
  if (pair)
    launch_pdl(posconv_kernel<true>, dim3(std::min(tiles, sm_count)), dim3(PC_THREADS), PC_SMEM_TOTAL, st, op.a_hi, op.a_lo,
               h->pos.map_hi, h->pos.map_lo, op.o_map, op.o_map31, op.p);
  else
    launch_pdl(posconv_kernel<false>, dim3(std::min(tiles, sm_count)), dim3(PC_THREADS), PC_SMEM_TOTAL, st, op.a_hi, op.a_lo,
               h->pos.map_hi, h->pos.map_lo, op.o_map, op.o_map31, op.p);
  CUDA_TRY(h, cudaGetLastError());
  return SYL_OK;
}

// cudaFuncSetAttribute applies to the CURRENT device's copy of the function, so the >48 KB dynamic shared memory
// opt-in is made once per device (a second handle on cuda:1 in the same process needs its own), under a lock.
std::mutex g_attr_mutex;
uint64_t g_attrs_set_mask = 0;       // bit d: attributes set on device d
int ensure_attrs(syl_handle* h) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  const int d = h ? h->device : 0;
  if (d < 64 && (g_attrs_set_mask >> d) & 1) return SYL_OK;
  CUDA_TRY(h, cudaFuncSetAttribute(gemm3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_TOTAL));
  CUDA_TRY(h, cudaFuncSetAttribute(posconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_TOTAL));
  CUDA_TRY(h, cudaFuncSetAttribute(posconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_TOTAL));
  CUDA_TRY(h, (cudaFuncSetAttribute(attention7_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL)));
  CUDA_TRY(h, (cudaFuncSetAttribute(attention7_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL)));
#ifdef SYL_DIAG
  CUDA_TRY(h, (cudaFuncSetAttribute(attention7_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL)));
#endif
  if (d < 64) g_attrs_set_mask |= uint64_t(1) << d;
  return SYL_OK;
}

int posconv_base_offset_mode() {
  // Measured on B200 (profiles/r01_posconv_base_offset.md): the UMMA swizzle XOR is a function of the absolute
  // shared-memory address, exactly like TMA's, so a descriptor whose start is shifted by whole 128-byte rows needs
  // base_offset = 0.  Setting (addr >> 7) & 7 there produces wrong results.  The env var only exists to re-run
  // that experiment.
#ifdef SYL_DIAG
  return diag_env("SYL_POSCONV_BASE_OFFSET", 0);
#else
  return 0;
#endif
}

// ------------------------------------------------------------------------------------------------
// plan: tensor maps + GEMM parameters for one (batch, t_samp, workspace, hidden) combination
// ------------------------------------------------------------------------------------------------
int build_plan(syl_handle* h, int batch, int t_samp, void* ws, float* hidden) {
  for (int i = 0; i < syl_handle::kPlans; ++i) {
    Plan& c = h->plans[i];
    if (c.valid && c.batch == batch && c.t_samp == t_samp && c.ws == ws) {
      h->plan_cur = i;
      c.last_use = ++h->plan_clock;
      return SYL_OK;
    }
  }
  int victim = 0;
  for (int i = 0; i < syl_handle::kPlans; ++i) {
    if (!h->plans[i].valid) { victim = i; break; }
    if (h->plans[i].last_use < h->plans[victim].last_use) victim = i;
  }
  if (h->trim && batch > GEMM2_MAX_TRIM_BATCHES)
    return fail(h, SYL_E_ARG, "trimmed mode handles at most %d utterances per call, got %d", GEMM2_MAX_TRIM_BATCHES, batch);
  h->plan_cur = victim;
  Plan& pl = h->plans[h->plan_cur];
  pl.valid = false;
  pl.order_valid = false;
  pl.uses = 0;
  pl.last_use = ++h->plan_clock;
  for (auto& g : pl.graphs) cudaGraphExecDestroy(g.exec);
  pl.graphs.clear();
  pl.batch = batch;
  pl.t_samp = t_samp;
  pl.ws = ws;
  pl.hidden = hidden;
  pl.lay = make_layout(batch, t_samp);
  const WsLayout& L = pl.lay;
  const int T = L.T, M = batch * T;
  const bool split_proj = h->mode & SYL_SPLIT_FPROJ, split_pos = h->mode & SYL_SPLIT_POS, split_enc = h->mode & SYL_SPLIT_ENC;

  // conv1..conv6: A = overlapping-row view of the previous channels-last activation
  for (int i = 1; i <= 6; ++i) {
    GemmOp& op = pl.conv[i - 1];
    op.p = base_params();
    op.w = &h->convw[i - 1];
    const int k = kConvK[i], s = kConvS[i], Lin = L.L[i - 1], Lout = L.L[i];
    if (!make_a_maps(h, op, at<__half>(ws, L.act_hi[i - 1]), at<__half>(ws, L.act_lo[i - 1]), k * kC, Lout, batch,
                     (uint64_t)s * kC, (uint64_t)Lin * kC))
      return SYL_E_CUDA;
    op.p.N = kC;
    op.p.n_pass = conv_split(h->mode, i) ? 3 : 1;
    op.p.act = 1;
    if (h->trim) {      // only the rows the valid frames read; the rest of a computed tile is written as zero
      op.p.valid_rows = at<int32_t>(ws, L.valid) + (1 + i) * batch;
      op.p.skip_invalid_tiles = 1;
    }
    // the lo half of this layer's output is only needed if the NEXT conv runs split
    const bool ok = (i < 6) ? make_o_maps(h, op, nullptr, at<__half>(ws, L.act_hi[i]),
                                          conv_split(h->mode, i + 1) ? at<__half>(ws, L.act_lo[i]) : nullptr, kC)
                            : make_o_maps(h, op, at<float>(ws, L.conv6), nullptr, nullptr, kC);
    if (!ok) return SYL_E_CUDA;
  }
  // feature projection: LN(512) output -> 768, bias, zero padded frames (modeling_hubert.py:229, :429-432);
  // tiled per utterance so that valid_rows indexes the batch
  {
    GemmOp& op = pl.proj;
    op.p = base_params();
    op.w = &h->proj;
    if (!make_a_maps(h, op, at<__half>(ws, L.ln_hi), at<__half>(ws, L.ln_lo), kC, T, batch, kC, (uint64_t)T * kC))
      return SYL_E_CUDA;
    op.p.N = kH;
    op.p.n_pass = split_proj ? 3 : 1;
    op.p.bias = h->proj.bias;
    op.p.valid_rows = at<int32_t>(ws, L.valid);
    if (!make_o_maps(h, op, at<float>(ws, L.h), at<__half>(ws, L.h16_hi), split_pos ? at<__half>(ws, L.h16_lo) : nullptr, kH))
      return SYL_E_CUDA;
  }
  // positional conv (posconv.cuh): activation window resident in smem, weights stream per tap
  {
    PosOp& op = pl.pos;
    uint64_t dims[3] = {(uint64_t)kH, (uint64_t)T, (uint64_t)batch};
    uint64_t str[3] = {1, (uint64_t)kH, (uint64_t)T * kH};
    if (!make_tmap_f16(&op.a_hi, at<__half>(ws, L.h16_hi), 3, dims, str, PC_WIN_ROWS / 2, &h->err)) return SYL_E_CUDA;
    if (!make_tmap_f16(&op.a_lo, at<__half>(ws, L.h16_lo), 3, dims, str, PC_WIN_ROWS / 2, &h->err)) return SYL_E_CUDA;
    if (!make_tmap(&op.o_map, at<float>(ws, L.pos), 4, 3, dims, str, 16, 32, 64, &h->err)) return SYL_E_CUDA;
    if (!make_tmap(&op.o_map31, at<float>(ws, L.pos), 4, 3, dims, str, 16, 31, 64, &h->err)) return SYL_E_CUDA;
    op.p.T = T;
    op.p.batches = batch;
    op.p.n_pass = split_pos ? 3 : 1;
    op.p.bias = h->pos.bias;
    op.p.use_base_offset = posconv_base_offset_mode();
  }
  // encoder layers
  const int nl = h->n_layers;
  // encoder GEMMs: one flat [B T, K] matrix, or (trimmed mode) tiled per utterance so that whole tiles of padding drop out
  const int enc_rows = h->trim ? T : M, enc_batches = h->trim ? batch : 1;
  auto enc_trim = [&](GemmOp& op) {
    if (!h->trim) return;
    op.p.valid_rows = at<int32_t>(ws, L.valid);
    op.p.skip_invalid_tiles = 1;
  };
  pl.qkv.assign(nl, GemmOp());
  pl.out.assign(nl, GemmOp());
  pl.ffn1.assign(nl, GemmOp());
  pl.ffn2.assign(nl, GemmOp());
  for (int l = 0; l < nl; ++l) {
    const LayerW& w = h->layers[l];
    {
      GemmOp& op = pl.qkv[l];
      op.p = base_params();
      op.w = &w.qkv;
      if (!make_a_maps(h, op, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo), kH, enc_rows, enc_batches, kH, (uint64_t)enc_rows * kH)) return SYL_E_CUDA;
      enc_trim(op);
      op.p.N = 3 * kH;
      op.p.n_pass = split_enc ? 3 : 1;
      op.p.bias = w.qkv.bias;
      op.p.col_scale = 0.125f;        // head_dim ** -0.5, exact (modeling_hubert.py:287)
      op.p.col_scale_limit = kH;
      if (!make_o_maps(h, op, nullptr, at<__half>(ws, L.qkv), nullptr, 3 * kH)) return SYL_E_CUDA;
    }
    {
      GemmOp& op = pl.out[l];
      op.p = base_params();
      op.w = &w.out;
      if (!make_a_maps(h, op, at<__half>(ws, L.ctx_hi), at<__half>(ws, L.ctx_lo), kH, enc_rows, enc_batches, kH, (uint64_t)enc_rows * kH)) return SYL_E_CUDA;
      enc_trim(op);
      op.p.N = kH;
      op.p.n_pass = split_enc ? 3 : 1;
      op.p.bias = w.out.bias;
      if (!make_o_maps(h, op, at<float>(ws, L.pre), nullptr, nullptr, kH)) return SYL_E_CUDA;
    }
    {
      GemmOp& op = pl.ffn1[l];
      op.p = base_params();
      op.w = &w.ffn1;
      if (!make_a_maps(h, op, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo), kH, enc_rows, enc_batches, kH, (uint64_t)enc_rows * kH)) return SYL_E_CUDA;
      enc_trim(op);
      op.p.N = kF;
      op.p.n_pass = split_enc ? 3 : 1;
      op.p.bias = w.ffn1.bias;
      op.p.act = 1;
      if (!make_o_maps(h, op, nullptr, at<__half>(ws, L.mid_hi), split_enc ? at<__half>(ws, L.mid_lo) : nullptr, kF)) return SYL_E_CUDA;
    }
    {
      GemmOp& op = pl.ffn2[l];
      op.p = base_params();
      op.w = &w.ffn2;
      if (!make_a_maps(h, op, at<__half>(ws, L.mid_hi), at<__half>(ws, L.mid_lo), kF, enc_rows, enc_batches, kF, (uint64_t)enc_rows * kF)) return SYL_E_CUDA;
      enc_trim(op);
      op.p.N = kH;
      op.p.n_pass = split_enc ? 3 : 1;
      op.p.bias = w.ffn2.bias;
      if (!make_o_maps(h, op, at<float>(ws, L.pre), nullptr, nullptr, kH)) return SYL_E_CUDA;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)3 * kH, (uint64_t)T, (uint64_t)batch};
    uint64_t str[3] = {1, (uint64_t)3 * kH, (uint64_t)T * 3 * kH};
    if (!make_tmap_f16(&pl.attn_map, at<__half>(ws, L.qkv), 3, dims, str, 128, &h->err)) return SYL_E_CUDA;
    uint64_t od[3] = {(uint64_t)kH, (uint64_t)T, (uint64_t)batch};
    uint64_t os[3] = {1, (uint64_t)kH, (uint64_t)T * kH};
    if (!make_tmap(&pl.ctx_hi_map, at<__half>(ws, L.ctx_hi), 2, 3, od, os, 64, 32, 128, &h->err)) return SYL_E_CUDA;
    if (!make_tmap(&pl.ctx_lo_map, at<__half>(ws, L.ctx_lo), 2, 3, od, os, 64, 32, 128, &h->err)) return SYL_E_CUDA;
  }
  pl.valid = true;
  return SYL_OK;
}

template <int D>
void launch_ln(const float* x, const float* add, const __half* add_hi, const __half* add_lo, const float* g, const float* b, int rows,
               float* of, __half* ohi, __half* olo, cudaStream_t st, const int32_t* valid = nullptr, int T = 1) {
  constexpr int warps = 8;   // rows per block; 2 and 4 measured equal (profiles/r03_variants_ab.md)
  launch_pdl(layernorm_rows_kernel<D>, dim3((rows + warps - 1) / warps), dim3(warps * 32), 0, st, x, add, add_hi, add_lo, g, b, rows, of,
             ohi, olo, valid, T);
}

long long* g_attn_trace = nullptr;   // set by syl_attention_trace for one launch
int g_attn_trace_cap = 0;

int launch_attention(const CUtensorMap& qkv, const CUtensorMap& o_hi, const CUtensorMap& o_lo, const int32_t* kv_len,
                     int B, int T, int out_lo, int sm_count, cudaStream_t st, int trim = 0, const int32_t* order = nullptr) {
  AttnParams ap;
  ap.trace = g_attn_trace;
  ap.trace_cap = g_attn_trace_cap;
  ap.T = T;
  ap.batches = B;
  ap.heads = kHeads;
  ap.model_dim = kH;
  ap.kv_len = kv_len;
  ap.out_lo = out_lo;
  ap.trim = trim && kv_len != nullptr;
  ap.order = order;
  // experiment switches (profiles/r02_attention.md): SYL_ATTN_POLY = exp2 pairs (out of every four) computed on the
  // FMA pipe (0 or 1), SYL_ATTN_DEBUG = arithmetic-removal probes
#ifdef SYL_DIAG
  static const int debug = diag_env("SYL_ATTN_DEBUG", 0), poly = diag_env("SYL_ATTN_POLY", 1);
#else
  constexpr int debug = 0, poly = 1;
#endif
  ap.debug = debug;
  const int q_tiles = (T + ATT_BQ - 1) / ATT_BQ;
  const int items = B * kHeads * ((q_tiles + ATT_QT - 1) / ATT_QT);
  const int grid = std::min(items, sm_count);
#ifdef SYL_DIAG
  if (ap.trace) {
    launch_pdl(attention7_kernel<0, true>, dim3(grid), dim3(ATT7_THREADS), ATT7_SMEM_TOTAL, st, qkv, o_hi, o_lo, ap);
    return launch_ok() ? SYL_OK : SYL_E_CUDA;
  }
#endif
  if (poly == 1)
    launch_pdl(attention7_kernel<1, false>, dim3(grid), dim3(ATT7_THREADS), ATT7_SMEM_TOTAL, st, qkv, o_hi, o_lo, ap);
  else
    launch_pdl(attention7_kernel<0, false>, dim3(grid), dim3(ATT7_THREADS), ATT7_SMEM_TOTAL, st, qkv, o_hi, o_lo, ap);
  return launch_ok() ? SYL_OK : SYL_E_CUDA;
}

int run_frontend(syl_handle* h, const float* wav, cudaStream_t st) {
  Plan& pl = h->plans[h->plan_cur];
  const WsLayout& L = pl.lay;
  void* ws = pl.ws;
  const int B = pl.batch, L0 = L.L[0];
  {
    StageTimer tm(h, ST_CONV0, st);
    const int chunks = (L0 + MOM_T_PER_BLOCK - 1) / MOM_T_PER_BLOCK;
    launch_pdl(conv0_moments_kernel, dim3(chunks, B), dim3(MOM_THREADS), 0, st, wav, pl.t_samp, L0, at<double>(ws, L.mom));
    launch_pdl(conv0_gn_coeff_kernel, dim3(B), dim3(kC), 0, st, at<double>(ws, L.mom), chunks, h->conv0_w, h->gn_g, h->gn_b, L0,
               at<float>(ws, L.gn_scale), at<float>(ws, L.gn_shift));
    __half* hi = at<__half>(ws, L.act_hi[0]);
    __half* lo = conv_split(h->mode, 1) ? at<__half>(ws, L.act_lo[0]) : nullptr;
    const int32_t* needed0 = h->trim ? at<int32_t>(ws, L.valid) + B : nullptr;   // rows of conv0 the valid frames read
    if (lo)
      launch_pdl(conv0_mma_kernel<true>, dim3((L0 + C0M_T - 1) / C0M_T, B), dim3(C0M_THREADS), 0, st, wav, pl.t_samp, L0,
                 h->conv0_bfrag, at<float>(ws, L.gn_scale), at<float>(ws, L.gn_shift), hi, lo, needed0);
    else
      launch_pdl(conv0_mma_kernel<false>, dim3((L0 + C0M_T - 1) / C0M_T, B), dim3(C0M_THREADS), 0, st, wav, pl.t_samp, L0,
                 h->conv0_bfrag, at<float>(ws, L.gn_scale), at<float>(ws, L.gn_shift), hi, lo, needed0);
  }
  CUDA_TRY(h, cudaGetLastError());
  {
    StageTimer tm(h, ST_CONV1, st);    // conv1 alone: half of the conv stack's FLOPs, the launch bench.py's roofline quotes
    int rc = launch_gemm(h, pl.conv[0], st, h->sm_count);
    if (rc) return rc;
  }
  StageTimer tm(h, ST_CONV, st);
  for (int i = 1; i < 6; ++i) {
    int rc = launch_gemm(h, pl.conv[i], st, h->sm_count);
    if (rc) return rc;
  }
  return SYL_OK;
}

// one post-LN encoder layer (modeling_hubert.py:388-405); residual adds are fused into the LayerNorm kernels
int run_layer(syl_handle* h, int l, float* h_out, cudaStream_t st) {
  Plan& pl = h->plans[h->plan_cur];
  const WsLayout& L = pl.lay;
  void* ws = pl.ws;
  const int B = pl.batch, T = L.T, M = B * T;
  const LayerW& w = h->layers[l];
  const bool split_enc = h->mode & SYL_SPLIT_ENC;
  int rc;
  {
    StageTimer tm(h, ST_QKV, st);
    if ((rc = launch_gemm(h, pl.qkv[l], st, h->sm_count))) return rc;
  }
  {
    StageTimer tm(h, ST_ATTN, st);
    if (launch_attention(pl.attn_map, pl.ctx_hi_map, pl.ctx_lo_map, at<int32_t>(ws, L.valid), B, T, split_enc, h->sm_count, st, h->trim,
                         h->trim && pl.order_valid ? at<int32_t>(ws, L.valid) + 8 * B : nullptr))
      return fail(h, SYL_E_CUDA, "attention launch failed: %s", launch_err());
  }
  {
    StageTimer tm(h, ST_OUT, st);
    if ((rc = launch_gemm(h, pl.out[l], st, h->sm_count))) return rc;
  }
  {
    StageTimer tm(h, ST_LN_ENC, st);   // h = LN(h + attn); the residual stream is the fp16 pair (h16_hi, h16_lo), in place
    launch_ln<kH>(at<float>(ws, L.pre), nullptr, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo),
                  w.ln1_g, w.ln1_b, M, nullptr, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo), st,
                  h->trim ? at<int32_t>(ws, L.valid) : nullptr, T);
  }
  {
    StageTimer tm(h, ST_FFN1, st);
    if ((rc = launch_gemm(h, pl.ffn1[l], st, h->sm_count))) return rc;
  }
  {
    StageTimer tm(h, ST_FFN2, st);
    if ((rc = launch_gemm(h, pl.ffn2[l], st, h->sm_count))) return rc;
  }
  {
    StageTimer tm(h, ST_LN_ENC, st);   // h = LN(h + ffn); fp32 only where the caller wants the layer output (h_out)
    launch_ln<kH>(at<float>(ws, L.pre), nullptr, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo),
                  w.ln2_g, w.ln2_b, M, h_out, at<__half>(ws, L.h16_hi), at<__half>(ws, L.h16_lo), st,
                  h->trim ? at<int32_t>(ws, L.valid) : nullptr, T);
  }
  CUDA_TRY(h, cudaGetLastError());
  return SYL_OK;
}

int run_segment(const float* states, int B, int T, float thr_norm, float thr_merge, int32_t* seg, int32_t* seg_count,
                float* seg_feat, int max_seg, float* nsq, int32_t* scratch, cudaStream_t st) {
  const int rows = B * T;
  float* pw = nsq + rows;      // powf(nsq, .5f) per frame, right behind the squared norms
  launch_pdl(frame_sqnorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, states, rows, T, thr_norm, nsq, pw, scratch);
  launch_pdl(segment_kernel, dim3((T + SEG_CHUNK - 1) / SEG_CHUNK, B), dim3(SEG_SCAN_THREADS), 0, st, states, pw, T, thr_merge, seg,
             seg_count, max_seg, scratch);
  if (seg_feat) launch_pdl(segment_pool_kernel, dim3(std::min(max_seg, SEG_POOL_GRID), B), dim3(192), 0, st, states, T, seg, seg_count, max_seg, seg_feat);
  return launch_ok() ? SYL_OK : SYL_E_CUDA;
}

}  // namespace

#ifdef SYL_DIAG
template <int N>
static int run_mma_probe(int iters, int ctas, long long* out_dev, cudaStream_t st) {
  if (cudaFuncSetAttribute(mma_probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024 + 64) != cudaSuccess)
    return SYL_E_CUDA;
  mma_probe_kernel<N><<<ctas, 128, 65536 + 1024 + 64, st>>>(iters, out_dev);
  return launch_ok() ? SYL_OK : SYL_E_CUDA;
}
#endif

// ================================================================================================
// exported C ABI
// ================================================================================================
extern "C" {

const char* syl_last_error(const syl_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int syl_num_frames(int n_samples) {
  int L[7];
  conv_lengths(n_samples, L);
  return L[6];
}

int syl_create(syl_handle** out, int device, int n_layers, int mode) {
  if (!out || n_layers < 1 || n_layers > 48) return fail(nullptr, SYL_E_ARG, "syl_create: bad arguments");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return fail(nullptr, SYL_E_CUDA, "syl_create: CUDA device %d not available (%d devices); there is no CPU fallback",
                device, count);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, SYL_E_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, SYL_E_CUDA, "syl_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  syl_handle* h = new syl_handle();
  h->device = device;
  h->n_layers = n_layers;
  if (mode & SYL_SPLIT_CONV) mode |= SYL_SPLIT_CONV2 | SYL_SPLIT_CONV3 | SYL_SPLIT_CONV4 | SYL_SPLIT_CONV5 | SYL_SPLIT_CONV6;
  if (mode & SYL_SPLIT_PROJ) mode |= SYL_SPLIT_FPROJ | SYL_SPLIT_POS;
  h->trim = (mode & SYL_TRIM_PADDING) != 0;
  h->mode = mode & ~SYL_TRIM_PADDING;
  h->sm_count = prop.multiProcessorCount;
  *out = h;
  return SYL_OK;
}

void syl_destroy(syl_handle* h) {
  if (!h) return;
  {
  DeviceGuard dg(h->device);
  for (auto& kv : h->raw) cudaFree(kv.second.first);
  for (void* p : h->owned) cudaFree(p);
  for (cudaEvent_t e : h->pool) cudaEventDestroy(e);
  for (Plan& pl : h->plans)
    for (auto& g : pl.graphs) cudaGraphExecDestroy(g.exec);
  (void)cudaGetLastError();   // teardown never leaves a pending error for the next runtime user of this thread
  }
  delete h;
}

int syl_load_weight(syl_handle* h, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype) {
  if (!h || !name || !dev_ptr || !shape || ndim < 1 || ndim > 4) return fail(h, SYL_E_ARG, "syl_load_weight: bad arguments");
  std::lock_guard<std::mutex> lock(h->mu);
  if (dtype != SYL_DTYPE_F32) return fail(h, SYL_E_ARG, "syl_load_weight: only fp32 tensors are accepted");
  if (h->finalized) return fail(h, SYL_E_STATE, "syl_load_weight after syl_finalize");
  DeviceGuard dg(h->device);
  if (dg.err != cudaSuccess) return fail(h, SYL_E_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(dg.err));
  size_t n = 1;
  std::vector<int64_t> shp(shape, shape + ndim);
  for (int64_t d : shp) n *= (size_t)d;
  std::string key(name);
  // legacy weight-norm naming
  const std::string pre = "encoder.pos_conv_embed.conv.";
  if (key == pre + "weight_g") key = pre + "parametrizations.weight.original0";
  if (key == pre + "weight_v") key = pre + "parametrizations.weight.original1";
  float* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, n * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(p, dev_ptr, n * sizeof(float), cudaMemcpyDeviceToDevice));
  auto it = h->raw.find(key);
  if (it != h->raw.end()) cudaFree(it->second.first);
  h->raw[key] = std::make_pair(p, shp);
  return SYL_OK;
}

int syl_finalize(syl_handle* h) {
  SYL_ENTER();
  if (!h) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->finalized) return SYL_OK;
  DeviceGuard dg(h->device);
  if (dg.err != cudaSuccess) return fail(h, SYL_E_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(dg.err));
  int rc = ensure_attrs(h);
  if (rc) return rc;
  const std::string fe = "feature_extractor.conv_layers.";
  if (!(h->conv0_w = copy_vec(h, fe + "0.conv.weight", (size_t)kC * 10))) return SYL_E_STATE;
  if (!(h->conv0_bfrag = dev_alloc<uint4>(h, 64 * 32))) return fail(h, SYL_E_CUDA, "cudaMalloc failed");
  conv0_bfrag_kernel<<<8, 256>>>(h->conv0_w, h->conv0_bfrag);
  if (!(h->gn_g = copy_vec(h, fe + "0.layer_norm.weight", kC))) return SYL_E_STATE;
  if (!(h->gn_b = copy_vec(h, fe + "0.layer_norm.bias", kC))) return SYL_E_STATE;
  for (int i = 1; i <= 6; ++i) {
    const int k = kConvK[i];
    const float* w = raw_ptr(h, fe + std::to_string(i) + ".conv.weight", (size_t)kC * kC * k);
    if (!w) return SYL_E_STATE;
    PackedLinear& pw = h->convw[i - 1];
    pw.N = kC;
    pw.K = k * kC;
    pw.hi = dev_alloc<__half>(h, (size_t)kC * kC * k);
    pw.lo = dev_alloc<__half>(h, (size_t)kC * kC * k);
    if (!pw.hi || !pw.lo) return fail(h, SYL_E_CUDA, "cudaMalloc failed");
    pack_conv_w_kernel<<<grid_for((size_t)kC * kC * k), 256>>>(w, k, pw.hi, pw.lo);
    if (!make_weight_maps(h, pw, 256)) return SYL_E_CUDA;
  }
  if (!(h->fp_ln_g = copy_vec(h, "feature_projection.layer_norm.weight", kC))) return SYL_E_STATE;
  if (!(h->fp_ln_b = copy_vec(h, "feature_projection.layer_norm.bias", kC))) return SYL_E_STATE;
  if ((rc = pack_linear(h, h->proj, {"feature_projection.projection.weight"}, {"feature_projection.projection.bias"}, kH, kC)))
    return rc;
  {
    const std::string pc = "encoder.pos_conv_embed.conv.";
    const float* g = raw_ptr(h, pc + "parametrizations.weight.original0", kPosK);
    const float* v = raw_ptr(h, pc + "parametrizations.weight.original1", (size_t)kH * kPosCg * kPosK);
    if (!g || !v) return SYL_E_STATE;
    float* norm = dev_alloc<float>(h, kPosK);
    PackedLinear& pw = h->pos;
    pw.N = kH;
    pw.K = kPosK * 64;
    pw.hi = dev_alloc<__half>(h, (size_t)kH * kPosK * 64);
    pw.lo = dev_alloc<__half>(h, (size_t)kH * kPosK * 64);
    if (!norm || !pw.hi || !pw.lo) return fail(h, SYL_E_CUDA, "cudaMalloc failed");
    if (!(pw.bias = copy_vec(h, pc + "bias", kH))) return SYL_E_STATE;
    pos_tap_norm_kernel<<<kPosK, 256>>>(v, norm);
    pack_pos_w_kernel<<<grid_for((size_t)kH * kPosK * 64), 256>>>(v, g, norm, pw.hi, pw.lo);
    if (!make_weight_maps(h, pw, 48)) return SYL_E_CUDA;
  }
  if (!(h->enc_ln_g = copy_vec(h, "encoder.layer_norm.weight", kH))) return SYL_E_STATE;
  if (!(h->enc_ln_b = copy_vec(h, "encoder.layer_norm.bias", kH))) return SYL_E_STATE;
  h->layers.resize(h->n_layers);
  for (int l = 0; l < h->n_layers; ++l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    LayerW& w = h->layers[l];
    if ((rc = pack_linear(h, w.qkv, {p + "attention.q_proj.weight", p + "attention.k_proj.weight", p + "attention.v_proj.weight"},
                          {p + "attention.q_proj.bias", p + "attention.k_proj.bias", p + "attention.v_proj.bias"}, kH, kH)))
      return rc;
    if ((rc = pack_linear(h, w.out, {p + "attention.out_proj.weight"}, {p + "attention.out_proj.bias"}, kH, kH))) return rc;
    if ((rc = pack_linear(h, w.ffn1, {p + "feed_forward.intermediate_dense.weight"}, {p + "feed_forward.intermediate_dense.bias"}, kF, kH)))
      return rc;
    if ((rc = pack_linear(h, w.ffn2, {p + "feed_forward.output_dense.weight"}, {p + "feed_forward.output_dense.bias"}, kH, kF)))
      return rc;
    if (!(w.ln1_g = copy_vec(h, p + "layer_norm.weight", kH))) return SYL_E_STATE;
    if (!(w.ln1_b = copy_vec(h, p + "layer_norm.bias", kH))) return SYL_E_STATE;
    if (!(w.ln2_g = copy_vec(h, p + "final_layer_norm.weight", kH))) return SYL_E_STATE;
    if (!(w.ln2_b = copy_vec(h, p + "final_layer_norm.bias", kH))) return SYL_E_STATE;
  }
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaGetLastError());
  for (auto& kv : h->raw) cudaFree(kv.second.first);
  h->raw.clear();
  h->finalized = true;
  return SYL_OK;
}

size_t syl_workspace_bytes(const syl_handle* h, int batch, int t_samp_max) {
  (void)h;
  if (batch <= 0 || t_samp_max < 400) return 0;
  return make_layout(batch, t_samp_max).total;
}

int syl_set_active_layers(syl_handle* h, int n) {
  if (!h) return SYL_E_ARG;
  h->active_layers = n;
  return SYL_OK;
}

int syl_forward_launch_count(const syl_handle* h, int with_segmentation) {
  if (!h) return 0;
  const int nl = (h->active_layers >= 0 && h->active_layers < h->n_layers) ? h->active_layers : h->n_layers;
  // valid_frames, moments, gn_coeff, conv0_mma, 6 conv GEMMs, LN512, proj, pos, LN ; per layer 4 GEMM + attn + 2 LN
  return 4 + 6 + 4 + nl * 7 + (with_segmentation ? 3 : 0) + (h->trim ? 1 : 0);   // + length_order_kernel in trimmed mode
}

// enqueue every kernel of one forward on `st` (eagerly, or into a stream capture)
static int enqueue_forward(syl_handle* h, Plan& pl, const float* wav, const int32_t* n_samples, float* hidden,
                           int32_t* seg, int32_t* seg_count, float* seg_feat, int max_seg, float thr_norm,
                           float thr_merge, cudaStream_t st) {
  const WsLayout& L = pl.lay;
  void* workspace = pl.ws;
  const int batch = pl.batch;
  const int T = L.T, M = batch * T;
  const bool split_proj = h->mode & SYL_SPLIT_FPROJ, split_enc = h->mode & SYL_SPLIT_ENC;
  int rc;
  valid_frames_kernel<<<(batch + 127) / 128, 128, 0, st>>>(n_samples, batch, T, at<int32_t>(workspace, L.valid));
  pl.order_valid = h->trim;
  if (h->trim)
    length_order_kernel<<<1, 128, 0, st>>>(at<int32_t>(workspace, L.valid), batch, at<int32_t>(workspace, L.valid) + 8 * batch);
  if ((rc = run_frontend(h, wav, st))) return rc;
  {
    StageTimer tm(h, ST_LN, st);
    launch_ln<kC>(at<float>(workspace, L.conv6), nullptr, nullptr, nullptr, h->fp_ln_g, h->fp_ln_b, M, nullptr, at<__half>(workspace, L.ln_hi),
                  split_proj ? at<__half>(workspace, L.ln_lo) : nullptr, st, h->trim ? at<int32_t>(workspace, L.valid) : nullptr, T);
  }
  {
    StageTimer tm(h, ST_PROJ, st);
    if ((rc = launch_gemm(h, pl.proj, st, h->sm_count))) return rc;
  }
  {
    StageTimer tm(h, ST_POS, st);
    if ((rc = launch_posconv(h, pl.pos, st, h->sm_count))) return rc;
  }
  const int nl = (h->active_layers >= 0 && h->active_layers < h->n_layers) ? h->active_layers : h->n_layers;
  // h = LN(h + pos)   (modeling_hubert.py:441-442); with zero layers this is already the output
  {
    StageTimer tm(h, ST_LN, st);
    launch_ln<kH>(at<float>(workspace, L.h), at<float>(workspace, L.pos), nullptr, nullptr, h->enc_ln_g, h->enc_ln_b, M,
                  nl == 0 ? hidden : nullptr, at<__half>(workspace, L.h16_hi), at<__half>(workspace, L.h16_lo), st,
                  h->trim ? at<int32_t>(workspace, L.valid) : nullptr, T);
  }
  CUDA_TRY(h, cudaGetLastError());
  for (int l = 0; l < nl; ++l) {
    // between layers the residual stream exists only as the fp16 pair (h16_hi, h16_lo); the last layer also writes fp32
    float* out = (l == nl - 1) ? hidden : nullptr;
    if ((rc = run_layer(h, l, out, st))) return rc;
  }
  if (seg) {
    StageTimer tm(h, ST_SEG, st);
    rc = run_segment(hidden, batch, T, thr_norm, thr_merge, seg, seg_count, seg_feat, max_seg,
                     at<float>(workspace, L.nsq), at<int32_t>(workspace, L.seg_scratch), st);
    if (rc) return fail(h, rc, "segmentation launch failed: %s", launch_err());
  }
  return SYL_OK;
}

int syl_forward(syl_handle* h, const float* wav, const int32_t* n_samples, int batch, int t_samp_max, float* hidden,
                int32_t* seg, int32_t* seg_count, float* seg_feat, int max_seg, float thr_norm, float thr_merge,
                void* workspace, size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!h) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->finalized) return fail(h, SYL_E_STATE, "syl_forward before syl_finalize");
  if (!wav || !hidden || !workspace || batch <= 0) return fail(h, SYL_E_ARG, "syl_forward: null pointer or empty batch");
  if (t_samp_max < 400) return fail(h, SYL_E_ARG, "syl_forward: need at least 400 samples (one frame), got %d", t_samp_max);
  const size_t need = syl_workspace_bytes(h, batch, t_samp_max);
  if (workspace_bytes < need) return fail(h, SYL_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return fail(h, SYL_E_ARG, "workspace must be 1024-byte aligned");
  if (seg && (!seg_count || max_seg <= 0)) return fail(h, SYL_E_ARG, "syl_forward: seg needs seg_count and max_seg > 0");
  DeviceGuard dg(h->device);
  if (dg.err != cudaSuccess) return fail(h, SYL_E_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(dg.err));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = build_plan(h, batch, t_samp_max, workspace, hidden);
  if (rc) return rc;
  Plan& pl = h->plans[h->plan_cur];

  // The forward is ~80 launches of fixed shape; replaying it as one CUDA graph removes the per-launch host cost
  // (2.8 ms of CPU per step, measured) and the launch gaps between the short kernels.  Graphs are keyed by the
  // full argument set; per-stage profiling needs real events, so it uses the eager path.
  // (the legacy default stream cannot be captured)
  // a shape seen for the first time runs eagerly: capturing + instantiating a graph costs more than one eager forward,
  // and callers with ever-changing shapes (mixed-length lists, length buckets) would pay it on every call
  const bool graphs = h->use_graphs && !h->profile && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
                      pl.uses++ > 0;
  if (graphs) {
    for (const Plan::GraphEntry& g : pl.graphs) {
      if (g.key[0] == wav && g.key[1] == n_samples && g.key[2] == hidden && g.key[3] == seg && g.key[4] == seg_count &&
          g.key[5] == seg_feat && g.max_seg == max_seg && g.thr_norm == thr_norm && g.thr_merge == thr_merge &&
          g.active_layers == h->active_layers) {
        CUDA_TRY(h, cudaGraphLaunch(g.exec, st));
        return SYL_OK;
      }
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
    const bool began = cs == cudaStreamCaptureStatusNone && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (!began) cudaGetLastError();   // not capturable: clear the error, launch eagerly below
    if (began) {
      rc = enqueue_forward(h, pl, wav, n_samples, hidden, seg, seg_count, seg_feat, max_seg, thr_norm, thr_merge, st);
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc == SYL_OK && ce == cudaSuccess && graph) {
        cudaGraphExec_t exec = nullptr;
        if (cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          cudaGraphDestroy(graph);
          if (pl.graphs.size() >= 8) {
            cudaGraphExecDestroy(pl.graphs.front().exec);
            pl.graphs.erase(pl.graphs.begin());
          }
          Plan::GraphEntry e{{wav, n_samples, hidden, seg, seg_count, seg_feat}, max_seg, h->active_layers, thr_norm, thr_merge, exec};
          pl.graphs.push_back(e);
          CUDA_TRY(h, cudaGraphLaunch(exec, st));
          return SYL_OK;
        }
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();   // capture failed: clear the error and fall through to the eager path
      if (rc != SYL_OK) return rc;
    }
  }
  return enqueue_forward(h, pl, wav, n_samples, hidden, seg, seg_count, seg_feat, max_seg, thr_norm, thr_merge, st);
}

int syl_set_graph_mode(syl_handle* h, int on) {
  if (!h) return SYL_E_ARG;
  h->use_graphs = on != 0;
  return SYL_OK;
}

int syl_conv_frontend(syl_handle* h, const float* wav, int batch, int t_samp_max, float* feats, void* workspace,
                      size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!h) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->finalized) return fail(h, SYL_E_STATE, "syl_conv_frontend before syl_finalize");
  if (!wav || !feats || !workspace || batch <= 0 || t_samp_max < 400) return fail(h, SYL_E_ARG, "syl_conv_frontend: bad arguments");
  const size_t need = syl_workspace_bytes(h, batch, t_samp_max);
  if (workspace_bytes < need) return fail(h, SYL_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  DeviceGuard dg(h->device);
  if (dg.err != cudaSuccess) return fail(h, SYL_E_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(dg.err));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = build_plan(h, batch, t_samp_max, workspace, nullptr);
  if (rc) return rc;
  {
    const WsLayout& L0 = h->plans[h->plan_cur].lay;
    valid_frames_kernel<<<(batch + 127) / 128, 128, 0, st>>>(nullptr, batch, L0.T, at<int32_t>(workspace, L0.valid));
  }
  if ((rc = run_frontend(h, wav, st))) return rc;
  const WsLayout& L = h->plans[h->plan_cur].lay;
  CUDA_TRY(h, cudaMemcpyAsync(feats, at<float>(workspace, L.conv6), (size_t)batch * L.T * kC * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  return SYL_OK;
}

int syl_encoder_layer(syl_handle* h, int layer, const float* h_in, const int32_t* valid_frames, int batch, int T,
                      float* h_out, void* workspace, size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!h) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->finalized) return fail(h, SYL_E_STATE, "syl_encoder_layer before syl_finalize");
  if (!h_in || !h_out || !workspace || batch <= 0 || T <= 0 || layer < 0 || layer >= h->n_layers)
    return fail(h, SYL_E_ARG, "syl_encoder_layer: bad arguments");
  // smallest sample count that yields exactly T frames
  int n = T;
  for (int i = 6; i >= 0; --i) n = (n - 1) * kConvS[i] + kConvK[i];
  const size_t need = syl_workspace_bytes(h, batch, n);
  if (workspace_bytes < need) return fail(h, SYL_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  DeviceGuard dg(h->device);
  if (dg.err != cudaSuccess) return fail(h, SYL_E_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(dg.err));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = build_plan(h, batch, n, workspace, nullptr);
  if (rc) return rc;
  const WsLayout& L = h->plans[h->plan_cur].lay;
  const size_t M = (size_t)batch * T;
  if (valid_frames)
    CUDA_TRY(h, cudaMemcpyAsync(at<int32_t>(workspace, L.valid), valid_frames, batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  else
    fill_i32_kernel<<<(batch + 127) / 128, 128, 0, st>>>(at<int32_t>(workspace, L.valid), batch, T);
  CUDA_TRY(h, cudaMemcpyAsync(at<float>(workspace, L.h), h_in, M * kH * sizeof(float), cudaMemcpyDeviceToDevice, st));
  split_f32_kernel<<<grid_for(M * kH), 256, 0, st>>>(h_in, at<__half>(workspace, L.h16_hi), at<__half>(workspace, L.h16_lo), M * kH);
  return run_layer(h, layer, h_out, st);
}

int syl_attention(const void* qkv_f16, const int32_t* kv_len, int batch, int T, void* out_f16, void* stream) {
  SYL_ENTER();
  std::string err;
  if (!qkv_f16 || !out_f16 || batch <= 0 || T <= 0) return fail(nullptr, SYL_E_ARG, "syl_attention: bad arguments");
  if (cudaFuncSetAttribute(attention7_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL) != cudaSuccess ||
      cudaFuncSetAttribute(attention7_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL) != cudaSuccess
#ifdef SYL_DIAG
      || cudaFuncSetAttribute(attention7_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT7_SMEM_TOTAL) != cudaSuccess
#endif
  )
    return fail(nullptr, SYL_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
  CUtensorMap map, omap;
  uint64_t dims[3] = {(uint64_t)3 * kH, (uint64_t)T, (uint64_t)batch};
  uint64_t str[3] = {1, (uint64_t)3 * kH, (uint64_t)T * 3 * kH};
  uint64_t od[3] = {(uint64_t)kH, (uint64_t)T, (uint64_t)batch};
  uint64_t os[3] = {1, (uint64_t)kH, (uint64_t)T * kH};
  if (!make_tmap_f16(&map, qkv_f16, 3, dims, str, 128, &err) || !make_tmap(&omap, out_f16, 2, 3, od, os, 64, 32, 128, &err))
    return fail(nullptr, SYL_E_CUDA, "%s", err.c_str());
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (launch_attention(map, omap, omap, kv_len, batch, T, 0, sms, reinterpret_cast<cudaStream_t>(stream)) != SYL_OK)
    return fail(nullptr, SYL_E_CUDA, "attention launch failed: %s", launch_err());
  return SYL_OK;
}

#ifdef SYL_DIAG
int syl_attention_trace(const void* qkv_f16, const int32_t* kv_len, int batch, int T, void* out_f16, void* trace_dev,
                        int trace_cap, void* stream) {
  if (!trace_dev || trace_cap <= 0) return fail(nullptr, SYL_E_ARG, "syl_attention_trace: bad arguments");
  g_attn_trace = reinterpret_cast<long long*>(trace_dev);
  g_attn_trace_cap = trace_cap;
  const int rc = syl_attention(qkv_f16, kv_len, batch, T, out_f16, stream);
  g_attn_trace = nullptr;
  g_attn_trace_cap = 0;
  return rc;
}
#endif

size_t syl_pcm16_workspace_bytes(int batch, int t_samp_max) {
  if (batch <= 0 || t_samp_max <= 0) return 0;
  return (size_t)batch * ((t_samp_max + PCM_CHUNK - 1) / PCM_CHUNK) * 2 * sizeof(double);
}

int syl_prepare_pcm16(const int16_t* pcm, const int64_t* offsets, const int32_t* n_samples, int batch, int t_samp_max,
                      int normalize, float* wav_out, void* workspace, size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!pcm || !offsets || !n_samples || !wav_out || !workspace || batch <= 0 || t_samp_max <= 0)
    return fail(nullptr, SYL_E_ARG, "syl_prepare_pcm16: bad arguments");
  if (workspace_bytes < syl_pcm16_workspace_bytes(batch, t_samp_max))
    return fail(nullptr, SYL_E_WORKSPACE, "syl_prepare_pcm16: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int chunks = (t_samp_max + PCM_CHUNK - 1) / PCM_CHUNK;
  double* part = reinterpret_cast<double*>(workspace);
  if (normalize) pcm16_stats_kernel<int16_t><<<dim3(chunks, batch), PCM_THREADS, 0, st>>>(pcm, offsets, n_samples, chunks, part);
  pcm16_apply_kernel<int16_t><<<dim3(chunks, batch), PCM_THREADS, 0, st>>>(pcm, offsets, n_samples, chunks, part, normalize, t_samp_max, wav_out);
  return launch_ok() ? SYL_OK : fail(nullptr, SYL_E_CUDA, "syl_prepare_pcm16 launch failed: %s", launch_err());
}

int syl_prepare_f32(const float* wav, const int64_t* offsets, const int32_t* n_samples, int batch, int t_samp_max,
                    int normalize, float* wav_out, void* workspace, size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!wav || !offsets || !n_samples || !wav_out || !workspace || batch <= 0 || t_samp_max <= 0)
    return fail(nullptr, SYL_E_ARG, "syl_prepare_f32: bad arguments");
  if (workspace_bytes < syl_pcm16_workspace_bytes(batch, t_samp_max))
    return fail(nullptr, SYL_E_WORKSPACE, "syl_prepare_f32: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int chunks = (t_samp_max + PCM_CHUNK - 1) / PCM_CHUNK;
  double* part = reinterpret_cast<double*>(workspace);
  if (normalize) pcm16_stats_kernel<float><<<dim3(chunks, batch), PCM_THREADS, 0, st>>>(wav, offsets, n_samples, chunks, part);
  pcm16_apply_kernel<float><<<dim3(chunks, batch), PCM_THREADS, 0, st>>>(wav, offsets, n_samples, chunks, part, normalize, t_samp_max, wav_out);
  return launch_ok() ? SYL_OK : fail(nullptr, SYL_E_CUDA, "syl_prepare_f32 launch failed: %s", launch_err());
}

int syl_resample(const float* wav_in, const int32_t* n_in, int batch, int t_in_max, const float* kernel, int orig_g, int new_g,
                 int width, float* wav_out, int32_t* n_out, int t_out_max, void* stream) {
  SYL_ENTER();
  if (!wav_in || !n_in || !kernel || !wav_out || batch <= 0 || t_in_max <= 0 || t_out_max <= 0 || orig_g <= 0 || new_g <= 0 || width < 0)
    return fail(nullptr, SYL_E_ARG, "syl_resample: bad arguments");
  resample_kernel<<<dim3((t_out_max + 255) / 256, batch), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      wav_in, n_in, t_in_max, kernel, orig_g, new_g, width, wav_out, n_out, t_out_max);
  return launch_ok() ? SYL_OK : fail(nullptr, SYL_E_CUDA, "syl_resample launch failed: %s", launch_err());
}

int syl_kmeans_assign(const float* feats, int n, const float* centroids, int K, int normalize, int32_t* idx_out,
                      float* dist_out, void* stream) {
  SYL_ENTER();
  if (n == 0) return SYL_OK;
  if (!feats || !centroids || !idx_out || n < 0 || K <= 0) return fail(nullptr, SYL_E_ARG, "syl_kmeans_assign: bad arguments");
  kmeans_assign_kernel<<<(n + KM_ROWS - 1) / KM_ROWS, KM_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      feats, n, centroids, K, normalize, idx_out, dist_out);
  return launch_ok() ? SYL_OK : fail(nullptr, SYL_E_CUDA, "syl_kmeans_assign launch failed: %s", launch_err());
}

size_t syl_segment_workspace_bytes(int batch, int T) {
  if (batch <= 0 || T <= 0) return 0;
  return (((size_t)batch * T * 2 * sizeof(float) + 1023) & ~size_t(1023)) + (size_t)batch * 6 * (T + 1) * sizeof(int32_t);
}

int syl_segment(const float* states, int batch, int T, float thr_norm, float thr_merge, int32_t* seg,
                int32_t* seg_count, float* seg_feat, int max_seg, void* workspace, size_t workspace_bytes,
                void* stream) {
  SYL_ENTER();
  if (!states || !seg || !seg_count || !workspace || batch <= 0 || T <= 0 || max_seg <= 0) return SYL_E_ARG;
  if (workspace_bytes < syl_segment_workspace_bytes(batch, T)) return SYL_E_WORKSPACE;
  float* nsq = reinterpret_cast<float*>(workspace);
  int32_t* scratch = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) +
                                                (((size_t)batch * T * 2 * sizeof(float) + 1023) & ~size_t(1023)));
  return run_segment(states, batch, T, thr_norm, thr_merge, seg, seg_count, seg_feat, max_seg, nsq, scratch,
                     reinterpret_cast<cudaStream_t>(stream));
}

size_t syl_gemm_workspace_bytes(int M, int N, int K) {
  auto al = [](size_t b) { return (b + 1023) & ~size_t(1023); };
  return 2 * al((size_t)M * K * 2) + 2 * al((size_t)N * K * 2);
}

int syl_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, float* out, int M, int N,
                 int K, int n_pass, int act, void* workspace, size_t workspace_bytes, void* stream) {
  SYL_ENTER();
  if (!A || !W || !out || !workspace || M <= 0 || N <= 0 || K <= 0) return fail(nullptr, SYL_E_ARG, "syl_gemm_f32: bad arguments");
  if (N % 256 != 0 || K % 64 != 0) return fail(nullptr, SYL_E_ARG, "syl_gemm_f32: N must be a multiple of 256 and K of 64");
  if (n_pass != 1 && n_pass != 3) return fail(nullptr, SYL_E_ARG, "syl_gemm_f32: n_pass must be 1 or 3");
  if (workspace_bytes < syl_gemm_workspace_bytes(M, N, K)) return fail(nullptr, SYL_E_WORKSPACE, "syl_gemm_f32: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto al = [](size_t b) { return (b + 1023) & ~size_t(1023); };
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __half* a_hi = reinterpret_cast<__half*>(ws);
  __half* a_lo = reinterpret_cast<__half*>(ws + al((size_t)M * K * 2));
  __half* w_hi = reinterpret_cast<__half*>(ws + 2 * al((size_t)M * K * 2));
  __half* w_lo = reinterpret_cast<__half*>(ws + 2 * al((size_t)M * K * 2) + al((size_t)N * K * 2));
  split_f32_kernel<<<grid_for((size_t)M * K), 256, 0, st>>>(A, a_hi, a_lo, (size_t)M * K);
  split_f32_kernel<<<grid_for((size_t)N * K), 256, 0, st>>>(W, w_hi, w_lo, (size_t)N * K);
  std::string err;
  uint64_t wd[2] = {(uint64_t)K, (uint64_t)N}, wsd[2] = {1, (uint64_t)K};
  GemmOp op;
  op.p = base_params();
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 1}, str[3] = {1, (uint64_t)K, (uint64_t)M * K};
  if (!make_tmap_f16(&op.a_hi, a_hi, 3, dims, str, 128, &err) || !make_tmap_f16(&op.a_lo, a_lo, 3, dims, str, 128, &err))
    return fail(nullptr, SYL_E_CUDA, "%s", err.c_str());
  memset(&op.o3_hi, 0, sizeof(CUtensorMap));
  memset(&op.o3_lo, 0, sizeof(CUtensorMap));
  {
    uint64_t od[3] = {(uint64_t)N, (uint64_t)M, 1}, os[3] = {1, (uint64_t)N, (uint64_t)M * N};
    if (!make_tmap(&op.o3_f32, out, 4, 3, od, os, 16, 32, 64, &err)) return fail(nullptr, SYL_E_CUDA, "%s", err.c_str());
  }
  op.p.rows_per_batch = M;
  op.p.batches = 1;
  op.p.N = N;
  op.p.kb_per_pass = K / GEMM_BLOCK_K;
  op.p.n_pass = n_pass;
  op.p.bias = bias;
  op.p.act = act;
  op.p.out_f32 = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CUtensorMap b2_hi, b2_lo;     // this CTA's half of the B tile
  if (!make_tmap_f16(&b2_hi, w_hi, 2, wd, wsd, 128, &err) || !make_tmap_f16(&b2_lo, w_lo, 2, wd, wsd, 128, &err))
    return fail(nullptr, SYL_E_CUDA, "%s", err.c_str());
  if (cudaFuncSetAttribute(gemm3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_TOTAL) != cudaSuccess)
    return fail(nullptr, SYL_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (launch_gemm3_raw(op, b2_hi, b2_lo, st, sms) != SYL_OK)
    return fail(nullptr, SYL_E_CUDA, "gemm launch failed: %s", launch_err());
  if (residual) add_inplace_kernel<<<grid_for((size_t)M * N), 256, 0, st>>>(out, residual, (size_t)M * N);
  return launch_ok() ? SYL_OK : fail(nullptr, SYL_E_CUDA, "launch failed: %s", launch_err());
}

#ifdef SYL_DIAG
int syl_gemm_set_trace(void* trace_dev) {
  g_gemm_trace = reinterpret_cast<long long*>(trace_dev);
  return SYL_OK;
}

int syl_mma_probe(int n, int iters, int ctas, void* cycles_out_dev, void* stream) {
  SYL_ENTER();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long* out = reinterpret_cast<long long*>(cycles_out_dev);
  switch (n) {
    case 16: return run_mma_probe<16>(iters, ctas, out, st);
    case 32: return run_mma_probe<32>(iters, ctas, out, st);
    case 48: return run_mma_probe<48>(iters, ctas, out, st);
    case 64: return run_mma_probe<64>(iters, ctas, out, st);
    case 96: return run_mma_probe<96>(iters, ctas, out, st);
    case 128: return run_mma_probe<128>(iters, ctas, out, st);
    case 256: return run_mma_probe<256>(iters, ctas, out, st);
    default: return SYL_E_ARG;
  }
}
#endif

int syl_powf_half(const float* x, float* y, int64_t n, void* stream) {
  SYL_ENTER();
  if (!x || !y || n <= 0) return SYL_E_ARG;
  powf_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n);
  return launch_ok() ? SYL_OK : SYL_E_CUDA;
}

int syl_saturation_scan(syl_handle* h, unsigned long long* count_dev, void* stream) {
  SYL_ENTER();
  if (!h || !count_dev) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (!h->plans[h->plan_cur].valid) return fail(h, SYL_E_STATE, "syl_saturation_scan: no forward has run yet");
  const Plan& pl = h->plans[h->plan_cur];
  const WsLayout& L = pl.lay;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t B = pl.batch, M = B * L.T;
  CUDA_TRY(h, cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), st));
  auto scan = [&](size_t off, size_t n) {
    count_saturated_kernel<<<grid_for(n), 256, 0, st>>>(at<__half>(pl.ws, off), n, count_dev);
  };
  for (int i = 0; i < 6; ++i) scan(L.act_hi[i], B * L.L[i] * kC);     // conv0 .. conv5 activations
  scan(L.ln_hi, M * kC);
  scan(L.h16_hi, M * kH);                                            // residual stream (last layer's state)
  scan(L.qkv, M * 3 * kH);
  scan(L.ctx_hi, M * kH);
  scan(L.mid_hi, M * kF);
  return launch_ok() ? SYL_OK : fail(h, SYL_E_CUDA, "syl_saturation_scan launch failed: %s", launch_err());
}

int syl_num_stages(void) { return ST_COUNT; }

const char* syl_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }

int syl_profile_enable(syl_handle* h, int on) {
  if (!h) return SYL_E_ARG;
  h->profile = on != 0;
  return SYL_OK;
}

int syl_profile_read(syl_handle* h, float* ms, int* counts) {
  if (!h || !ms || !counts) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  for (int i = 0; i < ST_COUNT; ++i) {
    ms[i] = 0.0f;
    counts[i] = 0;
  }
  for (const ProfRec& r : h->recs) {
    CUDA_TRY(h, cudaEventSynchronize(r.b));
    float t = 0.0f;
    CUDA_TRY(h, cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.stage] += t;
    counts[r.stage] += 1;
  }
  h->recs.clear();
  h->pool_used = 0;
  return SYL_OK;
}

int syl_read_stage(syl_handle* h, const char* name, float* out, size_t n_floats, void* stream) {
  SYL_ENTER();
  if (!h || !name || !out) return SYL_E_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard dg(h->device);
  if (!h->plans[h->plan_cur].valid) return fail(h, SYL_E_STATE, "syl_read_stage: no forward has run yet");
  const Plan& pl = h->plans[h->plan_cur];
  const WsLayout& L = pl.lay;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const std::string s(name);
  const size_t B = pl.batch, M = B * L.T;
  if (s.size() == 5 && s.compare(0, 4, "conv") == 0 && s[4] >= '0' && s[4] <= '5') {
    const int i = s[4] - '0';
    const size_t n = B * L.L[i] * kC;
    if (n_floats < n) return fail(h, SYL_E_ARG, "syl_read_stage: output too small (%zu < %zu)", n_floats, n);
    const bool has_lo = conv_split(h->mode, i + 1);
    join_f16_kernel<<<grid_for(n), 256, 0, st>>>(at<__half>(pl.ws, L.act_hi[i]), has_lo ? at<__half>(pl.ws, L.act_lo[i]) : nullptr,
                                                 out, n);
    return SYL_OK;
  }
  const float* src = nullptr;
  size_t n = 0;
  if (s == "conv6") { src = at<float>(pl.ws, L.conv6); n = M * kC; }
  else if (s == "pos") { src = at<float>(pl.ws, L.pos); n = M * kH; }
  else if (s == "pre") { src = at<float>(pl.ws, L.pre); n = M * kH; }
  else if (s == "h") {      // the residual stream lives as an fp16 pair
    if (n_floats < M * kH) return fail(h, SYL_E_ARG, "syl_read_stage: output too small");
    join_f16_kernel<<<grid_for(M * kH), 256, 0, st>>>(at<__half>(pl.ws, L.h16_hi), at<__half>(pl.ws, L.h16_lo), out, M * kH);
    return SYL_OK;
  }
  else return fail(h, SYL_E_ARG, "syl_read_stage: unknown stage '%s'", name);
  if (n_floats < n) return fail(h, SYL_E_ARG, "syl_read_stage: output too small (%zu < %zu)", n_floats, n);
  CUDA_TRY(h, cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SYL_OK;
}

}  // extern "C"
