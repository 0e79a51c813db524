/* sylber_b200 - C ABI of the B200-native Segmenter forward path.
 *
 * The reference (Berkeley-Speech-Group/sylber) has no FFI: its boundary is Python-level, three calls inside
 * Segmenter.__call__ (sylber/model/sylber.py:63-138):
 *     :122  self.speech_model(batch_tensor, attention_mask=...).last_hidden_state      -> syl_forward (hidden)
 *     :126  get_segment(states, norm_threshold, merge_threshold)                       -> syl_forward (seg, seg_count)
 *     :133  states[s:e].mean(0)                                                        -> syl_forward (seg_feat)
 * and the weights enter through load_state_dict at :51-52                              -> syl_load_weight / syl_finalize.
 * This header is what a ctypes / cffi binding on the reference side would bind (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer argument named *_dev / wav / hidden / seg... is a DEVICE pointer on the handle's GPU unless
 *     stated otherwise; the caller owns all of them, including the workspace.  The library owns only its packed
 *     weights and never allocates per call.
 *   - functions return 0 on success or a negative SYL_E_* code; nothing throws across the ABI; the message of the
 *     last failure is available from syl_last_error().
 *   - work is enqueued on the CUDA stream passed in (a cudaStream_t cast to void*); the library never
 *     synchronises the device except inside syl_finalize().
 *   - a handle belongs to one device; entry points that take a handle switch to that device for the duration of the
 *     call and restore the caller's current device before returning.  Handle-less entry points (syl_segment,
 *     syl_attention, syl_kmeans_assign, syl_prepare_*, syl_resample, syl_gemm_f32, syl_powf_half) launch on the
 *     CURRENT device - the one their pointer arguments must live on.
 *   - entry points that take a handle serialise on a mutex inside it (plan cache, CUDA-graph cache and profiling
 *     records are per-handle state), so several host threads may share a handle; forwards that are in flight at the
 *     same time must use different workspaces and output buffers.
 *   - the library reads no environment variables.  Experiment switches and the diagnostic entry points exist only
 *     in the -DSYL_DIAG build (libsylber_b200_diag.so).
 *   - there is no CPU fallback: without an sm_100 GPU every compute entry point fails with SYL_E_CUDA.
 */
#ifndef SYLBER_B200_H
#define SYLBER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct syl_handle syl_handle;

#define SYL_OK 0
#define SYL_E_ARG (-1)      /* bad argument / shape */
#define SYL_E_STATE (-2)    /* called in the wrong state (e.g. forward before finalize, missing weight) */
#define SYL_E_CUDA (-3)     /* CUDA runtime / driver failure */
#define SYL_E_WORKSPACE (-4) /* workspace too small */

/* precision mode = bit mask of the GEMM sites that run split precision (fp16 hi + lo operands, 3 tensor-core
 * passes, ~fp32 accuracy) instead of a single fp16 pass.  DESIGN.md explains the measured error budget. */
#define SYL_SPLIT_CONV 1     /* shorthand: conv2..conv6 of the feature encoder (= SYL_SPLIT_CONV2 | ... | SYL_SPLIT_CONV6) */
#define SYL_SPLIT_CONV1 8    /* conv1, half of the conv stack's FLOPs */
#define SYL_SPLIT_PROJ 2     /* shorthand: feature projection + positional conv (= SYL_SPLIT_FPROJ | SYL_SPLIT_POS) */
#define SYL_SPLIT_ENC 4      /* encoder linear layers (QKV, out-proj, FFN) */
#define SYL_SPLIT_CONV2 16   /* individual sites, for error-budget experiments (tests/mode_sweep.py) */
#define SYL_SPLIT_CONV3 32
#define SYL_SPLIT_CONV4 64
#define SYL_SPLIT_CONV5 128
#define SYL_SPLIT_CONV6 256
#define SYL_SPLIT_FPROJ 512
#define SYL_SPLIT_POS 1024
/* Presets.  Measured on B200 against the fp32 CPU oracle (relative Frobenius error of the final hidden states, bar
 * 1e-3; device ms per step of batch 32 x 10 s; clips whose SEGMENTS equal the fp32 reference's on configs 2 / 5,
 * profiles/r03_segment_agreement.md): fast 5.1e-4 / 4.9 ms / 24 of 32, 8 of 16; parity 4.1e-4 / 5.1 ms / 27 of 32,
 * 10 of 16; strict 3.0e-4 / 7.5 ms; exact 1.2e-5 / 10.9 ms / 32 of 32, 16 of 16.  FAST is the default since round 3:
 * splitting conv4-6 and the projection costs 6 % of the step for 1e-4 of state error and a handful of threshold
 * decisions, and both sit at half the tolerance.  EXACT is the preset for segment identity with the fp32 reference. */
#define SYL_MODE_FAST 0                                                      /* single-pass fp16 everywhere (default) */
#define SYL_MODE_PARITY (SYL_SPLIT_CONV4 | SYL_SPLIT_CONV5 | SYL_SPLIT_CONV6 | SYL_SPLIT_FPROJ)
#define SYL_MODE_STRICT (SYL_SPLIT_CONV | SYL_SPLIT_CONV1 | SYL_SPLIT_PROJ)  /* whole front end split */
#define SYL_MODE_EXACT (SYL_SPLIT_CONV | SYL_SPLIT_CONV1 | SYL_SPLIT_PROJ | SYL_SPLIT_ENC)   /* every GEMM site split */

/* Option bit of the mode argument (not a precision site).  TRIMMED MODE, an opt-in deviation from the reference's
 * padding semantics (sylber/model/sylber.py:93-126): the reference computes, returns and segments the padded frames of
 * a short utterance in a batch; with this bit the frames beyond an utterance's valid length are NOT computed - conv,
 * GEMM and attention tiles that lie entirely in the padding are skipped, hidden rows >= valid_frames[b] come back as
 * zeros and therefore carry no segments.  Valid frames are unchanged: they still see the batch-wide padded length
 * through the conv-0 GroupNorm statistics, exactly as in the reference (tests/test_gpu_trim.py). */
#define SYL_TRIM_PADDING 4096

#define SYL_DTYPE_F32 0

/* ---- lifecycle: replaces HubertModel(config) + load_state_dict (sylber/model/sylber.py:41,51-54) ---- */
int syl_create(syl_handle** out, int device, int n_layers, int mode);
/* name: HubertModel state_dict key; dev_ptr: fp32 device tensor, copied.  Both pos-conv namings are accepted
 * (parametrizations.weight.original0/1 and weight_g/weight_v). */
int syl_load_weight(syl_handle* h, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype);
/* folds weight-norm, packs GEMM operands; fails (SYL_E_STATE) naming the first missing tensor */
int syl_finalize(syl_handle* h);
void syl_destroy(syl_handle* h);
const char* syl_last_error(const syl_handle* h);   /* h may be NULL: last error of syl_create */

/* ---- shapes ---- */
int syl_num_frames(int n_samples);                  /* HF _get_feat_extract_output_lengths (modeling_hubert.py:675-688) */
size_t syl_workspace_bytes(const syl_handle* h, int batch, int t_samp_max);

/* ---- the hot path: sylber/model/sylber.py:122-133 ----
 *   wav        [batch, t_samp_max] fp32, zero padded              (sylber.py:93-117)
 *   n_samples  [batch] int32, valid samples per utterance         (what the reference's attention_mask encodes)
 *   hidden     [batch, T, 768] fp32 out, T = syl_num_frames(t_samp_max)     == last_hidden_state
 *   seg        [batch, max_seg, 2] int32 out, frame indices [start, end)    == get_segment(...)
 *   seg_count  [batch] int32 out (may exceed max_seg: then only the first max_seg rows were written)
 *   seg_feat   [batch, max_seg, 768] fp32 out                               == states[s:e].mean(0)
 * seg / seg_count / seg_feat may all be NULL to run the encoder only.
 */
int syl_forward(syl_handle* h, const float* wav, const int32_t* n_samples, int batch, int t_samp_max, float* hidden,
                int32_t* seg, int32_t* seg_count, float* seg_feat, int max_seg, float thr_norm, float thr_merge,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-stage entry points (unit tests, ncu) ---- */
/* conv feature encoder, modeling_hubert.py:203-213: wav -> [batch, T, 512] fp32 (channels last) */
int syl_conv_frontend(syl_handle* h, const float* wav, int batch, int t_samp_max, float* feats, void* workspace,
                      size_t workspace_bytes, void* stream);
/* one post-LN encoder layer, modeling_hubert.py:388-405: h_in/h_out [batch, T, 768] fp32 */
int syl_encoder_layer(syl_handle* h, int layer, const float* h_in, const int32_t* valid_frames, int batch, int T,
                      float* h_out, void* workspace, size_t workspace_bytes, void* stream);
/* attention core: qkv [batch*T, 2304] fp16 (Q pre-scaled by 1/8), out [batch*T, 768] fp16 */
int syl_attention(const void* qkv_f16, const int32_t* kv_len, int batch, int T, void* out_f16, void* stream);
#ifdef SYL_DIAG
/* diagnostic build only: the same launch with CTA 0 logging clock64 stamps of its MMA thread and softmax warpgroups into
 * trace_dev[7][trace_cap] (int64 device memory, zero it first); decoded by tools/attn_trace.py */
int syl_attention_trace(const void* qkv_f16, const int32_t* kv_len, int batch, int T, void* out_f16, void* trace_dev,
                        int trace_cap, void* stream);
#endif
/* ---- the steps either side of the path ----
 * Front door, replaces the host preprocessing of sylber/model/sylber.py:83-87 (file branch) and :93-118 (padding):
 * pcm holds the utterances' 16 kHz int16 samples back to back on the device, utterance b = pcm[offsets[b] ..
 * offsets[b] + n_samples[b]); wav_out[b, :] = (x / 32768 - mean) / std (unbiased std, like torch.std) when
 * normalize != 0, else x / 32768; zero padded to t_samp_max.  workspace >= syl_pcm16_workspace_bytes. */
size_t syl_pcm16_workspace_bytes(int batch, int t_samp_max);
int syl_prepare_pcm16(const int16_t* pcm, const int64_t* offsets, const int32_t* n_samples, int batch, int t_samp_max,
                      int normalize, float* wav_out, void* workspace, size_t workspace_bytes, void* stream);
/* the same for fp32 samples (e.g. the output of syl_resample) */
int syl_prepare_f32(const float* wav, const int64_t* offsets, const int32_t* n_samples, int batch, int t_samp_max,
                    int normalize, float* wav_out, void* workspace, size_t workspace_bytes, void* stream);
/* Band-limited sinc resampling, replaces torchaudio.transforms.Resample(sr, 16000) (sylber/model/sylber.py:85):
 * wav_in [batch, t_in_max] fp32 with n_in[b] valid samples, kernel [new_g, 2 * width + orig_g] from
 * sylber_b200/resample.py (frequencies reduced by their gcd); wav_out [batch, t_out_max], zero filled beyond
 * n_out[b] = ceil(new_g * n_in[b] / orig_g) (n_out may be null). */
int syl_resample(const float* wav_in, const int32_t* n_in, int batch, int t_in_max, const float* kernel, int orig_g, int new_g,
                 int width, float* wav_out, int32_t* n_out, int t_out_max, void* stream);
/* Back door, replaces the codebook search of KMQuantizer.get_indices (sylber/model/quantizer.py:86-110):
 * idx_out[r] = argmin_k |x_r - c_k|^2 over centroids [K, 768] (first minimum wins); normalize != 0 applies
 * x / sqrt(sum x^2 + 1e-8) * 6 first (quantizer.py:99); dist_out (optional) receives the squared distances. */
int syl_kmeans_assign(const float* feats, int n, const float* centroids, int K, int normalize, int32_t* idx_out,
                      float* dist_out, void* stream);
/* segmentation + pooling on given states [batch, T, 768] fp32; workspace >= syl_segment_workspace_bytes */
size_t syl_segment_workspace_bytes(int batch, int T);
int syl_segment(const float* states, int batch, int T, float thr_norm, float thr_merge, int32_t* seg,
                int32_t* seg_count, float* seg_feat, int max_seg, void* workspace, size_t workspace_bytes,
                void* stream);
/* generic tensor-core GEMM on fp32 inputs: out[M,N] = act(A[M,K] W[N,K]^T + bias) + residual */
size_t syl_gemm_workspace_bytes(int M, int N, int K);
int syl_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, float* out, int M, int N,
                 int K, int n_pass, int act, void* workspace, size_t workspace_bytes, void* stream);
#ifdef SYL_DIAG
/* diagnostic build only: cycles for `iters` x 4 back-to-back tcgen05.mma (M=128, N=n, K=16) issued by one thread of each of
 * `ctas` CTAs; writes the elapsed clock64 ticks of CTA 0 to cycles_out_dev (one int64) */
int syl_mma_probe(int n, int iters, int ctas, void* cycles_out_dev, void* stream);
/* diagnostic build only: every following GEMM launch writes the clock64 timeline of its CTA 0 to trace_dev (128 int64;
 * slots in csrc/gemm_tc.cuh); null switches the probe off */
int syl_gemm_set_trace(void* trace_dev);
#endif
/* y[i] = powf(x[i], 0.5f) as glibc computes it (the fp64 replay used by the segmentation kernel) */
int syl_powf_half(const float* x, float* y, int64_t n, void* stream);
/* copy an intermediate of the most recent syl_forward / syl_conv_frontend out as fp32.
 * names: conv0..conv6 ([batch, L_i, 512]), proj, pos, enc_in ([batch, T, 768]) */
int syl_read_stage(syl_handle* h, const char* name, float* out, size_t n_floats, void* stream);
/* fp16 range check (diagnostic, off the hot path): scans the fp16 activation buffers the most recent forward on this
 * handle left in its workspace (conv0-5 outputs, LayerNorm(512) output, residual stream, QKV, attention context, FFN
 * intermediate - the last layer's for the per-layer ones) and writes to *count_dev (one uint64 on the device) how many
 * elements sit at the saturation value +-65504 or are not finite.  Every fp16 store of the forward saturates instead
 * of overflowing, so a non-zero count means this input + checkpoint leaves the fp16 range somewhere. */
int syl_saturation_scan(syl_handle* h, unsigned long long* count_dev, void* stream);
/* run only the first n encoder layers in syl_forward (debug / per-layer parity); n < 0 restores all */
int syl_set_active_layers(syl_handle* h, int n);
/* how many kernels of this library syl_forward launches for the given shape (bench.py reports it) */
int syl_forward_launch_count(const syl_handle* h, int with_segmentation);

/* syl_forward replays the whole launch sequence as one CUDA graph when it is called again with the same argument
 * set (default on; off = always launch eagerly).  Profiling (below) always uses the eager path. */
int syl_set_graph_mode(syl_handle* h, int on);

/* per-stage device timing: when enabled, syl_forward brackets each stage with CUDA events on the launch stream.
 * syl_profile_read waits for the recorded events, returns the accumulated milliseconds and region counts per
 * stage since the previous read (arrays of syl_num_stages() entries) and resets the accumulation. */
int syl_num_stages(void);
const char* syl_stage_name(int stage);
int syl_profile_enable(syl_handle* h, int on);
int syl_profile_read(syl_handle* h, float* ms, int* counts);

#ifdef __cplusplus
}
#endif
#endif /* SYLBER_B200_H */
