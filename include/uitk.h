/*
 * uitk.h — C ABI of the B200 (sm_100a) UiT inference kernels.
 *
 * The reference (RicherMans/UIT_Mobile) is pure Python and has no native/FFI interface; its extension point is
 * the `models` package name lookup + the nn.Module protocol (models/__init__.py:1-2, models/uit.py:252-493).
 * This library is what the Python mirror of that module (`uit_mobile_b200/models/uit.py`) binds with ctypes.
 * Every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are marked `d_`, host pointers `h_`.
 *   - returns 0 (UITK_OK) or a negative UITK_E* code; `uitk_last_error()` gives a thread-local message.
 *   - never allocates device memory, never synchronises, never frees: the caller owns all buffers, passes a
 *     workspace and the CUDA stream (`cudaStream_t` as void*) to launch on, and keeps buffers alive until
 *     the stream has passed the launch.  Re-entrant across streams/threads.
 *   - all tensors are dense row-major fp32 unless stated.
 */
#ifndef UITK_H_
#define UITK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define UITK_API __attribute__((visibility("default")))
#else
#define UITK_API
#endif

#define UITK_VERSION 230

#define UITK_OK 0
#define UITK_EINVAL (-1)    /* bad shape / argument */
#define UITK_EALIGN (-2)    /* misaligned pointer */
#define UITK_ECUDA (-3)     /* CUDA runtime error (message holds cudaGetErrorString) */
#define UITK_EARCH (-4)     /* device is not sm_100 */
#define UITK_ENOSPACE (-5)  /* workspace / blob too small */

/* Fixed geometry of the hot path (models/uit.py:287-308, 581-635). */
#define UITK_N_FFT 512
#define UITK_HOP 160
#define UITK_N_MELS 64
#define UITK_N_FREQS 257
#define UITK_EMBED 128
#define UITK_PATCH 16
#define UITK_MAX_TOKENS 24
#define UITK_INNER 32
#define UITK_HIDDEN 384

/* Encoder arithmetic. */
#define UITK_PREC_FP32 0   /* fp32 CUDA-core GEMMs (validation path) */
#define UITK_PREC_BF16 1   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate/residual/LN/softmax */

/* UITBase variants behind the same API (models/uit.py:89-178, 181-203, 389-412). */
#define UITK_ATTN_BNECK 0  /* BNeckAttention: qkv 128->96, 2 heads x 16, proj 32->128 (uit.py:89-122) */
#define UITK_ATTN_FULL 1   /* Attention: qkv 128->384, 2 heads x 64, proj 128->128 (uit.py:124-178, causal=False) */
#define UITK_ACT_RELU 0
#define UITK_ACT_GELU 1    /* nn.GELU (exact erf), the UITBase default (uit.py:338) */
#define UITK_POOL_MEAN 0   /* mean over tokens (uit.py:402-404) */
#define UITK_POOL_TOKEN 1  /* cls token + token_pos_embed prepended, head on token 0 (uit.py:389-392, 399-401) */
#define UITK_POOL_DM 2     /* mean over frequency, head per time step, mean of the scores (uit.py:405-412) */

UITK_API int uitk_version(void);
UITK_API const char* uitk_last_error(void);

/* Number of CUDA kernels this library has launched in this process (monotonic; for bench accounting). */
UITK_API uint64_t uitk_kernel_launches(void);

/* Number of STFT frames for a clip of L samples: 1 + L/160 (torch.stft center=True). */
UITK_API int64_t uitk_num_frames(int64_t L);

/* Number of eval crops and tokens per crop for T frames (models/uit.py:468-481, 43-74). */
UITK_API int uitk_num_crops(int64_t T, int target_length);
UITK_API int uitk_tokens_per_crop(int64_t T, int target_length);

/* ---- front-end constants -------------------------------------------------------------------------------
 * Packs the module's persistent buffers `front_end.0.spectrogram.window[512]` and
 * `front_end.0.mel_scale.fb[257,64]` (state_dict entries, Q9) plus FFT twiddles into the device layout of the
 * log-mel kernel.  Host -> host; the caller uploads the blob.  Call with h_blob == NULL to get the size. */
UITK_API size_t uitk_frontend_blob_bytes(const float* h_fb);
UITK_API int uitk_pack_frontend(const float* h_window, const float* h_fb, void* h_blob, size_t blob_bytes);

/* ---- log-mel front-end ------------------------------------------------------------------------------------
 * Replaces front_end = MelSpectrogram(16 kHz, n_fft 512, win 512, hop 160, 64 mels, center, reflect, power 2)
 * -> AmplitudeToDB('power') WITHOUT the top-dB clamp (models/uit.py:298-308, 455; torchaudio
 * functional.spectrogram / MelScale / amplitude_to_DB).
 *   d_wav      [B, L] fp32, row stride ld_wav elements (ld_wav < L gives overlapping sliding windows)
 *   d_db       [B, 64, T] fp32, T = 1 + L/160: 10*log10(max(mel, 1e-10)), NOT yet clamped
 *   d_max_pow  one uint32 = bit pattern of the running max mel POWER (non-negative float, so unsigned
 *              order == float order).  Caller zero-initialises; the kernel atomically maxes into it.
 *              The reference's single batch-global cutoff (Q2) is max_db - 120 with
 *              max_db = 10*log10(max(max_pow, 1e-10)); it is applied by uitk_clamp_db / uitk_encoder, after
 *              the caller has had the chance to all-reduce(max) the word across GPUs.
 *   d_min_pow  optional (may be NULL): same for the running MIN mel power; caller initialises to 0x7f800000 (+inf).
 *              A pipelined caller that encodes chunk i with the running maximum of chunks <= i uses it to prove
 *              afterwards that the clamp was inactive (min_db >= final max_db - 120) and the result therefore exact. */
UITK_API int uitk_logmel(const float* d_wav, int64_t B, int64_t L, int64_t ld_wav, const void* d_frontend_blob,
                float* d_db, uint32_t* d_max_pow, uint32_t* d_min_pow, void* stream);

/* Same front-end on 16-bit PCM (x = pcm / 32768: the reference's own ingest normalisation, dataset.py:44-46 and
 * torchaudio.load in inference.py:52).  Bit-identical to uitk_logmel on the converted floats; halves the input bytes. */
UITK_API int uitk_logmel_i16(const int16_t* d_pcm, int64_t B, int64_t L, int64_t ld_pcm, const void* d_frontend_blob,
                    float* d_db, uint32_t* d_max_pow, uint32_t* d_min_pow, void* stream);

/* Sliding 1 s (or any) windows over ONE long stream, window hop a multiple of the STFT hop (the reference
 * has no such entry point - its callers slice the waveform and run front_end per slice, uit.py:455 / 468-472).
 * Produces exactly what uitk_logmel(d_stream, W, window, hop, ...) produces for the W = (n_samples - window)/hop + 1
 * overlapping windows - bit-identical d_db [W, 64, 1 + window/160] and max/min words - but computes every interior STFT
 * frame of the stream ONCE (a window frame t in [2, T-2) does not touch the window's reflect padding, so it is frame
 * w*hop/160 + t of the stream) and only the 4 edge frames per window separately.
 *   d_workspace  uitk_logmel_sliding_workspace_bytes(n_samples) bytes: the stream's log-mel [64, 1 + n_samples/160] */
UITK_API size_t uitk_logmel_sliding_workspace_bytes(int64_t n_samples);
UITK_API int uitk_logmel_sliding(const float* d_stream, int64_t n_samples, int64_t window, int64_t hop, const void* d_frontend_blob,
                        float* d_db, uint32_t* d_max_pow, uint32_t* d_min_pow, void* d_workspace, size_t workspace_bytes,
                        void* stream);

/* In-place top-dB clamp: db = max(db, 10*log10(max(max_pow,1e-10)) - top_db)  (amplitude_to_DB top_db=120).
 * d_min_pow (may be NULL): the batch's minimum power word from uitk_logmel; when given, the kernel decides on the device
 * whether any value lies under the cutoff and returns without touching memory if none does. */
UITK_API int uitk_clamp_db(float* d_db, int64_t n, const uint32_t* d_max_pow, const uint32_t* d_min_pow, float top_db, void* stream);

/* ---- encoder weights ----------------------------------------------------------------------------------------
 * Packs state_dict tensors (host fp32, reference layout) into the kernels' device layout.
 * `h_tensors` is an array of host pointers in the fixed order documented in uitk_encoder_tensor_names().
 * Host -> host; the caller uploads the blob. */
typedef struct {
  int depth;          /* 12 / 6 / 4 */
  int outputdim;      /* 537 */
  int grid_t;         /* time_pos_embed length (6 for target_length 102) */
  int precision;      /* UITK_PREC_* : the tensor-core megakernel runs the UiT-XS/XXS/XXXS configuration
                         (BNECK + RELU + MEAN); every other variant runs on the fp32 CUDA-core kernels */
  int attention;      /* UITK_ATTN_* */
  int act;            /* UITK_ACT_* */
  int pooling;        /* UITK_POOL_* */
  int reserved;       /* 0 */
} uitk_encoder_cfg;

UITK_API int uitk_encoder_num_tensors(int depth);
UITK_API const char* uitk_encoder_tensor_name(int depth, int index);   /* state_dict key of h_tensors[index] */
/* tokens per crop incl. the cls token of UITK_POOL_TOKEN */
UITK_API int uitk_tokens_total(const uitk_encoder_cfg* cfg, int64_t T, int target_length);
UITK_API size_t uitk_encoder_blob_bytes(const uitk_encoder_cfg* cfg);
UITK_API int uitk_pack_encoder(const uitk_encoder_cfg* cfg, const float* const* h_tensors, void* h_blob, size_t blob_bytes);

/* ---- encoder ----------------------------------------------------------------------------------------------------
 * Replaces init_bn + crop loop + forward_features + forward_head (models/uit.py:460-492, 379-412, 89-122,
 * 181-248).  UiT-XS/XXS/XXXS configuration (pooling='mean', BNeckAttention, ReLU MLP) on the tcgen05 megakernel when
 * cfg->precision == UITK_PREC_BF16; full Attention / GELU / pooling 'token' | 'dm' on the fp32 CUDA-core kernels.
 *   d_db        [B, 64, T] un-clamped dB from uitk_logmel; the clamp with *d_max_pow is fused into the load
 *   d_probs     [B, outputdim] sigmoid scores, crops reduced by mean (eval_avg 0) or max (1)
 *   workspace   uitk_encoder_workspace_bytes(cfg, B, T, target_length) bytes, 256-B aligned */
UITK_API size_t uitk_encoder_workspace_bytes(const uitk_encoder_cfg* cfg, int64_t B, int64_t T, int target_length);
UITK_API int uitk_encoder(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T,
                 int target_length, int eval_avg, const uint32_t* d_max_pow, float* d_probs,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/* Conditional exact re-run for callers that encoded SPECULATIVELY with a cutoff that was not yet batch-global (a rank's
 * local maximum while the all-reduce(MAX) of the word is still in flight; models/uit.py forward with a process group).
 * Same arguments as uitk_encoder plus three device words; every kernel of the launch decides ON THE DEVICE whether
 * anything could differ and returns immediately if not (no host synchronisation, no collective on the critical path):
 *   re-run  <=>  *d_max_used != *d_max_pow  and  dB(*d_min_pow) < dB(*d_max_pow) - 120
 * i.e. the speculative cutoff was lower than the true one AND some value of this call's input lies below the true cutoff.
 * Otherwise no value was (or would have been) clamped and the speculative scores are exactly the reference's.
 * Tensor-core configuration only (returns UITK_EINVAL for fp32 / variant configurations: encode after the all-reduce). */
UITK_API int uitk_encoder_fixup(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T,
                       int target_length, int eval_avg, const uint32_t* d_max_pow, const uint32_t* d_max_used,
                       const uint32_t* d_min_pow, float* d_probs, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- the batch-global top-dB scope (Q2) across the GPUs of one box, without a collective library call -----------------
 * Replaces `torch.distributed.all_reduce(max_pow, MAX)` (what a sharded caller of models/uit.py:455-462 needs so that
 * AmplitudeToDB's cutoff spans the global batch).  Every rank owns a slot of uitk_peer_words_slot_bytes() bytes (zeroed once)
 * that all peers have mapped (NVLink peer memory: CUDA IPC / torch symmetric memory);
 *   uitk_peer_words_publish: stores *d_word and then `epoch` (release, system scope) into this rank's slot        - 1 thread
 *   uitk_peer_words_collect: waits until the slot of every rank shows `epoch`, writes the MAX of the words        - 1 warp
 * d_peer_slots is a DEVICE array of n_ranks pointers (own slot included).  Epochs count 1, 2, 3, ... and every rank must call
 * publish(e) and collect(e) in this order on one stream for every e (the ring holds 4 epochs: a rank can be at most one epoch
 * ahead of a peer that still has to read).  The collect kernel spins on peer memory: launch it where the wait is short (after
 * the speculative uitk_encoder, before uitk_encoder_fixup). */
UITK_API size_t uitk_peer_words_slot_bytes(void);
UITK_API int uitk_peer_words_publish(uint32_t* d_my_slot, const uint32_t* d_word, uint32_t epoch, void* stream);
UITK_API int uitk_peer_words_collect(const uint32_t* const* d_peer_slots, int n_ranks, uint32_t epoch, uint32_t* d_word_out, void* stream);

/* ---- forward_features / forward_head (models/uit.py:379-412) as separate entry points -----------------------------
 * The reference exposes both methods; uitk_encoder fuses them.  These two keep the method-level contract:
 *   uitk_forward_features: d_spec [B, 64, T] = the ALREADY NORMALISED spectrogram the reference passes in (front_end +
 *       init_bn applied by the caller, e.g. through uitk_logmel / uitk_clamp_db / uitk_init_bn), T <= 16*grid_t + 15
 *       -> d_tokens [B, N, 128] after the final LayerNorm, N = 4 * ((T - 16)/16 + 1) (+1 cls token for UITK_POOL_TOKEN).
 *   uitk_forward_head: d_tokens [B, N, 128] -> d_probs [B, outputdim] (pooling per cfg, LayerNorm 1e-5, Linear, sigmoid).
 *       For UITK_POOL_DM the time-patch count is N / 4. */
UITK_API int uitk_init_bn(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T,
                          float* d_out, void* stream);
UITK_API size_t uitk_forward_features_workspace_bytes(const uitk_encoder_cfg* cfg, int64_t B, int64_t T);
UITK_API int uitk_forward_features(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_spec, int64_t B, int64_t T,
                          float* d_tokens, void* d_workspace, size_t workspace_bytes, void* stream);
UITK_API int uitk_forward_head(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_tokens, int64_t B,
                      int n_tokens, float* d_probs, void* stream);

/* ---- MobileNetV2 (models/mobilenetv2.py:66-178; the reference's distillation teacher / audio-tagging baseline) ------------
 * Eval forward of the default configuration (inverted_residual_setting of mobilenetv2.py:106-116, width_mult 1.0,
 * last_channel 1280): conv3x3/2 + BN + ReLU6, 17 inverted-residual blocks, conv1x1 + BN + ReLU6, mean over the mel axis,
 * Linear(1280 -> outputdim) per time step, sigmoid, mean over time.  fp32 CUDA-core kernels, eval BatchNorm folded at pack time.
 *   h_tensors   uitk_mnv2_num_tensors() host pointers in the order of uitk_mnv2_tensor_name(i) (state_dict keys)
 *   d_db        [B, 64, T] log-mel dB AFTER the top-dB clamp (uitk_logmel + uitk_clamp_db = the model's front_end)
 *   d_probs     [B, outputdim]
 *   workspace   uitk_mnv2_workspace_bytes(B, T) bytes, 256-B aligned (three activation buffers of <= 256 clips) */
UITK_API int uitk_mnv2_num_tensors(void);
UITK_API const char* uitk_mnv2_tensor_name(int index);
UITK_API size_t uitk_mnv2_blob_bytes(int outputdim);
UITK_API int uitk_pack_mnv2(int outputdim, const float* const* h_tensors, void* h_blob, size_t blob_bytes);
UITK_API size_t uitk_mnv2_workspace_bytes(int64_t B, int64_t T);
UITK_API int uitk_mnv2_forward(int outputdim, const void* d_blob, const float* d_db, int64_t B, int64_t T, float* d_probs,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/* Debug/validation taps (tests only): copy of the token activations [B*crops*tokens, 128] after patch embed
 * (stage 0) or after block i (stage i+1) is left in the workspace at this byte offset after uitk_encoder. */
UITK_API size_t uitk_encoder_tokens_offset(const uitk_encoder_cfg* cfg, int64_t B, int64_t T, int target_length);

/* Tests only: when enabled, the tensor-core encoder also writes the residual stream after the last block
 * (before the final LayerNorm) to the start of the workspace, like the fp32 path does. */
UITK_API void uitk_debug_taps(int enable);

/* Profiling only: when the library is built with -DUITK_TRACE (UITK_TRACE=1 python -m uit_mobile_b200.build), CTA 0 of
 * the tensor-core encoder stamps (stage id << 44 | SM clock) at every stage boundary; `which` 0 = compute thread 0,
 * 1 = the MMA-issuer thread.  Copies the first n (<= 4096) stamps of the last launch to host_out (synchronises).
 * Returns UITK_EINVAL in a normal build. */
UITK_API int uitk_debug_read_trace(long long* host_out, int which, int n);

/* Tests only: one 128 x N x K tcgen05 GEMM through the library's descriptor / TMEM / bulk-copy plumbing.
 * d_B_packed is bf16 in the K-major core-matrix layout documented in csrc/tc_ptx.cuh; C (+)= A B^T in fp32. */
UITK_API int uitk_selftest_umma(const float* d_A, const void* d_B_packed, const float* d_C_init, float* d_C, int N, int K,
                                void* stream);

/* Same GEMM with the A operand staged in TENSOR MEMORY (tcgen05.mma with a TMEM A operand, as the encoder's P V
 * attention product uses it): pins the packed-bf16 TMEM operand layout. */
UITK_API int uitk_selftest_umma_ts(const float* d_A, const void* d_B_packed, const float* d_C_init, float* d_C, int N, int K,
                                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UITK_H_ */
