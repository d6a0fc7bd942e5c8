// C ABI entry points that launch kernels (see include/uitk.h).  No allocation, no synchronisation.
#include "uitk_common.cuh"

using namespace uitk;

namespace {

int check_arch() {
  static thread_local int ok_dev = -1;
  int dev = 0;
  UITK_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev == ok_dev) return UITK_OK;
  int major = 0;
  UITK_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  UITK_REQUIRE(major == 10, UITK_EARCH, "libuitk is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  ok_dev = dev;
  return UITK_OK;
}

int g_debug_taps = 0;

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace

extern "C" {

int uitk_logmel(const float* d_wav, int64_t B, int64_t L, int64_t ld_wav, const void* d_frontend_blob, float* d_db,
                uint32_t* d_max_pow, uint32_t* d_min_pow, void* stream) {
  UITK_REQUIRE(d_wav && d_frontend_blob && d_db && d_max_pow, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(B >= 0, UITK_EINVAL, "negative batch");
  UITK_REQUIRE(L > UITK_N_FFT / 2, UITK_EINVAL, "reflect padding needs L > 256 samples (got %lld)", (long long)L);
  UITK_REQUIRE(L <= (1ll << 30), UITK_EINVAL, "clip too long (max 2^30 samples per row)");
  UITK_REQUIRE(ld_wav >= 1, UITK_EINVAL, "ld_wav must be >= 1");
  UITK_REQUIRE(aligned(d_wav, 4) && aligned(d_db, 4) && aligned(d_max_pow, 4) && aligned(d_frontend_blob, 16), UITK_EALIGN,
               "misaligned pointer");
  if (B == 0) return UITK_OK;
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return launch_logmel(d_wav, B, L, ld_wav, reinterpret_cast<const FrontendBlob*>(d_frontend_blob), d_db, d_max_pow, d_min_pow,
                       reinterpret_cast<cudaStream_t>(stream));
}

int uitk_logmel_i16(const int16_t* d_pcm, int64_t B, int64_t L, int64_t ld_pcm, const void* d_frontend_blob, float* d_db,
                    uint32_t* d_max_pow, uint32_t* d_min_pow, void* stream) {
  UITK_REQUIRE(d_pcm && d_frontend_blob && d_db && d_max_pow, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(B >= 0, UITK_EINVAL, "negative batch");
  UITK_REQUIRE(L > UITK_N_FFT / 2, UITK_EINVAL, "reflect padding needs L > 256 samples (got %lld)", (long long)L);
  UITK_REQUIRE(L <= (1ll << 30), UITK_EINVAL, "clip too long (max 2^30 samples per row)");
  UITK_REQUIRE(ld_pcm >= 1, UITK_EINVAL, "ld_pcm must be >= 1");
  UITK_REQUIRE(aligned(d_pcm, 2) && aligned(d_db, 4) && aligned(d_max_pow, 4) && aligned(d_frontend_blob, 16), UITK_EALIGN,
               "misaligned pointer");
  if (B == 0) return UITK_OK;
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return launch_logmel_i16(d_pcm, B, L, ld_pcm, reinterpret_cast<const FrontendBlob*>(d_frontend_blob), d_db, d_max_pow, d_min_pow,
                           reinterpret_cast<cudaStream_t>(stream));
}

size_t uitk_peer_words_slot_bytes(void) { return peer_words_slot_bytes(); }

int uitk_peer_words_publish(uint32_t* d_my_slot, const uint32_t* d_word, uint32_t epoch, void* stream) {
  UITK_REQUIRE(d_my_slot && d_word, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(epoch != 0, UITK_EINVAL, "epochs start at 1 (a zeroed slot must not look published)");
  UITK_REQUIRE(aligned(d_my_slot, 16) && aligned(d_word, 4), UITK_EALIGN, "misaligned pointer");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return launch_words_publish(d_my_slot, d_word, epoch, reinterpret_cast<cudaStream_t>(stream));
}

int uitk_peer_words_collect(const uint32_t* const* d_peer_slots, int n_ranks, uint32_t epoch, uint32_t* d_word_out, void* stream) {
  UITK_REQUIRE(d_peer_slots && d_word_out, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(n_ranks >= 1 && n_ranks <= 1024, UITK_EINVAL, "rank count %d out of range", n_ranks);
  UITK_REQUIRE(epoch != 0, UITK_EINVAL, "epochs start at 1");
  UITK_REQUIRE(aligned(d_peer_slots, 8) && aligned(d_word_out, 4), UITK_EALIGN, "misaligned pointer");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return launch_words_collect(d_peer_slots, n_ranks, epoch, d_word_out, reinterpret_cast<cudaStream_t>(stream));
}

size_t uitk_logmel_sliding_workspace_bytes(int64_t n_samples) {
  return n_samples < 0 ? 0 : (size_t)64 * (size_t)(1 + n_samples / UITK_HOP) * sizeof(float);
}

int uitk_logmel_sliding(const float* d_stream, int64_t n_samples, int64_t window, int64_t hop, const void* d_frontend_blob, float* d_db,
                        uint32_t* d_max_pow, uint32_t* d_min_pow, void* d_workspace, size_t workspace_bytes, void* stream) {
  UITK_REQUIRE(d_stream && d_frontend_blob && d_db && d_max_pow && d_workspace, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(hop >= UITK_HOP && hop % UITK_HOP == 0, UITK_EINVAL,
               "window hop %lld must be a positive multiple of the STFT hop (160); use uitk_logmel with ld_wav = hop otherwise", (long long)hop);
  UITK_REQUIRE(window >= 4 * UITK_HOP && window % UITK_HOP == 0 && window <= (1ll << 30), UITK_EINVAL,
               "window %lld must be a multiple of 160 samples, at least 640", (long long)window);
  UITK_REQUIRE(n_samples >= window && n_samples <= (1ll << 30), UITK_EINVAL, "stream shorter than one window (or longer than 2^30 samples)");
  // every stream frame computed below must be an interior frame of SOME window, or it would leak into the batch max / min
  UITK_REQUIRE(hop / UITK_HOP <= window / UITK_HOP - 3, UITK_EINVAL,
               "window hop %lld leaves stream frames that belong to no window (need hop <= window - 480); use uitk_logmel with ld_wav = hop",
               (long long)hop);
  UITK_REQUIRE(aligned(d_stream, 4) && aligned(d_db, 4) && aligned(d_max_pow, 4) && aligned(d_frontend_blob, 16) && aligned(d_workspace, 4),
               UITK_EALIGN, "misaligned pointer");
  UITK_REQUIRE(workspace_bytes >= uitk_logmel_sliding_workspace_bytes(n_samples), UITK_ENOSPACE, "workspace too small: need %zu, have %zu",
               uitk_logmel_sliding_workspace_bytes(n_samples), workspace_bytes);
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  const FrontendBlob* blob = reinterpret_cast<const FrontendBlob*>(d_frontend_blob);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t W = (n_samples - window) / hop + 1;          // windows
  const int64_t Tw = 1 + window / UITK_HOP;                  // frames per window
  const int64_t r = hop / UITK_HOP;                          // window hop in frames
  const int64_t U = 1 + n_samples / UITK_HOP;                // frames of the stream taken as one clip
  float* G = reinterpret_cast<float*>(d_workspace);          // [64][U]
  // A window frame t in [2, Tw-2) lies wholly inside the window: it IS frame w*r + t of the stream.  Compute the stream's
  // frames 2 .. (W-1)*r + Tw - 3 once (every one of them is an interior frame of some window, so they all count for the
  // batch max / min), scatter them into the windows, and compute only the 4 reflect-padded edge frames per window.
  rc = launch_logmel_frames(d_stream, 1, n_samples, n_samples, blob, G, 2, (W - 1) * r + Tw - 4, 64 * U, U, d_max_pow, d_min_pow, s);
  if (rc != UITK_OK) return rc;
  rc = launch_window_gather(G, U, d_db, W, (int)Tw, (int)r, s);
  if (rc != UITK_OK) return rc;
  rc = launch_logmel_frames(d_stream, W, window, hop, blob, d_db, 0, 2, 64 * Tw, Tw, d_max_pow, d_min_pow, s);
  if (rc != UITK_OK) return rc;
  return launch_logmel_frames(d_stream, W, window, hop, blob, d_db, Tw - 2, 2, 64 * Tw, Tw, d_max_pow, d_min_pow, s);
}

int uitk_clamp_db(float* d_db, int64_t n, const uint32_t* d_max_pow, const uint32_t* d_min_pow, float top_db, void* stream) {
  UITK_REQUIRE(d_db && d_max_pow, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(n >= 0, UITK_EINVAL, "negative size");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return launch_clamp_db(d_db, n, d_max_pow, d_min_pow, top_db, reinterpret_cast<cudaStream_t>(stream));
}

static int encoder_geometry(const uitk_encoder_cfg* cfg, int64_t B, int64_t T, int target_length, int64_t* rows) {
  UITK_REQUIRE(cfg, UITK_EINVAL, "null cfg");
  UITK_REQUIRE(B >= 0, UITK_EINVAL, "negative batch");
  UITK_REQUIRE(target_length >= 16 && target_length <= 16 * cfg->grid_t + 15, UITK_EINVAL,
               "target_length %d incompatible with time_pos_embed length %d", target_length, cfg->grid_t);
  UITK_REQUIRE(T >= 16, UITK_EINVAL, "need at least 16 frames (2400 samples) for one patch, got %lld", (long long)T);
  *rows = B * crops_for(T, target_length) * tokens_total_for(*cfg, T, target_length);
  return UITK_OK;
}

size_t uitk_encoder_workspace_bytes(const uitk_encoder_cfg* cfg, int64_t B, int64_t T, int target_length) {
  int64_t rows = 0;
  if (encoder_geometry(cfg, B, T, target_length, &rows) != UITK_OK) return 0;
  const int64_t RR = B * crops_for(T, target_length);
  if (tc_config(*cfg)) return encoder_tc_workspace_bytes(RR, RR * 24) + 256;       // tensor-core megakernel: 24 row slots per clip-crop
  return encoder_fp32_workspace_bytes(*cfg, RR * tokens_total_for(*cfg, T, target_length)) + 256;
}

size_t uitk_encoder_tokens_offset(const uitk_encoder_cfg* cfg, int64_t B, int64_t T, int target_length) {
  (void)cfg; (void)B; (void)T; (void)target_length;
  return 0;   // x[rows][128] is the first workspace region in both precisions
}

int uitk_encoder(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T,
                 int target_length, int eval_avg, const uint32_t* d_max_pow, float* d_probs, void* d_workspace,
                 size_t workspace_bytes, void* stream) {
  UITK_REQUIRE(cfg && d_encoder_blob && d_db && d_max_pow && d_probs && d_workspace, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(eval_avg == 0 || eval_avg == 1, UITK_EINVAL, "eval_avg must be 0 (mean) or 1 (max)");
  UITK_REQUIRE(aligned(d_workspace, 256) && aligned(d_encoder_blob, 256) && aligned(d_db, 4) && aligned(d_probs, 4), UITK_EALIGN,
               "misaligned pointer (workspace and blob need 256-byte alignment)");
  int64_t rows = 0;
  int rc = encoder_geometry(cfg, B, T, target_length, &rows);
  if (rc != UITK_OK) return rc;
  if (B == 0) return UITK_OK;
  rc = check_arch();
  if (rc != UITK_OK) return rc;
  EncoderArgs a{cfg, d_encoder_blob, d_db, B, T, target_length, eval_avg, d_max_pow, d_probs,
                d_workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream), g_debug_taps};
  UITK_REQUIRE(cfg->precision == UITK_PREC_FP32 || cfg->precision == UITK_PREC_BF16, UITK_EINVAL, "unknown precision %d", cfg->precision);
  return tc_config(*cfg) ? run_encoder_tc(a) : run_encoder_fp32(a);
}

int uitk_encoder_fixup(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T,
                       int target_length, int eval_avg, const uint32_t* d_max_pow, const uint32_t* d_max_used,
                       const uint32_t* d_min_pow, float* d_probs, void* d_workspace, size_t workspace_bytes, void* stream) {
  UITK_REQUIRE(cfg && d_encoder_blob && d_db && d_max_pow && d_max_used && d_min_pow && d_probs && d_workspace, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(eval_avg == 0 || eval_avg == 1, UITK_EINVAL, "eval_avg must be 0 (mean) or 1 (max)");
  UITK_REQUIRE(tc_config(*cfg), UITK_EINVAL, "uitk_encoder_fixup needs the tensor-core configuration (bf16, BNeckAttention, ReLU, mean pooling)");
  UITK_REQUIRE(aligned(d_workspace, 256) && aligned(d_encoder_blob, 256) && aligned(d_db, 4) && aligned(d_probs, 4), UITK_EALIGN,
               "misaligned pointer (workspace and blob need 256-byte alignment)");
  int64_t rows = 0;
  int rc = encoder_geometry(cfg, B, T, target_length, &rows);
  if (rc != UITK_OK) return rc;
  if (B == 0) return UITK_OK;
  rc = check_arch();
  if (rc != UITK_OK) return rc;
  EncoderArgs a{cfg, d_encoder_blob, d_db, B, T, target_length, eval_avg, d_max_pow, d_probs,
                d_workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream), 0};
  a.cond_used = d_max_used; a.cond_min = d_min_pow;
  return run_encoder_tc(a);
}

static int features_geometry(const uitk_encoder_cfg* cfg, int64_t B, int64_t T) {
  UITK_REQUIRE(cfg, UITK_EINVAL, "null cfg");
  UITK_REQUIRE(B >= 0, UITK_EINVAL, "negative batch");
  UITK_REQUIRE(T >= 16 && T <= 16 * cfg->grid_t + 15, UITK_EINVAL,
               "forward_features takes 16..%d frames (time_pos_embed has %d entries), got %lld", 16 * cfg->grid_t + 15, cfg->grid_t, (long long)T);
  return UITK_OK;
}

size_t uitk_forward_features_workspace_bytes(const uitk_encoder_cfg* cfg, int64_t B, int64_t T) {
  if (features_geometry(cfg, B, T) != UITK_OK) return 0;
  const int target = 16 * cfg->grid_t + 15;
  if (tc_config(*cfg)) return encoder_tc_workspace_bytes(B, B * 24) + 256;
  return encoder_fp32_workspace_bytes(*cfg, B * tokens_total_for(*cfg, T, target)) + 256;
}

int uitk_forward_features(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_spec, int64_t B, int64_t T,
                          float* d_tokens, void* d_workspace, size_t workspace_bytes, void* stream) {
  UITK_REQUIRE(cfg && d_encoder_blob && d_spec && d_tokens && d_workspace, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(aligned(d_workspace, 256) && aligned(d_encoder_blob, 256) && aligned(d_spec, 4) && aligned(d_tokens, 16), UITK_EALIGN,
               "misaligned pointer (workspace and blob need 256-byte alignment, tokens 16)");
  int rc = features_geometry(cfg, B, T);
  if (rc != UITK_OK) return rc;
  if (B == 0) return UITK_OK;
  rc = check_arch();
  if (rc != UITK_OK) return rc;
  static const uint32_t* no_word = nullptr;
  EncoderArgs a{cfg, d_encoder_blob, d_spec, B, T, 16 * cfg->grid_t + 15, 0, no_word, nullptr,
                d_workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream), 0};
  a.features_out = d_tokens;
  return tc_config(*cfg) ? run_encoder_tc(a) : run_encoder_fp32(a);
}

int uitk_forward_head(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_tokens, int64_t B, int n_tokens,
                      float* d_probs, void* stream) {
  UITK_REQUIRE(cfg && d_encoder_blob && d_tokens && d_probs, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(B >= 0, UITK_EINVAL, "negative batch");
  UITK_REQUIRE(n_tokens >= 1 && n_tokens <= UITK_MAX_TOKENS + 1, UITK_EINVAL, "n_tokens %d outside [1, %d]", n_tokens, UITK_MAX_TOKENS + 1);
  UITK_REQUIRE(cfg->pooling != UITK_POOL_DM || n_tokens % 4 == 0, UITK_EINVAL, "pooling='dm' needs 4 * t tokens, got %d", n_tokens);
  UITK_REQUIRE(aligned(d_tokens, 16) && aligned(d_encoder_blob, 256) && aligned(d_probs, 4), UITK_EALIGN, "misaligned pointer");
  if (B == 0) return UITK_OK;
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return run_forward_head(*cfg, d_encoder_blob, d_tokens, B, n_tokens, d_probs, reinterpret_cast<cudaStream_t>(stream));
}

int uitk_init_bn(const uitk_encoder_cfg* cfg, const void* d_encoder_blob, const float* d_db, int64_t B, int64_t T, float* d_out,
                 void* stream) {
  UITK_REQUIRE(cfg && d_encoder_blob && d_db && d_out, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(B >= 0 && T >= 1, UITK_EINVAL, "bad shape");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  const EncoderLayout lay = make_encoder_layout(*cfg);
  const float* W = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(d_encoder_blob) + sizeof(BlobHeader));
  return launch_init_bn(d_db, B, T, W + lay.bn_scale, W + lay.bn_shift, d_out, reinterpret_cast<cudaStream_t>(stream));
}

void uitk_debug_taps(int enable) { g_debug_taps = enable; }

int uitk_debug_read_trace(long long* host_out, int which, int n) {
  UITK_REQUIRE(host_out, UITK_EINVAL, "null pointer");
  return read_encoder_trace(host_out, which, n);
}

int uitk_selftest_umma(const float* d_A, const void* d_B_packed, const float* d_C_init, float* d_C, int N, int K, void* stream) {
  UITK_REQUIRE(d_A && d_B_packed && d_C, UITK_EINVAL, "null pointer");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return run_umma_selftest(d_A, d_B_packed, d_C_init, d_C, N, K, 0, reinterpret_cast<cudaStream_t>(stream));
}

int uitk_selftest_umma_ts(const float* d_A, const void* d_B_packed, const float* d_C_init, float* d_C, int N, int K, void* stream) {
  UITK_REQUIRE(d_A && d_B_packed && d_C, UITK_EINVAL, "null pointer");
  int rc = check_arch();
  if (rc != UITK_OK) return rc;
  return run_umma_selftest(d_A, d_B_packed, d_C_init, d_C, N, K, 1, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
