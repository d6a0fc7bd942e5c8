// Shared internals of libuitk (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/uitk.h"

namespace uitk {

void set_error(const char* fmt, ...);
void count_launches(int n);   // process-wide counter behind uitk_kernel_launches()

#define UITK_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      uitk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return UITK_ECUDA;                                                             \
    }                                                                                \
  } while (0)

#define UITK_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      uitk::set_error(__VA_ARGS__);     \
      return (code);                    \
    }                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// Front-end constant blob (device layout).  All fields 4 bytes; header first.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxMelWeights = 1792;   // packed filterbank entries kept in shared memory (HTK/64: 1024 incl. padding)

struct FrontendBlob {
  int magic;                 // 'UFE1'
  int n_weights;             // packed filterbank entries
  int pad0, pad1;
  float window[512];         // front_end.0.spectrogram.window
  float2 tw256[256];         // [k1*16 + lane] = exp(-2*pi*i*lane*k1/256)
  float2 tw512[256];         // exp(-2*pi*i*k/512)
  int mel_lo[64];            // first frequency bin of mel bin m's range, rounded down to a multiple of 4
  int mel_iters[4];          // float4 steps for the mel bins {16q .. 16q+15} (max range of the group / 4)
  int mel_qoff[4];           // float offset of group q's weights in mel_w
  int pad2[8];
  // group q, step i, lane j (mel bin 16q+j): 4 weights for bins lo+4i .. lo+4i+3 at mel_w[qoff + (i*16 + j)*4]; the 16
  // lanes of a frame group read 16 consecutive float4 (conflict-free), shorter ranges are zero padded
  float mel_w[kMaxMelWeights];
};
constexpr int kFrontendMagic = 0x55464531;

// ---------------------------------------------------------------------------------------------------------
// Encoder blob (fp32 section).  Offsets in floats from the start of the fp32 section.
// All Linear weights are stored TRANSPOSED, Wt[K][N] (K-major rows), so that K-slabs are contiguous.
// ---------------------------------------------------------------------------------------------------------
struct BlockOffsets {
  size_t ln1_w, ln1_b, qkv_wt, qkv_b, proj_wt, proj_b, ln2_w, ln2_b, fc1_wt, fc1_b, fc2_wt, fc2_b;
};

struct EncoderLayout {
  int depth, outputdim, grid_t;
  size_t bn_scale, bn_shift;          // [64] folded eval BatchNorm
  size_t patch_wt, patch_b;           // [256][128], [128]
  size_t time_pos, freq_pos;          // [grid_t][128], [4][128]
  size_t norm_w, norm_b, hln_w, hln_b;
  size_t cb_final;                    // [128] sum of all proj/fc2 biases (kept for blob compatibility; unused)
  size_t pos_tab;                     // [24][128] conv bias + time_pos[tok % 6] + freq_pos[tok / 6] (tensor-core path, 24-token crops)
  size_t head_wt, head_b;             // [128][outputdim_padded], [outputdim_padded]
  int outputdim_padded;
  size_t blocks;                      // first block
  size_t block_stride;
  BlockOffsets blk;                   // offsets relative to the block start
  size_t total_floats;
};

EncoderLayout make_encoder_layout(int depth, int outputdim, int grid_t);

struct BlobHeader {
  int magic;        // 'UEN1'
  int depth, outputdim, grid_t, precision;
  int reserved[3];
  unsigned long long fp32_offset;   // bytes from blob start
  unsigned long long bf16_offset;   // bytes from blob start (0 if absent)
  unsigned long long total_bytes;
  unsigned long long reserved2;
  unsigned char pad[192];           // header is 256 bytes so that the sections keep 256-byte alignment
};
static_assert(sizeof(BlobHeader) == 256, "BlobHeader must be 256 bytes");
constexpr int kEncoderMagic = 0x55454e31;

inline int crops_for(int64_t T, int target) { return T <= target ? 1 : (int)((T + target - 1) / target); }
inline int time_patches_for(int64_t T, int target) {
  int64_t tc = T <= target ? T : target;
  return (int)((tc - 16) / 16 + 1);
}

// kernels (launch wrappers), defined in the .cu files
int launch_logmel(const float* wav, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db,
                  uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s);
int launch_logmel_i16(const int16_t* pcm, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db,
                      uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s);
int launch_logmel_frames(const float* wav, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db, int64_t t0, int64_t tn,
                         int64_t out_bs, int64_t out_ms, uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s);
int launch_window_gather(const float* G, int64_t U, float* dbw, int64_t W, int Tw, int r, cudaStream_t s);
int launch_clamp_db(float* db, int64_t n, const uint32_t* max_pow, float top_db, cudaStream_t s);

struct EncoderArgs {
  const uitk_encoder_cfg* cfg;
  const void* blob;
  const float* db;
  int64_t B, T;
  int target_length, eval_avg;
  const uint32_t* max_pow;
  float* probs;
  void* ws;
  size_t ws_bytes;
  cudaStream_t stream;
  int debug_taps;
};
int run_encoder_fp32(const EncoderArgs& a);
int run_encoder_tc(const EncoderArgs& a);
int read_encoder_trace(long long* host_out, int which, int n);
int run_umma_selftest(const float* A, const void* Bp, const float* Cinit, float* C, int N, int K, int a_in_tmem, cudaStream_t s);
size_t encoder_tc_bf16_section_bytes(int depth);
size_t encoder_tc_block_bytes();
size_t encoder_tc_workspace_bytes(int64_t clip_crops, int64_t rows);
size_t encoder_fp32_workspace_bytes(int64_t rows);   // rows = B * crops * tokens

}  // namespace uitk
