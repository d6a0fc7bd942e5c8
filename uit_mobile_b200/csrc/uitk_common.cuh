// Shared internals of libuitk (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
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
// Mel projection on the tensor cores (mma.sync m16n8k8 tf32, fp32 accumulate): the 64 mel bins are 8 octets, the 257 frequency
// bins 33 groups of 8; octet o multiplies the groups that hold its non-zero filterbank entries.  For every (octet, group) the
// blob carries the B fragment of that 8x8 weight block in fp32 (the kernel splits it into tf32 hi + lo): entry [lane] = (b0, b1) with
// b0 = W[8g + 2*(lane%4)][8o + lane/4], b1 = W[8g + 2*(lane%4) + 1][8o + lane/4], W = 0.25 * fb (the kernel keeps the power
// spectrum as 4 |X|^2).  The blocks are dealt to the 8 warps of a CTA as SEGMENTS (a run of groups of one octet) so that every
// warp has about the same work: a heavy octet is cut in two or three, the earlier parts ("producers") hand their partial sums
// to the last ("owner") through shared memory, added in a fixed order (own + slot 0 + slot 1): the sum does not depend on timing.
constexpr int kMelOctets = 8;
constexpr int kMelGroups = 33;                 // ceil(257 / 8); group 32 holds the Nyquist bin alone
constexpr int kMelWarpBlocks = 40;             // blocks per warp (dense filterbank: 264 / 8 = 33)
constexpr int kMelSlots = 8;                   // partial-sum slots (at most 7 cuts between 8 warps)
constexpr int kMelSmemBlocks = 48;             // weight blocks the kernel keeps in shared memory (HTK/64: 40 + the zero block); a
                                               // denser filterbank is read from global memory instead
enum { kMelWhole = 0, kMelProducer = 1, kMelOwner = 2 };

// One weight block of a warp's list.  x = group | fin << 8 | role << 9 | octet << 11 | aux << 14, y = block index in mel_frag.
// fin: the block ends a run; role / octet / aux say what to do with the sums then:
//   whole     write the octet's dB values
//   producer  aux = slot | producers << 3: store the partial sums to the slot, arrive on named barrier 1 + octet
//   owner     aux = producers | slot0 << 2 | slot1 << 5: wait on the barrier, add slot0 (+ slot1), write the dB values
struct MelBlk {
  int x, y;
};

struct FrontendBlob {
  int magic;                 // 'UFE4'
  int n_blocks;              // (octet, group) weight blocks that follow the fixed part (+ one all-zero block at index n_blocks)
  int pad0, pad1;
  float window[512];         // front_end.0.spectrogram.window
  float2 tw256[256];         // [k1*16 + lane] = exp(-2*pi*i*lane*k1/256)
  float2 tw512[256];         // exp(-2*pi*i*k/512)
  int mel_nblk[kMelOctets];  // blocks of warp w (producer runs first, owner runs last: an owner never waits for a warp that waits)
  MelBlk mel_blk[kMelOctets][kMelWarpBlocks + 1];   // entry [mel_nblk] repeats the last one (the kernel reads one ahead)
  float2 mel_frag[1];        // [n_blocks + 1][32 lanes], variable length; the blocks of an octet are consecutive, groups ascending
};
inline size_t frontend_blob_bytes(int n_blocks) { return offsetof(FrontendBlob, mel_frag) + sizeof(float2) * 32 * (size_t)(n_blocks + 1); }
constexpr int kFrontendMagic = 0x55464535;

// ---------------------------------------------------------------------------------------------------------
// Encoder blob (fp32 section).  Offsets in floats from the start of the fp32 section.
// All Linear weights are stored TRANSPOSED, Wt[K][N] (K-major rows), so that K-slabs are contiguous.
// ---------------------------------------------------------------------------------------------------------
struct BlockOffsets {
  size_t ln1_w, ln1_b, qkv_wt, qkv_b, proj_wt, proj_b, ln2_w, ln2_b, fc1_wt, fc1_b, fc2_wt, fc2_b;
};

struct EncoderLayout {
  int depth, outputdim, grid_t;
  int qkv_n, inner;                   // 96 / 32 (BNeckAttention) or 384 / 128 (Attention)
  size_t bn_scale, bn_shift;          // [64] folded eval BatchNorm
  size_t patch_wt, patch_b;           // [256][128], [128]
  size_t time_pos, freq_pos;          // [grid_t][128], [4][128]
  size_t norm_w, norm_b, hln_w, hln_b;
  size_t cb_final;                    // [128] sum of all proj/fc2 biases (kept for blob compatibility; unused)
  size_t pos_tab;                     // [24][128] conv bias + time_pos[tok % 6] + freq_pos[tok / 6] (tensor-core path, 24-token crops)
  size_t cls_row;                     // [128] cls_token + token_pos_embed (pooling='token', uit.py:389-392)
  size_t ident_scale, ident_shift;    // [64] ones / zeros: "input already normalised" (uitk_forward_features)
  size_t head_wt, head_b;             // [128][outputdim_padded], [outputdim_padded]
  size_t head_frag;                   // head weight as mma.m16n8k8 tf32 B fragments, hi + lo: [outputdim_padded / 64][16 k-steps][8 n-tiles][32 lanes] float4
  int outputdim_padded;
  size_t blocks;                      // first block
  size_t block_stride;
  BlockOffsets blk;                   // offsets relative to the block start
  size_t total_floats;
};

EncoderLayout make_encoder_layout(const uitk_encoder_cfg& cfg);
inline bool tc_config(const uitk_encoder_cfg& c) {   // what the tensor-core megakernel implements
  return c.precision == UITK_PREC_BF16 && c.attention == UITK_ATTN_BNECK && c.act == UITK_ACT_RELU && c.pooling == UITK_POOL_MEAN;
}

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
int launch_clamp_db(float* db, int64_t n, const uint32_t* max_pow, const uint32_t* min_pow, float top_db, cudaStream_t s);

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
  // uitk_forward_features: db is an already normalised spectrogram (no clamp, no BatchNorm), no crops, and instead of the
  // head the final-LayerNorm tokens [B][tokens_total][128] are written to features_out
  float* features_out = nullptr;
  // uitk_encoder_fixup: {max_pow (true), max_used, min_pow}; every kernel returns at once unless fixup_needed()
  const uint32_t* cond_used = nullptr;
  const uint32_t* cond_min = nullptr;
};
// Device-side decision of uitk_encoder_fixup (see include/uitk.h): same dB map as the kernels' cutoff.
__device__ __forceinline__ bool fixup_needed(const uint32_t* max_true, const uint32_t* max_used, const uint32_t* min_pow) {
  const uint32_t mt = *max_true, mu = *max_used;
  if (mt == mu) return false;
  const float cutoff = 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(mt), 1e-10f)) - 120.f;
  const float min_db = 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(*min_pow), 1e-10f));
  return min_db < cutoff;
}
int launch_final_ln(const float* x, int64_t RR, int slots, int t_n, int slot_tn, const float* norm_w, const float* norm_b, float* out,
                    cudaStream_t s);
int launch_init_bn(const float* db, int64_t B, int64_t T, const float* scale, const float* shift, float* out, cudaStream_t s);
int run_forward_head(const uitk_encoder_cfg& cfg, const void* blob, const float* tokens, int64_t B, int n_tokens, float* probs,
                     cudaStream_t s);
int run_encoder_fp32(const EncoderArgs& a);
int run_encoder_tc(const EncoderArgs& a);
int read_encoder_trace(long long* host_out, int which, int n);
int run_umma_selftest(const float* A, const void* Bp, const float* Cinit, float* C, int N, int K, int a_in_tmem, cudaStream_t s);
size_t peer_words_slot_bytes();
int launch_words_publish(uint32_t* my_slot, const uint32_t* word, uint32_t epoch, cudaStream_t s);
int launch_words_collect(const uint32_t* const* peer_slots, int n, uint32_t epoch, uint32_t* out, cudaStream_t s);
size_t encoder_tc_bf16_section_bytes(int depth);
size_t encoder_tc_block_bytes();
size_t encoder_tc_workspace_bytes(int64_t clip_crops, int64_t rows);
size_t encoder_fp32_workspace_bytes(const uitk_encoder_cfg& cfg, int64_t rows);   // rows = B * crops * tokens_total
inline int tokens_total_for(const uitk_encoder_cfg& c, int64_t T, int target) {
  return 4 * time_patches_for(T, target) + (c.pooling == UITK_POOL_TOKEN ? 1 : 0);
}

}  // namespace uitk
