// Host-side packing of state_dict tensors into the kernels' device layouts (no CUDA calls here).
#include <math.h>
#include <stdarg.h>

#include <atomic>
#include <string>
#include <vector>

#include "uitk_common.cuh"

namespace uitk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long get_launches() { return g_launches.load(std::memory_order_relaxed); }

static size_t take(size_t& cur, size_t n) {
  const size_t off = cur;
  cur += (n + 63) / 64 * 64;   // 256-byte alignment of every tensor
  return off;
}

EncoderLayout make_encoder_layout(const uitk_encoder_cfg& cfg) {
  const int depth = cfg.depth, outputdim = cfg.outputdim, grid_t = cfg.grid_t;
  EncoderLayout l{};
  l.depth = depth; l.outputdim = outputdim; l.grid_t = grid_t;
  l.qkv_n = cfg.attention == UITK_ATTN_FULL ? 384 : 96;
  l.inner = cfg.attention == UITK_ATTN_FULL ? 128 : 32;
  l.outputdim_padded = (outputdim + 127) / 128 * 128;   // zero-padded to whole 128-column GEMM tiles
  size_t cur = 0;
  l.bn_scale = take(cur, 64); l.bn_shift = take(cur, 64);
  l.patch_wt = take(cur, 256 * 128); l.patch_b = take(cur, 128);
  l.time_pos = take(cur, (size_t)grid_t * 128); l.freq_pos = take(cur, 4 * 128);
  l.norm_w = take(cur, 128); l.norm_b = take(cur, 128);
  l.hln_w = take(cur, 128); l.hln_b = take(cur, 128);
  l.cb_final = take(cur, 128);
  l.pos_tab = take(cur, 24 * 128);
  l.cls_row = take(cur, 128);
  l.ident_scale = take(cur, 64); l.ident_shift = take(cur, 64);
  l.head_wt = take(cur, (size_t)128 * l.outputdim_padded); l.head_b = take(cur, l.outputdim_padded);
  l.head_frag = take(cur, (size_t)128 * l.outputdim_padded * 2);
  l.blocks = cur;
  size_t b = 0;
  l.blk.ln1_w = take(b, 128); l.blk.ln1_b = take(b, 128);
  l.blk.qkv_wt = take(b, (size_t)128 * l.qkv_n); l.blk.qkv_b = take(b, l.qkv_n);
  l.blk.proj_wt = take(b, (size_t)l.inner * 128); l.blk.proj_b = take(b, 128);
  l.blk.ln2_w = take(b, 128); l.blk.ln2_b = take(b, 128);
  l.blk.fc1_wt = take(b, 128 * 384); l.blk.fc1_b = take(b, 384);
  l.blk.fc2_wt = take(b, 384 * 128); l.blk.fc2_b = take(b, 128);
  l.block_stride = b;
  l.total_floats = cur + b * (size_t)depth;
  return l;
}

// Order of h_tensors[] (state_dict keys).  The front-end buffers and num_batches_tracked are not passed;
// cls_token / token_pos_embed are dead for pooling='mean' | 'dm' (Q4) and live for pooling='token'.
static const char* kFixedNames[] = {
    "init_bn.1.weight", "init_bn.1.bias", "init_bn.1.running_mean", "init_bn.1.running_var",
    "patch_embed.proj.weight", "patch_embed.proj.bias", "time_pos_embed", "freq_pos_embed",
    "norm.weight", "norm.bias", "outputlayer.0.weight", "outputlayer.0.bias",
    "outputlayer.1.weight", "outputlayer.1.bias", "cls_token", "token_pos_embed"};
static const char* kBlockNames[] = {
    "norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
    "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"};
constexpr int kNumFixed = 16, kNumBlock = 12;

static void transpose_into(float* dst, const float* w, int out_f, int in_f, int ld_dst) {
  // torch Linear weight [out_f][in_f] -> Wt[in_f][ld_dst]
  for (int o = 0; o < out_f; ++o)
    for (int k = 0; k < in_f; ++k) dst[(size_t)k * ld_dst + o] = w[(size_t)o * in_f + k];
}

// fp32 -> bf16 bits, round to nearest even (matches __float2bfloat16_rn for finite values)
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// W[(n0+n)*ldw + k0+k], n<N, k<K  ->  K-major core-matrix layout: dst[((k/8)*N + n)*8 + k%8]   (tc_ptx.cuh)
static void pack_kmajor(uint16_t* dst, const float* W, int ldw, int n0, int N, int k0, int K) {
  for (int k8 = 0; k8 < K / 8; ++k8)
    for (int n = 0; n < N; ++n)
      for (int i = 0; i < 8; ++i) dst[((size_t)k8 * N + n) * 8 + i] = f2bf(W[(size_t)(n0 + n) * ldw + k0 + k8 * 8 + i]);
}

static float bf2f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// [N x 8] bf16 tile (one K-major k-group): row n = (hi, mid, lo, 0, 0, 0, 0, 0) with hi + mid + lo == bias[n] to ~2^-24
static void pack_bias_tile(uint16_t* dst, const float* bias, int N) {
  for (int n = 0; n < N; ++n) {
    const uint16_t hi = f2bf(bias[n]);
    const float r1 = bias[n] - bf2f(hi);
    const uint16_t mid = f2bf(r1);
    const uint16_t lo = f2bf(r1 - bf2f(mid));
    uint16_t* d = dst + (size_t)n * 8;
    d[0] = hi; d[1] = mid; d[2] = lo;
    for (int i = 3; i < 8; ++i) d[i] = 0;
  }
}

static size_t fp32_section_bytes(const EncoderLayout& l) { return (l.total_floats * sizeof(float) + 1023) / 1024 * 1024; }

}  // namespace uitk

using namespace uitk;

extern "C" {

int uitk_version(void) { return UITK_VERSION; }
const char* uitk_last_error(void) { return get_error(); }
uint64_t uitk_kernel_launches(void) { return get_launches(); }

int64_t uitk_num_frames(int64_t L) { return 1 + L / UITK_HOP; }
int uitk_num_crops(int64_t T, int target_length) { return crops_for(T, target_length); }
int uitk_tokens_per_crop(int64_t T, int target_length) { return 4 * time_patches_for(T, target_length); }
int uitk_tokens_total(const uitk_encoder_cfg* cfg, int64_t T, int target_length) { return cfg ? tokens_total_for(*cfg, T, target_length) : 0; }

// (octet, 8-bin group) blocks of the filterbank that hold a non-zero entry: first group / group count per octet
static int mel_block_ranges(const float* h_fb, int* glo, int* gcnt) {
  int n = 0;
  for (int o = 0; o < kMelOctets; ++o) {
    int lo = -1, hi = -1;
    for (int k = 0; k < UITK_N_FREQS; ++k)
      for (int m = 8 * o; m < 8 * o + 8; ++m)
        if (h_fb[(size_t)k * UITK_N_MELS + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
    glo[o] = lo < 0 ? 0 : lo / 8;
    gcnt[o] = lo < 0 ? 0 : hi / 8 - lo / 8 + 1;
    n += gcnt[o];
  }
  return n;
}

size_t uitk_frontend_blob_bytes(const float* h_fb) {
  int glo[kMelOctets], gcnt[kMelOctets];
  return frontend_blob_bytes(h_fb ? mel_block_ranges(h_fb, glo, gcnt) : kMelOctets * kMelGroups);
}

int uitk_pack_frontend(const float* h_window, const float* h_fb, void* h_blob, size_t blob_bytes) {
  UITK_REQUIRE(h_window && h_fb && h_blob, UITK_EINVAL, "null pointer");
  int glo[kMelOctets], gcnt[kMelOctets];
  const int n_blocks = mel_block_ranges(h_fb, glo, gcnt);
  const size_t need = frontend_blob_bytes(n_blocks);
  UITK_REQUIRE(blob_bytes >= need, UITK_ENOSPACE, "front-end blob needs %zu bytes", need);
  FrontendBlob* fb = reinterpret_cast<FrontendBlob*>(h_blob);
  memset(fb, 0, need);
  fb->magic = kFrontendMagic;
  fb->n_blocks = n_blocks;
  memcpy(fb->window, h_window, sizeof(float) * 512);
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < 256; ++j) {
    // tw256 is stored as the table the kernel indexes: entry [k1*16 + lane] = exp(-2*pi*i*lane*k1/256), so that the
    // 16 lanes of a frame group read 16 consecutive entries (bank-conflict free)
    const int k1 = j >> 4, ln = j & 15;
    fb->tw256[j] = make_float2((float)cos(two_pi * (ln * k1) / 256.0), (float)-sin(two_pi * (ln * k1) / 256.0));
    fb->tw512[j] = make_float2((float)cos(two_pi * j / 512.0), (float)-sin(two_pi * j / 512.0));
  }
  int boff[kMelOctets];
  int b = 0;
  for (int o = 0; o < kMelOctets; ++o) {
    boff[o] = b;
    for (int u = 0; u < gcnt[o]; ++u, ++b)
      for (int lane = 0; lane < 32; ++lane) {
        const int tig = lane & 3, gid = lane >> 2, k0 = 8 * (glo[o] + u) + 2 * tig, m = 8 * o + gid;
        float w[2];
        for (int e = 0; e < 2; ++e)       // x 1/4: the kernel keeps the power spectrum as 4 |X|^2 (exact scaling)
          w[e] = k0 + e < UITK_N_FREQS ? 0.25f * h_fb[(size_t)(k0 + e) * UITK_N_MELS + m] : 0.f;
        fb->mel_frag[(size_t)b * 32 + lane] = make_float2(w[0], w[1]);
      }
  }
  // Deal the blocks to the 8 warps: walk the octets (heaviest end first) and cut the line of blocks into 8 runs of about equal
  // cost.  Cost model (instructions per warp and round): a block ~ kBlk, the dB epilogue of an octet ~ kOut, a partial-sum
  // hand-off ~ kXch.  A cut inside an octet makes the run that ends a warp's work the OWNER (it adds the partial sums and writes
  // the dB values, last thing it does) and the continuation in the next warp(s) a PRODUCER (first thing that warp does), so an
  // owner practically never waits.  At most two producers per octet.
  constexpr int kBlk = 24, kOut = 30, kXch = 12, kStage = 40;
  struct Task { int oct, role, g_lo, g_cnt, aux; };         // aux: producer -> its slot
  std::vector<Task> per_warp[kMelOctets];
  int total = 0;
  for (int o = 0; o < kMelOctets; ++o) total += kBlk * gcnt[o] + kOut;
  bool ok = false;
  for (int target = total / kMelOctets; target <= total + kOut && !ok; target += 4) {
    for (int w = 0; w < kMelOctets; ++w) per_warp[w].clear();
    int w = 0, load = kStage;                               // warp 0 first issues the next round's staging copy
    ok = true;
    for (int o = kMelOctets - 1; o >= 0 && ok; --o) {
      int remaining = gcnt[o], g = glo[o], producers = 0;
      bool owner_placed = false;
      for (;;) {
        const bool room = (int)per_warp[w].size() < 6;                          // runs per warp (8 octets: never reached)
        const int fixed = owner_placed ? kXch : kOut;                       // cost of this part besides its blocks
        if (room && load + kBlk * remaining + fixed <= target) {            // the rest of the octet fits here
          per_warp[w].push_back({o, owner_placed ? kMelProducer : kMelWhole, g, remaining, producers});
          load += kBlk * remaining + fixed;
          break;
        }
        int n = room ? (target - load - fixed - (owner_placed ? 0 : kXch)) / kBlk : 0;
        if (n > remaining - 1) n = remaining - 1;                           // something must be left for the next warp
        if (n >= 1 && (!owner_placed || producers < 1)) {
          per_warp[w].push_back({o, owner_placed ? kMelProducer : kMelOwner, g, n, producers});
          producers += owner_placed;
          owner_placed = true;
          remaining -= n; g += n;
        } else if (load == 0) {                             // an empty warp cannot take it and it cannot be cut (further)
          ok = false; break;
        }
        if (++w == kMelOctets) { ok = false; break; }
        load = 0;
      }
    }
  }
  if (!ok)                                                   // cannot happen (target = total always fits); keep a safe schedule anyway
    for (int w = 0; w < kMelOctets; ++w) { per_warp[w].clear(); per_warp[w].push_back({w, kMelWhole, glo[w], gcnt[w], 0}); }
  // emit the per-warp block lists
  int prod_slot[kMelOctets][2], n_prod[kMelOctets] = {0}, next_slot = 0;
  for (int w = 0; w < kMelOctets; ++w)
    for (const Task& t : per_warp[w])
      if (t.role == kMelProducer) prod_slot[t.oct][n_prod[t.oct]++] = next_slot++;
  UITK_REQUIRE(next_slot <= kMelSlots, UITK_EINVAL, "mel schedule needs %d hand-off slots (max %d)", next_slot, kMelSlots);
  for (int w = 0; w < kMelOctets; ++w) {
    int n = 0;
    for (int pass = 0; pass < 3; ++pass)                     // producers, whole octets, owners
      for (const Task& t : per_warp[w]) {
        if (t.role != (pass == 0 ? kMelProducer : pass == 1 ? kMelWhole : kMelOwner)) continue;
        int aux = 0;
        if (t.role == kMelProducer) aux = prod_slot[t.oct][t.aux] | (n_prod[t.oct] << 3);
        if (t.role == kMelOwner) aux = n_prod[t.oct] | (prod_slot[t.oct][0] << 2) | ((n_prod[t.oct] > 1 ? prod_slot[t.oct][1] : 0) << 5);
        const int cnt = t.g_cnt > 0 ? t.g_cnt : 1;           // an octet without weights: one all-zero block
        UITK_REQUIRE(n + cnt <= kMelWarpBlocks, UITK_EINVAL, "mel schedule: more than %d blocks for one warp", kMelWarpBlocks);
        for (int u = 0; u < cnt; ++u) {
          const int fin = u == cnt - 1;
          const int g = t.g_cnt > 0 ? t.g_lo + u : 0, blk = t.g_cnt > 0 ? boff[t.oct] + (t.g_lo - glo[t.oct]) + u : n_blocks;
          fb->mel_blk[w][n++] = MelBlk{g | (fin << 8) | (fin ? (t.role << 9) | (t.oct << 11) | (aux << 14) : 0), blk};
        }
      }
    fb->mel_nblk[w] = n;
    if (n > 0) fb->mel_blk[w][n] = fb->mel_blk[w][n - 1];
  }
  return UITK_OK;
}

int uitk_encoder_num_tensors(int depth) { return kNumFixed + kNumBlock * depth; }

const char* uitk_encoder_tensor_name(int depth, int index) {
  static thread_local std::string name;
  if (index < 0 || index >= uitk_encoder_num_tensors(depth)) return nullptr;
  if (index < kNumFixed) return kFixedNames[index];
  const int i = (index - kNumFixed) / kNumBlock, j = (index - kNumFixed) % kNumBlock;
  name = "blocks." + std::to_string(i) + "." + kBlockNames[j];
  return name.c_str();
}

static int check_cfg(const uitk_encoder_cfg* cfg) {
  UITK_REQUIRE(cfg, UITK_EINVAL, "null cfg");
  UITK_REQUIRE(cfg->depth >= 1 && cfg->depth <= 64, UITK_EINVAL, "depth %d out of range", cfg->depth);
  UITK_REQUIRE(cfg->outputdim >= 1 && cfg->outputdim <= 768, UITK_EINVAL, "outputdim %d out of range [1,768]", cfg->outputdim);
  UITK_REQUIRE(cfg->grid_t >= 1 && cfg->grid_t <= 6, UITK_EINVAL, "grid_t %d out of range [1,6]", cfg->grid_t);
  UITK_REQUIRE(cfg->precision == UITK_PREC_FP32 || cfg->precision == UITK_PREC_BF16, UITK_EINVAL, "bad precision %d", cfg->precision);
  UITK_REQUIRE(cfg->attention == UITK_ATTN_BNECK || cfg->attention == UITK_ATTN_FULL, UITK_EINVAL, "bad attention type %d", cfg->attention);
  UITK_REQUIRE(cfg->act == UITK_ACT_RELU || cfg->act == UITK_ACT_GELU, UITK_EINVAL, "bad activation %d", cfg->act);
  UITK_REQUIRE(cfg->pooling >= UITK_POOL_MEAN && cfg->pooling <= UITK_POOL_DM, UITK_EINVAL, "bad pooling %d", cfg->pooling);
  UITK_REQUIRE(cfg->reserved == 0, UITK_EINVAL, "cfg.reserved must be 0");
  return UITK_OK;
}

size_t uitk_encoder_blob_bytes(const uitk_encoder_cfg* cfg) {
  if (check_cfg(cfg) != UITK_OK) return 0;
  const EncoderLayout l = make_encoder_layout(*cfg);
  size_t n = sizeof(BlobHeader) + fp32_section_bytes(l);
  if (tc_config(*cfg)) n += encoder_tc_bf16_section_bytes(cfg->depth);
  return n;
}

int uitk_pack_encoder(const uitk_encoder_cfg* cfg, const float* const* t, void* h_blob, size_t blob_bytes) {
  int rc = check_cfg(cfg);
  if (rc != UITK_OK) return rc;
  UITK_REQUIRE(t && h_blob, UITK_EINVAL, "null pointer");
  const size_t need = uitk_encoder_blob_bytes(cfg);
  UITK_REQUIRE(blob_bytes >= need, UITK_ENOSPACE, "encoder blob needs %zu bytes, have %zu", need, blob_bytes);
  for (int i = 0; i < uitk_encoder_num_tensors(cfg->depth); ++i)
    UITK_REQUIRE(t[i], UITK_EINVAL, "tensor %d (%s) is null", i, uitk_encoder_tensor_name(cfg->depth, i));
  const EncoderLayout l = make_encoder_layout(*cfg);
  const bool tc = tc_config(*cfg);
  memset(h_blob, 0, need);
  BlobHeader* hdr = reinterpret_cast<BlobHeader*>(h_blob);
  hdr->magic = kEncoderMagic; hdr->depth = cfg->depth; hdr->outputdim = cfg->outputdim; hdr->grid_t = cfg->grid_t;
  hdr->precision = cfg->precision; hdr->fp32_offset = sizeof(BlobHeader); hdr->bf16_offset = 0; hdr->total_bytes = need;
  float* W = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(h_blob) + sizeof(BlobHeader));

  // eval BatchNorm folded the way ATen does it: alpha = w / sqrt(var + eps); beta = b - mean * alpha
  const float *bn_w = t[0], *bn_b = t[1], *bn_m = t[2], *bn_v = t[3];
  for (int m = 0; m < 64; ++m) {
    const float invstd = 1.f / sqrtf(bn_v[m] + 1e-5f);
    const float alpha = invstd * bn_w[m];
    W[l.bn_scale + m] = alpha;
    W[l.bn_shift + m] = bn_b[m] - bn_m[m] * alpha;
  }
  transpose_into(W + l.patch_wt, t[4], 128, 256, 128);                 // [128,1,16,16] -> [256][128], k = df*16+dt
  memcpy(W + l.patch_b, t[5], 128 * sizeof(float));
  for (int c = 0; c < 128; ++c) {
    for (int tt = 0; tt < cfg->grid_t; ++tt) W[l.time_pos + (size_t)tt * 128 + c] = t[6][(size_t)c * cfg->grid_t + tt];
    for (int f = 0; f < 4; ++f) W[l.freq_pos + (size_t)f * 128 + c] = t[7][(size_t)c * 4 + f];
  }
  memcpy(W + l.norm_w, t[8], 128 * sizeof(float)); memcpy(W + l.norm_b, t[9], 128 * sizeof(float));
  memcpy(W + l.hln_w, t[10], 128 * sizeof(float)); memcpy(W + l.hln_b, t[11], 128 * sizeof(float));
  transpose_into(W + l.head_wt, t[12], cfg->outputdim, 128, l.outputdim_padded);
  memcpy(W + l.head_b, t[13], cfg->outputdim * sizeof(float));
  {  // tensor-core head (head_tc_kernel): B fragments of Wt[k][n], tf32 hi + lo (the tensor core ignores the low 13 mantissa bits).
     // Lane (gid = lane / 4, tig = lane % 4) of k-step s, n-tile j, column tile ct holds k = 8s + 2 tig (+1), n = 64 ct + 8 j + gid.
    auto trunc = [](float f) { uint32_t u; memcpy(&u, &f, 4); u &= 0xffffe000u; memcpy(&f, &u, 4); return f; };
    float* F = W + l.head_frag;
    const float* Wt = W + l.head_wt;
    for (int ct = 0; ct < l.outputdim_padded / 64; ++ct)
      for (int s = 0; s < 16; ++s)
        for (int j = 0; j < 8; ++j)
          for (int lane = 0; lane < 32; ++lane) {
            const int k = 8 * s + 2 * (lane & 3), n = 64 * ct + 8 * j + (lane >> 2);
            const float w0 = Wt[(size_t)k * l.outputdim_padded + n], w1 = Wt[(size_t)(k + 1) * l.outputdim_padded + n];
            float* f = F + ((((size_t)ct * 16 + s) * 8 + j) * 32 + lane) * 4;
            f[0] = trunc(w0); f[1] = trunc(w1); f[2] = trunc(w0 - f[0]); f[3] = trunc(w1 - f[1]);
          }
  }
  for (int c = 0; c < 128; ++c) W[l.cls_row + c] = t[14][c] + t[15][c];      // cls_token + token_pos_embed (uit.py:390-391)
  for (int m = 0; m < 64; ++m) { W[l.ident_scale + m] = 1.f; W[l.ident_shift + m] = 0.f; }
  for (int i = 0; i < cfg->depth; ++i) {
    const float* const* b = t + kNumFixed + (size_t)i * kNumBlock;
    float* Wb = W + l.blocks + (size_t)i * l.block_stride;
    memcpy(Wb + l.blk.ln1_w, b[0], 128 * 4); memcpy(Wb + l.blk.ln1_b, b[1], 128 * 4);
    transpose_into(Wb + l.blk.qkv_wt, b[2], l.qkv_n, 128, l.qkv_n); memcpy(Wb + l.blk.qkv_b, b[3], l.qkv_n * 4);
    transpose_into(Wb + l.blk.proj_wt, b[4], 128, l.inner, 128); memcpy(Wb + l.blk.proj_b, b[5], 128 * 4);
    memcpy(Wb + l.blk.ln2_w, b[6], 128 * 4); memcpy(Wb + l.blk.ln2_b, b[7], 128 * 4);
    transpose_into(Wb + l.blk.fc1_wt, b[8], 384, 128, 384); memcpy(Wb + l.blk.fc1_b, b[9], 384 * 4);
    transpose_into(Wb + l.blk.fc2_wt, b[10], 128, 384, 128); memcpy(Wb + l.blk.fc2_b, b[11], 128 * 4);
  }
  // cb = running sum of all proj / fc2 biases (cb_final keeps the blob layout of earlier versions; no kernel reads it)
  std::vector<float> cb(128, 0.f);
  unsigned char* sec = reinterpret_cast<unsigned char*>(h_blob) + sizeof(BlobHeader) + fp32_section_bytes(l);
  if (tc) {
    hdr->bf16_offset = sizeof(BlobHeader) + fp32_section_bytes(l);
    for (int c = 0; c < 4; ++c)                                                          // patch weight, K quarters
      pack_kmajor(reinterpret_cast<uint16_t*>(sec + (size_t)c * 16384), t[4], 256, 0, 128, c * 64, 64);
  }
  for (int i = 0; i < cfg->depth; ++i) {
    const float* const* b = t + kNumFixed + (size_t)i * kNumBlock;
    if (tc) {
      unsigned char* blk = sec + 65536 + (size_t)i * encoder_tc_block_bytes();   // after the 4 x 16 KB patch chunks
      // LayerNorm affine folded into the consuming Linear: W' = W diag(gamma), b' = b + W beta (fp32, then bf16 for W')
      std::vector<float> wq(96 * 128), w1f(384 * 128), bq(96), b1f(384);
      for (int o = 0; o < 96; ++o) {
        float acc = b[3][o];
        for (int k = 0; k < 128; ++k) { wq[o * 128 + k] = b[2][o * 128 + k] * b[0][k]; acc += b[2][o * 128 + k] * b[1][k]; }
        bq[o] = acc;                                                            // qkv bias'
      }
      for (int o = 0; o < 384; ++o) {
        float acc = b[9][o];
        for (int k = 0; k < 128; ++k) { w1f[o * 128 + k] = b[8][o * 128 + k] * b[6][k]; acc += b[8][o * 128 + k] * b[7][k]; }
        b1f[o] = acc;                                                           // fc1 bias'
      }
      // Ring slots in the order the kernel consumes them (csrc/encoder_tc.cu).  Every Linear bias rides along as a
      // [N x 8] bf16 "bias tile" (hi, mid, lo split of the fp32 value in k = 0..2) that one extra MMA k-step multiplies
      // with a constant ones operand: the CUDA cores never touch a bias.
      uint16_t* w = reinterpret_cast<uint16_t*>(blk);
      pack_kmajor(w, wq.data(), 128, 0, 96, 0, 64); w += 96 * 64;               // slot: Wqkv' [96][128] K half 0 ...
      pack_bias_tile(w, bq.data(), 96); w += 96 * 8;                            //       ... + qkv bias' tile
      pack_kmajor(w, wq.data(), 128, 0, 96, 64, 64); w += 96 * 64;              // slot: Wqkv' K half 1
      pack_kmajor(w, b[4], 32, 0, 128, 0, 32); w += 128 * 32;                   // slot: Wproj [128][32] ...
      pack_bias_tile(w, b[5], 128); w += 128 * 8;                               //       ... + proj bias tile
      // MLP in 6 chunks of 64 hidden units: fc1 chunk = [64 x 128] K-major (8 k-steps) + its bias' tile,
      // fc2 chunk = [128 x 64] K-major (4 k-steps); the fc2 bias tile rides behind the first fc2 chunk
      auto w1 = [&](int c) {
        pack_kmajor(w, w1f.data(), 128, c * 64, 64, 0, 128); w += 64 * 128;
        pack_bias_tile(w, b1f.data() + c * 64, 64); w += 64 * 8;
      };
      auto w2 = [&](int c) {
        pack_kmajor(w, b[10], 384, 0, 128, c * 64, 64); w += 128 * 64;
        if (c == 0) { pack_bias_tile(w, b[11], 128); w += 128 * 8; }
      };
      w1(0); w1(1);                                                              // issue order of csrc/encoder_tc.cu: fc2 one chunk late
      for (int c = 1; c < 6; ++c) { if (c < 5) w1(c + 1); w2(c - 1); }
      w2(5);
      if ((size_t)(reinterpret_cast<unsigned char*>(w) - blk) != encoder_tc_block_bytes()) {
        set_error("internal: packed block is %zu bytes, kernel expects %zu", (size_t)(reinterpret_cast<unsigned char*>(w) - blk),
                  encoder_tc_block_bytes());
        return UITK_EINVAL;
      }
    }
    for (int c = 0; c < 128; ++c) cb[c] = (cb[c] + b[5][c]) + b[11][c];
  }
  memcpy(W + l.cb_final, cb.data(), 128 * 4);
  // tensor-core tile: 24 row slots per clip-crop, slot = band * 6 + tau; time slots beyond time_pos_embed are never live
  for (int tok = 0; tok < 24; ++tok)
    for (int c = 0; c < 128; ++c) {
      const float tp = tok % 6 < cfg->grid_t ? t[6][(size_t)c * cfg->grid_t + tok % 6] : 0.f;
      W[l.pos_tab + (size_t)tok * 128 + c] = (t[5][c] + tp) + t[7][(size_t)c * 4 + tok / 6];
    }
  return UITK_OK;
}

}  // extern "C"
