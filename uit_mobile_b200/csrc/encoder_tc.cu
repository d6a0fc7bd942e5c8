// Tensor-core encoder (UITK_PREC_BF16): ONE persistent megakernel runs patch embed + every transformer block for a
// tile of 128 token rows (5 clip-crops x 24 tokens) per CTA, with tcgen05.mma (bf16 x bf16 -> fp32 in TMEM).
// TWO CTAs are resident per SM (96 regs/thread, 110 KB SMEM, 256 TMEM columns each) so that one CTA's
// LayerNorm / softmax / ReLU work on the CUDA cores overlaps the other CTA's MMAs and weight streaming.
//
//   * the fp32 residual stream x[128 x 128] never leaves TENSOR MEMORY (columns 0..127): the proj and fc2 GEMMs
//     accumulate straight onto it (the residual add is the MMA's accumulate);
//   * every Linear bias is added by the tensor core: a [N x 8] bf16 "bias tile" (hi/mid/lo split of the fp32 bias,
//     pack.cu) rides behind the weights in the ring and one extra k-step multiplies it with a constant ones operand.
//     The CUDA cores never load or add a parameter: LayerNorm affines are folded into the next Linear at pack time;
//   * accumulators / TMEM operands live in columns 128..255: qkv (96) -> S_h (128) -> P_h (packed bf16, 64) + O_h (16)
//     in the attention phase; LayerNorm-2 output (packed bf16, 64) + fc1 chunk (64) in the MLP phase.  P_h and the
//     LayerNorm-2 output are TMEM A operands (TS-mode tcgen05.mma): no shared-memory store, read or proxy fence, and
//     the k-step runs at the tensor-core floor (an SS-mode k-step at N <= 128 is bound by the 128 B/cycle of SMEM);
//   * the remaining A operands (LayerNorm-1 output, Q/K/V^T, attention output, ReLU hidden chunks) are produced by the
//     CUDA cores straight from tcgen05.ld registers into K-major core-matrix shared-memory tiles (thread == row, so
//     every 16-byte store of a warp is contiguous: no bank conflicts, no swizzle needed);
//   * weights are pre-packed on the host in exactly that shared-memory layout and in consumption order, so a producer
//     warp streams them L2 -> SMEM with plain 1-D cp.async.bulk copies through a 3 x 18 KB mbarrier ring (full/empty,
//     slots released by tcgen05.commit);
//   * one MMA-issuer WARP (all 32 lanes run the warp-uniform issue code, one elected lane executes the tcgen05
//     instructions, so descriptors stay in uniform registers and UTCHMMA goes out back to back); the compute warps
//     hand operands over through non-blocking named-barrier arrivals and only ever wait on MMA-completion mbarriers;
//   * the MLP runs as 6 hidden chunks of 64 columns: "accumulator drained" is signalled as soon as the chunk is in
//     registers so fc1[c+1] runs under the ReLU epilogue of chunk c, fc2 is issued one chunk late together with the
//     next fc1 (the issue of a group blocks the issuer for about its execution time), hidden chunks cycle through a
//     3-deep ring;
//   * attention on the tensor cores: S_h = Q_h K_h^T for the whole tile (N = 128), fp32 softmax of each row's own 24
//     keys straight from TMEM, O_h = P_h V_h with the block-diagonal P_h in TMEM.
// LayerNorm, softmax, residual and all accumulation are fp32; only GEMM operands are rounded to bf16.
// Measured design inputs (scripts/microbench/tc_micro.cu, profiles/r1_tc_microbench.txt): SS k-step = max(~39, N/2)
// cycles, TS k-step = max(~11, N/2); fence.proxy.async ~120-140; mbarrier wake ~140-150; named barrier ~30.
// Reference semantics: models/uit.py:379-396 (features), 89-122 (attention), 181-248 (MLP, block).
#include <cstdlib>
#include <type_traits>

#include "tc_ptx.cuh"
#include "uitk_common.cuh"

namespace uitk {

namespace {

using namespace tc;

constexpr int kThreads = 320;          // 8 compute warps + TMA producer warp + MMA issuer warp
constexpr int kCompute = 256;
constexpr uint32_t kTile = 16384;                     // one [128 x 64] bf16 K-major weight tile = 4 MMA k-steps
constexpr uint32_t kBiasTile = 128 * 16;              // [128 x 8] bf16 bias tile (pack.cu: pack_bias_tile)
constexpr uint32_t kSlot = kTile + kBiasTile;         // ring slot: a weight tile and, for some, the bias tile behind it
constexpr int kSlots = 3;
constexpr uint32_t kQkvHalfBytes = 96 * 64 * 2;       // Wqkv, one K half: 12288
constexpr uint32_t kQkvBiasBytes = 96 * 16;           // 1536
constexpr uint32_t kProjBytes = 128 * 32 * 2;         // 8192
constexpr uint32_t kFc1BiasBytes = 64 * 16;           // bias tile of one 64-column fc1 chunk
// per block: [Wqkv k0 | bqkv] [Wqkv k1] [Wproj | bproj]  6 x [W1 chunk | b1 chunk]  6 x [W2 chunk] (+ b2 behind the first)
constexpr uint32_t kBlockBytes = 2 * kQkvHalfBytes + kQkvBiasBytes + kProjBytes + kBiasTile + 6 * (kTile + kFc1BiasBytes) + 6 * kTile + kBiasTile;
constexpr uint32_t kPatchBytes = 4 * kTile;           // 4 K-quarters of the patch weight

// shared memory map (bytes); one CTA uses 110.3 KB so that two fit on an SM
//   [0, 48 KB)   three 16 KB operand slots.  Attention phase: slot 0|1 = A (LayerNorm-1 output, 32 KB), later A_o + Q + K over
//                it; slot 2 = V^T.  MLP phase: 3-deep ring of hidden chunks (ReLU output, A operand of fc2).  Patch embed: the
//                [128 x 128] bf16 operand of one K half (32 KB).
//   [48 KB, ..)  weight ring: kSlots x 18 KB, filled by the producer warp with 1-D bulk copies in consumption order
constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_AO = OFF_A;                 // 8 KB attention output operand [128 x 32]
constexpr uint32_t OFF_Q = OFF_A + 8192;           // Q_h  [128 x 16] bf16 K-major, head h at + h*4096
constexpr uint32_t OFF_K = OFF_A + 16384;          // K_h  [128 x 16] (B operand: N = key row)
// V_h^T [16 x 128] (B operand: N = d, K = key row).  k-groups are 272 B apart (LBO is free in the descriptor): the 4 key
// groups a warp's 2-byte transposing stores hit then fall into distinct banks.  Head h at + h * kVtHead.
constexpr uint32_t OFF_VT = OFF_A + 32768;
constexpr uint32_t kVtLbo = 272, kVtHead = 16 * kVtLbo;
constexpr int kHSlots = 3;                         // hidden-chunk slots (16 KB each) at OFF_A + s * 16384
// P_h (softmax probabilities, block diagonal) never touches shared memory: it is written to TENSOR MEMORY (accumulator
// columns 0..63, packed bf16 pairs, over the dead S_h) and is the TMEM A operand of the P V product; O_h lands in
// accumulator columns 64..79.
constexpr uint32_t kColP = 0, kColO = 64;
// MLP phase: fc1 chunk accumulator (64 fp32 columns) and the LayerNorm-2 output as packed-bf16 TMEM A operand (64 columns)
constexpr uint32_t kColFc1 = 0, kColLn2 = 64;
// Patch embed: operand slot 2 is idle (the patch operand half takes slots 0|1), so the producer warp parks the per-token
// position table there ([24][128] fp32, rows 528 B apart: a quarter warp's 16-byte reads of 8 consecutive token rows then
// hit 8 distinct bank groups).  Read by every compute thread once per tile, before block 0 writes V^T over it.
constexpr uint32_t OFF_POS = 2 * 16384;
constexpr uint32_t kPosRow = 132;                  // floats per staged table row
static_assert(24 * kPosRow * 4 <= 16384, "position table must fit operand slot 2");
constexpr uint32_t OFF_RING = kHSlots * 16384;
// constant MMA operands: k-group of (1, 1, 1, 0, ..) rows = the A operand of every bias k-step, then 2 KB of zeros that
// serve as the second k-group of both the ones operand and every bias tile (their LBO points here: it must lie ABOVE the ring)
constexpr uint32_t OFF_CONST = OFF_RING + kSlots * kSlot;
constexpr uint32_t OFF_PART = OFF_CONST + 4096;             // 2 x 512 floats
constexpr uint32_t OFF_BAR = OFF_PART + 2 * 512 * 4;   // two alternating buffers of [2][128] float2 partial sums
constexpr uint32_t kSmemBytes = OFF_BAR + 256;
static_assert(kSmemBytes <= 115712, "two CTAs per SM need <= 113 KB of dynamic shared memory each");
constexpr uint32_t kTmemCols = 256;

enum { B_FULLW = 0, B_EMPTYW = kSlots, B_ACC = 2 * kSlots, B_X, B_FC1, B_POS, B_H /* kHSlots hidden slots */, B_COUNT = B_H + kHSlots };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier block");

struct TcParams {
  const unsigned char* wts;     // bf16 section
  const float* pos_tab;         // [24][128] conv bias + time_pos + freq_pos per token of a crop
  const float* bn_scale; const float* bn_shift;
  const float* norm_w; const float* norm_b;
  const float* db; const uint32_t* max_pow;
  int T, crops, tokens, t_n, target;   // tokens = 24 row SLOTS per clip-crop (4 mel bands x 6 time slots); t_n = valid time patches
  int RR, G, num_tiles, depth;
  int no_clamp;    // input is an already normalised spectrogram (uitk_forward_features): no top-dB clamp
  float* pooled;   // [RR][128]
  float* dbg_x;    // optional [RR*24][128]: residual stream after the last block (pre final LN), slot layout
  const uint32_t* c_used; const uint32_t* c_min;   // uitk_encoder_fixup: return at once unless fixup_needed(max_pow, c_used, c_min)
};

// Optional in-kernel timeline (build with -DUITK_TRACE): thread 0 of CTA 0 and its MMA-issuer thread stamp
// (id << 44 | clock) at every stage boundary; read back with uitk_debug_read_trace.  Compiled out by default.
// CAUTION when reading it: the stamps themselves cost the MMA-issuer warp ~10 % per block (it is the busiest warp), so
// stage SHARES are meaningful, absolute block times are not - decide between kernel variants with the normal build.
#ifdef UITK_TRACE
__device__ long long g_trace[3][4096];     // [2]: per CTA (globaltimer at start, at end, SM id, SM cycles start -> end), 4 words each
__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid() { int v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
// UITK_TRACE=2: stamps of the compute thread only (the issuer warp runs undisturbed: realistic block times)
#define TR(buf, id) do { if ((UITK_TRACE < 2 || (buf) == 0) && trace_on && tr_n < 4096) g_trace[buf][tr_n++] = ((long long)(id) << 44) | (clock64() & ((1ll << 44) - 1)); } while (0)
#else
#define TR(buf, id) do {} while (0)
#endif

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// the two warps that share TMEM lane quarter q (hsel 0 / 1): the only threads that exchange per-row partial results
__device__ __forceinline__ void bar_pair(int q) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); }
// compute warps -> MMA issuer hand-off on named barriers (ids 6..9 in rotation): the 256 compute threads arrive without
// blocking, the 32 issuer threads sync.  A named-barrier release costs ~30 cycles against ~140 for an mbarrier wake-up.
// FOUR ids: in the MLP the compute warps can be up to three signals ahead of the issuer (H-ready(c) still unconsumed while
// drained(c+1) and H-ready(c+1) arrive; the next signal needs fc1[c+2], which the issuer only issues after consuming
// drained(c+1)), and a barrier id must never collect arrivals of two signals at once.
constexpr int kReadyThreads = kCompute + 32;
__device__ __forceinline__ void ready_arrive(uint32_t sig) { asm volatile("bar.arrive %0, %1;" ::"r"(6 + (sig & 3)), "n"(kReadyThreads) : "memory"); }
__device__ __forceinline__ void ready_sync(uint32_t sig) { asm volatile("bar.sync %0, %1;" ::"r"(6 + (sig & 3)), "n"(kReadyThreads) : "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// wait executed by every thread of a warp: re-converge before the .sync.aligned tcgen05 instructions that follow
__device__ __forceinline__ void mbar_wait_all(uint64_t* bar, uint32_t parity) {
  mbar_wait(bar, parity);
  __syncwarp();
}

// Exchange of the two half-row partial sums (sum, sum of squares) of a row through `part`, mean / rstd of the full row.
// var = E[x^2] - mean^2 in fp32: |mean| is O(std) for these activations, the cancellation costs < 1e-6 relative.
__device__ __forceinline__ void finish_stats(float s, float ss, int hsel, int r, float eps, float* part, float& mean, float& rstd) {
  *reinterpret_cast<float2*>(part + (hsel * 128 + r) * 2) = make_float2(s, ss);
  bar_pair(r >> 5);
  const float2 oth = *reinterpret_cast<const float2*>(part + ((hsel ^ 1) * 128 + r) * 2);
  mean = (s + oth.x) * (1.f / 128.f);
  const float var = fmaxf((ss + oth.y) * (1.f / 128.f) - mean * mean, 0.f);
  rstd = rsqrtf(var + eps);
}

// Row statistics of this thread's row of x (two threads per row, 64 columns each), values not kept.
__device__ __forceinline__ void row_stats(uint32_t tx, int hsel, int r, float eps, float* part, float& mean, float& rstd) {
  float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);     // packed fp32x2 accumulators (even, odd columns)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float v[32];
    tmem_ld32(tx + hsel * 64 + j * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 xa = make_float2(v[i], v[i + 1]);
      s2 = add2(s2, xa);
      q2 = fma2(xa, xa, q2);
    }
  }
  finish_stats(s2.x + s2.y, q2.x + q2.y, hsel, r, eps, part, mean, rstd);
}

// LayerNorm of this thread's half row straight out of TMEM (ONE read: the 64 values stay in registers between the
// statistics and the normalisation), written as bf16 K-major core-matrix chunks.
// Only the normalisation (x - mean) * rstd happens here: the affine part (gamma, beta) is folded into the weights
// and bias of the Linear that consumes the operand when the weights are packed (W' = W diag(gamma), b' = b + W beta).
__device__ __forceinline__ void ln_to_operand(uint32_t tx, int hsel, int r, float eps, float* part, unsigned char* dst) {
  float v0[32], v1[32];
  tmem_ld32(tx + hsel * 64, v0);
  tmem_ld32(tx + hsel * 64 + 32, v1);
  tmem_ld_wait();
  float2 sa = make_float2(0.f, 0.f), sb = sa, qa = sa, qb = sa;       // independent chains
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 xa = make_float2(v0[i], v0[i + 1]), xb = make_float2(v1[i], v1[i + 1]);
    sa = add2(sa, xa); sb = add2(sb, xb);
    qa = fma2(xa, xa, qa); qb = fma2(xb, xb, qb);
  }
  float mean, rstd;
  finish_stats((sa.x + sa.y) + (sb.x + sb.y), (qa.x + qa.y) + (qb.x + qb.y), hsel, r, eps, part, mean, rstd);
  const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
  unsigned char* d = dst + hsel * 8 * 2048 + r * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o;
    float2 y;
    y = fma2(make_float2(v0[c * 8 + 0], v0[c * 8 + 1]), rs2, nm2); o.x = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v0[c * 8 + 2], v0[c * 8 + 3]), rs2, nm2); o.y = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v0[c * 8 + 4], v0[c * 8 + 5]), rs2, nm2); o.z = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v0[c * 8 + 6], v0[c * 8 + 7]), rs2, nm2); o.w = pack2_bf16(y.x, y.y);
    *reinterpret_cast<uint4*>(d + c * 2048) = o;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o;
    float2 y;
    y = fma2(make_float2(v1[c * 8 + 0], v1[c * 8 + 1]), rs2, nm2); o.x = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v1[c * 8 + 2], v1[c * 8 + 3]), rs2, nm2); o.y = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v1[c * 8 + 4], v1[c * 8 + 5]), rs2, nm2); o.z = pack2_bf16(y.x, y.y);
    y = fma2(make_float2(v1[c * 8 + 6], v1[c * 8 + 7]), rs2, nm2); o.w = pack2_bf16(y.x, y.y);
    *reinterpret_cast<uint4*>(d + (4 + c) * 2048) = o;
  }
}

// Same LayerNorm, but the normalised half row goes to TENSOR MEMORY as packed bf16 pairs (32 columns per thread, ONE
// tcgen05.st): the A operand of a TS-mode tcgen05.mma.  No shared-memory traffic and no proxy fence.
__device__ __forceinline__ void ln_to_tmem(uint32_t tx, uint32_t tdst, int hsel, int r, float eps, float* part) {
  float v0[32], v1[32];
  tmem_ld32(tx + hsel * 64, v0);
  tmem_ld32(tx + hsel * 64 + 32, v1);
  tmem_ld_wait();
  float2 sa = make_float2(0.f, 0.f), sb = sa, qa = sa, qb = sa;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 xa = make_float2(v0[i], v0[i + 1]), xb = make_float2(v1[i], v1[i + 1]);
    sa = add2(sa, xa); sb = add2(sb, xb);
    qa = fma2(xa, xa, qa); qb = fma2(xb, xb, qb);
  }
  float mean, rstd;
  finish_stats((sa.x + sa.y) + (sb.x + sb.y), (qa.x + qa.y) + (qb.x + qb.y), hsel, r, eps, part, mean, rstd);
  const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
  float pk[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 ya = fma2(make_float2(v0[2 * i], v0[2 * i + 1]), rs2, nm2);
    const float2 yb = fma2(make_float2(v1[2 * i], v1[2 * i + 1]), rs2, nm2);
    pk[i] = __uint_as_float(pack2_bf16(ya.x, ya.y));
    pk[16 + i] = __uint_as_float(pack2_bf16(yb.x, yb.y));
  }
  tmem_st32(tdst + hsel * 32, pk);
  tmem_st_wait();
}

// relu(v) for 8 consecutive hidden columns -> one 16-byte bf16 chunk (the fc1 bias is already in the accumulator)
__device__ __forceinline__ uint4 relu_pack8(const float* v) {
  uint4 o;
  o.x = pack2_relu_bf16(v[0], v[1]); o.y = pack2_relu_bf16(v[2], v[3]);
  o.z = pack2_relu_bf16(v[4], v[5]); o.w = pack2_relu_bf16(v[6], v[7]);
  return o;
}

// Patch gather of one K half: clamp(dB) * BatchNorm scale + shift -> bf16, written straight into the K-major operand tile.
// tokens == 24 (t_n == 6: 96 frames = 3 x 32 lanes, no tail predicate).  Element (mel row u, frame tt = lane + 32 j):
// k = (dfl0 + u) * 16 + (lane & 15), row = g * 24 + f * 6 + 2 j + (lane >> 4)  ->  shared offset = per-lane base
// + u * 4096 + j * 32 + g * 384 with compile-time u, j terms: the stores need no address arithmetic.
// The gather is bound by the round-trip latency of its loads (~1.5 k cycles each in the stage timeline), not by their count,
// so ALL (up to 5) clip-crops of the tile are loaded at once: 60 loads in flight per lane.  Deliberately NOT inlined: inside the
// megakernel the persistent state of the tile loop leaves room for a dozen loads only (ptxas spilled every load result to
// local memory, which serialises them); as a call, the state is saved once around it and the 60 values stay in registers.
template <bool kMasked>
__device__ __noinline__ void gather_half(const float* __restrict__ db, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                         unsigned char* A, int T, int crops, int target, int rr0, int g_cnt, int half, int warp, int lane,
                                         int t_n, float cutoff, uint64_t* wait_bar, uint32_t wait_parity) {
  const int f = warp >> 1, dfl0 = 4 * (warp & 1);
  const int mel0 = 16 * f + 8 * half + dfl0;
  unsigned char* lane_dst = A + (2 * dfl0 + ((lane & 15) >> 3)) * 2048 + (f * 6 + (lane >> 4)) * 16 + (lane & 7) * 2;
  float val[5][4][3];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    {
      const int rr = rr0 + min(g, g_cnt - 1);          // a ragged last tile re-loads its last clip-crop (never stored)
      int b = rr, start = 0;
      if (crops > 1) {
        b = rr / crops;
        start = (rr - b * crops) * target;
        if (start > T - target) start = T - target;
      }
      const float* src = db + ((size_t)b * 64 + mel0) * T + start + lane;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 3; ++j) val[g][u][j] = (!kMasked || lane + 32 * j < 16 * t_n) ? __ldg(src + 32 * j) : 0.f;
        src += T;
      }
    }
  }
  float sc[4], sh[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { sc[u] = __ldg(bn_scale + mel0 + u); sh[u] = __ldg(bn_shift + mel0 + u); }
  if (wait_bar != nullptr) mbar_wait_all(wait_bar, wait_parity);       // the operand tile is free (loads already in flight)
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    if (g < g_cnt) {
      unsigned char* d = lane_dst + g * (24 * 16);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float y = fmaf(fmaxf(val[g][u][j], cutoff), sc[u], sh[u]);
          *reinterpret_cast<__nv_bfloat16*>(d + u * 4096 + j * 32) =
              __float2bfloat16_rn((!kMasked || lane + 32 * j < 16 * t_n) ? y : 0.f);       // masked slots: zero patch
        }
    }
  }
}

// kMasked = false: the native geometry (1 s clips: 6 time patches, all 24 token slots live).
// kMasked = true : clips of 2400 .. 15 359 samples (1 .. 5 time patches).  The tile keeps the 24-slot layout (slot = band * 6 +
//   tau) so that every offset, the position table and the 5-clips-per-tile packing stay as they are; slots with tau >= t_n
//   carry a zero patch, are masked out of every softmax (probability exactly 0) and out of the token mean.  Wasted rows, but
//   the clip runs on the tensor cores instead of the 33x slower fp32 CUDA-core kernels.
template <bool kMasked>
__global__ void __launch_bounds__(kThreads, 2) encoder_tc_kernel(const TcParams p) {
  if (p.c_used != nullptr && !fixup_needed(p.max_pow, p.c_used, p.c_min)) return;   // uniform: before any allocation / barrier
  extern __shared__ __align__(1024) unsigned char smem[];
#ifdef UITK_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 1024) {
    g_trace[2][blockIdx.x * 4] = globaltimer_ns(); g_trace[2][blockIdx.x * 4 + 2] = smid(); g_trace[2][blockIdx.x * 4 + 3] = -clock64();
  }
#endif
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
  float* part = reinterpret_cast<float*>(smem + OFF_PART);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches stay uniform

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (tid < 256) {   // ones k-group (rows of 1, 1, 1, 0, 0, 0, 0, 0 in bf16) followed by the shared all-zero k-group
    *reinterpret_cast<uint4*>(smem + OFF_CONST + tid * 16) = tid < 128 ? make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =================================== weight producer =======================================
    // warp-uniform code, one elected lane issues the bulk copies (see the MMA issuer below for why)
    {
      uint32_t slot = 0, phase = 0;
      auto ring_load = [&](const unsigned char*& src, uint32_t bytes) {
        mbar_wait(&bars[B_EMPTYW + slot], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[B_FULLW + slot], bytes);
          bulk_g2s(smem + OFF_RING + slot * kSlot, src, bytes, &bars[B_FULLW + slot]);
        }
        __syncwarp();
        src += bytes;
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        // The dB rows of a tile come from HBM (the log-mel of a 4096-clip batch does not survive in L2 next to the waveforms)
        // and the patch gather is latency-bound on them: pull the NEXT tile's clips into L2 now, a whole tile ahead.
        if (p.crops == 1 && tile + (int)gridDim.x < p.num_tiles && elect_one()) {
          const int rr_next = (tile + (int)gridDim.x) * p.G;
          const int n_next = min(p.G, p.RR - rr_next);
          const size_t clip_bytes = (size_t)64 * p.T * sizeof(float);
          const uintptr_t a0 = reinterpret_cast<uintptr_t>(p.db) + (size_t)rr_next * clip_bytes;
          const uintptr_t a1 = (a0 + (size_t)n_next * clip_bytes) & ~(uintptr_t)15;
          for (uintptr_t a = (a0 + 15) & ~(uintptr_t)15; a < a1; a += 32768) {
            const uintptr_t left = a1 - a;
            bulk_prefetch_l2(reinterpret_cast<const void*>(a), (uint32_t)(left < 32768 ? left : 32768));
          }
        }
        __syncwarp();
        const unsigned char* wb = p.wts;
        for (int c = 0; c < 4; ++c) {
          ring_load(wb, kTile);
          if (c == 2) {
            // Patch quarter 2 re-uses the ring slot of the previous tile's last fc2 chunk: its "empty" barrier has just
            // completed, so every MMA of that tile is done and nothing reads operand slot 2 any more (the pooling stage
            // only uses slots 0|1).  Stage the position table there, one 512-byte row at a time.
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars[B_POS], 24 * 512);
              for (int t = 0; t < 24; ++t) bulk_g2s(smem + OFF_POS + t * (kPosRow * 4), p.pos_tab + t * 128, 512, &bars[B_POS]);
            }
            __syncwarp();
          }
        }
        for (int blk = 0; blk < p.depth; ++blk) {                 // same order as the MMA issuer consumes (pack.cu)
          ring_load(wb, kQkvHalfBytes + kQkvBiasBytes);
          ring_load(wb, kQkvHalfBytes);
          ring_load(wb, kProjBytes + kBiasTile);
          ring_load(wb, kTile + kFc1BiasBytes);                                                      // fc1[0]
          ring_load(wb, kTile + kFc1BiasBytes);                                                      // fc1[1]
          for (int c = 1; c < 6; ++c) {
            if (c < 5) ring_load(wb, kTile + kFc1BiasBytes);                                          // fc1[c+1]
            ring_load(wb, kTile + (c == 1 ? kBiasTile : 0u));                                         // fc2[c-1] (+ fc2 bias once)
          }
          ring_load(wb, kTile);                                                                      // fc2[5]
        }
      }
    }
  } else if (warp == 9) {
    // =================================== MMA issuer =======================================
    // One warp issues every tcgen05.mma of the CTA.  ALL 32 lanes run this (warp-uniform) code and one elected lane
    // executes the tcgen05 instructions: the descriptors then live in uniform registers and the UTCHMMA instructions go
    // out back to back.  (Issuing from inside an `if (lane == 0)` region makes the compiler wrap every MMA in an
    // R2UR + ELECT waterfall loop, ~75-135 cycles per instruction: measured, scripts/microbench/tc_micro.cu.)
    // It waits for (a) the "operand ready" signal of the 256 compute threads (alternating mbarriers) and (b) the weight
    // chunk in the ring; the compute warps never block on weights, only on the completion barriers (ACC / X / FC1 / H).
    {
      const uint32_t sA = smem_u32(smem + OFF_A), sRing = smem_u32(smem + OFF_RING);
      const uint32_t sOnes = smem_u32(smem + OFF_CONST), sZero = sOnes + 2048;
      const uint32_t sQ = smem_u32(smem + OFF_Q), sK = smem_u32(smem + OFF_K), sVT = smem_u32(smem + OFF_VT);
      constexpr uint32_t ID128 = make_idesc_bf16(128, 128), ID96 = make_idesc_bf16(128, 96), ID64 = make_idesc_bf16(128, 64);
      constexpr uint32_t ID16 = make_idesc_bf16(128, 16);
      uint32_t cslot = 0, cphase = 0, sig = 0;
#ifdef UITK_TRACE
      const bool trace_on = blockIdx.x == 0 && lane == 0;
      int tr_n = 0;
#endif
      bool prewaited = false;
      // the next ring slot is almost always full long before the operand is ready: take that mbarrier wait (~90 cycles
      // even when already complete) off the critical path by doing it BEFORE blocking on the ready barrier
      auto wait_ready = [&]() {
        if (!prewaited) { mbar_wait(&bars[B_FULLW + cslot], cphase); prewaited = true; }
        ready_sync(sig);
        ++sig;
        tc_fence_after();
        TR(1, 1);
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      // KSTEPS MMAs (K = 16 each) of A[128 x 16*KSTEPS] (k-groups 2048 B apart) with the B chunk in the current ring slot
      // bias_off >= 0: the slot also carries a bias tile at that byte offset -> one more k-step  D += ones * bias^T
      auto mbar_wait_ring = [&]() {
        if (!prewaited) mbar_wait(&bars[B_FULLW + cslot], cphase);
        prewaited = false;
        tc_fence_after();
        TR(1, 2);
      };
      auto mma_from_ring = [&](uint32_t d_tmem, uint32_t a_base, uint32_t idesc, uint32_t b_lbo, auto ksteps_c, bool accum_first,
                               int bias_off = -1) {
        constexpr int KSTEPS = decltype(ksteps_c)::value;
        mbar_wait_ring();
        const uint32_t b_base = sRing + cslot * kSlot;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks)
            umma_bf16(d_tmem, make_smem_desc(a_base + ks * 4096, 2048, 128), make_smem_desc(b_base + ks * 2 * b_lbo, b_lbo, 128), idesc,
                      (accum_first || ks > 0) ? 1u : 0u);
          if (bias_off >= 0)
            umma_bf16(d_tmem, make_smem_desc(sOnes, 2048, 128), make_smem_desc(b_base + bias_off, sZero - (b_base + bias_off), 128), idesc, 1u);
          umma_commit(&bars[B_EMPTYW + cslot]);
        }
        __syncwarp();
        TR(1, 3);
        if (++cslot == kSlots) { cslot = 0; cphase ^= 1; }
      };
      using K2 = std::integral_constant<int, 2>;
      using K4 = std::integral_constant<int, 4>;
      using K8 = std::integral_constant<int, 8>;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int half = 0; half < 2; ++half) {                                // patch embed, one K half (128 of 256) at a time
          wait_ready();                                                       // operand half gathered into A
          for (int c = 0; c < 2; ++c) mma_from_ring(tmem, sA + c * 16384, ID128, 2048, K4{}, half + c > 0);
          commit(&bars[B_ACC]);
        }
        for (int blk = 0; blk < p.depth; ++blk) {
          wait_ready();                                                       // LN1 output in A
          mma_from_ring(tmem + 128, sA, ID96, 1536, K4{}, false, kQkvHalfBytes);
          mma_from_ring(tmem + 128, sA + 16384, ID96, 1536, K4{}, true);
          commit(&bars[B_ACC]);
          {                                                                   // attention GEMMs (no weights involved)
            for (int h = 0; h < 2; ++h) {
              wait_ready();                                                   // Q/K/V^T operands written (h=0) / ACC drained (h=1)
              if (elect_one()) {
                umma_bf16(tmem + 128, make_smem_desc(sQ + h * 4096, 2048, 128), make_smem_desc(sK + h * 4096, 2048, 128), ID128, 0u);   // S_h = Q_h K_h^T
                umma_commit(&bars[B_ACC]);
              }
              __syncwarp();
              wait_ready();                                                   // P_h written, S_h consumed
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)                                // O_h = P_h V_h  (N = 16, K = 128), P from TMEM
                  umma_bf16_ts(tmem + 128 + kColO, tmem + 128 + kColP + ks * 8, make_smem_desc(sVT + h * kVtHead + ks * 2 * kVtLbo, kVtLbo, 128),
                               ID16, ks > 0 ? 1u : 0u);
                umma_commit(&bars[B_ACC]);
              }
              __syncwarp();
            }
          }
          wait_ready();                                                       // attention output in A_o
          mma_from_ring(tmem, sA /* == A_o */, ID128, 2048, K2{}, true, kProjBytes);   // x += o Wproj^T + bproj
          commit(&bars[B_X]);
          wait_ready();                                                       // LN2 output in tensor memory (acc columns 64..127)
          // MLP in 6 chunks of 64 hidden units.  fc1 reads its A operand (the LayerNorm output, packed bf16) from TENSOR
          // MEMORY: no shared-memory A traffic, and at N = 64 a TS-mode k-step takes 32 cycles (SS mode: 48, SMEM-bound).
          auto fc1 = [&]() {
            mbar_wait_ring();
            const uint32_t b_base = sRing + cslot * kSlot;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                umma_bf16_ts(tmem + 128 + kColFc1, tmem + 128 + kColLn2 + ks * 8, make_smem_desc(b_base + ks * 2048, 1024, 128), ID64, ks > 0 ? 1u : 0u);
              umma_bf16(tmem + 128 + kColFc1, make_smem_desc(sOnes, 2048, 128), make_smem_desc(b_base + kTile, sZero - (b_base + kTile), 128), ID64, 1u);
              umma_commit(&bars[B_EMPTYW + cslot]);
              umma_commit(&bars[B_FC1]);
            }
            __syncwarp();
            TR(1, 3);
            if (++cslot == kSlots) { cslot = 0; cphase ^= 1; }
          };
          // The issue of an MMA group blocks this warp for about as long as the group executes, so fc2 runs ONE CHUNK
          // LATE: on "accumulator c drained" the issuer sends fc1[c+1] and fc2[c-1] (whose hidden slot was signalled long
          // ago) in one go and is back waiting before drained(c+1) arrives; the hidden ring is 3 deep, so nobody waits.
          auto fc2 = [&](int c) {
            mma_from_ring(tmem, sA + (c % kHSlots) * 16384, ID128, 2048, K4{}, true, c == 0 ? (int)kTile : -1);   // x += H_c W2_c^T (+ b2 once)
            commit(&bars[B_H + (c % kHSlots)]);
          };
          fc1();                                                              // fc1[0]
          wait_ready();                                                       // accumulator 0 drained into registers
          fc1();                                                              // fc1[1] runs while the ReLU epilogue of chunk 0 does
          for (int c = 1; c < 6; ++c) {
            wait_ready();                                                     // hidden slot of chunk c-1 written
            wait_ready();                                                     // accumulator c drained
            if (c < 5) fc1();                                                 // fc1[c+1]
            fc2(c - 1);
          }
          wait_ready();                                                       // hidden slot of chunk 5 written
          fc2(5);
        }
      }
    }
  } else {
    // =================================== compute warps =======================================
    const int q = warp & 3, hsel = warp >> 2;
    const int r = q * 32 + lane;                                   // row == TMEM lane
    const uint32_t tx = tmem + ((uint32_t)(q * 32) << 16);         // X columns 0..127
    const uint32_t tacc = tx + 128;                                // accumulator columns 128..255
    uint32_t ph_acc = 0, ph_x = 0, ph_fc1 = 0, ph_h = 0, ph_pos = 0;      // ph_h: one phase bit per hidden slot
    uint32_t sig = 0;                     // "operand ready" signal counter (mirrors the issuer's)
#ifdef UITK_TRACE
    const bool trace_on = blockIdx.x == 0 && tid == 0;
    int tr_n = 0;
#endif
    // operand written by this thread is visible to the async proxy, its TMEM reads are done: tell the MMA issuer
    // (every thread fences its own writes; __syncwarp orders the warp; ONE lane arrives -> 8 arrivals per signal instead
    // of 256 serialized shared-memory atomics)
    auto signal_ready = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      ready_arrive(sig);
      ++sig;
    };
    // same signal stream, but nothing was written to shared memory: this thread's TMEM reads are done
    auto signal_drained = [&]() {
      tc_fence_before();
      ready_arrive(sig);
      ++sig;
    };

    const float cutoff = p.no_clamp ? -INFINITY : 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(*p.max_pow), 1e-10f)) - 120.f;
    const int tokens = p.tokens;          // 24 slots per clip-crop
    const int t_n = kMasked ? p.t_n : 6;  // valid time patches (slots with tau >= t_n are masked)

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int rr0 = tile * p.G;
      const int g_cnt = min(p.G, p.RR - rr0);
      const int rows_valid = g_cnt * tokens;
      TR(0, 1);

      // ---------------- patch embed: gather (clamp + BatchNorm) -> bf16 A, one K half [128 x 128] at a time ----------------
      // k = df * 16 + dt (df = mel & 15): K half h holds df in [8h, 8h + 8).  Warp w gathers token row f = w >> 1 and the
      // 4 mel rows 16 f + 8 h + 4 (w & 1) + u of every clip-crop of the tile (12 coalesced loads in flight per lane).
      for (int i = tid; i < (128 - rows_valid) * 16; i += kCompute) {        // zero the padding rows (both K halves use the same bytes)
        const int rz = rows_valid + i / 16, k8 = i % 16;
        *reinterpret_cast<uint4*>(smem + OFF_A + k8 * 2048 + rz * 16) = make_uint4(0, 0, 0, 0);
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        // half 1: its loads are issued first, THEN the wait for the MMAs of half 0 (which still read A), then the stores
        gather_half<kMasked>(p.db, p.bn_scale, p.bn_shift, smem + OFF_A, p.T, p.crops, p.target, rr0, g_cnt, half, warp, lane, t_n, cutoff,
                             half == 1 ? &bars[B_ACC] : nullptr, ph_acc);
        if (half == 1) ph_acc ^= 1;
        signal_ready();
      }
      TR(0, 2);
      mbar_wait_all(&bars[B_ACC], ph_acc); ph_acc ^= 1;
      tc_fence_after();
      TR(0, 3);
      {   // x += conv bias + time_pos[tau] + freq_pos[f]   (uit.py:380-383; one pre-added table row per token), back to TMEM.
          // The table row comes from the staged copy in operand slot 2 (thread == row: from global memory these were 32
          // different cache lines per warp load, ~6.8 k cycles per tile).
        mbar_wait_all(&bars[B_POS], ph_pos); ph_pos ^= 1;
        const float* pt = reinterpret_cast<const float*>(smem + OFF_POS) + (r % tokens) * kPosRow;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[32];
          const int c0 = hsel * 64 + j * 32;
          tmem_ld32(tx + c0, v);
          tmem_ld_wait();
          if (r < rows_valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 pb = *reinterpret_cast<const float4*>(pt + c0 + i);
              v[i] += pb.x; v[i + 1] += pb.y; v[i + 2] += pb.z; v[i + 3] += pb.w;
            }
          }
          tmem_st32(tx + c0, v);
        }
        tmem_st_wait();
      }

      // ---------------- transformer blocks ----------------
      for (int blk = 0; blk < p.depth; ++blk) {
        TR(0, 10);

        // LN1 -> A ; qkv = A Wqkv^T  (two K halves through the ring)
        ln_to_operand(tx, hsel, r, 1e-6f, part, smem + OFF_A);
        signal_ready();
        TR(0, 11);
        mbar_wait_all(&bars[B_ACC], ph_acc); ph_acc ^= 1;
        tc_fence_after();
        TR(0, 12);
        {
          // ---------- tensor-core attention: S_h = Q_h K_h^T and O_h = P_h V_h on tcgen05, softmax straight from TMEM ----------
          const bool valid = r < rows_valid;
          const int g = valid ? r / 24 : 0;
          {   // qkv (+bias) -> bf16 operands.  Accumulator columns: q 0..31 | k 32..63 | v 64..95 (two heads of 16 each).
              // hsel 0 takes q + v of head 0, hsel 1 takes k + v of head 1: four 16-byte stores and 16 transposing 2-byte stores each
            float v[32], w[16];
            tmem_ld32(tacc + hsel * 32, v);
            tmem_ld16(tacc + 64 + hsel * 16, w);
            tmem_ld_wait();
            unsigned char* qk = smem + (hsel == 0 ? OFF_Q : OFF_K) + r * 16;   // Q_h / K_h: head c >> 1 at + 4096, k-group c & 1 at + 2048
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(qk + (c >> 1) * 4096 + (c & 1) * 2048) = pack8_bf16(v + c * 8);
            unsigned char* vt = smem + OFF_VT + hsel * kVtHead + (r >> 3) * kVtLbo + (r & 7) * 2;   // V_h^T[d][key r]
#pragma unroll
            for (int d = 0; d < 16; ++d) *reinterpret_cast<__nv_bfloat16*>(vt + d * 16) = __float2bfloat16_rn(w[d]);
          }
          signal_ready();
          TR(0, 13);
          constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;   // softmax(0.125 * s) through exp2
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            mbar_wait_all(&bars[B_ACC], ph_acc); ph_acc ^= 1;            // S_h in ACC
            tc_fence_after();
            TR(0, 14);
            // tcgen05.ld takes ONE (warp-uniform) column address, but the 32 rows of a warp straddle two clips: load the key
            // window of BOTH clips (one wait) and let every lane keep the one that belongs to its row.  The two warps of a
            // row pair split the 24 keys 12 / 12 (hsel 0: keys 0..11, hsel 1: keys 12..23): equal work on both.
            const int g_lo = (q * 32) / 24;                              // clip of the warp's first row (warp-uniform)
            const bool two = g_lo < 4;                                   // rows 96..127: clip 4 only (the rest is padding)
            float sv[12];
            {
              float t0[12], t1[12];
              tmem_ld8(tacc + g_lo * 24 + hsel * 12, t0); tmem_ld4(tacc + g_lo * 24 + hsel * 12 + 8, t0 + 8);
              if (two) { tmem_ld8(tacc + g_lo * 24 + 24 + hsel * 12, t1); tmem_ld4(tacc + g_lo * 24 + 24 + hsel * 12 + 8, t1 + 8); }
              tmem_ld_wait();
              const bool first = g == g_lo;
#pragma unroll
              for (int i = 0; i < 12; ++i) sv[i] = valid ? ((first || !two) ? t0[i] : t1[i]) : 0.f;
            }
            if (kMasked) {   // key slot hsel * 12 + i is live iff its time index i % 6 < t_n (slots 0 and 12 always are)
#pragma unroll
              for (int i = 0; i < 12; ++i)
                if (i % 6 >= t_n) sv[i] = -INFINITY;
            }
            float m_loc, l_loc;
            {
              const float m0 = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3])), m1 = fmaxf(fmaxf(sv[4], sv[5]), fmaxf(sv[6], sv[7]));
              const float m2 = fmaxf(fmaxf(sv[8], sv[9]), fmaxf(sv[10], sv[11]));
              m_loc = fmaxf(fmaxf(m0, m1), m2);
              const float mb = -m_loc * kScaleLog2e;
#pragma unroll
              for (int i = 0; i < 12; ++i) sv[i] = ex2_approx(fmaf(sv[i], kScaleLog2e, mb));
              l_loc = (((sv[0] + sv[1]) + (sv[2] + sv[3])) + ((sv[4] + sv[5]) + (sv[6] + sv[7]))) + ((sv[8] + sv[9]) + (sv[10] + sv[11]));
            }
            float* ex = part + h * 512;
            *reinterpret_cast<float2*>(ex + (hsel * 128 + r) * 2) = make_float2(m_loc, l_loc);
            bar_pair(q);       // the partner thread of this row is in the other warp of the pair; its S_h reads are done too
            const float2 oth = *reinterpret_cast<const float2*>(ex + ((hsel ^ 1) * 128 + r) * 2);
            const float M = fmaxf(m_loc, oth.x);
            const float f_own = ex2_approx((m_loc - M) * kScaleLog2e), f_oth = ex2_approx((oth.x - M) * kScaleLog2e);
            const float f = valid ? __fdividef(f_own, l_loc * f_own + oth.y * f_oth) : 0.f;
            // P_h -> tensor memory.  Row r holds its clip's 24 probabilities in packed columns [12 g, 12 g + 12) (this thread:
            // 6 of them, at + 6 hsel) and zeros elsewhere.  tcgen05.st takes one (warp-uniform) column address, so every clip's
            // column block is stored by the whole warp: in the (at most two) blocks of the warp's own clips the lanes of that
            // clip store their values and the others zeros; the remaining blocks are all zeros.  All S_h reads of the CTA
            // happened before the exchange barrier above, so overwriting S_h is safe.
            uint32_t pk[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) pk[j] = pack2_bf16(sv[2 * j] * f, sv[2 * j + 1] * f);
            const uint32_t zero6[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              const uint32_t col = tacc + kColP + 12 * c + 6 * hsel;
              if (c == g_lo || c == g_lo + 1) {                           // warp-uniform
                const bool mine = valid && g == c;
                uint32_t z[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) z[j] = mine ? pk[j] : 0u;
                tmem_st4(col, z); tmem_st2(col + 4, z + 4);
              } else {
                tmem_st4(col, zero6); tmem_st2(col + 4, zero6 + 4);
              }
            }
            if (hsel == 1) tmem_st4(tacc + kColP + 60, zero6);
            tmem_st_wait();
            signal_drained();                                             // P_h complete (tensor memory), S_h consumed
            TR(0, 15);
            mbar_wait_all(&bars[B_ACC], ph_acc); ph_acc ^= 1;            // O_h in ACC columns 64..79
            tc_fence_after();
            TR(0, 16);
            float o[8];
            tmem_ld8(tacc + kColO + hsel * 8, o);
            tmem_ld_wait();
            *reinterpret_cast<uint4*>(smem + OFF_AO + (h * 2 + hsel) * 2048 + r * 16) = pack8_bf16(o);
            if (h == 0) signal_drained();                                 // ACC drained: S_1 may overwrite it
            TR(0, 17);
          }
        }
        signal_ready();
        TR(0, 18);
        mbar_wait_all(&bars[B_X], ph_x); ph_x ^= 1;
        tc_fence_after();
        TR(0, 19);

        // LN2 -> tensor memory ; 6 hidden chunks: hidden_c = relu(LN2 W1_c^T + b1_c) ; x += hidden_c W2_c^T
        ln_to_tmem(tx, tacc + kColLn2, hsel, r, 1e-6f, part + 512);
        signal_drained();                  // nothing went to shared memory: no proxy fence
        TR(0, 20);
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          mbar_wait_all(&bars[B_FC1], ph_fc1); ph_fc1 ^= 1;
          tc_fence_after();
          TR(0, 21);
          float v[32];                                                    // this thread's 32 hidden columns of the chunk
          tmem_ld32(tacc + kColFc1 + hsel * 32, v);
          tmem_ld_wait();
          signal_drained();                                               // fc1[c+1] may overwrite the accumulator
          TR(0, 24);
          uint4 hv[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) hv[cc] = relu_pack8(v + cc * 8);
          const int hs = c % kHSlots;                                     // hidden slot
          if (c >= kHSlots) { mbar_wait_all(&bars[B_H + hs], (ph_h >> hs) & 1); ph_h ^= 1u << hs; }   // fc2[c-4] finished reading it
          unsigned char* H = smem + OFF_A + hs * 16384 + hsel * 8192 + r * 16;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(H + cc * 2048) = hv[cc];
          signal_ready();
          TR(0, 22);
        }
        // block end: fc2[3..5] complete (slots 0, 1, 2) => x is final for this block, the operand slots are reusable
#pragma unroll
        for (int hs = 0; hs < kHSlots; ++hs) { mbar_wait_all(&bars[B_H + hs], (ph_h >> hs) & 1); ph_h ^= 1u << hs; }
        tc_fence_after();
        TR(0, 23);
      }

      // ---------------- final LayerNorm (eps 1e-6) + token mean -> pooled[rr][128] ----------------
      // The affine part commutes with the token mean: pooled = gamma * mean_t((x - mu) rstd) + beta.  The normalised rows go
      // through a [128][64] fp32 staging tile (32 KB, 16-B chunks XOR-swizzled by row), one column half per pass.
      {
        float* Y = reinterpret_cast<float*>(smem + OFF_A);
        float mean, rstd;
        row_stats(tx, hsel, r, 1e-6f, part, mean, rstd);
        const float invn = 1.f / (float)(4 * t_n);
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if (hsel == pass) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float v[32];
              tmem_ld32(tx + hsel * 64 + j * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const int kl = j * 32 + i;                                  // column within the half
                if (p.dbg_x != nullptr && r < rows_valid)
                  *reinterpret_cast<float4*>(p.dbg_x + ((size_t)rr0 * tokens + r) * 128 + hsel * 64 + kl) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                float4 o;
                o.x = (v[i] - mean) * rstd; o.y = (v[i + 1] - mean) * rstd;
                o.z = (v[i + 2] - mean) * rstd; o.w = (v[i + 3] - mean) * rstd;
                *reinterpret_cast<float4*>(&Y[r * 64 + ((((kl >> 2) ^ (r & 7)) << 2))]) = o;
              }
            }
          }
          tc_fence_before();
          bar_compute();
          const int col = tid & 63, k = pass * 64 + col;
          const float gam = __ldg(p.norm_w + k), bet = __ldg(p.norm_b + k);
          for (int g = tid >> 6; g < g_cnt; g += 4) {
            float acc = 0.f;
            for (int t = 0; t < tokens; ++t) {
              if (kMasked && t % 6 >= t_n) continue;
              const int row = g * tokens + t;
              acc += Y[row * 64 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3))];
            }
            p.pooled[(size_t)(rr0 + g) * 128 + k] = fmaf(gam, acc * invn, bet);
          }
          bar_compute();     // Y is free again (next pass / next tile's gather)
        }
        TR(0, 30);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef UITK_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 1024) { g_trace[2][blockIdx.x * 4 + 1] = globaltimer_ns(); g_trace[2][blockIdx.x * 4 + 3] += clock64(); }
#endif
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int read_encoder_trace(long long* host_out, int which, int n) {
#ifdef UITK_TRACE
  if (which < 0 || which > 2 || n < 0 || n > 4096) return UITK_EINVAL;
  UITK_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * sizeof(long long), (size_t)which * 4096 * sizeof(long long)));
  return UITK_OK;
#else
  (void)host_out; (void)which; (void)n;
  set_error("libuitk was built without -DUITK_TRACE");
  return UITK_EINVAL;
#endif
}

size_t encoder_tc_bf16_section_bytes(int depth) { return (size_t)kPatchBytes + (size_t)depth * kBlockBytes; }
size_t encoder_tc_block_bytes() { return kBlockBytes; }

size_t encoder_tc_workspace_bytes(int64_t clip_crops, int64_t rows) {
  return align_up((size_t)rows * 128 * 4, 256) + align_up((size_t)clip_crops * 128 * 4, 256);
}

int launch_head_pooled(const float* pooled, int64_t B, int crops, const float* W, const EncoderLayout& lay, int outputdim, int eval_max,
                       float* probs, cudaStream_t s, const uint32_t* c_true, const uint32_t* c_used, const uint32_t* c_min);

int run_encoder_tc(const EncoderArgs& a) {
  const uitk_encoder_cfg& cfg = *a.cfg;
  const bool feat = a.features_out != nullptr;
  const int crops = feat ? 1 : crops_for(a.T, a.target_length);
  const int t_n = time_patches_for(a.T, a.target_length);
  constexpr int kSlots24 = 24;                  // row slots per clip-crop: 4 mel bands x 6 time slots, t_n of them live per band
  const int64_t RR = a.B * crops;
  UITK_REQUIRE(RR * kSlots24 < (1ll << 31) - 256, UITK_EINVAL, "too many token rows for one call; chunk the batch");
  UITK_REQUIRE(t_n >= 1 && t_n <= cfg.grid_t && t_n <= 6, UITK_EINVAL, "%d time patches exceed time_pos_embed length %d", t_n, cfg.grid_t);
  const EncoderLayout lay = make_encoder_layout(cfg);
  const unsigned char* blob = reinterpret_cast<const unsigned char*>(a.blob);
  const float* W = reinterpret_cast<const float*>(blob + sizeof(BlobHeader));
  const size_t bf16_off = sizeof(BlobHeader) + align_up(lay.total_floats * sizeof(float), 1024);

  UITK_REQUIRE(encoder_tc_workspace_bytes(RR, RR * kSlots24) <= a.ws_bytes, UITK_ENOSPACE, "workspace too small: need %zu, have %zu",
               encoder_tc_workspace_bytes(RR, RR * kSlots24), a.ws_bytes);
  unsigned char* ws = reinterpret_cast<unsigned char*>(a.ws);
  float* dbg_x = reinterpret_cast<float*>(ws);
  float* pooled = reinterpret_cast<float*>(ws + align_up((size_t)RR * kSlots24 * 128 * 4, 256));

  TcParams p{};
  p.wts = blob + bf16_off;
  p.pos_tab = W + lay.pos_tab;
  p.bn_scale = W + (feat ? lay.ident_scale : lay.bn_scale); p.bn_shift = W + (feat ? lay.ident_shift : lay.bn_shift);
  p.norm_w = W + lay.norm_w; p.norm_b = W + lay.norm_b;
  p.db = a.db; p.max_pow = a.max_pow; p.no_clamp = feat ? 1 : 0;
  p.T = (int)a.T; p.crops = crops; p.tokens = kSlots24; p.t_n = t_n; p.target = a.target_length;
  p.RR = (int)RR; p.G = 128 / kSlots24; p.num_tiles = (int)((RR + p.G - 1) / p.G); p.depth = cfg.depth;
  p.pooled = pooled;
  p.dbg_x = (feat || (a.debug_taps & 1)) ? dbg_x : nullptr;
  p.c_used = a.cond_used; p.c_min = a.cond_min;

  int dev = 0, sms = 0;
  UITK_CHECK_CUDA(cudaGetDevice(&dev));
  UITK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int resident = 2 * sms;                       // two CTAs per SM
#ifdef UITK_TRACE
  if (const char* e = getenv("UITK_TC_GRID")) resident = atoi(e);      // profiling builds only: e.g. 148 = one CTA per SM
#endif
  const int grid = p.num_tiles < resident ? p.num_tiles : resident;
  if (t_n == 6) {
    UITK_CHECK_CUDA(cudaFuncSetAttribute(encoder_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    encoder_tc_kernel<false><<<grid, kThreads, kSmemBytes, a.stream>>>(p);
  } else {
    UITK_CHECK_CUDA(cudaFuncSetAttribute(encoder_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    encoder_tc_kernel<true><<<grid, kThreads, kSmemBytes, a.stream>>>(p);
  }
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  if (feat)   // tokens after the final LayerNorm, compacted from the 24-slot tile layout (uit.py:395)
    return launch_final_ln(dbg_x, RR, kSlots24, t_n, 6, W + lay.norm_w, W + lay.norm_b, a.features_out, a.stream);
  return launch_head_pooled(pooled, a.B, crops, W, lay, cfg.outputdim, a.eval_avg, a.probs, a.stream, a.max_pow, a.cond_used, a.cond_min);
}

}  // namespace uitk
