// K1: fused log-mel front-end for sm_100a.
//
// Replaces  MelSpectrogram(n_fft 512, win 512, hop 160, center/reflect, power 2, 64 HTK mels) -> 10*log10(max(.,1e-10))
// (models/uit.py:298-308, 455; torchaudio functional.spectrogram :123-144, MelScale :407-419, amplitude_to_DB :390).
// The batch-global top-dB clamp (Q2) is NOT applied here: the kernel only tracks the global maximum.
//
// The B * T frames of the batch are one flat list cut into rounds of 16 consecutive frames (a round may straddle two
// clips, so T = 101 wastes no frame slot); one CTA = 7 consecutive rounds, 256 threads, 3 CTAs per SM (70 KB of shared memory each).
// Per round:
//  * staging: the round's <= 3264 samples arrive with one or two 1-D bulk copies (cp.async.bulk + mbarrier, issued by one
//    lane while the previous round runs its mel phase); only the reflect padding at the clip edges is written by threads.
//  * FFT phase: a frame's 512-point real DFT is a 256-point complex FFT (z[n] = x[2n] + i x[2n+1]) by 16 threads (half a warp),
//    16 complex points each in registers: radix-16 pass, W256 twiddle, 16x16 transpose through padded shared memory
//    (warp-synchronous), second radix-16 pass, real-FFT unpack two bins at a time (k and 256 - k share one partner shuffle),
//    |X|^2 stored as the frame's power row.
//  * mel phase: P[16 frames][257] x W[257][64] on the tensor cores (mma.sync m16n8k8 tf32; both operands split hi + lo, three
//    products, fp32 accumulate: ~2^-21 relative).  Only the 40 8x8 blocks of W that hold filterbank entries are multiplied; they
//    are dealt to the 8 warps as balanced runs at pack time (uitk_common.cuh).  The A fragments come straight out of the
//    bin-interleaved, XOR-swizzled power rows (one conflict-free 16-byte load per block), the dB values go from the
//    accumulator fragments to global memory in 32-byte runs.
// HBM sees every sample once (frames overlap 3.2x; the overlap is absorbed by the staging buffer) and every output once:
// 4*L + 4*64*T algorithmic bytes/clip.  The legacy mma.sync path is deliberate: the contraction is 0.5 % of a tensor core's
// time either way; what it buys is that the 2 x 257 power / weight values per frame no longer cross the LSU twice (the
// CUDA-core version of this phase was 48 % of the kernel's shared-memory wavefronts).
#include "tc_ptx.cuh"
#include "uitk_common.cuh"

namespace uitk {

namespace {

constexpr int kThreads = 256;
constexpr int kFramesPerRound = 16;
constexpr int kSamplesPerRound = UITK_HOP * (kFramesPerRound - 1) + UITK_N_FFT;  // 2912: 16 frames of ONE clip
// A round that straddles a clip boundary stages two sample ranges back to back: the second clip's first frame starts
// where the first clip's last frame ends, i.e. every slot of the second clip is shifted by n_fft - hop = 352 samples.
constexpr int kStraddleShift = UITK_N_FFT - UITK_HOP;                            // 352
constexpr int kStageFloats = kSamplesPerRound + kStraddleShift;                  // 3264
constexpr int kExStride = 288;   // float2 per frame group: 16 x 17 used by the transpose; 576 words = 0 mod 32 (the power rows
                                 // that reuse the tile are XOR-swizzled per frame slot instead)

// cos / sin (2 pi m / 32), m = 0..15
__device__ constexpr float kCos32[16] = {1.f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                                         0.55557023301960222f, 0.38268343236508977f, 0.19509032201612827f, 0.f, -0.19509032201612827f,
                                         -0.38268343236508977f, -0.55557023301960222f, -0.70710678118654752f, -0.83146961230254524f,
                                         -0.92387953251128674f, -0.98078528040323044f};
__device__ constexpr float kSin32[16] = {0.f, 0.19509032201612827f, 0.38268343236508977f, 0.55557023301960222f, 0.70710678118654752f,
                                         0.83146961230254524f, 0.92387953251128674f, 0.98078528040323044f, 1.f, 0.98078528040323044f,
                                         0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960222f,
                                         0.38268343236508977f, 0.19509032201612827f};

// Complex arithmetic on packed fp32x2 (FADD2 / FMUL2 / FFMA2): a complex add is ONE instruction, a complex multiply TWO
// (ptxas folds the (x,x) / (y,y) broadcasts and the (-w.y, w.x) swizzle into operand selectors).
using tc::add2; using tc::fma2; using tc::mul2; using tc::sub2;
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return fma2(make_float2(a.y, a.y), make_float2(-w.y, w.x), mul2(make_float2(a.x, a.x), w));
}

// forward radix-4 butterfly (e^{-2 pi i nk/4})
__device__ __forceinline__ void fft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), u = csub(b, d);
  const float2 us = make_float2(u.y, u.x);                        // (b - d) * (-i) = (u.y, -u.x)
  a = cadd(t0, t2); c = csub(t0, t2);
  b = fma2(us, make_float2(1.f, -1.f), t1);
  d = fma2(us, make_float2(-1.f, 1.f), t1);
}

// in-register forward 16-point DFT, natural order in and out
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // twiddles W16^(n1*k1): v[n1 + 4*k1]
  v[1 + 4] = cmul(v[1 + 4], make_float2(c1, -s1));    // W^1
  v[1 + 8] = cmul(v[1 + 8], make_float2(h, -h));      // W^2
  v[1 + 12] = cmul(v[1 + 12], make_float2(s1, -c1));  // W^3
  v[2 + 4] = cmul(v[2 + 4], make_float2(h, -h));      // W^2
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);    // W^4 = -i
  v[2 + 12] = cmul(v[2 + 12], make_float2(-h, -h));   // W^6
  v[3 + 4] = cmul(v[3 + 4], make_float2(s1, -c1));    // W^3
  v[3 + 8] = cmul(v[3 + 8], make_float2(-h, -h));     // W^6
  v[3 + 12] = cmul(v[3 + 12], make_float2(-c1, s1));  // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // V[k1 + 4*k2] sits in v[4*k1 + k2]: 4x4 register transpose
  float2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[2]; v[2] = v[8]; v[8] = t;
  t = v[3]; v[3] = v[12]; v[12] = t;
  t = v[6]; v[6] = v[9]; v[9] = t;
  t = v[7]; v[7] = v[13]; v[13] = t;
  t = v[11]; v[11] = v[14]; v[14] = t;
}

// TIn = float (the reference's input contract) or int16_t (PCM ingest: x = pcm / 32768, dataset.py:44-46 /
// torchaudio.load normalisation; the 2^-15 scale is folded into the window, which is exact, so both instantiations
// produce bit-identical results for the same audio).
// Where the 16 frame slots of round r come from.  straddle == 1 (T >= 16): slots are 16 consecutive frames of the flat
// list, nA of them in clip cA (frames tA ..) and the rest in clip cA + 1 (frames 0 ..).  Otherwise rounds never cross a
// clip (ceil(T / 16) rounds per clip, dead slots in the last one).
struct Round {
  long long cA;     // first clip
  int tA, nA, nB;   // first frame in cA, live slots in cA, live slots in cA + 1
  int pad;
};
// (clip, first frame) of a round; stepped from round to round without divisions (one 64-bit division per CTA)
struct RoundPos {
  long long c;
  int t;
};
__device__ __forceinline__ RoundPos round_pos(long long r, int T, int straddle, int rounds_per_clip) {
  RoundPos P;
  if (straddle) {
    const long long F0 = r * kFramesPerRound;
    P.c = F0 / T;
    P.t = (int)(F0 - P.c * T);
  } else {
    P.c = r / rounds_per_clip;
    P.t = (int)(r - P.c * rounds_per_clip) * kFramesPerRound;
  }
  return P;
}
__device__ __forceinline__ RoundPos next_pos(RoundPos P, int T, int straddle) {
  P.t += kFramesPerRound;
  if (P.t >= T) { P.t = straddle ? P.t - T : 0; ++P.c; }      // T >= 16 in straddle mode: at most one wrap
  return P;
}
__device__ __forceinline__ Round round_at(RoundPos P, int T, long long B, int straddle) {
  Round R;
  R.cA = P.c; R.tA = P.t;
  R.nA = min(kFramesPerRound, T - P.t);
  R.nB = (straddle && P.c + 1 < B) ? kFramesPerRound - R.nA : 0;
  return R;
}

struct SmemLayout {
  float window[512];
  float x[kStageFloats];             // the round's samples; round r+1 is bulk-copied in while round r runs its mel phase
  float2 ex[kFramesPerRound * kExStride];   // per frame group: 16x17 transpose tile, then the frame's 257 powers (swizzled)
  float4 part[kMelSlots][32];        // partial mel sums of a split octet: producer warp(s) -> owner warp
  MelBlk blk[kMelOctets][kMelWarpBlocks + 1];      // } contiguous, same order as in the blob: one copy loop
  float2 w[kMelSmemBlocks][32];                    // } mel weight fragments
  float red[16];
  Round desc[2];                     // where the frame slots of the round come from (written by warp 0 with the staging)
  uint64_t full;                     // mbarrier: "x landed"
};
static_assert(sizeof(MelBlk) * kMelOctets * (kMelWarpBlocks + 1) % 16 == 0, "block lists are copied as uint4");

// D(16x8, fp32) += A(16x8, tf32, row) * B(8x8, tf32, col).  Fragments (lane = 4 * gid + tig): a0 = A[gid][tig], a1 = A[gid+8][tig],
// a2 = A[gid][tig+4], a3 = A[gid+8][tig+4]; b0 = B[tig][gid], b1 = B[tig+4][gid]; d0/d1 = D[gid][2 tig, +1], d2/d3 = D[gid+8][..].
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 10 log10(max(x, 1e-10)) = 3.0103 log2(.): the argument is a normal number, so the flush-to-zero form (no denormal rescue) is exact
// to the same ~1e-6 dB as __log2f
__device__ __forceinline__ float power_to_db(float x) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(x, 1e-10f)));
  return 3.01029995663981195f * l;
}

template <typename TIn>
__global__ void __launch_bounds__(kThreads, 3)
logmel_kernel(const TIn* __restrict__ wav, long long B, long long L, long long ld, int T, int t0, long long out_bs, long long out_ms,
              long long num_rounds, int straddle, int rpc,
              const FrontendBlob* __restrict__ blob, float* __restrict__ db, uint32_t* __restrict__ max_pow,
              uint32_t* __restrict__ min_pow) {
  // T = frames computed per clip: frames t0 .. t0 + T - 1 of the clip's 1 + L/160.  Output element (clip, mel, frame t) goes
  // to db[clip * out_bs + mel * out_ms + t].
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);

  const int tid = threadIdx.x;
  const int g = (tid >> 5) + ((tid & 16) >> 1);   // frame slot in the round: warp p runs slots p (lanes 0-15) and p + 8 (lanes 16-31)
  const int j = tid & 15;      // lane within the frame group
  const int rounds_per_clip = (T + kFramesPerRound - 1) / kFramesPerRound;
  const long long r_begin = (long long)blockIdx.x * rpc;
  const long long r_end = r_begin + rpc < num_rounds ? r_begin + rpc : num_rounds;

  constexpr bool kPcm = sizeof(TIn) == 2;
  constexpr int kVec = 16 / (int)sizeof(TIn);          // samples per 16 bytes
  if (tid == 0) {
    tc::mbar_init(&S.full, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 512; i += kThreads) S.window[i] = kPcm ? blob->window[i] * (1.f / 32768.f) : blob->window[i];
  // block lists + weight fragments (the fragments only if they fit: a denser filterbank than HTK/64 is read from global memory)
  const bool w_in_smem = blob->n_blocks + 1 <= kMelSmemBlocks;
  {
    const int n16 = (int)(sizeof(S.blk) / 16) + (w_in_smem ? (blob->n_blocks + 1) * 32 * (int)sizeof(float2) / 16 : 0);
    const uint4* src = reinterpret_cast<const uint4*>(&blob->mel_blk[0][0]);
    uint4* dst = reinterpret_cast<uint4*>(&S.blk[0][0]);
    for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  float tmax = 0.f, tmin = INFINITY;
  // thread-constant twiddles kept in registers: W256^(j*2^i) (the other powers are products of these) and W512^j
  const float2 w1 = blob->tw256[1 * 16 + j], w2 = blob->tw256[2 * 16 + j], w4 = blob->tw256[4 * 16 + j], w8 = blob->tw256[8 * 16 + j];
  const float2 wj512 = blob->tw512[j];
  const int Li = (int)L;

  // Stage the samples of round r into S.x (warp 0 only).  A round reads one sample range per clip it touches (two when it
  // straddles a clip boundary); the part of a range that lies inside its clip is ONE 1-D bulk copy (cp.async.bulk, completion
  // on S.full) issued by lane 0, the few samples of the reflect padding at the clip edges (and everything, if the clip's
  // rows are not 16-byte aligned) are written by the warp's lanes with the reflect index map (no edge repeat).
  auto stage = [&](long long r, RoundPos P) {
    if (r >= r_end || tid >= 32) return;
    const Round R = round_at(P, T, B, straddle);
    if (tid == 0) S.desc[(r - r_begin) & 1] = R;
    TIn* dst = reinterpret_cast<TIn*>(S.x);                // raw samples (PCM uses half of the buffer)
    const int lenA = UITK_HOP * (R.nA - 1) + UITK_N_FFT;
    int len[2], s0[2], lo[2], hi[2];
    len[0] = lenA; len[1] = R.nB > 0 ? UITK_HOP * (R.nB - 1) + UITK_N_FFT : 0;
    s0[0] = (t0 + R.tA) * UITK_HOP - UITK_N_FFT / 2; s0[1] = t0 * UITK_HOP - UITK_N_FFT / 2;
    uint32_t tx = 0;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const TIn* clip = wav + (R.cA + p) * ld;
      lo[p] = max(0, -s0[p]);                                 // s0 is a multiple of 32 samples: lo keeps the 16-byte alignment
      hi[p] = min(len[p], Li - s0[p]);
      hi[p] = lo[p] + ((hi[p] - lo[p]) & ~(kVec - 1));
      if (hi[p] < lo[p] || (reinterpret_cast<uintptr_t>(clip) & 15) != 0) lo[p] = hi[p] = 0;
      tx += (uint32_t)(hi[p] - lo[p]) * (uint32_t)sizeof(TIn);
    }
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(&S.full, tx);
#pragma unroll
      for (int p = 0; p < 2; ++p)
        if (hi[p] > lo[p])
          tc::bulk_g2s(dst + (p ? lenA : 0) + lo[p], wav + (R.cA + p) * ld + s0[p] + lo[p], (uint32_t)(hi[p] - lo[p]) * (uint32_t)sizeof(TIn),
                       &S.full);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (hi[p] - lo[p] == len[p]) continue;                  // the usual case: nothing left for the scalar path
      const TIn* clip = wav + (R.cA + p) * ld;
      TIn* d = dst + (p ? lenA : 0);
      for (int i = tid; i < len[p] - (hi[p] - lo[p]); i += 32) {
        const int ii = i < lo[p] ? i : i + (hi[p] - lo[p]);
        int id = s0[p] + ii;
        if (id < 0) id = -id;
        if (id >= Li) id = 2 * (Li - 1) - id;
        d[ii] = (id >= 0 && id < Li) ? __ldg(clip + id) : TIn(0);
      }
    }
  };
  RoundPos pos = round_pos(r_begin, T, straddle, rounds_per_clip);
  stage(r_begin, pos);

  // mel phase: lane = 4 * gid + tig (mma fragment coordinates); this warp's segments of the filterbank (uitk_common.cuh)
  const int lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int mel_nblk = blob->mel_nblk[warp];
  const long long orow_tig = (long long)(2 * tig) * out_ms;
  // Power rows: the two frame slots p, p + 8 of warp p share one row, bin-interleaved, that reuses the warp's first transpose tile:
  // P[p + 8 h][k] at float 2 * (k ^ swz(p)) + h.  A 16-byte load at bins (b, b + 1) is then exactly the A fragment of mma rows
  // p / p + 8 for logical k = tig / tig + 4 <-> bins b = 8G + 2 tig, b + 1; the stores of a warp's two frame groups hit even / odd
  // banks; the XOR (8 bins for odd p) keeps the two rows of a quarter-warp load in different bank halves.
  // k = j + 16 m  ->  (j ^ swz) + 16 m;  256 - k = (16 - j) + 16 (15 - m)  ->  ((16 - j) ^ swz) + 16 (15 - m)   (swz = 0 | 8, also for j = 0)
  float* prow_w = reinterpret_cast<float*>(S.ex + (g & 7) * kExStride) + (g >> 3);
  float* const pk_base = prow_w + 2 * (j ^ ((g & 1) << 3));
  float* const pn_base = prow_w + 2 * ((16 - j) ^ ((g & 1) << 3));
  const float* const prow_r = reinterpret_cast<const float*>(S.ex + gid * kExStride) + 4 * tig;   // group G at + 16 * (G ^ (gid & 1))
  const float2* const wtab = (w_in_smem ? &S.w[0][0] : blob->mel_frag) + lane;

  for (long long r = r_begin; r < r_end; ++r) {
    tc::mbar_wait(&S.full, (uint32_t)((r - r_begin) & 1));
    __syncthreads();   // scalar-staged edge samples of S.x and the round descriptor visible; previous round's power rows are consumed
    const TIn* sx = reinterpret_cast<const TIn*>(S.x);
    const Round R = S.desc[(r - r_begin) & 1];
    if (tid < 32) pos = next_pos(pos, T, straddle);    // only the staging warp tracks the position
    const int n_live = R.nA + R.nB;
    const bool warp_live = (g & 7) < n_live;           // warp-uniform: the warp's first frame group is live

    if (warp_live) {
    // ---- windowed load: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1], n = j + 16 m
    float2 v[16];
    const TIn* xf = sx + g * UITK_HOP + (g >= R.nA ? kStraddleShift : 0);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int n = j + 16 * m;
      float2 xx;
      if (kPcm) {
        const short2 xs = *reinterpret_cast<const short2*>(xf + 2 * n);
        xx = make_float2((float)xs.x, (float)xs.y);
      } else {
        xx = *reinterpret_cast<const float2*>(xf + 2 * n);
      }
      const float2 ww = *reinterpret_cast<const float2*>(S.window + 2 * n);
      v[m] = mul2(xx, ww);
    }
    fft16(v);                                   // over m -> k1
    {   // v[k1] *= W256^(j*k1), powers composed from w1, w2, w4, w8 (<= 3 roundings)
      const float2 w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
      v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
      v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7); v[8] = cmul(v[8], w8);
      v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
      v[12] = cmul(v[12], cmul(w8, w4)); v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6));
      v[15] = cmul(v[15], cmul(w8, w7));
    }
    float2* e = S.ex + g * kExStride;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) e[k1 * 17 + j] = v[k1];
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = e[j * 17 + n1];   // this thread now owns k1 = j
    __syncwarp();
    fft16(v);                                   // over n1 -> k2 ; v[k2] = Z[j + 16 k2]

    // ---- real-FFT unpack + power, two bins per step.  With A = Z[k], B = conj(Z[256-k]), G = -i W512^k:
    //   2 X[k] = (A + B) + G (A - B)        2 conj(X[256-k]) = (A + B) - G (A - B)
    // so the thread that owns k = j + 16 m (m < 8) produces the powers of k AND of 256 - k from one partner value:
    // Z[256-k] is register 15 - m of lane 16 - j of this frame group (lane 0: its own register (16 - m) & 15), fetched with ONE
    // shuffle pair.  Lane 0 adds k = 128 (its own partner).  G = (-i W512^j) * W32^m with compile-time W32^m.  Powers are kept
    // as 4 |X|^2; the factor 1/4 is folded into the mel weights (exact: power of two).
    const float2 G0 = make_float2(wj512.y, -wj512.x);             // -i * W512^j
    const int pl = (16 - j) & 15;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float2 zk = v[m];
      float2 src = v[15 - m];
      if (j == 0) src = v[(16 - m) & 15];
      float2 zn;
      zn.x = __shfl_sync(0xffffffffu, src.x, pl, 16);
      zn.y = __shfl_sync(0xffffffffu, src.y, pl, 16);
      const float2 S2 = fma2(zn, make_float2(1.f, -1.f), zk);     // A + B
      const float2 D2 = fma2(zn, make_float2(-1.f, 1.f), zk);     // A - B
      const float2 Gm = m == 0 ? G0 : cmul(G0, make_float2(kCos32[m], -kSin32[m]));
      const float2 Tm = cmul(D2, Gm);
      const float2 Xp = cadd(S2, Tm), Xn = csub(S2, Tm);
      pk_base[32 * m] = fmaf(Xp.x, Xp.x, Xp.y * Xp.y);                      // 4 |X[k]|^2,       k = j + 16 m
      pn_base[32 * (15 - m)] = fmaf(Xn.x, Xn.x, Xn.y * Xn.y);               // 4 |X[256-k]|^2
    }
    if (j == 0) pk_base[32 * 8] = 4.f * fmaf(v[8].x, v[8].x, v[8].y * v[8].y);   // X[128] = conj(Z[128])
    }   // warp_live
    const MelBlk* blk = S.blk[warp];
    MelBlk me = blk[0];                                // first weight block of the mel phase: loaded under the barrier wait
    float2 wf = wtab[me.y * 32];
    __syncthreads();
    stage(r + 1, pos);                                 // S.x is free: the next round's samples land under the mel phase

    // ---- mel projection on the tensor cores + dB.  D[16 frame slots][8 mel bins of an octet] = P[16][8 bins] * W[8 bins][8]
    // summed over the segment's bin groups; P and W as tf32 hi + lo, three products (hi*hi + lo*hi + hi*lo: ~2^-21 relative).
    // Every warp walks its list of weight blocks (uitk_common.cuh); a block with the fin bit ends a run of one octet.
    {
      long long orow[2];                               // output offsets (mel 2 tig of octet 0) of this lane's two frame slots
      const long long rowA = R.cA * out_bs + t0 + R.tA, rowB = (R.cA + 1) * out_bs + t0 - R.nA;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = gid + 8 * h;                     // mma rows gid / gid + 8 are frame slots gid / gid + 8
        orow[h] = (s < R.nA ? rowA : rowB) + s + orow_tig;
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acl[4] = {0.f, 0.f, 0.f, 0.f}, acw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int b = 0; b < mel_nblk; ++b) {
        const MelBlk nx = blk[b + 1];
        const float2 wn = wtab[nx.y * 32];                           // next block's weights under this block's math
        // (P[gid][k], P[gid + 8][k], P[gid][k + 1], P[gid + 8][k + 1]), k = 8 G + 2 tig  ==  (a0, a1, a2, a3)
        const float4 pa = *reinterpret_cast<const float4*>(prow_r + 16 * ((me.x & 0xff) ^ (gid & 1)));
        const uint32_t a0 = __float_as_uint(pa.x) & 0xffffe000u, a1 = __float_as_uint(pa.y) & 0xffffe000u;
        const uint32_t a2 = __float_as_uint(pa.z) & 0xffffe000u, a3 = __float_as_uint(pa.w) & 0xffffe000u;
        const uint32_t l0 = __float_as_uint(pa.x - __uint_as_float(a0)), l1 = __float_as_uint(pa.y - __uint_as_float(a1));
        const uint32_t l2 = __float_as_uint(pa.z - __uint_as_float(a2)), l3 = __float_as_uint(pa.w - __uint_as_float(a3));
        const uint32_t b0 = __float_as_uint(wf.x) & 0xffffe000u, b1 = __float_as_uint(wf.y) & 0xffffe000u;
        const uint32_t c0 = __float_as_uint(wf.x - __uint_as_float(b0)), c1 = __float_as_uint(wf.y - __uint_as_float(b1));
        mma_tf32(acc, a0, a1, a2, a3, b0, b1);
        mma_tf32(acl, l0, l1, l2, l3, b0, b1);
        mma_tf32(acw, a0, a1, a2, a3, c0, c1);
        if (me.x & 0x100) {                            // end of a run (warp-uniform)
          const int role = (me.x >> 9) & 3, oct = (me.x >> 11) & 7, aux = me.x >> 14;
          float m[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { m[i] = acc[i] + (acl[i] + acw[i]); acc[i] = acl[i] = acw[i] = 0.f; }
          if (role == kMelProducer) {                  // early part of a split octet: hand the partial sums to the owner
            S.part[aux & 7][lane] = make_float4(m[0], m[1], m[2], m[3]);
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + oct), "r"(32 + 32 * (aux >> 3)) : "memory");
          } else {
            if (role == kMelOwner) {
              asm volatile("bar.sync %0, %1;" ::"r"(1 + oct), "r"(32 + 32 * (aux & 3)) : "memory");
              const float4 q = S.part[(aux >> 2) & 7][lane];
              m[0] += q.x; m[1] += q.y; m[2] += q.z; m[3] += q.w;
              if ((aux & 3) == 2) {
                const float4 q2 = S.part[(aux >> 5) & 7][lane];
                m[0] += q2.x; m[1] += q2.y; m[2] += q2.z; m[3] += q2.w;
              }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (gid + 8 * h < n_live) {
                float* o = db + orow[h] + (long long)(8 * oct) * out_ms;
                const float m0 = m[2 * h], m1 = m[2 * h + 1];
                tmax = fmaxf(tmax, fmaxf(m0, m1)); tmin = fminf(tmin, fminf(m0, m1));
                o[0] = power_to_db(m0);
                o[out_ms] = power_to_db(m1);
              }
            }
          }
        }
        me = nx; wf = wn;
      }
    }
  }

  // ---- global max of the mel power (non-negative: uint order == float order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
  }
  __syncthreads();
  if ((tid & 31) == 0) { S.red[tid >> 5] = tmax; S.red[8 + (tid >> 5)] = tmin; }
  __syncthreads();
  if (tid == 0 && r_end > r_begin) {
    float m = S.red[0], mn = S.red[8];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) { m = fmaxf(m, S.red[w]); mn = fminf(mn, S.red[8 + w]); }
    atomicMax(max_pow, __float_as_uint(m));
    if (min_pow != nullptr) atomicMin(min_pow, __float_as_uint(fmaxf(mn, 0.f)));
  }
}

__global__ void clamp_db_kernel(float* __restrict__ db, long long n, const uint32_t* __restrict__ max_pow,
                                const uint32_t* __restrict__ min_pow, float top_db) {
  const float cutoff = 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(*max_pow), 1e-10f)) - top_db;   // same map as the kernel
  // the smallest dB value of the batch is this map of the minimum power word: nothing to clamp -> nothing to read or write
  if (min_pow != nullptr && 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(*min_pow), 1e-10f)) >= cutoff) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) db[i] = fmaxf(db[i], cutoff);
}

}  // namespace

template <typename TIn>
static int launch_logmel_t(const TIn* wav, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db,
                           uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s, int64_t t0 = 0, int64_t tn = -1, int64_t out_bs = -1,
                           int64_t out_ms = -1) {
  const int64_t Tall = 1 + L / UITK_HOP;
  const int64_t T = tn < 0 ? Tall : tn;                                      // frames computed per clip
  if (out_ms < 0) out_ms = Tall;
  if (out_bs < 0) out_bs = 64 * Tall;
  if (T == 0 || B == 0) return UITK_OK;
  const size_t smem = sizeof(SmemLayout);
  static_assert(3 * (sizeof(SmemLayout) + 1024) <= 228 * 1024, "three CTAs per SM");
  UITK_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // rounds of 16 consecutive frames of the flat frame list (T >= 16), else per-clip rounds
  const int straddle = T >= kFramesPerRound ? 1 : 0;
  const int64_t num_rounds = straddle ? (B * T + kFramesPerRound - 1) / kFramesPerRound : B * ((T + kFramesPerRound - 1) / kFramesPerRound);
  // 7 consecutive rounds per CTA.  Measured (scripts/logmel_time.py): persistent CTAs (one range per resident CTA) are
  // 5 % SLOWER - the three CTAs of an SM then run their FFT (FMA-bound) and mel (LSU-bound) phases in lock step, while
  // ordinary launch order staggers them; 4 / 14 / 28 rounds per CTA are 1-4 % slower than 7.
  const int rpc = 7;
  const int64_t grid64 = (num_rounds + rpc - 1) / rpc;
  UITK_REQUIRE(grid64 < (1ll << 31), UITK_EINVAL, "too many frames for one launch");
  const int grid = (int)grid64;
  logmel_kernel<TIn><<<grid, kThreads, smem, s>>>(wav, (long long)B, (long long)L, (long long)ld, (int)T, (int)t0, (long long)out_bs,
                                                  (long long)out_ms, (long long)num_rounds, straddle, rpc, blob, db, max_pow, min_pow);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

namespace {

// db_w[w][m][t] = G[m][w * r + t] for the interior frames t in [2, Tw - 2) of every window (sliding-window reuse, see api.cu)
__global__ void window_gather_kernel(const float* __restrict__ G, long long U, float* __restrict__ dbw, long long W, int Tw, int r) {
  const int inner = Tw - 4;
  const long long total = W * 64 * inner;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int t = (int)(i % inner) + 2;
    const long long wm = i / inner;
    const int m = (int)(wm & 63);
    const long long w = wm >> 6;
    dbw[(w * 64 + m) * Tw + t] = __ldg(G + (long long)m * U + w * r + t);
  }
}

}  // namespace

int launch_logmel_frames(const float* wav, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db, int64_t t0, int64_t tn,
                         int64_t out_bs, int64_t out_ms, uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s) {
  return launch_logmel_t<float>(wav, B, L, ld, blob, db, max_pow, min_pow, s, t0, tn, out_bs, out_ms);
}

int launch_window_gather(const float* G, int64_t U, float* dbw, int64_t W, int Tw, int r, cudaStream_t s) {
  if (W == 0 || Tw <= 4) return UITK_OK;
  const long long total = W * 64 * (long long)(Tw - 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  window_gather_kernel<<<(int)blocks, 256, 0, s>>>(G, (long long)U, dbw, (long long)W, Tw, r);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

int launch_logmel(const float* wav, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db,
                  uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s) {
  return launch_logmel_t<float>(wav, B, L, ld, blob, db, max_pow, min_pow, s);
}

int launch_logmel_i16(const int16_t* pcm, int64_t B, int64_t L, int64_t ld, const FrontendBlob* blob, float* db,
                      uint32_t* max_pow, uint32_t* min_pow, cudaStream_t s) {
  return launch_logmel_t<int16_t>(pcm, B, L, ld, blob, db, max_pow, min_pow, s);
}

int launch_clamp_db(float* db, int64_t n, const uint32_t* max_pow, const uint32_t* min_pow, float top_db, cudaStream_t s) {
  if (n == 0) return UITK_OK;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 16) blocks = 148 * 16;
  clamp_db_kernel<<<blocks, 256, 0, s>>>(db, (long long)n, max_pow, min_pow, top_db);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

}  // namespace uitk
