// K1: fused log-mel front-end for sm_100a.
//
// Replaces  MelSpectrogram(n_fft 512, win 512, hop 160, center/reflect, power 2, 64 HTK mels) -> 10*log10(max(.,1e-10))
// (models/uit.py:298-308, 455; torchaudio functional.spectrogram :123-144, MelScale :407-419, amplitude_to_DB :390).
// The batch-global top-dB clamp (Q2) is NOT applied here: the kernel only tracks the global maximum.
//
// The B * T frames of the batch are one flat list cut into rounds of 16 consecutive frames (a round may straddle two
// clips, so T = 101 wastes no frame slot); one CTA = 7 consecutive rounds, 3 CTAs per SM (75.9 KB of shared memory each).
// A frame's 512-point real DFT is computed as a
// 256-point complex FFT (z[n] = x[2n] + i x[2n+1]) by 16 threads (half a warp), each holding 16 complex points in
// registers: radix-16 pass over registers, W256 twiddle, 16x16 transpose through shared memory (warp-synchronous, padded
// rows), second radix-16 pass, real-FFT unpack, |X|^2, sparse mel (<=2 non-zero filters per bin -> packed ranges), dB.
// The samples of round r+1 are staged with 16-byte cp.async while round r computes.  HBM sees every sample once (frames
// overlap 3.2x; the overlap is absorbed by the staging buffer / L1) and every output once: 4*L + 4*64*T algorithmic
// bytes/clip.
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
constexpr int kExStride = 280;   // float2 per frame group: 16 x 17 used; 560 words = 16 mod 32 so that the two
                                 // frame groups of a warp use complementary banks for 32-bit accesses

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

struct SmemLayout {
  float window[512];
  int mel_lo[64], mel_iters[4], mel_qoff[4];
  float x[2][kStageFloats];          // double buffered: cp.async stages round r+1 while round r computes
  float2 ex[kFramesPerRound * kExStride];
  float out[64 * 17];
  float red[16];
  // followed by mel_w[n_weights]
};

// TIn = float (the reference's input contract) or int16_t (PCM ingest: x = pcm / 32768, dataset.py:44-46 /
// torchaudio.load normalisation; the 2^-15 scale is folded into the window, which is exact, so both instantiations
// produce bit-identical results for the same audio).
// Where the 16 frame slots of round r come from.  straddle == 1 (T >= 16): slots are 16 consecutive frames of the flat
// list, nA of them in clip cA (frames tA ..) and the rest in clip cA + 1 (frames 0 ..).  Otherwise rounds never cross a
// clip (ceil(T / 16) rounds per clip, dead slots in the last one).
struct Round {
  long long cA;     // first clip
  int tA, nA, nB;   // first frame in cA, live slots in cA, live slots in cA + 1
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
  float* s_melw = reinterpret_cast<float*>(smem_raw + sizeof(SmemLayout));

  const int tid = threadIdx.x;
  const int g = tid >> 4;      // frame slot in the round
  const int j = tid & 15;      // lane within the frame group
  const int rounds_per_clip = (T + kFramesPerRound - 1) / kFramesPerRound;
  const long long r_begin = (long long)blockIdx.x * rpc;
  const long long r_end = r_begin + rpc < num_rounds ? r_begin + rpc : num_rounds;

  constexpr bool kPcm = sizeof(TIn) == 2;
  constexpr int kVec = 16 / (int)sizeof(TIn);          // samples per 16-byte cp.async
  for (int i = tid; i < 512; i += kThreads) S.window[i] = kPcm ? blob->window[i] * (1.f / 32768.f) : blob->window[i];
  if (tid < 64) S.mel_lo[tid] = blob->mel_lo[tid];
  if (tid < 4) { S.mel_iters[tid] = blob->mel_iters[tid]; S.mel_qoff[tid] = blob->mel_qoff[tid]; }
  const int nw = blob->n_weights;
  for (int i = tid; i < nw; i += kThreads) s_melw[i] = blob->mel_w[i];

  float tmax = 0.f, tmin = INFINITY;
  // thread-constant twiddles kept in registers: W256^(j*2^i) (the other powers are products of these) and W512^j
  const float2 w1 = blob->tw256[1 * 16 + j], w2 = blob->tw256[2 * 16 + j], w4 = blob->tw256[4 * 16 + j], w8 = blob->tw256[8 * 16 + j];
  const float2 wj512 = blob->tw512[j];
  const int Li = (int)L;

  // Stage the samples of round r into S.x[buf]: 16-byte cp.async for groups that lie inside their clip (and are 16-B
  // aligned), synchronous loads with the reflect index map (no edge repeat) for the few groups at the clip edges.
  auto stage = [&](long long r, RoundPos P, int buf) {
    if (r < r_end) {
      const Round R = round_at(P, T, B, straddle);
      const int lenA = UITK_HOP * (R.nA - 1) + UITK_N_FFT;                       // floats of clip cA's range
      const int total = R.nB > 0 ? lenA + UITK_HOP * (R.nB - 1) + UITK_N_FFT : lenA;
      const TIn* clipA = wav + R.cA * ld;
      const TIn* clipB = clipA + ld;
      const bool okA = (reinterpret_cast<uintptr_t>(clipA) & 15) == 0, okB = (reinterpret_cast<uintptr_t>(clipB) & 15) == 0;
      const int s0A = (t0 + R.tA) * UITK_HOP - UITK_N_FFT / 2;
      const int s0B = t0 * UITK_HOP - UITK_N_FFT / 2;
      TIn* dst = reinterpret_cast<TIn*>(S.x[buf]);           // raw samples (PCM uses half of the buffer)
      for (int i = tid * kVec; i < total; i += kThreads * kVec) {
        const bool inA = i < lenA;                           // lenA is a multiple of kVec: a group never spans both clips
        const TIn* clip = inA ? clipA : clipB;
        const int idx = inA ? s0A + i : s0B + (i - lenA);
        if ((inA ? okA : okB) && idx >= 0 && idx + kVec - 1 < Li) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(clip + idx)
                       : "memory");
        } else {
#pragma unroll
          for (int e = 0; e < kVec; ++e) {
            int id = idx + e;
            if (id < 0) id = -id;
            if (id >= Li) id = 2 * (Li - 1) - id;
            dst[i + e] = (id >= 0 && id < Li) ? __ldg(clip + id) : TIn(0);
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  RoundPos pos = round_pos(r_begin, T, straddle, rounds_per_clip);
  stage(r_begin, pos, 0);

  int buf = 0;
  for (long long r = r_begin; r < r_end; ++r, buf ^= 1) {
    const RoundPos pos_next = next_pos(pos, T, straddle);
    stage(r + 1, pos_next, buf ^ 1);                   // buffer last read two barriers ago
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();   // S.x[buf] (and the constants) visible; previous round's S.out readers are done
    const TIn* sx = reinterpret_cast<const TIn*>(S.x[buf]);
    const Round R = round_at(pos, T, B, straddle);
    pos = pos_next;
    const bool live = g < R.nA + R.nB;
    const bool warp_live = (g & ~1) < R.nA + R.nB;     // warp-uniform: the warp's first frame group is live

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

    // ---- real-FFT unpack + power: X[k] = E[k] + W512^k O[k], k = j + 16 m.  The partner Z[256-k] lives in lane
    // (16-j) of this frame group at register 15-m (lane 0: its own register (16-m)&15): one shuffle, static indices.
    // W512^k = W512^j * W32^m with W32^m compile-time constants.
    // 2 X[k] = (A + B) + G (A - B) with A = Z[k], B = conj(Z[256-k]), G = -i W512^k; G[m+1] = G[m] * W32 (<= 16 roundings).
    // The factor 1/4 of |X|^2 is folded into the packed mel weights (exact: power of two).
    float p[16];
    float2 G = make_float2(wj512.y, -wj512.x);                     // -i * W512^j
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 zk = v[m];
      float2 zn;
      zn.x = __shfl_sync(0xffffffffu, v[15 - m].x, (16 - j) & 15, 16);
      zn.y = __shfl_sync(0xffffffffu, v[15 - m].y, (16 - j) & 15, 16);
      if (j == 0) zn = v[(16 - m) & 15];
      const float2 S2 = fma2(zn, make_float2(1.f, -1.f), zk);     // A + B
      const float2 D2 = fma2(zn, make_float2(-1.f, 1.f), zk);     // A - B
      const float2 X2 = cadd(S2, cmul(D2, G));
      p[m] = fmaf(X2.x, X2.x, X2.y * X2.y);                       // 4 |X[k]|^2
      G = cmul(G, make_float2(kCos32[1], -kSin32[1]));
    }
    __syncwarp();
    float* pf = reinterpret_cast<float*>(e);
#pragma unroll
    for (int m = 0; m < 16; ++m) pf[j + 16 * m] = p[m];
    if (j == 0) {
      const float ny = 2.f * (v[0].x - v[0].y);   // X[256] = Re Z0 - Im Z0 (x2: powers are kept as 4 |X|^2)
      pf[256] = ny * ny;
    }
    __syncwarp();

    // ---- sparse mel + dB
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = j + 16 * q;
      const int iters = S.mel_iters[q];
      const float4* wq = reinterpret_cast<const float4*>(s_melw + S.mel_qoff[q]) + j;     // lane-interleaved weights
      const float4* pq = reinterpret_cast<const float4*>(pf + S.mel_lo[m]);               // 4-aligned range start
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int i = 0; i < iters; ++i) {            // uniform trip count per group; short ranges carry zero weights
        const float4 w4 = wq[i * 16];
        const float4 p4 = pq[i];
        a0 = fmaf(w4.x, p4.x, a0); a1 = fmaf(w4.y, p4.y, a1);
        a2 = fmaf(w4.z, p4.z, a2); a3 = fmaf(w4.w, p4.w, a3);
      }
      const float acc = (a0 + a1) + (a2 + a3);
      if (live) { tmax = fmaxf(tmax, acc); tmin = fminf(tmin, acc); }
      S.out[m * 17 + g] = 3.01029995663981195f * __log2f(fmaxf(acc, 1e-10f));   // 10 log10(x); |err| ~1e-6 dB
    }
    }   // warp_live
    __syncthreads();
    {   // S.out [64 mel][16 slots] -> db: this thread stores slot gg = tid & 15 of mel rows (tid >> 4) + 16 it
      const int gg = tid & 15;
      if (gg < R.nA + R.nB) {
        const bool inA = gg < R.nA;
        float* o = db + (inA ? R.cA : R.cA + 1) * out_bs + (tid >> 4) * out_ms + t0 + (inA ? R.tA + gg : gg - R.nA);
        const float* so = S.out + (tid >> 4) * 17 + gg;
#pragma unroll
        for (int it = 0; it < 4; ++it) o[it * 16 * out_ms] = so[it * 16 * 17];
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
  const size_t smem = sizeof(SmemLayout) + sizeof(float) * kMaxMelWeights;
  static_assert(3 * (sizeof(SmemLayout) + sizeof(float) * kMaxMelWeights + 1024) <= 228 * 1024, "three CTAs per SM");
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
