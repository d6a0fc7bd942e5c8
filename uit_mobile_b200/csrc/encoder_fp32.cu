// fp32 CUDA-core encoder (UITK_PREC_FP32): the exact-arithmetic validation path of the UiT encoder.
//
// Unfused on purpose: every stage is a small kernel whose output can be compared with the oracle, and the
// tensor-core path (encoder_tc.cu) is checked against it on the device.  Stages (models/uit.py):
//   patch GEMM  = top-dB clamp + eval BatchNorm (460-462) + Conv2d 16x16/16 as [rows,256]x[256,128] + pos (380-388)
//   per block   = LN1+qkv GEMM | attention (89-122) | proj GEMM + residual | LN2+fc1 GEMM+ReLU | fc2 GEMM + residual
//   head        = final LN(1e-6), token mean, LN(1e-5), Linear 128->outputdim, sigmoid, crop mean/max (395-404, 468-488)
// It is also the implementation of the UITBase variants the tensor-core megakernel does not cover: full
// Attention (2 heads x 64, uit.py:124-178), GELU MLP (uit.py:338), pooling='token' (cls row, uit.py:389-392, 399-401) and
// pooling='dm' (uit.py:405-412), and of uitk_forward_features / uitk_forward_head (uit.py:379-412).
#include "uitk_common.cuh"

namespace uitk {

namespace {

enum { PRO_PLAIN = 0, PRO_LN = 1, PRO_PATCH = 2 };
enum { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_RESID = 2, EPI_PATCH = 3, EPI_SIGMOID = 4, EPI_BIAS_GELU = 5 };

struct GemmParams {
  const float* A; int lda;
  const float* Wt; int ldw;
  const float* bias;
  float* C; int ldc;
  int M, K;
  int n_valid;   // EPI_SIGMOID: columns >= n_valid are padding and are not stored
  // PRO_LN
  const float* ln_w; const float* ln_b; float ln_eps;
  // PRO_PATCH / EPI_PATCH
  const float* db; int T; int crops; int tokens; int t_n; int target;
  const float* bn_scale; const float* bn_shift; const uint32_t* max_pow;
  const float* time_pos; const float* freq_pos;
  int no_clamp;    // PRO_PATCH: the input is an already normalised spectrogram (uitk_forward_features)
  int out_tt;      // EPI_PATCH: rows per crop in C (tokens, +1 when a cls row leads every crop) ...
  int out_off;     // ... and the row offset of the first patch token inside a crop (1 with a cls row)
};

constexpr int BM = 128, KS = 32, AS_LD = 36;

template <int BN, int PRO, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(GemmParams p) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BM * AS_LD];
  __shared__ __align__(16) float Ws[KS * BN];
  __shared__ float s_mean[BM], s_rstd[BM];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  if (PRO == PRO_LN) {   // K == 128: one float4 per lane
    for (int r = warp * 16; r < warp * 16 + 16; ++r) {
      const int row = row0 + r;
      float mean = 0.f, rstd = 0.f;
      if (row < p.M) {
        const float4 x = *reinterpret_cast<const float4*>(p.A + (size_t)row * p.lda + lane * 4);
        float s = x.x + x.y + x.z + x.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean = s * (1.f / 128.f);
        const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd = rsqrtf(q * (1.f / 128.f) + p.ln_eps);
      }
      if (lane == 0) { s_mean[r] = mean; s_rstd[r] = rstd; }
    }
  }
  float cutoff = 0.f;
  if (PRO == PRO_PATCH) cutoff = p.no_clamp ? -INFINITY : 3.01029995663981195f * __log2f(fmaxf(__uint_as_float(*p.max_pow), 1e-10f)) - 120.f;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += KS) {
    __syncthreads();
    // ---- A slab [128][32]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      const int r = idx >> 3, c4 = idx & 7;
      const int row = row0 + r;
      const int k = k0 + c4 * 4;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.M) {
        if (PRO == PRO_PATCH) {
          const int rr = row / p.tokens, tok = row - rr * p.tokens;
          const int b = rr / p.crops, c = rr - b * p.crops;
          int start = 0;
          if (p.crops > 1) { start = c * p.target; if (start > p.T - p.target) start = p.T - p.target; }
          const int f = tok / p.t_n, tau = tok - f * p.t_n;
          const int mel = f * 16 + (k >> 4);
          const int t = start + tau * 16 + (k & 15);
          const float* src = p.db + ((size_t)b * 64 + mel) * p.T + t;
          const float sc = p.bn_scale[mel], sh = p.bn_shift[mel];
          val.x = fmaf(fmaxf(__ldg(src + 0), cutoff), sc, sh);
          val.y = fmaf(fmaxf(__ldg(src + 1), cutoff), sc, sh);
          val.z = fmaf(fmaxf(__ldg(src + 2), cutoff), sc, sh);
          val.w = fmaf(fmaxf(__ldg(src + 3), cutoff), sc, sh);
        } else {
          val = *reinterpret_cast<const float4*>(p.A + (size_t)row * p.lda + k);
          if (PRO == PRO_LN) {
            const float mean = s_mean[r], rstd = s_rstd[r];
            const float4 g = *reinterpret_cast<const float4*>(p.ln_w + k);
            const float4 be = *reinterpret_cast<const float4*>(p.ln_b + k);
            val.x = (val.x - mean) * rstd * g.x + be.x;
            val.y = (val.y - mean) * rstd * g.y + be.y;
            val.z = (val.z - mean) * rstd * g.z + be.z;
            val.w = (val.w - mean) * rstd * g.w + be.w;
          }
        }
      }
      *reinterpret_cast<float4*>(&As[r * AS_LD + c4 * 4]) = val;
    }
    // ---- W slab [32][BN]
    for (int idx = tid; idx < KS * BN / 4; idx += 256) {
      const int kk = idx / (BN / 4), c4 = idx - kk * (BN / 4);
      *reinterpret_cast<float4*>(&Ws[kk * BN + c4 * 4]) =
          *reinterpret_cast<const float4*>(p.Wt + (size_t)(k0 + kk) * p.ldw + n0 + c4 * 4);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KS; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&As[(ty * 8 + i) * AS_LD + kk]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float bv[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Ws[(kk + k) * BN + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = k == 0 ? a[i].x : (k == 1 ? a[i].y : (k == 2 ? a[i].z : a[i].w));
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + ty * 8 + i;
    if (row >= p.M) continue;
    int f = 0, tau = 0;
    size_t crow = row;
    if (EPI == EPI_PATCH) {
      const int rr = row / p.tokens, tok = row - rr * p.tokens;
      f = tok / p.t_n; tau = tok - f * p.t_n;
      crow = (size_t)rr * p.out_tt + p.out_off + tok;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx + 16 * j;
      float v = acc[i][j] + p.bias[col];
      if (EPI == EPI_SIGMOID) {
        if (col < p.n_valid) p.C[(size_t)row * p.ldc + col] = 1.f / (1.f + expf(-v));
        continue;
      }
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_BIAS_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));   // nn.GELU() (exact erf)
      if (EPI == EPI_PATCH) { v += p.time_pos[tau * 128 + col]; v += p.freq_pos[f * 128 + col]; }
      float* dst = p.C + crow * p.ldc + col;
      if (EPI == EPI_BIAS_RESID) v = *dst + v;
      *dst = v;
    }
  }
}

// One warp per (clip-crop, head); lane i < tokens owns query row i.  qkv rows are [q(2 x HD) | k(2 x HD) | v(2 x HD)];
// HD = 16 (BNeckAttention) or 64 (Attention); tokens <= 25.  CLIPS clip-crops per CTA (shared-memory budget).
template <int HD, int CLIPS>
__global__ void __launch_bounds__(CLIPS * 64) attention_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                               int RR, int tokens, float scale) {
  constexpr int kMaxTok = UITK_MAX_TOKENS + 1, ROW = 6 * HD;
  __shared__ float s[CLIPS][kMaxTok * ROW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr0 = blockIdx.x * CLIPS;
  for (int lr = 0; lr < CLIPS; ++lr) {
    const int rr = rr0 + lr;
    if (rr >= RR) break;
    const float* src = qkv + (size_t)rr * tokens * ROW;
    for (int i = tid; i < tokens * ROW; i += CLIPS * 64) s[lr][i] = src[i];
  }
  __syncthreads();
  const int lr = warp >> 1, head = warp & 1;
  const int rr = rr0 + lr;
  if (rr >= RR || lane >= tokens) return;
  const float* base = s[lr];
  const float* qi = base + lane * ROW + head * HD;
  float sc[kMaxTok];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxTok; ++j) {
    float a = 0.f;
    if (j < tokens) {
      const float* kj = base + j * ROW + 2 * HD + head * HD;
#pragma unroll
      for (int d = 0; d < HD; ++d) a = fmaf(qi[d], kj[d], a);
      a *= scale;
      mx = fmaxf(mx, a);
    }
    sc[j] = a;
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxTok; ++j) {
    sc[j] = j < tokens ? expf(sc[j] - mx) : 0.f;
    sum += sc[j];
  }
  const float inv = 1.f / sum;
  float* dst = o + ((size_t)rr * tokens + lane) * (2 * HD) + head * HD;
#pragma unroll 1
  for (int d0 = 0; d0 < HD; d0 += 16) {
    float out[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) out[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxTok; ++j) {
      if (j < tokens) {
        const float pj = sc[j] * inv;
        const float* vj = base + j * ROW + 4 * HD + head * HD + d0;
#pragma unroll
        for (int d = 0; d < 16; ++d) out[d] = fmaf(pj, vj[d], out[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 16; d += 4) *reinterpret_cast<float4*>(dst + d0 + d) = make_float4(out[d], out[d + 1], out[d + 2], out[d + 3]);
  }
}

// cls rows of pooling='token': x[rr * tt + 0][:] = cls_token + token_pos_embed (uit.py:389-392)
__global__ void cls_fill_kernel(float* __restrict__ x, int RR, int tt, const float* __restrict__ cls_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < RR * 128) x[(size_t)(i >> 7) * tt * 128 + (i & 127)] = cls_row[i & 127];
}

// Final LayerNorm (eps 1e-6, affine) of token rows, one warp per output row: out[rr][j] = LN(x[rr * slots + map(j)]),
// map(j) = (j / t_n) * slot_tn + j % t_n  (identity when slot_tn == t_n; the tensor-core tile keeps 6 time slots per mel band).
__global__ void __launch_bounds__(256) final_ln_kernel(const float* __restrict__ x, long long rows_out, int n_out, int slots, int t_n,
                                                       int slot_tn, const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows_out) return;
  const long long rr = row / n_out;
  const int j = (int)(row - rr * n_out);
  const int src = slot_tn == t_n ? j : (j / t_n) * slot_tn + j % t_n;
  const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)rr * slots + src) * 128 + lane * 4);
  float s = v.x + v.y + v.z + v.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 128.f);
  const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
  float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-6f);
  const float4 g = *reinterpret_cast<const float4*>(w + lane * 4), be = *reinterpret_cast<const float4*>(b + lane * 4);
  *reinterpret_cast<float4*>(out + (size_t)row * 128 + lane * 4) =
      make_float4(d0 * rstd * g.x + be.x, d1 * rstd * g.y + be.y, d2 * rstd * g.z + be.z, d3 * rstd * g.w + be.w);
}

// eval-mode init_bn on a [B, 64, T] log-mel (uit.py:310-313, 460-462): out = x * scale[mel] + shift[mel]
__global__ void init_bn_kernel(const float* __restrict__ db, long long n, int T, const float* __restrict__ scale,
                               const float* __restrict__ shift, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int mel = (int)((i / T) & 63);
    out[i] = fmaf(db[i], scale[mel], shift[mel]);
  }
}

// One CTA per clip: [final LN per token,] pooling, head LN, Linear + sigmoid, reduce over crops.
//   APPLY_NORM: x holds the residual stream and the final LayerNorm (eps 1e-6) is applied here; otherwise x already holds
//               forward_features' output (uitk_forward_head).
//   pooling (uit.py:398-412): MEAN  - one group of all tokens;  TOKEN - one group = row 0 (the cls token);
//               DM - t_n groups {f * t_n + tau | f < 4}: frequency mean per time step, head + sigmoid per step, mean of the scores.
template <bool APPLY_NORM>
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ x, int crops, int tokens, int pooling, int t_n,
                                                   const float* __restrict__ norm_w, const float* __restrict__ norm_b,
                                                   const float* __restrict__ hln_w, const float* __restrict__ hln_b,
                                                   const float* __restrict__ head_wt, const float* __restrict__ head_b,
                                                   int outputdim, int ld_head, int eval_max, float* __restrict__ probs) {
  __shared__ float part[8][128];
  __shared__ float pooled[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t b = blockIdx.x;
  float accp[3] = {0.f, 0.f, 0.f};
  if (eval_max) accp[0] = accp[1] = accp[2] = -INFINITY;
  const int n_groups = pooling == UITK_POOL_DM ? t_n : 1;
  const float4 g = *reinterpret_cast<const float4*>(norm_w + lane * 4);
  const float4 be = *reinterpret_cast<const float4*>(norm_b + lane * 4);
  for (int c = 0; c < crops; ++c) {
    const float* xc = x + (b * crops + c) * (size_t)tokens * 128;
    float crop_p[3] = {0.f, 0.f, 0.f};
    for (int grp = 0; grp < n_groups; ++grp) {
      // rows of this group: first, step, count
      int first = 0, step = 1, count = tokens;
      if (pooling == UITK_POOL_TOKEN) count = 1;
      if (pooling == UITK_POOL_DM) { first = grp; step = t_n; count = 4; }
      float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = warp; i < count; i += 8) {
        const float4 v = *reinterpret_cast<const float4*>(xc + (size_t)(first + i * step) * 128 + lane * 4);
        if (APPLY_NORM) {
          float s = v.x + v.y + v.z + v.w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * (1.f / 128.f);
          const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
          float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-6f);
          ps.x += d0 * rstd * g.x + be.x; ps.y += d1 * rstd * g.y + be.y;
          ps.z += d2 * rstd * g.z + be.z; ps.w += d3 * rstd * g.w + be.w;
        } else {
          ps.x += v.x; ps.y += v.y; ps.z += v.z; ps.w += v.w;
        }
      }
      __syncthreads();   // previous group's readers of pooled/part are done
      *reinterpret_cast<float4*>(&part[warp][lane * 4]) = ps;
      __syncthreads();
      if (warp == 0) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 t4 = *reinterpret_cast<const float4*>(&part[w][lane * 4]);
          m.x += t4.x; m.y += t4.y; m.z += t4.z; m.w += t4.w;
        }
        const float invn = 1.f / (float)count;
        m.x *= invn; m.y *= invn; m.z *= invn; m.w *= invn;
        float s = m.x + m.y + m.z + m.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / 128.f);
        const float d0 = m.x - mean, d1 = m.y - mean, d2 = m.z - mean, d3 = m.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-5f);
        const float4 hg = *reinterpret_cast<const float4*>(hln_w + lane * 4);
        const float4 hb = *reinterpret_cast<const float4*>(hln_b + lane * 4);
        *reinterpret_cast<float4*>(&pooled[lane * 4]) =
            make_float4(d0 * rstd * hg.x + hb.x, d1 * rstd * hg.y + hb.y, d2 * rstd * hg.z + hb.z, d3 * rstd * hg.w + hb.w);
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int oidx = tid + 256 * r;
        if (oidx < outputdim) {
          float z = head_b[oidx];
#pragma unroll 8
          for (int k = 0; k < 128; ++k) z = fmaf(pooled[k], __ldg(head_wt + (size_t)k * ld_head + oidx), z);
          crop_p[r] += 1.f / (1.f + expf(-z));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float pr = n_groups == 1 ? crop_p[r] : crop_p[r] / (float)n_groups;
      accp[r] = eval_max ? fmaxf(accp[r], pr) : accp[r] + pr;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int oidx = tid + 256 * r;
    if (oidx < outputdim) probs[b * outputdim + oidx] = eval_max ? accp[r] : (crops == 1 ? accp[r] : accp[r] / (float)crops);
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

size_t encoder_fp32_workspace_bytes(const uitk_encoder_cfg& cfg, int64_t rows) {
  const size_t r = (size_t)rows;
  const size_t qkv_n = cfg.attention == UITK_ATTN_FULL ? 384 : 96, inner = cfg.attention == UITK_ATTN_FULL ? 128 : 32;
  return align_up(r * 128 * 4, 256) + align_up(r * qkv_n * 4, 256) + align_up(r * inner * 4, 256) + align_up(r * 384 * 4, 256);
}

int launch_final_ln(const float* x, int64_t RR, int slots, int t_n, int slot_tn, const float* norm_w, const float* norm_b, float* out,
                    cudaStream_t s) {
  // n_out rows per crop: the compact token count (cls row included when slots is the compact count + 1)
  const int n_out = slot_tn == t_n ? slots : 4 * t_n;
  const long long rows = (long long)RR * n_out;
  if (rows == 0) return UITK_OK;
  final_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, rows, n_out, slots, t_n, slot_tn, norm_w, norm_b, out);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

int launch_init_bn(const float* db, int64_t B, int64_t T, const float* scale, const float* shift, float* out, cudaStream_t s) {
  const long long n = (long long)B * 64 * T;
  if (n == 0) return UITK_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  init_bn_kernel<<<(int)blocks, 256, 0, s>>>(db, n, (int)T, scale, shift, out);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

int run_forward_head(const uitk_encoder_cfg& cfg, const void* blob, const float* tokens, int64_t B, int n_tokens, float* probs,
                     cudaStream_t s) {
  const EncoderLayout lay = make_encoder_layout(cfg);
  const float* W = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(blob) + sizeof(BlobHeader));
  const int t_n = cfg.pooling == UITK_POOL_DM ? n_tokens / 4 : 0;
  head_kernel<false><<<(unsigned)B, 256, 0, s>>>(tokens, 1, n_tokens, cfg.pooling, t_n, W + lay.norm_w, W + lay.norm_b, W + lay.hln_w,
                                                 W + lay.hln_b, W + lay.head_wt, W + lay.head_b, cfg.outputdim, lay.outputdim_padded, 0, probs);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

int run_encoder_fp32(const EncoderArgs& a) {
  const uitk_encoder_cfg& cfg = *a.cfg;
  const bool feat = a.features_out != nullptr;
  const int crops = feat ? 1 : crops_for(a.T, a.target_length);
  const int t_n = time_patches_for(a.T, a.target_length);
  const int tokens = 4 * t_n;                                          // patch tokens per crop
  const int tt = tokens + (cfg.pooling == UITK_POOL_TOKEN ? 1 : 0);    // rows per crop (cls row first)
  const int64_t RR = a.B * crops;
  const int64_t M64 = RR * tt;
  UITK_REQUIRE(M64 < (1ll << 31) - 256, UITK_EINVAL, "too many token rows for one call (%lld); chunk the batch", (long long)M64);
  UITK_REQUIRE(cfg.outputdim <= 768, UITK_EINVAL, "outputdim %d > 768 unsupported by the head kernel", cfg.outputdim);
  UITK_REQUIRE(t_n <= cfg.grid_t, UITK_EINVAL, "%d time patches exceed time_pos_embed length %d", t_n, cfg.grid_t);
  UITK_REQUIRE(a.cond_used == nullptr, UITK_EINVAL, "uitk_encoder_fixup is implemented by the tensor-core configuration only");
  const int M = (int)M64, Mp = (int)(RR * tokens);

  const EncoderLayout lay = make_encoder_layout(cfg);
  // the fp32 section starts right after the header (pack.cu)
  const float* W = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(a.blob) + sizeof(BlobHeader));
  const bool full = cfg.attention == UITK_ATTN_FULL;
  const int qkv_n = lay.qkv_n, inner = lay.inner;

  unsigned char* ws = reinterpret_cast<unsigned char*>(a.ws);
  float* x = reinterpret_cast<float*>(ws); ws += align_up((size_t)M * 128 * 4, 256);
  float* qkv = reinterpret_cast<float*>(ws); ws += align_up((size_t)M * qkv_n * 4, 256);
  float* o = reinterpret_cast<float*>(ws); ws += align_up((size_t)M * inner * 4, 256);
  float* h = reinterpret_cast<float*>(ws);
  UITK_REQUIRE(encoder_fp32_workspace_bytes(cfg, M) <= a.ws_bytes, UITK_ENOSPACE, "workspace too small: need %zu, have %zu",
               encoder_fp32_workspace_bytes(cfg, M), a.ws_bytes);

  cudaStream_t s = a.stream;
  const int mb = (M + BM - 1) / BM;
  GemmParams p{};
  p.M = Mp;
  // ---- patch embed
  p.K = 256; p.Wt = W + lay.patch_wt; p.ldw = 128; p.bias = W + lay.patch_b; p.C = x; p.ldc = 128;
  p.db = a.db; p.T = (int)a.T; p.crops = crops; p.tokens = tokens; p.t_n = t_n; p.target = a.target_length;
  p.bn_scale = W + (feat ? lay.ident_scale : lay.bn_scale); p.bn_shift = W + (feat ? lay.ident_shift : lay.bn_shift);
  p.max_pow = a.max_pow; p.no_clamp = feat ? 1 : 0;
  p.time_pos = W + lay.time_pos; p.freq_pos = W + lay.freq_pos;
  p.out_tt = tt; p.out_off = tt - tokens;
  gemm_kernel<128, PRO_PATCH, EPI_PATCH><<<dim3((Mp + BM - 1) / BM, 1), 256, 0, s>>>(p);
  int launches = 1;
  if (tt != tokens) {
    cls_fill_kernel<<<(unsigned)((RR * 128 + 255) / 256), 256, 0, s>>>(x, (int)RR, tt, W + lay.cls_row);
    ++launches;
  }

  for (int i = 0; i < cfg.depth; ++i) {
    const float* Wb = W + lay.blocks + (size_t)i * lay.block_stride;
    GemmParams g{};
    g.M = M;
    // LN1 + qkv
    g.A = x; g.lda = 128; g.K = 128; g.Wt = Wb + lay.blk.qkv_wt; g.ldw = qkv_n; g.bias = Wb + lay.blk.qkv_b;
    g.C = qkv; g.ldc = qkv_n; g.ln_w = Wb + lay.blk.ln1_w; g.ln_b = Wb + lay.blk.ln1_b; g.ln_eps = 1e-6f;
    if (full) {
      gemm_kernel<128, PRO_LN, EPI_BIAS><<<dim3(mb, 3), 256, 0, s>>>(g);
      attention_kernel<64, 1><<<(unsigned)RR, 64, 0, s>>>(qkv, o, (int)RR, tt, 0.125f);           // (128 // 2) ** -0.5
    } else {
      gemm_kernel<96, PRO_LN, EPI_BIAS><<<dim3(mb, 1), 256, 0, s>>>(g);
      attention_kernel<16, 4><<<(unsigned)((RR + 3) / 4), 256, 0, s>>>(qkv, o, (int)RR, tt, 0.125f);  // scale from the un-bottlenecked head dim (Q3)
    }
    // proj + residual
    g.A = o; g.lda = inner; g.K = inner; g.Wt = Wb + lay.blk.proj_wt; g.ldw = 128; g.bias = Wb + lay.blk.proj_b;
    g.C = x; g.ldc = 128;
    gemm_kernel<128, PRO_PLAIN, EPI_BIAS_RESID><<<dim3(mb, 1), 256, 0, s>>>(g);
    // LN2 + fc1 + activation
    g.A = x; g.lda = 128; g.K = 128; g.Wt = Wb + lay.blk.fc1_wt; g.ldw = 384; g.bias = Wb + lay.blk.fc1_b;
    g.C = h; g.ldc = 384; g.ln_w = Wb + lay.blk.ln2_w; g.ln_b = Wb + lay.blk.ln2_b;
    if (cfg.act == UITK_ACT_GELU) gemm_kernel<128, PRO_LN, EPI_BIAS_GELU><<<dim3(mb, 3), 256, 0, s>>>(g);
    else gemm_kernel<128, PRO_LN, EPI_BIAS_RELU><<<dim3(mb, 3), 256, 0, s>>>(g);
    // fc2 + residual
    g.A = h; g.lda = 384; g.K = 384; g.Wt = Wb + lay.blk.fc2_wt; g.ldw = 128; g.bias = Wb + lay.blk.fc2_b;
    g.C = x; g.ldc = 128;
    gemm_kernel<128, PRO_PLAIN, EPI_BIAS_RESID><<<dim3(mb, 1), 256, 0, s>>>(g);
  }
  count_launches(launches + 5 * cfg.depth);
  UITK_CHECK_CUDA(cudaGetLastError());
  if (feat) return launch_final_ln(x, RR, tt, t_n, t_n, W + lay.norm_w, W + lay.norm_b, a.features_out, s);
  head_kernel<true><<<(unsigned)a.B, 256, 0, s>>>(x, crops, tt, cfg.pooling, t_n, W + lay.norm_w, W + lay.norm_b, W + lay.hln_w,
                                                  W + lay.hln_b, W + lay.head_wt, W + lay.head_b, cfg.outputdim, lay.outputdim_padded,
                                                  a.eval_avg, a.probs);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

// Head for the tensor-core path: pooled features [B*crops][128] -> head LayerNorm(1e-5) -> Linear -> sigmoid -> crop
// mean/max.  8 clips per CTA so that every head-weight element fetched from L2 feeds 8 FMAs.
namespace {
constexpr int kHeadClips = 8;
__global__ void __launch_bounds__(256) head_pooled_kernel(const float* __restrict__ pooled, long long B, int crops,
                                                          const float* __restrict__ hln_w, const float* __restrict__ hln_b,
                                                          const float* __restrict__ head_wt, const float* __restrict__ head_b,
                                                          int outputdim, int ld_head, int eval_max, float* __restrict__ probs,
                                                          const uint32_t* c_true, const uint32_t* c_used, const uint32_t* c_min) {
  if (c_used != nullptr && !fixup_needed(c_true, c_used, c_min)) return;      // uitk_encoder_fixup: nothing could differ
  __shared__ __align__(16) float pn[kHeadClips][128];      // normalised features
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b0 = (long long)blockIdx.x * kHeadClips;
  float accp[3][kHeadClips];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < kHeadClips; ++c) accp[r][c] = eval_max ? -INFINITY : 0.f;
  const float4 hg = *reinterpret_cast<const float4*>(hln_w + lane * 4);
  const float4 hb = *reinterpret_cast<const float4*>(hln_b + lane * 4);
  for (int c = 0; c < crops; ++c) {
    __syncthreads();
    {   // warp w normalises clip b0 + w
      const long long b = b0 + warp;
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B) m = *reinterpret_cast<const float4*>(pooled + (b * crops + c) * 128 + lane * 4);
      float s = m.x + m.y + m.z + m.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.f / 128.f);
      const float d0 = m.x - mean, d1 = m.y - mean, d2 = m.z - mean, d3 = m.w - mean;
      float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-5f);
      *reinterpret_cast<float4*>(&pn[warp][lane * 4]) =
          make_float4(d0 * rstd * hg.x + hb.x, d1 * rstd * hg.y + hb.y, d2 * rstd * hg.z + hb.z, d3 * rstd * hg.w + hb.w);
    }
    __syncthreads();
    float z[3][kHeadClips];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int oidx = tid + 256 * r;
      const float bias = oidx < outputdim ? head_b[oidx] : 0.f;
#pragma unroll
      for (int cc = 0; cc < kHeadClips; ++cc) z[r][cc] = bias;
    }
    for (int k = 0; k < 128; k += 4) {
      float4 pv[kHeadClips];
#pragma unroll
      for (int cc = 0; cc < kHeadClips; ++cc) pv[cc] = *reinterpret_cast<const float4*>(&pn[cc][k]);   // broadcast loads
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int oidx = tid + 256 * r;
        if (oidx < outputdim) {
          const float w0 = __ldg(head_wt + (size_t)(k + 0) * ld_head + oidx), w1 = __ldg(head_wt + (size_t)(k + 1) * ld_head + oidx);
          const float w2 = __ldg(head_wt + (size_t)(k + 2) * ld_head + oidx), w3 = __ldg(head_wt + (size_t)(k + 3) * ld_head + oidx);
#pragma unroll
          for (int cc = 0; cc < kHeadClips; ++cc)
            z[r][cc] = fmaf(pv[cc].w, w3, fmaf(pv[cc].z, w2, fmaf(pv[cc].y, w1, fmaf(pv[cc].x, w0, z[r][cc]))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < kHeadClips; ++cc) {
        const float pr = 1.f / (1.f + expf(-z[r][cc]));
        accp[r][cc] = eval_max ? fmaxf(accp[r][cc], pr) : accp[r][cc] + pr;
      }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int oidx = tid + 256 * r;
    if (oidx >= outputdim) continue;
#pragma unroll
    for (int cc = 0; cc < kHeadClips; ++cc)
      if (b0 + cc < B) probs[(b0 + cc) * outputdim + oidx] = eval_max ? accp[r][cc] : accp[r][cc] / (float)crops;
  }
}
}  // namespace

// Single-crop head: probs[B][outputdim] = sigmoid(LayerNorm_1e-5(pooled[B][128]) W^T + b) as one GEMM on the tensor cores
// (mma.sync m16n8k8 tf32; the fp32 CUDA-core version of round 1 took 35 us of the encoder's 435), both operands split
// hi + lo and three products per k-step (hi*hi + lo*hi + hi*lo: ~2^-21 relative, i.e. fp32-grade logits in front of the sigmoid).
// 64 clips x 64 classes per CTA of 4 warps; warp w normalises and multiplies its own 16 rows, so the CTA needs no barrier.  The
// weight fragments come pre-split and in fragment order from the blob (pack.cu), one 16-byte load per lane, k-step and n-tile,
// loaded one k-step ahead.  35 us -> see profiles/README.md.
namespace {
constexpr int kTLd = 136;      // row stride of the normalised features: 8 (mod 32) floats, so the 8-byte fragment loads are conflict-free
__device__ __forceinline__ void mma_tf32_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128) head_tc_kernel(const float* __restrict__ pooled, int B, const float* __restrict__ hln_w,
                                                      const float* __restrict__ hln_b, const float4* __restrict__ head_frag,
                                                      const float* __restrict__ head_b, int outputdim, float* __restrict__ probs,
                                                      const uint32_t* c_true, const uint32_t* c_used, const uint32_t* c_min) {
  if (c_used != nullptr && !fixup_needed(c_true, c_used, c_min)) return;      // uitk_encoder_fixup: nothing could differ
  __shared__ __align__(16) float As[64 * kTLd];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int row0 = blockIdx.x * 64 + warp * 16, ct = blockIdx.y;
  const float4* fr = head_frag + (size_t)ct * 16 * 8 * 32 + lane;
  float4 bcur[8], bnext[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bcur[j] = __ldg(fr + j * 32);
  const float4 hg = __ldg(reinterpret_cast<const float4*>(hln_w + lane * 4));
  const float4 hb = __ldg(reinterpret_cast<const float4*>(hln_b + lane * 4));
  float* Aw = As + warp * 16 * kTLd;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {                    // LayerNorm(1e-5) of this warp's 16 rows (one float4 per lane)
    const int row = row0 + i;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < B) m = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)row * 128 + lane * 4));
    float sum = m.x + m.y + m.z + m.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / 128.f);
    const float d0 = m.x - mean, d1 = m.y - mean, d2 = m.z - mean, d3 = m.w - mean;
    float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / 128.f) + 1e-5f);
    *reinterpret_cast<float4*>(&Aw[i * kTLd + lane * 4]) =
        make_float4(d0 * rstd * hg.x + hb.x, d1 * rstd * hg.y + hb.y, d2 * rstd * hg.z + hb.z, d3 * rstd * hg.w + hb.w);
  }
  __syncwarp();
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const float* a_lo_row = Aw + gid * kTLd + 2 * tig;       // logical k = tig / tig + 4  <->  features 8 s + 2 tig, + 1
  const float* a_hi_row = a_lo_row + 8 * kTLd;
#pragma unroll 1
  for (int s = 0; s < 16; ++s) {
    if (s + 1 < 16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bnext[j] = __ldg(fr + ((s + 1) * 8 + j) * 32);
    }
    const float2 p0 = *reinterpret_cast<const float2*>(a_lo_row + 8 * s), p1 = *reinterpret_cast<const float2*>(a_hi_row + 8 * s);
    const uint32_t a0 = __float_as_uint(p0.x) & 0xffffe000u, a1 = __float_as_uint(p1.x) & 0xffffe000u;
    const uint32_t a2 = __float_as_uint(p0.y) & 0xffffe000u, a3 = __float_as_uint(p1.y) & 0xffffe000u;
    const uint32_t l0 = __float_as_uint(p0.x - __uint_as_float(a0)), l1 = __float_as_uint(p1.x - __uint_as_float(a1));
    const uint32_t l2 = __float_as_uint(p0.y - __uint_as_float(a2)), l3 = __float_as_uint(p1.y - __uint_as_float(a3));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mma_tf32_acc(acc[j], a0, a1, a2, a3, __float_as_uint(bcur[j].x), __float_as_uint(bcur[j].y));
      mma_tf32_acc(acc[j], l0, l1, l2, l3, __float_as_uint(bcur[j].x), __float_as_uint(bcur[j].y));
      mma_tf32_acc(acc[j], a0, a1, a2, a3, __float_as_uint(bcur[j].z), __float_as_uint(bcur[j].w));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bcur[j] = bnext[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = ct * 64 + j * 8 + 2 * tig;
    const float b0 = __ldg(head_b + col), b1 = __ldg(head_b + col + 1);          // head_b is zero padded to the padded width
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + gid + 8 * h;
      if (row >= B) continue;
      if (col < outputdim) probs[(size_t)row * outputdim + col] = 1.f / (1.f + expf(-(acc[j][2 * h] + b0)));
      if (col + 1 < outputdim) probs[(size_t)row * outputdim + col + 1] = 1.f / (1.f + expf(-(acc[j][2 * h + 1] + b1)));
    }
  }
}
}  // namespace

int launch_head_pooled(const float* pooled, int64_t B, int crops, const float* W, const EncoderLayout& lay, int outputdim,
                       int eval_max, float* probs, cudaStream_t s, const uint32_t* c_true, const uint32_t* c_used,
                       const uint32_t* c_min) {
  if (crops == 1 && B < (1ll << 31) - 256) {
    // single crop: one tiled GEMM [B,128] x [128,outputdim] with the head LayerNorm as prologue and the sigmoid as epilogue
    head_tc_kernel<<<dim3((unsigned)((B + 63) / 64), (unsigned)((outputdim + 63) / 64)), 128, 0, s>>>(
        pooled, (int)B, W + lay.hln_w, W + lay.hln_b, reinterpret_cast<const float4*>(W + lay.head_frag), W + lay.head_b, outputdim, probs,
        c_true, c_used, c_min);
    count_launches(1);
    UITK_CHECK_CUDA(cudaGetLastError());
    return UITK_OK;
  }
  head_pooled_kernel<<<(unsigned)((B + kHeadClips - 1) / kHeadClips), 256, 0, s>>>(
      pooled, (long long)B, crops, W + lay.hln_w, W + lay.hln_b, W + lay.head_wt, W + lay.head_b, outputdim, lay.outputdim_padded,
      eval_max, probs, c_true, c_used, c_min);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

}  // namespace uitk
