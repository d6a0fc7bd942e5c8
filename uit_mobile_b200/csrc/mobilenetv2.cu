// MobileNetV2 teacher / audio-tagging baseline: models/mobilenetv2.py:66-178, eval forward
//   front_end (log-mel + top-dB clamp: uitk_logmel / uitk_clamp_db)  ->  [B, 1, 64, T]
//   features: conv3x3/2 (1 -> 32) + BN + ReLU6, 17 inverted-residual blocks (mobilenetv2.py:30-64), conv1x1 (320 -> 1280) + BN + ReLU6,
//             AdaptiveAvgPool2d((1, None)) = mean over the mel axis
//   classifier: Linear(1280 -> outputdim) per time step, sigmoid, mean over time.
// fp32 CUDA-core kernels, activations NHWC ([clip][mel][time][channel]) so that the 1x1 convolutions are plain GEMMs over the
// pixels.  Eval BatchNorm is folded at pack time: scale into the convolution weights, shift kept as the bias.  The model is the
// reference's distillation teacher, not its deployed path: correctness (fp32 parity) first, no tensor-core variant.
#include <math.h>

#include <string>
#include <vector>

#include "uitk_common.cuh"

namespace uitk {

namespace {

enum { L_STEM = 0, L_PW, L_DW, L_POOL, L_CLS };
struct Layer {
  int type, cin, cout, stride, relu6;
  int src, dst;        // activation buffers: 0 = A (block input / output), 1 = B (expanded), 2 = C (depthwise output)
  int residual;        // pointwise: add the block input (dst == A, in place)
  size_t w, b;         // float offsets into the blob (weights [K][N] / [9][C] / [9][32], folded shift [N])
};

// The default inverted_residual_setting (mobilenetv2.py:106-116), width_mult 1.0
const int kSetting[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2}, {6, 32, 3, 2}, {6, 64, 4, 2}, {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};
constexpr int kLast = 1280;
constexpr int kMnv2Magic = 0x554d5631;   // 'UMV1'

struct Net {
  std::vector<Layer> layers;
  size_t total_floats;
};

Net make_net(int outputdim) {
  Net n;
  size_t cur = 64;                                              // 256-byte header
  auto take = [&](size_t k) { size_t o = cur; cur += (k + 63) / 64 * 64; return o; };
  auto add = [&](int type, int cin, int cout, int stride, int relu6, int src, int dst, int res) {
    Layer l{type, cin, cout, stride, relu6, src, dst, res, 0, 0};
    const size_t wn = type == L_STEM ? (size_t)9 * cout : type == L_DW ? (size_t)9 * cout : (size_t)cin * cout;
    if (type != L_POOL) { l.w = take(wn); l.b = take(cout); }
    n.layers.push_back(l);
  };
  add(L_STEM, 1, 32, 2, 1, -1, 0, 0);
  int inp = 32;
  for (const auto& s : kSetting)
    for (int i = 0; i < s[2]; ++i) {
      const int stride = i == 0 ? s[3] : 1, hidden = inp * s[0], oup = s[1];
      int src = 0;
      if (s[0] != 1) { add(L_PW, inp, hidden, 1, 1, 0, 1, 0); src = 1; }
      add(L_DW, hidden, hidden, stride, 1, src, 2, 0);
      add(L_PW, hidden, oup, 1, 0, 2, 0, stride == 1 && inp == oup);
      inp = oup;
    }
  add(L_PW, inp, kLast, 1, 1, 0, 1, 0);
  add(L_POOL, kLast, kLast, 1, 0, 1, 2, 0);
  add(L_CLS, kLast, outputdim, 1, 0, 2, 0, 0);
  n.total_floats = cur;
  return n;
}

inline int down(int v, int stride) { return stride == 1 ? v : (v - 1) / 2 + 1; }     // 3x3, padding 1

// ---- kernels ---------------------------------------------------------------------------------------------------------------
// stem: in [B][64][T] (one channel) -> out [B][Ho][Wo][32], 3x3 stride 2 padding 1, + shift, ReLU6.  One thread = one pixel x 4 channels.
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ in, long long B, int H, int W, int Ho, int Wo,
                                                   const float* __restrict__ w, const float* __restrict__ shift, float* __restrict__ out) {
  const long long total = B * Ho * Wo * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i & 7);
    long long p = i >> 3;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const long long b = p / Ho;
    float4 acc = __ldg(reinterpret_cast<const float4*>(shift) + c4);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int x = 2 * wo - 1 + kw;
        if (x < 0 || x >= W) continue;
        const float v = __ldg(in + (b * H + h) * W + x);
        const float4 f = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * 32) + c4);
        acc.x = fmaf(v, f.x, acc.x); acc.y = fmaf(v, f.y, acc.y); acc.z = fmaf(v, f.z, acc.z); acc.w = fmaf(v, f.w, acc.w);
      }
    }
    acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
    acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// depthwise 3x3, padding 1, stride 1 | 2, + shift, ReLU6.  NHWC, weights [9][C]; one thread = one output pixel x 4 channels.
__global__ void __launch_bounds__(256) dw_kernel(const float* __restrict__ in, long long B, int H, int W, int C, int stride, int Ho, int Wo,
                                                 const float* __restrict__ w, const float* __restrict__ shift, float* __restrict__ out) {
  const int C4 = C >> 2;
  const long long total = B * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long p = i / C4;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const long long b = p / Ho;
    float4 acc = __ldg(reinterpret_cast<const float4*>(shift) + c4);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = stride * ho - 1 + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int x = stride * wo - 1 + kw;
        if (x < 0 || x >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + ((b * H + h) * W + x) * C) + c4);
        const float4 f = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C) + c4);
        acc.x = fmaf(v.x, f.x, acc.x); acc.y = fmaf(v.y, f.y, acc.y); acc.z = fmaf(v.z, f.z, acc.z); acc.w = fmaf(v.w, f.w, acc.w);
      }
    }
    acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
    acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// pointwise convolution / Linear: out[M][N] = epi(in[M][K] * w[K][N] + shift[N]) (+ out[M][N] when RES: the block input, in place).
// 64 x 64 tile per CTA, K in chunks of 8 (every channel count of the network is a multiple of 8), 4 x 4 outputs per thread.
enum { EPI_NONE = 0, EPI_RELU6 = 1, EPI_SIGMOID = 2 };
template <int EPI, bool RES>
__global__ void __launch_bounds__(256) pw_kernel(const float* __restrict__ in, long long M, int K, int N, const float* __restrict__ w,
                                                 const float* __restrict__ shift, float* __restrict__ out) {
  __shared__ __align__(16) float As[8][64 + 4];     // [k][m]
  __shared__ __align__(16) float Ws[8][64];         // [k][n]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lm = tid >> 2, lk = (tid & 3) * 2;      // A loader: row lm, k pair lk
  const int wk = tid >> 5, wn = (tid & 31) * 2;     // W loader: row wk, column pair wn
  for (int k0 = 0; k0 < K; k0 += 8) {
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
    if (m0 + lm < M) a = __ldg(reinterpret_cast<const float2*>(in + (m0 + lm) * K + k0 + lk));
    if (n0 + wn < N) b.x = __ldg(w + (size_t)(k0 + wk) * N + n0 + wn);           // scalar: N may be odd (outputdim 537)
    if (n0 + wn + 1 < N) b.y = __ldg(w + (size_t)(k0 + wk) * N + n0 + wn + 1);
    __syncthreads();                                // previous chunk consumed
    As[lk][lm] = a.x; As[lk + 1][lm] = a.y;
    *reinterpret_cast<float2*>(&Ws[wk][wn]) = b;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + __ldg(shift + n);
      if (EPI == EPI_RELU6) v = fminf(fmaxf(v, 0.f), 6.f);
      if (EPI == EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (RES) v += out[m * N + n];
      out[m * N + n] = v;
    }
  }
}

// AdaptiveAvgPool2d((1, None)): in [B][H][W][C] -> out [B][W][C] (mean over the mel axis)
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ in, long long B, int H, int W, int C, float* __restrict__ out) {
  const long long total = B * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int x = (int)(p % W);
    const long long b = p / W;
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += in[((b * H + h) * W + x) * C + c];
    out[i] = s / (float)H;
  }
}

// x.mean(1) over the time steps: in [B][W][N] -> out [B][N]
__global__ void __launch_bounds__(256) time_mean_kernel(const float* __restrict__ in, long long B, int W, int N, float* __restrict__ out) {
  const long long total = B * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long b = i / N;
    float s = 0.f;
    for (int x = 0; x < W; ++x) s += in[(b * W + x) * N + n];
    out[i] = s / (float)W;
  }
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 64 ? 148 * 64 : g));
}

constexpr int kChunk = 256;            // clips per pass: bounds the workspace (3 activation buffers of the chunk)

size_t per_clip_floats(int64_t T) {    // largest activation of one clip, over all layers
  int H = 64, W = (int)T;
  size_t mx = 0;
  const Net net = make_net(1);
  for (const Layer& l : net.layers) {
    if (l.type == L_STEM || l.type == L_DW) { H = down(H, l.stride); W = down(W, l.stride); }
    size_t n = (size_t)H * W * l.cout;
    if (l.type == L_POOL) n = (size_t)W * l.cout;
    if (l.type != L_CLS && n > mx) mx = n;
  }
  return mx;
}

}  // namespace

}  // namespace uitk

using namespace uitk;

extern "C" {

int uitk_mnv2_num_tensors(void) {
  const Net net = make_net(1);
  int n = 2;
  for (const Layer& l : net.layers)
    if (l.type == L_STEM || l.type == L_PW || l.type == L_DW) n += 5;
  return n;
}

// state_dict key of h_tensors[index] (mobilenetv2.py module tree): conv weight, then BatchNorm weight / bias / running_mean /
// running_var per convolution, in network order; classifier.1.weight / bias last.
const char* uitk_mnv2_tensor_name(int index) {
  static thread_local std::string name;
  std::vector<std::string> names;
  auto unit = [&](const std::string& conv, const std::string& bn) {
    names.push_back(conv + ".weight");
    for (const char* s : {".weight", ".bias", ".running_mean", ".running_var"}) names.push_back(bn + s);
  };
  unit("features.0.0", "features.0.1");
  int f = 1;
  for (const auto& s : kSetting)
    for (int i = 0; i < s[2]; ++i, ++f) {
      const std::string p = "features." + std::to_string(f) + ".conv.";
      if (s[0] != 1) {
        unit(p + "0.0", p + "0.1"); unit(p + "1.0", p + "1.1"); unit(p + "2", p + "3");
      } else {
        unit(p + "0.0", p + "0.1"); unit(p + "1", p + "2");
      }
    }
  unit("features." + std::to_string(f) + ".0", "features." + std::to_string(f) + ".1");
  names.push_back("classifier.1.weight");
  names.push_back("classifier.1.bias");
  if (index < 0 || index >= (int)names.size()) return nullptr;
  name = names[index];
  return name.c_str();
}

size_t uitk_mnv2_blob_bytes(int outputdim) { return outputdim >= 1 ? make_net(outputdim).total_floats * sizeof(float) : 0; }

int uitk_pack_mnv2(int outputdim, const float* const* t, void* h_blob, size_t blob_bytes) {
  UITK_REQUIRE(t && h_blob, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(outputdim >= 1, UITK_EINVAL, "outputdim must be positive");
  const Net net = make_net(outputdim);
  UITK_REQUIRE(blob_bytes >= net.total_floats * sizeof(float), UITK_ENOSPACE, "MobileNetV2 blob needs %zu bytes", net.total_floats * sizeof(float));
  float* W = reinterpret_cast<float*>(h_blob);
  memset(W, 0, net.total_floats * sizeof(float));
  reinterpret_cast<int*>(W)[0] = kMnv2Magic;
  reinterpret_cast<int*>(W)[1] = outputdim;
  int ti = 0;
  for (const Layer& l : net.layers) {
    if (l.type == L_POOL) continue;
    if (l.type == L_CLS) {                                     // Linear weight [N][K] -> [K][N], bias as is
      for (int n = 0; n < l.cout; ++n)
        for (int k = 0; k < l.cin; ++k) W[l.w + (size_t)k * l.cout + n] = t[ti][(size_t)n * l.cin + k];
      memcpy(W + l.b, t[ti + 1], sizeof(float) * l.cout);
      ti += 2;
      continue;
    }
    const float *cw = t[ti], *g = t[ti + 1], *be = t[ti + 2], *mu = t[ti + 3], *var = t[ti + 4];
    ti += 5;
    for (int c = 0; c < l.cout; ++c) {                         // eval BatchNorm (eps 1e-5) folded: y = conv * scale + shift
      const float scale = g[c] / sqrtf(var[c] + 1e-5f);
      W[l.b + c] = be[c] - mu[c] * scale;
      if (l.type == L_PW) {
        for (int k = 0; k < l.cin; ++k) W[l.w + (size_t)k * l.cout + c] = cw[(size_t)c * l.cin + k] * scale;      // [N][K][1][1] -> [K][N]
      } else {
        for (int q = 0; q < 9; ++q) W[l.w + (size_t)q * l.cout + c] = cw[(size_t)c * 9 + q] * scale;              // [C][1][3][3] -> [9][C]
      }
    }
  }
  return UITK_OK;
}

size_t uitk_mnv2_workspace_bytes(int64_t B, int64_t T) {
  if (B < 0 || T < 1) return 0;
  const int64_t nb = B < kChunk ? B : kChunk;
  return 3 * (per_clip_floats(T) * (size_t)nb + 64) * sizeof(float) + 256;
}

// d_db: [B, 64, T] log-mel dB AFTER the top-dB clamp (front_end of mobilenetv2.py:146-154)  ->  d_probs [B, outputdim]
int uitk_mnv2_forward(int outputdim, const void* d_blob, const float* d_db, int64_t B, int64_t T, float* d_probs, void* d_workspace,
                      size_t workspace_bytes, void* stream) {
  UITK_REQUIRE(d_blob && d_db && d_probs && d_workspace, UITK_EINVAL, "null pointer");
  UITK_REQUIRE(B >= 0 && T >= 1 && outputdim >= 1, UITK_EINVAL, "bad shape");
  UITK_REQUIRE(reinterpret_cast<uintptr_t>(d_workspace) % 256 == 0 && reinterpret_cast<uintptr_t>(d_blob) % 256 == 0, UITK_EALIGN,
               "workspace and blob need 256-byte alignment");
  UITK_REQUIRE(workspace_bytes >= uitk_mnv2_workspace_bytes(B, T), UITK_ENOSPACE, "workspace too small: need %zu, have %zu",
               uitk_mnv2_workspace_bytes(B, T), workspace_bytes);
  if (B == 0) return UITK_OK;
  int dev = 0, major = 0;
  UITK_CHECK_CUDA(cudaGetDevice(&dev));
  UITK_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  UITK_REQUIRE(major == 10, UITK_EARCH, "libuitk is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const Net net = make_net(outputdim);
  const float* W = reinterpret_cast<const float*>(d_blob);
  const size_t per = (per_clip_floats(T) * (size_t)(B < kChunk ? B : kChunk) + 63) / 64 * 64;
  float* buf[3];
  for (int i = 0; i < 3; ++i) buf[i] = reinterpret_cast<float*>(d_workspace) + i * per;
  for (int64_t b0 = 0; b0 < B; b0 += kChunk) {
    const long long nb = B - b0 < kChunk ? B - b0 : kChunk;
    int H = 64, Wd = (int)T;
    for (const Layer& l : net.layers) {
      const float* src = l.src < 0 ? d_db + b0 * 64 * T : buf[l.src];
      float* dst = buf[l.dst];
      switch (l.type) {
        case L_STEM: {
          const int Ho = down(H, 2), Wo = down(Wd, 2);
          stem_kernel<<<grid_for(nb * Ho * Wo * 8), 256, 0, s>>>(src, nb, H, Wd, Ho, Wo, W + l.w, W + l.b, dst);
          H = Ho; Wd = Wo;
          break;
        }
        case L_DW: {
          const int Ho = down(H, l.stride), Wo = down(Wd, l.stride);
          dw_kernel<<<grid_for(nb * Ho * Wo * (l.cout / 4)), 256, 0, s>>>(src, nb, H, Wd, l.cout, l.stride, Ho, Wo, W + l.w, W + l.b, dst);
          H = Ho; Wd = Wo;
          break;
        }
        case L_PW: {
          const long long M = nb * H * Wd;
          const dim3 g((unsigned)((M + 63) / 64), (unsigned)((l.cout + 63) / 64));
          if (l.relu6) pw_kernel<EPI_RELU6, false><<<g, 256, 0, s>>>(src, M, l.cin, l.cout, W + l.w, W + l.b, dst);
          else if (l.residual) pw_kernel<EPI_NONE, true><<<g, 256, 0, s>>>(src, M, l.cin, l.cout, W + l.w, W + l.b, dst);
          else pw_kernel<EPI_NONE, false><<<g, 256, 0, s>>>(src, M, l.cin, l.cout, W + l.w, W + l.b, dst);
          break;
        }
        case L_POOL:
          pool_kernel<<<grid_for(nb * Wd * l.cout), 256, 0, s>>>(src, nb, H, Wd, l.cout, dst);
          break;
        case L_CLS: {
          const long long M = nb * Wd;
          const dim3 g((unsigned)((M + 63) / 64), (unsigned)((l.cout + 63) / 64));
          pw_kernel<EPI_SIGMOID, false><<<g, 256, 0, s>>>(src, M, l.cin, l.cout, W + l.w, W + l.b, dst);
          count_launches(1);
          time_mean_kernel<<<grid_for(nb * l.cout), 256, 0, s>>>(dst, nb, Wd, l.cout, d_probs + b0 * outputdim);
          break;
        }
      }
      count_launches(1);
      UITK_CHECK_CUDA(cudaGetLastError());
    }
  }
  return UITK_OK;
}

}  // extern "C"
