// UMMA plumbing self-test (tests only): one 128 x N x K tcgen05 GEMM through the same descriptor / TMEM / bulk-copy
// helpers the encoder megakernel uses (tc_ptx.cuh).
#include "tc_ptx.cuh"
#include "uitk_common.cuh"

namespace uitk {

namespace {

using namespace tc;

// ---------------------------------------------------------------------------------------------------------------
// UMMA plumbing self-test: C[128 x N] (+)= A[128 x K] * Bp^T, Bp already packed (bf16, K-major core-matrix layout).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) umma_selftest_kernel(const float* __restrict__ A, const unsigned char* __restrict__ Bp,
                                                               const float* __restrict__ Cinit, float* __restrict__ C, int N, int K, int a_in_tmem) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sAp = smem;                   // up to 128 x 256 bf16 = 64 KB
  unsigned char* sBp = smem + 65536;           // up to 128 x 256 bf16 = 64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 131072 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hsel = warp >> 2, r = q * 32 + lane;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tx = tmem + ((uint32_t)(q * 32) << 16);
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)N * K * 2;
    mbar_arrive_expect_tx(&bars[0], bytes);
    bulk_g2s(sBp, Bp, bytes, &bars[0]);
  }
  // A: thread (r, hsel) converts its half of the k-groups (shared memory) or packed column blocks (tensor memory)
  const int k8n = K / 8;
  if (a_in_tmem) {
    for (int c0 = hsel * 4; c0 < K / 2; c0 += 8) {        // 4 packed columns = 8 consecutive k
      uint32_t v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = pack2_bf16(A[(size_t)r * K + 2 * (c0 + i)], A[(size_t)r * K + 2 * (c0 + i) + 1]);
      tmem_st4(tx + 256 + c0, v);
    }
    tmem_st_wait();
  } else
  for (int k8 = hsel; k8 < k8n; k8 += 2) {
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = A[(size_t)r * K + k8 * 8 + i];
    *reinterpret_cast<uint4*>(sAp + k8 * 2048 + r * 16) = pack8_bf16(y);
  }
  if (Cinit != nullptr) {
    for (int c0 = hsel * 32; c0 < N; c0 += 64) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (c0 + i < N) ? Cinit[(size_t)r * N + c0 + i] : 0.f;
      tmem_st32(tx + c0, v);
    }
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t lbo_b = (uint32_t)N * 16;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t db = make_smem_desc(smem_u32(sBp) + ks * 2 * lbo_b, lbo_b, 128);
      const uint32_t acc = (Cinit != nullptr || ks > 0) ? 1u : 0u;
      if (a_in_tmem) umma_bf16_ts(tmem, tmem + 256 + ks * 8, db, idesc, acc);
      else umma_bf16(tmem, make_smem_desc(smem_u32(sAp) + ks * 4096, 2048, 128), db, idesc, acc);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  __syncwarp();
  tc_fence_after();
  for (int c0 = hsel * 32; c0 < N; c0 += 64) {
    float v[32];
    tmem_ld32(tx + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) C[(size_t)r * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

int run_umma_selftest(const float* A, const void* Bp, const float* Cinit, float* C, int N, int K, int a_in_tmem, cudaStream_t s) {
  UITK_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, UITK_EINVAL, "selftest: bad N/K");
  const int smem = 131072 + 256;
  UITK_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 256, smem, s>>>(A, reinterpret_cast<const unsigned char*>(Bp), Cinit, C, N, K, a_in_tmem);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

}  // namespace uitk
