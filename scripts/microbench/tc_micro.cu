// Microbenchmarks that size the tensor-core encoder's design: tcgen05.ld throughput vs. number of warps,
// MMA issue->commit->wake round trip, fence.proxy.async cost, cross-warp mbarrier hand-off.
// Build+run on the GPU box: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tc_micro scripts/microbench/tc_micro.cu && /tmp/tc_micro
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../uit_mobile_b200/csrc/tc_ptx.cuh"

using namespace uitk::tc;

__global__ void __launch_bounds__(512, 1) micro(long long* out, int nw_ld, int mma_n, int mma_k, int a_tmem) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  __shared__ long long stamp[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // ---- A: tcgen05.ld.32x32b.x32 + wait, 64 iterations, warps [0, nw_ld) ----
  long long tA = 0;
  {
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    if (warp < nw_ld) {
      const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int it = 0; it < 64; ++it) {
        float v[32];
        tmem_ld32(base + ((it * 32 + (warp >> 2) * 128) & 511), v);
        tmem_ld_wait();
        acc += v[0] + v[31];
      }
    }
    const long long t1 = clock64();
    if (acc == 123.f) out[100] = 1;
    if (tid == 0) tA = t1 - t0;
    __syncthreads();
  }
  // ---- A2: same but two loads in flight (x32 + x32 then one wait) ----
  long long tA2 = 0;
  {
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    if (warp < nw_ld) {
      const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int it = 0; it < 32; ++it) {
        float v[32], w[32];
        tmem_ld32(base + ((it * 64 + (warp >> 2) * 128) & 511), v);
        tmem_ld32(base + ((it * 64 + 32 + (warp >> 2) * 128) & 511), w);
        tmem_ld_wait();
        acc += v[0] + w[31];
      }
    }
    const long long t1 = clock64();
    if (acc == 123.f) out[100] = 1;
    if (tid == 0) tA2 = t1 - t0;
    __syncthreads();
  }
  // ---- B: MMA round trip.  The whole (provably uniform) warp 0 runs the loop, one elected lane issues: descriptors
  // stay in uniform registers and the UTCHMMA instructions are issued back to back (no R2UR waterfall loops).
  long long tB = 0, tB2 = 0;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, mma_n);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 32768);
    uint32_t ph = 0;
    for (int rep = 0; rep < 4; ++rep) {
      const long long t0 = clock64();
      if (a_tmem) {
#pragma unroll 8
        for (int ks = 0; ks < mma_k; ++ks) {
          const uint64_t db = make_smem_desc(sB + (ks & 7) * 2 * (mma_n * 16), mma_n * 16, 128);
          const uint32_t ta = tmem + 256 + (ks & 7) * 8, acc = ks > 0;
          if (elect_one())
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      } else {
#pragma unroll 8
        for (int ks = 0; ks < mma_k; ++ks) {
          const uint64_t db = make_smem_desc(sB + (ks & 7) * 2 * (mma_n * 16), mma_n * 16, 128);
          if (elect_one()) umma_bf16(tmem, make_smem_desc(sA + (ks & 7) * 4096, 2048, 128), db, idesc, ks > 0);
        }
      }
      if (elect_one()) umma_commit(&bars[0]);
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bars[0], ph); ph ^= 1;
      tc_fence_after();
      const long long t2 = clock64();
      tB = t2 - t0; tB2 = t1 - t0;
    }
  }
  __syncthreads();
  // ---- C: st.shared + fence.proxy.async ----
  long long tC = 0;
  {
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < 16; ++it) {
      *reinterpret_cast<uint4*>(smem + 65536 + tid * 16) = make_uint4(it, 0, 0, 0);
      fence_proxy_async_smem();
    }
    const long long t1 = clock64();
    if (tid == 0) tC = (t1 - t0) / 16;
    __syncthreads();
  }
  // ---- D: cross-warp mbarrier hand-off: warp 1 lane 0 arrives, warp 0 lane 0 (already waiting) wakes ----
  long long tD = 0;
  {
    __syncthreads();
    if (tid == 32) {
      for (int i = 0; i < 2000; ++i) asm volatile("nanosleep.u32 20;");
      stamp[0] = clock64();
      mbar_arrive(&bars[1]);
    }
    if (tid == 0) {
      mbar_wait(&bars[1], 0);
      stamp[1] = clock64();
    }
    __syncthreads();
    if (tid == 0) tD = stamp[1] - stamp[0];
  }
  // ---- E: named barrier among 256 threads ----
  long long tE = 0;
  {
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 256) for (int i = 0; i < 16; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
    const long long t1 = clock64();
    if (tid == 0) tE = (t1 - t0) / 16;
    __syncthreads();
  }
  if (tid == 0 && blockIdx.x == 0) { out[0] = tA; out[1] = tA2; out[2] = tB; out[3] = tB2; out[4] = tC; out[5] = tD; out[6] = tE; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 1024); long long h[8];
  cudaFuncSetAttribute(micro, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int nws[] = {1, 2, 4, 8, 16};
  for (int grid : {1, 148}) {
    for (int nw : nws) {
      micro<<<grid, 512, 128 * 1024>>>(d, nw, 128, 1, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      printf("grid %3d  ld warps %2d: [ld x32 + wait] %6.1f cyc/iter (%.1f B/cyc/SM)   [2 x ld x32 + wait] %6.1f cyc/iter (%.1f B/cyc/SM)\n", grid, nw,
             h[0] / 64.0, nw * 4096.0 * 64 / h[0], h[1] / 32.0, nw * 8192.0 * 32 / h[1]);
    }
  }
  for (int ts : {0, 1}) for (int n : {16, 32, 64, 128, 256}) for (int k : {1, 8, 24}) {
    if (ts && n == 256) continue;
    micro<<<1, 512, 128 * 1024>>>(d, 1, n, k, ts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("MMA %s M128 N%3d x %2d k-steps: issue+commit %5lld cyc, round trip (issue -> commit -> mbarrier wake) %5lld cyc\n", ts ? "TS(A in TMEM)" : "SS", n, k, h[3], h[2]);
  }
  // ---- the front-end north_star sketches: the 512-point DFT as tcgen05 contractions, factored 16 x 32.  Operands bf16
  // hi + lo, 3 products per k-step (2^-16 relative; fp32-grade needs hi/mid/lo = 6 products).  Only the MMA issue stream of a
  // 128-row tile is timed, operands already in SMEM / TMEM - no framing, windowing, twiddles, operand splits, TMEM round trips.
  //   stage 1: 32 real 16-point DFTs per frame: rows = 4 frames x 32, K = 16 (1 k-step), N = 32        -> 3 k-steps / 4 frames (SS)
  //   stage 2: 16 complex 32-point DFTs per frame: rows = 8 frames x 16, K = 64 (4 k-steps), N = 64    -> 12 k-steps / 8 frames (TS)
  {
    double cyc_frame[2] = {0, 0};
    for (int prod : {3, 6}) {
      micro<<<148, 512, 128 * 1024>>>(d, 1, 32, 32 * prod, 0);             // 32 tiles of stage 1 back to back
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      const double s1 = (double)h[2] / (32.0 * 4.0);                         // cycles per frame
      micro<<<148, 512, 128 * 1024>>>(d, 1, 64, 8 * 4 * prod, 1);           // 8 tiles of stage 2
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      const double s2 = (double)h[2] / (8.0 * 8.0);
      cyc_frame[prod == 6] = s1 + s2;
      printf("two-stage tcgen05 DFT, %d products/k-step: stage 1 %.1f + stage 2 %.1f = %.1f tensor-pipe cycles per frame and SM -> %.3f ms for the bench's "
             "413 696 frames on 148 SMs at 1.965 GHz (MMA issue stream ALONE)\n", prod, s1, s2, s1 + s2, (s1 + s2) * 413696.0 / 148.0 / 1.965e6);
    }
  }
  printf("st.shared + fence.proxy.async: %lld cyc; mbarrier arrive -> waiter wake: %lld cyc; bar.sync 256 threads: %lld cyc\n", h[4], h[5], h[6]);
  return 0;
}
