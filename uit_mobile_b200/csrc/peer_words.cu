// Exchange of the batch-maximum word between the GPUs of one box WITHOUT a collective library call: every rank publishes its
// word in a small buffer that its peers have mapped (NVLink peer memory: torch symmetric memory / CUDA IPC), and reads the
// peers' words with plain system-scope loads.  Replaces the ncclAllReduce(MAX) of one 32-bit word per step: an NCCL kernel needs a
// free CTA slot (the persistent encoder leaves none until its last tile wave) and ~40 us of host time per enqueue.
//
//   slot layout (per rank, in ITS OWN memory, mapped by every peer): kRing entries of {max bits, epoch, pad, pad}
//   publish(e):  entry[e % kRing].max = word;  fence;  entry[e % kRing].epoch = e            (1 thread)
//   collect(e):  lane r spins until peer r's entry[e % kRing].epoch == e, reads its max, warp max -> out word   (1 warp)
// A rank publishes epoch e + 1 only after its own collect(e) (same stream), and collect(e) needs every peer's publish(e): no rank
// is ever more than one epoch ahead of a peer that still has to read, so kRing = 4 entries never wrap onto unread data.
// Non-negative floats order like their bit patterns (the word is the bit pattern of the largest mel power): integer max.
#include "uitk_common.cuh"

namespace uitk {

namespace {

constexpr int kRing = 4;

__global__ void words_publish_kernel(uint32_t* __restrict__ my_slot, const uint32_t* __restrict__ word, uint32_t epoch) {
  uint32_t* e = my_slot + (epoch % kRing) * 4;
  const uint32_t w = *word;
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(e), "r"(w) : "memory");
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(e + 1), "r"(epoch) : "memory");   // orders the word before the epoch
}

__global__ void words_collect_kernel(const uint32_t* const* __restrict__ peer_slots, int n, uint32_t epoch, uint32_t* __restrict__ out) {
  const int lane = threadIdx.x;
  uint32_t w = 0;
  for (int r = lane; r < n; r += 32) {
    const uint32_t* e = peer_slots[r] + (epoch % kRing) * 4;
    uint32_t seen;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(e + 1) : "memory");
      if (seen == epoch) break;
      if ((spins & 1023u) == 1023u) {     // a peer that never publishes (crashed rank, mismatched call sequence): fail loudly, do not hang
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ll) __trap();      // 20 s
      }
    }
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(e) : "memory");
    w = max(w, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if (lane == 0) *out = w;
}

}  // namespace

size_t peer_words_slot_bytes() { return (size_t)kRing * 4 * sizeof(uint32_t); }

int launch_words_publish(uint32_t* my_slot, const uint32_t* word, uint32_t epoch, cudaStream_t s) {
  words_publish_kernel<<<1, 1, 0, s>>>(my_slot, word, epoch);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

int launch_words_collect(const uint32_t* const* peer_slots, int n, uint32_t epoch, uint32_t* out, cudaStream_t s) {
  words_collect_kernel<<<1, 32, 0, s>>>(peer_slots, n, epoch, out);
  count_launches(1);
  UITK_CHECK_CUDA(cudaGetLastError());
  return UITK_OK;
}

}  // namespace uitk
