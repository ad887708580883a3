// Shared definitions of the NVLink peer-memory arena (see peer.cu for the layout and the protocol).
#pragma once
#include "common.cuh"

namespace b200ssl {
namespace peer {


constexpr int kMaxWorld = 16;
constexpr int kMaxExchange = 8;
constexpr size_t kFlagBytes = 1024;                 // flags[kMaxExchange][kMaxWorld] u64
constexpr size_t kCtlBytes = 4096;
constexpr int kPeerThreads = 256;
constexpr int kChunkBytes = 16384;                  // one CTA moves 16 KB: 4 x 16 B per thread
constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct LocalCtl {                                   // at arena + kFlagBytes
  unsigned long long epoch[kMaxExchange];
  unsigned int done[kMaxExchange];                  // CTAs of the running launch that have finished
  unsigned int pushed[kMaxExchange][kMaxWorld];     // CTAs that have finished pushing to one destination
  unsigned int timeouts;                            // sticky: a wait gave up (results of that step are garbage)
};
static_assert(sizeof(LocalCtl) <= kCtlBytes - kFlagBytes, "control block overflows its page");

// Exchange ids 0..5 belong to the caller (PeerArena regions); the directly addressed bank uses the last two:
constexpr int kXSmoothDone = 6;                     // "my smoothing pass of step e has read every shard"
constexpr int kXEnqueueDone = 7;                    // "my rows of step e are in their shards"

__device__ __forceinline__ LocalCtl* local_ctl(uint8_t* arena) { return reinterpret_cast<LocalCtl*>(arena + kFlagBytes); }
__device__ __forceinline__ unsigned long long* flag_of(uint8_t* arena, int x, int src) {
  return reinterpret_cast<unsigned long long*>(arena) + x * kMaxWorld + src;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_cg(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }



// Waits until flags[x][s] >= epoch for the sources s handled by the calling thread.
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch, LocalCtl* ctl) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > kTimeoutNs) {
      atomicAdd(&ctl->timeouts, 1u);
      return;
    }
  }
}

}  // namespace peer
}  // namespace b200ssl
