// Shared device/host helpers for libb200ssl (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "b200ssl.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200ssl is written for sm_100a (B200) only"
#endif

namespace b200ssl {

// ---- host side error plumbing (api.cu) -------------------------------------
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// Optional device-side timing probe (tools/kernel_timeline.py): when a buffer is registered with
// b200ssl_debug_set_timing_buffer(), instrumented kernels store clock64() stamps per CTA:
// dbg[cta * 16 + slot].  NULL (the default) costs one predictable branch per stamp.
constexpr size_t kDebugRegion = 4096 * 16;   // u64 slots per instrumented kernel (region index = its PDL_* tag)
unsigned long long* debug_timing_buffer(int kernel_tag);
#define B200SSL_STAMP(dbg, cta, slot)                                            \
  do {                                                                          \
    if ((dbg) != nullptr) (dbg)[(size_t)(cta) * 16 + (slot)] = clock64();       \
  } while (0)
// wall-clock variant (%globaltimer, ns): comparable across SMs -- start skew and tail of a launch
#define B200SSL_STAMP_NS(dbg, cta, slot)                                         \
  do {                                                                          \
    if ((dbg) != nullptr) {                                                     \
      unsigned long long t_;                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
      (dbg)[(size_t)(cta) * 16 + (slot)] = t_;                                  \
    }                                                                           \
  } while (0)

// Workspace layout (caller zero-fills it once): [0,256) ticket counters of the
// row kernels, [256, 256+64K) per-row-tile tickets of the similarity kernels,
// then float partials.  Tickets reset themselves, partials are scratch.
constexpr size_t kWsTicketBytes = 256;
constexpr size_t kWsTicket2Bytes = 64 * 1024;
constexpr size_t kWsHeaderBytes = kWsTicketBytes + kWsTicket2Bytes;
constexpr int kMaxRowCtas = 4 * kNumSMs;
int smooth_nsplit(long long rows, long long bank_rows, int* tiles_per_split);
int contrast_nsplit(long long rows, int modes, int* tiles_per_split);
size_t contrast_workspace_floats(long long rows, int dim);
// SMs the head's kernels should confine themselves to (0 = no limit): set by ModelEMA(overlap=True), whose capped update leaves
// that many SMs free for the whole update -- a head kernel that takes more would hand its surplus SMs to the update's pending
// CTAs when it retires, and the kernels after it would find no whole SM (b200ssl_set_head_sm_budget)
extern int g_head_sm_budget;
size_t smooth_tc_workspace_floats(long long rows, long long bank_rows, int classes);
size_t smooth_tc_f32_workspace_bytes(long long rows, long long bank_rows, int classes);   // fold partials + split operand copies (fp32 storage)
size_t contrast_tc_workspace_floats(long long rows);

// ---- programmatic dependent launch (PDL) -------------------------------------
// Every hot kernel (a) lets the next kernel of the stream be scheduled as soon as it has started itself and
// (b) waits for the complete previous kernel (incl. its memory flush) before it touches global memory.  The
// launch latency, barrier init, TMEM allocation and descriptor prefetch of kernel N+1 then overlap the tail
// of kernel N -- at the reference's sizes the head is a chain of ~5 launch-latency-bound kernels per step.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// cudaLaunchKernelEx with the PDL attribute and an optional thread-block cluster.
enum { PDL_SMOOTH = 0, PDL_ROWS = 1, PDL_CONTRAST_FWD = 2, PDL_CONTRAST_BWD = 3, PDL_EMA = 4, PDL_PEER = 5 };
constexpr int kPdlDefaultMask = 15;   // the four head kernels; measured: the HBM-bound EMA launch loses 4 us when it is scheduled early
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int tag, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, dim3 cluster,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  static const int pdl_mask = getenv("B200SSL_PDL_MASK") ? atoi(getenv("B200SSL_PDL_MASK")) : kPdlDefaultMask;   // A/B aid
  if ((pdl_mask >> tag) & 1) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dtype helpers ---------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct Vec16<__half> { static constexpr int N = 8; };

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void unpack16(const uint4& v, float (&out)[4], float) {
  out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
  out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w);
}
__device__ __forceinline__ void unpack16(const uint4& v, float (&out)[8], __nv_bfloat16) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(w[i] << 16);
    out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack16(const float (&in)[4], float) {
  return make_uint4(__float_as_uint(in[0]), __float_as_uint(in[1]), __float_as_uint(in[2]), __float_as_uint(in[3]));
}
__device__ __forceinline__ uint4 pack16(const float (&in)[8], __nv_bfloat16) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void unpack16(const uint4& v, float (&out)[8], __half) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
    out[2 * i] = __low2float(h);
    out[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack16(const float (&in)[8], __half) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(in[2 * i], in[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// Copy `count` contiguous elements global -> shared (as fp32) with 128-bit
// loads when the global address is 16-byte aligned; all threads of the CTA.
template <typename T>
__device__ __forceinline__ void tile_g2s(const T* __restrict__ g, float* __restrict__ s, int count) {
  constexpr int N = Vec16<T>::N;
  const int tid = threadIdx.x, nt = blockDim.x;
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int nvec = count / N;
    for (int v = tid; v < nvec; v += nt) {
      float f[N];
      unpack16(ldg128(g + (size_t)v * N), f, T());
#pragma unroll
      for (int i = 0; i < N; ++i) s[v * N + i] = f[i];
    }
    for (int i = nvec * N + tid; i < count; i += nt) s[i] = to_f32<T>(g[i]);
  } else {
    for (int i = tid; i < count; i += nt) s[i] = to_f32<T>(g[i]);
  }
}
// shared (fp32) -> global, same vectorisation rule.
template <typename T>
__device__ __forceinline__ void tile_s2g(const float* __restrict__ s, T* __restrict__ g, int count) {
  constexpr int N = Vec16<T>::N;
  const int tid = threadIdx.x, nt = blockDim.x;
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int nvec = count / N;
    for (int v = tid; v < nvec; v += nt) {
      float f[N];
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = s[v * N + i];
      stg128(g + (size_t)v * N, pack16(f, T()));
    }
    for (int i = nvec * N + tid; i < count; i += nt) g[i] = from_f32<T>(s[i]);
  } else {
    for (int i = tid; i < count; i += nt) g[i] = from_f32<T>(s[i]);
  }
}

// ---- sub-warp (LPR lanes per row) butterfly reductions -----------------------
template <int LPR> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPR> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// (value, index) arg-max with first-index tie-break (torch.max semantics).
template <int LPR> __device__ __forceinline__ void group_argmax(float& v, int& i) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

// Fold `nsplit` partial vectors (each `count4` float4 long, `stride4` float4 apart) in
// split order and hand float4 #i of the total to f(i, sum).  The loads of 8 splits are
// issued back to back (independent, L2-only) so the fold costs a few round trips instead
// of nsplit * count4 / blockDim serialised ones.
template <typename F>
__device__ __forceinline__ void fold_splits_vec4(const float4* __restrict__ part, size_t stride4, int nsplit, int count4, F f) {
  for (int i = threadIdx.x; i < count4; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < nsplit; s0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (s0 + u < nsplit) v[u] = __ldcg(part + (size_t)(s0 + u) * stride4 + i);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (s0 + u < nsplit) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    f(i, acc);
  }
}

// The same fold with the splits dealt to thread groups: `count4` float4 items, P = blockDim / count4 groups, group g adds the
// splits [g * per, (g + 1) * per) in order (8 loads in flight), the partial sums meet in `scratch` ([P][count4] float4 of
// shared memory) and are added in group order.  A long fold (tens of splits of a short vector) costs ceil(nsplit / (8 P))
// L2 round trips instead of ceil(nsplit / 8).  Fixed order for a fixed launch geometry: deterministic.  All threads of
// the CTA must call it.
template <typename F>
__device__ __forceinline__ void fold_splits_wide(const float4* __restrict__ part, size_t stride4, int nsplit, int count4, float4* scratch,
                                                 F f) {
  const int nt = blockDim.x;
  if (count4 <= 0) return;
  int P = nt / count4;
  if (P > (nsplit + 7) / 8) P = (nsplit + 7) / 8;
  if (P <= 1) {                                              // short fold or long vector: the plain strided loop
    fold_splits_vec4(part, stride4, nsplit, count4, f);
    return;
  }
  const int per = (nsplit + P - 1) / P;
  const int item = threadIdx.x % count4, grp = threadIdx.x / count4;
  if (grp < P) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int s_end = min(nsplit, (grp + 1) * per);
    for (int s0 = grp * per; s0 < s_end; s0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (s0 + u < s_end) v[u] = __ldcg(part + (size_t)(s0 + u) * stride4 + item);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (s0 + u < s_end) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    scratch[grp * count4 + item] = acc;
  }
  __syncthreads();
  if (threadIdx.x < count4) {
    float4 t = scratch[threadIdx.x];
    for (int g = 1; g < P; ++g) {
      const float4 v = scratch[g * count4 + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    f(threadIdx.x, t);
  }
  __syncthreads();
}

// Deterministic grid-wide sum of NV floats per CTA: every CTA stores its
// partials, takes a ticket; the last CTA to arrive adds all partials in a
// fixed order and returns true with the totals in `total` (valid in thread 0).
// The ticket resets itself, so the workspace stays reusable.
template <int NV>
__device__ __forceinline__ bool grid_reduce_last(const float (&mine)[NV], float* partials, unsigned* ticket,
                                                 float (&total)[NV]) {
  __shared__ bool s_last;
  __shared__ float s_red[NV][32];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) partials[(size_t)blockIdx.x * NV + v] = mine[v];
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.f;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += __ldcg(&partials[(size_t)b * NV + v]);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = warp_sum(acc[v]);
    if (lane == 0) s_red[v][warp] = acc[v];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t = 0.f;
      for (int w = 0; w < nwarp; ++w) t += s_red[v][w];
      total[v] = t;
    }
    *ticket = 0u;
  }
  return true;
}

}  // namespace b200ssl
