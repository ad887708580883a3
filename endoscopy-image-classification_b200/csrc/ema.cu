// K8: multi-tensor EMA (code/ema.py:51-62) -- one launch over a device-resident
// block table instead of 4 eager launches per state-dict tensor.
//
// HBM-bound: 3 x sizeof(elem) algorithmic bytes per unique state element (read
// e, read m, write e).  Each table row is one <= kEmaBlockElems chunk of one
// tensor and carries the two base pointers, so a CTA needs a single dependent
// 32-byte load before its 128-bit streaming loads can issue; every thread has
// 2 x kUnroll 16-byte loads in flight.
//
// Bit-exactness: the eager reference rounds after each of mul, mul,
// add, with both python scalars rounded to the tensor's opmath type first, so
// the update is __fmul_rn, __fmul_rn, __fadd_rn -- never an FMA.  `repeat`
// re-applies the update in registers for storages that appear more than once
// in state_dict() (custom_model.py:194-200).
#include <stdlib.h>

#include "common.cuh"

namespace b200ssl {
namespace {

constexpr int kEmaThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }

template <typename T>
__device__ __forceinline__ float ema_step(float e, float m, float d, float o, int repeat) {
  for (int r = 0; r < repeat; ++r)
    e = round_to<T>(__fadd_rn(round_to<T>(__fmul_rn(d, e)), round_to<T>(__fmul_rn(o, m))));
  return e;
}

template <typename T> struct VecN { static constexpr int N = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ float load1(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float load1<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void store1(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void store1<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename T>
__device__ __forceinline__ void ema_block_float(T* __restrict__ e, const T* __restrict__ m, int count, float d, float o,
                                                int repeat, int mode) {
  constexpr int N = VecN<T>::N;
  const int tid = threadIdx.x;
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(m)) & 15u) == 0) {
    const int nvec = count / N;
    for (int base = 0; base < nvec; base += kEmaThreads * kUnroll) {
      uint4 ev[kUnroll], mv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int v = base + u * kEmaThreads + tid;
        if (v < nvec) {
          mv[u] = ldg128(m + (size_t)v * N);
          if (mode == 0) ev[u] = ld_stream(e + (size_t)v * N);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int v = base + u * kEmaThreads + tid;
        if (v < nvec) {
          if (mode == 0) {
            float ef[N], mf[N];
            unpack16(ev[u], ef, T());
            unpack16(mv[u], mf, T());
#pragma unroll
            for (int i = 0; i < N; ++i) ef[i] = ema_step<T>(ef[i], mf[i], d, o, repeat);
            stg128(e + (size_t)v * N, pack16(ef, T()));
          } else {
            stg128(e + (size_t)v * N, mv[u]);
          }
        }
      }
    }
    done = nvec * N;
  }
  for (int i = done + tid; i < count; i += kEmaThreads) {
    const float mf = load1(m + i);
    store1(e + i, mode == 0 ? ema_step<T>(load1(e + i), mf, d, o, repeat) : mf);
  }
}

// integer buffers (num_batches_tracked): fp32 arithmetic, truncating copy_ (quirk Q3)
template <typename I>
__device__ __forceinline__ void ema_block_int(I* __restrict__ e, const I* __restrict__ m, int count, float d, float o,
                                              int repeat, int mode) {
  for (int i = threadIdx.x; i < count; i += kEmaThreads) {
    I ev = e[i];
    const I mv = m[i];
    if (mode == 0) {
      for (int r = 0; r < repeat; ++r) ev = (I)__fadd_rn(__fmul_rn(d, (float)ev), __fmul_rn(o, (float)mv));
    } else {
      ev = mv;
    }
    e[i] = ev;
  }
}

// One instantiation per floating storage type; integer buffers ride along in
// every launch's first pass (FIRST), other float types are skipped.
template <typename T, int DT>
__global__ void __launch_bounds__(kEmaThreads, 4) ema_multi_tensor_kernel(const b200ssl_ema_block* __restrict__ blocks,
                                                                          int n_blocks, float d, float o, int mode,
                                                                          int do_ints) {
  pdl_launch_dependents();
  pdl_wait();                                               // the weights may still be written by the previous kernel
  for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const uint4 lo = ldg128(&blocks[b]);
    const uint4 hi = ldg128(reinterpret_cast<const char*>(&blocks[b]) + 16);
    void* e = reinterpret_cast<void*>(((unsigned long long)lo.y << 32) | lo.x);
    const void* m = reinterpret_cast<const void*>(((unsigned long long)lo.w << 32) | lo.z);
    const int count = (int)hi.x, dtype = (int)hi.y, repeat = (int)hi.z;
    if (dtype == DT) {
      ema_block_float<T>((T*)e, (const T*)m, count, d, o, repeat, mode);
    } else if (do_ints) {
      if (dtype == B200SSL_I64) ema_block_int<long long>((long long*)e, (const long long*)m, count, d, o, repeat, mode);
      else if (dtype == B200SSL_I32) ema_block_int<int>((int*)e, (const int*)m, count, d, o, repeat, mode);
      else if (dtype == B200SSL_U8) ema_block_int<unsigned char>((unsigned char*)e, (const unsigned char*)m, count, d, o, repeat, mode);
    }
  }
}

// ---- the update next to other kernels: a fixed set of SMs left alone ------------------------------------------------------
// `mask` (one bit per SM id) marks SMs that the update must not use: a CTA that finds itself on one exits at once, the
// others fetch 4096-element chunks from a device-side counter until the table is done.  Which SMs are free for the kernels
// next to the update no longer depends on who was placed first.  probe_sm_set_kernel finds a set that clusters fit in.
__device__ __forceinline__ unsigned smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}

template <typename T, int DT>
__global__ void __launch_bounds__(kEmaThreads, 4) ema_masked_kernel(const b200ssl_ema_block* __restrict__ blocks, int n_blocks, float d,
                                                                    float o, int mode, int do_ints, const unsigned* __restrict__ mask,
                                                                    unsigned* sched) {
  __shared__ int s_next[2];
  pdl_launch_dependents();
  pdl_wait();                                               // the weights may still be written by the previous kernel
  const unsigned sm = smid();
  const bool reserved = (__ldg(&mask[sm >> 5]) >> (sm & 31)) & 1u;
  if (!reserved) {
    if (threadIdx.x == 0) s_next[0] = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    for (int it = 0;; ++it) {
      const int b = s_next[it & 1];
      if (b >= n_blocks) break;
      if (threadIdx.x == 0) s_next[(it + 1) & 1] = (int)atomicAdd(&sched[0], 1u);    // next chunk, fetched under this one
      const uint4 lo = ldg128(&blocks[b]);
      const uint4 hi = ldg128(reinterpret_cast<const char*>(&blocks[b]) + 16);
      void* e = reinterpret_cast<void*>(((unsigned long long)lo.y << 32) | lo.x);
      const void* m = reinterpret_cast<const void*>(((unsigned long long)lo.w << 32) | lo.z);
      const int count = (int)hi.x, dtype = (int)hi.y, repeat = (int)hi.z;
      if (dtype == DT) {
        ema_block_float<T>((T*)e, (const T*)m, count, d, o, repeat, mode);
      } else if (do_ints) {
        if (dtype == B200SSL_I64) ema_block_int<long long>((long long*)e, (const long long*)m, count, d, o, repeat, mode);
        else if (dtype == B200SSL_I32) ema_block_int<int>((int*)e, (const int*)m, count, d, o, repeat, mode);
        else if (dtype == B200SSL_U8) ema_block_int<unsigned char>((unsigned char*)e, (const unsigned char*)m, count, d, o, repeat, mode);
      }
      __syncthreads();
    }
  }
  // the last CTA of the grid re-arms the scheduler for the next launch (stream order; CUDA-graph replay safe)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
      sched[0] = 0u;
      sched[1] = 0u;
      __threadfence();
    }
  }
}

// `gridDim.x` CTAs in clusters of `cluster` CTAs, one CTA per SM (shared-memory request), all resident at once: each marks
// its SM in `mask`.  The marked set is one in which that many clusters fit side by side.
__global__ void probe_sm_set_kernel(unsigned* mask, unsigned* counter) {
  const unsigned sm = smid();
  if (threadIdx.x == 0) {
    atomicOr(&mask[sm >> 5], 1u << (sm & 31));
    __threadfence();
    atomicAdd(counter, 1u);
    for (unsigned it = 0; *reinterpret_cast<volatile unsigned*>(counter) < gridDim.x && it < (1u << 22); ++it) {}   // bounded: never hangs
  }
  __syncthreads();
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

// one thread that returns after `ns` nanoseconds: put ahead of an overlapped update on its side stream, it lets the kernel
// queued at the same time on the other stream place its CTAs first (see b200ssl_stream_delay)
__global__ void delay_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

static int ema_launch(const char* fn, const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype, int32_t do_ints,
                      float decay, float one_minus_decay, int32_t mode, int32_t max_ctas, void* stream) {
  static_assert(sizeof(b200ssl_ema_block) == 32, "block table row must be 32 bytes");
  if (!blocks) return fail(B200SSL_E_NULL, "%s: NULL block table", fn);
  if (reinterpret_cast<uintptr_t>(blocks) & 15u) return fail(B200SSL_E_ALIGN, "%s: block table must be 16-byte aligned", fn);
  if (n_blocks <= 0) return fail(B200SSL_E_SHAPE, "%s: n_blocks must be > 0", fn);
  if (mode != 0 && mode != 1) return fail(B200SSL_E_ARG, "%s: mode %d (0 update, 1 set)", fn, mode);
  if (max_ctas < 0) return fail(B200SSL_E_ARG, "%s: max_ctas %d < 0", fn, max_ctas);
  // 4 resident CTAs of 256 threads per SM; max_ctas caps the grid (see b200ssl_ema_multi_tensor_ctas)
  const int max_grid = (max_ctas > 0 && max_ctas < kNumSMs * 4) ? max_ctas : kNumSMs * 4;
  const int grid = n_blocks < max_grid ? n_blocks : max_grid;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaSuccess;
  switch (float_dtype) {
    case B200SSL_F32:
      e = launch_pdl(PDL_EMA, ema_multi_tensor_kernel<float, B200SSL_F32>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks, n_blocks,
                     decay, one_minus_decay, mode, do_ints);
      break;
    case B200SSL_BF16:
      e = launch_pdl(PDL_EMA, ema_multi_tensor_kernel<__nv_bfloat16, B200SSL_BF16>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks,
                     n_blocks, decay, one_minus_decay, mode, do_ints);
      break;
    case B200SSL_F16:
      e = launch_pdl(PDL_EMA, ema_multi_tensor_kernel<__half, B200SSL_F16>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks, n_blocks,
                     decay, one_minus_decay, mode, do_ints);
      break;
    default:
      return fail(B200SSL_E_DTYPE, "%s: float_dtype %d (want F32, BF16 or F16)", fn, float_dtype);
  }
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

extern "C" int b200ssl_ema_multi_tensor(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype,
                                        int32_t do_ints, float decay, float one_minus_decay, int32_t mode,
                                        void* stream) {
  return ema_launch("b200ssl_ema_multi_tensor", blocks, n_blocks, float_dtype, do_ints, decay, one_minus_decay, mode, 0, stream);
}

extern "C" int b200ssl_ema_multi_tensor_ctas(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype,
                                             int32_t do_ints, float decay, float one_minus_decay, int32_t mode,
                                             int32_t max_ctas, void* stream) {
  return ema_launch("b200ssl_ema_multi_tensor_ctas", blocks, n_blocks, float_dtype, do_ints, decay, one_minus_decay, mode, max_ctas, stream);
}

extern "C" int b200ssl_stream_delay(int64_t nanoseconds, void* stream) {
  if (nanoseconds < 0 || nanoseconds > 1000000) return fail(B200SSL_E_ARG, "b200ssl_stream_delay: %lld ns outside [0, 1e6]", (long long)nanoseconds);
  if (nanoseconds == 0) return 0;
  delay_kernel<<<1, 1, 0, as_stream(stream)>>>((unsigned long long)nanoseconds);
  return check_launch("b200ssl_stream_delay");
}

extern "C" int b200ssl_probe_sm_set(int32_t n_clusters, int32_t cluster, uint32_t* mask8, uint32_t* counter, void* stream) {
  const char* fn = "b200ssl_probe_sm_set";
  if (!mask8 || !counter) return fail(B200SSL_E_NULL, "%s: NULL buffer", fn);
  if (n_clusters < 1 || cluster < 1 || cluster > 8 || (cluster & (cluster - 1)) || n_clusters * cluster > kNumSMs / 2)
    return fail(B200SSL_E_ARG, "%s: %d clusters of %d (cluster a power of two <= 8, at most half the SMs)", fn, n_clusters, cluster);
  const size_t smem = 160 * 1024;                          // more than half an SM: one CTA per SM
  cudaError_t e = cudaFuncSetAttribute(probe_sm_set_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    e = launch_pdl(PDL_EMA, probe_sm_set_kernel, dim3((unsigned)(n_clusters * cluster)), dim3(32), smem, as_stream(stream),
                   dim3((unsigned)cluster, 1, 1), mask8, counter);
  if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

extern "C" int b200ssl_ema_multi_tensor_masked(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype, int32_t do_ints,
                                               float decay, float one_minus_decay, int32_t mode, const uint32_t* sm_mask8,
                                               uint32_t* sched2, void* stream) {
  const char* fn = "b200ssl_ema_multi_tensor_masked";
  if (!blocks || !sm_mask8 || !sched2) return fail(B200SSL_E_NULL, "%s: NULL argument", fn);
  if (reinterpret_cast<uintptr_t>(blocks) & 15u) return fail(B200SSL_E_ALIGN, "%s: block table must be 16-byte aligned", fn);
  if (n_blocks <= 0) return fail(B200SSL_E_SHAPE, "%s: n_blocks must be > 0", fn);
  if (mode != 0 && mode != 1) return fail(B200SSL_E_ARG, "%s: mode %d (0 update, 1 set)", fn, mode);
  const int grid = kNumSMs * 4;                            // four CTAs per SM everywhere; the ones on masked SMs leave at once
  cudaStream_t st = as_stream(stream);
  cudaError_t e;
  switch (float_dtype) {
    case B200SSL_F32:
      e = launch_pdl(PDL_EMA, ema_masked_kernel<float, B200SSL_F32>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks, n_blocks, decay,
                     one_minus_decay, mode, do_ints, sm_mask8, sched2);
      break;
    case B200SSL_BF16:
      e = launch_pdl(PDL_EMA, ema_masked_kernel<__nv_bfloat16, B200SSL_BF16>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks,
                     n_blocks, decay, one_minus_decay, mode, do_ints, sm_mask8, sched2);
      break;
    case B200SSL_F16:
      e = launch_pdl(PDL_EMA, ema_masked_kernel<__half, B200SSL_F16>, dim3(grid), dim3(kEmaThreads), 0, st, dim3(1, 1, 1), blocks, n_blocks,
                     decay, one_minus_decay, mode, do_ints, sm_mask8, sched2);
      break;
    default:
      return fail(B200SSL_E_DTYPE, "%s: float_dtype %d (want F32, BF16 or F16)", fn, float_dtype);
  }
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}
