// Micro-benchmark (B200): throughput and round-trip latency of the small tcgen05.mma shapes K3 uses, operands in shared memory
// (128B swizzle, garbage data).  Per "unit": GEMM1 = 4 x (M128 N128 K16), GEMM2 = 8 x (M128 N32 K16), one commit each.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../endoscopy-image-classification_b200/csrc -I../../include -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace b200ssl;

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand read from tensor memory (lane = row, one 32-bit column = two bf16 of K)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}


// kind::f16 instruction descriptor with explicit operand majors (0 = K-major, 1 = MN-major), MN-major SW128 descriptor
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return tc::idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// the gradient GEMMs of the contrastive backward: 16 x (M128 N64 K16) per tile; AMN / BMN = operand read MN-major
template <int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) kg(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); abort_flag = 0; tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t base = tc::smem_u32(smem);
    constexpr uint32_t idesc = idesc_bf16(128, 64, AMN, BMN);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int b = it & 1;
      const uint64_t aK = tc::smem_desc_sw128(base), aM = smem_desc_sw128_mn(base, 16384);
      const uint64_t bK = tc::smem_desc_sw128(base + 65536), bM = smem_desc_sw128_mn(base + 65536, 0);
      if (leader) {
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t a = AMN ? aM + (uint64_t)(2 * part * 1024 + 128 * kk) : aK + (uint64_t)((2 * part + (kk >> 2)) * 1024 + 2 * (kk & 3));
            const uint64_t bb = BMN ? bM + 128 * kk : bK + 2 * (kk & 3) + (kk >> 2) * 512;
            tc::mma_bf16_ss(tmem + 256, a, bb, idesc, (it | part | kk) != 0);
          }
        tc::mma_commit(&bar[b]);
      }
      __syncwarp();
      if (it >= 1) tc::mbar_wait(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1, &abort_flag);
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = abort_flag; }
  if (warp == 0) { tc::tcgen05_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int AMN, int BMN>
void rung(const char* name, long long* d_out) {
  const int iters = 1024;
  cudaFuncSetAttribute(kg<AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  kg<AMN, BMN><<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaDeviceSynchronize();
  kg<AMN, BMN><<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"variant\": \"%s\", \"clocks_per_16_mma\": %.1f, \"abort\": %lld, \"err\": \"%s\"}\n", name, (double)h[0] / iters, h[1], cudaGetErrorString(e));
}

// MODE bit 0: GEMM1, bit 1: GEMM2; SYNC: 0 = two units in flight (throughput), 1 = wait after every unit (latency)
template <int MODE, int SYNC, int N2>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2][2];     // [gemm][unit parity]
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0][0], 1); tc::mbar_init(&bar[0][1], 1); tc::mbar_init(&bar[1][0], 1); tc::mbar_init(&bar[1][1], 1); abort_flag = 0; tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t base = tc::smem_u32(smem);
    constexpr uint32_t idesc1 = tc::idesc_bf16_f32(128, 128), idesc2 = tc::idesc_bf16_f32(128, N2);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int b = it & 1;
      if (MODE & 1) {
        const uint64_t a = tc::smem_desc_sw128(base), q = tc::smem_desc_sw128(base + 16384 + b * 16384);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc::mma_bf16_ss(tmem + b * 128, a + 2 * kk, q + 2 * kk, idesc1, kk > 0);
          tc::mma_commit(&bar[0][b]);
        }
        __syncwarp();
      }
      if (MODE & 4) {
        const uint64_t qb = tc::smem_desc_sw128(base + 114688 + b * 16384);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_bf16_ts(tmem + 256, tmem + 384 + b * 64 + kk * 8, qb + (kk >> 2) * (N2 * 8) + 2 * (kk & 3), idesc2, (it | kk) != 0);
          tc::mma_commit(&bar[1][b]);
        }
        __syncwarp();
      }
      if (MODE & 2) {
        const uint64_t pa = tc::smem_desc_sw128(base + 49152 + b * 32768), qb = tc::smem_desc_sw128(base + 114688 + b * 16384);
        if (leader) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_bf16_ss(tmem + 256, pa + kb * 1024 + 2 * kk, qb + kb * (N2 * 8) + 2 * kk, idesc2, (it | kb | kk) != 0);
          tc::mma_commit(&bar[1][b]);
        }
        __syncwarp();
      }
      // SYNC: wait for this unit's commits; otherwise for the previous unit's (two units in flight)
      const int w = SYNC ? it : it - 1;
      if (w >= 0) {
        if (MODE & 1) tc::mbar_wait(&bar[0][w & 1], (w >> 1) & 1, &abort_flag);
        if (MODE & 6) tc::mbar_wait(&bar[1][w & 1], (w >> 1) & 1, &abort_flag);
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = abort_flag; }
  if (warp == 0) { tc::tcgen05_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int MODE, int SYNC, int N2>
void run(const char* name, long long* d_out) {
  const int iters = 1024;
  cudaFuncSetAttribute(k<MODE, SYNC, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<MODE, SYNC, N2><<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaDeviceSynchronize();
  k<MODE, SYNC, N2><<<148, 128, 200 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"variant\": \"%s\", \"clocks_per_unit\": %.1f, \"abort\": %lld, \"err\": \"%s\"}\n", name, (double)h[0] / iters, h[1], cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaMemset(d_out, 0, 16);
  run<1, 0, 32>("GEMM1 4x(128x128x16), pipelined", d_out);
  run<2, 0, 32>("GEMM2 8x(128x32x16), pipelined", d_out);
  run<2, 0, 64>("GEMM2 8x(128x64x16), pipelined", d_out);
  run<3, 0, 32>("GEMM1+GEMM2, pipelined", d_out);
  run<4, 0, 32>("GEMM2 A-in-TMEM 8x(128x32x16), pipelined", d_out);
  run<5, 0, 32>("GEMM1 + GEMM2 A-in-TMEM, pipelined", d_out);
  run<4, 1, 32>("GEMM2 A-in-TMEM, round trip every unit", d_out);
  run<1, 1, 32>("GEMM1, commit round trip every unit", d_out);
  run<2, 1, 32>("GEMM2, commit round trip every unit", d_out);
  run<3, 1, 32>("GEMM1+GEMM2, round trip every unit", d_out);
  rung<0, 0>("grad GEMM 16x(128x64x16), A K-major, B K-major", d_out);
  rung<0, 1>("grad GEMM 16x(128x64x16), A K-major, B MN-major (dF0)", d_out);
  rung<1, 1>("grad GEMM 16x(128x64x16), A MN-major, B MN-major (dF1)", d_out);
  return 0;
}
