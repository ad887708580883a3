// 64x64 FFMA tile primitives shared by the exact-fp32 similarity kernels
// (bank.cu K3, contrast.cu K6).  256 threads; thread (ty, tx) = (tid>>4, tid&15)
// owns the 4x4 outputs (ty+16i, tx+16j): with rows padded to an odd stride the
// operand reads are bank-conflict free (A rows broadcast across the 16 tx lanes,
// B rows are 16 consecutive rows at stride D+1).
#pragma once
#include "common.cuh"

namespace b200ssl {

constexpr int kTM = 64;
constexpr int kTN = 64;
constexpr int kTileThreads = 256;

// rows_valid x D contiguous elements -> smem[rows_total][ld] as fp32, zero padded.
// D must be a multiple of 8 so 16-byte vectors never straddle a row.
template <typename T>
__device__ __forceinline__ void load_tile_padded(const T* __restrict__ g, int rows_valid, int rows_total, int D, int ld,
                                                 float* __restrict__ s) {
  constexpr int N = Vec16<T>::N;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nvec_row = D / N;
  const int nvec = rows_total * nvec_row;
  const bool aligned = (reinterpret_cast<uintptr_t>(g) & 15u) == 0;
  for (int v = tid; v < nvec; v += nt) {
    const int r = v / nvec_row, cv = v - r * nvec_row;
    float f[N];
    if (r < rows_valid) {
      if (aligned) {
        unpack16(ldg128(g + (size_t)r * D + cv * N), f, T());
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = to_f32<T>(g[(size_t)r * D + cv * N + i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) s[r * ld + cv * N + i] = f[i];
  }
}

// count_valid contiguous elements -> smem (dense), zero padded up to count_total.
template <typename T>
__device__ __forceinline__ void load_tile_dense(const T* __restrict__ g, int count_valid, int count_total,
                                                float* __restrict__ s) {
  tile_g2s(g, s, count_valid);
  for (int i = count_valid + threadIdx.x; i < count_total; i += blockDim.x) s[i] = 0.f;
}

// acc[i][j] = <A[ty+16i, :], B[tx+16j, :]> over K (fp32 FFMA, ascending k).
__device__ __forceinline__ void tile_dot_4x4(const float* __restrict__ As, const float* __restrict__ Bs, int ld, int K,
                                             int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[(ty + 16 * i) * ld + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[(tx + 16 * j) * ld + k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

}  // namespace b200ssl
