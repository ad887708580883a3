// K6: CoMatch graph-contrastive loss, exact-fp32 FFMA path (code/comatch.py:199-213
// and its autograd backward, closed form in SURVEY 8a row a7).
//
// Nothing of size [rows, rows] is ever written: each 64x64 tile of
//   S = F0 F1^T (embedding similarity) and Qraw = probs probs^T (pseudo-label graph)
// is recomputed from shared memory where it is needed.
//   fwd : pass 1 -> rowsum_i = sum_j exp(S_ij/tau), qsum_i = sum_j Qm_ij
//         pass 2 -> loss_i = -sum_j log(P_ij+1e-7) Qn_ij,  r_i = sum_j G_ij P_ij
//   bwd : dZ = P o (G - r);  dF0 = dZ F1 / tau (row CTAs);  dF1 = dZ^T F0 / tau (column CTAs)
// with P = exp(S/tau)/rowsum, Qm = threshold(diag1(Qraw)), Qn = Qm/qsum,
// G = -Qn/(P+1e-7)/rows.
#include <math.h>

#include "common.cuh"
#include "tiles.cuh"

namespace b200ssl {
namespace {

struct ContrastParams {
  const void* f0; const void* f1; const float* probs;
  long long rows; int D, C; float tau, th;
  float* stats;              // [3][rows]: rowsum, qsum, r
  float* out; float* partials; unsigned* ticket;
  const float* upstream; void* g0; void* g1;
};

__device__ __forceinline__ void load_probs_padded(const float* __restrict__ g, int rows_valid, int C, float* __restrict__ s) {
  const int ld = C + 1;
  for (int e = threadIdx.x; e < kTM * C; e += blockDim.x) {
    const int r = e / C, c = e - r * C;
    s[r * ld + c] = (r < rows_valid) ? g[(size_t)r * C + c] : 0.f;
  }
}

// Qm for one element: diagonal forced to 1, then thresholded (comatch.py:205-208).
__device__ __forceinline__ float graph_weight(float qraw, bool diag, float th) {
  const float q = diag ? 1.f : qraw;
  return (q >= th) ? q : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(kTileThreads) contrast_fwd_kernel(const ContrastParams p) {
  extern __shared__ float smem[];
  const int D = p.D, C = p.C, ldd = D + 1, ldc = C + 1;
  float* As = smem;
  float* Bs = As + kTM * ldd;
  float* Pi = Bs + kTN * ldd;
  float* Pj = Pi + kTM * ldc;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long i0 = (long long)blockIdx.x * kTM;
  const int mrows = (int)min((long long)kTM, p.rows - i0);
  const long long ntiles = (p.rows + kTN - 1) / kTN;
  load_tile_padded(static_cast<const T*>(p.f0) + i0 * D, mrows, kTM, D, ldd, As);
  load_probs_padded(p.probs + i0 * C, mrows, C, Pi);

  float rs[4] = {0.f, 0.f, 0.f, 0.f}, qs[4] = {0.f, 0.f, 0.f, 0.f};
  float li[4] = {0.f, 0.f, 0.f, 0.f}, rr[4] = {0.f, 0.f, 0.f, 0.f};
  const float inv_rows = 1.0f / (float)p.rows;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    for (long long jt = 0; jt < ntiles; ++jt) {
      const long long j0 = jt * kTN;
      const int ncols = (int)min((long long)kTN, p.rows - j0);
      __syncthreads();
      load_tile_padded(static_cast<const T*>(p.f1) + j0 * D, ncols, kTN, D, ldd, Bs);
      load_probs_padded(p.probs + j0 * C, ncols, C, Pj);
      __syncthreads();
      float s[4][4], q[4][4];
      tile_dot_4x4(As, Bs, ldd, D, ty, tx, s);
      tile_dot_4x4(Pi, Pj, ldc, C, ty, tx, q);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = tx + 16 * j;
          if (col >= ncols) continue;
          const float e = expf(__fdiv_rn(s[i][j], p.tau));                       // :200
          const float qm = graph_weight(q[i][j], (i0 + ty + 16 * i) == (j0 + col), p.th);
          if (pass == 0) {
            rs[i] += e;
            qs[i] += qm;
          } else if (qm != 0.f) {
            const float P = __fdiv_rn(e, rs[i]);                                 // :201
            const float qn = __fdiv_rn(qm, qs[i]);                               // :209
            li[i] -= logf(P + 1e-7f) * qn;                                       // :212
            rr[i] += qn * __fdiv_rn(P, P + 1e-7f);
          }
        }
    }
    if (pass == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
          qs[i] += __shfl_xor_sync(0xffffffffu, qs[i], o);
        }
    }
  }
  float acc[1] = {0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      li[i] += __shfl_xor_sync(0xffffffffu, li[i], o);
      rr[i] += __shfl_xor_sync(0xffffffffu, rr[i], o);
    }
    const int r = ty + 16 * i;
    if (tx == 0 && r < mrows) {
      p.stats[i0 + r] = rs[i];
      p.stats[p.rows + i0 + r] = qs[i];
      p.stats[2 * p.rows + i0 + r] = -rr[i] * inv_rows;
      acc[0] += li[i];
    }
  }
  // CTA sum -> grid sum (deterministic)
  __shared__ float s_w[kTileThreads / 32];
  const float t = warp_sum(acc[0]);
  if ((tid & 31) == 0) s_w[tid >> 5] = t;
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kTileThreads / 32; ++w) c += s_w[w];
    acc[0] = c;
  }
  float total[1];
  if (grid_reduce_last<1>(acc, p.partials, p.ticket, total) && tid == 0) p.out[0] = total[0] / (float)p.rows;  // :213
}

// grid = (tiles, 2): y==0 -> dF0 of row tile x ; y==1 -> dF1 of column tile x.
template <typename T, int ND>
__global__ void __launch_bounds__(kTileThreads) contrast_bwd_kernel(const ContrastParams p) {
  extern __shared__ float smem[];
  const int D = p.D, C = p.C, ldd = D + 1, ldc = C + 1;
  float* As = smem;                    // F0 tile (rows i)
  float* Bs = As + kTM * ldd;          // F1 tile (rows j)
  float* Pi = Bs + kTN * ldd;
  float* Pj = Pi + kTM * ldc;
  float* Zs = Pj + kTN * ldc;          // dZ tile [64][65]
  float* St = Zs + kTM * (kTN + 1);    // stats of the i rows: [3][64]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const bool colmode = blockIdx.y == 1;
  const long long own0 = (long long)blockIdx.x * kTM;
  const int nown = (int)min((long long)kTM, p.rows - own0);
  const long long ntiles = (p.rows + kTN - 1) / kTN;
  const float inv_rows = 1.0f / (float)p.rows;
  if (!colmode) {
    load_tile_padded(static_cast<const T*>(p.f0) + own0 * D, nown, kTM, D, ldd, As);
    load_probs_padded(p.probs + own0 * C, nown, C, Pi);
    for (int e = tid; e < 3 * kTM; e += blockDim.x) {
      const int w = e / kTM, r = e - w * kTM;
      St[e] = (r < nown) ? p.stats[w * p.rows + own0 + r] : 1.f;
    }
  } else {
    load_tile_padded(static_cast<const T*>(p.f1) + own0 * D, nown, kTN, D, ldd, Bs);
    load_probs_padded(p.probs + own0 * C, nown, C, Pj);
  }
  float acc[ND];
#pragma unroll
  for (int m = 0; m < ND; ++m) acc[m] = 0.f;
  const int orow = tid & 63, dg = tid >> 6;

  for (long long t = 0; t < ntiles; ++t) {
    const long long o0 = t * kTN;
    const int nother = (int)min((long long)kTN, p.rows - o0);
    __syncthreads();
    if (!colmode) {
      load_tile_padded(static_cast<const T*>(p.f1) + o0 * D, nother, kTN, D, ldd, Bs);
      load_probs_padded(p.probs + o0 * C, nother, C, Pj);
    } else {
      load_tile_padded(static_cast<const T*>(p.f0) + o0 * D, nother, kTM, D, ldd, As);
      load_probs_padded(p.probs + o0 * C, nother, C, Pi);
      for (int e = tid; e < 3 * kTM; e += blockDim.x) {
        const int w = e / kTM, r = e - w * kTM;
        St[e] = (r < nother) ? p.stats[w * p.rows + o0 + r] : 1.f;
      }
    }
    __syncthreads();
    const long long i0 = colmode ? o0 : own0, j0 = colmode ? own0 : o0;
    const int ni = colmode ? nother : nown, nj = colmode ? nown : nother;
    float s[4][4], q[4][4];
    tile_dot_4x4(As, Bs, ldd, D, ty, tx, s);
    tile_dot_4x4(Pi, Pj, ldc, C, ty, tx, q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ri = ty + 16 * i;
      const float rsum = St[ri], qsum = St[kTM + ri], r_i = St[2 * kTM + ri];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = tx + 16 * j;
        float dz = 0.f;
        if (ri < ni && col < nj) {
          const float P = __fdiv_rn(expf(__fdiv_rn(s[i][j], p.tau)), rsum);
          const float qm = graph_weight(q[i][j], (i0 + ri) == (j0 + col), p.th);
          const float G = (qm != 0.f) ? -__fdiv_rn(__fdiv_rn(qm, qsum), P + 1e-7f) * inv_rows : 0.f;
          dz = P * (G - r_i);
        }
        Zs[ri * (kTN + 1) + col] = dz;
      }
    }
    __syncthreads();
    if (!colmode) {
      for (int k = 0; k < kTN; ++k) {
        const float z = Zs[orow * (kTN + 1) + k];
#pragma unroll
        for (int m = 0; m < ND; ++m) {
          const int d = dg + 4 * m;
          if (d < D) acc[m] = fmaf(z, Bs[k * ldd + d], acc[m]);
        }
      }
    } else {
      for (int k = 0; k < kTM; ++k) {
        const float z = Zs[k * (kTN + 1) + orow];
#pragma unroll
        for (int m = 0; m < ND; ++m) {
          const int d = dg + 4 * m;
          if (d < D) acc[m] = fmaf(z, As[k * ldd + d], acc[m]);
        }
      }
    }
  }
  const float up = p.upstream ? *p.upstream : 1.f;
  T* out = static_cast<T*>(colmode ? p.g1 : p.g0);
  if (orow < nown) {
#pragma unroll
    for (int m = 0; m < ND; ++m) {
      const int d = dg + 4 * m;
      if (d < D) out[(own0 + orow) * D + d] = from_f32<T>(__fdiv_rn(acc[m], p.tau) * up);
    }
  }
}

int check_contrast(const char* fn, long long rows, int D, int C, int dtype, float tau) {
  if (rows <= 0) return fail(B200SSL_E_SHAPE, "%s: rows must be > 0", fn);
  if (D < 8 || D > B200SSL_MAX_EMB_DIM || D % 8) return fail(B200SSL_E_SHAPE, "%s: dim %d must be a multiple of 8 in [8,%d]", fn, D, B200SSL_MAX_EMB_DIM);
  if (C < 2 || C > 128) return fail(B200SSL_E_SHAPE, "%s: classes %d outside [2,128]", fn, C);
  if (dtype != B200SSL_F32 && dtype != B200SSL_BF16) return fail(B200SSL_E_DTYPE, "%s: dtype %d", fn, dtype);
  if (!(tau > 0.f)) return fail(B200SSL_E_ARG, "%s: temperature must be > 0", fn);
  return 0;
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_contrast_fwd(const void* feats_s0, const void* feats_s1, const float* probs, int64_t rows,
                                    int32_t dim, int32_t classes, int32_t dtype, float temperature, float contrast_th,
                                    float* stats, float* out_scalar, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  const char* fn = "b200ssl_contrast_fwd";
  if (int e = check_contrast(fn, rows, dim, classes, dtype, temperature)) return e;
  if (!feats_s0 || !feats_s1 || !probs || !stats || !out_scalar) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  const long long tiles = (rows + kTM - 1) / kTM;
  const size_t need = kWsHeaderBytes + sizeof(float) * (size_t)tiles;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(B200SSL_E_ALIGN, "%s: workspace NULL or not 256-byte aligned", fn);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  ContrastParams p{};
  p.f0 = feats_s0; p.f1 = feats_s1; p.probs = probs; p.rows = rows; p.D = dim; p.C = classes;
  p.tau = temperature; p.th = contrast_th; p.stats = stats; p.out = out_scalar;
  p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  p.ticket = reinterpret_cast<unsigned*>(workspace) + 4;
  const size_t smem = ((size_t)(kTM + kTN) * (dim + 1) + (size_t)(kTM + kTN) * (classes + 1)) * sizeof(float);
  cudaError_t e = cudaSuccess;
  if (dtype == B200SSL_F32) {
    auto k = contrast_fwd_kernel<float>;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k<<<(unsigned)tiles, kTileThreads, smem, as_stream(stream)>>>(p);
  } else {
    auto k = contrast_fwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k<<<(unsigned)tiles, kTileThreads, smem, as_stream(stream)>>>(p);
  }
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

extern "C" int b200ssl_contrast_bwd(const void* feats_s0, const void* feats_s1, const float* probs, const float* stats,
                                    int64_t rows, int32_t dim, int32_t classes, int32_t dtype, float temperature,
                                    float contrast_th, const float* upstream, void* grad_f0, void* grad_f1,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_contrast_bwd";
  (void)workspace; (void)workspace_bytes;
  if (int e = check_contrast(fn, rows, dim, classes, dtype, temperature)) return e;
  if (!feats_s0 || !feats_s1 || !probs || !stats || !grad_f0 || !grad_f1) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  ContrastParams p{};
  p.f0 = feats_s0; p.f1 = feats_s1; p.probs = probs; p.rows = rows; p.D = dim; p.C = classes;
  p.tau = temperature; p.th = contrast_th; p.stats = const_cast<float*>(stats);
  p.upstream = upstream; p.g0 = grad_f0; p.g1 = grad_f1;
  const long long tiles = (rows + kTM - 1) / kTM;
  const size_t smem = ((size_t)(kTM + kTN) * (dim + 1) + (size_t)(kTM + kTN) * (classes + 1) + kTM * (kTN + 1) + 3 * kTM) * sizeof(float);
  dim3 grid((unsigned)tiles, 2);
  cudaError_t e = cudaSuccess;
#define LAUNCH_BWD(T, ND)                                                                                    \
  do {                                                                                                       \
    auto k = contrast_bwd_kernel<T, ND>;                                                                     \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) k<<<grid, kTileThreads, smem, as_stream(stream)>>>(p);                              \
  } while (0)
  if (dtype == B200SSL_F32) {
    if (dim <= 64) LAUNCH_BWD(float, 16); else if (dim <= 128) LAUNCH_BWD(float, 32); else LAUNCH_BWD(float, 64);
  } else {
    if (dim <= 64) LAUNCH_BWD(__nv_bfloat16, 16); else if (dim <= 128) LAUNCH_BWD(__nv_bfloat16, 32); else LAUNCH_BWD(__nv_bfloat16, 64);
  }
#undef LAUNCH_BWD
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}
