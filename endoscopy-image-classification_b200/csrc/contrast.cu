// K6: CoMatch graph-contrastive loss, exact-fp32 FFMA path (code/comatch.py:199-213
// and its autograd backward, closed form in SURVEY 8a row a7).
//
// Nothing of size [rows, rows] is ever written: each 64x64 tile of
//   S = F0 F1^T (embedding similarity) and Qraw = probs probs^T (pseudo-label graph)
// is recomputed from shared memory where it is needed.
//   stats : rowsum_i = sum_j exp(S_ij/tau), qsum_i = sum_j Qm_ij
//   loss  : loss_i = -sum_j log(P_ij+1e-7) Qn_ij,  r_i = sum_j G_ij P_ij
//   bwd   : dZ = P o (G - r);  dF0 = dZ F1 / tau (row CTAs);  dF1 = dZ^T F0 / tau (column CTAs)
// with P = exp(S/tau)/rowsum, Qm = threshold(diag1(Qraw)), Qn = Qm/qsum,
// G = -Qn/(P+1e-7)/rows.
// Every kernel splits the streamed dimension over gridDim.y CTAs so that the
// reference's sizes (rows = 448 -> 7 tiles) still fill the 148 SMs; the split
// partials are folded by the last CTA of a tile in split order (deterministic).
#include <math.h>

#include "common.cuh"
#include "tiles.cuh"

namespace b200ssl {
namespace {

struct ContrastParams {
  const void* f0; const void* f1; const float* probs;
  long long rows, rows_pad; int D, C; float tau, th;
  float* stats;              // [3][rows]: rowsum, qsum, r
  float* out;
  float* part;               // split partials
  unsigned* tile_tickets;    // per (kernel, tile) tickets
  unsigned* grid_ticket; float* grid_part;
  int nsplit, tiles_per_split;
  const float* upstream; float factor; void* g0; void* g1;
  const float* loss_u; float lambda_u, lambda_c; float* total_out;
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

// fold the 16 column lanes (tx) of a row
__device__ __forceinline__ float fold16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Last-arriving split CTA of a tile returns true (all partials visible).
__device__ __forceinline__ bool tile_last(unsigned* ticket, int nsplit) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)nsplit - 1);
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// PASS 0: row statistics.  PASS 1: per-row loss and r (needs the statistics).
template <typename T, int PASS>
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
  const int split = blockIdx.y;
  const long long jt0 = (long long)split * p.tiles_per_split, jt1 = min(ntiles, jt0 + p.tiles_per_split);
  load_tile_padded(static_cast<const T*>(p.f0) + i0 * D, mrows, kTM, D, ldd, As);
  load_probs_padded(p.probs + i0 * C, mrows, C, Pi);

  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};   // (rs, qs) or (loss, rr)
  float rs[4], qs[4];
  if (PASS == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 16 * i;
      rs[i] = (r < mrows) ? p.stats[i0 + r] : 1.f;
      qs[i] = (r < mrows) ? p.stats[p.rows + i0 + r] : 1.f;
    }
  }
  for (long long jt = jt0; jt < jt1; ++jt) {
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
        if (PASS == 0) {
          a0[i] += e;
          a1[i] += qm;
        } else if (qm != 0.f) {
          const float P = __fdiv_rn(e, rs[i]);                                 // :201
          const float qn = __fdiv_rn(qm, qs[i]);                               // :209
          a0[i] -= logf(P + 1e-7f) * qn;                                       // :212
          a1[i] += qn * __fdiv_rn(P, P + 1e-7f);
        }
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { a0[i] = fold16(a0[i]); a1[i] = fold16(a1[i]); }
  // publish this split's partials: part[(pass*2+w)][split][rows_pad]
  float* base = p.part + (size_t)(PASS * 2) * p.nsplit * p.rows_pad;
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 16 * i;
      base[(size_t)split * p.rows_pad + i0 + r] = a0[i];
      base[(size_t)(p.nsplit + split) * p.rows_pad + i0 + r] = a1[i];
    }
  }
  if (!tile_last(p.tile_tickets + PASS * gridDim.x + blockIdx.x, p.nsplit)) return;
  float lsum[1] = {0.f};
  // the two per-row quantities of this pass, folded over the splits (vectorised, split order)
  __shared__ float s_fold[2][kTM];
#pragma unroll
  for (int w = 0; w < 2; ++w)
    fold_splits_vec4(reinterpret_cast<const float4*>(base + (size_t)w * p.nsplit * p.rows_pad + i0), (size_t)p.rows_pad / 4,
                     p.nsplit, kTM / 4, [&](int i, float4 v) {
                       s_fold[w][4 * i] = v.x; s_fold[w][4 * i + 1] = v.y; s_fold[w][4 * i + 2] = v.z; s_fold[w][4 * i + 3] = v.w;
                     });
  __syncthreads();
  for (int r = tid; r < mrows; r += blockDim.x) {
    const float t0 = s_fold[0][r], t1 = s_fold[1][r];
    if (PASS == 0) {
      p.stats[i0 + r] = t0;
      p.stats[p.rows + i0 + r] = t1;
    } else {
      p.stats[2 * p.rows + i0 + r] = -t1 / (float)p.rows;
      lsum[0] += t0;
    }
  }
  if (tid == 0) p.tile_tickets[PASS * gridDim.x + blockIdx.x] = 0u;
  if (PASS == 0) return;
  // loss: sum over the rows of this tile, then over tiles (last tile-finisher folds, fixed order)
  __shared__ float s_w[kTileThreads / 32];
  const float t = warp_sum(lsum[0]);
  if ((tid & 31) == 0) s_w[tid >> 5] = t;
  __syncthreads();
  __shared__ bool s_glast;
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kTileThreads / 32; ++w) c += s_w[w];
    p.grid_part[blockIdx.x] = c;
    __threadfence();
    s_glast = (atomicAdd(p.grid_ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_glast) return;
  __threadfence();
  float v = 0.f;
  for (unsigned b = tid; b < gridDim.x; b += blockDim.x) v += __ldcg(p.grid_part + b);
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) s_w[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kTileThreads / 32; ++w) c += s_w[w];
    const float lc = c / (float)p.rows;                                          // :213
    p.out[0] = lc;
    if (p.total_out) p.total_out[0] = p.lambda_u * (p.loss_u ? *p.loss_u : 0.f) + p.lambda_c * lc;  // :222
    *p.grid_ticket = 0u;
  }
}

// grid = (tiles, nsplit, 2): z==0 -> dF0 of row tile x ; z==1 -> dF1 of column tile x.
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
  const bool colmode = blockIdx.z == 1;
  const int split = blockIdx.y;
  const long long own0 = (long long)blockIdx.x * kTM;
  const int nown = (int)min((long long)kTM, p.rows - own0);
  const long long ntiles = (p.rows + kTN - 1) / kTN;
  const long long t0 = (long long)split * p.tiles_per_split, t1 = min(ntiles, t0 + p.tiles_per_split);
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

  for (long long t = t0; t < t1; ++t) {
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
  const float up = (p.upstream ? *p.upstream : 1.f) * p.factor;
  T* out = static_cast<T*>(colmode ? p.g1 : p.g0);
  if (p.nsplit == 1) {
    if (orow < nown) {
#pragma unroll
      for (int m = 0; m < ND; ++m) {
        const int d = dg + 4 * m;
        if (d < D) out[(own0 + orow) * D + d] = from_f32<T>(__fdiv_rn(acc[m], p.tau) * up);
      }
    }
    return;
  }
  // split partials [mode][split][rows_pad][D] -> last split CTA of (mode, tile) folds them in order
  float* base = p.part + (size_t)(colmode ? 1 : 0) * p.nsplit * p.rows_pad * D;
#pragma unroll
  for (int m = 0; m < ND; ++m) {
    const int d = dg + 4 * m;
    if (d < D) base[((size_t)split * p.rows_pad + own0 + orow) * D + d] = acc[m];
  }
  unsigned* ticket = p.tile_tickets + (2 + (colmode ? 1 : 0)) * gridDim.x + blockIdx.x;
  if (!tile_last(ticket, p.nsplit)) return;
  fold_splits_vec4(reinterpret_cast<const float4*>(base + (size_t)own0 * D), (size_t)p.rows_pad * D / 4, p.nsplit,
                   nown * D / 4, [&](int i, float4 v) {
                     const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                     for (int j = 0; j < 4; ++j) out[own0 * D + 4 * i + j] = from_f32<T>(__fdiv_rn(vv[j], p.tau) * up);
                   });
  if (tid == 0) *ticket = 0u;
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

// number of CTAs the streamed dimension is split over (shared with api.cu workspace sizing)
int contrast_nsplit(long long rows, int modes, int* tiles_per_split) {
  const long long tiles = (rows + kTM - 1) / kTM;
  long long want = (2 * kNumSMs + tiles * modes - 1) / (tiles * modes);
  if (want < 1) want = 1;
  if (want > tiles) want = tiles;
  const long long tps = (tiles + want - 1) / want;
  if (tiles_per_split) *tiles_per_split = (int)tps;
  return (int)((tiles + tps - 1) / tps);
}

size_t contrast_workspace_floats(long long rows, int dim) {
  const long long tiles = (rows + kTM - 1) / kTM;
  const long long rows_pad = tiles * kTM;
  const size_t fwd = (size_t)4 * contrast_nsplit(rows, 1, nullptr) * rows_pad;
  const int nb = contrast_nsplit(rows, 2, nullptr);
  const size_t bwd = nb > 1 ? (size_t)2 * nb * rows_pad * dim : 0;
  return (size_t)((tiles + 3) & ~3LL) + (fwd > bwd ? fwd : bwd);
}

}  // namespace b200ssl

using namespace b200ssl;

namespace b200ssl {
int contrast_fwd_tc(const void* f0, const void* f1, const void* probs_hl, long long rows, int classes, float temperature,
                    float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                    float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int contrast_bwd_tc(const void* f0, const void* f1, const void* probs_hl, const float* stats, long long rows, int classes,
                    float temperature, float contrast_th, const float* upstream, float factor, void* g0, void* g1,
                    void* scale_grad, long long scale_numel, const float* scale_up, float scale_factor,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);
int contrast_fwd_tc_f32(const float* f0, const float* f1, const float* probs, long long rows, int classes, float temperature,
                        float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                        float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int contrast_bwd_tc_f32(const float* f0, const float* f1, const float* probs, const float* stats, long long rows, int classes,
                        float temperature, float contrast_th, const float* upstream, float factor, float* g0, float* g1,
                        float* scale_grad, long long scale_numel, const float* scale_up, float scale_factor,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern int g_f32_simt;     // bank.cu: fp32 storage on the exact-fp32 FFMA tiles (A/B) instead of the split-operand tensor-core kernels
}
// fp32 storage, 64-wide embeddings: the tensor-core kernels on bf16 hi + mid operands (contrast_tc.cu)
static bool use_tc_f32(const void* f0, const void* f1, int dim, int classes, int dtype, const void* g0 = nullptr, const void* g1 = nullptr) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  return dtype == B200SSL_F32 && dim == 64 && classes <= 32 && !g_f32_simt && al(f0) && al(f1) && al(g0) && al(g1);
}
// bf16 embeddings with 128-byte rows + the hi/lo probability split: tcgen05 path (contrast_tc.cu)
static bool use_tc(const void* f0, const void* f1, const void* probs_hl, int dim, int classes, int dtype, const void* g0 = nullptr,
                   const void* g1 = nullptr) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  return dtype == B200SSL_BF16 && dim == 64 && classes <= 32 && probs_hl && al(f0) && al(f1) && al(probs_hl) && al(g0) && al(g1);
}

static int contrast_setup(const char* fn, ContrastParams& p, int modes, void* workspace, size_t workspace_bytes) {
  const long long tiles = (p.rows + kTM - 1) / kTM;
  p.rows_pad = tiles * kTM;
  p.nsplit = contrast_nsplit(p.rows, modes, &p.tiles_per_split);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(B200SSL_E_ALIGN, "%s: workspace NULL or not 256-byte aligned", fn);
  const size_t need = kWsHeaderBytes + sizeof(float) * contrast_workspace_floats(p.rows, p.D);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  if ((size_t)tiles * 4 * sizeof(unsigned) > kWsTicket2Bytes) return fail(B200SSL_E_SHAPE, "%s: too many row tiles", fn);
  p.tile_tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kWsTicketBytes);
  p.grid_ticket = reinterpret_cast<unsigned*>(workspace) + 4;
  p.grid_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  p.part = p.grid_part + ((tiles + 3) & ~3LL);   // keep the split partials 16-byte aligned (float4 folds)
  return 0;
}

extern "C" int b200ssl_contrast_fwd(const void* feats_s0, const void* feats_s1, const float* probs, const void* probs_hl,
                                    int64_t rows,
                                    int32_t dim, int32_t classes, int32_t dtype, float temperature, float contrast_th,
                                    float* stats, float* out_scalar, const float* loss_u, float lambda_u,
                                    float lambda_c, float* total_out, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  const char* fn = "b200ssl_contrast_fwd";
  if (int e = check_contrast(fn, rows, dim, classes, dtype, temperature)) return e;
  if (!feats_s0 || !feats_s1 || !probs || !stats || !out_scalar) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(B200SSL_E_ALIGN, "%s: workspace NULL or not 256-byte aligned", fn);
  if (use_tc(feats_s0, feats_s1, probs_hl, dim, classes, dtype))
    return contrast_fwd_tc(feats_s0, feats_s1, probs_hl, rows, classes, temperature, contrast_th, stats, out_scalar, loss_u,
                           lambda_u, lambda_c, total_out, workspace, workspace_bytes, as_stream(stream));
  if (use_tc_f32(feats_s0, feats_s1, dim, classes, dtype))
    return contrast_fwd_tc_f32(static_cast<const float*>(feats_s0), static_cast<const float*>(feats_s1), probs, rows, classes, temperature,
                               contrast_th, stats, out_scalar, loss_u, lambda_u, lambda_c, total_out, workspace, workspace_bytes,
                               as_stream(stream));
  ContrastParams p{};
  p.f0 = feats_s0; p.f1 = feats_s1; p.probs = probs; p.rows = rows; p.D = dim; p.C = classes;
  p.tau = temperature; p.th = contrast_th; p.stats = stats; p.out = out_scalar;
  p.loss_u = loss_u; p.lambda_u = lambda_u; p.lambda_c = lambda_c; p.total_out = total_out;
  if (int e = contrast_setup(fn, p, 1, workspace, workspace_bytes)) return e;
  const long long tiles = (rows + kTM - 1) / kTM;
  const size_t smem = ((size_t)(kTM + kTN) * (dim + 1) + (size_t)(kTM + kTN) * (classes + 1)) * sizeof(float);
  dim3 grid((unsigned)tiles, (unsigned)p.nsplit);
  cudaError_t e = cudaSuccess;
#define LAUNCH_FWD(T)                                                                                          \
  do {                                                                                                         \
    auto k0 = contrast_fwd_kernel<T, 0>;                                                                       \
    auto k1 = contrast_fwd_kernel<T, 1>;                                                                       \
    if (smem > 48 * 1024) {                                                                                    \
      e = cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    }                                                                                                          \
    if (e == cudaSuccess) {                                                                                    \
      k0<<<grid, kTileThreads, smem, as_stream(stream)>>>(p);                                                  \
      k1<<<grid, kTileThreads, smem, as_stream(stream)>>>(p);                                                  \
    }                                                                                                          \
  } while (0)
  if (dtype == B200SSL_F32) LAUNCH_FWD(float); else LAUNCH_FWD(__nv_bfloat16);
#undef LAUNCH_FWD
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

extern "C" int b200ssl_contrast_bwd(const void* feats_s0, const void* feats_s1, const float* probs, const void* probs_hl,
                                    const float* stats, int64_t rows, int32_t dim, int32_t classes, int32_t dtype, float temperature,
                                    float contrast_th, const float* upstream, float factor, void* grad_f0, void* grad_f1,
                                    void* scale_grad, int64_t scale_numel, const float* scale_upstream, float scale_factor,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_contrast_bwd";
  if (int e = check_contrast(fn, rows, dim, classes, dtype, temperature)) return e;
  if (!feats_s0 || !feats_s1 || !probs || !stats || !grad_f0 || !grad_f1) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(B200SSL_E_ALIGN, "%s: workspace NULL or not 256-byte aligned", fn);
  if (use_tc(feats_s0, feats_s1, probs_hl, dim, classes, dtype, grad_f0, grad_f1)) {
    const bool piggy = scale_grad && (reinterpret_cast<uintptr_t>(scale_grad) & 15u) == 0;   // vector path needs 16-byte alignment
    if (scale_grad && !piggy)
      if (int e = b200ssl_scale_inplace(scale_grad, scale_numel, dtype, scale_upstream, scale_factor, stream)) return e;
    return contrast_bwd_tc(feats_s0, feats_s1, probs_hl, stats, rows, classes, temperature, contrast_th, upstream, factor,
                           grad_f0, grad_f1, piggy ? scale_grad : nullptr, scale_numel, scale_upstream, scale_factor, workspace,
                           workspace_bytes, as_stream(stream));
  }
  if (use_tc_f32(feats_s0, feats_s1, dim, classes, dtype, grad_f0, grad_f1))
    return contrast_bwd_tc_f32(static_cast<const float*>(feats_s0), static_cast<const float*>(feats_s1), probs, stats, rows, classes,
                               temperature, contrast_th, upstream, factor, static_cast<float*>(grad_f0), static_cast<float*>(grad_f1),
                               static_cast<float*>(scale_grad), scale_numel, scale_upstream, scale_factor, workspace, workspace_bytes,
                               as_stream(stream));
  if (scale_grad)      // exact-fp32 path: the scaling is its own (tiny) launch
    if (int e = b200ssl_scale_inplace(scale_grad, scale_numel, dtype, scale_upstream, scale_factor, stream)) return e;
  ContrastParams p{};
  p.f0 = feats_s0; p.f1 = feats_s1; p.probs = probs; p.rows = rows; p.D = dim; p.C = classes;
  p.tau = temperature; p.th = contrast_th; p.stats = const_cast<float*>(stats);
  p.upstream = upstream; p.factor = factor; p.g0 = grad_f0; p.g1 = grad_f1;
  if (int e = contrast_setup(fn, p, 2, workspace, workspace_bytes)) return e;
  const long long tiles = (rows + kTM - 1) / kTM;
  const size_t smem = ((size_t)(kTM + kTN) * (dim + 1) + (size_t)(kTM + kTN) * (classes + 1) + kTM * (kTN + 1) + 3 * kTM) * sizeof(float);
  dim3 grid((unsigned)tiles, (unsigned)p.nsplit, 2);
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
