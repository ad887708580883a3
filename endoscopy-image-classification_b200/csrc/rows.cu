// Row kernels of the SSL head: every one of them reads [rows, classes] logits
// with 128-bit loads into shared memory, gives LPR lanes to a row, reduces with
// warp shuffles and emits forward scalars AND the backward gradient in the same
// launch.  HBM-bound by construction (each input read once, each output written
// once); at the reference's sizes (rows <= 14336, 23 classes) they are launch-
// latency floored -- see DESIGN.md.
//
//   K1  fixmatch_head_kernel   code/loss.py:126-164 (+ F.cross_entropy :119, bwd)
//   f2  labeled_ce_kernel      code/loss.py:103-119, 308-364 (+ bwd)
//   K2  comatch_da_kernel      code/comatch.py:167-173
//   K4/K7 comatch_finalize_kernel  code/comatch.py:174-176,182,184-185,216-220 (+ bwd)
#include <cooperative_groups.h>
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace b200ssl {
namespace {

constexpr int kRowThreads = 128;
constexpr int kRowWarps = kRowThreads / 32;

template <int LPR, int EPL>
struct RowCfg {
  static constexpr int kRowsPerWarp = 32 / LPR;
  static constexpr int kRowsPerTile = kRowWarps * kRowsPerWarp;
};

// ---- per-row primitives (LPR lanes cooperate, lane gl owns c = gl + k*LPR) ----
template <int LPR, int EPL>
__device__ __forceinline__ void row_load(const float* srow, int C, int gl, bool valid, float (&x)[EPL]) {
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int c = gl + k * LPR;
    x[k] = (valid && c < C) ? srow[c] : -INFINITY;
  }
}

// FAST = false: IEEE division / accurate expf, logf (fp32 storage: bit-level parity with the eager reference).
// FAST = true : MUFU approximations (bf16 storage: the inputs carry 8 mantissa bits, tolerance 1e-2).
template <bool FAST> __device__ __forceinline__ float fdiv(float a, float b) { return FAST ? __fdividef(a, b) : __fdiv_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fexp(float a) { return FAST ? __expf(a) : expf(a); }
template <bool FAST> __device__ __forceinline__ float flog(float a) { return FAST ? __logf(a) : logf(a); }

// e[k] = expf(x[k]-max) (0 for padding), returns max and sum over the row.
template <int LPR, int EPL, bool FAST = false>
__device__ __forceinline__ void row_softmax_stats(const float (&x)[EPL], float (&e)[EPL], float& mx, float& sum) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < EPL; ++k) m = fmaxf(m, x[k]);
  m = group_max<LPR>(m);
  if (m == -INFINITY) m = 0.f;  // fully padded (invalid) row
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    e[k] = (x[k] == -INFINITY) ? 0.f : fexp<FAST>(x[k] - m);
    s += e[k];
  }
  sum = group_sum<LPR>(s);
  mx = m;
}

template <int LPR, int EPL>
__device__ __forceinline__ void row_argmax(const float (&p)[EPL], int C, int gl, float& best, int& besti) {
  float bv = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int c = gl + k * LPR;
    if (c < C && p[k] > bv) { bv = p[k]; bi = c; }  // ascending c: strict > keeps the first
  }
  group_argmax<LPR>(bv, bi);
  best = bv;
  besti = bi;
}

// value of column `col` of a row held in registers (exact: one lane contributes).
// Used instead of re-reading shared memory that sibling lanes may already be
// overwriting with gradients.
template <int LPR, int EPL>
__device__ __forceinline__ float row_pick(const float (&x)[EPL], int gl, int col) {
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k)
    if (gl + k * LPR == col) v = x[k];
  return group_sum<LPR>(v);
}

// block-level sum of NV per-thread values -> thread 0
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV]) {
  __shared__ float s_part[NV][kRowWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float t = warp_sum(v[i]);
    if (lane == 0) s_part[i][warp] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t = 0.f;
      for (int w = 0; w < kRowWarps; ++w) t += s_part[i][w];
      v[i] = t;
    }
  }
}

// ============================================================ K1 ============
struct HeadParams {
  const void* w; const void* s; const void* s2;
  void* gs; void* gs2;
  long long rows; int C;
  float thr, inv_T; int hard;
  float* out; long long* idx; float* mask;
  float* partials; unsigned* ticket;
};

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) fixmatch_head_kernel(const HeadParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sw = smem;
  float* ss = sw + ROWS * C;
  float* ss2 = ss + ROWS * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const bool dual = p.s2 != nullptr;
  const float inv_rows = 1.0f / (float)p.rows;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  float acc[3] = {0.f, 0.f, 0.f};  // loss, mask count, loss2

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    const int cnt = nrows * C;
    tile_g2s(static_cast<const T*>(p.w) + row0 * C, sw, cnt);
    tile_g2s(static_cast<const T*>(p.s) + row0 * C, ss, cnt);
    if (dual) tile_g2s(static_cast<const T*>(p.s2) + row0 * C, ss2, cnt);
    __syncthreads();

    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    float x[EPL], e[EPL], pw[EPL];
    float mx, sum;
    // ---- weak view: pseudo label (loss.py:151-154) ----
    row_load<LPR, EPL>(sw + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL>(x, e, mx, sum);
#pragma unroll
    for (int k = 0; k < EPL; ++k) pw[k] = __fdiv_rn(e[k], sum);
    float pmax; int idx;
    row_argmax<LPR, EPL>(pw, C, gl, pmax, idx);
    const float m = (valid && pmax >= p.thr) ? 1.f : 0.f;
    float tsum = 1.f;
    if (!p.hard) {  // soft targets softmax(w/T) (extension, quirk Q4)
#pragma unroll
      for (int k = 0; k < EPL; ++k) x[k] = x[k] * p.inv_T;
      row_softmax_stats<LPR, EPL>(x, e, mx, sum);
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < EPL; ++k) { pw[k] = __fdiv_rn(e[k], sum); t += pw[k]; }
      tsum = group_sum<LPR>(t);
    }
    if (valid && gl == 0) {
      if (p.idx) p.idx[row0 + r] = idx;
      if (p.mask) p.mask[row0 + r] = m;
      acc[1] += m;
    }
    const float g = m * inv_rows;
    // ---- strong view(s): masked CE + gradient (loss.py:157-164, :119) ----
#pragma unroll 1
    for (int head = 0; head < (dual ? 2 : 1); ++head) {
      float* srow = (head ? ss2 : ss) + r * C;
      row_load<LPR, EPL>(srow, C, gl, valid, x);
      row_softmax_stats<LPR, EPL>(x, e, mx, sum);
      const float logsum = logf(sum);
      float lrow;
      if (p.hard) {
        const float sidx = row_pick<LPR, EPL>(x, gl, valid ? idx : -1);
        lrow = -((sidx - mx) - logsum);
      } else {
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < EPL; ++k)
          if (gl + k * LPR < C && valid) d += -pw[k] * ((x[k] - mx) - logsum);
        lrow = group_sum<LPR>(d);
      }
      if (valid && gl == 0) acc[head ? 2 : 0] += lrow * m;
      if (valid) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int c = gl + k * LPR;
          if (c < C) {
            const float sm = __fdiv_rn(e[k], sum);
            float gr;
            if (p.hard) gr = (c == idx) ? __fadd_rn(__fmul_rn(sm, g), -g) : __fmul_rn(sm, g);
            else gr = g * (sm * tsum - pw[k]);
            srow[c] = gr;
          }
        }
      }
    }
    __syncthreads();
    tile_s2g(ss, static_cast<T*>(p.gs) + row0 * C, cnt);
    if (dual) tile_s2g(ss2, static_cast<T*>(p.gs2) + row0 * C, cnt);
    __syncthreads();
  }
  block_sum<3>(acc);
  float total[3];
  if (grid_reduce_last<3>(acc, p.partials, p.ticket, total) && threadIdx.x == 0) {
    const float n = (float)p.rows;
    p.out[0] = total[0] / n;
    p.out[1] = total[1] / n;
    p.out[2] = total[2] / n;
  }
}

// ============================================================ f2 ============
// Targets: F.cross_entropy's ignore_index (-100) drops a row (zero loss, zero gradient, not counted in the mean);
// any other label outside [0, C) is a caller bug: the row is dropped too and the sticky counter `bad_labels`
// (workspace slot kWsBadLabelSlot, read by b200ssl_bad_label_count) is raised instead of reading out of bounds.
constexpr int kIgnoreIndex = -100;
constexpr int kWsBadLabelSlot = 60;
__device__ __forceinline__ int checked_label(long long y, int C, unsigned* bad) {
  if (y >= 0 && y < C) return (int)y;
  if (y != kIgnoreIndex) atomicAdd(bad, 1u);
  return -1;
}

struct LabeledParams {
  const void* x; const long long* y; const float* cw; void* gx;
  long long rows; int C; int poly; float eps;
  float* out; float* partials; unsigned* ticket; unsigned* bad_labels;
};

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) labeled_ce_kernel(const LabeledParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sx = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  // normaliser: rows (poly / unweighted) or sum of w[y] (F.cross_entropy weighted mean)
  __shared__ float s_norm;
  if (!p.poly) {
    float v[1] = {0.f};
    for (long long i = threadIdx.x; i < p.rows; i += blockDim.x) {
      const long long yi = p.y[i];
      if (yi >= 0 && yi < C) v[0] += p.cw ? p.cw[yi] : 1.f;
    }
    block_sum<1>(v);
    if (threadIdx.x == 0) s_norm = v[0];
  } else if (threadIdx.x == 0) {
    s_norm = (float)p.rows;
  }
  __syncthreads();
  const float inv_norm = 1.0f / s_norm;
  float acc[1] = {0.f};
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    const int cnt = nrows * C;
    tile_g2s(static_cast<const T*>(p.x) + row0 * C, sx, cnt);
    __syncthreads();
    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    float* srow = sx + r * C;
    float x[EPL], e[EPL];
    float mx, sum;
    row_load<LPR, EPL>(srow, C, gl, valid, x);
    row_softmax_stats<LPR, EPL>(x, e, mx, sum);
    int y = -1;
    if (valid && gl == 0) y = checked_label(p.y[row0 + r], C, p.bad_labels);
    y = __shfl_sync(0xffffffffu, y, rw * LPR);                 // lane 0 of the row group validated (and counted) it
    const bool live = valid && y >= 0;
    const float wy = (live && p.cw) ? p.cw[y] : 1.f;
    const float xy = row_pick<LPR, EPL>(x, gl, live ? y : -1);
    const float ce = -((xy - mx) - logf(sum));
    const float pt = __fdiv_rn(expf(xy - mx), sum);
    const float rowl = p.poly ? (wy * ce + p.eps * (1.f - pt)) : wy * ce;
    if (live && gl == 0) acc[0] += rowl;
    const float coef = live ? (p.poly ? (wy + p.eps * pt) : wy) * inv_norm : 0.f;
    if (valid) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        if (c < C) srow[c] = coef * (__fdiv_rn(e[k], sum) - (c == y ? 1.f : 0.f));
      }
    }
    __syncthreads();
    tile_s2g(sx, static_cast<T*>(p.gx) + row0 * C, cnt);
    __syncthreads();
  }
  block_sum<1>(acc);
  float total[1];
  if (grid_reduce_last<1>(acc, p.partials, p.ticket, total) && threadIdx.x == 0) p.out[0] = total[0] / s_norm;
}


// ---- f2, un-reduced: per-row losses (reduction='none' / 'sum', soft targets) ------------------------------------------
// code/loss.py:118-124 + PolyLoss reduction='none' (:357-359): loss_rows[i] and the gradient of loss_rows[i] w.r.t. its
// own logits row (unit upstream); autograd's per-row upstream gradient is applied by scale_rows_kernel.
struct RowCeParams {
  const void* x; const long long* y; const float* soft; const float* cw; void* gx; float* loss_rows;
  long long rows; int C; int poly; float eps; unsigned* bad_labels;
};

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) ce_rows_kernel(const RowCeParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sx = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    const int cnt = nrows * C;
    tile_g2s(static_cast<const T*>(p.x) + row0 * C, sx, cnt);
    __syncthreads();
    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    float* srow = sx + r * C;
    float x[EPL], e[EPL];
    float mx, sum;
    row_load<LPR, EPL>(srow, C, gl, valid, x);
    row_softmax_stats<LPR, EPL>(x, e, mx, sum);
    const float logsum = logf(sum);
    float rowl;
    if (p.soft) {                                            // loss.py:120-124: sum_c -t_c log_softmax(x)_c
      float t[EPL], d = 0.f, ts = 0.f;
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        t[k] = (valid && c < C) ? p.soft[(row0 + r) * C + c] : 0.f;
        if (valid && c < C) { d += -t[k] * ((x[k] - mx) - logsum); ts += t[k]; }
      }
      rowl = group_sum<LPR>(d);
      const float tsum = group_sum<LPR>(ts);
      if (valid) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int c = gl + k * LPR;
          if (c < C) srow[c] = __fdiv_rn(e[k], sum) * tsum - t[k];
        }
      }
    } else {
      int y = -1;
      if (valid && gl == 0) y = checked_label(p.y[row0 + r], C, p.bad_labels);
      y = __shfl_sync(0xffffffffu, y, rw * LPR);
      const bool live = valid && y >= 0;
      const float wy = (live && p.cw) ? p.cw[y] : 1.f;
      const float xy = row_pick<LPR, EPL>(x, gl, live ? y : -1);
      const float ce = -((xy - mx) - logsum);
      const float pt = __fdiv_rn(expf(xy - mx), sum);
      rowl = live ? (p.poly ? (wy * ce + p.eps * (1.f - pt)) : wy * ce) : 0.f;
      const float coef = live ? (p.poly ? (wy + p.eps * pt) : wy) : 0.f;
      if (valid) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int c = gl + k * LPR;
          if (c < C) srow[c] = coef * (__fdiv_rn(e[k], sum) - (c == y ? 1.f : 0.f));
        }
      }
    }
    if (valid && gl == 0) p.loss_rows[row0 + r] = rowl;
    __syncthreads();
    tile_s2g(sx, static_cast<T*>(p.gx) + row0 * C, cnt);
    __syncthreads();
  }
}

// out[i, :] = in[i, :] * row_scale[i]   (out-of-place: the stash survives a second backward)
template <typename T>
__global__ void scale_rows_kernel(const T* in, T* out, long long rows, int C, const float* row_scale) {
  const long long n = rows * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = from_f32<T>(to_f32<T>(in[i]) * row_scale[i / C]);
}

// ============================================================ f4 ============
// Evaluation head (code/fixmatch.py:148-168 per batch): mean cross-entropy of the batch (F.cross_entropy 'mean'),
// softmax -> arg-max (first index on ties, over the probabilities like np.argmax of F.softmax) and the confusion
// matrix [target, prediction] accumulated with integer atomics -- every metric of utils.calculate_metrics
// (code/utils.py:38-55) is a function of that matrix, so an evaluation pass ends with ONE device-to-host copy.
struct EvalParams {
  const void* x; const long long* y; long long rows; int C;
  unsigned long long* confusion; float* loss_out; long long* pred;
  float* partials; unsigned* ticket; unsigned* bad_labels;
};

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) eval_head_kernel(const EvalParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sx = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  float acc[2] = {0.f, 0.f};                                  // sum of CE, rows counted
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    tile_g2s(static_cast<const T*>(p.x) + row0 * C, sx, nrows * C);
    __syncthreads();
    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    float x[EPL], e[EPL], pr[EPL];
    float mx, sum;
    row_load<LPR, EPL>(sx + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL>(x, e, mx, sum);
#pragma unroll
    for (int k = 0; k < EPL; ++k) pr[k] = __fdiv_rn(e[k], sum);
    float pmax; int idx;
    row_argmax<LPR, EPL>(pr, C, gl, pmax, idx);               // fixmatch.py:160,166
    int y = -1;
    if (valid && gl == 0) y = checked_label(p.y[row0 + r], C, p.bad_labels);
    y = __shfl_sync(0xffffffffu, y, rw * LPR);
    const bool live = valid && y >= 0;
    const float xy = row_pick<LPR, EPL>(x, gl, live ? y : -1);
    const float ce = -((xy - mx) - logf(sum));                // fixmatch.py:156
    if (valid && gl == 0) {
      if (p.pred) p.pred[row0 + r] = idx;
      if (live) {
        acc[0] += ce;
        acc[1] += 1.f;
        atomicAdd(&p.confusion[(size_t)y * C + idx], 1ull);
      }
    }
    __syncthreads();
  }
  block_sum<2>(acc);
  float total[2];
  if (grid_reduce_last<2>(acc, p.partials, p.ticket, total) && threadIdx.x == 0) p.loss_out[0] = total[0] / total[1];
}

// ============================================================ K2 ============
struct DaParams {
  const void* w; long long rows; int C;
  float* ring; int* state; int window; float* prob_avg; float* col_mean;
  float* partials; unsigned* ticket;
};

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) comatch_da_kernel(const DaParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sw = smem;             // [ROWS*C]
  float* scol = sw + ROWS * C;  // [C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  for (int c = threadIdx.x; c < C; c += blockDim.x) scol[c] = 0.f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    tile_g2s(static_cast<const T*>(p.w) + row0 * C, sw, nrows * C);
    __syncthreads();
    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    float x[EPL], e[EPL];
    float mx, sum;
    row_load<LPR, EPL>(sw + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL>(x, e, mx, sum);
    if (valid) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        if (c < C) sw[r * C + c] = __fdiv_rn(e[k], sum);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float t = scol[c];
      for (int rr = 0; rr < nrows; ++rr) t += sw[rr * C + c];
      scol[c] = t;
    }
    __syncthreads();
  }
  // publish per-CTA column sums, last CTA folds them and updates the history
  __shared__ bool s_last;
  for (int c = threadIdx.x; c < C; c += blockDim.x) p.partials[(size_t)blockIdx.x * C + c] = scol[c];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int count_old = p.state[0], head = p.state[1];
  const int count = min(count_old + 1, p.window);
  const int head_new = (head + 1) % p.window;
  // All loads of the fold are issued up front (one thread per (class, lane) pair would
  // serialise dependent round trips per class): thread c owns class c.
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (unsigned b0 = 0; b0 < gridDim.x; b0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (b0 + u < gridDim.x) ? __ldcg(&p.partials[(size_t)(b0 + u) * C + c]) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) t += v[u];
    }
    const float mean = t / (float)p.rows;  // probs.mean(0), comatch.py:169
    // history mean, oldest -> newest (comatch.py:172); the newest entry is `mean`
    float h = 0.f;
    for (int a0 = 0; a0 < count - 1; a0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int a = a0 + u;
        const int slot = (head_new - count + a + 2 * p.window) % p.window;
        v[u] = (a < count - 1) ? p.ring[(size_t)slot * C + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) h += v[u];
    }
    h += mean;
    p.ring[(size_t)head * C + c] = mean;
    p.prob_avg[c] = h / (float)count;
    if (p.col_mean) p.col_mean[c] = mean;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.state[0] = count;
    p.state[1] = head_new;
    *p.ticket = 0u;
  }
}

// ====================================================== K2b + K4 + K7 =======
struct FinalizeParams {
  const void* w; const void* s0;
  const float* prob_avg; const float* rowsum; const float* numer;
  long long rows; int C;
  float alpha, one_minus_alpha, thr, gamma;
  float* probs; float* probs_orig; float* scores; long long* lbs; float* mask;
  void* gs0; float* out; float* partials; unsigned* ticket;
  int numer_ld, rowsum_ld;   // row strides of numer / rowsum (C and 1, or W and W for the packed [rows, W] layout)
  __nv_bfloat16* hl;   // optional [rows, 64] bf16: hi(32) | lo(32) split of probs for the tensor-core graph kernel
};

__device__ __forceinline__ float pow_gamma(float b, float gamma) {
  if (gamma == 2.f) return b * b;
  if (gamma == 1.f) return b;
  if (gamma == 0.f) return 1.f;
  return powf(b, gamma);
}

// numer rows [row0, row0+nrows) -> dense smem tile; vectorised when the rows are contiguous (ld == C)
__device__ __forceinline__ void load_numer_tile(const float* __restrict__ numer, int ld, long long row0, int nrows, int C, float* s) {
  if (ld == C) { tile_g2s(numer + row0 * C, s, nrows * C); return; }
  for (int i = threadIdx.x; i < nrows * C; i += blockDim.x) {
    const int r = i / C, c = i - r * C;
    s[i] = numer[(row0 + r) * ld + c];
  }
}

// One row of the CoMatch finalisation (LPR lanes): pseudo-label from the weak logits in `sw`, alpha-mix
// with the smoothing sums (`sn`, rowsum), confidence mask, focal soft-CE + gradient against the strong
// logits in `ss`.  Leaves probs in sw, probs_orig in so, the gradient in ss; adds (loss, mask) to acc.
template <int LPR, int EPL, bool FAST = false>
__device__ __forceinline__ void finalize_row(const FinalizeParams& p, const float* __restrict__ prob_avg, float* sw, float* ss,
                                             float* so, const float* sn, long long row0, int r, bool valid, int gl, bool smooth,
                                             float inv_rows, float (&acc)[2], __nv_bfloat16* shl = nullptr) {
  const int C = p.C;
    float x[EPL], e[EPL], pr[EPL];
    float mx, sum;
    // softmax + DA divide + renormalise (comatch.py:163,174-176)
    row_load<LPR, EPL>(sw + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL, FAST>(x, e, mx, sum);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const int c = gl + k * LPR;
      pr[k] = (c < C) ? fdiv<FAST>(fdiv<FAST>(e[k], sum), prob_avg[c]) : 0.f;
      q += pr[k];
    }
    q = group_sum<LPR>(q);
    const float rs = (smooth && valid) ? p.rowsum[(row0 + r) * p.rowsum_ld] : 1.f;
    float psum = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const int c = gl + k * LPR;
      const float po = fdiv<FAST>(pr[k], q);
      if (valid && c < C) so[r * C + c] = po;
      float v = po;
      if (smooth && valid && c < C)  // comatch.py:181-182
        v = __fadd_rn(__fmul_rn(p.alpha, po), __fmul_rn(p.one_minus_alpha, fdiv<FAST>(sn[r * C + c], rs)));
      pr[k] = (c < C) ? v : 0.f;
      psum += pr[k];
    }
    psum = group_sum<LPR>(psum);
    float score; int lb;
    row_argmax<LPR, EPL>(pr, C, gl, score, lb);          // comatch.py:184
    const float m = (valid && score >= p.thr) ? 1.f : 0.f;  // :185
    if (valid) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        if (c < C) sw[r * C + c] = pr[k];
      }
      if (p.hl) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int c = gl + k * LPR;
          if (c < 32) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(pr[k]);
            __nv_bfloat16* dst = shl ? shl + r * 64 + c : p.hl + (row0 + r) * 64 + c;
            dst[0] = hi;
            dst[32] = __float2bfloat16_rn(pr[k] - __bfloat162float(hi));
          }
        }
      }
      if (gl == 0) {
        if (p.scores) p.scores[row0 + r] = score;
        if (p.lbs) p.lbs[row0 + r] = lb;
        if (p.mask) p.mask[row0 + r] = m;
        acc[1] += m;
      }
    }
    // focal soft-CE on the strong view (comatch.py:216-220) + gradient
    row_load<LPR, EPL>(ss + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL, FAST>(x, e, mx, sum);
    const float logsum = flog<FAST>(sum);
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k)
      if (valid && gl + k * LPR < C) d += ((x[k] - mx) - logsum) * pr[k];
    d = group_sum<LPR>(d);
    const float logp = -d * m;
    const float pp = fexp<FAST>(-logp);
    const float om = 1.f - pp;
    const float lrow = pow_gamma(om, p.gamma) * logp;
    if (valid && gl == 0) acc[0] += lrow;
    float dl = pow_gamma(om, p.gamma);
    if (p.gamma != 0.f) dl += p.gamma * pow_gamma(om, p.gamma - 1.f) * pp * logp;
    const float coef = dl * inv_rows * m;
    if (valid) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        if (c < C) ss[r * C + c] = coef * (fdiv<FAST>(e[k], sum) * psum - pr[k]);
      }
    }
}

template <typename T, int LPR, int EPL>
__global__ void __launch_bounds__(kRowThreads) comatch_finalize_kernel(const FinalizeParams p) {
  using Cfg = RowCfg<LPR, EPL>;
  constexpr int ROWS = Cfg::kRowsPerTile;
  extern __shared__ float smem[];
  const int C = p.C;
  float* sw = smem;            // weak logits  -> probs
  float* ss = sw + ROWS * C;   // strong logits -> grad
  float* so = ss + ROWS * C;   // probs_orig
  float* sn = so + ROWS * C;   // numer
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  const bool smooth = p.rowsum != nullptr && p.numer != nullptr;
  const float inv_rows = 1.0f / (float)p.rows;
  const long long ntiles = (p.rows + ROWS - 1) / ROWS;
  float acc[2] = {0.f, 0.f};
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * ROWS;
    const int nrows = (int)min((long long)ROWS, p.rows - row0);
    const int cnt = nrows * C;
    tile_g2s(static_cast<const T*>(p.w) + row0 * C, sw, cnt);
    tile_g2s(static_cast<const T*>(p.s0) + row0 * C, ss, cnt);
    if (smooth) load_numer_tile(p.numer, p.numer_ld, row0, nrows, C, sn);
    __syncthreads();
    const int r = warp * Cfg::kRowsPerWarp + rw;
    const bool valid = r < nrows;
    finalize_row<LPR, EPL>(p, p.prob_avg, sw, ss, so, sn, row0, r, valid, gl, smooth, inv_rows, acc);
    __syncthreads();
    tile_s2g(sw, p.probs + row0 * C, cnt);
    tile_s2g(so, p.probs_orig + row0 * C, cnt);
    tile_s2g(ss, static_cast<T*>(p.gs0) + row0 * C, cnt);
    __syncthreads();
  }
  block_sum<2>(acc);
  float total[2];
  if (grid_reduce_last<2>(acc, p.partials, p.ticket, total) && threadIdx.x == 0) {
    p.out[0] = total[0] / (float)p.rows;
    p.out[1] = total[1] / (float)p.rows;
  }
}

// ================================================= fused K2 + K2b/K4/K7 (+ K5) ==================
// One thread-block cluster does the whole row phase of a CoMatch step for small batches: softmax
// column means -> (cluster exchange through distributed shared memory) -> DA history + prob_avg
// -> finalisation of every row -> optional ring-buffer enqueue -> loss / mask-mean (second
// exchange).  No global tickets, no second and third launch: at the reference's sizes the
// separate kernels cost 12 + 10 + 5 us of pure latency chains (profiles/README.md).
constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedRows = kFusedWarps * 4;      // LPR = 8 -> 4 rows per warp, 64 rows per pass
constexpr int kFusedMaxCluster = 8;

struct FusedParams {
  FinalizeParams f;
  float* ring; int* state; int window;
  // optional enqueue of [unlabeled-weak ; labeled] rows (qf == nullptr: skip)
  void* qf; void* qp; void* qpt; const void* fu; const void* fx; const long long* tx;
  long long n_x; int D; long long* ptr_state; long long K;
  int CL; long long rows_per_cta;
  unsigned long long* dbg;
  int onehot_tail;   // probs_orig has rows + n_x rows: fill the tail with onehot(targets_x) (block for the sharded enqueue)
};

// Bank row g of the (local) ring.
template <typename T>
struct BankRow { T* qf; T* qp; T* qpt; long long row, ld; };
template <typename T>
__device__ __forceinline__ BankRow<T> bank_row(const FusedParams& p, long long g) {
  return {static_cast<T*>(p.qf), static_cast<T*>(p.qp), static_cast<T*>(p.qpt), g, p.K};
}

template <typename T>
__global__ void __launch_bounds__(kFusedThreads) comatch_rows_fused_kernel(const FusedParams p) {
  namespace cg = cooperative_groups;
  constexpr int LPR = 8, EPL = 4, ROWS = kFusedRows;
  constexpr bool FAST = !std::is_same<T, float>::value;       // bf16 storage: MUFU approximations are inside the budget
  extern __shared__ float smem[];
  const FinalizeParams& f = p.f;
  const int C = f.C;
  float* sw = smem;
  float* ss = sw + ROWS * C;
  float* so = ss + ROWS * C;
  float* sn = so + ROWS * C;
  float* sall = sn + ROWS * C;      // [8][32] column sums pushed by every CTA of the cluster
  float* savg = sall + kFusedMaxCluster * 32;   // [32] prob_avg
  float* sfin = savg + 32;          // [8][2]  (loss, mask) pushed into rank 0
  __nv_bfloat16* shl = reinterpret_cast<__nv_bfloat16*>(sfin + kFusedMaxCluster * 2);   // [64][64] bf16 hi/lo staging
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gl = lane % LPR, rw = lane / LPR;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = p.CL;
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const long long r_lo = min(f.rows, (long long)crank * p.rows_per_cta), r_hi = min(f.rows, r_lo + p.rows_per_cta);
  const bool single = (r_hi - r_lo) <= ROWS;                 // all rows of this CTA fit one pass: tiles are loaded once
  const bool smooth = f.rowsum != nullptr && f.numer != nullptr;
  const float inv_rows = 1.0f / (float)f.rows;
  pdl_launch_dependents();
  pdl_wait();                                                // previous kernel complete: global memory may be touched now
  const long long ptr = p.qf ? *reinterpret_cast<volatile long long*>(p.ptr_state) : 0;
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 0);

  // DA history: everything that does not depend on this batch is fetched up front (comatch.py:169-173)
  int count = 1, head = 0;
  float h_old = 0.f;
  if (tid < C) {
    const int count_old = p.state[0];
    head = p.state[1];
    count = min(count_old + 1, p.window);
    const int head_new = (head + 1) % p.window;
    for (int a0 = 0; a0 < count - 1; a0 += 8) {               // oldest -> newest, the newest entry is this batch's mean
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int a = a0 + u;
        const int slot = (head_new - count + a + 2 * p.window) % p.window;
        v[u] = (a < count - 1) ? p.ring[(size_t)slot * C + tid] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) h_old += v[u];
    }
  }
  if (single) {
    const int nrows = (int)(r_hi - r_lo), cnt = nrows * C;
    tile_g2s(static_cast<const T*>(f.w) + r_lo * C, sw, cnt);
    tile_g2s(static_cast<const T*>(f.s0) + r_lo * C, ss, cnt);
    if (smooth) load_numer_tile(f.numer, f.numer_ld, r_lo, nrows, C, sn);
  }
  // ---- phase 1: softmax column sums of the rows this CTA owns (comatch.py:163,169) ----
  float colsum = 0.f;                                        // thread c < C owns class c
  for (long long row0 = r_lo; row0 < r_hi; row0 += ROWS) {
    const int nrows = (int)min((long long)ROWS, r_hi - row0);
    __syncthreads();
    if (!single) tile_g2s(static_cast<const T*>(f.w) + row0 * C, sw, nrows * C);
    __syncthreads();
    const int r = warp * 4 + rw;
    const bool valid = r < nrows;
    float x[EPL], e[EPL];
    float mx, sum;
    row_load<LPR, EPL>(sw + r * C, C, gl, valid, x);
    row_softmax_stats<LPR, EPL, FAST>(x, e, mx, sum);
    if (valid) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int c = gl + k * LPR;
        if (c < C) so[r * C + c] = fdiv<FAST>(e[k], sum);
      }
    }
    __syncthreads();
    if (tid < C)
      for (int rr = 0; rr < nrows; ++rr) colsum += so[rr * C + tid];
  }
  if (tid < C) {                                             // push this CTA's column sums to every CTA of the cluster
    for (int r = 0; r < CL; ++r) (CL > 1 ? cluster.map_shared_rank(sall, r) : sall)[crank * 32 + tid] = colsum;
  }
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 1);
  if (CL > 1) cluster.sync(); else __syncthreads();
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 2);
  // ---- DA history -> prob_avg, computed identically by every CTA ----
  if (tid < C) {
    float tot = 0.f;
    for (int r = 0; r < CL; ++r) tot += sall[r * 32 + tid];  // rank order, local reads
    const float mean = tot / (float)f.rows;
    savg[tid] = (h_old + mean) / (float)count;
    if (crank == 0) {
      p.ring[(size_t)head * C + tid] = mean;                 // slot `head` leaves the window: nobody reads it this step
      const_cast<float*>(f.prob_avg)[tid] = savg[tid];
    }
  }
  __syncthreads();
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 3);

  // ---- phase 2: finalise every row, store, enqueue ----
  float acc[2] = {0.f, 0.f};
  constexpr int epv = 16 / (int)sizeof(T);
  const int vec_per_row = p.D * (int)sizeof(T) / 16;
  for (long long row0 = r_lo; row0 < r_hi; row0 += ROWS) {
    const int nrows = (int)min((long long)ROWS, r_hi - row0);
    const int cnt = nrows * C;
    if (!single) {
      __syncthreads();
      tile_g2s(static_cast<const T*>(f.w) + row0 * C, sw, cnt);
      tile_g2s(static_cast<const T*>(f.s0) + row0 * C, ss, cnt);
      if (smooth) load_numer_tile(f.numer, f.numer_ld, row0, nrows, C, sn);
      __syncthreads();
    }
    if (tid == 0) B200SSL_STAMP(p.dbg, crank, 4);
    const int r = warp * 4 + rw;
    finalize_row<LPR, EPL, FAST>(f, savg, sw, ss, so, sn, row0, r, r < nrows, gl, smooth, inv_rows, acc, f.hl ? shl : nullptr);
    __syncthreads();
    if (tid == 0) B200SSL_STAMP(p.dbg, crank, 5);
    if (f.hl)                                             // [rows, 64] bf16 rows are 128 B: one 16-byte store per thread
      for (int i = tid; i < nrows * 8; i += kFusedThreads)
        reinterpret_cast<uint4*>(f.hl + row0 * 64)[i] = reinterpret_cast<const uint4*>(shl)[i];
    tile_s2g(sw, f.probs + row0 * C, cnt);
    tile_s2g(so, f.probs_orig + row0 * C, cnt);
    tile_s2g(ss, static_cast<T*>(f.gs0) + row0 * C, cnt);
    if (tid == 0) B200SSL_STAMP(p.dbg, crank, 8);
    if (p.qf) {     // unlabeled-weak rows of this pass -> bank rows (ptr + row) % K     (comatch.py:187-196)
      const long long g0 = (ptr + row0) % p.K;
      const BankRow<T> b0 = bank_row<T>(p, g0);
      // common case: the pass does not wrap and starts on a 16-byte boundary -> 128-bit stores only
      if (g0 + nrows <= p.K && b0.row + nrows <= b0.ld && b0.row % epv == 0 && nrows % epv == 0 && b0.ld % epv == 0) {
        for (int i = tid; i < nrows * vec_per_row; i += kFusedThreads)
          reinterpret_cast<uint4*>(b0.qf + b0.row * p.D)[i] = ldg128(static_cast<const T*>(p.fu) + row0 * p.D + (long long)i * epv);
        for (int i = tid; i < cnt / epv; i += kFusedThreads)
          reinterpret_cast<uint4*>(b0.qp + b0.row * C)[i] = pack16(*reinterpret_cast<const float(*)[epv]>(so + i * epv), T());
        if (p.qpt) {
          const int chunks = nrows / epv;
          for (int i = tid; i < C * chunks; i += kFusedThreads) {
            const int c = i / chunks, ch = i - c * chunks;
            float v[epv];
#pragma unroll
            for (int k = 0; k < epv; ++k) v[k] = so[(ch * epv + k) * C + c];
            *reinterpret_cast<uint4*>(b0.qpt + (size_t)c * b0.ld + b0.row + ch * epv) = pack16(v, T());
          }
        }
      } else {
        for (int i = tid; i < nrows * vec_per_row; i += kFusedThreads) {
          const int rr = i / vec_per_row, v = i - rr * vec_per_row;
          const BankRow<T> b = bank_row<T>(p, (ptr + row0 + rr) % p.K);
          const uint4 val = ldg128(static_cast<const T*>(p.fu) + (row0 + rr) * p.D + v * epv);
          *reinterpret_cast<uint4*>(b.qf + b.row * p.D + v * epv) = val;
        }
        for (int i = tid; i < cnt; i += kFusedThreads) {
          const int rr = i / C, c = i - rr * C;
          const BankRow<T> b = bank_row<T>(p, (ptr + row0 + rr) % p.K);
          b.qp[b.row * C + c] = from_f32<T>(so[i]);
        }
        if (p.qpt)                                          // transposed copy: consecutive threads -> consecutive bank rows
          for (int i = tid; i < cnt; i += kFusedThreads) {
            const int c = i / nrows, rr = i - c * nrows;
            const BankRow<T> b = bank_row<T>(p, (ptr + row0 + rr) % p.K);
            b.qpt[(size_t)c * b.ld + b.row] = from_f32<T>(so[rr * C + c]);
          }
      }
    }
  }
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 9);
  const int n_x = (int)p.n_x;
  if (p.qf && n_x > 0) {       // labeled rows: [feats_x ; onehot(targets_x)], spread over the CTAs
    const long long g0 = (ptr + f.rows) % p.K;
    const int i0 = crank * kFusedThreads + tid, istep = CL * kFusedThreads;
    const BankRow<T> b0 = bank_row<T>(p, g0);
    if (g0 + n_x <= p.K && b0.row + n_x <= b0.ld && b0.row % epv == 0 && n_x % epv == 0 && b0.ld % epv == 0) {   // 128-bit stores only
      for (int i = i0; i < n_x * vec_per_row; i += istep)
        reinterpret_cast<uint4*>(b0.qf + b0.row * p.D)[i] = ldg128(static_cast<const T*>(p.fx) + (size_t)i * epv);
      for (int i = i0; i < n_x * C / epv; i += istep) {
        float v[epv];
#pragma unroll
        for (int k = 0; k < epv; ++k) {
          const int e = i * epv + k, rr = e / C;
          v[k] = (e - rr * C == (int)p.tx[rr]) ? 1.f : 0.f;
        }
        reinterpret_cast<uint4*>(b0.qp + b0.row * C)[i] = pack16(v, T());
      }
      if (p.qpt) {
        const int chunks = n_x / epv;
        for (int i = i0; i < C * chunks; i += istep) {
          const int c = i / chunks, ch = i - c * chunks;
          float v[epv];
#pragma unroll
          for (int k = 0; k < epv; ++k) v[k] = (c == (int)p.tx[ch * epv + k]) ? 1.f : 0.f;
          *reinterpret_cast<uint4*>(b0.qpt + (size_t)c * b0.ld + b0.row + ch * epv) = pack16(v, T());
        }
      }
    } else {
      for (int i = i0; i < n_x * vec_per_row; i += istep) {
        const int rr = i / vec_per_row, v = i - rr * vec_per_row;
        const BankRow<T> b = bank_row<T>(p, (ptr + f.rows + rr) % p.K);
        const uint4 val = ldg128(static_cast<const T*>(p.fx) + (size_t)rr * p.D + v * epv);
        *reinterpret_cast<uint4*>(b.qf + b.row * p.D + v * epv) = val;
      }
      for (int i = i0; i < n_x * C; i += istep) {
        const int rr = i / C, c = i - rr * C;
        const BankRow<T> b = bank_row<T>(p, (ptr + f.rows + rr) % p.K);
        const T val = from_f32<T>(c == (int)p.tx[rr] ? 1.f : 0.f);
        b.qp[b.row * C + c] = val;
        if (p.qpt) b.qpt[(size_t)c * b.ld + b.row] = val;
      }
    }
  }
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 6);
  if (p.onehot_tail) {      // [probs_orig ; onehot(targets_x)] = the probability block of the enqueue (comatch.py:188-189)
    for (int i = crank * kFusedThreads + tid; i < n_x * C; i += CL * kFusedThreads) {
      const int rr = i / C, c = i - rr * C;
      f.probs_orig[(f.rows + rr) * C + c] = (c == (int)p.tx[rr]) ? 1.f : 0.f;
    }
  }
  // ---- loss / mask mean: CTA sum pushed into rank 0, which folds the CTAs in rank order ----
  __shared__ float s_part[2][kFusedWarps];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float t = warp_sum(acc[i]);
    if (lane == 0) s_part[i][warp] = t;
  }
  __syncthreads();
  if (tid < 2) {
    float t = 0.f;
    for (int w = 0; w < kFusedWarps; ++w) t += s_part[tid][w];
    (CL > 1 ? cluster.map_shared_rank(sfin, 0) : sfin)[crank * 2 + tid] = t;
  }
  if (CL > 1) cluster.sync(); else __syncthreads();          // last exchange: afterwards only rank 0 reads, and only its own smem
  if (crank == 0 && tid < 2) {
    float t = 0.f;
    for (int r = 0; r < CL; ++r) t += sfin[r * 2 + tid];
    f.out[tid] = t / (float)f.rows;
  }
  if (crank == 0 && tid == 0) {
    p.state[0] = count;
    p.state[1] = (head + 1) % p.window;
    if (p.qf) p.ptr_state[0] = (ptr + f.rows + p.n_x) % p.K;       // comatch.py:196
  }
  if (tid == 0) B200SSL_STAMP(p.dbg, crank, 7);
}

// ---- grad *= *scale -----------------------------------------------------------
template <typename T>
__global__ void scale_kernel(T* g, long long n, const float* scale, float factor) {
  const float sc = *scale * factor;
  constexpr int N = Vec16<T>::N;
  const long long nvec = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? n / N : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float f[N];
    unpack16(*reinterpret_cast<const uint4*>(g + v * N), f, T());
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] *= sc;
    *reinterpret_cast<uint4*>(g + v * N) = pack16(f, T());
  }
  for (long long i = nvec * N + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    g[i] = from_f32<T>(to_f32<T>(g[i]) * sc);
}

// ---- launch helpers -------------------------------------------------------------
struct RowLaunch { int grid; size_t smem; };

template <int LPR, int EPL>
RowLaunch row_launch(long long rows, int C, int ntile_arrays, int extra_floats = 0) {
  constexpr int ROWS = RowCfg<LPR, EPL>::kRowsPerTile;
  const long long ntiles = (rows + ROWS - 1) / ROWS;
  RowLaunch l;
  l.grid = (int)(ntiles < kMaxRowCtas ? ntiles : kMaxRowCtas);
  if (l.grid < 1) l.grid = 1;
  l.smem = ((size_t)ntile_arrays * ROWS * C + extra_floats) * sizeof(float);
  return l;
}

int check_rows(const char* fn, long long rows, int C, int dtype) {
  if (rows <= 0) return fail(B200SSL_E_SHAPE, "%s: rows must be > 0 (got %lld)", fn, rows);
  if (C < 2 || C > B200SSL_MAX_CLASSES) return fail(B200SSL_E_SHAPE, "%s: classes %d outside [2, %d]", fn, C, B200SSL_MAX_CLASSES);
  if (dtype != B200SSL_F32 && dtype != B200SSL_BF16) return fail(B200SSL_E_DTYPE, "%s: dtype %d (want f32=0 or bf16=1)", fn, dtype);
  return 0;
}
int check_ws(const char* fn, void* ws, size_t have, size_t need) {
  if (!ws) return fail(B200SSL_E_NULL, "%s: workspace is NULL", fn);
  if (reinterpret_cast<uintptr_t>(ws) & 255u) return fail(B200SSL_E_ALIGN, "%s: workspace must be 256-byte aligned", fn);
  if (have < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, have, need);
  return 0;
}

// dispatch on (dtype, classes) -> (T, LPR, EPL)
#define B200SSL_ROW_DISPATCH(DTYPE, C, ...)                                  \
  do {                                                                         \
    if ((DTYPE) == B200SSL_F32) {                                              \
      using T = float;                                                         \
      if ((C) <= 32) { constexpr int LPR = 8, EPL = 4; __VA_ARGS__; }                 \
      else if ((C) <= 128) { constexpr int LPR = 32, EPL = 4; __VA_ARGS__; }          \
      else { constexpr int LPR = 32, EPL = 32; __VA_ARGS__; }                         \
    } else {                                                                   \
      using T = __nv_bfloat16;                                                 \
      if ((C) <= 32) { constexpr int LPR = 8, EPL = 4; __VA_ARGS__; }                 \
      else if ((C) <= 128) { constexpr int LPR = 32, EPL = 4; __VA_ARGS__; }          \
      else { constexpr int LPR = 32, EPL = 32; __VA_ARGS__; }                         \
    }                                                                          \
  } while (0)

template <typename K>
int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_fixmatch_head_fwd_bwd(const void* logits_w, const void* logits_s, const void* logits_s2,
                                             void* grad_s, void* grad_s2, int64_t rows, int32_t classes,
                                             int32_t dtype, float p_cutoff, float inv_T, int32_t use_hard_labels,
                                             float* out_scalars, int64_t* idx, float* mask, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_fixmatch_head_fwd_bwd";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits_w || !logits_s || !grad_s || !out_scalars) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if ((logits_s2 == nullptr) != (grad_s2 == nullptr)) return fail(B200SSL_E_NULL, "%s: logits_s2 and grad_s2 go together", fn);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes + sizeof(float) * 3 * kMaxRowCtas)) return e;
  HeadParams p{logits_w, logits_s, logits_s2, grad_s, grad_s2, rows, classes, p_cutoff, inv_T,
               use_hard_labels ? 1 : 0, out_scalars, reinterpret_cast<long long*>(idx), mask,
               reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes),
               reinterpret_cast<unsigned*>(workspace)};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 3);
    auto k = fixmatch_head_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_labeled_ce_fwd_bwd(const void* logits, const int64_t* targets, const float* class_weights,
                                          void* grad, int64_t rows, int32_t classes, int32_t dtype, int32_t poly,
                                          float epsilon, float* out_scalar, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  const char* fn = "b200ssl_labeled_ce_fwd_bwd";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits || !targets || !grad || !out_scalar) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes + sizeof(float) * kMaxRowCtas)) return e;
  LabeledParams p{logits, reinterpret_cast<const long long*>(targets), class_weights, grad, rows, classes,
                  poly ? 1 : 0, poly ? epsilon : 0.f, out_scalar,
                  reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes),
                  reinterpret_cast<unsigned*>(workspace) + 1, reinterpret_cast<unsigned*>(workspace) + kWsBadLabelSlot};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 1);
    auto k = labeled_ce_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_ce_rows_fwd_bwd(const void* logits, const int64_t* targets, const float* soft_targets,
                                       const float* class_weights, void* grad_unit, float* loss_rows, int64_t rows,
                                       int32_t classes, int32_t dtype, int32_t poly, float epsilon, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_ce_rows_fwd_bwd";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits || !grad_unit || !loss_rows) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if ((targets == nullptr) == (soft_targets == nullptr)) return fail(B200SSL_E_ARG, "%s: exactly one of targets / soft_targets", fn);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes)) return e;
  RowCeParams p{logits, reinterpret_cast<const long long*>(targets), soft_targets, class_weights, grad_unit, loss_rows, rows,
                classes, poly ? 1 : 0, poly ? epsilon : 0.f, reinterpret_cast<unsigned*>(workspace) + kWsBadLabelSlot};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 1);
    auto k = ce_rows_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_scale_rows(const void* grad_in, void* grad_out, int64_t rows, int32_t classes, int32_t dtype,
                                  const float* row_scale, void* stream) {
  const char* fn = "b200ssl_scale_rows";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!grad_in || !grad_out || !row_scale) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  long long blocks = (rows * classes + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (dtype == B200SSL_F32)
    scale_rows_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>(static_cast<const float*>(grad_in), static_cast<float*>(grad_out), rows, classes, row_scale);
  else
    scale_rows_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(grad_in), static_cast<__nv_bfloat16*>(grad_out), rows, classes, row_scale);
  return check_launch(fn);
}

extern "C" int b200ssl_eval_head(const void* logits, const int64_t* targets, int64_t rows, int32_t classes, int32_t dtype,
                                 uint64_t* confusion, float* loss_out, int64_t* pred, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  const char* fn = "b200ssl_eval_head";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits || !targets || !confusion || !loss_out) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes + sizeof(float) * 2 * kMaxRowCtas)) return e;
  EvalParams p{logits, reinterpret_cast<const long long*>(targets), rows, classes, reinterpret_cast<unsigned long long*>(confusion),
               loss_out, reinterpret_cast<long long*>(pred), reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes),
               reinterpret_cast<unsigned*>(workspace) + 5, reinterpret_cast<unsigned*>(workspace) + kWsBadLabelSlot};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 1);
    auto k = eval_head_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_bad_label_count(void* workspace, uint32_t* count, int32_t reset) {
  if (!workspace || !count) return fail(B200SSL_E_NULL, "b200ssl_bad_label_count: NULL pointer");
  unsigned* slot = reinterpret_cast<unsigned*>(workspace) + kWsBadLabelSlot;
  cudaError_t e = cudaMemcpy(count, slot, sizeof(uint32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && reset) e = cudaMemset(slot, 0, sizeof(uint32_t));
  return e == cudaSuccess ? 0 : fail((int)e, "b200ssl_bad_label_count: %s", cudaGetErrorString(e));
}

extern "C" int b200ssl_comatch_da(const void* logits_u_w, int64_t rows, int32_t classes, int32_t dtype,
                                  float* da_ring, int32_t* da_state, int32_t window, float* prob_avg,
                                  float* col_mean_out, void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_comatch_da";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits_u_w || !da_ring || !da_state || !prob_avg) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (window < 1 || window > B200SSL_DA_WINDOW_MAX) return fail(B200SSL_E_ARG, "%s: window %d outside [1,%d]", fn, window, B200SSL_DA_WINDOW_MAX);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes + sizeof(float) * (size_t)classes * kNumSMs)) return e;
  DaParams p{logits_u_w, rows, classes, da_ring, da_state, window, prob_avg, col_mean_out,
             reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes),
             reinterpret_cast<unsigned*>(workspace) + 2};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 1, classes);
    if (l.grid > kNumSMs) l.grid = kNumSMs;
    auto k = comatch_da_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_comatch_finalize(const void* logits_u_w, const void* logits_u_s0, const float* prob_avg,
                                        const float* rowsum, const float* numer, int32_t rowsum_ld, int32_t numer_ld,
                                        int64_t rows, int32_t classes,
                                        int32_t dtype, float alpha, float one_minus_alpha, float thr, float gamma,
                                        float* probs, float* probs_orig, void* probs_hl, float* scores, int64_t* lbs,
                                        float* mask, void* grad_s0, float* out_scalars, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_comatch_finalize";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (!logits_u_w || !logits_u_s0 || !prob_avg || !probs || !probs_orig || !grad_s0 || !out_scalars)
    return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if ((rowsum == nullptr) != (numer == nullptr)) return fail(B200SSL_E_NULL, "%s: rowsum and numer go together", fn);
  if (probs_hl && classes > 32) return fail(B200SSL_E_SHAPE, "%s: probs_hl needs classes <= 32", fn);
  if (int e = check_ws(fn, workspace, workspace_bytes, kWsHeaderBytes + sizeof(float) * 2 * kMaxRowCtas)) return e;
  FinalizeParams p{logits_u_w, logits_u_s0, prob_avg, rowsum, numer, rows, classes, alpha, one_minus_alpha, thr, gamma,
                   probs, probs_orig, scores, reinterpret_cast<long long*>(lbs), mask, grad_s0, out_scalars,
                   reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes),
                   reinterpret_cast<unsigned*>(workspace) + 3, numer_ld > 0 ? numer_ld : classes, rowsum_ld > 0 ? rowsum_ld : 1,
                   static_cast<__nv_bfloat16*>(probs_hl)};
  B200SSL_ROW_DISPATCH(dtype, classes, {
    RowLaunch l = row_launch<LPR, EPL>(rows, classes, 4);
    auto k = comatch_finalize_kernel<T, LPR, EPL>;
    if (int e = set_smem(k, l.smem)) return e;
    k<<<l.grid, kRowThreads, l.smem, as_stream(stream)>>>(p);
  });
  return check_launch(fn);
}

extern "C" int b200ssl_scale_inplace(void* grad, int64_t numel, int32_t dtype, const float* scale, float factor,
                                     void* stream) {
  const char* fn = "b200ssl_scale_inplace";
  if (!grad || !scale) return fail(B200SSL_E_NULL, "%s: NULL pointer", fn);
  if (numel <= 0) return fail(B200SSL_E_SHAPE, "%s: numel must be > 0", fn);
  const int threads = 256;
  long long blocks = (numel / 4 + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (dtype == B200SSL_F32) scale_kernel<float><<<(int)blocks, threads, 0, as_stream(stream)>>>(static_cast<float*>(grad), numel, scale, factor);
  else if (dtype == B200SSL_BF16) scale_kernel<__nv_bfloat16><<<(int)blocks, threads, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(grad), numel, scale, factor);
  else return fail(B200SSL_E_DTYPE, "%s: dtype %d", fn, dtype);
  return check_launch(fn);
}

extern "C" int b200ssl_comatch_rows_fused(const void* logits_u_w, const void* logits_u_s0, const float* rowsum,
                                          const float* numer, int32_t rowsum_ld, int32_t numer_ld, int64_t rows,
                                          int32_t classes, int32_t dtype, float alpha,
                                          float one_minus_alpha, float thr, float gamma, float* da_ring, int32_t* da_state,
                                          int32_t window, float* prob_avg, float* probs, float* probs_orig, void* probs_hl,
                                          float* scores, int64_t* lbs, float* mask, void* grad_s0, float* out_scalars,
                                          void* queue_feats, void* queue_probs, void* queue_probs_t, const void* feats_u_w,
                                          const void* feats_x, const int64_t* targets_x, int64_t n_x, int32_t dim,
                                          int64_t* ptr_state, int64_t bank_rows, int32_t onehot_tail, void* stream) {
  const char* fn = "b200ssl_comatch_rows_fused";
  if (int e = check_rows(fn, rows, classes, dtype)) return e;
  if (classes > 32) return fail(B200SSL_E_SHAPE, "%s: classes %d > 32 (use the separate kernels)", fn, classes);
  if (!logits_u_w || !logits_u_s0 || !da_ring || !da_state || !prob_avg || !probs || !probs_orig || !grad_s0 || !out_scalars)
    return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if ((rowsum == nullptr) != (numer == nullptr)) return fail(B200SSL_E_NULL, "%s: rowsum and numer go together", fn);
  if (window < 1 || window > B200SSL_DA_WINDOW_MAX) return fail(B200SSL_E_ARG, "%s: window %d outside [1,%d]", fn, window, B200SSL_DA_WINDOW_MAX);
  if (queue_feats) {
    const size_t es = dtype == B200SSL_F32 ? 4 : 2;
    if (!queue_probs || !feats_u_w || !ptr_state || (n_x > 0 && (!feats_x || !targets_x)))
      return fail(B200SSL_E_NULL, "%s: NULL enqueue tensor", fn);
    if (dim < 1 || (dim * es) % 16 || bank_rows <= 0 || rows + n_x > bank_rows || n_x < 0)
      return fail(B200SSL_E_SHAPE, "%s: enqueue needs 16-byte rows and rows + n_x <= bank_rows", fn);
    if ((reinterpret_cast<uintptr_t>(queue_feats) | reinterpret_cast<uintptr_t>(feats_u_w) | reinterpret_cast<uintptr_t>(feats_x)) & 15u)
      return fail(B200SSL_E_ALIGN, "%s: embeddings must be 16-byte aligned", fn);
  }
  FusedParams p{};
  p.f = FinalizeParams{logits_u_w, logits_u_s0, prob_avg, rowsum, numer, rows, classes, alpha, one_minus_alpha, thr, gamma,
                       probs, probs_orig, scores, reinterpret_cast<long long*>(lbs), mask, grad_s0, out_scalars, nullptr, nullptr,
                       numer_ld > 0 ? numer_ld : classes, rowsum_ld > 0 ? rowsum_ld : 1, static_cast<__nv_bfloat16*>(probs_hl)};
  p.ring = da_ring; p.state = da_state; p.window = window;
  p.qf = queue_feats; p.qp = queue_probs; p.qpt = queue_probs_t; p.fu = feats_u_w; p.fx = feats_x;
  p.tx = reinterpret_cast<const long long*>(targets_x); p.n_x = n_x; p.D = dim;
  p.ptr_state = reinterpret_cast<long long*>(ptr_state); p.K = bank_rows;
  p.onehot_tail = onehot_tail ? 1 : 0;
  p.dbg = debug_timing_buffer(PDL_ROWS);
  if (onehot_tail && (n_x > 0 && !targets_x)) return fail(B200SSL_E_NULL, "%s: onehot_tail needs targets_x", fn);
  long long cl = (rows + kFusedRows - 1) / kFusedRows;
  if (cl > kFusedMaxCluster) cl = kFusedMaxCluster;
  int clp = 1;
  while (clp * 2 <= cl) clp *= 2;
  if (clp < cl) clp *= 2;                      // round up to a power of two (cluster sizes 1, 2, 4, 8)
  p.CL = clp;
  p.rows_per_cta = (((rows + clp - 1) / clp) + 7) & ~7LL;          // multiples of 8 rows keep the tiles 16-byte aligned
  const size_t smem = ((size_t)4 * kFusedRows * classes + kFusedMaxCluster * 32 + 32 + kFusedMaxCluster * 2) * sizeof(float) +
                      (size_t)kFusedRows * 64 * sizeof(__nv_bfloat16) + 16;
  cudaError_t e;
  if (dtype == B200SSL_F32)
    e = launch_pdl(PDL_ROWS, comatch_rows_fused_kernel<float>, dim3((unsigned)clp, 1, 1), dim3(kFusedThreads, 1, 1), smem, as_stream(stream),
                   dim3((unsigned)clp, 1, 1), p);
  else
    e = launch_pdl(PDL_ROWS, comatch_rows_fused_kernel<__nv_bfloat16>, dim3((unsigned)clp, 1, 1), dim3(kFusedThreads, 1, 1), smem,
                   as_stream(stream), dim3((unsigned)clp, 1, 1), p);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}
