// CoMatch memory bank: exact-fp32 smoothing partial sums (K3, SIMT path) and the
// ring-buffer enqueue (K5).
//
// K3 here is the *fp32-exact* path: the reference's torch.mm is true fp32 (TF32
// off), and single-pass TF32 misses the 1e-5 parity budget on per-element probs
// (SURVEY H3), so fp32 storage uses FFMA tiles.  bf16 storage is served by the
// tcgen05/TMEM kernel in bank_tc.cu.  Either way A[rows, K] is never
// materialised: S tile -> exp -> (row-sum, E.Qp) is fused per 64x64 tile
// (comatch.py:180-181).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "tiles.cuh"

namespace b200ssl {
namespace {

struct SmoothParams {
  const void* f; const void* qf; const void* qp;
  long long rows, bank_rows; int D, C; float tau;
  float* rowsum; float* numer;
  float* part; unsigned* tickets;  // split-K partials [nsplit][rows_pad][1+C], per row-tile tickets
  int nsplit, tiles_per_split, W; long long rows_pad;
  int numer_ld, rowsum_ld;
};

// grid = (row tiles, nsplit); 256 threads; NACC = ceil(maxC/4) numerator
// accumulators per thread (thread owns row tid&63, columns (tid>>6) + 4m).
template <typename T, int NACC>
__global__ void __launch_bounds__(kTileThreads) bank_smooth_simt_kernel(const SmoothParams p) {
  extern __shared__ float smem[];
  const int D = p.D, C = p.C, ldd = D + 1;
  float* As = smem;                 // [TM][D+1]  queries
  float* Bs = As + kTM * ldd;       // [TN][D+1]  bank keys
  float* Es = Bs + kTN * ldd;       // [TM][TN+1] exp(S/tau)
  float* Ps = Es + kTM * (kTN + 1); // [TN][C]    bank probabilities
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const long long i0 = (long long)blockIdx.x * kTM;
  const int mrows = (int)min((long long)kTM, p.rows - i0);
  const int split = blockIdx.y;
  const long long ktile0 = (long long)split * p.tiles_per_split;
  const long long nktiles = (p.bank_rows + kTN - 1) / kTN;
  const long long ktile1 = min(nktiles, ktile0 + p.tiles_per_split);

  load_tile_padded(static_cast<const T*>(p.f) + i0 * D, mrows, kTM, D, ldd, As);

  float rs[4] = {0.f, 0.f, 0.f, 0.f};
  float nacc[NACC];
#pragma unroll
  for (int m = 0; m < NACC; ++m) nacc[m] = 0.f;
  const int prow = tid & 63, cg = tid >> 6;

  for (long long kt = ktile0; kt < ktile1; ++kt) {
    const long long k0 = kt * kTN;
    const int nk = (int)min((long long)kTN, p.bank_rows - k0);
    load_tile_padded(static_cast<const T*>(p.qf) + k0 * D, nk, kTN, D, ldd, Bs);
    load_tile_dense(static_cast<const T*>(p.qp) + k0 * C, nk * C, kTN * C, Ps);
    __syncthreads();
    float acc[4][4];
    tile_dot_4x4(As, Bs, ldd, D, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = tx + 16 * j;
        const float e = (col < nk) ? expf(__fdiv_rn(acc[i][j], p.tau)) : 0.f;  // comatch.py:180
        Es[(ty + 16 * i) * (kTN + 1) + col] = e;
        rs[i] += e;
      }
    __syncthreads();
    for (int k = 0; k < kTN; ++k) {
      const float e = Es[prow * (kTN + 1) + k];
#pragma unroll
      for (int m = 0; m < NACC; ++m) {
        const int c = cg + 4 * m;
        if (c < C) nacc[m] = fmaf(e, Ps[k * C + c], nacc[m]);
      }
    }
    __syncthreads();
  }
  // row sums: fold the 16 column-lanes (tx) of each row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
  }
  const bool direct = p.nsplit == 1;
  const int W = p.W;   // floats per row of a split partial: [rowsum, numer[0..C), pad]
  float* pbase = p.part + (size_t)split * p.rows_pad * W;
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 16 * i;
      if (direct) { if (r < mrows) p.rowsum[(i0 + r) * p.rowsum_ld] = rs[i]; }
      else pbase[(i0 + r) * W] = rs[i];
    }
  }
#pragma unroll
  for (int m = 0; m < NACC; ++m) {
    const int c = cg + 4 * m;
    if (c < C) {
      if (direct) { if (prow < mrows) p.numer[(i0 + prow) * p.numer_ld + c] = nacc[m]; }
      else pbase[(i0 + prow) * W + 1 + c] = nacc[m];
    }
  }
  if (direct) return;
  // last split of this row tile folds the partials in split order (deterministic)
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&p.tickets[blockIdx.x], 1u) == (unsigned)p.nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  fold_splits_vec4(reinterpret_cast<const float4*>(p.part + (size_t)i0 * W), (size_t)p.rows_pad * W / 4, p.nsplit,
                   mrows * W / 4, [&](int i, float4 v) {
                     const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                     for (int j = 0; j < 4; ++j) {
                       const int e = 4 * i + j, row = e / W, col = e - row * W;
                       if (col == 0) p.rowsum[(i0 + row) * p.rowsum_ld] = vv[j];
                       else if (col <= C) p.numer[(i0 + row) * p.numer_ld + col - 1] = vv[j];
                     }
                   });
  if (tid == 0) p.tickets[blockIdx.x] = 0u;
}

// ---- K5 --------------------------------------------------------------------------
struct EnqueueParams {
  void* qpt;  // optional transposed, class-padded copy [32][shard_rows] (tensor-core path)
  void* qf; void* qp; const void* fu; const void* fx; const float* po; const long long* tx;
  long long n_u, n_x; int D, C;
  long long ptr, block_offset, K, shard_begin, shard_rows;
  long long* ptr_state; long long advance;  // device-resident {write pointer, ticket}; see b200ssl.h
};

template <typename T>
__global__ void __launch_bounds__(256) bank_enqueue_kernel(const EnqueueParams p) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // The write pointer lives on the device when ptr_state is given (CUDA-graph replays
  // must not bake it in); every CTA reads it before taking a ticket, the last CTA to
  // finish advances it: ptr = (ptr + n) % K  (comatch.py:196).
  __shared__ long long s_ptr;
  if (threadIdx.x == 0) s_ptr = p.ptr_state ? *reinterpret_cast<volatile long long*>(p.ptr_state) : p.ptr;
  __syncthreads();
  const long long ptr = s_ptr;
  if (p.ptr_state && threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(p.ptr_state + 1), 1ull);
    if (t == gridDim.x - 1) {
      p.ptr_state[1] = 0;
      if (p.advance) p.ptr_state[0] = (ptr + p.advance) % p.K;
      __threadfence();
    }
  }
  if (r >= p.n_u + p.n_x) return;
  const long long g = (ptr + p.block_offset + r) % p.K;  // comatch.py:194-196 with wrap
  const long long local = g - p.shard_begin;
  if (local < 0 || local >= p.shard_rows) return;
  const bool lab = r >= p.n_u;                              // rows: [unlabeled-weak ; labeled]  (:187)
  const T* src = lab ? static_cast<const T*>(p.fx) + (r - p.n_u) * p.D : static_cast<const T*>(p.fu) + r * p.D;
  T* df = static_cast<T*>(p.qf) + local * p.D;
  for (int d = lane; d < p.D; d += 32) df[d] = src[d];
  T* dp = static_cast<T*>(p.qp) + local * p.C;
  T* dt = p.qpt ? static_cast<T*>(p.qpt) + local : nullptr;
  const int y = lab ? (int)p.tx[r - p.n_u] : -1;            // one-hot (:188)
  const float* sp = lab ? nullptr : p.po + r * p.C;         // probs_orig (:189)
  for (int c = lane; c < p.C; c += 32) {
    const T v = from_f32<T>(lab ? (c == y ? 1.f : 0.f) : sp[c]);
    dp[c] = v;
    if (dt) dt[(size_t)c * p.shard_rows] = v;
  }
}

}  // namespace

// shared with api.cu (workspace sizing)
int smooth_nsplit(long long rows, long long bank_rows, int* tiles_per_split) {
  const long long row_tiles = (rows + kTM - 1) / kTM;
  const long long ktiles = (bank_rows + kTN - 1) / kTN;
  long long want = (2 * kNumSMs + row_tiles - 1) / row_tiles;
  if (want < 1) want = 1;
  if (want > ktiles) want = ktiles;
  if (want > 64) want = 64;
  const long long tps = (ktiles + want - 1) / want;
  const long long nsplit = (ktiles + tps - 1) / tps;
  if (tiles_per_split) *tiles_per_split = (int)tps;
  return (int)nsplit;
}

}  // namespace b200ssl

using namespace b200ssl;

namespace b200ssl {
int g_f32_simt = getenv("B200SSL_K3_F32_SIMT") != nullptr;   // also read by contrast.cu
int bank_smooth_tc(const void* feats, const void* queue_feats, const void* queue_probs_t, long long rows,
                   long long bank_rows, int classes, float temperature, float* rowsum, float* numer, int rowsum_ld,
                   int numer_ld, const b200ssl_bank_shards* shards, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int bank_smooth_tc_f32(const float* feats, const float* queue_feats, const float* queue_probs, long long rows, long long bank_rows,
                       int classes, float temperature, float* rowsum, float* numer, int rowsum_ld, int numer_ld, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);
}

extern "C" void b200ssl_debug_set_k3_f32_simt(int32_t on) { b200ssl::g_f32_simt = on != 0; }

extern "C" int b200ssl_bank_smooth_partial(const void* feats_u_w, const void* queue_feats, const void* queue_probs,
                                           const void* queue_probs_t, int64_t rows, int64_t bank_rows, int32_t dim, int32_t classes,
                                           int32_t dtype, float temperature, float* rowsum, float* numer,
                                           int32_t rowsum_ld, int32_t numer_ld, const b200ssl_bank_shards* shards,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "b200ssl_bank_smooth_partial";
  if (!feats_u_w || !rowsum || !numer || (!shards && (!queue_feats || !queue_probs))) return fail(B200SSL_E_NULL, "%s: NULL tensor", fn);
  if (shards) {
    // directly addressed sharded bank: tensor-core path only
    if (shards->world < 2 || shards->world > 8 || shards->rank < 0 || shards->rank >= shards->world || !shards->arenas_host ||
        !shards->arenas_dev || shards->shard_rows <= 0 || shards->shard_rows % 8 ||
        bank_rows != shards->shard_rows * (shards->replicated ? 1 : shards->world))
      return fail(B200SSL_E_ARG, "%s: bad shard table (2..8 ranks, shard_rows a multiple of 8, bank_rows = world*shard_rows or, "
                                 "replicated, = shard_rows)", fn);
    if (dtype != B200SSL_BF16 || dim != 64 || classes > 31 || classes < 2 || (reinterpret_cast<uintptr_t>(feats_u_w) & 15u) ||
        ((shards->feats_offset | shards->probs_t_offset) & 127u))
      return fail(B200SSL_E_DTYPE, "%s: the directly addressed sharded bank needs bf16, dim 64, classes <= 31, aligned rows", fn);
    if (!(temperature > 0.f) || rows <= 0 || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u))
      return fail(B200SSL_E_ARG, "%s: rows / temperature / workspace", fn);
    return bank_smooth_tc(feats_u_w, nullptr, nullptr, rows, bank_rows, classes, temperature, rowsum, numer,
                          rowsum_ld > 0 ? rowsum_ld : 1, numer_ld > 0 ? numer_ld : classes, shards, workspace, workspace_bytes,
                          as_stream(stream));
  }
  if (rows <= 0 || bank_rows <= 0) return fail(B200SSL_E_SHAPE, "%s: rows=%lld bank_rows=%lld", fn, (long long)rows, (long long)bank_rows);
  if (dim < 8 || dim > B200SSL_MAX_EMB_DIM || dim % 8) return fail(B200SSL_E_SHAPE, "%s: dim %d must be a multiple of 8 in [8,%d]", fn, dim, B200SSL_MAX_EMB_DIM);
  if (classes < 2 || classes > 128) return fail(B200SSL_E_SHAPE, "%s: classes %d outside [2,128]", fn, classes);
  if (!(temperature > 0.f)) return fail(B200SSL_E_ARG, "%s: temperature must be > 0", fn);
  if (dtype != B200SSL_F32 && dtype != B200SSL_BF16) return fail(B200SSL_E_DTYPE, "%s: dtype %d", fn, dtype);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(B200SSL_E_ALIGN, "%s: workspace NULL or not 256-byte aligned", fn);
  // bf16 bank with 128-byte rows: tcgen05 / TMEM / TMA kernel (bank_tc.cu); everything else: exact-fp32 FFMA tiles
  if (dtype == B200SSL_BF16 && dim == 64 && classes <= 31 && queue_probs_t && bank_rows % 8 == 0 &&
      !(reinterpret_cast<uintptr_t>(feats_u_w) & 15u) && !(reinterpret_cast<uintptr_t>(queue_feats) & 15u) &&
      !(reinterpret_cast<uintptr_t>(queue_probs_t) & 15u))
    return bank_smooth_tc(feats_u_w, queue_feats, queue_probs_t, rows, bank_rows, classes, temperature, rowsum, numer,
                          rowsum_ld > 0 ? rowsum_ld : 1, numer_ld > 0 ? numer_ld : classes, nullptr, workspace, workspace_bytes,
                          as_stream(stream));
  // fp32 storage with 64-wide embeddings: the same tensor-core kernel on bf16 hi + mid operands (1e-6 on the smoothed
  // probabilities, the reference's fp32 torch.mm is 3e-7); B200SSL_K3_F32_SIMT=1 keeps the exact-fp32 FFMA tiles (A/B)
  if (dtype == B200SSL_F32 && dim == 64 && classes <= 31 && !g_f32_simt && !(reinterpret_cast<uintptr_t>(feats_u_w) & 15u) &&
      !(reinterpret_cast<uintptr_t>(queue_feats) & 15u))
    return bank_smooth_tc_f32(static_cast<const float*>(feats_u_w), static_cast<const float*>(queue_feats),
                              static_cast<const float*>(queue_probs), rows, bank_rows, classes, temperature, rowsum, numer,
                              rowsum_ld > 0 ? rowsum_ld : 1, numer_ld > 0 ? numer_ld : classes, workspace, workspace_bytes, as_stream(stream));
  SmoothParams p{};
  p.f = feats_u_w; p.qf = queue_feats; p.qp = queue_probs;
  p.rows = rows; p.bank_rows = bank_rows; p.D = dim; p.C = classes; p.tau = temperature;
  p.rowsum = rowsum; p.numer = numer;
  p.rowsum_ld = rowsum_ld > 0 ? rowsum_ld : 1; p.numer_ld = numer_ld > 0 ? numer_ld : classes;
  p.nsplit = smooth_nsplit(rows, bank_rows, &p.tiles_per_split);
  const long long row_tiles = (rows + kTM - 1) / kTM;
  p.rows_pad = row_tiles * kTM;
  p.W = (1 + classes + 3) & ~3;
  const size_t need = kWsHeaderBytes + (p.nsplit > 1 ? (size_t)p.nsplit * p.rows_pad * p.W * sizeof(float) : 0);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  if ((size_t)row_tiles * 4 > kWsTicket2Bytes) return fail(B200SSL_E_SHAPE, "%s: too many row tiles", fn);
  p.tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kWsTicketBytes);
  p.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  const size_t smem = ((size_t)(kTM + kTN) * (dim + 1) + kTM * (kTN + 1) + (size_t)kTN * classes) * sizeof(float);
  dim3 grid((unsigned)row_tiles, (unsigned)p.nsplit);
  cudaError_t e = cudaSuccess;
#define LAUNCH_SMOOTH(T, NACC)                                                                               \
  do {                                                                                                       \
    auto k = bank_smooth_simt_kernel<T, NACC>;                                                               \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) k<<<grid, kTileThreads, smem, as_stream(stream)>>>(p);                              \
  } while (0)
  if (dtype == B200SSL_F32) {
    if (classes <= 32) LAUNCH_SMOOTH(float, 8); else LAUNCH_SMOOTH(float, 32);
  } else {
    if (classes <= 32) LAUNCH_SMOOTH(__nv_bfloat16, 8); else LAUNCH_SMOOTH(__nv_bfloat16, 32);
  }
#undef LAUNCH_SMOOTH
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

extern "C" int b200ssl_bank_enqueue(void* queue_feats, void* queue_probs, void* queue_probs_t, const void* feats_u_w,
                                    const void* feats_x,
                                    const float* probs_orig, const int64_t* targets_x, int64_t n_u, int64_t n_x,
                                    int32_t dim, int32_t classes, int32_t dtype, int64_t ptr, int64_t* ptr_state,
                                    int64_t advance, int64_t block_offset, int64_t bank_rows_global,
                                    int64_t shard_begin, int64_t shard_rows, void* stream) {
  const char* fn = "b200ssl_bank_enqueue";
  if (!queue_feats || !queue_probs) return fail(B200SSL_E_NULL, "%s: NULL bank", fn);
  if (n_u < 0 || n_x < 0 || n_u + n_x <= 0) return fail(B200SSL_E_SHAPE, "%s: n_u=%lld n_x=%lld", fn, (long long)n_u, (long long)n_x);
  if ((n_u > 0 && (!feats_u_w || !probs_orig)) || (n_x > 0 && (!feats_x || !targets_x))) return fail(B200SSL_E_NULL, "%s: NULL rows", fn);
  if (dim < 1 || classes < 2) return fail(B200SSL_E_SHAPE, "%s: dim=%d classes=%d", fn, dim, classes);
  if (advance < 0 || (advance && !ptr_state)) return fail(B200SSL_E_ARG, "%s: advance needs ptr_state", fn);
  if (bank_rows_global <= 0 || ptr < 0 || ptr >= bank_rows_global || block_offset < 0 || shard_begin < 0 ||
      shard_rows <= 0 || shard_begin + shard_rows > bank_rows_global)
    return fail(B200SSL_E_ARG, "%s: bad ring geometry (K=%lld ptr=%lld off=%lld shard=[%lld,+%lld))", fn,
                (long long)bank_rows_global, (long long)ptr, (long long)block_offset, (long long)shard_begin, (long long)shard_rows);
  if (queue_probs_t && classes > 32) return fail(B200SSL_E_SHAPE, "%s: transposed bank needs classes <= 32", fn);
  EnqueueParams p{queue_probs_t, queue_feats, queue_probs, feats_u_w, feats_x, probs_orig, reinterpret_cast<const long long*>(targets_x),
                  n_u, n_x, dim, classes, ptr, block_offset, bank_rows_global, shard_begin, shard_rows,
                  reinterpret_cast<long long*>(ptr_state), advance};
  const long long n = n_u + n_x;
  const int grid = (int)((n + 7) / 8);
  if (dtype == B200SSL_F32) bank_enqueue_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(p);
  else if (dtype == B200SSL_BF16) bank_enqueue_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(p);
  else return fail(B200SSL_E_DTYPE, "%s: dtype %d", fn, dtype);
  return check_launch(fn);
}
