// K3 on the 5th-generation tensor cores: memory-smoothing partial sums against the
// bf16 bank (code/comatch.py:180-181), FlashAttention-shaped, never materialising
// A[rows, K]:
//
//   TMA (128B swizzle)        tcgen05.mma kind::f16           tcgen05.ld + MUFU
//   F tile [128 x 64]   -+->  S[128 x 128] = F Qf^T (TMEM) -> E = exp2(S * log2e/tau)
//   Qf tile [128 x 64]  -+                                     rowsum += E ; P = bf16(E) -> smem (swizzled)
//   QpT tile [32 x 128] ---->  numer[128 x 32] += P QpT^T (TMEM accumulator, K = 128 keys)
//
// Warp roles (192 threads): warp 0 = TMEM allocator + TMA producer, warp 1 = MMA
// issuer (one thread), warps 2..5 = epilogue (thread <-> TMEM lane <-> query row).
// S is double-buffered in TMEM so GEMM1 of key tile t+1 overlaps the exp of tile t;
// the bank is split over gridDim.y CTAs and the last CTA of a row tile folds the
// split partials in order (deterministic).
//
// Bank layout for this path: queue_feats [K, 64] bf16 row-major (a row is exactly one
// 128-byte swizzle row) and a transposed, class-padded copy of the probabilities
// queue_probs_t [32, K] bf16 (K-major B operand of the second GEMM), both written by
// b200ssl_bank_enqueue.
#include <math.h>

#include "common.cuh"
#include "tc.cuh"

namespace b200ssl {
namespace tc {

// ---- host: tensor map encode through the driver entry point (no -lcuda) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200SSL_E_ARG, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (row_stride_bytes & 15u))
    return fail(B200SSL_E_ALIGN, "TMA needs a 16-byte aligned base and row pitch (base %p, pitch %llu)", base,
                (unsigned long long)row_stride_bytes);
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200SSL_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tc

namespace {

constexpr int kBM = 128;          // queries per CTA  (UMMA M)
constexpr int kBN = 128;          // keys per tile    (UMMA N of GEMM1 / K of GEMM2)
constexpr int kCP = 32;           // classes, padded  (UMMA N of GEMM2)
constexpr int kStages = 2;
constexpr int kTcThreads = 192;
constexpr uint32_t kTileA = kBM * 128;                  // 16 KB: [128][64] bf16
constexpr uint32_t kTileQf = kBN * 128;                 // 16 KB
constexpr uint32_t kSubQp = kCP * 128;                  //  4 KB: [32][64] bf16
constexpr uint32_t kTileQp = 2 * kSubQp;                //  8 KB
constexpr uint32_t kSubP = kBM * 128;                   // 16 KB: [128][64] bf16
constexpr uint32_t kTileP = 2 * kSubP;                  // 32 KB
constexpr uint32_t kSmemData = kTileA + kStages * (kTileQf + kTileQp) + kTileP;   // 96 KB
constexpr uint32_t kTmemCols = 512;                     // S[0] 0..127, S[1] 128..255, numer 256..287
constexpr size_t kSmemRequest = 120 * 1024;             // > half an SM: one CTA per SM (it owns all TMEM columns)

struct SmoothTcParams {
  long long rows, bank_rows, rows_pad;
  int C, W, nsplit, tiles_per_split;   // W = round_up(1 + C, 4): floats per row of a split partial
  float scale;                    // log2(e) / temperature
  float* rowsum; float* numer;
  float* part; unsigned* tickets;
};

enum { BAR_A = 0, BAR_KV_FULL = 1, BAR_KV_EMPTY = 3, BAR_S_FULL = 5, BAR_S_EMPTY = 7, BAR_P_FULL = 9, BAR_P_EMPTY = 10,
       BAR_ACC = 11, BAR_COUNT = 12 };

__global__ void __launch_bounds__(kTcThreads, 1)
bank_smooth_tc_kernel(const __grid_constant__ CUtensorMap tm_f, const __grid_constant__ CUtensorMap tm_qf,
                      const __grid_constant__ CUtensorMap tm_qpt, const SmoothTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sQf = sA + kTileA;
  uint8_t* sQp = sQf + kStages * kTileQf;
  uint8_t* sP = sQp + kStages * kTileQp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTileP);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x, split = blockIdx.y;
  const long long nktiles = (p.bank_rows + kBN - 1) / kBN;
  const long long kt0 = (long long)split * p.tiles_per_split;
  const int T = (int)(min(nktiles, kt0 + p.tiles_per_split) - kt0);

  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[BAR_A], 1);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bars[BAR_KV_FULL + s], 1);
      tc::mbar_init(&bars[BAR_KV_EMPTY + s], 1);
      tc::mbar_init(&bars[BAR_S_FULL + s], 1);
      tc::mbar_init(&bars[BAR_S_EMPTY + s], 128);
    }
    tc::mbar_init(&bars[BAR_P_FULL], 128);
    tc::mbar_init(&bars[BAR_P_EMPTY], 1);
    tc::mbar_init(&bars[BAR_ACC], 1);
    *abort_flag = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
  if (warp == 1 && lane == 0) {
    tc::tma_prefetch_desc(&tm_f);
    tc::tma_prefetch_desc(&tm_qf);
    tc::tma_prefetch_desc(&tm_qpt);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&bars[BAR_A], kTileA);
      tc::tma_load_2d(sA, &tm_f, 0, row_tile * kBM, &bars[BAR_A]);
      for (int t = 0; t < T; ++t) {
        const int s = t % kStages;
        if (t >= kStages) tc::mbar_wait(&bars[BAR_KV_EMPTY + s], ((t / kStages) - 1) & 1, abort_flag);
        const int key0 = (int)((kt0 + t) * kBN);
        tc::mbar_arrive_expect_tx(&bars[BAR_KV_FULL + s], kTileQf + kTileQp);
        tc::tma_load_2d(sQf + s * kTileQf, &tm_qf, 0, key0, &bars[BAR_KV_FULL + s]);
        tc::tma_load_2d(sQp + s * kTileQp, &tm_qpt, key0, 0, &bars[BAR_KV_FULL + s]);
        tc::tma_load_2d(sQp + s * kTileQp + kSubQp, &tm_qpt, key0 + 64, 0, &bars[BAR_KV_FULL + s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = tc::idesc_bf16_f32(kBM, kBN);
      constexpr uint32_t idesc2 = tc::idesc_bf16_f32(kBM, kCP);
      const uint64_t a_desc = tc::smem_desc_sw128(tc::smem_u32(sA));
      tc::mbar_wait(&bars[BAR_A], 0, abort_flag);
      auto gemm1 = [&](int t) {       // S[b] = F Qf^T   (K = 64 -> 4 x UMMA_K 16)
        const int s = t % kStages, b = t & 1;
        tc::mbar_wait(&bars[BAR_KV_FULL + s], (t / kStages) & 1, abort_flag);
        if (t >= 2) tc::mbar_wait(&bars[BAR_S_EMPTY + b], ((t >> 1) - 1) & 1, abort_flag);
        tc::tcgen05_fence_after();
        const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(sQf + s * kTileQf));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem + b * kBN, a_desc + 2 * k, b_desc + 2 * k, idesc1, k > 0);
        tc::mma_commit(&bars[BAR_S_FULL + b]);
      };
      gemm1(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) gemm1(t + 1);
        const int s = t % kStages;
        tc::mbar_wait(&bars[BAR_P_FULL], t & 1, abort_flag);
        tc::tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {   // numer += P QpT^T   (K = 128 keys -> 2 sub-tiles x 4 x UMMA_K 16)
          const uint64_t pa = tc::smem_desc_sw128(tc::smem_u32(sP + kb * kSubP));
          const uint64_t qb = tc::smem_desc_sw128(tc::smem_u32(sQp + s * kTileQp + kb * kSubQp));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem + 2 * kBN, pa + 2 * k, qb + 2 * k, idesc2, (t | kb | k) != 0);
        }
        tc::mma_commit(&bars[BAR_KV_EMPTY + s]);
        tc::mma_commit(&bars[BAR_P_EMPTY]);
      }
      tc::mma_commit(&bars[BAR_ACC]);
    }
  } else {
    // ================= epilogue: thread <-> TMEM lane <-> query row =================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are visible to this warp
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float rowsum = 0.f;
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      tc::mbar_wait(&bars[BAR_S_FULL + b], (t >> 1) & 1, abort_flag);
      if (t >= 1) tc::mbar_wait(&bars[BAR_P_EMPTY], (t - 1) & 1, abort_flag);
      tc::tcgen05_fence_after();
      const long long key0 = (kt0 + t) * kBN;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[32];
        tc::tmem_ld_32x32(lane_addr + b * kBN + c4 * 32, r);
        tc::tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = (key0 + c4 * 32 + j < p.bank_rows) ? exp2f(__uint_as_float(r[j]) * p.scale) : 0.f;
          const float e1 = (key0 + c4 * 32 + j + 1 < p.bank_rows) ? exp2f(__uint_as_float(r[j + 1]) * p.scale) : 0.f;
          const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
          rowsum += __low2float(h) + __high2float(h);     // the weights GEMM2 actually uses
          packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int sub = c4 >> 1, chunk = (c4 & 1) * 4 + q;
          uint4 v = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
          *reinterpret_cast<uint4*>(sP + sub * kSubP + tc::sw128_offset(r_in, chunk)) = v;
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars[BAR_S_EMPTY + b]);
      tc::fence_proxy_async_smem();               // generic-proxy writes of P -> visible to the tensor core
      tc::mbar_arrive(&bars[BAR_P_FULL]);
    }
    tc::mbar_wait(&bars[BAR_ACC], 0, abort_flag);
    tc::tcgen05_fence_after();
    uint32_t r[32];
    tc::tmem_ld_32x32(lane_addr + 2 * kBN, r);
    tc::tmem_ld_wait();
    const long long grow = (long long)row_tile * kBM + r_in;
    if (p.nsplit == 1) {
      if (grow < p.rows) {
        p.rowsum[grow] = rowsum;
#pragma unroll
        for (int c = 0; c < kCP; ++c)
          if (c < p.C) p.numer[grow * p.C + c] = __uint_as_float(r[c]);
      }
    } else {
      // split partial, row-interleaved [split][rows_pad][W]: element 0 = rowsum, 1..C = numer
      float4* o = reinterpret_cast<float4*>(p.part + ((size_t)split * p.rows_pad + grow) * p.W);
      float vals[kCP + 4];
      vals[0] = rowsum;
#pragma unroll
      for (int c = 0; c < kCP; ++c) vals[1 + c] = __uint_as_float(r[c]);
#pragma unroll
      for (int q = 0; q < (kCP + 4) / 4; ++q)
        if (4 * q < p.W) o[q] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
    }
    tc::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem, kTmemCols);
  }
  if (p.nsplit == 1) return;
  // last split CTA of this row tile folds the partials in split order (deterministic)
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.tickets[row_tile], 1u) == (unsigned)p.nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const long long i0 = (long long)row_tile * kBM;
  const int mrows = (int)min((long long)kBM, p.rows - i0);
  const int W = p.W, C = p.C;
  fold_splits_vec4(reinterpret_cast<const float4*>(p.part + (size_t)i0 * W), (size_t)p.rows_pad * W / 4, p.nsplit,
                   mrows * W / 4, [&](int i, float4 v) {
                     const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                     for (int j = 0; j < 4; ++j) {
                       const int e = 4 * i + j, row = e / W, col = e - row * W;
                       if (col == 0) p.rowsum[i0 + row] = vv[j];
                       else if (col <= C) p.numer[(i0 + row) * C + col - 1] = vv[j];
                     }
                   });
  if (threadIdx.x == 0) p.tickets[row_tile] = 0u;
}

}  // namespace

int smooth_tc_nsplit(long long rows, long long bank_rows, int* tiles_per_split) {
  const long long row_tiles = (rows + kBM - 1) / kBM;
  const long long ktiles = (bank_rows + kBN - 1) / kBN;
  long long want = (kNumSMs + row_tiles - 1) / row_tiles;     // one CTA per SM
  if (want < 1) want = 1;
  if (want > ktiles) want = ktiles;
  const long long tps = (ktiles + want - 1) / want;
  if (tiles_per_split) *tiles_per_split = (int)tps;
  return (int)((ktiles + tps - 1) / tps);
}

size_t smooth_tc_workspace_floats(long long rows, long long bank_rows, int classes) {
  const long long row_tiles = (rows + kBM - 1) / kBM;
  const int ns = smooth_tc_nsplit(rows, bank_rows, nullptr);
  return ns > 1 ? (size_t)ns * row_tiles * kBM * ((1 + classes + 3) & ~3) : 0;
}

// bf16, dim 64, classes <= 32, bank rows a multiple of 8: the tensor-core path.
int bank_smooth_tc(const void* feats, const void* queue_feats, const void* queue_probs_t, long long rows,
                   long long bank_rows, int classes, float temperature, float* rowsum, float* numer, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_bank_smooth_partial[tcgen05]";
  SmoothTcParams p{};
  p.rows = rows; p.bank_rows = bank_rows; p.C = classes; p.W = (1 + classes + 3) & ~3;
  p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.rowsum = rowsum; p.numer = numer;
  p.nsplit = smooth_tc_nsplit(rows, bank_rows, &p.tiles_per_split);
  const long long row_tiles = (rows + kBM - 1) / kBM;
  p.rows_pad = row_tiles * kBM;
  const size_t need = kWsHeaderBytes + sizeof(float) * smooth_tc_workspace_floats(rows, bank_rows, classes);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  if ((size_t)row_tiles * 4 > kWsTicket2Bytes) return fail(B200SSL_E_SHAPE, "%s: too many row tiles", fn);
  p.tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kWsTicketBytes);
  p.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  CUtensorMap tm_f, tm_qf, tm_qpt;
  if (int e = tc::make_tmap_bf16_2d(&tm_f, feats, (uint64_t)rows, 64, 128, kBM, 64)) return e;
  if (int e = tc::make_tmap_bf16_2d(&tm_qf, queue_feats, (uint64_t)bank_rows, 64, 128, kBN, 64)) return e;
  if (int e = tc::make_tmap_bf16_2d(&tm_qpt, queue_probs_t, kCP, (uint64_t)bank_rows, (uint64_t)bank_rows * 2, kCP, 64)) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bank_smooth_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemRequest);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
    attr_set = true;
  }
  static_assert(kSmemData + 1024 + 256 <= kSmemRequest, "shared memory budget");
  dim3 grid((unsigned)row_tiles, (unsigned)p.nsplit);
  bank_smooth_tc_kernel<<<grid, kTcThreads, kSmemRequest, stream>>>(tm_f, tm_qf, tm_qpt, p);
  return check_launch(fn);
}

}  // namespace b200ssl
