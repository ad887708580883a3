// K3 on the 5th-generation tensor cores: memory-smoothing partial sums against the bank
// (code/comatch.py:180-181), FlashAttention-shaped, never materialising A[rows, K]:
//
//   TMA (128B swizzle)        tcgen05.mma kind::f16           tcgen05.ld + FMUL2 + MUFU
//   F tile [128 x 64]   -+->  S[128 x 128] = F Qf^T (TMEM) -> E = exp2(S * log2e/tau)
//   Qf tile [128 x 64]  -+                                     P = bf16(E) -> back into TMEM (tcgen05.st)
//   QpT tile [32 x 128] ---->  [numer | rowsum][128 x 32] += P [QpT | 1]^T   (A = P read from TMEM, K = 128 keys)
//
// Warp roles (576 threads): warp 0 = TMEM allocator + TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue (4 threads
// per TMEM lane = query row, 32 key columns each).  S and P are double-buffered in tensor memory so GEMM1 of unit j+1 and
// GEMM2 of unit j-1 overlap the exponentials of unit j; the bank is split over a cluster of CTAs whose partials are folded
// through DSMEM in rank order (deterministic).
//
// What bounds it (profiles/r02_k3_experiments.md, tools/micro/): a unit is 12 SMALL MMAs (628 clocks of tensor pipe, 43-88 each)
// and 16384 exponentials (1096 clocks of MUFU).  The MMA warp therefore runs its loop in warp-uniform control flow and issues
// from one elected lane -- inside `if (lane == 0)` the compiler rebuilds every tcgen05.mma operand with an ELECT / R2UR sequence,
// ~17 dependent instructions per MMA, and THAT bounded the round-1 kernel, not the exponentials.  Keeping P in tensor memory
// halves the shared-memory traffic of a unit and removes the st.shared + fence.proxy.async of every epilogue thread.  Measured
// dead ends: an FMA-pipe polynomial for part of the exponentials (pays nothing until the kernel is MUFU-bound; 8 of 32 slots are
// a wash now), two epilogue groups on alternate tiles, issuing the TMEM read of the next unit early (twice).
//
// fp32 storage (NT = 2): the same kernel on bf16 hi + mid operands made by split_operands_kernel (three cross-term MMA groups
// per GEMM, P_hi / P_mid in TMEM, unit accumulators summed in registers) -- 1e-6 of fp64, see the kernel.
//
// Bank layout for this path: queue_feats [K, 64] bf16 row-major (a row is exactly one 128-byte swizzle row) and a transposed,
// class-padded copy of the probabilities queue_probs_t [32, K] bf16 (K-major B operand of the second GEMM, row C = ones ->
// row sums), both written by b200ssl_bank_enqueue.
#include <math.h>

#include <cooperative_groups.h>

#include "common.cuh"
#include "peer.cuh"
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
  // The driver API needs a context that is current in the CALLING thread.  autograd runs backward on its
  // own worker thread, where our launch can be the first CUDA call: bind the primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
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

namespace cg = cooperative_groups;

constexpr int kBM = 128;          // queries per row tile (UMMA M)
constexpr int kBN = 128;          // keys per tile    (UMMA N of GEMM1 / K of GEMM2)
constexpr int kCP = 32;           // classes + the "ones" column, padded (UMMA N of GEMM2)
constexpr int kMaxStages = 6;      // key tiles in flight: 6 with P in tensor memory; with P in shared memory 5 while they fit (mt <= 2), else 4
                                  // three tiles can be in use while the next ones load (remote shards: NVLink latency)
constexpr int kMaxMT = 4;         // row tiles one CTA can serve from ONE staged key tile ("row loop"): a remote key tile
                                  // then crosses NVLink once per step instead of once per row tile
constexpr int kEpiWarps = 16;     // 4 per TMEM lane quarter: each thread owns 32 of the 128 key columns of its row
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTcThreads = 64 + kEpiThreads;   // warp 0 TMA + TMEM alloc, warp 1 MMA, warps 2.. epilogue
constexpr int kMaxCluster = 8;
constexpr int kMaxSeg = 8;        // bank segments = shards of a rank-sharded bank (1 = the whole bank is local)
constexpr uint32_t kTileA = kBM * 128;                  // 16 KB: [128][64] bf16
constexpr uint32_t kTileQf = kBN * 128;                 // 16 KB
constexpr uint32_t kSubQp = kCP * 128;                  //  4 KB: [32][64] bf16
constexpr uint32_t kTileQp = 2 * kSubQp;                //  8 KB
constexpr uint32_t kSubP = kBM * 128;                   // 16 KB: [128][64] bf16
constexpr uint32_t kTileP = 2 * kSubP;                  // 32 KB
constexpr int stages_for(int mt, bool ptmem = true) { return ptmem ? kMaxStages : mt <= 2 ? 5 : 4; }   // the most that fit next to mt query tiles
constexpr uint32_t smem_stages_n(int stages, bool ptmem = true) { return stages * (kTileQf + kTileQp) + (ptmem ? 0 : 2 * kTileP); }   // P double buffered
constexpr uint32_t smem_stages(int mt, bool ptmem = true) { return smem_stages_n(stages_for(mt, ptmem), ptmem); }   // 144 KB (P in TMEM) / 184 / 160 KB
constexpr uint32_t kTmemCols = 512;                     // S[0] 0..127, S[1] 128..255, [numer | rowsum] of row tile m at 256 + 32 m,
constexpr uint32_t kTmemP = 384;                        // bf16 P[0] 384..447, P[1] 448..511 (A operand of GEMM2 in tensor memory)
constexpr int kRedLd = 36;                              // floats per row of a reduction tile (16-byte rows, 4-way bank spread)
constexpr uint32_t kRedTile = kBM * kRedLd * 4;         // 18 KB per row tile, staged over the drained pipeline buffers
constexpr size_t smem_request_n(int mt, int stages, bool ptmem = true) { return 1024 + (size_t)mt * kTileA + smem_stages_n(stages, ptmem) + 512; }
constexpr size_t smem_request(int mt, bool ptmem = true) { return smem_request_n(mt, stages_for(mt, ptmem), ptmem); }   // P in TMEM: 161.5 .. 209.5 KB
constexpr size_t kSmemMax = smem_request(kMaxMT, false) > smem_request(kMaxMT, true) ? smem_request(kMaxMT, false) : smem_request(kMaxMT, true);

struct BankMaps {                 // one pair of tensor maps per shard; remote shards are peer-mapped NVLink addresses
  CUtensorMap qf[kMaxSeg];
  CUtensorMap qpt[kMaxSeg];
};

struct SmoothTcParams {
  long long rows, rows_pad;
  int row_tiles;                  // ceil(rows / 128)
  int mt;                         // row tiles per CTA
  int stages;                     // key tiles in flight (3..5)
  int nseg, tps;                  // key tiles are enumerated shard by shard: tile kt -> (kt / tps, kt % tps)
  uint8_t* const* arenas;         // non-NULL: directly addressed sharded bank (peer.cuh flags / epochs)
  int rank, world, seg_first;     // seg_first: segment the tile enumeration starts at (the own shard)
  int C, W;                       // W = round_up(C + 1, 4): [numer 0..C-1, rowsum] per row of a partial
  int nsplit, cluster, nouter;    // nsplit = cluster * nouter CTAs share one group of row tiles
  float scale;                    // log2(e) / temperature
  float s_min;                    // lower clamp of S for the polynomial exponentials (-126 / scale)
  float* rowsum; float* numer; int rowsum_ld, numer_ld;
  float* part; unsigned* tickets;
  unsigned long long* dbg;
};

enum { BAR_A = 0, BAR_KV_FULL = 1, BAR_KV_EMPTY = BAR_KV_FULL + kMaxStages, BAR_S_FULL = BAR_KV_EMPTY + kMaxStages,
       BAR_S_EMPTY = BAR_S_FULL + 2, BAR_P_FULL = BAR_S_EMPTY + 2, BAR_P_EMPTY = BAR_P_FULL + 2,
       BAR_ACC = BAR_P_EMPTY + 2, BAR_COUNT };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^(s * scale) on the FMA pipe (the MUFU unit makes 16 exponentials per clock and SM, a quarter of what the two GEMMs of a
// key tile could consume): round-to-nearest split x = j + f through the 1.5 * 2^23 trick, degree-3 minimax polynomial of
// 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding of P that follows), exponent patched in with one
// integer shift-add.  s is clamped from below so that j stays inside the exponent range (the reference's exp underflows
// there); above the range the reference's own fp32 exp overflows too.
__device__ __forceinline__ float ex2_poly(float s, float scale, float s_min) {
  constexpr float kMagic = 12582912.f;                      // 1.5 * 2^23
  s = fmaxf(s, s_min);
  const float t = fmaf(s, scale, kMagic);                   // low mantissa bits = round(x)
  const float f = fmaf(s, scale, -(t - kMagic));            // x - round(x)
  float p = fmaf(0x1.c3f76p-5f, f, 0x1.f0de1ap-3f);
  p = fmaf(p, f, 0x1.62f31ap-1f);
  p = fmaf(p, f, 0x1.fff692p-1f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));    // (magic << 23) == 0 (mod 2^32)
}

// NPOLY of the 32 exponentials a thread makes per S tile go through ex2_poly, spread evenly between the MUFU ones.
template <int NPOLY>
__device__ __forceinline__ constexpr bool poly_slot(int i) { return NPOLY > 0 && ((i + 1) * NPOLY) / 32 != (i * NPOLY) / 32; }

// PTMEM: P = bf16(exp) goes back into tensor memory and is the A operand of the second GEMM (tcgen05.mma with A in TMEM);
// false = the round-2 A/B form, P through swizzled shared memory (st.shared + fence.proxy.async, SS-form MMA).
// NT: bf16 terms per operand.  1 = bf16 storage.  2 = fp32 storage split on the fly into bf16 hi + mid (csrc/bank_tc.cu:
// split_operands_kernel): S = F_mid Q_hi^T + F_hi Q_mid^T + F_hi Q_hi^T, and likewise [numer | rowsum] from P = P_hi + P_mid
// (both in tensor memory, single buffered) against Qp_hi / Qp_mid -- 2^-17 per operand, 1e-6 on the smoothed probabilities
// (the reference's fp32 torch.mm: 3e-7), at 36 instead of 12 MMAs per unit.
template <int NPOLY, bool PTMEM, int NT = 1>
__global__ void __launch_bounds__(kTcThreads, 1)
bank_smooth_tc_kernel(const __grid_constant__ CUtensorMap tm_f, const __grid_constant__ BankMaps maps, const SmoothTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                       // [mt][128][64] bf16 query tiles
  const int kStages = p.stages;
  static_assert(NT == 1 || (NT == 2 && PTMEM && NPOLY == 0), "split operands: P in tensor memory, MUFU exponentials");
  constexpr uint32_t tA = NT * kTileA, tQf = NT * kTileQf, tQp = NT * kTileQp;    // all terms of one tile, term-major
  uint8_t* sQf = sA + (size_t)p.mt * tA;
  uint8_t* sQp = sQf + kStages * tQf;
  uint8_t* sP = sQp + kStages * tQp;
  float* sRed = reinterpret_cast<float*>(sQf);             // [mt][128][kRedLd] after the pipeline has drained
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (PTMEM ? 0 : 2 * kTileP));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int row_group = blockIdx.x, split = blockIdx.y;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int tile0 = row_group * p.mt;                      // first row tile of this CTA
  const int M = min(p.mt, p.row_tiles - tile0);            // row tiles it serves (>= 1)
  if (threadIdx.x == 0) {
    B200SSL_STAMP(p.dbg, cta, 0);
    B200SSL_STAMP_NS(p.dbg, cta, 10);
  }
  pdl_launch_dependents();                                 // the next kernel may start its prologue
  const long long nktiles = (long long)p.nseg * p.tps;
  // Key tiles are dealt round-robin to the splits: tile u = split + t*nsplit of the enumeration that starts at the
  // OWN shard, so every CTA begins with local tiles while its remote ones are already in flight.
  const int T = (int)((nktiles - split + p.nsplit - 1) / p.nsplit);      // >= 1: nsplit <= nktiles
  const int J = T * M;                                                   // (key tile, row tile) units, row tile fastest
  auto tile_of = [&](int t, int* seg) {                                  // -> first bank row of the tile inside its shard
    const long long u = split + (long long)t * p.nsplit;
    const int si = (int)(u / p.tps);
    *seg = (p.seg_first + si) % p.nseg;
    return (int)(u - (long long)si * p.tps) * kBN;
  };

  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[BAR_A], 1);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bars[BAR_KV_FULL + s], 1);
      tc::mbar_init(&bars[BAR_KV_EMPTY + s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bars[BAR_S_FULL + s], 1);
      tc::mbar_init(&bars[BAR_S_EMPTY + s], kEpiWarps);        // one arrival per epilogue warp
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bars[BAR_P_FULL + s], kEpiWarps);
      tc::mbar_init(&bars[BAR_P_EMPTY + s], 1);
    }
    tc::mbar_init(&bars[BAR_ACC], 1);
    *abort_flag = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
  if (warp == 1 && lane == 0) {
    tc::tma_prefetch_desc(&tm_f);
    int seg0;
    tile_of(0, &seg0);
    tc::tma_prefetch_desc(&maps.qf[seg0]);
    tc::tma_prefetch_desc(&maps.qpt[seg0]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                              // previous kernel complete: global memory may be touched now
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, cta, 1);     // setup (barrier init, TMEM alloc) done

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&bars[BAR_A], (uint32_t)M * tA);
      for (int m = 0; m < M; ++m)
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) tc::tma_load_2d(sA + (size_t)m * tA + tt * kTileA, &tm_f, 64 * tt, (tile0 + m) * kBM, &bars[BAR_A]);
    }
    if (p.arenas) {
      // multi-rank bank: every rank's enqueue of the previous step must have landed (its flag was published a
      // contrastive forward + backward + EMA ago, so this normally falls through).  One lane per peer: the
      // acquire loads overlap instead of costing one L2 round trip each.
      uint8_t* mine = p.arenas[p.rank];
      peer::LocalCtl* ctl = peer::local_ctl(mine);
      const unsigned long long need = *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[peer::kXSmoothDone]);
      if (lane < p.world && lane != p.rank) peer::wait_flag(peer::flag_of(mine, peer::kXEnqueueDone, lane), need, ctl);
      __syncwarp();
    }
    if (lane == 0) {
      for (int t = 0, s = 0, lap = 0; t < T; ++t) {
        if (t >= kStages) tc::mbar_wait(&bars[BAR_KV_EMPTY + s], (lap - 1) & 1, abort_flag);
        int seg;
        const int key0 = tile_of(t, &seg);
        tc::mbar_arrive_expect_tx(&bars[BAR_KV_FULL + s], tQf + tQp);
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {                   // term tt: columns [64 tt, +64) of a key row, rows [32 tt, +32) of QpT
          tc::tma_load_2d(sQf + s * tQf + tt * kTileQf, &maps.qf[seg], 64 * tt, key0, &bars[BAR_KV_FULL + s]);
          tc::tma_load_2d(sQp + s * tQp + tt * kTileQp, &maps.qpt[seg], key0, 32 * tt, &bars[BAR_KV_FULL + s]);
          tc::tma_load_2d(sQp + s * tQp + tt * kTileQp + kSubQp, &maps.qpt[seg], key0 + 64, 32 * tt, &bars[BAR_KV_FULL + s]);
        }
        if (++s == kStages) { s = 0; ++lap; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp runs the loop (warp-uniform control flow), one elected lane issues =========
    // The issue path of this warp bounds the kernel: a unit is 12 small MMAs (64 / 16 clocks of tensor work each) plus
    // three commits, so every instruction between two tcgen05.mma counts (profiles/r02_k3_experiments.md).
    {
      const bool leader = tc::elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      constexpr uint32_t idesc1 = tc::idesc_bf16_f32(kBM, kBN);
      constexpr uint32_t idesc2 = tc::idesc_bf16_f32(kBM, kCP);
      constexpr int kPairs = NT == 2 ? 3 : 1;
      const uint32_t a_u32 = tc::smem_u32(sA), qf_u32 = tc::smem_u32(sQf), qp_u32 = tc::smem_u32(sQp), p_u32 = tc::smem_u32(sP);
      tc::mbar_wait(&bars[BAR_A], 0, abort_flag);
      if (threadIdx.x == 32) B200SSL_STAMP(p.dbg, cta, 2);  // query tiles landed (TMA)
      // unit j = (key tile t, row tile m), m fastest: S[j & 1] = F_m Qf_t^T   (K = 64 -> 4 x UMMA_K 16).
      // (t, m) and the stage index are carried as counters -- a runtime division per GEMM cost 17 % at rows 3584 x K 65536.
      auto gemm1 = [&](int j, int m, int s, int lap) {
        const int b = j & 1;
        if (m == 0) tc::mbar_wait(&bars[BAR_KV_FULL + s], lap, abort_flag);
        if (j >= 2) tc::mbar_wait(&bars[BAR_S_EMPTY + b], ((j >> 1) - 1) & 1, abort_flag);   // S of unit j-2 is in registers
        tc::tcgen05_fence_after();
        const uint64_t a_desc = tc::smem_desc_sw128(a_u32 + (uint32_t)m * tA);
        const uint64_t b_desc = tc::smem_desc_sw128(qf_u32 + (uint32_t)s * tQf);
        if (leader) {
#pragma unroll
          for (int pr = 0; pr < kPairs; ++pr) {             // (term of F, term of Q): (mid, hi), (hi, mid), (hi, hi) -- small products first
            const uint64_t ta = (NT == 2 && pr == 0) ? (kTileA >> 4) : 0, tb = (NT == 2 && pr == 1) ? (kTileQf >> 4) : 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem_u + b * kBN, a_desc + ta + 2 * k, b_desc + tb + 2 * k, idesc1, (pr | k) != 0);
          }
          tc::mma_commit(&bars[BAR_S_FULL + b]);
        }
        __syncwarp();
      };
      auto gemm2 = [&](int j, int t, int m, int s) {   // [numer | rowsum]_m += P [QpT | 1]^T   (K = 128 keys -> 2 sub-tiles x 4 x UMMA_K 16)
        const int pb = NT == 2 ? 0 : (j & 1);               // split operands: one P buffer holding [P_hi | P_mid]
        tc::mbar_wait(&bars[BAR_P_FULL + pb], (NT == 2 ? j : (j >> 1)) & 1, abort_flag);
        tc::tcgen05_fence_after();
        const uint64_t pa0 = tc::smem_desc_sw128(p_u32 + (uint32_t)pb * kTileP);
        const uint64_t qb0 = tc::smem_desc_sw128(qp_u32 + (uint32_t)s * tQp);
        if (leader) {
#pragma unroll
          for (int pr = 0; pr < kPairs; ++pr) {             // (term of P, term of Qp): (mid, hi), (hi, mid), (hi, hi)
            const int ta = (NT == 2 && pr == 0) ? 1 : 0;
            const uint64_t tb = (NT == 2 && pr == 1) ? (kTileQp >> 4) : 0;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (PTMEM)
                  tc::mma_bf16_ts(tmem_u + 2 * kBN + m * kCP, tmem_u + kTmemP + (NT == 2 ? ta : pb) * (kBN / 2) + (kb * 4 + k) * 8,
                                  qb0 + tb + (uint64_t)(kb * (kSubQp >> 4) + 2 * k), idesc2, ((NT == 2 ? 0 : t) | pr | kb | k) != 0);
                else
                  tc::mma_bf16_ss(tmem_u + 2 * kBN + m * kCP, pa0 + (uint64_t)(kb * (kSubP >> 4) + 2 * k), qb0 + (uint64_t)(kb * (kSubQp >> 4) + 2 * k),
                                  idesc2, (t | kb | k) != 0);
              }
            }
          }
          if (m == M - 1) tc::mma_commit(&bars[BAR_KV_EMPTY + s]);   // the key tile has served every row tile
          tc::mma_commit(&bars[BAR_P_EMPTY + pb]);
        }
        __syncwarp();
      };
      int m1 = 0, s1 = 0, ph1 = 0;                           // (row tile, stage, stage-ring lap) of the next GEMM1
      auto next1 = [&]() { if (++m1 == M) { m1 = 0; if (++s1 == kStages) { s1 = 0; ph1 ^= 1; } } };
      int t2 = 0, m2 = 0, s2 = 0;                            // ... of the next GEMM2
      gemm1(0, m1, s1, ph1);
      next1();
      for (int j = 0; j < J; ++j) {
        if (j + 1 < J) {                                     // S is double buffered: GEMM1 of unit j+1 overlaps the exp of unit j
          gemm1(j + 1, m1, s1, ph1);
          next1();
        }
        gemm2(j, t2, m2, s2);
        if (++m2 == M) { m2 = 0; ++t2; if (++s2 == kStages) s2 = 0; }
      }
      if (leader) tc::mma_commit(&bars[BAR_ACC]);
      __syncwarp();
    }
  } else {
    // ===== epilogue: four threads per TMEM lane (query row), 32 of the 128 key columns each =====
    const int quarter = warp & 3, colq = (warp - 2) >> 2;   // TMEM lanes [32*quarter, +32) are visible to this warp
    const int half = colq >> 1, c2 = colq & 1;              // P sub-tile (64 keys) and 32-column group inside it
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    // shared-window addresses of this thread's four 16-byte P chunks (loop invariant; st.shared, not generic stores)
    const uint32_t p_base = tc::smem_u32(sP) + half * kSubP;
    uint32_t p_off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) p_off[q] = p_base + tc::sw128_offset(r_in, c2 * 4 + q);
    if (threadIdx.x == 64) {                                // first S tile ready (TMA + GEMM1): stamped outside the unit loop
      tc::mbar_wait(&bars[BAR_S_FULL], 0, abort_flag);
      B200SSL_STAMP(p.dbg, cta, 3);
    }
    __syncwarp();
    uint32_t r[32];
    uint32_t w[16], wm[16];                                 // packed bf16 pairs of P (hi; split operands: + mid)
    // Split operands (1e-5 parity bar): the tensor core adds into its fp32 accumulator with truncation, a bias that grows with
    // the length of the chain (measured 8e-5 after 100 units).  There every unit starts a fresh accumulator (24 MMAs) and the
    // unit tiles are summed here in registers, round-to-nearest: 8 of the 32 [numer | rowsum] columns of the row per thread.
    float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto exp_pack = [&](int e) {                                                                    // comatch.py:180
      const float s0 = __uint_as_float(r[2 * e]), s1 = __uint_as_float(r[2 * e + 1]);
      float e0, e1;
      if (poly_slot<NPOLY>(2 * e) || poly_slot<NPOLY>(2 * e + 1)) {
        e0 = poly_slot<NPOLY>(2 * e) ? ex2_poly(s0, p.scale, p.s_min) : ex2_approx(s0 * p.scale);
        e1 = poly_slot<NPOLY>(2 * e + 1) ? ex2_poly(s1, p.scale, p.s_min) : ex2_approx(s1 * p.scale);
      } else {
        float x0, x1;                                       // one packed multiply for the pair (bit-identical to two FMULs)
        tc::mul_f32x2(x0, x1, s0, s1, p.scale);
        e0 = ex2_approx(x0);
        e1 = ex2_approx(x1);
      }
      const __nv_bfloat162 hh = __floats2bfloat162_rn(e0, e1);
      w[e] = *reinterpret_cast<const uint32_t*>(&hh);
      if (NT == 2) {                                        // what bf16 dropped, as a second bf16: P = P_hi + P_mid to 2^-17
        const __nv_bfloat162 mm = __floats2bfloat162_rn(e0 - __low2float(hh), e1 - __high2float(hh));
        wm[e] = *reinterpret_cast<const uint32_t*>(&mm);
      }
    };
    for (int j = 0; j < J; ++j) {
      const int b = j & 1;
      // keys beyond the bank need no mask: their QpT columns (incl. the ones column) are TMA zero fill.
      // The S read is split in two 16-column loads: the second is in flight during the exponentials of the first.
      // (issuing the first half of unit j+1 ahead of the P store of unit j was measured 15 % SLOWER, twice: r02_k3_experiments.md)
      tc::mbar_wait(&bars[BAR_S_FULL + b], (j >> 1) & 1, abort_flag);
      tc::tcgen05_fence_after();
      tc::tmem_ld_32x16<0>(lane_addr + b * kBN + colq * 32, r);
      tc::tmem_ld_wait(r);
      tc::tmem_ld_32x16<16>(lane_addr + b * kBN + colq * 32 + 16, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) exp_pack(e);
      tc::tmem_ld_wait(r);
      tc::tcgen05_fence_before();
      tc::mbar_arrive_warp(&bars[BAR_S_EMPTY + b], lane);   // S[b] is in registers: GEMM1 of unit j+2 may overwrite it
#pragma unroll
      for (int e = 8; e < 16; ++e) exp_pack(e);
      // the P buffer this unit writes must have been consumed (unit j-2; split operands: j-1) -- only now, after the exponentials
      const int pb = NT == 2 ? 0 : b;
      if (NT == 2 ? j >= 1 : j >= 2) tc::mbar_wait(&bars[BAR_P_EMPTY + pb], (NT == 2 ? j - 1 : (j >> 1) - 1) & 1, abort_flag);
      if (NT == 2 && j >= 1) {                              // GEMM2 of unit j-1 has retired: bank its [numer | rowsum] tile (see acc8)
        uint32_t t8[8];
        tc::tcgen05_fence_after();
        tc::tmem_ld_32x8(lane_addr + 2 * kBN + colq * 8, t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc8[i] += __uint_as_float(t8[i]);
      }
      if (PTMEM) {
        tc::tcgen05_fence_after();
        tc::tmem_st_32x16(lane_addr + kTmemP + (NT == 2 ? 0 : b) * (kBN / 2) + colq * 16, w);
        if (NT == 2) tc::tmem_st_32x16(lane_addr + kTmemP + (kBN / 2) + colq * 16, wm);
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) tc::st_shared_v4(p_off[q] + b * kTileP, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        tc::fence_proxy_async_smem();               // generic-proxy writes of P -> visible to the tensor core
      }
      tc::mbar_arrive_warp(&bars[BAR_P_FULL + pb], lane);
    }
    if (threadIdx.x == 64) B200SSL_STAMP(p.dbg, cta, 4);   // last exp tile done
    tc::mbar_wait(&bars[BAR_ACC], 0, abort_flag);           // all MMAs retired: pipeline smem is free, accumulators final
    if (threadIdx.x == 64) B200SSL_STAMP(p.dbg, cta, 5);
    tc::tcgen05_fence_after();
    if (NT == 2) {                                          // last unit's tile, then this thread's 8 columns of the (single) row tile
      uint32_t t8[8];
      tc::tmem_ld_32x8(lane_addr + 2 * kBN + colq * 8, t8);
      float4* dst = reinterpret_cast<float4*>(sRed + (size_t)r_in * kRedLd + colq * 8);
      dst[0] = make_float4(acc8[0] + __uint_as_float(t8[0]), acc8[1] + __uint_as_float(t8[1]), acc8[2] + __uint_as_float(t8[2]),
                           acc8[3] + __uint_as_float(t8[3]));
      dst[1] = make_float4(acc8[4] + __uint_as_float(t8[4]), acc8[5] + __uint_as_float(t8[5]), acc8[6] + __uint_as_float(t8[6]),
                           acc8[7] + __uint_as_float(t8[7]));
    } else if (colq < M) {                                  // warp group colq stages row tile colq: [numer | rowsum] -> fp32 tile
      uint32_t r[32];
      tc::tmem_ld_32x32(lane_addr + 2 * kBN + colq * kCP, r);
      tc::tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(sRed + ((size_t)colq * kBM + r_in) * kRedLd);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                             __uint_as_float(r[4 * q + 3]));
    }
    tc::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem, kTmemCols);
  }
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, cta, 6);      // accumulators staged, TMEM released
  if (p.arenas && threadIdx.x == 0) {
    // every key tile of this CTA has been consumed.  The last CTA of the grid tells the peers that this rank no
    // longer reads the shards of this step: their enqueue may overwrite rows.  Nothing this rank WROTE has to be
    // visible with the flag, so plain (relaxed) system-scope stores do -- no fence on the way to the fold.
    uint8_t* mine = p.arenas[p.rank];
    peer::LocalCtl* ctl = peer::local_ctl(mine);
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[peer::kXSmoothDone]) + 1;
    if (atomicAdd(&ctl->done[peer::kXSmoothDone], 1u) == gridDim.x * gridDim.y - 1) {
      ctl->done[peer::kXSmoothDone] = 0;
      for (int s = 0; s < p.world; ++s)
        if (s != p.rank) peer::st_relaxed_sys(peer::flag_of(p.arenas[s], peer::kXSmoothDone, p.rank), epoch);
      *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[peer::kXSmoothDone]) = epoch;
    }
  }

  // ---- fold the splits of this row group: first inside the cluster through distributed shared memory ----
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = p.cluster;
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  if (CL > 1) cluster.sync();                              // every CTA's reduction tiles are complete and visible
  const int RB = kBM / CL;                                 // rows of every row tile this CTA reduces
  const int W = p.W, C = p.C;
  const int outer = split / CL;
  const float* peer[kMaxCluster];
#pragma unroll
  for (int r = 0; r < kMaxCluster; ++r) peer[r] = (CL > 1 && r < CL) ? cluster.map_shared_rank(sRed, r) : sRed;
  const int W4 = W / 4;
  for (int idx = threadIdx.x; idx < M * RB * W4; idx += blockDim.x) {
    const int m = idx / (RB * W4), rem = idx - m * (RB * W4);
    const int rr = rem / W4, q4 = rem - rr * W4;
    const int row = crank * RB + rr;
    float4 v[kMaxCluster];
#pragma unroll
    for (int r = 0; r < kMaxCluster; ++r)                  // all remote loads in flight at once
      if (r < CL) v[r] = *reinterpret_cast<const float4*>(peer[r] + ((size_t)m * kBM + row) * kRedLd + 4 * q4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kMaxCluster; ++r)                  // rank order: deterministic
      if (r < CL) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    const long long grow = (long long)(tile0 + m) * kBM + row;
    const float vv[4] = {acc.x, acc.y, acc.z, acc.w};
    if (p.nouter == 1) {
      if (grow < p.rows) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 4 * q4 + j;
          if (col < C) p.numer[grow * p.numer_ld + col] = vv[j];
          else if (col == C) p.rowsum[grow * p.rowsum_ld] = vv[j];       // the ones column: sum_k P_ik
        }
      }
    } else {
      *reinterpret_cast<float4*>(p.part + ((size_t)outer * p.rows_pad + grow) * W + 4 * q4) = acc;
    }
  }
  if (CL > 1) cluster.sync();                              // nobody leaves while its smem is still being read
  if (threadIdx.x == 0) {
    B200SSL_STAMP(p.dbg, cta, 7);                          // cluster fold done
    B200SSL_STAMP_NS(p.dbg, cta, 11);
  }
  if (p.nouter == 1) return;

  // ---- then across clusters: the slice "rows [crank*RB, +RB) of every row tile of this group" is folded, in cluster order,
  // by whichever of the `nouter` CTAs that wrote it arrives last -- the outer folds of a launch run on CL CTAs per row group
  // in parallel, each ONE pass over all its row tiles (items = M * RB * W / 4 float4, dealt to thread groups that each add a
  // contiguous range of the partials with 8 loads in flight; the group sums meet in shared memory in group order) ----
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&p.tickets[row_group * kMaxCluster + crank], 1u) == (unsigned)p.nouter - 1;
  __syncthreads();
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, cta, 8);      // ticket taken
  if (!s_last) return;
  __threadfence();
  {
    const int per_tile = RB * W4, items = M * per_tile;
    float4* scratch = reinterpret_cast<float4*>(sRed);     // the cluster is past its last DSMEM read
    const int nt = blockDim.x;
    int P = items <= nt ? nt / items : 1;                  // thread groups (items > nt: one group strides over the items)
    if (P > (p.nouter + 7) / 8) P = (p.nouter + 7) / 8;
    const int per = (p.nouter + P - 1) / P;
    const size_t stride = (size_t)p.rows_pad * W;          // floats between two partials
    auto src_of = [&](int item, long long* grow) -> const float* {
      const int m = item / per_tile, rem = item - m * per_tile;
      const int rr = rem / W4, q4 = rem - rr * W4;
      *grow = (long long)(tile0 + m) * kBM + crank * RB + rr;
      return p.part + (size_t)(*grow) * W + 4 * q4;
    };
    auto emit = [&](int item, long long grow, const float4& v) {
      if (grow >= p.rows) return;
      const int q4 = item % W4;
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 4 * q4 + j;
        if (col < C) p.numer[grow * p.numer_ld + col] = vv[j];
        else if (col == C) p.rowsum[grow * p.rowsum_ld] = vv[j];
      }
    };
    auto range_sum = [&](const float* src, int s_begin, int s_end) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s0 = s_begin; s0 < s_end; s0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (s0 + u < s_end) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(s0 + u) * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (s0 + u < s_end) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      return acc;
    };
    if (P <= 1) {
      for (int item = threadIdx.x; item < items; item += nt) {
        long long grow;
        const float* src = src_of(item, &grow);
        emit(item, grow, range_sum(src, 0, p.nouter));
      }
    } else {
      const int item = threadIdx.x % items, grp = threadIdx.x / items;
      long long grow;
      const float* src = src_of(item, &grow);
      if (grp < P) scratch[grp * items + item] = range_sum(src, grp * per, min(p.nouter, (grp + 1) * per));
      __syncthreads();
      if ((int)threadIdx.x < items) {
        float4 t = scratch[threadIdx.x];
        for (int g = 1; g < P; ++g) {
          const float4 v = scratch[g * items + threadIdx.x];
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        emit(item, grow, t);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.tickets[row_group * kMaxCluster + crank] = 0u;
    B200SSL_STAMP(p.dbg, cta, 9);                          // outer fold done
    B200SSL_STAMP_NS(p.dbg, cta, 12);
  }
}

// ---- fp32 storage -> bf16 hi + mid operands (one launch per K3 call; 2^-17 relative per element) -------------------------
// x = hi + mid + O(2^-17 |x|): hi = bf16(x), mid = bf16(x - hi).  Rows of the queries and of the bank become [hi(64) | mid(64)]
// (term tt of a row = TMA columns [64 tt, +64)); the probabilities become the transposed, class-padded [2][32][Kp] with the
// ones row (-> row sums) in term 0.
struct SplitParams {
  const float* f; long long rows; __nv_bfloat16* f2;        // queries [rows, 64] -> [rows, 128]
  const float* qf; long long K; __nv_bfloat16* qf2;         // bank    [K, 64]    -> [K, 128]
  const float* qp; int C; long long Kp; __nv_bfloat16* qpt2; // probs   [K, C]     -> [2][32][Kp]
};

__device__ __forceinline__ void split8(const float* src, __nv_bfloat16* hi, __nv_bfloat16* mid) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float h[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __bfloat162float(__float2bfloat16_rn(x[i]));
    m[i] = x[i] - h[i];
  }
  *reinterpret_cast<uint4*>(hi) = pack16(h, __nv_bfloat16());
  *reinterpret_cast<uint4*>(mid) = pack16(m, __nv_bfloat16());
}

__global__ void __launch_bounds__(256) split_operands_kernel(const SplitParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long nthreads = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nvec = (p.rows + p.K) * 8;                // 8 floats per item
  for (long long v = tid; v < nvec; v += nthreads) {
    const long long row = v >> 3;
    const int c8 = (int)(v & 7) * 8;
    if (row < p.rows) split8(p.f + row * 64 + c8, p.f2 + row * 128 + c8, p.f2 + row * 128 + 64 + c8);
    else split8(p.qf + (row - p.rows) * 64 + c8, p.qf2 + (row - p.rows) * 128 + c8, p.qf2 + (row - p.rows) * 128 + 64 + c8);
  }
  for (long long k = tid; k < p.Kp; k += nthreads) {        // coalesced over k for every class row
    for (int c = 0; c < kCP; ++c) {
      const float x = k < p.K ? (c < p.C ? __ldg(p.qp + k * p.C + c) : c == p.C ? 1.f : 0.f) : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      p.qpt2[(size_t)c * p.Kp + k] = h;
      p.qpt2[(size_t)(kCP + c) * p.Kp + k] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  }
}

struct SmoothPlan { int mt, cluster, nouter; };

// A/B aids (tools/k3_tune.py): b200ssl_debug_set_k3 at run time (or B200SSL_K3_MT / B200SSL_K3_POLY in the environment).
int g_force_mt = getenv("B200SSL_K3_MT") ? atoi(getenv("B200SSL_K3_MT")) : 0;          // > 0: row tiles per CTA
int g_force_cluster = 0, g_force_nouter = 0;                                           // > 0: cluster size / clusters per row group
int g_force_stages = getenv("B200SSL_K3_STAGES") ? atoi(getenv("B200SSL_K3_STAGES")) : 0;   // 3..5 key-tile stages
int g_p_tmem = getenv("B200SSL_K3_P_SMEM") ? 0 : 1;                                    // 0: P through shared memory (A/B)
int g_force_poly = getenv("B200SSL_K3_POLY") ? atoi(getenv("B200SSL_K3_POLY")) : -1;   // exponentials (of 32) on the FMA pipe

// Clusters of `cl` CTAs (one CTA per SM: the kernel takes more than half an SM's shared memory) that the chip runs at once.
// A cluster lives inside one GPC, so the SMs a GPC has beyond a multiple of `cl` stay idle: 148 / 74 / 33 / 15 on B200.
// Asked from the driver once per size; the table is the fallback without a device (CPU tests of the plan).
int max_active_clusters(int cl) {
  static int cache[kMaxCluster + 1] = {0};
  if (cache[cl]) return cache[cl];
  int n = cl == 1 ? 148 : cl == 2 ? 74 : cl == 4 ? 33 : 15;     // measured on B200 (b200ssl_debug_max_active_clusters)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
    cudaFuncSetAttribute(bank_smooth_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, (unsigned)(cl * kNumSMs), 1);
    cfg.blockDim = dim3(kTcThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_request(1);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)cl; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, bank_smooth_tc_kernel<0, true>, &cfg) == cudaSuccess && q > 0) n = q;
    else (void)cudaGetLastError();
  } else {
    (void)cudaGetLastError();
  }
  cache[cl] = n;
  return n;
}

// Cost of a plan in microseconds (measured constants, profiles/r02_k3_*): a wave of CTAs pays its prologue (barrier init,
// TMEM allocation, first TMA round trip) and mt * T S-tile units; the DSMEM fold of a cluster and the ticketed global fold
// of `no` partials come on top.  Waves are counted in CLUSTERS the chip can hold at once.
double plan_cost(long long row_tiles, long long ktiles, int mt, int cl, long long no, double unit_scale = 1.0) {
  const long long groups = (row_tiles + mt - 1) / mt;
  const long long waves = (groups * no + max_active_clusters(cl) - 1) / max_active_clusters(cl);
  const long long T = (ktiles + cl * no - 1) / (cl * no);
  // outer fold: the last CTA of a (row group, cluster rank) slice adds `no` partials of mt * 128 / cl rows in one pass, 8 loads
  // in flight per thread group (576 / items groups)
  const int items = mt * 6 * kBM / cl;
  const long long groups_t = kTcThreads / items > 1 ? kTcThreads / items : 1;
  const long long rounds = ((no + groups_t - 1) / groups_t + 7) / 8;
  const double fold = no > 1 ? 3.0 + 0.9 * (double)rounds * (items > kTcThreads ? (double)((items + kTcThreads - 1) / kTcThreads) : 1.0) : 0.0;
  return (double)waves * (1.8 + 0.62 * unit_scale * (double)mt * (double)T) + 1.5 + (cl > 1 ? 2.5 : 0.5) + fold + 0.4 * (mt - 1);   // + extra query tiles to stage
}

// How a launch is cut: `mt` row tiles per CTA, and per group of row tiles a cluster of `cluster` CTAs (power of two <= 8,
// folded through DSMEM) times `nouter` clusters (folded through global partials) that share the key tiles.
// `nt` = 2 (fp32 storage as bf16 hi + mid): a unit is 36 instead of 12 MMAs, ~1.6 x the time of a bf16 unit.
SmoothPlan smooth_tc_plan(long long rows, long long ktiles, bool remote, int nt = 1) {
  const long long row_tiles = (rows + kBM - 1) / kBM;
  SmoothPlan best{1, 1, 1};
  double best_cost = 1e300;
  // Shards read over NVLink: one CTA serves every row tile from the key tile it staged, so a remote tile crosses the link
  // once per step (row loop), whenever the row tiles fit one CTA.
  if (nt == 2) {                                            // split operands: the unit tiles are summed in registers, one row tile per CTA
    SmoothPlan b1{1, 1, 1};
    double bc = 1e300;
    for (int cl = 1; cl <= kMaxCluster; cl *= 2) {
      if (cl > ktiles) break;
      const long long no_max = ktiles / cl < 148 ? ktiles / cl : 148;
      for (long long no = 1; no <= no_max; ++no) {
        const double c = plan_cost(row_tiles, ktiles, 1, cl, no, 1.6);
        if (c < bc - 1e-9) { bc = c; b1 = SmoothPlan{1, cl, (int)no}; }
      }
    }
    return b1;
  }
  const int mt_lo = g_force_mt > 0 ? (g_force_mt < kMaxMT ? g_force_mt : kMaxMT) : (remote && row_tiles <= kMaxMT) ? (int)row_tiles : 1;
  const int mt_hi = (g_force_mt > 0 || (remote && row_tiles <= kMaxMT)) ? mt_lo : kMaxMT;
  // with an SM budget (g_head_sm_budget): also the best plan of at most that many CTAs; it wins while it costs <= 1.5 x the best
  SmoothPlan best_b{1, 1, 1};
  double best_b_cost = 1e300;
  const bool forced = g_force_mt > 0 || g_force_cluster > 0 || g_force_nouter > 0;
  for (int mt = mt_lo; mt <= mt_hi; ++mt) {
    if (mt > row_tiles && mt > mt_lo) break;
    const long long groups = (row_tiles + mt - 1) / mt;
    for (int cl = 1; cl <= kMaxCluster; cl *= 2) {
      if (cl > ktiles) break;
      if (g_force_cluster > 0 && cl != g_force_cluster) continue;
      const long long no_max = ktiles / cl < 148 ? ktiles / cl : 148;
      for (long long no = 1; no <= no_max; ++no) {
        if (g_force_nouter > 0 && no != g_force_nouter) continue;
        const double c = plan_cost(row_tiles, ktiles, mt, cl, no, nt == 2 ? 1.6 : 1.0);
        if (c < best_cost - 1e-9) { best_cost = c; best = SmoothPlan{mt, cl, (int)no}; }
        if (g_head_sm_budget > 0 && groups * cl * no <= g_head_sm_budget && c < best_b_cost - 1e-9) { best_b_cost = c; best_b = SmoothPlan{mt, cl, (int)no}; }
      }
    }
  }
  // (not for shards read over NVLink: there the launch lives off the number of tile streams in flight -- measured at N = 8,
  // K = 65536: 100 us/step with the unconstrained plan, 128 with 48 CTAs)
  if (!forced && !remote && g_head_sm_budget > 0 && best_b_cost <= 1.5 * best_cost) return best_b;
  return best;
}

template <int NPOLY, bool PTMEM, int NT = 1>
cudaError_t launch_smooth(const SmoothTcParams& p, const CUtensorMap& tm_f, const BankMaps& maps, dim3 grid, size_t smem, cudaStream_t stream) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(bank_smooth_tc_kernel<NPOLY, PTMEM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return e;
    attr_smem = kSmemMax;
  }
  return launch_pdl(PDL_SMOOTH, bank_smooth_tc_kernel<NPOLY, PTMEM, NT>, grid, dim3(kTcThreads, 1, 1), smem, stream, dim3(1, (unsigned)p.cluster, 1), tm_f,
                    maps, p);
}

}  // namespace

size_t smooth_tc_workspace_floats(long long rows, long long bank_rows, int classes) {
  const long long row_tiles = (rows + kBM - 1) / kBM;
  size_t need = 0;
  for (int remote = 0; remote < 2; ++remote) {             // the sizing call does not know where the bank lives
    // shards whose rows are not a multiple of the key tile add up to one partial tile each
    for (int extra = 0; extra <= (remote ? kMaxSeg : 0); extra += kMaxSeg) {
      const SmoothPlan pl = smooth_tc_plan(rows, (bank_rows + kBN - 1) / kBN + extra, remote != 0);
      const long long groups = (row_tiles + pl.mt - 1) / pl.mt;
      const size_t n = pl.nouter > 1 ? (size_t)pl.nouter * groups * pl.mt * kBM * ((1 + classes + 3) & ~3) : 0;
      if (n > need) need = n;
    }
  }
  return need;
}

namespace {
// smem of the split-operand form: 32 KB per query tile, 48 KB per key-tile stage; 3 stages while they fit (mt <= 2), else 2
constexpr int stages_for_split(int mt) { return mt <= 2 ? 3 : 2; }
constexpr size_t smem_request_split(int mt) { return 1024 + (size_t)mt * 2 * kTileA + (size_t)stages_for_split(mt) * 2 * (kTileQf + kTileQp) + 512; }
static_assert(smem_request_split(kMaxMT) <= kSmemMax && smem_request_split(2) <= kSmemMax, "shared memory budget (split operands)");
static_assert((size_t)kMaxMT * kRedTile <= (size_t)stages_for_split(kMaxMT) * 2 * (kTileQf + kTileQp), "reduction tiles over the drained key stages (split operands)");
}  // namespace

// dim 64, classes <= 31: the tensor-core path.  nt = 1: bf16 operands (bank rows a multiple of 8).  nt = 2: operands split into
// bf16 hi + mid -- `feats` / `queue_feats` are [*, 128] bf16, `queue_probs_t` is [2][32][qpt_ld] bf16 (bank_smooth_tc_f32).
static int bank_smooth_tc_impl(int nt, long long qpt_ld, const void* feats, const void* queue_feats, const void* queue_probs_t, long long rows,
                   long long bank_rows, int classes, float temperature, float* rowsum, float* numer, int rowsum_ld,
                   int numer_ld, const b200ssl_bank_shards* sh, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_bank_smooth_partial[tcgen05]";
  SmoothTcParams p{};
  const bool direct = sh && !sh->replicated;               // shards read in place; a replicated ring is one local segment
  const long long seg_rows = direct ? sh->shard_rows : bank_rows;
  p.nseg = direct ? sh->world : 1;
  p.tps = (int)((seg_rows + kBN - 1) / kBN);
  if (sh) {
    p.arenas = reinterpret_cast<uint8_t* const*>(sh->arenas_dev); p.rank = sh->rank; p.world = sh->world;
    p.seg_first = direct ? sh->rank : 0;
  }
  p.rows = rows; p.C = classes; p.W = (1 + classes + 3) & ~3;
  p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.s_min = -126.f / p.scale;
  p.rowsum = rowsum; p.numer = numer; p.rowsum_ld = rowsum_ld; p.numer_ld = numer_ld; p.dbg = debug_timing_buffer(PDL_SMOOTH);
  const SmoothPlan pl = smooth_tc_plan(rows, (long long)p.nseg * p.tps, direct, nt);
  p.mt = pl.mt; p.cluster = pl.cluster; p.nouter = pl.nouter;
  p.nsplit = p.cluster * p.nouter;
  const long long row_tiles = (rows + kBM - 1) / kBM;
  const long long groups = (row_tiles + p.mt - 1) / p.mt;
  p.row_tiles = (int)row_tiles;
  p.rows_pad = groups * p.mt * kBM;
  const size_t need = kWsHeaderBytes + sizeof(float) * (p.nouter > 1 ? (size_t)p.nouter * p.rows_pad * p.W : 0);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  if ((size_t)groups * kMaxCluster * 4 > kWsTicket2Bytes) return fail(B200SSL_E_SHAPE, "%s: too many row tiles", fn);
  p.tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kWsTicketBytes);
  p.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  CUtensorMap tm_f;
  BankMaps maps;
  if (int e = tc::make_tmap_bf16_2d(&tm_f, feats, (uint64_t)rows, 64 * nt, 128 * nt, kBM, 64)) return e;
  for (int s = 0; s < kMaxSeg; ++s) {
    const int src = direct ? (s < p.nseg ? s : 0) : (sh ? sh->rank : 0);   // unused entries repeat a valid one (never dereferenced)
    const void* qf = sh ? reinterpret_cast<const void*>(sh->arenas_host[src] + sh->feats_offset) : queue_feats;
    const void* qpt = sh ? reinterpret_cast<const void*>(sh->arenas_host[src] + sh->probs_t_offset) : queue_probs_t;
    if (int e = tc::make_tmap_bf16_2d(&maps.qf[s], qf, (uint64_t)seg_rows, 64 * nt, 128 * nt, kBN, 64)) return e;
    if (int e = tc::make_tmap_bf16_2d(&maps.qpt[s], qpt, kCP * nt, (uint64_t)seg_rows, (uint64_t)(nt == 2 ? qpt_ld : seg_rows) * 2, kCP, 64)) return e;
  }
  static_assert(kSmemMax <= 227 * 1024 && smem_request(2, false) <= 227 * 1024 && smem_request(1, false) <= 227 * 1024, "shared memory budget");
  static_assert((size_t)kMaxMT * kRedTile <= smem_stages(kMaxMT, true) && (size_t)kMaxMT * kRedTile <= smem_stages(kMaxMT, false),
                "the reduction tiles must fit in the drained pipeline buffers");
  static_assert(2 * kBN + kMaxMT * kCP <= (int)kTmemCols, "TMEM budget");
  // Share of the exponentials computed on the FMA pipe (of 32 per thread and S tile).  Measured on B200 (tools/k3_tune.py,
  // profiles/r02_k3_experiments.md): while the MMA issue path bounded the kernel every polynomial slot ADDED ~25 clocks per
  // S tile; now 8 of 32 are a wash (81.5 vs 82.1 us at rows 3584 x K 65536) and 16 lose -- the default stays 0, the knob for A/B.
  const int npoly = g_force_poly >= 0 ? g_force_poly : 0;
  const dim3 grid((unsigned)groups, (unsigned)p.nsplit, 1);
  cudaError_t e;
  if (nt == 2) {
    p.stages = stages_for_split(p.mt);
    e = launch_smooth<0, true, 2>(p, tm_f, maps, grid, smem_request_split(p.mt), stream);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
    return check_launch(fn);
  }
  p.stages = stages_for(p.mt, g_p_tmem != 0);
  if (g_force_stages >= 3 && g_force_stages < p.stages) p.stages = g_force_stages;
  const size_t smem = smem_request_n(p.mt, p.stages, g_p_tmem != 0);
  if (!g_p_tmem) e = npoly >= 8 ? launch_smooth<8, false>(p, tm_f, maps, grid, smem, stream) : launch_smooth<0, false>(p, tm_f, maps, grid, smem, stream);
  else e = npoly >= 16 ? launch_smooth<16, true>(p, tm_f, maps, grid, smem, stream)
           : npoly >= 8 ? launch_smooth<8, true>(p, tm_f, maps, grid, smem, stream)
                        : launch_smooth<0, true>(p, tm_f, maps, grid, smem, stream);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

int bank_smooth_tc(const void* feats, const void* queue_feats, const void* queue_probs_t, long long rows,
                   long long bank_rows, int classes, float temperature, float* rowsum, float* numer, int rowsum_ld,
                   int numer_ld, const b200ssl_bank_shards* sh, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return bank_smooth_tc_impl(1, 0, feats, queue_feats, queue_probs_t, rows, bank_rows, classes, temperature, rowsum, numer, rowsum_ld, numer_ld,
                             sh, workspace, workspace_bytes, stream);
}

// Room behind the fold partials for the bf16 hi + mid copies of the queries, the bank and the transposed probabilities.
static size_t split_region(long long rows, long long bank_rows, size_t* off_f2, size_t* off_qf2, size_t* off_qpt2, long long* kp) {
  const long long Kp = (bank_rows + 7) & ~7LL;              // 16-byte row pitch of the transposed probabilities
  auto up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const size_t f2 = up((size_t)rows * 256), qf2 = up((size_t)bank_rows * 256), qpt2 = up((size_t)2 * kCP * Kp * 2);
  if (off_f2) { *off_f2 = 0; *off_qf2 = f2; *off_qpt2 = f2 + qf2; *kp = Kp; }
  return f2 + qf2 + qpt2;
}

size_t smooth_tc_f32_workspace_bytes(long long rows, long long bank_rows, int classes) {
  size_t part = 0;
  const SmoothPlan pl = smooth_tc_plan(rows, (bank_rows + kBN - 1) / kBN, false, 2);
  const long long row_tiles = (rows + kBM - 1) / kBM, groups = (row_tiles + pl.mt - 1) / pl.mt;
  if (pl.nouter > 1) part = sizeof(float) * (size_t)pl.nouter * groups * pl.mt * kBM * ((1 + classes + 3) & ~3);
  return ((part + 1023) & ~(size_t)1023) + 2048 + split_region(rows, bank_rows, nullptr, nullptr, nullptr, nullptr);   // + alignment slack of the tail carve
}

// fp32 storage (the reference's own precision, code/comatch.py:180-181 is a true-fp32 torch.mm): one pre-pass splits the queries,
// the bank and the probabilities into bf16 hi + mid inside the workspace, then the tensor-core kernel runs on the split operands.
// dim 64, classes <= 31, any number of bank rows.
int bank_smooth_tc_f32(const float* feats, const float* queue_feats, const float* queue_probs, long long rows, long long bank_rows,
                       int classes, float temperature, float* rowsum, float* numer, int rowsum_ld, int numer_ld, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_bank_smooth_partial[tcgen05, fp32 storage]";
  const size_t need = kWsHeaderBytes + smooth_tc_f32_workspace_bytes(rows, bank_rows, classes);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  size_t o_f2, o_qf2, o_qpt2;
  long long Kp;
  const size_t split_bytes = split_region(rows, bank_rows, &o_f2, &o_qf2, &o_qpt2, &Kp);
  char* base = static_cast<char*>(workspace) + ((workspace_bytes - split_bytes) & ~(size_t)1023);    // the tail of the workspace
  SplitParams sp{feats, rows, reinterpret_cast<__nv_bfloat16*>(base + o_f2), queue_feats, bank_rows,
                 reinterpret_cast<__nv_bfloat16*>(base + o_qf2), queue_probs, classes, Kp, reinterpret_cast<__nv_bfloat16*>(base + o_qpt2)};
  const long long items = (rows + bank_rows) * 8;
  long long blocks = (items + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  cudaError_t e = launch_pdl(PDL_SMOOTH, split_operands_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, dim3(1, 1, 1), sp);
  if (e != cudaSuccess) return fail((int)e, "%s: split launch: %s", fn, cudaGetErrorString(e));
  return bank_smooth_tc_impl(2, Kp, sp.f2, sp.qf2, sp.qpt2, rows, bank_rows, classes, temperature, rowsum, numer, rowsum_ld, numer_ld, nullptr,
                             workspace, (size_t)(base - static_cast<char*>(workspace)), stream);
}

}  // namespace b200ssl

extern "C" void b200ssl_debug_set_k3(int32_t row_tiles_per_cta, int32_t cluster, int32_t clusters_per_row_group, int32_t poly_of_32) {
  b200ssl::g_force_mt = row_tiles_per_cta;
  b200ssl::g_force_cluster = cluster;
  b200ssl::g_force_nouter = clusters_per_row_group;
  // poly_of_32 < 0: defaults; otherwise exponentials (of 32) on the FMA pipe, + 100: P through shared memory (A/B)
  b200ssl::g_p_tmem = poly_of_32 < 0 ? 1 : (poly_of_32 / 100 != 1);
  b200ssl::g_force_poly = poly_of_32 < 0 ? -1 : poly_of_32 % 100;
}

extern "C" int b200ssl_debug_max_active_clusters(int32_t cluster) {
  return (cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8) ? b200ssl::max_active_clusters(cluster) : 0;
}

// Launch geometry of the tensor-core K3 for a problem size (host only; CPU tests and tools).
extern "C" int b200ssl_debug_smooth_plan(int64_t rows, int64_t bank_rows, int32_t remote_shards, int32_t* out_mt_cluster_nouter) {
  if (!out_mt_cluster_nouter || rows <= 0 || bank_rows <= 0) return b200ssl::fail(B200SSL_E_ARG, "b200ssl_debug_smooth_plan: bad argument");
  const b200ssl::SmoothPlan pl = b200ssl::smooth_tc_plan(rows, (bank_rows + b200ssl::kBN - 1) / b200ssl::kBN, remote_shards != 0);
  out_mt_cluster_nouter[0] = pl.mt;
  out_mt_cluster_nouter[1] = pl.cluster;
  out_mt_cluster_nouter[2] = pl.nouter;
  return 0;
}
