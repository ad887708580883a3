// K6 on the 5th-generation tensor cores (bf16 embeddings, dim 64, classes <= 32):
// graph-contrastive loss and its gradient (code/comatch.py:199-213 + autograd).
//
// Per 128 x 128 tile (i rows of F0, j rows of F1) the MMA thread issues
//   S = F0_i F1_j^T                                   4 x tcgen05.mma  (K = 64)
//   Q = Hi_i Hi_j^T + Hi_i Lo_j^T + Lo_i Hi_j^T      6 x tcgen05.mma  (K = 32 each)
// into TMEM (S: columns 0..127, Q: 128..255).  Q uses the bf16 hi/lo split of the fp32
// pseudo-label probabilities (probs_hl [rows, 64] = [hi(32) | lo(32)], written by the
// finalize kernel) so the 0.8 graph threshold sees ~fp32 accuracy although the operands
// are bf16.  512 epilogue threads (four per TMEM lane, 32 columns each) read S/Q with
// tcgen05.ld and do the exp / threshold / log math; the MMA warp runs in warp-uniform control
// flow and issues from one elected lane (tc::elect_one: operands stay in uniform registers).
//
// One thread-block CLUSTER owns one 128-row strip; its CTAs split the streamed dimension
// and exchange their per-row partials through distributed shared memory, in rank order
// (deterministic, no global round trips, no atomics):
//   forward (one launch):  pass A  rowsum_i = sum_j exp(S/tau), qsum_i = sum_j Qm
//                          -- cluster exchange --
//                          pass B  loss_i, r_i   (dense: exp / log / rcp for every element; with one
//                                  tile per CTA the S/Q tile is simply re-read from TMEM)
//   backward (one launch): dZ = P o (G - r) -> bf16 hi + lo -> swizzled smem -> third MMA group
//                          blockIdx.z = 0: dF0_i += dZ F1_j      (A = dZ K-major,  B = F1 tile MN-major)
//                          blockIdx.z = 1: dF1_j += dZ^T F0_i    (A = dZ MN-major, B = F0 tile MN-major)
//                          -- cluster fold of the [128 x 64] fp32 accumulators --
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"
#include "tc.cuh"

namespace b200ssl {
namespace {

namespace cg = cooperative_groups;

constexpr int kT = 128;                       // tile edge (UMMA M and N)
constexpr int kCtEpiWarps = 16;               // 4 per TMEM lane quarter: each thread owns 32 of the 128 columns of its row
constexpr int kCtThreads = 64 + 32 * kCtEpiWarps;   // warp 0 TMA/alloc, warp 1 MMA, warps 2..17 epilogue
constexpr int kMaxCl = 8;
constexpr uint32_t kTileF = kT * 128;         // 16 KB: [128][64] bf16
constexpr uint32_t kSubZ = kT * 128;          // 16 KB: dZ columns [64*kb, +64)
constexpr int kStatRows = 8;                  // per source rank: [value 0/1][column group 0..3][128 rows]
// NT = bf16 terms per embedding: 1 = bf16 storage; 2 = fp32 storage split into bf16 hi + mid by split_contrast_kernel (S from
// three cross terms, gradient GEMMs from dZ_hi F_hi + dZ_hi F_mid + dZ_lo F_hi; precise reciprocals; pairs within 2e-5 of the
// graph threshold recomputed in fp32 from the probabilities: 1e-5 parity with the reference's fp32 arithmetic).
// Shared memory: own tile (NT F terms + probs hi/lo), key-tile stages of the same shape (two; the backward with NT = 2 has
// room for one and runs its tiles in series), the backward's dZ hi + lo pair (64 KB), barriers, then the exchange area
// (forward: two [8][8][128] fp32 stat areas; backward: the [128][64] fp32 gather buffer).
// Gradient GEMMs of the fp32-storage backward: sums with heavy cancellation (sum_j dZ_ij = 0), two bf16 terms per operand
// leave 1.5e-5 -- so there dZ and the streamed embeddings carry THREE terms (hi, mid, lo; six cross products, 2e-8).
template <int NT, bool BWD> constexpr int stage_terms() { return (NT == 2 && BWD) ? 3 : NT; }
template <int NT, bool BWD> constexpr int z_terms() { return !BWD ? 0 : NT == 2 ? 3 : 2; }
template <int NT> constexpr uint32_t own_bytes() { return (NT + 1) * kTileF; }
template <int NT, bool BWD> constexpr uint32_t stage_bytes() { return (stage_terms<NT, BWD>() + 1) * kTileF; }
template <int NT, bool BWD> constexpr int n_stages() { return (NT == 2 && BWD) ? 1 : 2; }
template <int NT, bool BWD> constexpr size_t smem_ct_request() {
  return 1024 + (size_t)own_bytes<NT>() + (size_t)stage_bytes<NT, BWD>() * n_stages<NT, BWD>() + (size_t)z_terms<NT, BWD>() * 2 * kSubZ + 512 +
         (!BWD ? (size_t)2 * kMaxCl * kStatRows * kT * sizeof(float)     // forward: two stat areas
               : NT == 2 ? 0 : (size_t)kT * 64 * sizeof(float));         // backward: gather buffer (fp32 storage: it reuses the dZ buffers)
}
static_assert(smem_ct_request<1, false>() <= 227 * 1024 && smem_ct_request<2, false>() <= 227 * 1024 &&
              smem_ct_request<1, true>() <= 227 * 1024 && smem_ct_request<2, true>() <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemColsCt = 512;         // S 0..127, Q 128..255; forward: second S/Q buffer 256..511; backward: dF accumulator 256..319
constexpr int kAccLd = 68;                    // floats per row of the staged accumulator (16-byte rows, bank spread)

enum { CB_OWN = 0, CB_KV_FULL = 1, CB_KV_EMPTY = 3, CB_SQ_FULL = 5, CB_SQ_EMPTY = 7, CB_Z_FULL = 9, CB_Z_EMPTY = 10,
       CB_ACC = 11, CB_COUNT = 12 };     // SQ_FULL / SQ_EMPTY: two S/Q buffers in the forward, one (index 0) in the backward

struct ContrastTcParams {
  long long rows;
  int C, cluster;
  float scale;            // log2(e) / tau
  float inv_tau, th;
  float* stats;           // [3][rows]: rowsum, qsum, r
  float* out; unsigned* grid_ticket; float* grid_part;
  const float* loss_u; float lambda_u, lambda_c; float* total_out;
  const float* upstream; float factor; void* g0; void* g1;
  const float* probs;     // NT = 2: fp32 pseudo-label probabilities [rows, C] (exact Q for pairs at the graph threshold)
  unsigned long long* dbg;
  // optional piggy-backed task of the backward launch: sgrad[i] *= (*sup) * sfactor  (the stashed focal-CE gradient)
  void* sgrad; long long snumel; const float* sup; float sfactor;     // sgrad: bf16 (NT = 1) or fp32 (NT = 2)
};

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / x for x > 0 on the FMA pipe (the epilogues are bound by the MUFU unit): exponent-flip seed (12 % off) and three Newton
// steps -> 6e-8 relative.  (Two steps, 2.4e-4, are inside the bf16 budget too -- but the forward now closes r_i = sum qn P / (P + eps)
// in exact arithmetic while the backward evaluates P / (P + eps) per pair, and the two have to cancel in dZ = P (G - r).)
__device__ __forceinline__ float rcp_fma(float x) {
  float y = __uint_as_float(0x7EF311C7u - __float_as_uint(x));
  y = y * fmaf(-x, y, 2.0f);
  y = y * fmaf(-x, y, 2.0f);
  return y * fmaf(-x, y, 2.0f);
}

// kind::f16 instruction descriptor with explicit operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return tc::idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
// MN-major SW128 operand: rows are K, 64 contiguous MN elements per 128-byte row, 8-row groups 1024 B apart
// (SBO), further 64-wide MN blocks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// reciprocal of a positive number: FMA-pipe Newton (2.4e-4) for bf16 storage, MUFU seed + one Newton step (~1e-7) for fp32
template <int NT> __device__ __forceinline__ float rcp_pos(float x) {
  if (NT == 1) return rcp_fma(x);
  const float y = rcpa(x);
  return y * fmaf(-x, y, 2.0f);
}
// q_ij = p_i . p_j in fp32 (fp32 storage only): the tensor-core Q (bf16 hi / lo probabilities) is good to ~1e-5 -- enough to
// rule a pair out of the pseudo-label graph, not for the 1e-5 parity of what the positives contribute (q enters the loss and
// dZ linearly).  Every pair at or above the threshold (~1 / classes of them) takes its q from here, as rarely on the other
// side of the threshold as any other fp32 summation order of the reference's product.
__device__ __forceinline__ float exact_q(const float* probs, int C, long long i, long long j) {
  const float* a = probs + i * C;
  const float* b = probs + j * C;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(__ldg(a + c), __ldg(b + c), acc);
  return acc;
}

struct Smem {
  uint8_t *ownF, *ownP, *stage, *z;
  uint64_t* bars;
  uint32_t* tmem_slot;
  volatile int* abort_flag;
  float* stat;            // [2 passes][8 source ranks][8][128]: per-row partials (value 0/1 x column group 0..3) pushed by the cluster
};

template <int NT, bool BWD>
__device__ __forceinline__ Smem carve(uint8_t* smem_raw) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.ownF = smem;                                            // NT term tiles
  s.ownP = s.ownF + NT * kTileF;
  s.stage = s.ownP + kTileF;                                // per stage: NT term tiles of F, then the probs hi/lo tile
  s.z = s.stage + n_stages<NT, BWD>() * stage_bytes<NT, BWD>();
  s.bars = reinterpret_cast<uint64_t*>(s.z + z_terms<NT, BWD>() * 2 * kSubZ);
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bars + CB_COUNT);
  s.abort_flag = reinterpret_cast<volatile int*>(s.tmem_slot + 1);
  s.stat = (BWD && NT == 2) ? reinterpret_cast<float*>(s.z) : reinterpret_cast<float*>(s.bars + CB_COUNT + 2);
  return s;
}

__device__ __forceinline__ uint32_t setup(const Smem& sm, int warp, int lane, const CUtensorMap* a, const CUtensorMap* b,
                                          const CUtensorMap* c) {
  if (threadIdx.x == 0) {
    tc::mbar_init(&sm.bars[CB_OWN], 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&sm.bars[CB_KV_FULL + s], 1);
      tc::mbar_init(&sm.bars[CB_KV_EMPTY + s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&sm.bars[CB_SQ_FULL + s], 1);
      tc::mbar_init(&sm.bars[CB_SQ_EMPTY + s], kCtEpiWarps);  // one arrival per epilogue warp
    }
    tc::mbar_init(&sm.bars[CB_Z_FULL], kCtEpiWarps);
    tc::mbar_init(&sm.bars[CB_Z_EMPTY], 1);
    tc::mbar_init(&sm.bars[CB_ACC], 1);
    *sm.abort_flag = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(sm.tmem_slot, kTmemColsCt);
  if (warp == 1 && lane == 0) {
    tc::tma_prefetch_desc(a);
    tc::tma_prefetch_desc(b);
    tc::tma_prefetch_desc(c);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  pdl_wait();                                               // previous kernel complete: global memory may be touched now
  return *sm.tmem_slot;
}

// S = A_F B_F^T (4 MMAs) and Q = hi.hi + hi.lo + lo.hi (6 MMAs) into TMEM columns [0,128) and [128,256)
struct SqDesc { uint64_t aF, bF, aP, bP; };
// built by every lane of the MMA warp (warp-uniform values), consumed by the elected issuer
__device__ __forceinline__ SqDesc sq_desc(const uint8_t* aFp, const uint8_t* bFp, const uint8_t* aPp, const uint8_t* bPp) {
  return SqDesc{tc::smem_desc_sw128(tc::smem_u32(aFp)), tc::smem_desc_sw128(tc::smem_u32(bFp)),
                tc::smem_desc_sw128(tc::smem_u32(aPp)), tc::smem_desc_sw128(tc::smem_u32(bPp))};
}
template <int NT>
__device__ __forceinline__ void issue_sq(uint32_t tmem, const SqDesc& d) {
  constexpr uint32_t idesc_sq = idesc_bf16(kT, kT, 0, 0);
  const uint64_t aF = d.aF, bF = d.bF, aP = d.aP, bP = d.bP;
#pragma unroll
  for (int pr = 0; pr < (NT == 2 ? 4 : 1); ++pr) {          // split embeddings: (mid, mid), (mid, hi), (hi, mid), (hi, hi) -- small products first
    // all four cross terms: the logits sit in an exponent (x 1 / tau), the 2^-18 of mid x mid is worth its four MMAs here
    const uint64_t ta = (NT == 2 && pr <= 1) ? (kTileF >> 4) : 0, tb = (NT == 2 && (pr == 0 || pr == 2)) ? (kTileF >> 4) : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem, aF + ta + 2 * k, bF + tb + 2 * k, idesc_sq, (pr | k) != 0);
  }
  // hi = columns 0..31 (byte 0), lo = columns 32..63 (byte 64 -> +4 descriptor units)
#pragma unroll
  for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 2 * k, bP + 2 * k, idesc_sq, k > 0);
#pragma unroll
  for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 2 * k, bP + 4 + 2 * k, idesc_sq, true);
#pragma unroll
  for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 4 + 2 * k, bP + 2 * k, idesc_sq, true);
}

// ============================================================== forward ==============================
// grid = (row tiles, CL), cluster (1, CL, 1): cluster rank c streams the j tiles [c*nt/CL, (c+1)*nt/CL).
template <int NT>
__global__ void __launch_bounds__(kCtThreads, 1)
contrast_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1,
                       const __grid_constant__ CUtensorMap tm_ph, const ContrastTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve<NT, false>(smem_raw);
  constexpr uint32_t kStageBytes = stage_bytes<NT, false>();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int own_tile = blockIdx.x, CL = p.cluster;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const long long ntiles = (p.rows + kT - 1) / kT;
  const long long t0 = ntiles * crank / CL;
  const int T = (int)(ntiles * (crank + 1) / CL - t0);      // >= 1 (CL <= ntiles)
  const int U = (T == 1) ? 1 : 2 * T;                       // tile visits: pass A, then pass B recomputes unless T == 1
  const int ctaid = blockIdx.y * gridDim.x + blockIdx.x;
  pdl_launch_dependents();
  const uint32_t tmem = setup(sm, warp, lane, &tm_f0, &tm_f1, &tm_ph);
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 0);
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 1);
  float* statA = sm.stat;                                   // [src rank][8][128]
  float* statB = sm.stat + kMaxCl * kStatRows * kT;

  // per-thread epilogue state (valid in warps 2..17)
  const int quarter = warp & 3, colq = (warp - 2) >> 2;
  const int r_in = quarter * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  const long long gi = (long long)own_tile * kT + r_in;
  float inv_rs = 1.f, inv_qs = 1.f;

  const int rows_i = (int)p.rows, gi_i = (int)gi;
  // pass 0 accumulators of a thread: mo[0] = sum E, mo[1] = sum qm, mo[2] = sum qm*y (y = log2 of E), mo[3..6] = sum qm * E^-k
  // (k = 1..4), mo[7] = min of E over the pairs of the graph -- everything pass B needs, if the Taylor series below holds
  float mo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 3.0e38f};
  auto moments = [&](float y, float E, float qm) {
    const float ei = rcp_pos<NT>(E), ei2 = ei * ei;
    mo[1] += qm;
    mo[2] = fmaf(qm, y, mo[2]);
    const float qe = qm * ei;
    mo[3] += qe;
    mo[4] = fmaf(qe, ei, mo[4]);
    mo[5] = fmaf(qe, ei2, mo[5]);
    mo[6] = fmaf(qe * ei, ei2, mo[6]);
    mo[7] = qm > 0.f ? fminf(mo[7], E) : mo[7];
  };
  auto epilogue_tile = [&](int pass, int buf, long long j0ll, float& a0, float& a1) {
    const int j0 = (int)j0ll;
    const uint32_t sq_addr = lane_addr + buf * (2 * kT);
    // interior tile: every (row, column) is inside the matrix and off the diagonal -> no per-element checks
    const bool interior = (own_tile * kT + kT <= rows_i) && (j0 + kT <= rows_i) && (j0 != own_tile * kT);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {                           // 2 x 16 columns: S and Q of a half fit the register budget (96 at 576 threads)
    const int col0 = colq * 32 + h * 16;
    uint32_t sv[32], qv[32];                                // (entries 0..15 used)
    tc::tmem_ld_32x16<0>(sq_addr + col0, sv);
    tc::tmem_ld_32x16<0>(sq_addr + kT + col0, qv);
    tc::tmem_ld_wait(sv);
    tc::tmem_ld_wait(qv);
    if (pass == 0) {
      if (interior) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float q = __uint_as_float(qv[j]);
          if (NT == 2 && q > p.th - 2e-5f) q = exact_q(p.probs, p.C, gi, j0 + col0 + j);
          const float y = __uint_as_float(sv[j]) * p.scale, E = ex2a(y);
          mo[0] += E;                                                             // comatch.py:200
          moments(y, E, (q >= p.th) ? q : 0.f);                                   // :206-208
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int gj = j0 + col0 + j;
          const bool ok = (gi_i < rows_i) && (gj < rows_i);
          float q = __uint_as_float(qv[j]);
          if (NT == 2 && ok && q > p.th - 2e-5f) q = exact_q(p.probs, p.C, gi, gj);
          q = (gi_i == gj) ? 1.f : q;                                             // fill_diagonal_(1)  :205
          const float y = __uint_as_float(sv[j]) * p.scale, E = ex2a(y);
          mo[0] += ok ? E : 0.f;
          moments(y, E, (ok && q >= p.th) ? q : 0.f);
        }
      }
    } else {
      // pass B, dense: every element pays exp / log (2 MUFU ops; the reciprocal runs on the FMA pipe); the positives of the pseudo-label graph are only
      // ~1/classes of them, but compacting them first (a shared-memory list per thread, round 2) cost 2.5 x more than
      // the arithmetic it saved
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float q = __uint_as_float(qv[j]);
        bool pos;
        if (interior) {
          if (NT == 2 && q > p.th - 2e-5f) q = exact_q(p.probs, p.C, gi, j0 + col0 + j);
          pos = q >= p.th;
        } else {
          const int gj = j0 + col0 + j;
          const bool ok = (gi_i < rows_i) && (gj < rows_i);
          if (NT == 2 && ok && q > p.th - 2e-5f) q = exact_q(p.probs, p.C, gi, gj);
          q = (gi_i == gj) ? 1.f : q;
          pos = ok && (q >= p.th);
        }
        const float P = ex2a(__uint_as_float(sv[j]) * p.scale) * inv_rs;                // :201
        const float qn = pos ? q * inv_qs : 0.f;                                        // :209
        a0 -= lg2a(P + 1e-7f) * 0.6931471805599453f * qn;                               // :212
        a1 += qn * P * rcp_pos<NT>(P + 1e-7f);
      }
    }
    }
  };

  float a0 = 0.f, a1 = 0.f;
  float* scratch = statB;                                   // [8 values][4 column groups][128 rows]: idle until a pass B pushes
  __shared__ float rowres[8 * kT];                          // [8][128]: cluster-wide row results after exchange A
  bool need_b = false;
#pragma unroll 1
  for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1 && !need_b) break;                       // the closed form below covered pass B
    const int u_begin = phase == 0 ? 0 : T, u_end = phase == 0 ? T : U;     // phase 1 is empty when T == 1
    if (warp == 0) {
      if (lane == 0) {
        if (phase == 0) {
          tc::mbar_arrive_expect_tx(&sm.bars[CB_OWN], own_bytes<NT>());
#pragma unroll
          for (int tt = 0; tt < NT; ++tt) tc::tma_load_2d(sm.ownF + tt * kTileF, &tm_f0, 64 * tt, own_tile * kT, &sm.bars[CB_OWN]);
          tc::tma_load_2d(sm.ownP, &tm_ph, 0, own_tile * kT, &sm.bars[CB_OWN]);
        }
        for (int u = u_begin; u < u_end; ++u) {
          const int s = u & 1;
          if (u >= 2) tc::mbar_wait(&sm.bars[CB_KV_EMPTY + s], ((u >> 1) - 1) & 1, sm.abort_flag);
          const int o0 = (int)((t0 + (u % T)) * kT);
          tc::mbar_arrive_expect_tx(&sm.bars[CB_KV_FULL + s], kStageBytes);
#pragma unroll
          for (int tt = 0; tt < NT; ++tt)
            tc::tma_load_2d(sm.stage + s * kStageBytes + tt * kTileF, &tm_f1, 64 * tt, o0, &sm.bars[CB_KV_FULL + s]);
          tc::tma_load_2d(sm.stage + s * kStageBytes + NT * kTileF, &tm_ph, 0, o0, &sm.bars[CB_KV_FULL + s]);
        }
      }
    } else if (warp == 1) {
      // the whole warp runs the loop (warp-uniform control flow), one elected lane issues the MMAs
      const bool leader = tc::elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      if (phase == 0) tc::mbar_wait(&sm.bars[CB_OWN], 0, sm.abort_flag);
      for (int u = u_begin; u < u_end; ++u) {
        const int s = u & 1;                                // key-tile stage and S/Q buffer of this visit
        tc::mbar_wait(&sm.bars[CB_KV_FULL + s], (u >> 1) & 1, sm.abort_flag);
        if (u >= 2) tc::mbar_wait(&sm.bars[CB_SQ_EMPTY + s], ((u >> 1) - 1) & 1, sm.abort_flag);   // S/Q of visit u-2 consumed
        tc::tcgen05_fence_after();
        const uint8_t* stF = sm.stage + s * kStageBytes;
        const SqDesc d = sq_desc(sm.ownF, stF, sm.ownP, stF + NT * kTileF);
        if (leader) {
          issue_sq<NT>(tmem_u + s * (2 * kT), d);          // the MMAs of visit u+1 run under the epilogue of visit u
          tc::mma_commit(&sm.bars[CB_SQ_FULL + s]);
          tc::mma_commit(&sm.bars[CB_KV_EMPTY + s]);
        }
        __syncwarp();
      }
    } else {
      if (phase == 1) {                                     // exact pass B: full row statistics from exchange A
        inv_rs = 1.0f / rowres[0 * kT + r_in];
        inv_qs = 1.0f / rowres[1 * kT + r_in];
        a0 = a1 = 0.f;
        if (T == 1) {                                       // the only S/Q tile of this CTA is still in TMEM
          tc::tcgen05_fence_after();
          epilogue_tile(1, 0, t0 * kT, a0, a1);
        }
      }
      for (int u = u_begin; u < u_end; ++u) {
        tc::mbar_wait(&sm.bars[CB_SQ_FULL + (u & 1)], (u >> 1) & 1, sm.abort_flag);
        if (u == 0 && threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 2);     // first S/Q tile ready (TMA + 10 MMAs)
        tc::tcgen05_fence_after();
        epilogue_tile(phase, u & 1, (t0 + (u % T)) * kT, a0, a1);
        tc::tcgen05_fence_before();
        if (!(T == 1)) tc::mbar_arrive_warp(&sm.bars[CB_SQ_EMPTY + (u & 1)], lane);
      }
      const int RBx = kT / CL;
      if (phase == 0) {
        // the four column groups of a row meet in shared memory (fixed order), then ONE thread per row pushes the eight
        // row partials of this CTA to every CTA of the cluster (all of them need the full row statistics)
#pragma unroll
        for (int v = 0; v < 8; ++v) scratch[(v * 4 + colq) * kT + r_in] = mo[v];
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCtEpiWarps) : "memory");
        if (colq == 0) {
          float sum[8];
#pragma unroll
          for (int v = 0; v < 7; ++v)
            sum[v] = (scratch[(v * 4 + 0) * kT + r_in] + scratch[(v * 4 + 1) * kT + r_in]) +
                     (scratch[(v * 4 + 2) * kT + r_in] + scratch[(v * 4 + 3) * kT + r_in]);
          sum[7] = fminf(fminf(scratch[(28 + 0) * kT + r_in], scratch[(28 + 1) * kT + r_in]),
                         fminf(scratch[(28 + 2) * kT + r_in], scratch[(28 + 3) * kT + r_in]));
          for (int r = 0; r < CL; ++r) {
            float* st = statA + crank * kStatRows * kT;
            if (CL > 1) st = cluster.map_shared_rank(st, r);
#pragma unroll
            for (int v = 0; v < 8; ++v) st[v * kT + r_in] = sum[v];
          }
        }
      } else {
        // exact pass B: partials only to the CTA that folds this row
        for (int r = 0; r < CL; ++r) {
          if (r != r_in / RBx) continue;
          float* st = statB + crank * kStatRows * kT;
          if (CL > 1) st = cluster.map_shared_rank(st, r);
          st[(0 + colq) * kT + r_in] = a0;
          st[(4 + colq) * kT + r_in] = a1;
        }
      }
      if (threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 3 + 2 * phase);     // pass A (3) / pass B (5) done
    }
    if (CL > 1) cluster.sync(); else __syncthreads();       // partials of this pass visible cluster-wide
    if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 4 + 2 * phase);       // after exchange A (4) / B (6)
    if (phase == 0) {
      // Row results from every CTA's partials (rank order; every CTA of the cluster computes the same numbers) and the
      // decision whether the closed form holds: with z = eps * rs / E,
      //   log(P + eps) = ln2 * y - log(rs) + log1p(z),   P / (P + eps) = 1 / (1 + z),
      // and sum_j qm z^k = (eps rs)^k sum_j qm E^-k, so a degree-4 Taylor series of both needs only the moments of pass A.
      // It holds while the largest z of a row (its smallest E among the pairs of the graph) stays below kZLim; otherwise
      // (never seen with unit-norm embeddings below ~1e5 rows) the whole cluster runs the exact second pass.
      constexpr float kZLim = NT == 2 ? 0.03f : 0.25f;      // z^5 / 5: 5e-9 / 2e-4 absolute on a log of magnitude >= 1
      int bad = 0;
      if (warp >= 2 && colq == 0) {
        float v8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 3.0e38f};
        for (int r = 0; r < CL; ++r) {
          const float* st = statA + r * kStatRows * kT;
#pragma unroll
          for (int v = 0; v < 7; ++v) v8[v] += st[v * kT + r_in];
          v8[7] = fminf(v8[7], st[7 * kT + r_in]);
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) rowres[v * kT + r_in] = v8[v];
        if (crank == 0 && gi < p.rows) { p.stats[gi] = v8[0]; p.stats[p.rows + gi] = v8[1]; }
        bad = (gi < p.rows) && (1e-7f * v8[0] > kZLim * v8[7]);
      }
      need_b = __syncthreads_or(bad) != 0;                  // (also orders rowres for every reader below)
      if (threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 9);                  // row statistics assembled
    }
  }
  // ---- loss_i and r_i: cluster rank c folds rows [c*RB, (c+1)*RB) in rank order ----
  const int RB = kT / CL;
  float lsum = 0.f;
  if (threadIdx.x < RB) {
    const int row = crank * RB + threadIdx.x;
    float li = 0.f, rr = 0.f;
    if (need_b) {
      for (int r = 0; r < CL; ++r) {                        // rank order, local reads
        const float* st = statB + r * kStatRows * kT;
        li += (st[0 * kT + row] + st[1 * kT + row]) + (st[2 * kT + row] + st[3 * kT + row]);
        rr += (st[4 * kT + row] + st[5 * kT + row]) + (st[6 * kT + row] + st[7 * kT + row]);
      }
    } else {
      const float rs = rowres[0 * kT + row], qs = rowres[1 * kT + row], m1 = rowres[2 * kT + row];
      const float c = 1e-7f * rs, c2 = c * c;
      const float s1 = c * rowres[3 * kT + row], s2 = c2 * rowres[4 * kT + row], s3 = c2 * c * rowres[5 * kT + row],
                  s4 = c2 * c2 * rowres[6 * kT + row];
      const float inv_q = 1.0f / qs;
      // loss_i = -sum_j qn log(P + eps)   (:209-212);   r_i = sum_j qn P / (P + eps)
      li = -inv_q * (0.6931471805599453f * m1 - logf(rs) * qs + (s1 - 0.5f * s2 + s3 * (1.0f / 3.0f) - 0.25f * s4));
      rr = inv_q * (qs - s1 + s2 - s3 + s4);
    }
    const long long g = (long long)own_tile * kT + row;
    if (g < p.rows) {
      p.stats[2 * p.rows + g] = -rr / (float)p.rows;
      lsum = li;
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();                                          // all exchanges were pushes: nothing remote is read after exchange B
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem, kTmemColsCt);
  }
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 7);     // rows folded, TMEM freed
  // ---- loss: CTA sum, then last CTA of the grid folds all CTA partials in order ----
  __shared__ float s_w[kCtThreads / 32];
  __shared__ bool s_glast;
  const int tid = threadIdx.x;
  const unsigned ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  lsum = warp_sum(lsum);
  if (lane == 0) s_w[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kCtThreads / 32; ++w) c += s_w[w];
    p.grid_part[cta] = c;
    __threadfence();
    s_glast = (atomicAdd(p.grid_ticket, 1u) == ncta - 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 8);     // grid ticket taken
  if (!s_glast) return;
  __threadfence();
  float v = 0.f;
  for (unsigned b = tid; b < ncta; b += blockDim.x) v += __ldcg(p.grid_part + b);
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kCtThreads / 32; ++w) c += s_w[w];
    const float lc = c / (float)p.rows;                                         // :213
    p.out[0] = lc;
    if (p.total_out) p.total_out[0] = p.lambda_u * (p.loss_u ? *p.loss_u : 0.f) + p.lambda_c * lc;   // :222
    *p.grid_ticket = 0u;
  }
}

// ============================================================== backward =============================
// grid = (tiles, CL, 2), cluster (1, CL, 1): z = 0 -> dF0 of row tile x, z = 1 -> dF1 of column tile x.
template <int NT>
__global__ void __launch_bounds__(kCtThreads, 1)
contrast_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1,
                       const __grid_constant__ CUtensorMap tm_ph, const ContrastTcParams p) {
  pdl_launch_dependents();
  if (blockIdx.z == 2) {                                    // whole clusters of this z-slice only scale the stashed gradient
    pdl_wait();
    const float sc = (p.sup ? *p.sup : 1.f) * p.sfactor;
    const long long stride = (long long)gridDim.x * gridDim.y * blockDim.x;
    const long long first = (long long)(blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    if (NT == 2) {                                          // fp32 storage
      float* g = static_cast<float*>(p.sgrad);
      for (long long i = first; i < p.snumel; i += stride) g[i] *= sc;
      return;
    }
    __nv_bfloat16* g = static_cast<__nv_bfloat16*>(p.sgrad);
    const long long nvec = p.snumel / 8;
    for (long long v = first; v < nvec; v += stride) {
      float f[8];
      unpack16(*reinterpret_cast<const uint4*>(g + v * 8), f, __nv_bfloat16());
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= sc;
      *reinterpret_cast<uint4*>(g + v * 8) = pack16(f, __nv_bfloat16());
    }
    for (long long i = nvec * 8 + first; i < p.snumel; i += stride) g[i] = __float2bfloat16_rn(__bfloat162float(g[i]) * sc);
    return;
  }
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve<NT, true>(smem_raw);
  constexpr uint32_t kStageBytes = stage_bytes<NT, true>();
  constexpr int kStages = n_stages<NT, true>();             // 1: tiles in series (S/Q of tile t+1 cannot be issued under tile t)
  constexpr int kSTerms = stage_terms<NT, true>();          // bf16 terms of the streamed embeddings (3 with fp32 storage)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const bool colmode = blockIdx.z == 1;                     // own tile = j rows of F1
  const int own_tile = blockIdx.x, CL = p.cluster;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const long long ntiles = (p.rows + kT - 1) / kT;
  const long long t0 = ntiles * crank / CL;
  const int T = (int)(ntiles * (crank + 1) / CL - t0);
  const int ctaid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const uint32_t tmem = setup(sm, warp, lane, &tm_f0, &tm_f1, &tm_ph);
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 0);
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 1);
  float* sGather = sm.stat;                                 // [source rank][RB rows][64] fp32 = 32 KB, written by the peers
  float fin[16];                                            // epilogue threads: their 16 columns of the finished gradient tile
  const float up = (p.upstream ? *p.upstream : 1.f) * p.factor * p.inv_tau;   // fetched early: off the critical tail

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&sm.bars[CB_OWN], own_bytes<NT>());
#pragma unroll
      for (int tt = 0; tt < NT; ++tt) tc::tma_load_2d(sm.ownF + tt * kTileF, colmode ? &tm_f1 : &tm_f0, 64 * tt, own_tile * kT, &sm.bars[CB_OWN]);
      tc::tma_load_2d(sm.ownP, &tm_ph, 0, own_tile * kT, &sm.bars[CB_OWN]);
      for (int t = 0; t < T; ++t) {
        const int s = t % kStages;
        if (t >= kStages) tc::mbar_wait(&sm.bars[CB_KV_EMPTY + s], (t / kStages - 1) & 1, sm.abort_flag);
        const int o0 = (int)((t0 + t) * kT);
        tc::mbar_arrive_expect_tx(&sm.bars[CB_KV_FULL + s], kStageBytes);
#pragma unroll
        for (int tt = 0; tt < kSTerms; ++tt)
          tc::tma_load_2d(sm.stage + s * kStageBytes + tt * kTileF, colmode ? &tm_f0 : &tm_f1, 64 * tt, o0, &sm.bars[CB_KV_FULL + s]);
        tc::tma_load_2d(sm.stage + s * kStageBytes + kSTerms * kTileF, &tm_ph, 0, o0, &sm.bars[CB_KV_FULL + s]);
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (warp-uniform control flow), one elected lane issues: 10 + 16 MMAs per tile
    const bool leader = tc::elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    constexpr uint32_t idesc_row = idesc_bf16(kT, 64, 0, 1);    // dF0: A = dZ (K-major),  B = F1 tile (MN-major)
    constexpr uint32_t idesc_col = idesc_bf16(kT, 64, 1, 1);    // dF1: A = dZ (MN-major), B = F0 tile (MN-major)
    tc::mbar_wait(&sm.bars[CB_OWN], 0, sm.abort_flag);
    const uint32_t zaddr = tc::smem_u32(sm.z);
    // rows of S/Q are always the i side (F0), columns the j side (F1).  S/Q of tile t+1 is issued as soon as the epilogue
    // holds tile t in registers (SQ_EMPTY), i.e. under the dZ arithmetic of tile t; the gradient MMAs of tile t follow.
    auto sq = [&](int t) {
      const int s = t % kStages;
      const uint8_t* stF = sm.stage + s * kStageBytes;
      tc::mbar_wait(&sm.bars[CB_KV_FULL + s], (t / kStages) & 1, sm.abort_flag);
      if (t >= 1) tc::mbar_wait(&sm.bars[CB_SQ_EMPTY], (t - 1) & 1, sm.abort_flag);
      tc::tcgen05_fence_after();
      const SqDesc d = !colmode ? sq_desc(sm.ownF, stF, sm.ownP, stF + kSTerms * kTileF) : sq_desc(stF, sm.ownF, stF + kSTerms * kTileF, sm.ownP);
      if (leader) {
        issue_sq<NT>(tmem_u, d);
        tc::mma_commit(&sm.bars[CB_SQ_FULL]);
      }
      __syncwarp();
    };
    sq(0);
    for (int t = 0; t < T; ++t) {
      const int s = t % kStages;
      const uint8_t* stF = sm.stage + s * kStageBytes;
      if (kStages == 2 && t + 1 < T) sq(t + 1);
      tc::mbar_wait(&sm.bars[CB_Z_FULL], t & 1, sm.abort_flag);
      tc::tcgen05_fence_after();
      const uint64_t bB = smem_desc_sw128_mn(tc::smem_u32(stF), 0);
      const uint64_t aZk = tc::smem_desc_sw128(zaddr);            // dZ K-major: sub-tile q at + q * kSubZ
      const uint64_t aZm = smem_desc_sw128_mn(zaddr, kSubZ);      // dZ MN-major: hi pair at 0, lo pair at + 2 * kSubZ
      if (leader) {
        // (dZ term, F term).  bf16 storage: (hi, hi), (lo, hi) -- dZ = hi + lo, two bf16 terms ~ 16 mantissa bits.
        // fp32 storage: (mid, mid), (hi, lo), (lo, hi), (hi, mid), (mid, hi), (hi, hi) -- three terms each, small products first
        constexpr int kPairs = NT == 2 ? 6 : 2;
        constexpr int zt[6] = {1, 0, 2, 0, 1, 0}, ft[6] = {1, 2, 0, 1, 0, 0};
#pragma unroll
        for (int pr = 0; pr < kPairs; ++pr) {
          const int part = NT == 2 ? zt[pr] : pr;
          const uint64_t bT = bB + (NT == 2 ? (uint64_t)ft[pr] * (kTileF >> 4) : 0);
          if (!colmode) {
            // dF0[i, d] += sum_j dZ[i, j] F1[j, d]
#pragma unroll
            for (int k = 0; k < 8; ++k)
              tc::mma_bf16_ss(tmem_u + 2 * kT, aZk + (uint64_t)((2 * part + (k >> 2)) * (kSubZ >> 4) + 2 * (k & 3)), bT + 128 * k, idesc_row,
                              ((NT == 2 ? 0 : t) | k | pr) != 0);
          } else {
            // dF1[j, d] += sum_i dZ[i, j] F0[i, d]
#pragma unroll
            for (int k = 0; k < 8; ++k)
              tc::mma_bf16_ss(tmem_u + 2 * kT, aZm + (uint64_t)(2 * part * (kSubZ >> 4) + 128 * k), bT + 128 * k, idesc_col,
                              ((NT == 2 ? 0 : t) | k | pr) != 0);
          }
        }
        tc::mma_commit(&sm.bars[CB_KV_EMPTY + s]);
        tc::mma_commit(&sm.bars[CB_Z_EMPTY]);
      }
      __syncwarp();
      if (kStages == 1 && t + 1 < T) sq(t + 1);             // one stage: the next tile's S/Q only after this tile's gradient MMAs
    }
    if (leader) tc::mma_commit(&sm.bars[CB_ACC]);
    __syncwarp();
  } else {
    // ===== epilogue: 4 threads per TMEM lane (row i), 32 columns (j) each =====
    const int quarter = warp & 3, colq = (warp - 2) >> 2;
    const int half = colq >> 1, c2 = colq & 1;              // dZ sub-tile (64 columns) and 32-column group inside it
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const float inv_rows = 1.0f / (float)p.rows;
    const int rows_i = (int)p.rows;
    int gi = colmode ? 0 : own_tile * kT + r_in;
    float inv_rs = 1.f, qscale = 0.f, r_i = 0.f;            // qscale = -1 / (qsum_i * rows)
    auto load_stats = [&]() {
      inv_rs = 1.f; qscale = 0.f; r_i = 0.f;
      if (gi < rows_i) {
        inv_rs = 1.0f / p.stats[gi];
        qscale = -inv_rows / p.stats[p.rows + gi];
        r_i = p.stats[2 * p.rows + gi];
      }
    };
    if (!colmode) load_stats();
    const uint32_t z_u32 = tc::smem_u32(sm.z);
    // fp32 storage (1e-5 parity bar): the tensor core truncates when it adds into its fp32 accumulator, a bias that grows with
    // the length of the chain.  There every tile starts a fresh accumulator (24 MMAs) and the tile results are summed here
    // in registers, round-to-nearest: 16 of the 64 gradient columns of the row per thread.
    float acc16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc16[i] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int o0 = (int)((t0 + t) * kT);
      const int i0 = colmode ? o0 : own_tile * kT, j0 = colmode ? own_tile * kT : o0;
      if (colmode) { gi = o0 + r_in; load_stats(); }
      const bool interior = (i0 + kT <= rows_i) && (j0 + kT <= rows_i) && (i0 != j0);
      tc::mbar_wait(&sm.bars[CB_SQ_FULL], t & 1, sm.abort_flag);
      if (t == 0 && threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 2);       // first S/Q tile ready
      tc::tcgen05_fence_after();
      {
        const int col0 = colq * 32;
        uint32_t sv[32], qv[32];
        tc::tmem_ld_32x32(lane_addr + col0, sv);
        tc::tmem_ld_32x32(lane_addr + kT + col0, qv);
        tc::tmem_ld_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive_warp(&sm.bars[CB_SQ_EMPTY], lane);  // S/Q are in registers: the MMA warp may issue the next tile's
        if constexpr (NT == 2) {
          // fp32 storage: dZ as THREE bf16 terms, formed and stored 8 columns at a time (register budget).  The tiles run in
          // series here, so the dZ buffers are free by now: the gradient MMAs of tile t-1 retired before S/Q of tile t were issued
          if (t >= 1) {
            tc::mbar_wait(&sm.bars[CB_Z_EMPTY], (t - 1) & 1, sm.abort_flag);
            uint32_t av[32];                                // bank the gradient tile of t-1 (see acc16)
            tc::tcgen05_fence_after();
            tc::tmem_ld_32x16<0>(lane_addr + 2 * kT + colq * 16, av);
            tc::tmem_ld_wait(av);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc16[i] += __uint_as_float(av[i]);
            tc::tcgen05_fence_before();
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {                    // pairs of the graph: q from the fp32 probabilities
            const float q = __uint_as_float(qv[j]);
            if (q > p.th - 2e-5f && gi < rows_i && j0 + col0 + j < rows_i) qv[j] = __float_as_uint(exact_q(p.probs, p.C, gi, j0 + col0 + j));
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t ph[4], pm[4], pl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float d[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int j = 8 * q4 + 2 * k + u, gj = j0 + col0 + j;
                const float s = __uint_as_float(sv[j]);
                float q = __uint_as_float(qv[j]);
                bool ok = true;
                if (!interior) {
                  ok = (gi < rows_i) && (gj < rows_i);
                  q = (gi == gj) ? 1.f : q;                                       // fill_diagonal_(1)
                }
                const float qm = (q >= p.th) ? q : 0.f;
                const float P = ex2a(s * p.scale) * inv_rs;
                d[u] = ok ? P * (qm * qscale * rcp_pos<2>(P + 1e-7f) - r_i) : 0.f;
              }
              const __nv_bfloat162 h = __floats2bfloat162_rn(d[0], d[1]);
              const float r0 = d[0] - __low2float(h), r1 = d[1] - __high2float(h);
              const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
              const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - __low2float(m), r1 - __high2float(m));
              ph[k] = *reinterpret_cast<const uint32_t*>(&h);
              pm[k] = *reinterpret_cast<const uint32_t*>(&m);
              pl[k] = *reinterpret_cast<const uint32_t*>(&l);
            }
            const uint32_t off = half * kSubZ + tc::sw128_offset(r_in, c2 * 4 + q4);
            tc::st_shared_v4(z_u32 + off, ph[0], ph[1], ph[2], ph[3]);
            tc::st_shared_v4(z_u32 + 2 * kSubZ + off, pm[0], pm[1], pm[2], pm[3]);
            tc::st_shared_v4(z_u32 + 4 * kSubZ + off, pl[0], pl[1], pl[2], pl[3]);
          }
        } else {
        uint32_t packed[16], packed_lo[16];
        // dZ = P (G - r),  G = -(qm / qsum) / (P + 1e-7) / rows   (branch-free: qm = 0 off the graph).  Interior tiles (all
        // but the diagonal and the ragged edge) skip every per-element bound / diagonal test: two instantiations of the
        // loop instead of a uniform branch and three selects per element (ncu, round 1: 35 instructions per element).
        auto dz_of = [&](float s, float q) {
          const float qm = (q >= p.th) ? q : 0.f;
          const float P = ex2a(s * p.scale) * inv_rs;
          return P * (qm * qscale * rcp_pos<NT>(P + 1e-7f) - r_i);
        };

        auto pack_pair = [&](int k, float d0, float d1) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
          const __nv_bfloat162 l = __floats2bfloat162_rn(d0 - __low2float(h), d1 - __high2float(h));
          packed[k] = *reinterpret_cast<const uint32_t*>(&h);
          packed_lo[k] = *reinterpret_cast<const uint32_t*>(&l);
        };
        if (interior) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            pack_pair(k, dz_of(__uint_as_float(sv[2 * k]), __uint_as_float(qv[2 * k])),
                      dz_of(__uint_as_float(sv[2 * k + 1]), __uint_as_float(qv[2 * k + 1])));
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float d[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = 2 * k + u, gj = j0 + col0 + j;
              const bool ok = (gi < rows_i) && (gj < rows_i);
              const float q = (gi == gj) ? 1.f : __uint_as_float(qv[j]);       // fill_diagonal_(1)
              d[u] = ok ? dz_of(__uint_as_float(sv[j]), q) : 0.f;
            }
            pack_pair(k, d[0], d[1]);
          }
        }
        // the dZ buffer must have been consumed by the gradient MMAs of the previous tile -- only now, after the arithmetic
        if (t >= 1) tc::mbar_wait(&sm.bars[CB_Z_EMPTY], (t - 1) & 1, sm.abort_flag);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = c2 * 4 + q4;                     // 16-byte chunk inside the 64-column sub-tile `half`
          const uint4 v = make_uint4(packed[4 * q4], packed[4 * q4 + 1], packed[4 * q4 + 2], packed[4 * q4 + 3]);
          const uint4 vl = make_uint4(packed_lo[4 * q4], packed_lo[4 * q4 + 1], packed_lo[4 * q4 + 2], packed_lo[4 * q4 + 3]);
          tc::st_shared_v4(z_u32 + half * kSubZ + tc::sw128_offset(r_in, chunk), v.x, v.y, v.z, v.w);
          tc::st_shared_v4(z_u32 + (2 + half) * kSubZ + tc::sw128_offset(r_in, chunk), vl.x, vl.y, vl.z, vl.w);
        }
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive_warp(&sm.bars[CB_Z_FULL], lane);
      if (t == T - 1 && threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 3);   // last dZ tile written
    }
    tc::mbar_wait(&sm.bars[CB_ACC], 0, sm.abort_flag);       // all MMAs retired: dZ smem is free, accumulator final
    if (threadIdx.x == 64) B200SSL_STAMP(p.dbg, ctaid, 4);   // gradient accumulator complete
    tc::tcgen05_fence_after();
    {   // accumulator (+ the tiles banked in registers) -> fin[]: this thread's 16 fp32 columns of its row
      uint32_t av[32];
      tc::tmem_ld_32x16<0>(lane_addr + 2 * kT + colq * 16, av);
      tc::tmem_ld_wait(av);
#pragma unroll
      for (int i = 0; i < 16; ++i) fin[i] = __uint_as_float(av[i]) + acc16[i];
    }
    tc::tcgen05_fence_before();
  }
  // fp32 storage: the gather buffer reuses the dZ buffers, which a CTA's gradient MMAs read until its last tile -- so nobody
  // pushes before EVERY CTA of the cluster has retired its MMAs.  (bf16 storage has a gather buffer of its own: no extra sync.)
  if (NT == 2) { if (CL > 1) cluster.sync(); else __syncthreads(); }
  if (warp >= 2) {
    // -> the CTA that folds this row: thread (row, colq) pushes its 16 fp32 columns into the owner's gather buffer
    // [source rank][local row][64] through distributed shared memory
    const int quarter = warp & 3, colq = (warp - 2) >> 2, r_in = quarter * 32 + lane;
    const int RBx = kT / CL, owner = r_in / RBx, lrow = r_in - owner * RBx;
    float* base = sGather + ((size_t)crank * RBx + lrow) * 64;
    float4* dst = reinterpret_cast<float4*>(CL > 1 ? cluster.map_shared_rank(base, owner) : base);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)                          // 16-byte chunk index XOR (row & 7): bank spread at the destination
      dst[(colq * 4 + q4) ^ (lrow & 7)] = make_float4(fin[4 * q4], fin[4 * q4 + 1], fin[4 * q4 + 2], fin[4 * q4 + 3]);
  }
  if (CL > 1) cluster.sync(); else __syncthreads();          // every slice has landed; afterwards only local reads
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 5);
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem, kTmemColsCt);
  }
  // ---- fold: this CTA owns rows [crank*RB, (crank+1)*RB) and adds the CL pushed slices in rank order ----
  const int RB = kT / CL;
  if (NT == 2) {                                            // fp32 gradients: 16 x (4 fp32 = 16 B) per row
    float* out = static_cast<float*>(colmode ? p.g1 : p.g0);
    for (int idx = threadIdx.x; idx < RB * 16; idx += blockDim.x) {
      const int rr = idx >> 4, c4 = idx & 15;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kMaxCl; ++r)
        if (r < CL) {
          const float4 v = reinterpret_cast<const float4*>(sGather + ((size_t)r * RB + rr) * 64)[c4 ^ (rr & 7)];
          f.x += v.x; f.y += v.y; f.z += v.z; f.w += v.w;
        }
      const long long grow = (long long)own_tile * kT + crank * RB + rr;
      if (grow < p.rows) *reinterpret_cast<float4*>(out + grow * 64 + 4 * c4) = make_float4(f.x * up, f.y * up, f.z * up, f.w * up);
    }
    if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 6);
    return;
  }
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(colmode ? p.g1 : p.g0);
  for (int idx = threadIdx.x; idx < RB * 8; idx += blockDim.x) {     // 8 x (8 bf16 = 16 B) per row
    const int rr = idx >> 3, c8 = idx & 7;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < kMaxCl; ++r)
      if (r < CL) {
        const float4* row = reinterpret_cast<const float4*>(sGather + ((size_t)r * RB + rr) * 64);
        const float4 lo = row[(2 * c8) ^ (rr & 7)];
        const float4 hi = row[(2 * c8 + 1) ^ (rr & 7)];
        f[0] += lo.x; f[1] += lo.y; f[2] += lo.z; f[3] += lo.w;
        f[4] += hi.x; f[5] += hi.y; f[6] += hi.z; f[7] += hi.w;
      }
    const long long grow = (long long)own_tile * kT + crank * RB + rr;
    if (grow < p.rows) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= up;
      *reinterpret_cast<uint4*>(out + grow * 64 + 8 * c8) = pack16(f, __nv_bfloat16());
    }
  }
  if (threadIdx.x == 0) B200SSL_STAMP(p.dbg, ctaid, 6);     // fold done
}

// Cluster size of a launch with `tiles * slices` clusters: the widest cluster (shortest per-CTA tile loop) whose clusters
// are all resident at once.  A cluster lives inside one GPC, so a B200 runs only 15 clusters of 8 one-CTA-per-SM blocks at a
// time (33 of 4, 74 of 2 -- b200ssl_debug_max_active_clusters); a launch that needs several waves of clusters pays the
// prologue and the folds once per wave (28 strips x 8 at rows 3584 were two forward and four backward waves in round 1).
int ct_cluster(long long rows, int slices) {
  const long long tiles = (rows + kT - 1) / kT;
  static const int cap[4][2] = {{8, 15}, {4, 33}, {2, 74}, {1, 148}};
  // with an SM budget of the head (g_head_sm_budget): the widest cluster whose launch stays inside it, if any
  if (g_head_sm_budget > 0)
    for (const auto& c : cap)
      if (c[0] <= tiles && tiles * slices <= c[1] && tiles * slices * c[0] <= g_head_sm_budget) return c[0];
  for (const auto& c : cap)
    if (c[0] <= tiles && tiles * slices <= c[1]) return c[0];
  return 1;
}

// ---- fp32 storage -> bf16 operands (one launch ahead of the forward and of the backward kernel) ----------------------
// embeddings [rows, 64] fp32 -> [rows, 192] bf16 = [hi(64) | mid(64) | lo(64)];  probabilities [rows, C] fp32 -> [rows, 64] bf16 =
// [hi(32) | lo(32)], classes >= C zero (what the row kernels emit as probs_hl for bf16 steps)
struct SplitCtParams {
  const float* f0; const float* f1; const float* probs; long long rows; int C;
  __nv_bfloat16* f0s; __nv_bfloat16* f1s; __nv_bfloat16* hl;
};

__global__ void __launch_bounds__(256) split_contrast_kernel(const SplitCtParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long nthreads = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long v = tid; v < p.rows * 16; v += nthreads) {           // 8 floats per item, both embedding matrices
    const long long row = v >> 4;
    const int which = (int)(v >> 3) & 1, c8 = (int)(v & 7) * 8;
    const float* src = (which ? p.f1 : p.f0) + row * 64 + c8;
    __nv_bfloat16* dst = (which ? p.f1s : p.f0s) + row * 192 + c8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float h[8], m[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h[i] = __bfloat162float(__float2bfloat16_rn(x[i]));
      m[i] = __bfloat162float(__float2bfloat16_rn(x[i] - h[i]));
      l[i] = (x[i] - h[i]) - m[i];
    }
    *reinterpret_cast<uint4*>(dst) = pack16(h, __nv_bfloat16());
    *reinterpret_cast<uint4*>(dst + 64) = pack16(m, __nv_bfloat16());
    *reinterpret_cast<uint4*>(dst + 128) = pack16(l, __nv_bfloat16());
  }
  for (long long v = tid; v < p.rows * 32; v += nthreads) {
    const long long row = v >> 5;
    const int c = (int)(v & 31);
    const float x = c < p.C ? __ldg(p.probs + row * p.C + c) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    p.hl[row * 64 + c] = h;
    p.hl[row * 64 + 32 + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

static size_t ct_split_bytes(long long rows) { return (size_t)rows * (384 + 384 + 128) + 3 * 1024; }

// carves the split copies out of the tail of the workspace and launches the pre-pass
static int ct_split(const char* fn, const float* f0, const float* f1, const float* probs, long long rows, int classes, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream, SplitCtParams* out) {
  const size_t need = kWsHeaderBytes + sizeof(float) * (size_t)((rows + kT - 1) / kT) * kMaxCl + ct_split_bytes(rows);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  auto up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  char* base = static_cast<char*>(workspace) + ((workspace_bytes - ct_split_bytes(rows)) & ~(size_t)1023);
  SplitCtParams sp{f0, f1, probs, rows, classes, reinterpret_cast<__nv_bfloat16*>(base),
                   reinterpret_cast<__nv_bfloat16*>(base + up((size_t)rows * 384)),
                   reinterpret_cast<__nv_bfloat16*>(base + 2 * up((size_t)rows * 384))};
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  cudaError_t e = launch_pdl(PDL_CONTRAST_FWD, split_contrast_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, dim3(1, 1, 1), sp);
  if (e != cudaSuccess) return fail((int)e, "%s: split launch: %s", fn, cudaGetErrorString(e));
  *out = sp;
  return 0;
}

}  // namespace

size_t contrast_tc_workspace_floats(long long rows) {
  const long long tiles = (rows + kT - 1) / kT;
  return (size_t)tiles * kMaxCl + (ct_split_bytes(rows) + 3) / 4;    // per-CTA loss partials + the split copies of fp32 storage
}

template <int NT, bool BWD, typename K>
static int ct_attr(const char* fn, K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ct_request<NT, BWD>());
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return 0;
}

static int ct_maps(CUtensorMap* m, const void* f0, const void* f1, const void* ph, long long rows, int nt) {
  const int terms = nt == 2 ? 3 : 1;                       // fp32 storage: [hi | mid | lo] per row (the forward reads hi, mid only)
  if (int e = tc::make_tmap_bf16_2d(&m[0], f0, (uint64_t)rows, 64 * terms, 128 * terms, kT, 64)) return e;
  if (int e = tc::make_tmap_bf16_2d(&m[1], f1, (uint64_t)rows, 64 * terms, 128 * terms, kT, 64)) return e;
  return tc::make_tmap_bf16_2d(&m[2], ph, (uint64_t)rows, 64, 128, kT, 64);
}

template <int NT, bool BWD, typename K>
static int ct_launch(const char* fn, int tag, K kernel, dim3 grid, int cluster, cudaStream_t stream, const CUtensorMap* m,
                     const ContrastTcParams& p) {
  cudaError_t e = launch_pdl(tag, kernel, grid, dim3(kCtThreads, 1, 1), smem_ct_request<NT, BWD>(), stream, dim3(1, (unsigned)cluster, 1), m[0],
                             m[1], m[2], p);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

// nt = 1: f0 / f1 bf16 [rows, 64];  nt = 2: fp32 storage, f0 / f1 / probs_hl are the split copies made by ct_split
static int contrast_fwd_tc_impl(int nt, const void* f0, const void* f1, const void* probs_hl, const float* probs, long long rows, int classes,
                                float temperature, float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u,
                                float lambda_c, float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_contrast_fwd[tcgen05]";
  ContrastTcParams p{};
  p.rows = rows; p.C = classes; p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.inv_tau = 1.0f / temperature; p.th = contrast_th; p.stats = stats; p.out = out_scalar; p.probs = probs;
  p.loss_u = loss_u; p.lambda_u = lambda_u; p.lambda_c = lambda_c; p.total_out = total_out;
  p.cluster = ct_cluster(rows, 1); p.dbg = debug_timing_buffer(PDL_CONTRAST_FWD);
  const size_t need = kWsHeaderBytes + sizeof(float) * (size_t)((rows + kT - 1) / kT) * kMaxCl;
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  p.grid_ticket = reinterpret_cast<unsigned*>(workspace) + 4;
  p.grid_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  CUtensorMap m[3];
  if (int e = ct_maps(m, f0, f1, probs_hl, rows, nt)) return e;
  const dim3 grid((unsigned)((rows + kT - 1) / kT), (unsigned)p.cluster, 1);
  static bool attr[2] = {false, false};
  if (nt == 2) {
    if (!attr[1]) { if (int e = ct_attr<2, false>(fn, contrast_tc_fwd_kernel<2>)) return e; attr[1] = true; }
    return ct_launch<2, false>(fn, PDL_CONTRAST_FWD, contrast_tc_fwd_kernel<2>, grid, p.cluster, stream, m, p);
  }
  if (!attr[0]) { if (int e = ct_attr<1, false>(fn, contrast_tc_fwd_kernel<1>)) return e; attr[0] = true; }
  return ct_launch<1, false>(fn, PDL_CONTRAST_FWD, contrast_tc_fwd_kernel<1>, grid, p.cluster, stream, m, p);
}

int contrast_fwd_tc(const void* f0, const void* f1, const void* probs_hl, long long rows, int classes, float temperature,
                    float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                    float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return contrast_fwd_tc_impl(1, f0, f1, probs_hl, nullptr, rows, classes, temperature, contrast_th, stats, out_scalar, loss_u, lambda_u,
                              lambda_c, total_out, workspace, workspace_bytes, stream);
}

// fp32 storage (dim 64, classes <= 32): split pre-pass into the workspace, then the tensor-core kernel on the split operands
int contrast_fwd_tc_f32(const float* f0, const float* f1, const float* probs, long long rows, int classes, float temperature,
                        float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                        float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SplitCtParams sp;
  if (int e = ct_split("b200ssl_contrast_fwd[tcgen05, fp32 storage]", f0, f1, probs, rows, classes, workspace, workspace_bytes, stream, &sp)) return e;
  return contrast_fwd_tc_impl(2, sp.f0s, sp.f1s, sp.hl, probs, rows, classes, temperature, contrast_th, stats, out_scalar, loss_u, lambda_u,
                              lambda_c, total_out, workspace, (size_t)(reinterpret_cast<char*>(sp.f0s) - static_cast<char*>(workspace)), stream);
}

static int contrast_bwd_tc_impl(int nt, const void* f0, const void* f1, const void* probs_hl, const float* probs, const float* stats,
                                long long rows, int classes, float temperature, float contrast_th, const float* upstream, float factor,
                                void* g0, void* g1, void* scale_grad, long long scale_numel, const float* scale_up, float scale_factor,
                                cudaStream_t stream) {
  const char* fn = "b200ssl_contrast_bwd[tcgen05]";
  ContrastTcParams p{};
  p.rows = rows; p.C = classes; p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.inv_tau = 1.0f / temperature; p.th = contrast_th; p.stats = const_cast<float*>(stats); p.probs = probs;
  p.upstream = upstream; p.factor = factor; p.g0 = g0; p.g1 = g1;
  p.cluster = ct_cluster(rows, scale_grad ? 3 : 2); p.dbg = debug_timing_buffer(PDL_CONTRAST_BWD);
  p.sgrad = scale_grad; p.snumel = scale_numel; p.sup = scale_up; p.sfactor = scale_factor;
  CUtensorMap m[3];
  if (int e = ct_maps(m, f0, f1, probs_hl, rows, nt)) return e;
  const dim3 grid((unsigned)((rows + kT - 1) / kT), (unsigned)p.cluster, scale_grad ? 3 : 2);
  static bool attr[2] = {false, false};
  if (nt == 2) {
    if (!attr[1]) { if (int e = ct_attr<2, true>(fn, contrast_tc_bwd_kernel<2>)) return e; attr[1] = true; }
    return ct_launch<2, true>(fn, PDL_CONTRAST_BWD, contrast_tc_bwd_kernel<2>, grid, p.cluster, stream, m, p);
  }
  if (!attr[0]) { if (int e = ct_attr<1, true>(fn, contrast_tc_bwd_kernel<1>)) return e; attr[0] = true; }
  return ct_launch<1, true>(fn, PDL_CONTRAST_BWD, contrast_tc_bwd_kernel<1>, grid, p.cluster, stream, m, p);
}

int contrast_bwd_tc(const void* f0, const void* f1, const void* probs_hl, const float* stats, long long rows, int classes,
                    float temperature, float contrast_th, const float* upstream, float factor, void* g0, void* g1,
                    void* scale_grad, long long scale_numel, const float* scale_up, float scale_factor,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  (void)workspace; (void)workspace_bytes;
  return contrast_bwd_tc_impl(1, f0, f1, probs_hl, nullptr, stats, rows, classes, temperature, contrast_th, upstream, factor, g0, g1,
                              scale_grad, scale_numel, scale_up, scale_factor, stream);
}

int contrast_bwd_tc_f32(const float* f0, const float* f1, const float* probs, const float* stats, long long rows, int classes,
                        float temperature, float contrast_th, const float* upstream, float factor, float* g0, float* g1,
                        float* scale_grad, long long scale_numel, const float* scale_up, float scale_factor,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SplitCtParams sp;
  if (int e = ct_split("b200ssl_contrast_bwd[tcgen05, fp32 storage]", f0, f1, probs, rows, classes, workspace, workspace_bytes, stream, &sp)) return e;
  return contrast_bwd_tc_impl(2, sp.f0s, sp.f1s, sp.hl, probs, stats, rows, classes, temperature, contrast_th, upstream, factor, g0, g1,
                              scale_grad, scale_numel, scale_up, scale_factor, stream);
}

}  // namespace b200ssl
