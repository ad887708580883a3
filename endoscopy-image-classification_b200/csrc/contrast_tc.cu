// K6 on the 5th-generation tensor cores (bf16 embeddings, dim 64, classes <= 32):
// graph-contrastive loss and its gradient (code/comatch.py:199-213 + autograd).
//
// Per 128 x 128 tile (i rows of F0, j rows of F1) the MMA thread issues
//   S = F0_i F1_j^T                      4 x tcgen05.mma  (K = 64)
//   Q = Hi_i Hi_j^T + Hi_i Lo_j^T + Lo_i Hi_j^T      6 x tcgen05.mma  (K = 32 each)
// into TMEM (S: columns 0..127, Q: 128..255).  Q uses the bf16 hi/lo split of the fp32
// pseudo-label probabilities (probs_hl [rows, 64] = [hi(32) | lo(32)], written by the
// finalize kernel) so the 0.8 graph threshold sees ~fp32 accuracy although the operands
// are bf16.  256 epilogue threads (two per TMEM lane, 64 columns each) read S/Q with
// tcgen05.ld and do the exp / threshold / log math.
//
//   MODE_STATS : rowsum_i = sum_j exp(S/tau), qsum_i = sum_j Qm
//   MODE_LOSS  : loss_i, r_i  (exp/log only where Qm != 0: the graph is ~1% dense)
//   MODE_BWD   : dZ = P o (G - r) -> bf16 -> swizzled smem -> third MMA group
//                blockIdx.z = 0: dF0_i += dZ F1_j      (A = dZ K-major,  B = F1 tile MN-major)
//                blockIdx.z = 1: dF1_j += dZ^T F0_i    (A = dZ MN-major, B = F0 tile MN-major)
// The streamed dimension is split over gridDim.y CTAs; split partials are folded by the
// last CTA of a tile in split order (deterministic).
#include <math.h>

#include "common.cuh"
#include "tc.cuh"

namespace b200ssl {
namespace {

constexpr int kT = 128;                       // tile edge (UMMA M and N)
constexpr int kCtThreads = 320;               // warp 0 TMA/alloc, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t kTileF = kT * 128;         // 16 KB: [128][64] bf16
constexpr uint32_t kStageBytes = 2 * kTileF;  // F tile + probs hi/lo tile
constexpr uint32_t kSubZ = kT * 128;          // 16 KB: dZ columns [64*kb, +64)
constexpr uint32_t kSmemCt = 2 * kTileF + 2 * kStageBytes + 4 * kSubZ;   // 160 KB (dZ is kept as a bf16 hi + lo pair)
constexpr size_t kSmemCtRequest = kSmemCt + 1024 + 512;
constexpr uint32_t kTmemColsCt = 512;         // S 0..127, Q 128..255, dF accumulator 256..319

enum { MODE_STATS = 0, MODE_LOSS = 1, MODE_BWD = 2 };
enum { CB_OWN = 0, CB_KV_FULL = 1, CB_KV_EMPTY = 3, CB_SQ_FULL = 5, CB_SQ_EMPTY = 6, CB_Z_FULL = 7, CB_Z_EMPTY = 8,
       CB_ACC = 9, CB_COUNT = 10 };

struct ContrastTcParams {
  long long rows, rows_pad;
  int C, nsplit, tiles_per_split;
  float scale;            // log2(e) / tau
  float inv_tau, th;
  float* stats;           // [3][rows]: rowsum, qsum, r
  float* out; float* part; unsigned* tile_tickets; unsigned* grid_ticket; float* grid_part;
  const float* loss_u; float lambda_u, lambda_c; float* total_out;
  const float* upstream; float factor; void* g0; void* g1;
};

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

template <int MODE>
__global__ void __launch_bounds__(kCtThreads, 1)
contrast_tc_kernel(const __grid_constant__ CUtensorMap tm_f0, const __grid_constant__ CUtensorMap tm_f1,
                   const __grid_constant__ CUtensorMap tm_ph, const ContrastTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sOwnF = smem;                       // own tile: embeddings
  uint8_t* sOwnP = sOwnF + kTileF;             // own tile: probs hi/lo
  uint8_t* sStage = sOwnP + kTileF;            // 2 x (F tile, probs tile) of the streamed side
  uint8_t* sZ = sStage + 2 * kStageBytes;      // dZ tile: hi part (2 sub-tiles of 64 columns), then lo part
  uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + 4 * kSubZ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CB_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool colmode = (MODE == MODE_BWD) && blockIdx.z == 1;   // own tile = j rows of F1
  const int own_tile = blockIdx.x, split = blockIdx.y;
  const long long ntiles = (p.rows + kT - 1) / kT;
  const long long t0 = (long long)split * p.tiles_per_split;
  const int T = (int)(min(ntiles, t0 + p.tiles_per_split) - t0);

  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[CB_OWN], 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bars[CB_KV_FULL + s], 1);
      tc::mbar_init(&bars[CB_KV_EMPTY + s], 1);
    }
    tc::mbar_init(&bars[CB_SQ_FULL], 1);
    tc::mbar_init(&bars[CB_SQ_EMPTY], 256);
    tc::mbar_init(&bars[CB_Z_FULL], 256);
    tc::mbar_init(&bars[CB_Z_EMPTY], 1);
    tc::mbar_init(&bars[CB_ACC], 1);
    *abort_flag = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemColsCt);
  if (warp == 1 && lane == 0) {
    tc::tma_prefetch_desc(&tm_f0);
    tc::tma_prefetch_desc(&tm_f1);
    tc::tma_prefetch_desc(&tm_ph);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&bars[CB_OWN], 2 * kTileF);
      tc::tma_load_2d(sOwnF, colmode ? &tm_f1 : &tm_f0, 0, own_tile * kT, &bars[CB_OWN]);
      tc::tma_load_2d(sOwnP, &tm_ph, 0, own_tile * kT, &bars[CB_OWN]);
      for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        if (t >= 2) tc::mbar_wait(&bars[CB_KV_EMPTY + s], ((t >> 1) - 1) & 1, abort_flag);
        const int o0 = (int)((t0 + t) * kT);
        tc::mbar_arrive_expect_tx(&bars[CB_KV_FULL + s], kStageBytes);
        tc::tma_load_2d(sStage + s * kStageBytes, colmode ? &tm_f0 : &tm_f1, 0, o0, &bars[CB_KV_FULL + s]);
        tc::tma_load_2d(sStage + s * kStageBytes + kTileF, &tm_ph, 0, o0, &bars[CB_KV_FULL + s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc_sq = idesc_bf16(kT, kT, 0, 0);
      constexpr uint32_t idesc_row = idesc_bf16(kT, 64, 0, 1);    // dF0: A = dZ (K-major),  B = F1 tile (MN-major)
      constexpr uint32_t idesc_col = idesc_bf16(kT, 64, 1, 1);    // dF1: A = dZ (MN-major), B = F0 tile (MN-major)
      tc::mbar_wait(&bars[CB_OWN], 0, abort_flag);
      for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        uint8_t* stF = sStage + s * kStageBytes;
        uint8_t* stP = stF + kTileF;
        tc::mbar_wait(&bars[CB_KV_FULL + s], (t >> 1) & 1, abort_flag);
        if (t >= 1) tc::mbar_wait(&bars[CB_SQ_EMPTY], (t - 1) & 1, abort_flag);
        tc::tcgen05_fence_after();
        // rows of S/Q are always the i side (F0), columns the j side (F1)
        const uint64_t aF = tc::smem_desc_sw128(tc::smem_u32(colmode ? stF : sOwnF));
        const uint64_t bF = tc::smem_desc_sw128(tc::smem_u32(colmode ? sOwnF : stF));
        const uint64_t aP = tc::smem_desc_sw128(tc::smem_u32(colmode ? stP : sOwnP));
        const uint64_t bP = tc::smem_desc_sw128(tc::smem_u32(colmode ? sOwnP : stP));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem, aF + 2 * k, bF + 2 * k, idesc_sq, k > 0);
        // hi.hi + hi.lo + lo.hi ; hi = columns 0..31 (byte 0), lo = columns 32..63 (byte 64 -> +4)
#pragma unroll
        for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 2 * k, bP + 2 * k, idesc_sq, k > 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 2 * k, bP + 4 + 2 * k, idesc_sq, true);
#pragma unroll
        for (int k = 0; k < 2; ++k) tc::mma_bf16_ss(tmem + kT, aP + 4 + 2 * k, bP + 2 * k, idesc_sq, true);
        tc::mma_commit(&bars[CB_SQ_FULL]);
        if (MODE != MODE_BWD) {
          tc::mma_commit(&bars[CB_KV_EMPTY + s]);
        } else {
          tc::mbar_wait(&bars[CB_Z_FULL], t & 1, abort_flag);
          tc::tcgen05_fence_after();
          const uint32_t zaddr = tc::smem_u32(sZ);
          if (!colmode) {
            // dF0[i, d] += sum_j dZ[i, j] F1[j, d]
            const uint64_t bB = smem_desc_sw128_mn(tc::smem_u32(stF), 0);
#pragma unroll
            for (int part = 0; part < 2; ++part)       // dZ = hi + lo (two bf16 terms ~ 16 mantissa bits)
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t aZ = tc::smem_desc_sw128(zaddr + (2 * part + (k >> 2)) * kSubZ) + 2 * (k & 3);
                tc::mma_bf16_ss(tmem + 2 * kT, aZ, bB + 128 * k, idesc_row, (t | k | part) != 0);
              }
          } else {
            // dF1[j, d] += sum_i dZ[i, j] F0[i, d]
            const uint64_t bB = smem_desc_sw128_mn(tc::smem_u32(stF), 0);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              const uint64_t aZ = smem_desc_sw128_mn(zaddr + 2 * part * kSubZ, kSubZ);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                tc::mma_bf16_ss(tmem + 2 * kT, aZ + 128 * k, bB + 128 * k, idesc_col, (t | k | part) != 0);
            }
          }
          tc::mma_commit(&bars[CB_KV_EMPTY + s]);
          tc::mma_commit(&bars[CB_Z_EMPTY]);
        }
      }
      if (MODE == MODE_BWD) tc::mma_commit(&bars[CB_ACC]);
    }
  } else {
    // ================= epilogue: 2 threads per TMEM lane (row i), 64 columns (j) each =================
    const int ew = warp - 2;
    const int quarter = warp & 3, half = ew >> 2;
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float a0 = 0.f, a1 = 0.f;                  // (rowsum, qsum) or (loss, rr)
    float rs = 1.f, qs = 1.f, r_i = 0.f;
    const float inv_rows = 1.0f / (float)p.rows;
    long long gi = colmode ? 0 : (long long)own_tile * kT + r_in;   // global i of this thread's row
    if (MODE != MODE_STATS && !colmode && gi < p.rows) {
      rs = p.stats[gi];
      qs = p.stats[p.rows + gi];
      if (MODE == MODE_BWD) r_i = p.stats[2 * p.rows + gi];
    }
    for (int t = 0; t < T; ++t) {
      const long long o0 = (t0 + t) * kT;
      const long long j0 = colmode ? (long long)own_tile * kT : o0;
      if (colmode) {
        gi = o0 + r_in;
        rs = qs = 1.f; r_i = 0.f;
        if (gi < p.rows) { rs = p.stats[gi]; qs = p.stats[p.rows + gi]; r_i = p.stats[2 * p.rows + gi]; }
      }
      tc::mbar_wait(&bars[CB_SQ_FULL], t & 1, abort_flag);
      if (MODE == MODE_BWD && t >= 1) tc::mbar_wait(&bars[CB_Z_EMPTY], (t - 1) & 1, abort_flag);
      tc::tcgen05_fence_after();
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int col0 = half * 64 + c2 * 32;
        uint32_t sv[32], qv[32];
        tc::tmem_ld_32x32(lane_addr + col0, sv);
        tc::tmem_ld_32x32(lane_addr + kT + col0, qv);
        tc::tmem_ld_wait();
        uint32_t packed[16], packed_lo[16];
        float dz_even = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const long long gj = j0 + col0 + j;
          const bool ok = (gi < p.rows) && (gj < p.rows);
          float q = __uint_as_float(qv[j]);
          q = (gi == gj) ? 1.f : q;                         // fill_diagonal_(1)   comatch.py:205
          const float qm = (ok && q >= p.th) ? q : 0.f;     // pos_mask            :206-208
          float dz = 0.f;
          if (MODE == MODE_STATS) {
            const float e = ok ? exp2f(__uint_as_float(sv[j]) * p.scale) : 0.f;    // :200
            a0 += e;
            a1 += qm;
          } else if (MODE == MODE_LOSS) {
            if (qm != 0.f) {
              const float P = exp2f(__uint_as_float(sv[j]) * p.scale) / rs;          // :201
              const float qn = qm / qs;                                               // :209
              a0 -= __logf(P + 1e-7f) * qn;                                           // :212
              a1 += qn * (P / (P + 1e-7f));
            }
          } else {
            if (ok) {
              const float P = exp2f(__uint_as_float(sv[j]) * p.scale) / rs;
              const float G = (qm != 0.f) ? -(qm / qs) / (P + 1e-7f) * inv_rows : 0.f;
              dz = P * (G - r_i);
            }
            if (j & 1) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(dz_even, dz);
              const __nv_bfloat162 l = __floats2bfloat162_rn(dz_even - __low2float(h), dz - __high2float(h));
              packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
              packed_lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l);
            } else {
              dz_even = dz;
            }
          }
        }
        if (MODE == MODE_BWD) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = c2 * 4 + q4;               // 16-byte chunk inside the 64-column sub-tile `half`
            const uint4 v = make_uint4(packed[4 * q4], packed[4 * q4 + 1], packed[4 * q4 + 2], packed[4 * q4 + 3]);
            const uint4 vl = make_uint4(packed_lo[4 * q4], packed_lo[4 * q4 + 1], packed_lo[4 * q4 + 2], packed_lo[4 * q4 + 3]);
            *reinterpret_cast<uint4*>(sZ + half * kSubZ + tc::sw128_offset(r_in, chunk)) = v;
            *reinterpret_cast<uint4*>(sZ + (2 + half) * kSubZ + tc::sw128_offset(r_in, chunk)) = vl;
          }
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars[CB_SQ_EMPTY]);
      if (MODE == MODE_BWD) {
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&bars[CB_Z_FULL]);
      }
    }
    if (MODE != MODE_BWD) {
      // per-row partials of this (split, half): part[value][split*2+half][rows_pad]
      const size_t vstride = (size_t)2 * p.nsplit * p.rows_pad;
      float* base = p.part + (size_t)(split * 2 + half) * p.rows_pad + (size_t)own_tile * kT + r_in;
      base[0] = a0;
      base[vstride] = a1;
    } else {
      tc::mbar_wait(&bars[CB_ACC], 0, abort_flag);
      tc::tcgen05_fence_after();
      if (half == 0) {
        const float up = (p.upstream ? *p.upstream : 1.f) * p.factor * p.inv_tau;
        const long long grow = (long long)own_tile * kT + r_in;
        __nv_bfloat16* out = static_cast<__nv_bfloat16*>(colmode ? p.g1 : p.g0);
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t av[32];
          tc::tmem_ld_32x32(lane_addr + 2 * kT + c2 * 32, av);
          tc::tmem_ld_wait();
          if (p.nsplit == 1) {
            if (grow < p.rows) {
              uint4* o = reinterpret_cast<uint4*>(out + grow * 64 + c2 * 32);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(av[8 * q4 + e]) * up;
                o[q4] = pack16(f, __nv_bfloat16());
              }
            }
          } else {
            float4* o = reinterpret_cast<float4*>(p.part + (((size_t)(colmode ? 1 : 0) * p.nsplit + split) * p.rows_pad + grow) * 64 + c2 * 32);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              o[q4] = make_float4(__uint_as_float(av[4 * q4]), __uint_as_float(av[4 * q4 + 1]), __uint_as_float(av[4 * q4 + 2]),
                                  __uint_as_float(av[4 * q4 + 3]));
          }
        }
      }
    }
    tc::tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem, kTmemColsCt);
  }
  // ---------------- fold the split partials (last CTA of the tile) ----------------
  const int tid = threadIdx.x;
  const long long i0 = (long long)own_tile * kT;
  const int nown = (int)min((long long)kT, p.rows - i0);
  if (MODE == MODE_BWD && p.nsplit == 1) return;
  __shared__ bool s_last;
  unsigned* ticket = p.tile_tickets + (MODE == MODE_BWD ? 2 + (colmode ? 1 : 0) : MODE) * gridDim.x + own_tile;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)p.nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (MODE == MODE_BWD) {
    const float up = (p.upstream ? *p.upstream : 1.f) * p.factor * p.inv_tau;
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(colmode ? p.g1 : p.g0);
    const float* base = p.part + (size_t)(colmode ? 1 : 0) * p.nsplit * p.rows_pad * 64;
    fold_splits_vec4(reinterpret_cast<const float4*>(base + (size_t)i0 * 64), (size_t)p.rows_pad * 64 / 4, p.nsplit,
                     nown * 64 / 4, [&](int i, float4 v) {
                       const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * up, v.y * up);
                       const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * up, v.w * up);
                       uint2 w;
                       w.x = *reinterpret_cast<const uint32_t*>(&lo);
                       w.y = *reinterpret_cast<const uint32_t*>(&hi);
                       *reinterpret_cast<uint2*>(out + i0 * 64 + 4 * (size_t)i) = w;
                     });
    if (tid == 0) *ticket = 0u;
    return;
  }
  __shared__ float s_fold[2][kT];
  const size_t vstride = (size_t)2 * p.nsplit * p.rows_pad;
#pragma unroll
  for (int w = 0; w < 2; ++w)
    fold_splits_vec4(reinterpret_cast<const float4*>(p.part + w * vstride + i0), (size_t)p.rows_pad / 4, 2 * p.nsplit, kT / 4,
                     [&](int i, float4 v) {
                       s_fold[w][4 * i] = v.x; s_fold[w][4 * i + 1] = v.y; s_fold[w][4 * i + 2] = v.z; s_fold[w][4 * i + 3] = v.w;
                     });
  __syncthreads();
  float lsum = 0.f;
  for (int r = tid; r < nown; r += blockDim.x) {
    if (MODE == MODE_STATS) {
      p.stats[i0 + r] = s_fold[0][r];
      p.stats[p.rows + i0 + r] = s_fold[1][r];
    } else {
      p.stats[2 * p.rows + i0 + r] = -s_fold[1][r] / (float)p.rows;
      lsum += s_fold[0][r];
    }
  }
  if (tid == 0) *ticket = 0u;
  if (MODE == MODE_STATS) return;
  // loss: rows of this tile, then tiles (last tile-finisher folds in tile order)
  __shared__ float s_w[kCtThreads / 32];
  __shared__ bool s_glast;
  lsum = warp_sum(lsum);
  if (lane == 0) s_w[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int w = 0; w < kCtThreads / 32; ++w) c += s_w[w];
    p.grid_part[own_tile] = c;
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

int ct_nsplit(long long rows, int modes, int* tiles_per_split) {
  const long long tiles = (rows + kT - 1) / kT;
  long long want = (kNumSMs + tiles * modes - 1) / (tiles * modes);
  if (want < 1) want = 1;
  if (want > tiles) want = tiles;
  const long long tps = (tiles + want - 1) / want;
  if (tiles_per_split) *tiles_per_split = (int)tps;
  return (int)((tiles + tps - 1) / tps);
}

}  // namespace

size_t contrast_tc_workspace_floats(long long rows) {
  const long long tiles = (rows + kT - 1) / kT;
  const long long rows_pad = tiles * kT;
  const size_t fwd = (size_t)4 * ct_nsplit(rows, 1, nullptr) * rows_pad;
  const int nb = ct_nsplit(rows, 2, nullptr);
  const size_t bwd = nb > 1 ? (size_t)2 * nb * rows_pad * 64 : 0;
  return (size_t)((tiles + 3) & ~3LL) + (fwd > bwd ? fwd : bwd);
}

static int ct_setup(const char* fn, ContrastTcParams& p, int modes, void* workspace, size_t workspace_bytes) {
  const long long tiles = (p.rows + kT - 1) / kT;
  p.rows_pad = tiles * kT;
  p.nsplit = ct_nsplit(p.rows, modes, &p.tiles_per_split);
  const size_t need = kWsHeaderBytes + sizeof(float) * contrast_tc_workspace_floats(p.rows);
  if (workspace_bytes < need) return fail(B200SSL_E_WORKSPACE, "%s: workspace %zu < %zu bytes", fn, workspace_bytes, need);
  if ((size_t)tiles * 4 * sizeof(unsigned) > kWsTicket2Bytes) return fail(B200SSL_E_SHAPE, "%s: too many row tiles", fn);
  p.tile_tickets = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kWsTicketBytes);
  p.grid_ticket = reinterpret_cast<unsigned*>(workspace) + 4;
  p.grid_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWsHeaderBytes);
  p.part = p.grid_part + ((tiles + 3) & ~3LL);
  return 0;
}

template <typename K>
static int ct_attr(const char* fn, K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCtRequest);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
  return 0;
}

static int ct_maps(CUtensorMap* m, const void* f0, const void* f1, const void* ph, long long rows) {
  if (int e = tc::make_tmap_bf16_2d(&m[0], f0, (uint64_t)rows, 64, 128, kT, 64)) return e;
  if (int e = tc::make_tmap_bf16_2d(&m[1], f1, (uint64_t)rows, 64, 128, kT, 64)) return e;
  return tc::make_tmap_bf16_2d(&m[2], ph, (uint64_t)rows, 64, 128, kT, 64);
}

int contrast_fwd_tc(const void* f0, const void* f1, const void* probs_hl, long long rows, int classes, float temperature,
                    float contrast_th, float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                    float* total_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_contrast_fwd[tcgen05]";
  ContrastTcParams p{};
  p.rows = rows; p.C = classes; p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.inv_tau = 1.0f / temperature; p.th = contrast_th; p.stats = stats; p.out = out_scalar;
  p.loss_u = loss_u; p.lambda_u = lambda_u; p.lambda_c = lambda_c; p.total_out = total_out;
  if (int e = ct_setup(fn, p, 1, workspace, workspace_bytes)) return e;
  CUtensorMap m[3];
  if (int e = ct_maps(m, f0, f1, probs_hl, rows)) return e;
  static bool attr = false;
  if (!attr) {
    if (int e = ct_attr(fn, contrast_tc_kernel<MODE_STATS>)) return e;
    if (int e = ct_attr(fn, contrast_tc_kernel<MODE_LOSS>)) return e;
    attr = true;
  }
  dim3 grid((unsigned)((rows + kT - 1) / kT), (unsigned)p.nsplit, 1);
  contrast_tc_kernel<MODE_STATS><<<grid, kCtThreads, kSmemCtRequest, stream>>>(m[0], m[1], m[2], p);
  contrast_tc_kernel<MODE_LOSS><<<grid, kCtThreads, kSmemCtRequest, stream>>>(m[0], m[1], m[2], p);
  return check_launch(fn);
}

int contrast_bwd_tc(const void* f0, const void* f1, const void* probs_hl, const float* stats, long long rows, int classes,
                    float temperature, float contrast_th, const float* upstream, float factor, void* g0, void* g1,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* fn = "b200ssl_contrast_bwd[tcgen05]";
  ContrastTcParams p{};
  p.rows = rows; p.C = classes; p.scale = (float)(1.4426950408889634 / (double)temperature);
  p.inv_tau = 1.0f / temperature; p.th = contrast_th; p.stats = const_cast<float*>(stats);
  p.upstream = upstream; p.factor = factor; p.g0 = g0; p.g1 = g1;
  if (int e = ct_setup(fn, p, 2, workspace, workspace_bytes)) return e;
  CUtensorMap m[3];
  if (int e = ct_maps(m, f0, f1, probs_hl, rows)) return e;
  static bool attr = false;
  if (!attr) {
    if (int e = ct_attr(fn, contrast_tc_kernel<MODE_BWD>)) return e;
    attr = true;
  }
  dim3 grid((unsigned)((rows + kT - 1) / kT), (unsigned)p.nsplit, 2);
  contrast_tc_kernel<MODE_BWD><<<grid, kCtThreads, kSmemCtRequest, stream>>>(m[0], m[1], m[2], p);
  return check_launch(fn);
}

}  // namespace b200ssl
