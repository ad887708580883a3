// Thin inline-PTX layer over the Blackwell (sm_100a) tensor-core machinery used
// by the similarity kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld), shared-memory + instruction descriptors.
// Encodings follow the PTX ISA for sm_100a; field layouts were cross-checked
// against CUTLASS' cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the driver entry point is fetched at run time)
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ssl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must neither hang the GPU box nor fault it.  When the
// bound is hit the CTA-wide sticky flag `*abort` is raised, every later wait returns at
// once and the kernel runs to completion with garbage results (which parity tests catch).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if ((it & 1023u) == 1023u && (*abort || it > (1u << 22))) { *abort = 1; return; }
  }
}

// one arrival per warp: every lane has done its part (and its own fences) before the warp-level sync
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// One lane of a converged warp (elect.sync).  The MMA warps run their loops with ALL lanes in warp-uniform control flow and
// issue under this predicate: descriptor arithmetic then stays in uniform registers (a loop inside `if (lane == 0)` makes
// the compiler rebuild every tcgen05.mma operand with an ELECT / R2UR sequence, ~17 dependent instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- proxies / fences ----------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// {x0, x1} = {a0 * c, a1 * c} with one packed FMUL2 (sm_100 f32x2 arithmetic; each lane rounds like a scalar FMUL)
__device__ __forceinline__ void mul_f32x2(float& x0, float& x1, float a0, float a1, float c) {
  asm("{\n\t.reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\t"
      "mul.rn.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t}"
      : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(c));
}

// ---- TMA -------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates (c0 = innermost element index, c1 = row index).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- TMEM ------------------------------------------------------------------------
// Allocate `ncols` (power of two >= 32) TMEM columns; executed by ONE full warp.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (lane_base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns into r[O .. O+16)
template <int O>
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[O + 0]), "=r"(r[O + 1]), "=r"(r[O + 2]), "=r"(r[O + 3]), "=r"(r[O + 4]), "=r"(r[O + 5]), "=r"(r[O + 6]),
        "=r"(r[O + 7]), "=r"(r[O + 8]), "=r"(r[O + 9]), "=r"(r[O + 10]), "=r"(r[O + 11]), "=r"(r[O + 12]), "=r"(r[O + 13]),
        "=r"(r[O + 14]), "=r"(r[O + 15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
// registers -> 32 lanes x 16 columns (thread i of the warp writes lane lane_base + i)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait for a load that was issued EARLIER than the statement before it (software pipelining): the registers are
// in/out operands, so no use of them can be scheduled above the wait by the compiler.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
        "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
        "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
        "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :
      : "memory");
}

// ---- descriptors -------------------------------------------------------------------
// K-major operand tile in the canonical 128-byte-swizzled layout (what a TMA box with
// CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte inner extent produces): rows of 128 B,
// 8-row groups 1024 B apart (SBO), 16-byte chunk index XOR (row % 8).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                           // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// byte address of element (row, 16-byte chunk) inside such a tile (tile base 1024-B aligned)
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}
// kind::f16 instruction descriptor: BF16 x BF16 -> FP32, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by one thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand read from tensor memory (lane = row of A, one 32-bit column = two
// consecutive bf16 of K, low half first).  Half the shared-memory traffic of the SS form and no proxy fence for A.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// mbarrier arrive once all tcgen05 ops issued so far by this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- host: tensor maps -------------------------------------------------------------
// [rows, cols] row-major bf16 matrix, box = box_rows x 64 elements (128 B), 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols);

}  // namespace tc
}  // namespace b200ssl
