// Row exchanges of the rank-sharded memory bank over NVLink peer memory.
//
// Every rank owns one "arena" (cudaMalloc, exported through CUDA IPC and mapped by all peers of the node):
//
//   [0, 1 KB)      flags[exchange][source rank]  u64, written by the PEERS (st.release.sys), monotonic epochs
//   [1 KB, 4 KB)   local control: epoch[exchange], tickets, timeout counter -- only this rank touches it
//   [4 KB, ...)    staging regions, one per exchange: [parity 2][source rank R][slot bytes]
//
// One launch is the whole collective: push my rows into every peer's staging slot with plain 16-byte stores
// over NVLink, publish the epoch to the peer's flag, wait for the peers' flags and copy (all-gather) or fold
// in rank order (reduce-scatter) the staged rows into an ordinary local tensor.  At the bank's sizes
// (tens of KB per rank) this is one NVLink round trip (~3 us) where a NCCL collective costs its launch protocol
// (~20 us); the epochs live in device memory, so a CUDA graph replays it unchanged.
//
// Slot reuse: staging is double buffered on the epoch's parity.  A peer can run at most ONE epoch of the same
// exchange ahead of this rank (it needs this rank's flag of epoch e+1 to go further), so the slots of epoch e
// are never overwritten before the local copy-out of epoch e has run.
#include "common.cuh"
#include "peer.cuh"

namespace b200ssl {
namespace {

using namespace peer;

struct PeerParams {
  const uint8_t* src0; size_t bytes0;               // all-gather: my block = [src0 ; src1];  reduce-scatter: src0 = [R][bytes] fp32
  const uint8_t* src1; size_t bytes1;
  uint8_t* out;
  size_t bytes;                                     // per rank
  uint8_t* const* arenas;                           // device array [world] of arena bases (peer mapped), own included
  size_t region, slot;
  int x, rank, world, chunks;
};

__device__ __forceinline__ uint4 load_src(const PeerParams& p, size_t off) {     // 16-byte piece of [src0 ; src1]
  return off < p.bytes0 ? *reinterpret_cast<const uint4*>(p.src0 + off) : *reinterpret_cast<const uint4*>(p.src1 + (off - p.bytes0));
}

// grid (chunks, world): CTA (c, y) pushes chunk c to destination y, then serves chunk c of SOURCE y.
template <bool REDUCE>
__global__ void __launch_bounds__(kPeerThreads) peer_exchange_kernel(const PeerParams p) {
  pdl_launch_dependents();
  pdl_wait();
  uint8_t* mine = p.arenas[p.rank];
  LocalCtl* ctl = reinterpret_cast<LocalCtl*>(mine + kFlagBytes);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[p.x]) + 1;
  const size_t parity = (size_t)(epoch & 1);
  const int c = blockIdx.x, y = blockIdx.y, tid = threadIdx.x;
  const size_t c0 = (size_t)c * kChunkBytes;
  const size_t c1 = c0 + kChunkBytes < p.bytes ? c0 + kChunkBytes : p.bytes;

  // ---- push: my rows for destination y -> y's staging slot [parity][rank]
  if (y != p.rank) {
    uint8_t* dst = p.arenas[y] + p.region + (parity * p.world + p.rank) * p.slot;
    const size_t sbase = REDUCE ? (size_t)y * p.bytes : 0;       // reduce-scatter sends chunk y of the source
    for (size_t o = c0 + (size_t)tid * 16; o < c1; o += (size_t)kPeerThreads * 16)
      *reinterpret_cast<uint4*>(dst + o) = REDUCE ? *reinterpret_cast<const uint4*>(p.src0 + sbase + o) : load_src(p, o);
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();                                    // cumulative: the CTA's stores are visible before the flag
      if (atomicAdd(&ctl->pushed[p.x][y], 1u) == (unsigned)p.chunks - 1) {
        ctl->pushed[p.x][y] = 0;
        st_release_sys(reinterpret_cast<unsigned long long*>(p.arenas[y]) + p.x * kMaxWorld + p.rank, epoch);
      }
    }
  }

  // ---- receive
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(mine) + p.x * kMaxWorld;
  if (!REDUCE) {
    // all-gather: rows of source y -> out[y]
    uint8_t* o_base = p.out + (size_t)y * p.bytes;
    if (y == p.rank) {
      for (size_t o = c0 + (size_t)tid * 16; o < c1; o += (size_t)kPeerThreads * 16) *reinterpret_cast<uint4*>(o_base + o) = load_src(p, o);
    } else {
      if (tid == 0) wait_flag(flags + y, epoch, ctl);
      __syncthreads();
      const uint8_t* st = mine + p.region + (parity * p.world + y) * p.slot;
      for (size_t o = c0 + (size_t)tid * 16; o < c1; o += (size_t)kPeerThreads * 16) *reinterpret_cast<uint4*>(o_base + o) = ld_cg(st + o);
    }
  } else {
    // reduce-scatter: CTA (c, y) folds the y-th part of chunk c over all sources in rank order
    if (tid < p.world && tid != p.rank) wait_flag(flags + tid, epoch, ctl);
    __syncthreads();
    const size_t vecs = (c1 - c0) / 16;
    const size_t v0 = vecs * y / p.world, v1 = vecs * (y + 1) / p.world;
    for (size_t v = v0 + tid; v < v1; v += kPeerThreads) {
      const size_t o = c0 + v * 16;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < p.world; ++s) {
        const uint4 u = (s == p.rank) ? *reinterpret_cast<const uint4*>(p.src0 + (size_t)s * p.bytes + o)
                                      : ld_cg(mine + p.region + (parity * p.world + s) * p.slot + o);
        const float4 f = *reinterpret_cast<const float4*>(&u);
        if (s == 0) acc = f;
        else { acc.x = __fadd_rn(acc.x, f.x); acc.y = __fadd_rn(acc.y, f.y); acc.z = __fadd_rn(acc.z, f.z); acc.w = __fadd_rn(acc.w, f.w); }
      }
      *reinterpret_cast<float4*>(p.out + o) = acc;
    }
  }

  // ---- the last CTA of the launch commits the epoch
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctl->done[p.x], 1u) == gridDim.x * gridDim.y - 1) {
      ctl->done[p.x] = 0;
      *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[p.x]) = epoch;
    }
  }
}

int check_common(const char* fn, const void* src, const void* out, size_t bytes, const void* arenas, size_t region, size_t slot,
                 int x, int rank, int world) {
  if (!src || !out || !arenas) return fail(B200SSL_E_NULL, "%s: NULL pointer", fn);
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world) return fail(B200SSL_E_ARG, "%s: rank %d of %d (2..%d ranks)", fn, rank, world, kMaxWorld);
  if (x < 0 || x >= kMaxExchange) return fail(B200SSL_E_ARG, "%s: exchange id %d outside [0, %d)", fn, x, kMaxExchange);
  if (bytes == 0 || (bytes & 15u) || slot < bytes || (slot & 15u) || region < kCtlBytes || (region & 15u))
    return fail(B200SSL_E_ALIGN, "%s: bytes per rank %zu (slot %zu, region offset %zu) must be non-zero multiples of 16, slot >= bytes, region >= %zu",
                fn, bytes, slot, region, kCtlBytes);
  if ((reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return fail(B200SSL_E_ALIGN, "%s: 16-byte aligned rows required", fn);
  return 0;
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

extern "C" {

size_t b200ssl_peer_control_bytes(void) { return kCtlBytes; }

int b200ssl_peer_alloc(size_t bytes, void** arena, void* ipc_handle_64) {
  const char* fn = "b200ssl_peer_alloc";
  if (!arena || !ipc_handle_64) return fail(B200SSL_E_NULL, "%s: NULL pointer", fn);
  if (bytes < kCtlBytes) return fail(B200SSL_E_ARG, "%s: an arena holds at least its %zu control bytes", fn, kCtlBytes);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
  }
  memcpy(ipc_handle_64, &h, sizeof(h));
  *arena = p;
  return 0;
}

int b200ssl_peer_open(const void* ipc_handle_64, void** arena) {
  const char* fn = "b200ssl_peer_open";
  if (!arena || !ipc_handle_64) return fail(B200SSL_E_NULL, "%s: NULL pointer", fn);
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(arena, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaIpcOpenMemHandle: %s (peer access between the two GPUs is required)", fn, cudaGetErrorString(e));
  return 0;
}

int b200ssl_peer_close(void* arena) {
  if (!arena) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(arena);
  return e == cudaSuccess ? 0 : fail((int)e, "b200ssl_peer_close: %s", cudaGetErrorString(e));
}

int b200ssl_peer_free(void* arena) {
  if (!arena) return 0;
  cudaError_t e = cudaFree(arena);
  return e == cudaSuccess ? 0 : fail((int)e, "b200ssl_peer_free: %s", cudaGetErrorString(e));
}

int b200ssl_peer_timeouts(const void* arena, uint32_t* count) {
  if (!arena || !count) return fail(B200SSL_E_NULL, "b200ssl_peer_timeouts: NULL pointer");
  const LocalCtl* ctl = reinterpret_cast<const LocalCtl*>(static_cast<const uint8_t*>(arena) + kFlagBytes);
  cudaError_t e = cudaMemcpy(count, &ctl->timeouts, sizeof(uint32_t), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? 0 : fail((int)e, "b200ssl_peer_timeouts: %s", cudaGetErrorString(e));
}

int b200ssl_peer_all_gather(const void* src0, size_t bytes0, const void* src1, size_t bytes1, void* out, void* const* arenas,
                            size_t region_offset, size_t slot_bytes, int32_t exchange_id, int32_t rank, int32_t world, void* stream) {
  const char* fn = "b200ssl_peer_all_gather";
  const size_t bytes = bytes0 + bytes1;
  if (int rc = check_common(fn, src0, out, bytes, arenas, region_offset, slot_bytes, exchange_id, rank, world)) return rc;
  if ((bytes0 & 15u) || (bytes1 && (!src1 || (reinterpret_cast<uintptr_t>(src1) & 15u))))
    return fail(B200SSL_E_ALIGN, "%s: both source segments must be 16-byte aligned multiples of 16 bytes", fn);
  PeerParams p{static_cast<const uint8_t*>(src0), bytes0, static_cast<const uint8_t*>(src1), bytes1, static_cast<uint8_t*>(out), bytes,
               reinterpret_cast<uint8_t* const*>(arenas), region_offset, slot_bytes, exchange_id, rank, world,
               (int)((bytes + kChunkBytes - 1) / kChunkBytes)};
  cudaError_t e = launch_pdl(PDL_PEER, peer_exchange_kernel<false>, dim3((unsigned)p.chunks, (unsigned)world, 1), dim3(kPeerThreads, 1, 1), 0,
                             as_stream(stream), dim3(1, 1, 1), p);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

int b200ssl_peer_reduce_scatter_f32(const float* src, float* out, int64_t count_per_rank, void* const* arenas, size_t region_offset,
                                    size_t slot_bytes, int32_t exchange_id, int32_t rank, int32_t world, void* stream) {
  const char* fn = "b200ssl_peer_reduce_scatter_f32";
  if (count_per_rank <= 0) return fail(B200SSL_E_SHAPE, "%s: count_per_rank %lld", fn, (long long)count_per_rank);
  const size_t bytes = (size_t)count_per_rank * 4;
  if (int rc = check_common(fn, src, out, bytes, arenas, region_offset, slot_bytes, exchange_id, rank, world)) return rc;
  PeerParams p{reinterpret_cast<const uint8_t*>(src), bytes * world, nullptr, 0, reinterpret_cast<uint8_t*>(out), bytes,
               reinterpret_cast<uint8_t* const*>(arenas), region_offset, slot_bytes, exchange_id, rank, world,
               (int)((bytes + kChunkBytes - 1) / kChunkBytes)};
  cudaError_t e = launch_pdl(PDL_PEER, peer_exchange_kernel<true>, dim3((unsigned)p.chunks, (unsigned)world, 1), dim3(kPeerThreads, 1, 1), 0,
                             as_stream(stream), dim3(1, 1, 1), p);
  if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchKernelEx: %s", fn, cudaGetErrorString(e));
  return check_launch(fn);
}

}  // extern "C"
