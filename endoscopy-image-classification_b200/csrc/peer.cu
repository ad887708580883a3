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

// ---- enqueue of this rank's block into a peer-memory resident bank (b200ssl_bank_enqueue_peer) -----------------
// grid (ceil(n / 32), ndst): CTA (c, d) stages rows [32c, 32c+32) of the block [feats_u_w ; feats_x] /
// [probs_orig ; onehot(targets_x)] in shared memory and writes them into destination d -- copy (rank + d) % world of a
// replicated ring, or the shard that owns the rows.  Wide (it leaves the row kernel's 8 CTAs: remote stores are
// throughput-limited per SM) and meant for a side stream: the step's losses do not depend on it.
constexpr int kEnqRows = 32;
constexpr int kEnqThreads = 256;

struct EnqueuePeerParams {
  const __nv_bfloat16* fu; const __nv_bfloat16* fx; const float* po; const long long* tx;
  long long n_u, n_x, K, shard_rows;
  int C, rank, world, replicated;
  uint8_t* const* arenas; size_t qf_off, qp_off, qpt_off;
  long long* ptr_state;
};

__global__ void __launch_bounds__(kEnqThreads) bank_enqueue_peer_kernel(const EnqueuePeerParams p) {
  __shared__ uint4 sF[kEnqRows * 8];                         // [32][64] bf16
  __shared__ __align__(16) __nv_bfloat16 sP[kEnqRows * 32];  // [32][C] bf16, row-major like queue_probs
  const int tid = threadIdx.x, C = p.C;
  uint8_t* mine = p.arenas[p.rank];
  LocalCtl* ctl = local_ctl(mine);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[kXEnqueueDone]) + 1;
  const long long ptr0 = *reinterpret_cast<volatile long long*>(p.ptr_state);
  const long long n = p.n_u + p.n_x;
  const long long r0 = (long long)blockIdx.x * kEnqRows;
  const int nrows = (int)min((long long)kEnqRows, n - r0);
  // nobody may still be reading the rows this step overwrites: every rank's smoothing pass has published "reads done"
  if (tid < p.world && tid != p.rank) wait_flag(flag_of(mine, kXSmoothDone, tid), epoch, ctl);
  // stage the tile (local reads)
  for (int i = tid; i < nrows * 8; i += kEnqThreads) {
    const long long r = r0 + (i >> 3);
    const __nv_bfloat16* src = r < p.n_u ? p.fu + r * 64 : p.fx + (r - p.n_u) * 64;
    sF[i] = *reinterpret_cast<const uint4*>(src + (i & 7) * 8);
  }
  for (int i = tid; i < nrows * C; i += kEnqThreads) {
    const int rr = i / C, c = i - rr * C;
    const long long r = r0 + rr;
    sP[i] = __float2bfloat16_rn(r < p.n_u ? p.po[r * C + c] : (c == (int)p.tx[r - p.n_u] ? 1.f : 0.f));   // comatch.py:188-189
  }
  __syncthreads();
  // destination of the tile's first row
  const long long g0 = (ptr0 + (long long)p.rank * n + r0) % p.K;       // comatch.py:194-196, rank-major blocks
  const long long ld = p.replicated ? p.K : p.shard_rows;
  auto locate = [&](long long g, long long* row) -> uint8_t* {
    if (p.replicated) { *row = g; return p.arenas[(p.rank + blockIdx.y) % p.world]; }
    const int s = (int)(g / p.shard_rows);
    *row = g - (long long)s * p.shard_rows;
    return p.arenas[s];
  };
  long long row0;
  uint8_t* base0 = locate(g0, &row0);
  if (nrows == kEnqRows && (g0 & 7) == 0 && g0 + kEnqRows <= p.K && row0 + kEnqRows <= ld) {
    // aligned tile inside one shard: 128-bit stores only
    uint4* qf = reinterpret_cast<uint4*>(base0 + p.qf_off) + row0 * 8;
    for (int i = tid; i < kEnqRows * 8; i += kEnqThreads) qf[i] = sF[i];
    uint4* qp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base0 + p.qp_off) + row0 * C);
    for (int i = tid; i < kEnqRows * C / 8; i += kEnqThreads) qp[i] = reinterpret_cast<const uint4*>(sP)[i];
    __nv_bfloat16* qpt = reinterpret_cast<__nv_bfloat16*>(base0 + p.qpt_off);
    for (int i = tid; i < C * (kEnqRows / 8); i += kEnqThreads) {
      const int c = i / (kEnqRows / 8), j = i - c * (kEnqRows / 8);
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = sP[(8 * j + k) * C + c];
      *reinterpret_cast<uint4*>(qpt + (size_t)c * ld + row0 + 8 * j) = *reinterpret_cast<const uint4*>(v);
    }
  } else {
    for (int i = tid; i < nrows * 8; i += kEnqThreads) {
      long long row;
      uint8_t* base = locate((g0 + (i >> 3)) % p.K, &row);
      reinterpret_cast<uint4*>(base + p.qf_off)[row * 8 + (i & 7)] = sF[i];
    }
    for (int i = tid; i < nrows * C; i += kEnqThreads) {
      const int rr = i / C, c = i - rr * C;
      long long row;
      uint8_t* base = locate((g0 + rr) % p.K, &row);
      reinterpret_cast<__nv_bfloat16*>(base + p.qp_off)[row * C + c] = sP[i];
      reinterpret_cast<__nv_bfloat16*>(base + p.qpt_off)[(size_t)c * ld + row] = sP[i];
    }
  }
  // last CTA of the grid: all rows of this rank are out -> advance the ring pointer, publish "my rows are in"
  __syncthreads();
  __shared__ bool s_last;
  if (tid == 0) {
    __threadfence_system();                                  // cumulative over the CTA's (remote) stores
    s_last = atomicAdd(&ctl->done[kXEnqueueDone], 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) {
    __threadfence();
    ctl->done[kXEnqueueDone] = 0;
    p.ptr_state[0] = (ptr0 + (long long)p.world * n) % p.K;
    *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch[kXEnqueueDone]) = epoch;
  }
  if (tid < p.world && tid != p.rank) st_relaxed_sys(flag_of(p.arenas[tid], kXEnqueueDone, p.rank), epoch);
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

int b200ssl_bank_enqueue_peer(const void* feats_u_w, const void* feats_x, const float* probs_orig, const int64_t* targets_x,
                              int64_t n_u, int64_t n_x, int32_t dim, int32_t classes, int32_t dtype, int64_t* ptr_state,
                              const b200ssl_bank_shards* shards, void* stream) {
  const char* fn = "b200ssl_bank_enqueue_peer";
  if (!shards || !ptr_state) return fail(B200SSL_E_NULL, "%s: NULL shard table / ptr_state", fn);
  if (n_u < 0 || n_x < 0 || n_u + n_x <= 0) return fail(B200SSL_E_SHAPE, "%s: n_u=%lld n_x=%lld", fn, (long long)n_u, (long long)n_x);
  if ((n_u > 0 && (!feats_u_w || !probs_orig)) || (n_x > 0 && (!feats_x || !targets_x))) return fail(B200SSL_E_NULL, "%s: NULL rows", fn);
  if (dtype != B200SSL_BF16 || dim != 64 || classes < 2 || classes > 31)
    return fail(B200SSL_E_DTYPE, "%s: the peer-memory bank is bf16, dim 64, classes <= 31", fn);
  if (shards->world < 2 || shards->world > 8 || shards->rank < 0 || shards->rank >= shards->world || !shards->arenas_dev ||
      shards->shard_rows <= 0 || shards->shard_rows % 8)
    return fail(B200SSL_E_ARG, "%s: bad shard table", fn);
  if ((reinterpret_cast<uintptr_t>(feats_u_w) | reinterpret_cast<uintptr_t>(feats_x)) & 15u) return fail(B200SSL_E_ALIGN, "%s: embeddings must be 16-byte aligned", fn);
  EnqueuePeerParams p{};
  p.fu = static_cast<const __nv_bfloat16*>(feats_u_w); p.fx = static_cast<const __nv_bfloat16*>(feats_x); p.po = probs_orig;
  p.tx = reinterpret_cast<const long long*>(targets_x); p.n_u = n_u; p.n_x = n_x;
  p.shard_rows = shards->shard_rows; p.K = shards->shard_rows * (shards->replicated ? 1 : shards->world);
  p.C = classes; p.rank = shards->rank; p.world = shards->world; p.replicated = shards->replicated ? 1 : 0;
  p.arenas = reinterpret_cast<uint8_t* const*>(shards->arenas_dev);
  p.qf_off = shards->feats_offset; p.qp_off = shards->probs_offset; p.qpt_off = shards->probs_t_offset;
  p.ptr_state = reinterpret_cast<long long*>(ptr_state);
  if ((n_u + n_x) * shards->world > p.K) return fail(B200SSL_E_SHAPE, "%s: world*(n_u + n_x) > bank rows", fn);
  const dim3 grid((unsigned)((n_u + n_x + kEnqRows - 1) / kEnqRows), (unsigned)(p.replicated ? p.world : 1), 1);
  bank_enqueue_peer_kernel<<<grid, kEnqThreads, 0, as_stream(stream)>>>(p);
  return check_launch(fn);
}

}  // extern "C"
