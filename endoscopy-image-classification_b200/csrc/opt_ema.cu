// optimizer step + EMA in ONE multi-tensor pass over the trainable parameters.
//
// Replaces `optimizer.step()` (code/fixmatch.py:123; SGD-nesterov / Adam / AdamW of code/optimizer.py:43-51, with
// the no-decay parameter group of :13-27) immediately followed by `ema_model.update(model)` (fixmatch.py:127,
// ema.py:51-59) for the parameters.  The eager sequence reads each weight twice more than needed (the EMA re-reads
// what the optimizer just wrote) and issues one multi-tensor launch per elementary op; here a parameter element is
// read once (p, g, state, e) and written once (p, state, e):
//     Adam/AdamW + EMA  36 B/element instead of 28 + 12;   SGD-momentum + EMA  28 B instead of 20 + 12.
// Buffers and frozen parameters keep going through b200ssl_ema_multi_tensor.
//
// Arithmetic follows torch.optim's single-tensor formulas (torch/optim/{sgd,adam,adamw}.py) operation by operation
// in fp32; the step-dependent scalars (lr, lr/(1-beta1^t), sqrt(1-beta2^t), 1-lr*wd) are computed on the host in double like
// torch does and travel as kernel parameters (one 64-byte row per parameter group, at most 8 groups).
#include "common.cuh"

namespace b200ssl {
namespace {

constexpr int kOptThreads = 256;

struct GroupRow {   // == b200ssl_opt_group
  float lr, beta1, beta2, eps, weight_decay, step_size, bias2_sqrt, momentum;
  int kind, nesterov, first_step, reserved;
  float one_minus_beta1, one_minus_beta2, decay_factor, pad;
};
static_assert(sizeof(GroupRow) == 64 && sizeof(b200ssl_opt_group) == 64, "group row is 64 bytes");
static_assert(sizeof(b200ssl_opt_block) == 64, "block row is 64 bytes");
constexpr int kMaxGroups = 8;
struct GroupTable { GroupRow g[kMaxGroups]; };

struct Elem { float p, g, s1, s2, e; };

__device__ __forceinline__ void update_elem(Elem& x, const GroupRow& h, float d, float o, int ema_repeat, bool has_ema) {
  float g = x.g;
  if (h.kind == B200SSL_OPT_SGD) {
    if (h.weight_decay != 0.f) g = __fmaf_rn(h.weight_decay, x.p, g);              // grad.add(param, alpha=wd)
    if (h.momentum != 0.f) {
      x.s1 = h.first_step ? g : __fadd_rn(__fmul_rn(x.s1, h.momentum), g);          // buf.mul_(m).add_(grad)
      g = h.nesterov ? __fmaf_rn(h.momentum, x.s1, g) : x.s1;                       // grad.add(buf, alpha=m)
    }
    x.p = __fmaf_rn(-h.lr, g, x.p);                                                 // param.add_(grad, alpha=-lr)
  } else {
    if (h.kind == B200SSL_OPT_ADAMW) x.p = __fmul_rn(x.p, h.decay_factor);          // param.mul_(1 - lr*wd)
    else if (h.weight_decay != 0.f) g = __fmaf_rn(h.weight_decay, x.p, g);
    x.s1 = __fmaf_rn(h.one_minus_beta1, __fsub_rn(g, x.s1), x.s1);                  // exp_avg.lerp_(grad, 1-beta1)
    x.s2 = __fmaf_rn(__fmul_rn(h.one_minus_beta2, g), g, __fmul_rn(x.s2, h.beta2)); // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(x.s2), h.bias2_sqrt), h.eps); // sqrt / sqrt(1-beta2^t) + eps
    x.p = __fmaf_rn(-h.step_size, __fdiv_rn(x.s1, denom), x.p);                     // addcdiv_(exp_avg, denom, -lr/(1-beta1^t))
  }
  if (has_ema)
    for (int r = 0; r < ema_repeat; ++r) x.e = __fadd_rn(__fmul_rn(d, x.e), __fmul_rn(o, x.p));   // ema.py:53-56
}

// dev_groups != NULL: the group rows are read from device memory (a CUDA graph replays the launch while the host refreshes
// the step-dependent scalars with a 64-byte-per-group copy ahead of every replay); NULL: they travel as kernel parameters.
__global__ void __launch_bounds__(kOptThreads, 4) opt_ema_kernel(const b200ssl_opt_block* __restrict__ blocks, int n_blocks,
                                                                  const __grid_constant__ GroupTable groups,
                                                                  const GroupRow* __restrict__ dev_groups, float d, float o) {
  for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const b200ssl_opt_block blk = blocks[b];
    const GroupRow h = dev_groups ? dev_groups[blk.group & (kMaxGroups - 1)] : groups.g[blk.group & (kMaxGroups - 1)];
    float* p = static_cast<float*>(blk.param);
    const float* g = static_cast<const float*>(blk.grad);
    float* s1 = static_cast<float*>(blk.state1);
    float* s2 = static_cast<float*>(blk.state2);
    float* e = static_cast<float*>(blk.ema);
    const bool use_s1 = s1 != nullptr, use_s2 = s2 != nullptr, has_ema = e != nullptr;
    const int count = blk.count;
    const uintptr_t align = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1) |
                            reinterpret_cast<uintptr_t>(s2) | reinterpret_cast<uintptr_t>(e);
    const int nvec = (align & 15u) ? 0 : count / 4;
    for (int v = threadIdx.x; v < nvec; v += kOptThreads) {
      const float4 pv = reinterpret_cast<const float4*>(p)[v];
      const float4 gv = reinterpret_cast<const float4*>(g)[v];
      float4 av = use_s1 ? reinterpret_cast<const float4*>(s1)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 bv = use_s2 ? reinterpret_cast<const float4*>(s2)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 ev = has_ema ? reinterpret_cast<const float4*>(e)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      Elem x[4] = {{pv.x, gv.x, av.x, bv.x, ev.x}, {pv.y, gv.y, av.y, bv.y, ev.y}, {pv.z, gv.z, av.z, bv.z, ev.z}, {pv.w, gv.w, av.w, bv.w, ev.w}};
#pragma unroll
      for (int k = 0; k < 4; ++k) update_elem(x[k], h, d, o, blk.ema_repeat, has_ema);
      reinterpret_cast<float4*>(p)[v] = make_float4(x[0].p, x[1].p, x[2].p, x[3].p);
      if (use_s1) reinterpret_cast<float4*>(s1)[v] = make_float4(x[0].s1, x[1].s1, x[2].s1, x[3].s1);
      if (use_s2) reinterpret_cast<float4*>(s2)[v] = make_float4(x[0].s2, x[1].s2, x[2].s2, x[3].s2);
      if (has_ema) reinterpret_cast<float4*>(e)[v] = make_float4(x[0].e, x[1].e, x[2].e, x[3].e);
    }
    for (int i = nvec * 4 + threadIdx.x; i < count; i += kOptThreads) {
      Elem x{p[i], g[i], use_s1 ? s1[i] : 0.f, use_s2 ? s2[i] : 0.f, has_ema ? e[i] : 0.f};
      update_elem(x, h, d, o, blk.ema_repeat, has_ema);
      p[i] = x.p;
      if (use_s1) s1[i] = x.s1;
      if (use_s2) s2[i] = x.s2;
      if (has_ema) e[i] = x.e;
    }
  }
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_opt_ema_multi_tensor(const b200ssl_opt_block* blocks, int32_t n_blocks, const b200ssl_opt_group* groups,
                                            int32_t n_groups, float decay, float one_minus_decay, void* stream) {
  const char* fn = "b200ssl_opt_ema_multi_tensor";
  if (!blocks || !groups) return fail(B200SSL_E_NULL, "%s: NULL table", fn);
  if (reinterpret_cast<uintptr_t>(blocks) & 15u) return fail(B200SSL_E_ALIGN, "%s: the block table must be 16-byte aligned", fn);
  if (n_blocks <= 0 || n_groups <= 0 || n_groups > kMaxGroups)
    return fail(B200SSL_E_SHAPE, "%s: n_blocks=%d n_groups=%d (1..%d parameter groups)", fn, n_blocks, n_groups, kMaxGroups);
  GroupTable table{};
  memcpy(table.g, groups, sizeof(GroupRow) * n_groups);
  for (int i = 0; i < n_groups; ++i)
    if (table.g[i].kind < B200SSL_OPT_SGD || table.g[i].kind > B200SSL_OPT_ADAMW) return fail(B200SSL_E_ARG, "%s: group %d: kind %d", fn, i, table.g[i].kind);
  const int max_grid = kNumSMs * 4;
  const int grid = n_blocks < max_grid ? n_blocks : max_grid;
  opt_ema_kernel<<<grid, kOptThreads, 0, as_stream(stream)>>>(blocks, n_blocks, table, nullptr, decay, one_minus_decay);
  return check_launch(fn);
}

extern "C" int b200ssl_opt_ema_multi_tensor_dev(const b200ssl_opt_block* blocks, int32_t n_blocks, const b200ssl_opt_group* groups_dev,
                                                int32_t n_groups, float decay, float one_minus_decay, void* stream) {
  const char* fn = "b200ssl_opt_ema_multi_tensor_dev";
  if (!blocks || !groups_dev) return fail(B200SSL_E_NULL, "%s: NULL table", fn);
  if ((reinterpret_cast<uintptr_t>(blocks) | reinterpret_cast<uintptr_t>(groups_dev)) & 15u)
    return fail(B200SSL_E_ALIGN, "%s: the tables must be 16-byte aligned", fn);
  if (n_blocks <= 0 || n_groups <= 0 || n_groups > kMaxGroups)
    return fail(B200SSL_E_SHAPE, "%s: n_blocks=%d n_groups=%d (1..%d parameter groups)", fn, n_blocks, n_groups, kMaxGroups);
  GroupTable table{};
  const int max_grid = kNumSMs * 4;
  const int grid = n_blocks < max_grid ? n_blocks : max_grid;
  opt_ema_kernel<<<grid, kOptThreads, 0, as_stream(stream)>>>(blocks, n_blocks, table, reinterpret_cast<const GroupRow*>(groups_dev), decay,
                                                             one_minus_decay);
  return check_launch(fn);
}
