// SURVEY 8(f4), data side: the tail of the reference's view transforms on the device.
//
// code/dataset.py:24-109 builds every view on the host as  PIL ops -> ToTensor -> Normalize(mean, std)  and ships fp32
// NCHW tensors through the DataLoader and PCIe.  The pixel-exact, data-parallel part of that chain --
//     RandomHorizontalFlip            (dataset.py:35,46,66,73,84,95,105)
//     RandomCrop(size, padding, padding_mode='reflect')   (:36-38, :47-49)
//     ToTensor + Normalize            (:51-53, :107-109)
// -- runs here on uint8 HWC images (what PIL / the decoder produce): a batch crosses PCIe at 1 byte per sample instead of 4
// and lands in HBM as the normalised [N, 3, S, S] tensor the backbone reads.  The random decisions (flip, crop offset) stay
// with the caller's generator and arrive as per-image parameters, so a seeded run reproduces torchvision's views bit for
// bit: out = ((u8 / 255) - mean_c) / std_c with every operation rounded to fp32 like ToTensor().div(255), sub_(), div_().
// RandAugmentMC / ColorJitter (PIL resampling and lookup-table ops, randaugment.py:207-222) stay on the host: they sit
// between the crop and ToTensor and are not pixel-exactly reproducible outside PIL.
#include "common.cuh"

namespace b200ssl {
namespace {

struct ViewParams {
  const uint8_t* src;                 // [N, H, W, 3]
  void* dst;                          // [N, 3, S, S] fp32 or bf16
  const int32_t* flip;                // [N] 0/1 or NULL
  const int32_t* crop_xy;             // [N, 2] = (left, top) in the padded image or NULL (centre / identity)
  int N, H, W, S, pad;
  float mean[3], std[3];
};

__device__ __forceinline__ int reflect(int i, int n) {   // numpy.pad(mode='reflect'): no edge repeat
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// one thread = 4 horizontally adjacent output pixels x 3 channels: three 16-byte stores (fp32) per thread
template <typename T>
__global__ void __launch_bounds__(256) normalize_views_kernel(const ViewParams p) {
  const int S4 = p.S / 4;
  const long long total = (long long)p.N * p.S * S4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % S4);
    const int y = (int)((i / S4) % p.S);
    const int n = (int)(i / ((long long)S4 * p.S));
    const int left = p.crop_xy ? p.crop_xy[2 * n] : (p.W + 2 * p.pad - p.S) / 2;
    const int top = p.crop_xy ? p.crop_xy[2 * n + 1] : (p.H + 2 * p.pad - p.S) / 2;
    const bool fl = p.flip && p.flip[n];
    const int sy = reflect(top + y - p.pad, p.H);
    const uint8_t* row = p.src + ((size_t)n * p.H + sy) * p.W * 3;
    float v[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // the flip comes first in the reference's Compose: crop coordinates address the flipped image
      int sx = reflect(left + 4 * x4 + k - p.pad, p.W);
      if (fl) sx = p.W - 1 - sx;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[c][k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)row[3 * sx + c], 255.f), p.mean[c]), p.std[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T* out = static_cast<T*>(p.dst) + (((size_t)n * 3 + c) * p.S + y) * p.S + 4 * x4;
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(out) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[c][0], v[c][1]), b = __floats2bfloat162_rn(v[c][2], v[c][3]);
        *reinterpret_cast<uint2*>(out) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
    }
  }
}

}  // namespace
}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_normalize_views(const uint8_t* images_hwc, void* out_nchw, int32_t n, int32_t height, int32_t width,
                                       int32_t out_size, int32_t padding, const int32_t* flip, const int32_t* crop_xy,
                                       const float* mean3, const float* std3, int32_t out_dtype, void* stream) {
  const char* fn = "b200ssl_normalize_views";
  if (!images_hwc || !out_nchw || !mean3 || !std3) return fail(B200SSL_E_NULL, "%s: NULL pointer", fn);
  if (n <= 0 || height <= 0 || width <= 0 || out_size <= 0 || padding < 0) return fail(B200SSL_E_SHAPE, "%s: bad shape", fn);
  if (out_size % 4) return fail(B200SSL_E_SHAPE, "%s: out_size %d must be a multiple of 4 (16-byte stores)", fn, out_size);
  if (out_size > height + 2 * padding || out_size > width + 2 * padding || padding >= height || padding >= width)
    return fail(B200SSL_E_SHAPE, "%s: crop %d does not fit %dx%d padded by %d (reflect padding needs padding < size)", fn, out_size, height,
                width, padding);
  if (out_dtype != B200SSL_F32 && out_dtype != B200SSL_BF16) return fail(B200SSL_E_DTYPE, "%s: out_dtype %d", fn, out_dtype);
  if (reinterpret_cast<uintptr_t>(out_nchw) & 15u) return fail(B200SSL_E_ALIGN, "%s: output must be 16-byte aligned", fn);
  ViewParams p{images_hwc, out_nchw, flip, crop_xy, n, height, width, out_size, padding, {mean3[0], mean3[1], mean3[2]}, {std3[0], std3[1], std3[2]}};
  const long long total = (long long)n * out_size * (out_size / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  if (out_dtype == B200SSL_F32) normalize_views_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>(p);
  else normalize_views_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, as_stream(stream)>>>(p);
  return check_launch(fn);
}
