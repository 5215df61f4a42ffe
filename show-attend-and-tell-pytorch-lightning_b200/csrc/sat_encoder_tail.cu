// Encoder tail: the resize that turns the trunk's (1x1-conv) output into the annotation array.
// Reference: readme.md:118-121 appends nn.Upsample((s,s), mode="bilinear", align_corners=False) to the encoder of model.py:16-63;
// under bf16 autocast PyTorch runs it as  cast-to-fp32 -> upsample_bilinear2d<float> -> (later) cast-to-bf16  (three passes over the
// map, 0.5 ms of the BASELINE configs[1] step).  Here: ONE pass, NHWC in -> NHWC out = the [B,L,D] array the attention kernels
// stream, fp32 interpolation of the stored values with one rounding, and a deterministic gather-form backward (ATen's backward
// scatters with atomics).  HBM-bound: forward writes n*H2*W2*D elements, backward reads them once.
#include "sat_common.cuh"

namespace {

// source coordinate of output index o (PyTorch area_pixel_compute_source_index, align_corners=False, linear)
__device__ __forceinline__ void src_index(int o, float scale, int in, int& i0, int& i1, float& l1) {
  float s = scale * ((float)o + 0.5f) - 0.5f;
  s = s < 0.0f ? 0.0f : s;
  i0 = (int)s;
  i0 = i0 > in - 1 ? in - 1 : i0;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_fwd_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int h, int w, int H2, int W2, int D,
                                                         float sy, float sx) {
  constexpr int VN = Vec16<T>::N;
  const int NV = D / VN;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over [n, H2, W2, NV]
  if (idx >= (int64_t)n * H2 * W2 * NV) return;
  const int cv = (int)(idx % NV);
  int64_t r = idx / NV;
  const int x = (int)(r % W2); r /= W2;
  const int y = (int)(r % H2);
  const int64_t b = r / H2;
  int y0, y1, x0, x1;
  float ly, lx;
  src_index(y, sy, h, y0, y1, ly);
  src_index(x, sx, w, x0, x1, lx);
  const T* base = src + b * (int64_t)h * w * D + cv * VN;
  float a00[VN], a01[VN], a10[VN], a11[VN], o[VN];
  Vec16<T>::load(base + ((int64_t)y0 * w + x0) * D, a00);
  Vec16<T>::load(base + ((int64_t)y0 * w + x1) * D, a01);
  Vec16<T>::load(base + ((int64_t)y1 * w + x0) * D, a10);
  Vec16<T>::load(base + ((int64_t)y1 * w + x1) * D, a11);
  const float hy = 1.0f - ly, hx = 1.0f - lx;
#pragma unroll
  for (int k = 0; k < VN; ++k) o[k] = hy * (hx * a00[k] + lx * a01[k]) + ly * (hx * a10[k] + lx * a11[k]);      // ATen's grouping
  Vec16<T>::store(dst + idx * VN, o);
}

// weight with which output index o reads input index i along one axis
__device__ __forceinline__ float axis_weight(int o, int i, float scale, int in) {
  int i0, i1;
  float l1;
  src_index(o, scale, in, i0, i1, l1);
  return (i0 == i ? 1.0f - l1 : 0.0f) + (i1 == i ? l1 : 0.0f);
}

template <typename T>
__global__ void __launch_bounds__(256) resize_bwd_kernel(const T* __restrict__ d_dst, T* __restrict__ d_src, int n, int h, int w, int H2, int W2, int D,
                                                         float sy, float sx) {
  constexpr int VN = Vec16<T>::N;
  const int NV = D / VN;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over [n, h, w, NV]
  if (idx >= (int64_t)n * h * w * NV) return;
  const int cv = (int)(idx % NV);
  int64_t r = idx / NV;
  const int j = (int)(r % w); r /= w;
  const int i = (int)(r % h);
  const int64_t b = r / h;
  // output rows / columns that read input row i / column j: a contiguous range (the source coordinate is monotone)
  int ylo = H2, yhi = -1, xlo = W2, xhi = -1;
  for (int y = 0; y < H2; ++y)
    if (axis_weight(y, i, sy, h) != 0.0f) { ylo = y < ylo ? y : ylo; yhi = y; }
  for (int x = 0; x < W2; ++x)
    if (axis_weight(x, j, sx, w) != 0.0f) { xlo = x < xlo ? x : xlo; xhi = x; }
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.0f;
  const T* base = d_dst + b * (int64_t)H2 * W2 * D + cv * VN;
  for (int y = ylo; y <= yhi; ++y) {                 // fixed order: deterministic
    const float wy = axis_weight(y, i, sy, h);
    for (int x = xlo; x <= xhi; ++x) {
      const float wgt = wy * axis_weight(x, j, sx, w);
      float g[VN];
      Vec16<T>::load(base + ((int64_t)y * W2 + x) * D, g);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] += wgt * g[k];
    }
  }
  Vec16<T>::store(d_src + idx * VN, acc);
}

template <typename T>
int resize_impl(bool fwd, const void* a, void* o, int n, int h, int w, int H2, int W2, int D, cudaStream_t st) {
  const float sy = (float)h / (float)H2, sx = (float)w / (float)W2;
  const int NV = D / Vec16<T>::N;
  const int64_t total = (int64_t)n * (fwd ? (int64_t)H2 * W2 : (int64_t)h * w) * NV;
  if (total == 0) return 0;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (fwd) resize_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)a, (T*)o, n, h, w, H2, W2, D, sy, sx);
  else resize_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)a, (T*)o, n, h, w, H2, W2, D, sy, sx);
  SAT_COUNT_LAUNCH();
  SAT_LAUNCH_OK();
  return 0;
}

int resize_check(const void* a, void* o, int n, int h, int w, int H2, int W2, int D, int dtype) {
  SAT_REQUIRE(a && o, "sat_resize_nhwc: NULL pointer");
  SAT_REQUIRE(n >= 0 && h > 0 && w > 0 && H2 > 0 && W2 > 0 && D > 0, "sat_resize_nhwc: bad shape n=%d %dx%d -> %dx%d D=%d", n, h, w, H2, W2, D);
  SAT_REQUIRE(dtype == SAT_F32 || dtype == SAT_BF16, "sat_resize_nhwc: unknown dtype %d", dtype);
  SAT_REQUIRE(D % (dtype == SAT_F32 ? 4 : 8) == 0, "sat_resize_nhwc: D=%d must be a multiple of %d (16-byte channel vectors)", D, dtype == SAT_F32 ? 4 : 8);
  SAT_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(o)) & 15) == 0, "sat_resize_nhwc: pointers must be 16-byte aligned");
  return 0;
}

}  // namespace

extern "C" {

int sat_resize_nhwc_fwd(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t H2, int32_t W2, int32_t D, int32_t dtype, void* stream) {
  SAT_TRY(resize_check(src, dst, n, h, w, H2, W2, D, dtype));
  if (dtype == SAT_F32) return resize_impl<float>(true, src, dst, n, h, w, H2, W2, D, (cudaStream_t)stream);
  return resize_impl<bf16>(true, src, dst, n, h, w, H2, W2, D, (cudaStream_t)stream);
}

int sat_resize_nhwc_bwd(const void* d_dst, void* d_src, int32_t n, int32_t h, int32_t w, int32_t H2, int32_t W2, int32_t D, int32_t dtype, void* stream) {
  SAT_TRY(resize_check(d_dst, d_src, n, h, w, H2, W2, D, dtype));
  if (dtype == SAT_F32) return resize_impl<float>(false, d_dst, d_src, n, h, w, H2, W2, D, (cudaStream_t)stream);
  return resize_impl<bf16>(false, d_dst, d_src, n, h, w, H2, W2, D, (cudaStream_t)stream);
}

}  // extern "C"
