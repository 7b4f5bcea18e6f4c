// ovmr_b200 — input side of the hot path: the reference's `_transform`
// (clip/clip.py:73-80: Resize(n_px, BICUBIC) -> CenterCrop(n_px) -> ToTensor -> Normalize; Dassl's test transform,
// dassl/data/transforms/transforms.py:495-526, is the same chain) on the GPU, from decoded uint8 RGB pixels.
//
// Resize on a PIL image is Pillow's ImagingResample (a third-party dependency of the reference, Pillow >= 9): two
// separable passes — horizontal, then vertical — over uint8 data with per-output-pixel windows of normalised filter
// weights converted to 22-bit fixed point, each pass rounding to uint8:
//     ss = 1 << 21;  ss += px[xmin + k] * kk[k] ...;  out = clip8(ss >> 22)
// The windows / weights are computed on the host in double precision with Pillow's formulas (restated below from
// its published algorithm); the two passes run as CUDA kernels and are bit-exact against Pillow
// (tests/golden/preprocess.npz holds outputs of the reference's own _transform; tests/test_preprocess*.py).
// The centre crop is folded in: only pixels inside the crop window are produced.  Output is uint8 CHW, which
// ovmr_vit_forward_u8 consumes with ToTensor + Normalize fused into the patch load.
#include "../../include/ovmr_b200.h"

#include <math.h>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// Pillow's bicubic kernel (a = -0.5), support 2
inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
inline double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src HWC uint8 [H, W, 3] -> tmp HWC uint8 [rows, crop_w, 3] for source rows [y0, y0 + rows) and
// output columns [crop_left, crop_left + crop_w)
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, int W, int y0, int rows, const int* __restrict__ bounds,
                  const int* __restrict__ kk, int ksize, int crop_left, int crop_w, uint8_t* __restrict__ tmp) {
  const long long total = static_cast<long long>(rows) * crop_w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xo = static_cast<int>(i % crop_w), r = static_cast<int>(i / crop_w);
    const int xx = crop_left + xo;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int* k = kk + static_cast<long long>(xx) * ksize;
    const uint8_t* p = src + (static_cast<long long>(y0 + r) * W + xmin) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
      const int w = k[x];
      s0 += p[3 * x] * w;
      s1 += p[3 * x + 1] * w;
      s2 += p[3 * x + 2] * w;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = clip8(s0);
    o[1] = clip8(s1);
    o[2] = clip8(s2);
  }
}

// vertical pass: tmp HWC [rows, crop_w, 3] (source rows from y0) -> dst CHW uint8 [3, crop_h, crop_w]
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, int y0, const int* __restrict__ bounds, const int* __restrict__ kk,
                  int ksize, int crop_top, int crop_h, int crop_w, uint8_t* __restrict__ dst) {
  const long long total = static_cast<long long>(crop_h) * crop_w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xo = static_cast<int>(i % crop_w), yo = static_cast<int>(i / crop_w);
    const int yy = crop_top + yo;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int* k = kk + static_cast<long long>(yy) * ksize;
    const uint8_t* p = tmp + (static_cast<long long>(ymin - y0) * crop_w + xo) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
      const int w = k[y];
      const uint8_t* q = p + static_cast<long long>(y) * crop_w * 3;
      s0 += q[0] * w;
      s1 += q[1] * w;
      s2 += q[2] * w;
    }
    const long long plane = static_cast<long long>(crop_h) * crop_w;
    dst[i] = clip8(s0);
    dst[plane + i] = clip8(s1);
    dst[2 * plane + i] = clip8(s2);
  }
}

inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = static_cast<long long>(ovmr::num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

// Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c), host side, double precision.
int ovmr_resample_coeffs(int in_size, int out_size, int filter, int* bounds, int* kk, int kk_capacity) {
  if (in_size <= 0 || out_size <= 0 || (filter != 2 && filter != 3)) {
    ovmr::set_last_error("resample_coeffs: in_size=%d out_size=%d filter=%d (2 = bilinear, 3 = bicubic)", in_size, out_size, filter);
    return -1;
  }
  const double filter_support = filter == 3 ? 2.0 : 1.0;
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = filter_support * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  if (bounds == nullptr || kk == nullptr) return ksize;   // size query
  if (static_cast<long long>(out_size) * ksize > kk_capacity) {
    ovmr::set_last_error("resample_coeffs: kk capacity %d < %lld", kk_capacity, static_cast<long long>(out_size) * ksize);
    return -1;
  }
  double* pre = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double arg = (x + xmin - center + 0.5) * ss;
      const double w = filter == 3 ? bicubic_filter(arg) : bilinear_filter(arg);
      pre[x] = w;
      ww += w;
    }
    int* k = kk + static_cast<long long>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      double v = 0.0;
      if (x < xmax) v = ww != 0.0 ? pre[x] / ww : pre[x];
      k[x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << PRECISION_BITS)) : static_cast<int>(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  delete[] pre;
  return ksize;
}

int ovmr_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int out_h, int out_w, const int* xbounds, const int* xk,
                        int xksize, const int* ybounds, const int* yk, int yksize, const int* ybounds_host, int crop_top,
                        int crop_left, int crop_h, int crop_w, uint8_t* tmp, size_t tmp_bytes, uint8_t* dst_chw,
                        void* stream) {
  OVMR_REQUIRE(src_hwc && xbounds && xk && ybounds && yk && ybounds_host && tmp && dst_chw, "resize_crop: null argument");
  OVMR_REQUIRE(H > 0 && W > 0 && out_h > 0 && out_w > 0 && crop_h > 0 && crop_w > 0 && crop_top >= 0 && crop_left >= 0 &&
                   crop_top + crop_h <= out_h && crop_left + crop_w <= out_w,
               "resize_crop: bad geometry (%dx%d -> %dx%d, crop %dx%d at %d,%d)", H, W, out_h, out_w, crop_h, crop_w,
               crop_top, crop_left);
  // source rows the vertical pass of the crop window touches
  const int y0 = ybounds_host[2 * crop_top];
  const int y1 = ybounds_host[2 * (crop_top + crop_h - 1)] + ybounds_host[2 * (crop_top + crop_h - 1) + 1];
  OVMR_REQUIRE(y0 >= 0 && y1 <= H && y1 > y0, "resize_crop: inconsistent vertical bounds");
  const int rows = y1 - y0;
  OVMR_REQUIRE(tmp_bytes >= static_cast<size_t>(rows) * crop_w * 3, "resize_crop: tmp too small (%zu < %zu)", tmp_bytes,
               static_cast<size_t>(rows) * crop_w * 3);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  resample_h_kernel<<<grid_for(static_cast<long long>(rows) * crop_w, 256), 256, 0, st>>>(src_hwc, W, y0, rows, xbounds, xk,
                                                                                         xksize, crop_left, crop_w, tmp);
  OVMR_CHECK_CUDA(cudaGetLastError());
  resample_v_kernel<<<grid_for(static_cast<long long>(crop_h) * crop_w, 256), 256, 0, st>>>(tmp, y0, ybounds, yk, yksize,
                                                                                           crop_top, crop_h, crop_w, dst_chw);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(2);
  return 0;
}

}  // extern "C"
