// Per-sample photometric augmentations of the RESIDENT batch (the model input produced by
// sib_augment: normalised, already mirrored), replacing the DALI operators of the reference's
// train_pipeline between resize and crop_mirror_normalize (sota_imagenet/dali_dataloader.py:81-111):
//   gaussian_blur(window 11, sigma ~ U(0.5, 1.1))            -> sib_gaussian_blur
//   color_twist(contrast, brightness, hue, saturation)        \
//   hsv(saturation = 0) = grayscale                            > sib_pixel_ops (one pass, in place)
//   erase(re_count boxes, fill = DATA_MEAN)                   /
// Every operator of that chain is affine in the pixel value, and normalisation is affine too, so
// the same maps are applied in NORMALISED space with transformed coefficients (host side:
// data.BatchPixelAug); the [0, 255] clamp of the uint8 pipeline becomes a clamp at the normalised
// bounds.  Erase boxes are given in un-mirrored coordinates (DALI erases before the mirror) and are
// mirrored here for the samples whose crop box carries the flip flag.
// Layouts: 0 = bf16 NHWC with 4 channels (4th = 0), 1 = fp32 NCHW with 3 channels.
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

constexpr int kPixParams = 16;   // M[9] o[3] lo hi gray fill, followed by 4 floats per erase box

template <int LAYOUT>
__device__ __forceinline__ void load_px(const void* x, long n, int h, int w, int H, int W, float* v) {
  if (LAYOUT == 0) {
    const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x) +
                                                    (((long)n * H + h) * W + w) * 4);
    v[0] = __uint_as_float(q.x << 16);
    v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16);
  } else {
    const float* p = static_cast<const float*>(x) + ((long)n * 3 * H + h) * W + w;
    v[0] = p[0];
    v[1] = p[(long)H * W];
    v[2] = p[2l * H * W];
  }
}
template <int LAYOUT>
__device__ __forceinline__ void store_px(void* x, long n, int h, int w, int H, int W, const float* v) {
  if (LAYOUT == 0) {
    uint2 q;
    q.x = pack2(v[0], v[1]);
    q.y = pack2(v[2], 0.f);
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(x) + (((long)n * H + h) * W + w) * 4) = q;
  } else {
    float* p = static_cast<float*>(x) + ((long)n * 3 * H + h) * W + w;
    p[0] = v[0];
    p[(long)H * W] = v[1];
    p[2l * H * W] = v[2];
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
pixel_ops_kernel(void* __restrict__ x, const float* __restrict__ params, const int* __restrict__ crop_boxes,
                 int N, int H, int W, int nboxes) {
  const int stride = kPixParams + 4 * nboxes;
  const long total = (long)N * H * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int n = (int)(i / ((long)W * H));
    const float* p = params + (long)n * stride;
    float v[3], y[3];
    load_px<LAYOUT>(x, n, h, w, H, W, v);
    const float lo = p[12], hi = p[13];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float t = __fmaf_rn(p[3 * c + 2], v[2], __fmaf_rn(p[3 * c + 1], v[1], __fmaf_rn(p[3 * c], v[0], p[9 + c])));
      y[c] = fminf(fmaxf(t, lo), hi);
    }
    if (p[14] != 0.f) {     // grayscale: hsv(saturation = 0) keeps the luma (weights sum to 1)
      const float g = __fmaf_rn(0.114f, y[2], __fmaf_rn(0.587f, y[1], __fmul_rn(0.299f, y[0])));
      y[0] = y[1] = y[2] = g;
    }
    const bool flipped = crop_boxes != nullptr && crop_boxes[5 * n + 4] != 0;
    const int wu = flipped ? W - 1 - w : w;        // column in un-mirrored coordinates
    for (int b = 0; b < nboxes; ++b) {
      const float* e = p + kPixParams + 4 * b;
      if (h >= (int)e[0] && h < (int)e[2] && wu >= (int)e[1] && wu < (int)e[3]) {
        y[0] = y[1] = y[2] = p[15];
      }
    }
    store_px<LAYOUT>(x, n, h, w, H, W, y);
  }
}

// Separable 11-tap Gaussian, reflect-101 border.  One CTA per (sample, 16-row band): the band plus
// a 5-row halo is filtered horizontally into shared memory, then vertically to the output.
constexpr int kBlurR = 5;
constexpr int kBlurRows = 16;

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
gaussian_blur_kernel(const void* __restrict__ x, void* __restrict__ y, const float* __restrict__ sigma,
                     int N, int H, int W) {
  extern __shared__ float tile[];      // [kBlurRows + 2 * kBlurR][W][3]: horizontally filtered rows
  const int n = blockIdx.y;
  const int h0 = blockIdx.x * kBlurRows;
  const float sg = sigma[n];
  const int rows = min(kBlurRows, H - h0);
  if (sg <= 0.f) {                     // not selected: copy through
    for (int i = threadIdx.x; i < rows * W; i += blockDim.x) {
      float v[3];
      load_px<LAYOUT>(x, n, h0 + i / W, i % W, H, W, v);
      store_px<LAYOUT>(y, n, h0 + i / W, i % W, H, W, v);
    }
    return;
  }
  float wt[kBlurR + 1];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k <= kBlurR; ++k) {
    wt[k] = expf(-(float)(k * k) / (2.f * sg * sg));
    sum += k == 0 ? wt[k] : 2.f * wt[k];
  }
#pragma unroll
  for (int k = 0; k <= kBlurR; ++k) wt[k] /= sum;
  const int trows = rows + 2 * kBlurR;
  for (int i = threadIdx.x; i < trows * W; i += blockDim.x) {
    const int tr = i / W, w = i % W;
    const int hs = reflect101(h0 + tr - kBlurR, H);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = -kBlurR; k <= kBlurR; ++k) {
      float v[3];
      load_px<LAYOUT>(x, n, hs, reflect101(w + k, W), H, W, v);
      const float c = wt[k < 0 ? -k : k];
      acc[0] = __fmaf_rn(c, v[0], acc[0]);
      acc[1] = __fmaf_rn(c, v[1], acc[1]);
      acc[2] = __fmaf_rn(c, v[2], acc[2]);
    }
    tile[i * 3] = acc[0];
    tile[i * 3 + 1] = acc[1];
    tile[i * 3 + 2] = acc[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * W; i += blockDim.x) {
    const int r = i / W, w = i % W;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = -kBlurR; k <= kBlurR; ++k) {
      const float* t = tile + ((r + kBlurR + k) * W + w) * 3;
      const float c = wt[k < 0 ? -k : k];
      acc[0] = __fmaf_rn(c, t[0], acc[0]);
      acc[1] = __fmaf_rn(c, t[1], acc[1]);
      acc[2] = __fmaf_rn(c, t[2], acc[2]);
    }
    store_px<LAYOUT>(y, n, h0 + r, w, H, W, acc);
  }
}

}  // namespace sib

using namespace sib;

extern "C" int sib_pixel_ops(void* x, const float* params, const int* crop_boxes, int N, int H, int W,
                             int layout, int nboxes, void* stream) {
  SIB_CHECK(layout == 0 || layout == 1, "pixel_ops: layout must be 0 (bf16 NHWC4) or 1 (fp32 NCHW)");
  SIB_CHECK(nboxes >= 0 && nboxes <= 16, "pixel_ops: 0..16 erase boxes per sample (got %d)", nboxes);
  const long total = (long)N * H * W;
  long blocks = (total + 255) / 256;
  const long cap = (long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == 0) pixel_ops_kernel<0><<<(int)blocks, 256, 0, st>>>(x, params, crop_boxes, N, H, W, nboxes);
  else pixel_ops_kernel<1><<<(int)blocks, 256, 0, st>>>(x, params, crop_boxes, N, H, W, nboxes);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_gaussian_blur(const void* x, void* y, const float* sigma, int N, int H, int W,
                                 int layout, void* stream) {
  SIB_CHECK(layout == 0 || layout == 1, "gaussian_blur: layout must be 0 (bf16 NHWC4) or 1 (fp32 NCHW)");
  SIB_CHECK(x != y, "gaussian_blur: out of place only");
  const int smem = (kBlurRows + 2 * kBlurR) * W * 3 * (int)sizeof(float);
  SIB_CHECK(smem <= 200 * 1024, "gaussian_blur: image width %d too large", W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((H + kBlurRows - 1) / kBlurRows, N);
  if (layout == 0) {
    static bool configured = false;
    if (!configured) {
      SIB_CUDA(cudaFuncSetAttribute(gaussian_blur_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    gaussian_blur_kernel<0><<<grid, 256, smem, st>>>(x, y, sigma, N, H, W);
  } else {
    static bool configured = false;
    if (!configured) {
      SIB_CUDA(cudaFuncSetAttribute(gaussian_blur_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    gaussian_blur_kernel<1><<<grid, 256, smem, st>>>(x, y, sigma, N, H, W);
  }
  SIB_LAUNCH_CHECK();
  return 0;
}
