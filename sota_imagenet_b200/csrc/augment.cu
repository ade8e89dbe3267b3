// Synthetic-input GPU augmentation, standing in for the DALI train pipeline
// (reference dali_dataloader.py:65-74 random-resized-crop, :74 triangular resize,
// :113-122 crop_mirror_normalize with mean 127.5 / std 51, :123 one_hot), plus the
// stem input packer that turns the 3-channel image into the 64-channel row-pair layout
// the tensor-core stem convolution consumes.
//
// Crop boxes come from a counter-based Philox4x32-10 stream keyed by (seed, sample index) so
// the CPU oracle (oracle/augment_ref.c) reproduces them bit-exactly.
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

// ---- Philox4x32-10 (Salmon et al. 2011) ----
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// box = {x0, y0, w, h, flip}
__host__ __device__ inline void rrc_box(int H, int W, double min_area, double max_area,
                                        uint64_t seed, uint64_t sample, int box[5]) {
  const double kLogLo = -0.2876820724517809;   // ln 0.75
  const double kLogHi = 0.22314355131420976;   // ln 1.25
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t s0 = (uint32_t)sample, s1 = (uint32_t)(sample >> 32);
  uint32_t r[4];
  int bw = 0, bh = 0, bx = 0, by = 0;
  bool found = false;
  for (uint32_t attempt = 0; attempt < 100 && !found; ++attempt) {
    philox4x32_10(s0, s1, attempt, 0u, k0, k1, r);
    const double ua = (double)r[0] * (1.0 / 4294967296.0);
    const double ur = (double)r[1] * (1.0 / 4294967296.0);
    const double area = (min_area + (max_area - min_area) * ua) * (double)H * (double)W;
    const double ratio = exp(kLogLo + (kLogHi - kLogLo) * ur);
    const int w = (int)floor(sqrt(area * ratio) + 0.5);
    const int h = (int)floor(sqrt(area / ratio) + 0.5);
    if (w > 0 && h > 0 && w <= W && h <= H) {
      bw = w; bh = h;
      bx = (int)(r[2] % (uint32_t)(W - w + 1));
      by = (int)(r[3] % (uint32_t)(H - h + 1));
      found = true;
    }
  }
  if (!found) {
    // largest centred crop whose aspect lies inside [0.75, 1.25]
    const double in_ratio = (double)W / (double)H;
    if (in_ratio < 0.75) { bw = W; bh = (int)floor((double)W / 0.75 + 0.5); }
    else if (in_ratio > 1.25) { bh = H; bw = (int)floor((double)H * 1.25 + 0.5); }
    else { bw = W; bh = H; }
    if (bh > H) bh = H;
    if (bw > W) bw = W;
    bx = (W - bw) / 2;
    by = (H - bh) / 2;
  }
  philox4x32_10(s0, s1, 100u, 1u, k0, k1, r);
  box[0] = bx; box[1] = by; box[2] = bw; box[3] = bh; box[4] = (int)(r[0] & 1u);
}

__global__ void rrc_boxes_kernel(int* __restrict__ boxes, int B, int H, int W, double min_area,
                                 double max_area, uint64_t seed, uint64_t first_sample,
                                 int do_flip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int b[5];
  rrc_box(H, W, min_area, max_area, seed, first_sample + (uint64_t)i, b);
  if (!do_flip) b[4] = 0;
  for (int j = 0; j < 5; ++j) boxes[i * 5 + j] = b[j];
}

// Triangular (anti-aliased bilinear) resample of the crop to S x S, mirror, normalise.
// One thread per output pixel (3 channels).  out_mode 0: NHWC bf16 with 4 channels
// (4th = 0); out_mode 1: NCHW fp32 (the reference layout).
__global__ void __launch_bounds__(256)
augment_kernel(const uint8_t* __restrict__ src, const int* __restrict__ boxes, void* __restrict__ out,
               int B, int SH, int SW, int S, float mean, float inv_std, int out_mode) {
  const long total = (long)B * S * S;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S);
    const int oy = (int)((i / S) % S);
    const int n = (int)(i / ((long)S * S));
    const int* bx = boxes + n * 5;
    const int x0 = bx[0], y0 = bx[1], cw = bx[2], ch = bx[3], flip = bx[4];
    const int sx_out = flip ? (S - 1 - ox) : ox;
    const float scx = (float)cw / (float)S, scy = (float)ch / (float)S;
    const float supx = fmaxf(scx, 1.f), supy = fmaxf(scy, 1.f);
    const float cx = ((float)sx_out + 0.5f) * scx, cy = ((float)oy + 0.5f) * scy;
    const int xlo = (int)floorf(cx - supx), xhi = (int)ceilf(cx + supx);
    const int ylo = (int)floorf(cy - supy), yhi = (int)ceilf(cy + supy);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, wsum = 0.f;
    const uint8_t* img = src + (long)n * SH * SW * 3;
    for (int yy = ylo; yy < yhi; ++yy) {
      const float wy = fmaxf(0.f, 1.f - fabsf(((float)yy + 0.5f - cy) / supy));
      if (wy <= 0.f) continue;
      const int sy = min(max(yy, 0), ch - 1) + y0;
      for (int xx = xlo; xx < xhi; ++xx) {
        const float wx = fmaxf(0.f, 1.f - fabsf(((float)xx + 0.5f - cx) / supx));
        if (wx <= 0.f) continue;
        const int sx = min(max(xx, 0), cw - 1) + x0;
        const uint8_t* px = img + ((long)sy * SW + sx) * 3;
        const float w = wx * wy;
        acc0 = fmaf(w, (float)px[0], acc0);
        acc1 = fmaf(w, (float)px[1], acc1);
        acc2 = fmaf(w, (float)px[2], acc2);
        wsum += w;
      }
    }
    const float inv = 1.f / wsum;
    const float v0 = (acc0 * inv - mean) * inv_std;
    const float v1 = (acc1 * inv - mean) * inv_std;
    const float v2 = (acc2 * inv - mean) * inv_std;
    if (out_mode == 0) {
      uint2 o;
      o.x = pack2(v0, v1);
      o.y = pack2(v2, 0.f);
      reinterpret_cast<uint2*>(out)[i] = o;
    } else {
      float* o = static_cast<float*>(out);
      const long plane = (long)S * S;
      const long base = (long)n * 3 * plane + (long)oy * S + ox;
      o[base] = v0;
      o[base + plane] = v1;
      o[base + 2 * plane] = v2;
    }
  }
}

// Validation transform (reference dali_dataloader.py:146-160): resize so that the SHORTER side
// becomes `RS` (INTERP_TRIANGULAR, the other side scaled by the same ratio and rounded), centre
// crop S x S, normalise.  Resize and crop are fused: output pixel (oy, ox) is pixel
// (oy + oy0, ox + ox0) of the virtual RH x RW resized image; filter taps clamp at the IMAGE
// border (not at the crop window, unlike the train kernel whose crop happens before the resize).
__global__ void __launch_bounds__(256)
val_transform_kernel(const uint8_t* __restrict__ src, void* __restrict__ out, int B, int SH, int SW,
                     int S, int RH, int RW, int oy0, int ox0, float mean, float inv_std,
                     int out_mode) {
  const long total = (long)B * S * S;
  const float scx = (float)SW / (float)RW, scy = (float)SH / (float)RH;
  const float supx = fmaxf(scx, 1.f), supy = fmaxf(scy, 1.f);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S);
    const int oy = (int)((i / S) % S);
    const int n = (int)(i / ((long)S * S));
    const float cx = ((float)(ox + ox0) + 0.5f) * scx, cy = ((float)(oy + oy0) + 0.5f) * scy;
    const int xlo = (int)floorf(cx - supx), xhi = (int)ceilf(cx + supx);
    const int ylo = (int)floorf(cy - supy), yhi = (int)ceilf(cy + supy);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, wsum = 0.f;
    const uint8_t* img = src + (long)n * SH * SW * 3;
    for (int yy = ylo; yy < yhi; ++yy) {
      const float wy = fmaxf(0.f, 1.f - fabsf(((float)yy + 0.5f - cy) / supy));
      if (wy <= 0.f) continue;
      const int sy = min(max(yy, 0), SH - 1);
      for (int xx = xlo; xx < xhi; ++xx) {
        const float wx = fmaxf(0.f, 1.f - fabsf(((float)xx + 0.5f - cx) / supx));
        if (wx <= 0.f) continue;
        const int sx = min(max(xx, 0), SW - 1);
        const uint8_t* px = img + ((long)sy * SW + sx) * 3;
        const float w = wx * wy;
        acc0 = fmaf(w, (float)px[0], acc0);
        acc1 = fmaf(w, (float)px[1], acc1);
        acc2 = fmaf(w, (float)px[2], acc2);
        wsum += w;
      }
    }
    const float inv = 1.f / wsum;
    const float v0 = (acc0 * inv - mean) * inv_std;
    const float v1 = (acc1 * inv - mean) * inv_std;
    const float v2 = (acc2 * inv - mean) * inv_std;
    if (out_mode == 0) {
      uint2 o;
      o.x = pack2(v0, v1);
      o.y = pack2(v2, 0.f);
      reinterpret_cast<uint2*>(out)[i] = o;
    } else {
      float* o = static_cast<float*>(out);
      const long plane = (long)S * S;
      const long base = (long)n * 3 * plane + (long)oy * S + ox;
      o[base] = v0;
      o[base + plane] = v1;
      o[base + 2 * plane] = v2;
    }
  }
}

__global__ void one_hot_kernel(const long* __restrict__ labels, float* __restrict__ out, int B,
                               int C) {
  const long total = (long)B * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    out[i] = (labels[i / C] == c) ? 1.f : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Batch-level mixing on the resident batch: pt_clb.Mixup / pt_clb.Cutmix as combined by the
// reference's CutmixMixup (sota_imagenet/callbacks.py:232-247; the two parents live in the absent
// pytorch_tools package: restated, unpinned).
//   mode 0 (mixup):  out = lam * x + (1 - lam) * prev[perm[n]]          (fp32 math, one rounding)
//   mode 1 (cutmix): out = (h1 <= h < h2 && w1 <= w < w2) ? prev[perm[n]] : x
// LAYOUT 0: NHWC bf16, 8-element vectors along the (w, c) row; LAYOUT 1: NCHW fp32, float4 along w.
// `prev` may have another spatial extent (progressive resizing): [N][PH][PW][C] / [N][C][PH][PW].
template <int LAYOUT>
__global__ void __launch_bounds__(256)
mix_batch_kernel(const void* __restrict__ xv, const void* __restrict__ prevv,
                 const int* __restrict__ perm, void* __restrict__ outv, int N, int H, int W, int C,
                 int PH, int PW, int mode, float lam, float one_minus_lam, int h1, int w1, int h2,
                 int w2) {
  constexpr int V = LAYOUT == 0 ? 8 : 4;
  const int rowlen = LAYOUT == 0 ? W * C : W;            // elements of one vectorised row
  const int rowvecs = rowlen / V;
  const long rows = LAYOUT == 0 ? (long)N * H : (long)N * C * H;
  const long total = rows * rowvecs;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long)gridDim.x * blockDim.x) {
    const long row = v / rowvecs;
    const int col = (int)(v - row * rowvecs) * V;
    const int h = (int)(row % H);
    const long nc = row / H;                             // n (NHWC) or n*C + c (NCHW)
    const int n = LAYOUT == 0 ? (int)nc : (int)(nc / C);
    const int pn = perm[n];
    const long prow = LAYOUT == 0 ? ((long)pn * PH + h) * ((long)PW * C)
                                  : (((long)pn * C + (nc - (long)n * C)) * PH + h) * (long)PW;
    bool in[V];
    bool any = mode == 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int w = LAYOUT == 0 ? (col + j) / C : col + j;
      in[j] = mode == 1 && h >= h1 && h < h2 && w >= w1 && w < w2;
      any |= in[j];
    }
    if (LAYOUT == 0) {
      const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(xv) + row * rowlen + col;
      uint4 a = *reinterpret_cast<const uint4*>(x);
      if (any) {
        const uint4 b = *reinterpret_cast<const uint4*>(
            static_cast<const __nv_bfloat16*>(prevv) + prow + col);
        __nv_bfloat16* ae = reinterpret_cast<__nv_bfloat16*>(&a);
        const __nv_bfloat16* be = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (mode == 0)
            ae[j] = __float2bfloat16_rn(__fadd_rn(__fmul_rn(lam, __bfloat162float(ae[j])),
                                                  __fmul_rn(one_minus_lam, __bfloat162float(be[j]))));
          else if (in[j])
            ae[j] = be[j];
        }
      }
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(outv) + row * rowlen + col) = a;
    } else {
      const float* x = static_cast<const float*>(xv) + row * rowlen + col;
      float4 a = *reinterpret_cast<const float4*>(x);
      if (any) {
        const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(prevv) + prow + col);
        float* ae = reinterpret_cast<float*>(&a);
        const float* be = reinterpret_cast<const float*>(&b);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (mode == 0)
            ae[j] = __fadd_rn(__fmul_rn(lam, ae[j]), __fmul_rn(one_minus_lam, be[j]));
          else if (in[j])
            ae[j] = be[j];
        }
      }
      *reinterpret_cast<float4*>(static_cast<float*>(outv) + row * rowlen + col) = a;
    }
  }
}

// out[n][c] = w_self * t[n][c] + w_prev * prev_t[perm[n]][c]   (soft targets of mixup / cutmix)
__global__ void __launch_bounds__(256)
mix_targets_kernel(const float* __restrict__ t, const float* __restrict__ prev_t,
                   const int* __restrict__ perm, float* __restrict__ out, int N, int C,
                   float w_self, float w_prev) {
  const long total = (long)N * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / C);
    const int c = (int)(i - (long)n * C);
    out[i] = __fadd_rn(__fmul_rn(w_self, t[i]), __fmul_rn(w_prev, prev_t[(long)perm[n] * C + c]));
  }
}

// Stem packer.  For a KHxKW stride-2 convolution with padding (ph, pw) over a 3-channel
// image, output row p reads image rows 2p + r - ph.  Writing that row as 2*(p + a) + b with
// b in {0,1} turns the filter into NA row-pair taps over a packed tensor
//   Xq[n][j][q][(b, s, c)] = x[n][2j + b][2q + s - pw][c]      (zero outside the image)
// with KW*2*3 real channels padded to 64, i.e. an (NA x 1), stride-1 convolution with 64 input
// channels that the swizzled implicit-GEMM kernel handles directly.
// src_mode 0: NHWC bf16 4-channel, 1: NCHW fp32.
template <int KW>
__global__ void __launch_bounds__(256)
stem_pack_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ xq, int N, int H,
                 int W, int pw, int src_mode) {
  // one thread per packed pixel (n, j, q): gathers 2 rows x KW source pixels (each read once as
  // an 8-byte NHWC4 load, or 3 fp32 loads for NCHW input) and writes 64 channels = 8 x 16 bytes
  const int H2 = H >> 1, W2 = W >> 1;
  const long total = (long)N * H2 * W2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int q = (int)(i % W2);
    const int j = (int)((i / W2) % H2);
    const int n = (int)(i / ((long)W2 * H2));
    __nv_bfloat16 vals[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) vals[e] = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int y = 2 * j + b;
#pragma unroll
      for (int s = 0; s < KW; ++s) {
        const int x = 2 * q + s - pw;
        if (x < 0 || x >= W) continue;
        __nv_bfloat16* dst = vals + (b * KW + s) * 3;
        if (src_mode == 0) {
          const uint2 px = __ldg(reinterpret_cast<const uint2*>(
              static_cast<const __nv_bfloat16*>(src) + (((long)n * H + y) * W + x) * 4));
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&px);
          dst[0] = h[0]; dst[1] = h[1]; dst[2] = h[2];
        } else {
          const float* f = static_cast<const float*>(src);
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dst[c] = __float2bfloat16_rn(__ldg(f + (((long)n * 3 + c) * H + y) * W + x));
        }
      }
    }
    uint4* out = reinterpret_cast<uint4*>(xq + i * 64);
    const uint4* v4 = reinterpret_cast<const uint4*>(vals);
#pragma unroll
    for (int k = 0; k < 8; ++k) stg_stream(out + k, v4[k]);
  }
}

// Filter packing for the stem: Wq[k][a][(b,s,c)] = W[k][c][r = 2a + b - off][s], off chosen so
// that a = 0 covers the top padding rows; fp32 OIHW (contiguous) source.
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wq,
                                        int K, int KH, int KW, int NA, int off) {
  const int total = K * NA * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int chn = i & 63;
    const int a = (i >> 6) % NA;
    const int k = i / (64 * NA);
    float val = 0.f;
    if (chn < 2 * KW * 3) {
      const int b = chn / (KW * 3);
      const int rem = chn - b * KW * 3;
      const int s = rem / 3, c = rem - s * 3;
      const int r = 2 * a + b - off;
      if (r >= 0 && r < KH) val = w[(((long)k * 3 + c) * KH + r) * KW + s];
    }
    wq[i] = __float2bfloat16_rn(val);
  }
}

// inverse map for the gradient: dW[k][c][r][s] = dWq[k][a][(b,s,c)]
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dwq, float* __restrict__ dw,
                                         int K, int KH, int KW, int NA, int off, int accumulate) {
  const int total = K * 3 * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % KW;
    const int r = (i / KW) % KH;
    const int c = (i / (KW * KH)) % 3;
    const int k = i / (KW * KH * 3);
    const int rr = r + off;
    const int a = rr >> 1, b = rr & 1;
    const float g = dwq[((long)k * NA + a) * 64 + b * KW * 3 + s * 3 + c];
    dw[i] = accumulate ? dw[i] + g : g;
  }
}

static inline int ew_grid(long n, int threads) {
  long b = (n + threads - 1) / threads;
  long cap = (long)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace sib

using namespace sib;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int sib_rrc_boxes(int* boxes_dev, int B, int H, int W, double min_area, double max_area,
                             unsigned long long seed, unsigned long long first_sample, int do_flip,
                             void* stream) {
  rrc_boxes_kernel<<<(B + 127) / 128, 128, 0, ST(stream)>>>(boxes_dev, B, H, W, min_area, max_area,
                                                           seed, first_sample, do_flip);
  SIB_LAUNCH_CHECK();
  return 0;
}

// host twin of the device box generator (same code path compiled for the CPU)
extern "C" void sib_rrc_box_host(int H, int W, double min_area, double max_area,
                                 unsigned long long seed, unsigned long long sample, int* box5) {
  rrc_box(H, W, min_area, max_area, seed, sample, box5);
}

extern "C" int sib_augment(const void* src_u8, const int* boxes_dev, void* out, int B, int SH,
                           int SW, int S, float mean, float std, int out_mode, void* stream) {
  SIB_CHECK(out_mode == 0 || out_mode == 1, "augment: out_mode must be 0 (NHWC4 bf16) or 1 (NCHW f32)");
  const long total = (long)B * S * S;
  augment_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const uint8_t*>(src_u8), boxes_dev, out, B, SH, SW, S, mean, 1.f / std, out_mode);
  SIB_LAUNCH_CHECK();
  return 0;
}

// resized extent of the resize-shorter step and the centre-crop origin (host twin exported for
// the tests: sib_val_geometry_host)
__host__ __device__ static inline void val_geometry(int SH, int SW, int S, int RS, int* g4) {
  int RH, RW;
  if (SH <= SW) { RH = RS; RW = (int)floor((double)SW * RS / SH + 0.5); }
  else          { RW = RS; RH = (int)floor((double)SH * RS / SW + 0.5); }
  g4[0] = RH; g4[1] = RW;
  g4[2] = (int)floor(0.5 * (RH - S) + 0.5);   // oy0
  g4[3] = (int)floor(0.5 * (RW - S) + 0.5);   // ox0
}

extern "C" void sib_val_geometry_host(int SH, int SW, int S, int resize_shorter, int* g4_host) {
  val_geometry(SH, SW, S, resize_shorter, g4_host);
}

extern "C" int sib_val_transform(const void* src_u8, void* out, int B, int SH, int SW, int S,
                                 int resize_shorter, float mean, float std, int out_mode,
                                 void* stream) {
  SIB_CHECK(out_mode == 0 || out_mode == 1, "val_transform: out_mode must be 0 (NHWC4 bf16) or 1 (NCHW f32)");
  SIB_CHECK(resize_shorter >= S, "val_transform: resize_shorter %d smaller than the crop %d", resize_shorter, S);
  int g[4];
  val_geometry(SH, SW, S, resize_shorter, g);
  const long total = (long)B * S * S;
  val_transform_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const uint8_t*>(src_u8), out, B, SH, SW, S, g[0], g[1], g[2], g[3], mean, 1.f / std,
      out_mode);
  SIB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Ragged batches (real images of different sizes, records.pack_batch): one packed uint8 buffer,
// per-image byte offsets and {H, W}.  Same arithmetic as the uniform kernels above, with the
// image base pointer and extent looked up per sample.  mode 0: train (crop box, flip);
// mode 1: validation (resize shorter side to RS, centre crop).
namespace sib {

__global__ void rrc_boxes_ragged_kernel(int* __restrict__ boxes, const int* __restrict__ dims, int B,
                                        double min_area, double max_area, uint64_t seed,
                                        uint64_t first_sample, int do_flip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int b[5];
  rrc_box(dims[2 * i], dims[2 * i + 1], min_area, max_area, seed, first_sample + (uint64_t)i, b);
  if (!do_flip) b[4] = 0;
  for (int j = 0; j < 5; ++j) boxes[i * 5 + j] = b[j];
}

__global__ void __launch_bounds__(256)
resample_ragged_kernel(const uint8_t* __restrict__ packed, const long* __restrict__ offsets,
                       const int* __restrict__ dims, const int* __restrict__ boxes,
                       void* __restrict__ out, int B, int S, int RS, int mode, float mean,
                       float inv_std, int out_mode) {
  const long total = (long)B * S * S;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S);
    const int oy = (int)((i / S) % S);
    const int n = (int)(i / ((long)S * S));
    const int SH = dims[2 * n], SW = dims[2 * n + 1];
    const uint8_t* img = packed + offsets[n];
    // source window [x0, x0 + cw) x [y0, y0 + ch) that taps clamp to, sample scale and centre
    int x0, y0, cw, ch;
    float scx, scy, cx, cy;
    if (mode == 0) {
      const int* bx = boxes + n * 5;
      x0 = bx[0]; y0 = bx[1]; cw = bx[2]; ch = bx[3];
      const int sx_out = bx[4] ? (S - 1 - ox) : ox;
      scx = (float)cw / (float)S; scy = (float)ch / (float)S;
      cx = ((float)sx_out + 0.5f) * scx; cy = ((float)oy + 0.5f) * scy;
    } else {
      int g[4];
      val_geometry(SH, SW, S, RS, g);
      x0 = 0; y0 = 0; cw = SW; ch = SH;
      scx = (float)SW / (float)g[1]; scy = (float)SH / (float)g[0];
      cx = ((float)(ox + g[3]) + 0.5f) * scx; cy = ((float)(oy + g[2]) + 0.5f) * scy;
    }
    const float supx = fmaxf(scx, 1.f), supy = fmaxf(scy, 1.f);
    const int xlo = (int)floorf(cx - supx), xhi = (int)ceilf(cx + supx);
    const int ylo = (int)floorf(cy - supy), yhi = (int)ceilf(cy + supy);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, wsum = 0.f;
    for (int yy = ylo; yy < yhi; ++yy) {
      const float wy = fmaxf(0.f, 1.f - fabsf(((float)yy + 0.5f - cy) / supy));
      if (wy <= 0.f) continue;
      const int sy = min(max(yy, 0), ch - 1) + y0;
      for (int xx = xlo; xx < xhi; ++xx) {
        const float wx = fmaxf(0.f, 1.f - fabsf(((float)xx + 0.5f - cx) / supx));
        if (wx <= 0.f) continue;
        const int sx = min(max(xx, 0), cw - 1) + x0;
        const uint8_t* px = img + ((long)sy * SW + sx) * 3;
        const float w = wx * wy;
        acc0 = fmaf(w, (float)px[0], acc0);
        acc1 = fmaf(w, (float)px[1], acc1);
        acc2 = fmaf(w, (float)px[2], acc2);
        wsum += w;
      }
    }
    const float inv = 1.f / wsum;
    const float v0 = (acc0 * inv - mean) * inv_std;
    const float v1 = (acc1 * inv - mean) * inv_std;
    const float v2 = (acc2 * inv - mean) * inv_std;
    if (out_mode == 0) {
      uint2 o;
      o.x = pack2(v0, v1);
      o.y = pack2(v2, 0.f);
      reinterpret_cast<uint2*>(out)[i] = o;
    } else {
      float* o = static_cast<float*>(out);
      const long plane = (long)S * S;
      const long base = (long)n * 3 * plane + (long)oy * S + ox;
      o[base] = v0;
      o[base + plane] = v1;
      o[base + 2 * plane] = v2;
    }
  }
}

}  // namespace sib

extern "C" int sib_rrc_boxes_ragged(int* boxes_dev, const int* dims_dev, int B, double min_area,
                                    double max_area, unsigned long long seed,
                                    unsigned long long first_sample, int do_flip, void* stream) {
  rrc_boxes_ragged_kernel<<<(B + 127) / 128, 128, 0, ST(stream)>>>(boxes_dev, dims_dev, B, min_area,
                                                                  max_area, seed, first_sample, do_flip);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_augment_ragged(const void* packed_u8, const long* offsets_dev, const int* dims_dev,
                                  const int* boxes_dev, void* out, int B, int S, float mean, float std,
                                  int out_mode, void* stream) {
  SIB_CHECK(out_mode == 0 || out_mode == 1, "augment_ragged: out_mode must be 0 (NHWC4 bf16) or 1 (NCHW f32)");
  resample_ragged_kernel<<<ew_grid((long)B * S * S, 256), 256, 0, ST(stream)>>>(
      static_cast<const uint8_t*>(packed_u8), offsets_dev, dims_dev, boxes_dev, out, B, S, 0, 0, mean,
      1.f / std, out_mode);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_val_transform_ragged(const void* packed_u8, const long* offsets_dev,
                                        const int* dims_dev, void* out, int B, int S,
                                        int resize_shorter, float mean, float std, int out_mode,
                                        void* stream) {
  SIB_CHECK(out_mode == 0 || out_mode == 1, "val_transform_ragged: out_mode must be 0 or 1");
  SIB_CHECK(resize_shorter >= S, "val_transform_ragged: resize_shorter %d smaller than the crop %d",
            resize_shorter, S);
  resample_ragged_kernel<<<ew_grid((long)B * S * S, 256), 256, 0, ST(stream)>>>(
      static_cast<const uint8_t*>(packed_u8), offsets_dev, dims_dev, nullptr, out, B, S, resize_shorter,
      1, mean, 1.f / std, out_mode);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_one_hot(const long* labels, float* out, int B, int C, void* stream) {
  one_hot_kernel<<<ew_grid((long)B * C, 256), 256, 0, ST(stream)>>>(labels, out, B, C);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_mix_batch(const void* x, const void* prev, const int* perm_dev, void* out, int N,
                             int H, int W, int C, int PH, int PW, int layout, int mode, float lam,
                             float one_minus_lam, int h1, int w1, int h2, int w2, void* stream) {
  SIB_CHECK(layout == 0 || layout == 1, "mix_batch: layout must be 0 (NHWC bf16) or 1 (NCHW f32)");
  SIB_CHECK(mode == 0 || mode == 1, "mix_batch: mode must be 0 (mixup) or 1 (cutmix)");
  SIB_CHECK(mode == 1 || (PH == H && PW == W),
            "mix_batch: mixup needs the previous batch at the same size (%dx%d vs %dx%d)", PH, PW, H, W);
  SIB_CHECK(mode == 0 || (h1 >= 0 && w1 >= 0 && h2 <= (H < PH ? H : PH) && w2 <= (W < PW ? W : PW)),
            "mix_batch: box outside the common extent of the two batches");
  if (layout == 0) {
    SIB_CHECK((W * C) % 8 == 0 && (PW * C) % 8 == 0, "mix_batch: W*C must be a multiple of 8");
    const long total = (long)N * H * (W * C / 8);
    mix_batch_kernel<0><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
        x, prev, perm_dev, out, N, H, W, C, PH, PW, mode, lam, one_minus_lam, h1, w1, h2, w2);
  } else {
    SIB_CHECK(W % 4 == 0 && PW % 4 == 0, "mix_batch: W must be a multiple of 4");
    const long total = (long)N * C * H * (W / 4);
    mix_batch_kernel<1><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
        x, prev, perm_dev, out, N, H, W, C, PH, PW, mode, lam, one_minus_lam, h1, w1, h2, w2);
  }
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_mix_targets(const float* t, const float* prev_t, const int* perm_dev, float* out,
                               int N, int C, float w_self, float w_prev, void* stream) {
  mix_targets_kernel<<<ew_grid((long)N * C, 256), 256, 0, ST(stream)>>>(t, prev_t, perm_dev, out, N,
                                                                         C, w_self, w_prev);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_stem_pack(const void* src, void* xq, int N, int H, int W, int KW, int pad_w,
                             int src_mode, void* stream) {
  SIB_CHECK(H % 2 == 0 && W % 2 == 0, "stem_pack: image extent must be even (got %dx%d)", H, W);
  SIB_CHECK(2 * KW * 3 <= 64, "stem_pack: filter width %d too large", KW);
  const long total = (long)N * (H / 2) * (W / 2);
  if (KW == 7)
    stem_pack_kernel<7><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
        src, static_cast<__nv_bfloat16*>(xq), N, H, W, pad_w, src_mode);
  else if (KW == 3)
    stem_pack_kernel<3><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
        src, static_cast<__nv_bfloat16*>(xq), N, H, W, pad_w, src_mode);
  else
    return fail(1, "stem_pack: filter width %d not instantiated (3 or 7)", KW);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_stem_pack_weight(const float* w_oihw, void* wq, int K, int KH, int KW, int NA,
                                    int off, void* stream) {
  stem_pack_weight_kernel<<<ew_grid((long)K * NA * 64, 256), 256, 0, ST(stream)>>>(
      w_oihw, static_cast<__nv_bfloat16*>(wq), K, KH, KW, NA, off);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_stem_unpack_wgrad(const float* dwq, float* dw_oihw, int K, int KH, int KW,
                                     int NA, int off, int accumulate, void* stream) {
  stem_unpack_wgrad_kernel<<<ew_grid((long)K * 3 * KH * KW, 256), 256, 0, ST(stream)>>>(
      dwq, dw_oihw, K, KH, KW, NA, off, accumulate);
  SIB_LAUNCH_CHECK();
  return 0;
}
