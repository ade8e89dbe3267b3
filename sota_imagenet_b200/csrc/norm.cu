// Memory-bound NHWC bf16 kernels around the convolutions: BatchNorm statistics,
// finalize (+running stats), apply (+activation, +residual, +second BN), backward
// reduce / apply, 3x3/2 max-pool and global average pool.  All global traffic is
// 128-bit; per-channel sums use register accumulation -> smem -> fp32 atomics.
//
// Arithmetic follows nn.BatchNorm2d in training mode, which is what
// pytorch_tools' ABN / the reference's `patch_bn_mom` path run (reference train.py:76).
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

constexpr int kRedThreads = 256;

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == SIB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SIB_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}
__device__ __forceinline__ float act_grad(float out, int act, float slope) {
  if (act == SIB_ACT_RELU) return out > 0.f ? 1.f : 0.f;
  if (act == SIB_ACT_LEAKY) return out > 0.f ? 1.f : slope;
  return 1.f;
}

// ---------------------------------------------------------------------------
// per-channel sum / sum of squares of x[M][C]
// ---------------------------------------------------------------------------
// thread layout: tx = vector (8 channels) within the row, ty = row lane.
// Requires C % 8 == 0 and C/8 <= kRedThreads.
template <int NACC, class F>
__device__ __forceinline__ void channel_reduce(long M, int C, float* __restrict__ out, F body) {
  const int cvec = C >> 3;
  const int rpb = kRedThreads / cvec;       // rows handled per block iteration
  const int tx = threadIdx.x % cvec;
  const int ty = threadIdx.x / cvec;
  float acc[NACC][8];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
  if (ty < rpb) {
    for (long r = (long)blockIdx.x * rpb + ty; r < M; r += (long)gridDim.x * rpb)
      body(r, tx, acc);
  }
  __shared__ float red[NACC][kRedThreads * 8 / 8][8];   // [acc][thread][8]
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[a][threadIdx.x][j] = acc[a][j];
  __syncthreads();
  // first `cvec*8` threads each own one channel and sum over the row lanes
  for (int idx = threadIdx.x; idx < NACC * C; idx += kRedThreads) {
    const int a = idx / C;
    const int c = idx - a * C;
    const int v = c >> 3, j = c & 7;
    float s = 0.f;
    for (int y = 0; y < rpb; ++y) s += red[a][y * cvec + v][j];
    atomicAdd(out + (long)a * C + c, s);
  }
}

__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long M, int C, float* __restrict__ stats) {
  channel_reduce<2>(M, C, stats, [&](long r, int tx, float (&acc)[2][8]) {
    float f[8];
    unpack8(ldg_stream(x + r * C + tx * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += f[j];
      acc[1][j] += f[j] * f[j];
    }
  });
}

// ---------------------------------------------------------------------------
// finalize: stats -> mean / invstd / fused scale & shift, running-stat update
// ---------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_invstd,
                                   float* __restrict__ scale_shift, int C, float count,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[c] / count;
  float var = stats[C + c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + eps);
  mean_invstd[c] = mean;
  mean_invstd[C + c] = invstd;
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * invstd;
  scale_shift[C + c] = b - mean * g * invstd;
  if (running_mean) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// eval mode: scale/shift from running statistics
__global__ void bn_eval_scale_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var,
                                     float* __restrict__ scale_shift, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * invstd;
  scale_shift[C + c] = b - running_mean[c] * g * invstd;
}

// ---------------------------------------------------------------------------
// apply: y = act(x*scale + shift [+ res | + res*scale2 + shift2])
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ ss,
                const __nv_bfloat16* __restrict__ res, const float* __restrict__ ss2,
                __nv_bfloat16* __restrict__ y, long nvec, int C, int act, float slope) {
  const int cvec = C >> 3;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) * 8;
    float f[8], o[8];
    unpack8(ldg_stream(x + i * 8), f);
    const float4 sa = __ldg(reinterpret_cast<const float4*>(ss + c0));
    const float4 sb = __ldg(reinterpret_cast<const float4*>(ss + c0 + 4));
    const float4 ha = __ldg(reinterpret_cast<const float4*>(ss + C + c0));
    const float4 hb = __ldg(reinterpret_cast<const float4*>(ss + C + c0 + 4));
    const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], sc[j], sh[j]);
    if (res != nullptr) {
      float rr[8];
      unpack8(ldg_stream(res + i * 8), rr);
      if (ss2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] += fmaf(rr[j], __ldg(ss2 + c0 + j), __ldg(ss2 + C + c0 + j));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rr[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_fwd(o[j], act, slope);
    stg_stream(y + i * 8, pack8(o));
  }
}

// ---------------------------------------------------------------------------
// backward reduce: g = dy * act'(out);  sums[0] = sum g, sums[1] = sum g * xhat
// optional second BN (downsample branch): sums[2], sums[3] with x2 / mean_invstd2
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ out,
                     const __nv_bfloat16* __restrict__ x, const float* __restrict__ mi,
                     const __nv_bfloat16* __restrict__ x2, const float* __restrict__ mi2, long M,
                     int C, int act, float slope, float* __restrict__ sums) {
  if (x2 == nullptr) {
    channel_reduce<2>(M, C, sums, [&](long r, int tx, float (&acc)[2][8]) {
      float g[8], o[8], xv[8];
      const long off = r * C + tx * 8;
      unpack8(ldg_stream(dy + off), g);
      unpack8(ldg_stream(x + off), xv);
      if (act != SIB_ACT_NONE) {
        unpack8(ldg_stream(out + off), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        const float xh = (xv[j] - __ldg(mi + c)) * __ldg(mi + C + c);
        acc[0][j] += g[j];
        acc[1][j] += g[j] * xh;
      }
    });
  } else {
    channel_reduce<4>(M, C, sums, [&](long r, int tx, float (&acc)[4][8]) {
      float g[8], o[8], xv[8], xw[8];
      const long off = r * C + tx * 8;
      unpack8(ldg_stream(dy + off), g);
      unpack8(ldg_stream(x + off), xv);
      unpack8(ldg_stream(x2 + off), xw);
      if (act != SIB_ACT_NONE) {
        unpack8(ldg_stream(out + off), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        const float xh = (xv[j] - __ldg(mi + c)) * __ldg(mi + C + c);
        const float xh2 = (xw[j] - __ldg(mi2 + c)) * __ldg(mi2 + C + c);
        acc[0][j] += g[j];
        acc[1][j] += g[j] * xh;
        acc[2][j] += g[j];
        acc[3][j] += g[j] * xh2;
      }
    });
  }
}

// ---------------------------------------------------------------------------
// backward apply: dx = gamma*invstd * (g - sum_g/m - xhat * sum_gx/m)
// writes dx (for x), optionally dx2 (second BN) and g itself (residual-branch grad)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ out,
                    const __nv_bfloat16* __restrict__ x, const float* __restrict__ mi,
                    const float* __restrict__ gamma, const float* __restrict__ sums,
                    const __nv_bfloat16* __restrict__ x2, const float* __restrict__ mi2,
                    const float* __restrict__ gamma2, __nv_bfloat16* __restrict__ dx,
                    __nv_bfloat16* __restrict__ dx2, __nv_bfloat16* __restrict__ gout, long nvec,
                    int C, float inv_count, int act, float slope) {
  const int cvec = C >> 3;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) * 8;
    float g[8], o[8], xv[8], d[8];
    unpack8(ldg_stream(dy + i * 8), g);
    unpack8(ldg_stream(x + i * 8), xv);
    if (act != SIB_ACT_NONE) {
      unpack8(ldg_stream(out + i * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float mean = __ldg(mi + c), invstd = __ldg(mi + C + c);
      const float xh = (xv[j] - mean) * invstd;
      const float gm = gamma ? __ldg(gamma + c) : 1.f;
      d[j] = gm * invstd * (g[j] - __ldg(sums + c) * inv_count - xh * __ldg(sums + C + c) * inv_count);
    }
    stg_stream(dx + i * 8, pack8(d));
    if (x2 != nullptr) {
      float xw[8];
      unpack8(ldg_stream(x2 + i * 8), xw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const float mean = __ldg(mi2 + c), invstd = __ldg(mi2 + C + c);
        const float xh = (xw[j] - mean) * invstd;
        const float gm = gamma2 ? __ldg(gamma2 + c) : 1.f;
        d[j] = gm * invstd *
               (g[j] - __ldg(sums + 2 * C + c) * inv_count - xh * __ldg(sums + 3 * C + c) * inv_count);
      }
      stg_stream(dx2 + i * 8, pack8(d));
    }
    if (gout != nullptr) stg_stream(gout + i * 8, pack8(g));
  }
}

// dgamma = sum_gx, dbeta = sum_g  (accumulated into fp32 grads)
__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (accumulate) {
    if (dbeta) dbeta[c] += sums[c];
    if (dgamma) dgamma[c] += sums[C + c];
  } else {
    if (dbeta) dbeta[c] = sums[c];
    if (dgamma) dgamma[c] = sums[C + c];
  }
}

// ---------------------------------------------------------------------------
// 3x3 stride-2 pad-1 max pool, NHWC; idx stores the winning tap (0..8) per element
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int OH, int OW) {
  const int cvec = C >> 3;
  const long total = (long)N * OH * OW * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    long t = i / cvec;
    const int q = (int)(t % OW); t /= OW;
    const int p = (int)(t % OH);
    const int n = (int)(t / OH);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = p * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = q * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long)n * H + h) * W + w) * C + v * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > best[j]) { best[j] = f[j]; bi[j] = r * 3 + s; }
      }
    }
    stg_stream(y + i * 8, pack8(best));
    if (idx != nullptr) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = pk;
    }
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int OH, int OW) {
  const int cvec = C >> 3;
  const long total = (long)N * H * W * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    long t = i / cvec;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows (p,q) covering (h,w): p*2-1+r = h  =>  r = h - 2p + 1 in [0,3)
    for (int p = (h >> 1); p <= ((h + 1) >> 1); ++p) {
      if (p < 0 || p >= OH) continue;
      const int r = h - 2 * p + 1;
      if (r < 0 || r > 2) continue;
      for (int q = (w >> 1); q <= ((w + 1) >> 1); ++q) {
        if (q < 0 || q >= OW) continue;
        const int s = w - 2 * q + 1;
        if (s < 0 || s > 2) continue;
        const long o = ((((long)n * OH + p) * OW + q) * cvec + v) * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), g);
        const int tap = r * 3 + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int b = (j < 4 ? (pk.x >> (8 * j)) : (pk.y >> (8 * (j - 4)))) & 0xff;
          if (b == tap) acc[j] += g[j];
        }
      }
    }
    stg_stream(dx + i * 8, pack8(acc));
  }
}

// ---------------------------------------------------------------------------
// global average pool [N][HW][C] -> [N][C] and its backward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int HW,
               int C) {
  const int cvec = C >> 3;
  const long total = (long)N * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long n = i / cvec;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = 0; t < HW; ++t) {
      float f[8];
      unpack8(ldg_stream(x + ((n * HW + t) * C + v * 8)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    const float inv = 1.f / HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(acc);
  }
}

__global__ void __launch_bounds__(256)
gap_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int HW,
               int C) {
  const int cvec = C >> 3;
  const long total = (long)N * HW * cvec;
  const float inv = 1.f / HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long n = i / ((long)HW * cvec);
    float g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (n * cvec + v) * 8)), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv;
    stg_stream(dx + i * 8, pack8(g));
  }
}

// zero-insertion upsample of dY for strided 3x3 dgrad: up[n][p*s][q*s][:] = dy[n][p][q][:]
__global__ void __launch_bounds__(256)
upsample_zero_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ up, int N,
                     int OH, int OW, int C, int UH, int UW, int stride) {
  const int cvec = C >> 3;
  const long total = (long)N * UH * UW * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    long t = i / cvec;
    const int w = (int)(t % UW); t /= UW;
    const int h = (int)(t % UH);
    const int n = (int)(t / UH);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (h % stride == 0 && w % stride == 0 && h / stride < OH && w / stride < OW)
      val = ldg_stream(dy + ((((long)n * OH + h / stride) * OW + w / stride) * cvec + v) * 8);
    stg_stream(up + i * 8, val);
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

static int check_c(int C) {
  SIB_CHECK(C % 8 == 0 && C / 8 <= kRedThreads, "channel count %d must be a multiple of 8 and <= %d",
            C, kRedThreads * 8);
  return 0;
}

extern "C" int sib_bn_stats(const void* x, long M, int C, float* stats, void* stream) {
  if (int rc = check_c(C)) return rc;
  SIB_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * C, ST(stream)));
  const int rpb = kRedThreads / (C / 8);
  long blocks = (M + rpb - 1) / rpb;
  long cap = (long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bn_stats_kernel<<<(int)blocks, kRedThreads, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), M, C, stats);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_finalize(const float* stats, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float* mean_invstd,
                               float* scale_shift, int C, double count, float eps, float momentum,
                               void* stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(
      stats, gamma, beta, running_mean, running_var, mean_invstd, scale_shift, C, (float)count,
      eps, momentum);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_eval_scale(const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float* scale_shift, int C, float eps,
                                 void* stream) {
  bn_eval_scale_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(gamma, beta, running_mean,
                                                                running_var, scale_shift, C, eps);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_apply(const void* x, const float* scale_shift, const void* res,
                            const float* scale_shift2, void* y, long M, int C, int act,
                            float slope, void* stream) {
  SIB_CHECK(C % 8 == 0, "bn_apply: C %% 8 != 0");
  const long nvec = M * (C / 8);
  bn_apply_kernel<<<ew_grid(nvec, 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), scale_shift, static_cast<const __nv_bfloat16*>(res),
      scale_shift2, static_cast<__nv_bfloat16*>(y), nvec, C, act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_bwd_reduce(const void* dy, const void* out, const void* x,
                                 const float* mean_invstd, const void* x2,
                                 const float* mean_invstd2, long M, int C, int act, float slope,
                                 float* sums, void* stream) {
  if (int rc = check_c(C)) return rc;
  const int nacc = x2 ? 4 : 2;
  SIB_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * nacc * C, ST(stream)));
  const int rpb = kRedThreads / (C / 8);
  long blocks = (M + rpb - 1) / rpb;
  long cap = (long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bn_bwd_reduce_kernel<<<(int)blocks, kRedThreads, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(x), mean_invstd, static_cast<const __nv_bfloat16*>(x2),
      mean_invstd2, M, C, act, slope, sums);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_bwd_apply(const void* dy, const void* out, const void* x,
                                const float* mean_invstd, const float* gamma, const float* sums,
                                const void* x2, const float* mean_invstd2, const float* gamma2,
                                void* dx, void* dx2, void* gout, long M, int C, double count,
                                int act, float slope, void* stream) {
  SIB_CHECK(C % 8 == 0, "bn_bwd_apply: C %% 8 != 0");
  const long nvec = M * (C / 8);
  bn_bwd_apply_kernel<<<ew_grid(nvec, 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(x), mean_invstd, gamma, sums,
      static_cast<const __nv_bfloat16*>(x2), mean_invstd2, gamma2,
      static_cast<__nv_bfloat16*>(dx), static_cast<__nv_bfloat16*>(dx2),
      static_cast<__nv_bfloat16*>(gout), nvec, C, (float)(1.0 / count), act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_param_grad(const float* sums, float* dgamma, float* dbeta, int C,
                                 int accumulate, void* stream) {
  bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(sums, dgamma, dbeta, C,
                                                                accumulate);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C,
                                    void* stream) {
  SIB_CHECK(C % 8 == 0, "maxpool: C %% 8 != 0");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long total = (long)N * OH * OW * (C / 8);
  maxpool_fwd_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
      static_cast<uint8_t*>(idx), N, H, W, C, OH, OW);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W,
                                    int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "maxpool: C %% 8 != 0");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long total = (long)N * H * W * (C / 8);
  maxpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const uint8_t*>(idx),
      static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH, OW);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_gap_fwd(const void* x, void* y, int N, int HW, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "gap: C %% 8 != 0");
  gap_fwd_kernel<<<ew_grid((long)N * (C / 8), 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, HW, C);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_gap_bwd(const void* dy, void* dx, int N, int HW, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "gap: C %% 8 != 0");
  gap_bwd_kernel<<<ew_grid((long)N * HW * (C / 8), 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, HW, C);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_upsample_zero(const void* dy, void* up, int N, int OH, int OW, int C, int UH,
                                 int UW, int stride, void* stream) {
  SIB_CHECK(C % 8 == 0, "upsample_zero: C %% 8 != 0");
  const long total = (long)N * UH * UW * (C / 8);
  upsample_zero_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(up), N, OH, OW, C, UH, UW,
      stride);
  SIB_LAUNCH_CHECK();
  return 0;
}
