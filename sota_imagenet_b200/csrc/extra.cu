// Kernels used only by the BResNet-50 variant (reference configs/_old_configs/_first_attempts/
// BResNet50_encoder.yaml:44-51 kwargs on pytorch_tools.models.resnet50):
//   * BlurPool  (antialias: [1,2,1]x[1,2,1]/16 depthwise, stride 2, zero padding 1) fwd / bwd
//   * AvgPool2d(2,2) of the anti-aliased shortcut                                   fwd / bwd
//   * ECA attention: per-sample channel means -> conv1d(k=3) over channels -> sigmoid -> scale
//   * generic per-(sample, channel) scale / shift of an NHWC tensor (ECA scale, drop-connect)
//   * add + activation and its backward mask
// All NHWC bf16, 128-bit accesses.  These are memory-bound helpers, not tensor-core work.
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

static inline int ew_grid(long n, int threads) {
  long b = (n + threads - 1) / threads;
  long cap = (long)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

__device__ __forceinline__ float blur_w(int d) { return d == 1 ? 0.5f : 0.25f; }   // [1,2,1]/4

template <typename I>   // I = int when the vector count fits 31 bits (64-bit div/mod is ~10x the cost)
__global__ void __launch_bounds__(256)
blurpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                    int W, int C, int OH, int OW) {
  const int cvec = C >> 3;
  const I total = (I)N * OH * OW * cvec;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    I t = i / cvec;
    const int q = (int)(t % OW); t /= OW;
    const int p = (int)(t % OH);
    const int n = (int)(t / OH);
    float acc[8] = {};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * q - 1 + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long)n * H + h) * W + w) * C + v * 8)), f);
        const float wt = blur_w(r) * blur_w(s);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, f[j], acc[j]);
      }
    }
    stg_stream(y + (long)i * 8, pack8(acc));
  }
}

// Backward of the blur-pool on 2x2 input quads: the quad (2a + d, 2b + e) is fed by the four outputs
// (a | a+1, b | b+1) -- an even row / column only by its own output through the centre tap (1/2), an odd one by
// both neighbours through the outer taps (1/4 each) -- so one thread loads four gradients and writes four
// pixels (the per-pixel gather loaded 1 + 2 + 2 + 4 of them and ran its index arithmetic four times).
template <typename I>   // I = int when the vector count fits 31 bits (64-bit div/mod is ~10x the cost)
__global__ void __launch_bounds__(256)
blurpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                    int H, int W, int C, int OH, int OW) {
  const int cvec = C >> 3;
  const int HP = (H + 1) >> 1, WP = (W + 1) >> 1;
  const I total = (I)N * HP * WP * cvec;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    I t = i / cvec;
    const int b = (int)(t % WP); t /= WP;
    const int a = (int)(t % HP);
    const int n = (int)(t / HP);
    float g[2][2][8];
#pragma unroll
    for (int pi = 0; pi < 2; ++pi)
#pragma unroll
      for (int qi = 0; qi < 2; ++qi) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[pi][qi][j] = 0.f;
        if (a + pi < OH && b + qi < OW)
          unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((((long)n * OH + a + pi) * OW + b + qi) * cvec + v) * 8)),
                  g[pi][qi]);
      }
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // weights as in the forward: blur_w(r) * blur_w(s) with r = 1 for the even row (d = 0), r = 2 / 0 for the
      // odd row's two outputs; products of 1/2 and 1/4 are exact, the sums are accumulated in the same order
      // (p ascending, then q ascending) as the per-pixel gather did
      o[0][j] = 0.25f * g[0][0][j];
      o[1][j] = fmaf(0.125f, g[0][1][j], 0.125f * g[0][0][j]);
      o[2][j] = fmaf(0.125f, g[1][0][j], 0.125f * g[0][0][j]);
      o[3][j] = fmaf(0.0625f, g[1][1][j], fmaf(0.0625f, g[1][0][j], fmaf(0.0625f, g[0][1][j], 0.0625f * g[0][0][j])));
    }
    const int h = 2 * a, w = 2 * b;
    __nv_bfloat16* d0 = dx + ((((long)n * H + h) * W + w) * cvec + v) * 8;
    stg_stream(d0, pack8(o[0]));
    if (w + 1 < W) stg_stream(d0 + C, pack8(o[1]));
    if (h + 1 < H) {
      stg_stream(d0 + (long)W * C, pack8(o[2]));
      if (w + 1 < W) stg_stream(d0 + (long)W * C + C, pack8(o[3]));
    }
  }
}
template <typename I>   // I = int when the vector count fits 31 bits (64-bit div/mod is ~10x the cost)
__global__ void __launch_bounds__(256)
avgpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                    int W, int C) {
  const int cvec = C >> 3, OH = H >> 1, OW = W >> 1;
  const I total = (I)N * OH * OW * cvec;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    I t = i / cvec;
    const int q = (int)(t % OW); t /= OW;
    const int p = (int)(t % OH);
    const int n = (int)(t / OH);
    float acc[8] = {};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        float f[8];
        unpack8(ldg_stream(x + (((long)n * H + 2 * p + r) * W + 2 * q + s) * C + v * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    stg_stream(y + (long)i * 8, pack8(acc));
  }
}

// one thread per OUTPUT gradient: one load, four stores (H and W are even)
template <typename I>   // I = int when the vector count fits 31 bits (64-bit div/mod is ~10x the cost)
__global__ void __launch_bounds__(256)
avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                    int H, int W, int C) {
  const int cvec = C >> 3, OH = H >> 1, OW = W >> 1;
  const I total = (I)N * OH * OW * cvec;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    I t = i / cvec;
    const int q = (int)(t % OW); t /= OW;
    const int p = (int)(t % OH);
    const int n = (int)(t / OH);
    float g[8];
    unpack8(ldg_stream(dy + (long)i * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
    const uint4 o = pack8(g);
    __nv_bfloat16* d0 = dx + ((((long)n * H + 2 * p) * W + 2 * q) * cvec + v) * 8;
    stg_stream(d0, o);
    stg_stream(d0 + C, o);
    stg_stream(d0 + (long)W * C, o);
    stg_stream(d0 + (long)W * C + C, o);
  }
}

// out[n][c] = sum_hw a[n][hw][c] * (b ? b[n][hw][c] : 1)       one CTA per (sample, 64-channel group)
__global__ void __launch_bounds__(256)
chan_reduce_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                   float* __restrict__ out, int HW, int C, float scale) {
  __shared__ float red[32][8][8];
  const int n = blockIdx.y;
  const int v = blockIdx.x * 8 + (threadIdx.x & 7);     // 8-channel vector index
  const int lane_row = threadIdx.x >> 3;                // 32 row lanes
  const int cvec = C >> 3;
  float acc[8] = {};
  if (v < cvec) {
    for (int t = lane_row; t < HW; t += 32) {
      const long off = (((long)n * HW + t) * cvec + v) * 8;
      float fa[8];
      unpack8(ldg_stream(a + off), fa);
      if (b != nullptr) {
        float fb[8];
        unpack8(ldg_stream(b + off), fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(fa[j], fb[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += fa[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lane_row][threadIdx.x & 7][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int vv = threadIdx.x >> 3, j = threadIdx.x & 7;
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][vv][j];
    const int c = (blockIdx.x * 8 + vv) * 8 + j;
    if (c < C) out[(long)n * C + c] = s * scale;
  }
}

// Block-tail backward in one pass: g = dy * act'(y) is written out (it is also the shortcut
// gradient) and, per (sample, channel), s1 = sum_hw g and s2 = sum_hw g * x are reduced -- what the
// ECA gate's backward and (through them) the BatchNorm-backward sums of the block's last BN need.
// Same CTA layout as chan_reduce_kernel; two rows in flight per thread.
__global__ void __launch_bounds__(256)
act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                      const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ g,
                      float* __restrict__ s1, float* __restrict__ s2, int HW, int C, int act,
                      float slope) {
  __shared__ float red[2][32][8][8];
  const int n = blockIdx.y;
  const int v = blockIdx.x * 8 + (threadIdx.x & 7);
  const int lane_row = threadIdx.x >> 3;
  const int cvec = C >> 3;
  float a1[8] = {}, a2[8] = {};
  if (v < cvec) {
    for (int t = lane_row; t < HW; t += 64) {
      uint4 qd[2], qy[2], qx[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tt = t + 32 * u;
        if (tt < HW) {
          const long off = (((long)n * HW + tt) * cvec + v) * 8;
          qd[u] = ldg_stream(dy + off);
          qy[u] = ldg_stream(y + off);
          qx[u] = ldg_stream(x + off);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tt = t + 32 * u;
        if (tt < HW) {
          const long off = (((long)n * HW + tt) * cvec + v) * 8;
          float fd[8], fy[8], fx[8];
          unpack8(qd[u], fd);
          unpack8(qy[u], fy);
          unpack8(qx[u], fx);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float gg = fd[j];
            if (act == SIB_ACT_RELU) gg = fy[j] > 0.f ? gg : 0.f;
            else if (act == SIB_ACT_LEAKY) gg = fy[j] > 0.f ? gg : gg * slope;
            // sums of the values AS STORED (bf16), like every other statistic of the pipeline
            gg = __bfloat162float(__float2bfloat16_rn(gg));
            fd[j] = gg;
            a1[j] += gg;
            a2[j] = fmaf(gg, fx[j], a2[j]);
          }
          stg_stream(g + off, pack8(fd));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][lane_row][threadIdx.x & 7][j] = a1[j];
    red[1][lane_row][threadIdx.x & 7][j] = a2[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = threadIdx.x >> 6, vv = (threadIdx.x >> 3) & 7, j = threadIdx.x & 7;
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[k][r][vv][j];
    const int c = (blockIdx.x * 8 + vv) * 8 + j;
    if (c < C) (k == 0 ? s1 : s2)[(long)n * C + c] = s;
  }
}

// y[n][hw][c] = x * mul[n][c] (+ add[n][c])
__global__ void __launch_bounds__(256)
scale_nc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mul,
                const float* __restrict__ add, __nv_bfloat16* __restrict__ y, int N, int HW, int C) {
  const int cvec = C >> 3;
  const long total = (long)N * HW * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long n = i / ((long)HW * cvec);
    float f[8];
    unpack8(ldg_stream(x + i * 8), f);
    const float* m = mul + n * C + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= __ldg(m + j);
    if (add != nullptr) {
      const float* ad = add + n * C + v * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += __ldg(ad + j);
    }
    stg_stream(y + i * 8, pack8(f));
  }
}

// y[n][hw][c] = act(x[n][hw][c] * mul[n][c] + res[n][hw][c]): the BResNet block tail
// (ECA gate * drop-connect keep, shortcut add, activation) in one pass
__global__ void __launch_bounds__(256)
scale_add_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mul,
                     const float* __restrict__ add, const __nv_bfloat16* __restrict__ res,
                     __nv_bfloat16* __restrict__ y, int N, int HW, int C, int act, float slope) {
  const int cvec = C >> 3;
  const long total = (long)N * HW * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    const long n = i / ((long)HW * cvec);
    float f[8], r[8];
    unpack8(ldg_stream(x + i * 8), f);
    unpack8(ldg_stream(res + i * 8), r);
    float m[8], ad[8];
    load8f(mul + n * C + v * 8, m);
    if (add != nullptr) load8f(add + n * C + v * 8, ad);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // (x * mul + add) is the gated BatchNorm output when x is the RAW conv output and mul / add
      // carry the BN scale / shift times the gate; then the shortcut is added
      float t = add != nullptr ? fmaf(f[j], m[j], ad[j]) + r[j] : fmaf(f[j], m[j], r[j]);
      if (act == SIB_ACT_RELU) t = fmaxf(t, 0.f);
      else if (act == SIB_ACT_LEAKY) t = t > 0.f ? t : t * slope;
      f[j] = t;
    }
    stg_stream(y + i * 8, pack8(f));
  }
}

// ECA gate: s[n][c] = sigmoid(sum_k w[k] * p[n][c + k - 1])   (conv1d, kernel 3, zero padding 1)
__global__ void eca_gate_fwd_kernel(const float* __restrict__ p, const float* __restrict__ w,
                                    float* __restrict__ s, int N, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int c = i % C;
  float z = w[1] * p[i];
  if (c > 0) z += w[0] * p[i - 1];
  if (c + 1 < C) z += w[2] * p[i + 1];
  s[i] = 1.f / (1.f + __expf(-z));
}

// ds -> dz = ds*s*(1-s); dp[n][c] = sum_k w[k] * dz[n][c - k + 1]; dw[k] += sum dz[n][c] p[n][c+k-1]
__global__ void eca_gate_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ s,
                                    const float* __restrict__ p, const float* __restrict__ w,
                                    float* __restrict__ dp, float* __restrict__ dw, int N, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (i < N * C) {
    const int c = i % C;
    auto dz = [&](int idx) { const float sv = s[idx]; return ds[idx] * sv * (1.f - sv); };
    const float z1 = dz(i);
    float acc = w[1] * z1;
    if (c + 1 < C) acc += w[0] * dz(i + 1);
    if (c > 0) acc += w[2] * dz(i - 1);
    dp[i] = acc;
    g1 = z1 * p[i];
    if (c > 0) g0 = z1 * p[i - 1];
    if (c + 1 < C) g2 = z1 * p[i + 1];
  }
  g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(dw + 0, g0); atomicAdd(dw + 1, g1); atomicAdd(dw + 2, g2);
  }
}

// Fused block-tail coefficient kernels (BResNet, bresnet.FUSE_BN3_TAIL): all the [N][C]-sized algebra
// between the wide passes in ONE launch per direction instead of a dozen elementwise launches.
//   forward : p = pc * scale + shift (= mean_hw of bn3's output), gate = sigmoid(conv1d_k3(p)),
//             gate_k = gate * keep[n]; mul = gate_k * scale, add = gate_k * shift for scale_add_act
__global__ void eca_tail_fwd_kernel(const float* __restrict__ pc, const float* __restrict__ ss,
                                    const float* __restrict__ w, const float* __restrict__ keep,
                                    float* __restrict__ p, float* __restrict__ gate,
                                    float* __restrict__ gate_k, float* __restrict__ mul,
                                    float* __restrict__ add, int N, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int c = i % C, n = i / C;
  auto pv = [&](int cc) { return fmaf(pc[n * C + cc], ss[cc], ss[C + cc]); };
  const float p1 = pv(c);
  float z = w[1] * p1;
  if (c > 0) z += w[0] * pv(c - 1);
  if (c + 1 < C) z += w[2] * pv(c + 1);
  const float g = 1.f / (1.f + __expf(-z));
  const float gk = keep != nullptr ? g * keep[n] : g;
  p[i] = p1;
  gate[i] = g;
  gate_k[i] = gk;
  mul[i] = gk * ss[c];
  add[i] = gk * ss[C + c];
}

//   backward: from s1 = sum_hw g, s2 = sum_hw g * c3 (act_bwd_reduce):
//             ds = (scale * s2 + shift * s1) * keep      (d/d gate of y3 * gate * keep, y3 = c3*scale+shift)
//             dp = conv1d backward of ds * gate * (1 - gate); dw[k] += sum dz * p[c + k - 1]
//             add_nc = dp / HW                           (pooled-path gradient w.r.t. y3)
//             BatchNorm-backward sums of d = g * gate_k + add_nc:
//               sums[0][c] += gate_k * s1 + add_nc * HW
//               sums[1][c] += gate_k * invstd * (s2 - mean * s1) + add_nc * invstd * HW * (pc - mean)
__global__ void __launch_bounds__(256)
eca_tail_bwd_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                    const float* __restrict__ ss, const float* __restrict__ mi,
                    const float* __restrict__ pc, const float* __restrict__ p,
                    const float* __restrict__ gate, const float* __restrict__ gate_k,
                    const float* __restrict__ keep, const float* __restrict__ w,
                    float* __restrict__ add_nc, float* __restrict__ sums,
                    float* __restrict__ dw, int N, int C, float hw, int n_per) {
  // a block owns 32 channels of a RANGE of samples (blockIdx.y; 8 sample lanes): the per-channel sums
  // over its samples are combined in shared memory and published with one atomic per (block, channel,
  // sum) -- at most gridDim.y adds per address; the three filter gradients with 3 atomics per block.
  // The launcher sizes the grid to ~128 blocks: the former one-block-per-32-channels layout walked all
  // 256 samples in 32 dependent rounds of global loads (48 us per call on 8-64 blocks).
  __shared__ float red[2][8][32];
  __shared__ float redw[3][8];
  const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int n_begin = blockIdx.y * n_per;
  const int n_end = min(N, n_begin + n_per);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const float mean = mi[c], invstd = mi[C + c];
    const float w0 = w[0], w1 = w[1], w2 = w[2];
    for (int n = n_begin + nl; n < n_end; n += 8) {
      const int i = n * C + c;
      const float kp = keep != nullptr ? keep[n] : 1.f;
      auto dz = [&](int cc) {
        const int j = n * C + cc;
        const float ds = fmaf(ss[cc], s2[j], ss[C + cc] * s1[j]) * kp;
        const float sv = gate[j];
        return ds * sv * (1.f - sv);
      };
      const float z1 = dz(c);
      float dp = w1 * z1;
      if (c + 1 < C) dp += w0 * dz(c + 1);
      if (c > 0) dp += w2 * dz(c - 1);
      g1 = fmaf(z1, p[i], g1);
      if (c > 0) g0 = fmaf(z1, p[i - 1], g0);
      if (c + 1 < C) g2 = fmaf(z1, p[i + 1], g2);
      const float a = dp / hw;
      add_nc[i] = a;
      const float gk = gate_k[i];
      a0 += fmaf(gk, s1[i], a * hw);
      a1 += fmaf(gk, invstd * (s2[i] - mean * s1[i]), a * invstd * hw * (pc[i] - mean));
    }
  }
  red[0][nl][cl] = a0;
  red[1][nl][cl] = a1;
  g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
  if (cl == 0) { redw[0][nl] = g0; redw[1][nl] = g1; redw[2][nl] = g2; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int k = threadIdx.x >> 5;
    const int cc = blockIdx.x * 32 + cl;
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[k][r][cl];
    if (cc < C) {
      if (gridDim.y == 1) sums[k * C + cc] += t;     // only writer of these channels
      else atomicAdd(sums + k * C + cc, t);
    }
  } else if (threadIdx.x < 67) {
    const int k = threadIdx.x - 64;
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += redw[k][r];
    atomicAdd(dw + k, t);
  }
}

// y = act(a + b);   backward: g = dy * act'(y)
__global__ void __launch_bounds__(256)
add_act_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
               __nv_bfloat16* __restrict__ y, long nvec, int act, float slope) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8(ldg_stream(a + i * 8), fa);
    unpack8(ldg_stream(b + i * 8), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fa[j] + fb[j];
      if (act == SIB_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == SIB_ACT_LEAKY) v = v > 0.f ? v : v * slope;
      fa[j] = v;
    }
    stg_stream(y + i * 8, pack8(fa));
  }
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
               __nv_bfloat16* __restrict__ g, long nvec, int act, float slope) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long)gridDim.x * blockDim.x) {
    float fd[8], fy[8];
    unpack8(ldg_stream(dy + i * 8), fd);
    unpack8(ldg_stream(y + i * 8), fy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (act == SIB_ACT_RELU) fd[j] = fy[j] > 0.f ? fd[j] : 0.f;
      else if (act == SIB_ACT_LEAKY) fd[j] = fy[j] > 0.f ? fd[j] : fd[j] * slope;
    }
    stg_stream(g + i * 8, pack8(fd));
  }
}

// 3x3 stride-1 pad-1 max pool (the anti-aliased stem pools at stride 1, then BlurPool).
// A thread owns one 8-channel vector of one image COLUMN over a segment of rows and slides down it:
// per output row it loads the three pixels of the next input row once, reduces them to a row maximum
// (+ the tap column that won) and combines the three live row maxima -- 3 loads and ~50 compares per
// output instead of 9 loads and 72 compares, 32-bit index arithmetic outside the loop.  The winning
// tap r*3+s is the FIRST maximum in (r, s) scan order (strict >), as torch's max_pool2d indices.
__device__ __forceinline__ void mp_row_max(const __nv_bfloat16* __restrict__ x, long row_base, int C,
                                           int q, int W, bool row_ok, float* val, uint32_t& sel) {
#pragma unroll
  for (int j = 0; j < 8; ++j) val[j] = -INFINITY;
  sel = 0;                                    // 4 bits per channel: winning column tap s
  if (!row_ok) return;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int w = q - 1 + s;
    if (w < 0 || w >= W) continue;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + row_base + (long)w * C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (f[j] > val[j]) { val[j] = f[j]; sel = (sel & ~(0xfu << (4 * j))) | ((uint32_t)s << (4 * j)); }
  }
}

__global__ void __launch_bounds__(256)
maxpool3x3s1_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                        uint8_t* __restrict__ idx, int N, int H, int W, int C, int segs, int seg_rows) {
  const int cvec = C >> 3;
  const int total = N * segs * W * cvec;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cvec;
    int t = i / cvec;
    const int q = t % W; t /= W;
    const int seg = t % segs;
    const int n = t / segs;
    const int p0 = seg * seg_rows, p1 = min(H, p0 + seg_rows);
    const long img = (long)n * H * W * C + v * 8;          // + (h * W + w) * C
    float va[8], vb[8], vc[8];
    uint32_t sa, sb, sc;
    mp_row_max(x, img + (long)(p0 - 1) * W * C, C, q, W, p0 - 1 >= 0, va, sa);
    mp_row_max(x, img + (long)p0 * W * C, C, q, W, true, vb, sb);
    for (int p = p0; p < p1; ++p) {
      mp_row_max(x, img + (long)(p + 1) * W * C, C, q, W, p + 1 < H, vc, sc);
      float best[8];
      uint32_t bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        best[j] = va[j];
        bi[j] = (sa >> (4 * j)) & 3u;
        if (vb[j] > best[j]) { best[j] = vb[j]; bi[j] = 3u + ((sb >> (4 * j)) & 3u); }
        if (vc[j] > best[j]) { best[j] = vc[j]; bi[j] = 6u + ((sc >> (4 * j)) & 3u); }
      }
      const long o = img + ((long)p * W + q) * C;
      stg_stream(y + o, pack8(best));
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + o) = pk;
#pragma unroll
      for (int j = 0; j < 8; ++j) { va[j] = vb[j]; vb[j] = vc[j]; }
      sa = sb; sb = sc;
    }
  }
}

// Backward, same column-sliding layout: source row p holds, for this output column, three candidate
// pixels (q = w+1-s); a source whose winning tap is (r, s) sends its gradient to output row p-1+r.
// Three rolling accumulators (rows p-1, p, p+1); row p-1 is complete once source row p is consumed.
__global__ void __launch_bounds__(256)
maxpool3x3s1_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                        __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int segs,
                        int seg_rows) {
  const int cvec = C >> 3;
  const int total = N * segs * W * cvec;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % cvec;
    int t = i / cvec;
    const int w = t % W; t /= W;
    const int seg = t % segs;
    const int n = t / segs;
    const int h0 = seg * seg_rows, h1 = min(H, h0 + seg_rows);
    const long img = (long)n * H * W * C + v * 8;
    float aa[8] = {}, ab[8] = {}, ac[8];
    for (int p = h0 - 1; p <= h1; ++p) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ac[j] = 0.f;
      if (p >= 0 && p < H) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int q = w + 1 - s;
          if (q < 0 || q >= W) continue;
          const long o = img + ((long)p * W + q) * C;
          const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + o));
          float g[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t tap = ((j < 4 ? (pk.x >> (8 * j)) : (pk.y >> (8 * (j - 4)))) & 0xffu) - (uint32_t)s;
            aa[j] += tap == 0u ? g[j] : 0.f;      // r = 0 -> output row p - 1
            ab[j] += tap == 3u ? g[j] : 0.f;      // r = 1 -> row p
            ac[j] += tap == 6u ? g[j] : 0.f;      // r = 2 -> row p + 1
          }
        }
      }
      if (p - 1 >= h0) stg_stream(dx + img + ((long)(p - 1) * W + w) * C, pack8(aa));   // (p - 1 < h1 by the loop bound)
#pragma unroll
      for (int j = 0; j < 8; ++j) { aa[j] = ab[j]; ab[j] = ac[j]; }
    }
  }
}

}  // namespace sib

using namespace sib;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) static_cast<__nv_bfloat16*>(p)
#define CBF(p) static_cast<const __nv_bfloat16*>(p)

extern "C" int sib_blurpool_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "blurpool: C %% 8 != 0");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long total = (long)N * OH * OW * (C / 8);
  if (total < (1l << 31) - (1l << 24)) blurpool_fwd_kernel<int><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, H, W, C, OH, OW);
  else blurpool_fwd_kernel<long><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, H, W, C, OH, OW);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_blurpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "blurpool: C %% 8 != 0");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long total = (long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  if (total < (1l << 31) - (1l << 24)) blurpool_bwd_kernel<int><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), N, H, W, C, OH, OW);
  else blurpool_bwd_kernel<long><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), N, H, W, C, OH, OW);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_avgpool2_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  SIB_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: needs C %% 8 == 0 and even H, W");
  const long total = (long)N * (H / 2) * (W / 2) * (C / 8);
  if (total < (1l << 31) - (1l << 24)) avgpool2_fwd_kernel<int><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, H, W, C);
  else avgpool2_fwd_kernel<long><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, H, W, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_avgpool2_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
  SIB_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: needs C %% 8 == 0 and even H, W");
  const long total = (long)N * (H / 2) * (W / 2) * (C / 8);
  if (total < (1l << 31) - (1l << 24)) avgpool2_bwd_kernel<int><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), N, H, W, C);
  else avgpool2_bwd_kernel<long><<<ew_grid(total, 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), N, H, W, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_chan_reduce(const void* a, const void* b, float* out, int N, int HW, int C,
                               float scale, void* stream) {
  SIB_CHECK(C % 8 == 0, "chan_reduce: C %% 8 != 0");
  dim3 grid((C / 8 + 7) / 8, N);
  chan_reduce_kernel<<<grid, 256, 0, ST(stream)>>>(CBF(a), CBF(b), out, HW, C, scale);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_eca_tail_fwd(const float* pc, const float* scale_shift, const float* w, const float* keep,
                                float* p, float* gate, float* gate_k, float* mul, float* add, int N, int C,
                                void* stream) {
  eca_tail_fwd_kernel<<<(N * C + 255) / 256, 256, 0, ST(stream)>>>(pc, scale_shift, w, keep, p, gate, gate_k,
                                                                   mul, add, N, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
// sums [2][C] and dw [3] must be zero on entry (they are accumulated with atomics)
extern "C" int sib_eca_tail_bwd(const float* s1, const float* s2, const float* scale_shift,
                                const float* mean_invstd, const float* pc, const float* p, const float* gate,
                                const float* gate_k, const float* keep, const float* w, float* add_nc,
                                float* sums, float* dw, int N, int C, float hw, void* stream) {
  // ~128 blocks: channel groups x sample ranges (multiples of the 8 sample lanes); SIB_DETERMINISTIC=1
  // keeps one block per channel group (fixed summation order, no atomics on `sums`)
  static const bool det = [] { const char* e = getenv("SIB_DETERMINISTIC"); return e && e[0] == '1'; }();
  const int cg = (C + 31) / 32;
  int ns = det ? 1 : max(1, min(128 / cg, (N + 7) / 8));
  const int n_per = ((N + ns - 1) / ns + 7) / 8 * 8;
  ns = (N + n_per - 1) / n_per;
  eca_tail_bwd_kernel<<<dim3(cg, ns), 256, 0, ST(stream)>>>(s1, s2, scale_shift, mean_invstd, pc, p, gate,
                                                            gate_k, keep, w, add_nc, sums, dw, N, C, hw, n_per);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_act_bwd_reduce(const void* dy, const void* y, const void* x, void* g, float* s1,
                                  float* s2, int N, int HW, int C, int act, float slope, void* stream) {
  SIB_CHECK(C % 8 == 0, "act_bwd_reduce: C %% 8 != 0");
  dim3 grid((C / 8 + 7) / 8, N);
  act_bwd_reduce_kernel<<<grid, 256, 0, ST(stream)>>>(CBF(dy), CBF(y), CBF(x), BF(g), s1, s2, HW, C, act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_scale_nc(const void* x, const float* mul, const float* add, void* y, int N,
                            int HW, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "scale_nc: C %% 8 != 0");
  scale_nc_kernel<<<ew_grid((long)N * HW * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(x), mul, add,
                                                                               BF(y), N, HW, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_scale_add_act(const void* x, const float* mul, const float* add, const void* res,
                                 void* y, int N, int HW, int C, int act, float slope, void* stream) {
  SIB_CHECK(C % 8 == 0, "scale_add_act: C %% 8 != 0");
  scale_add_act_kernel<<<ew_grid((long)N * HW * (C / 8), 256), 256, 0, ST(stream)>>>(
      CBF(x), mul, add, CBF(res), BF(y), N, HW, C, act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_eca_gate_fwd(const float* p, const float* w, float* s, int N, int C,
                                void* stream) {
  eca_gate_fwd_kernel<<<(N * C + 255) / 256, 256, 0, ST(stream)>>>(p, w, s, N, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_eca_gate_bwd(const float* ds, const float* s, const float* p, const float* w,
                                float* dp, float* dw, int N, int C, void* stream) {
  eca_gate_bwd_kernel<<<(N * C + 255) / 256, 256, 0, ST(stream)>>>(ds, s, p, w, dp, dw, N, C);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_add_act(const void* a, const void* b, void* y, long n, int act, float slope,
                           void* stream) {
  SIB_CHECK(n % 8 == 0, "add_act: length %% 8 != 0");
  add_act_kernel<<<ew_grid(n / 8, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), BF(y), n / 8, act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_act_bwd(const void* dy, const void* y, void* g, long n, int act, float slope,
                           void* stream) {
  SIB_CHECK(n % 8 == 0, "act_bwd: length %% 8 != 0");
  act_bwd_kernel<<<ew_grid(n / 8, 256), 256, 0, ST(stream)>>>(CBF(dy), CBF(y), BF(g), n / 8, act, slope);
  SIB_LAUNCH_CHECK();
  return 0;
}
// row segments per image column so that ~600 k threads are in flight (each re-reads 2 halo rows)
static void maxpool_segments(int N, int H, int W, int C, int* segs, int* seg_rows) {
  const long cols = (long)N * W * (C / 8);
  long want = (600000 + cols - 1) / cols;
  if (want < 1) want = 1;
  if (want > (H + 7) / 8) want = (H + 7) / 8;          // at least 8 rows per segment
  *seg_rows = (int)((H + want - 1) / want);
  *segs = (H + *seg_rows - 1) / *seg_rows;
}

extern "C" int sib_maxpool3x3s1_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C,
                                    void* stream) {
  SIB_CHECK(C % 8 == 0, "maxpool3x3s1: C %% 8 != 0");
  int segs, seg_rows;
  maxpool_segments(N, H, W, C, &segs, &seg_rows);
  SIB_CHECK((long)N * segs * W * (C / 8) < (1l << 31) - (1l << 24), "maxpool3x3s1: tensor too large for 32-bit indexing");
  maxpool3x3s1_fwd_kernel<<<ew_grid((long)N * segs * W * (C / 8), 256), 256, 0, ST(stream)>>>(
      CBF(x), BF(y), static_cast<uint8_t*>(idx), N, H, W, C, segs, seg_rows);
  SIB_LAUNCH_CHECK();
  return 0;
}
extern "C" int sib_maxpool3x3s1_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W,
                                    int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "maxpool3x3s1: C %% 8 != 0");
  int segs, seg_rows;
  maxpool_segments(N, H, W, C, &segs, &seg_rows);
  SIB_CHECK((long)N * segs * W * (C / 8) < (1l << 31) - (1l << 24), "maxpool3x3s1: tensor too large for 32-bit indexing");
  maxpool3x3s1_bwd_kernel<<<ew_grid((long)N * segs * W * (C / 8), 256), 256, 0, ST(stream)>>>(
      CBF(dy), static_cast<const uint8_t*>(idx), BF(dx), N, H, W, C, segs, seg_rows);
  SIB_LAUNCH_CHECK();
  return 0;
}
