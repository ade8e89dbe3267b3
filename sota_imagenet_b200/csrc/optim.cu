// Optimizer + weight packing kernels.
//   * sib_sgd_step: coupled-L2 SGD with momentum / Nesterov over one flat fp32 parameter arena
//     (torch/optim/sgd.py semantics, reference arg_parser.py:136-138: torch.optim._multi_tensor.SGD);
//     one launch for all 161 tensors, also emits the bf16 copy the conv kernels read and an
//     optional EMA stream (pt_clb.ModelEma, reference train.py:112).
//   * sib_pack_dgrad_weights: table-driven repack of every conv filter from KRSC to the
//     tap-flipped [C][R][S][K] layout the dgrad implicit GEMM consumes.
//   * sib_weight_standardize: per-out-channel (w - mean) / sqrt(var + eps) (reference
//     model.py:91-100 analogue of pytorch_tools conv_to_ws_conv, train.py:66-67).
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

// per-segment hyper-parameters: segment i covers elements [seg_end[i-1], seg_end[i])
struct SgdSeg {
  long end;
  float lr, weight_decay, momentum, dampening;
  int nesterov;
  int pad;
};

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
           __nv_bfloat16* __restrict__ p_bf16, float* __restrict__ ema, float ema_decay,
           const SgdSeg* __restrict__ segs, int nseg, long n4, int first_step) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    const long e = i * 4;
    // segment lookup
    int sidx = 0, hi = nseg - 1;
    while (sidx < hi) {          // first segment whose end is > e
      const int mid = (sidx + hi) >> 1;
      if (e >= segs[mid].end) sidx = mid + 1; else hi = mid;
    }
    const SgdSeg sg = segs[sidx];
    float4 pv = *reinterpret_cast<float4*>(p + e);
    const float4 gv = *reinterpret_cast<const float4*>(g + e);
    float pa[4] = {pv.x, pv.y, pv.z, pv.w};
    float ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ba[4] = {0.f, 0.f, 0.f, 0.f};
    if (sg.momentum != 0.f && !first_step) {
      const float4 bv = *reinterpret_cast<const float4*>(buf + e);
      ba[0] = bv.x; ba[1] = bv.y; ba[2] = bv.z; ba[3] = bv.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float grad = ga[j] + sg.weight_decay * pa[j];
      if (sg.momentum != 0.f) {
        // first step: buf = grad (torch clones the gradient), afterwards buf = mu*buf + (1-damp)*grad
        ba[j] = first_step ? grad : sg.momentum * ba[j] + (1.f - sg.dampening) * grad;
        grad = sg.nesterov ? grad + sg.momentum * ba[j] : ba[j];
      }
      pa[j] -= sg.lr * grad;
    }
    *reinterpret_cast<float4*>(p + e) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    if (sg.momentum != 0.f)
      *reinterpret_cast<float4*>(buf + e) = make_float4(ba[0], ba[1], ba[2], ba[3]);
    if (p_bf16 != nullptr) {
      uint2 o;
      o.x = pack2(pa[0], pa[1]);
      o.y = pack2(pa[2], pa[3]);
      *reinterpret_cast<uint2*>(p_bf16 + e) = o;
    }
    if (ema != nullptr) {
      float4 ev = *reinterpret_cast<float4*>(ema + e);
      ev.x = ema_decay * ev.x + (1.f - ema_decay) * pa[0];
      ev.y = ema_decay * ev.y + (1.f - ema_decay) * pa[1];
      ev.z = ema_decay * ev.z + (1.f - ema_decay) * pa[2];
      ev.w = ema_decay * ev.w + (1.f - ema_decay) * pa[3];
      *reinterpret_cast<float4*>(ema + e) = ev;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MyNovograd (reference sota_imagenet/optimizers.py:35-161) over one flat arena.
//   norm_t   = sum(p_t^2)                 (whole tensor; the reference squares the WEIGHTS,
//                                          optimizers.py:136, not the gradients: kept as is)
//            | ||p_unit||_2               (unitwise_norm=True: per output unit, :18-22,133-134)
//   ema_norm = beta2*ema_norm + (1-beta2)*norm_t                     (:139-140)
//   ema_grad = beta1*ema_grad + (1-beta1)*g                          (:143,147)
//   p        = (p - lr * ema_grad / (sqrt(ema_norm) + eps)) * (1 - lr*wd)   (:144-145,155,158)
// ema_norm is one scalar per norm group here (the reference stores it expanded to the
// parameter's shape, :120-121; the host side exposes an expanded view).
// Three launches per arena: segmented sum of squares, per-group norm update, elementwise update.
struct NovoTensor {
  long begin;      // first element of the tensor in the arena
  long end;        // first element of the next tensor (alignment padding included, zeros)
  int unit_len;    // elements per norm group
  int ngroups;     // norm groups of this tensor
  int norm_base;   // index of its first group
  float lr, decay, beta1, one_minus_beta1, beta2, one_minus_beta2;
  int pad;
};

__device__ __forceinline__ int novo_find(const NovoTensor* __restrict__ tab, int nt, long e) {
  int lo = 0, hi = nt - 1;
  while (lo < hi) {            // first tensor whose end is > e
    const int mid = (lo + hi) >> 1;
    if (e >= tab[mid].end) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int novo_group(const NovoTensor& t, long e) {
  const long g = (e - t.begin) / t.unit_len;
  return t.norm_base + (int)(g < t.ngroups ? g : t.ngroups - 1);   // padding joins the last group (zeros)
}

constexpr int NOVO_ROUNDS = 8;   // float4 rounds a warp walks before its segmented reduction

__global__ void __launch_bounds__(256)
novograd_sumsq_kernel(const float* __restrict__ p, const NovoTensor* __restrict__ tab, int nt,
                      long n4, float* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long base4 = warp * (32 * NOVO_ROUNDS);
  int cur = INT_MAX;             // group this lane is accumulating
  float acc = 0.f;
  for (int r = 0; r < NOVO_ROUNDS; ++r) {
    const long i = base4 + r * 32 + lane;
    if (i >= n4) {               // ragged tail: flush, then sort after every live lane
      if (cur != INT_MAX && acc != 0.f) atomicAdd(sumsq + cur, acc);
      cur = INT_MAX;
      acc = 0.f;
      break;
    }
    const long e = i * 4;
    const NovoTensor t = tab[novo_find(tab, nt, e)];
    const float4 v = *reinterpret_cast<const float4*>(p + e);
    const int g0 = novo_group(t, e), g3 = novo_group(t, e + 3);
    if (g0 == g3) {
      const float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (g0 != cur) {
        if (cur != INT_MAX && acc != 0.f) atomicAdd(sumsq + cur, acc);
        cur = g0;
        acc = 0.f;
      }
      acc += s;
    } else {                     // a float4 straddling units (odd fan, e.g. 7*7*3)
      // keep the LAST element's group in the lane (lanes stay sorted by group for the reduction
      // below), push the others straight to memory
      if (cur != INT_MAX && acc != 0.f) atomicAdd(sumsq + cur, acc);
      cur = g3;
      acc = 0.f;
      const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = novo_group(t, e + j);
        if (gj == g3) acc += a[j] * a[j];
        else atomicAdd(sumsq + gj, a[j] * a[j]);
      }
    }
  }
  // lanes hold non-decreasing groups after the last round: segmented shuffle reduction
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float ov = __shfl_down_sync(full, acc, off);
    const int og = __shfl_down_sync(full, cur, off);
    if (lane + off < 32 && og == cur) acc += ov;
  }
  const int pg = __shfl_up_sync(full, cur, 1);
  if ((lane == 0 || pg != cur) && cur != INT_MAX && acc != 0.f) atomicAdd(sumsq + cur, acc);
}

// one CTA per tensor: fold this step's norms into the running ones, publish the denominators
__global__ void __launch_bounds__(128)
novograd_norm_kernel(const NovoTensor* __restrict__ tab, float* __restrict__ sumsq,
                     float* __restrict__ ema_norm, float* __restrict__ denom, float eps,
                     int unitwise) {
  const NovoTensor t = tab[blockIdx.x];
  for (int g = threadIdx.x; g < t.ngroups; g += blockDim.x) {
    const int i = t.norm_base + g;
    const float s = sumsq[i];
    const float nrm = unitwise ? sqrtf(s) : s;
    const float en = __fmaf_rn(t.one_minus_beta2, nrm, __fmul_rn(ema_norm[i], t.beta2));
    ema_norm[i] = en;
    denom[i] = sqrtf(en) + eps;
    sumsq[i] = 0.f;              // ready for the next step
  }
}

__global__ void __launch_bounds__(256)
novograd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ema_grad,
                __nv_bfloat16* __restrict__ p_bf16, float* __restrict__ ema, float ema_decay,
                const NovoTensor* __restrict__ tab, int nt, const float* __restrict__ denom,
                long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    const long e = i * 4;
    const NovoTensor t = tab[novo_find(tab, nt, e)];
    const float4 pv = *reinterpret_cast<const float4*>(p + e);
    const float4 gv = *reinterpret_cast<const float4*>(g + e);
    const float4 mv = *reinterpret_cast<const float4*>(ema_grad + e);
    float pa[4] = {pv.x, pv.y, pv.z, pv.w};
    const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w};
    const int g0 = novo_group(t, e), g3 = novo_group(t, e + 3);
    const float d0 = denom[g0];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = (g0 == g3) ? d0 : denom[novo_group(t, e + j)];
      ma[j] = __fmaf_rn(t.one_minus_beta1, ga[j], __fmul_rn(ma[j], t.beta1));
      pa[j] = __fmaf_rn(-t.lr, __fdiv_rn(ma[j], d), pa[j]);
      pa[j] = __fmul_rn(pa[j], t.decay);
    }
    *reinterpret_cast<float4*>(p + e) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<float4*>(ema_grad + e) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    if (p_bf16 != nullptr) {
      uint2 o;
      o.x = pack2(pa[0], pa[1]);
      o.y = pack2(pa[2], pa[3]);
      *reinterpret_cast<uint2*>(p_bf16 + e) = o;
    }
    if (ema != nullptr) {
      float4 ev = *reinterpret_cast<float4*>(ema + e);
      ev.x = ema_decay * ev.x + (1.f - ema_decay) * pa[0];
      ev.y = ema_decay * ev.y + (1.f - ema_decay) * pa[1];
      ev.z = ema_decay * ev.z + (1.f - ema_decay) * pa[2];
      ev.w = ema_decay * ev.w + (1.f - ema_decay) * pa[3];
      *reinterpret_cast<float4*>(ema + e) = ev;
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    uint2 o;
    o.x = pack2(v.x, v.y);
    o.y = pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i * 4) = o;
  }
}

// one entry per conv filter
struct PackEntry {
  long src_off;    // element offset of the KRSC filter in the bf16 arena
  long dst_off;    // element offset of the [C][R][S][K] flipped copy in the dgrad arena
  int K, RS, C;
  int block_begin; // first block of this entry
};

// 32x32 tile transpose per (tap): dst[c][RS-1-tap][k] = src[k][tap][c]
__global__ void __launch_bounds__(256)
pack_dgrad_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                  const PackEntry* __restrict__ tab, int nent) {
  __shared__ __nv_bfloat16 tile[32][33];
  int lo = 0, hi = nent - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackEntry e = tab[lo];
  int b = blockIdx.x - e.block_begin;
  const int ct = (e.C + 31) / 32, kt = (e.K + 31) / 32;
  const int tap = b / (ct * kt);
  b -= tap * ct * kt;
  const int k0 = (b / ct) * 32, c0 = (b % ct) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const __nv_bfloat16* s = src + e.src_off;
  __nv_bfloat16* d = dst + e.dst_off;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, c = c0 + tx;
    if (k < e.K && c < e.C) tile[r][tx] = s[((long)k * e.RS + tap) * e.C + c];
  }
  __syncthreads();
  const int ftap = e.RS - 1 - tap;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, k = k0 + tx;
    if (k < e.K && c < e.C) d[((long)c * e.RS + ftap) * e.K + k] = tile[tx][r];
  }
}

// Row-parity sub-filters of a 3x3 / stride-2 dgrad (conv.cu dgrad_s2_impl) from the flipped
// pack wd[c][fr][fs][k] (= w[k][2-fr][2-fs][c]):
//   sub1[(b,c)][dp][dq][k] = b ? wd[c][2dp][2dq][k] : (dq == 0 ? wd[c][2dp][1][k] : 0)
//   sub0[(b,c)][0 ][dq][k] = b ? wd[c][1  ][2dq][k] : (dq == 0 ? wd[c][1  ][1][k] : 0)
// one block per (a, b, c, dp, dq) row of K elements.
__global__ void __launch_bounds__(128)
pack_dgrad_s2_kernel(const __nv_bfloat16* __restrict__ wd, __nv_bfloat16* __restrict__ sub0,
                     __nv_bfloat16* __restrict__ sub1, int C, int K) {
  int row = blockIdx.x;                      // [0, 2C*2) -> sub0 rows, then [.., + 2C*4) -> sub1 rows
  const int n0 = 2 * C * 2;
  const int a = row >= n0 ? 1 : 0;
  if (a) row -= n0;
  const int taps = a ? 4 : 2;
  const int bc = row / taps, tap = row - bc * taps;
  const int b = bc / C, c = bc - b * C;
  const int dp = a ? (tap >> 1) : 0, dq = tap & 1;
  const int fr = a ? 2 * dp : 1;
  const bool zero = (b == 0 && dq == 1);
  const int fs = b ? 2 * dq : 1;
  const __nv_bfloat16* src = wd + (((long)c * 3 + fr) * 3 + fs) * K;
  __nv_bfloat16* dst = (a ? sub1 : sub0) + (long)row * K;
  for (int k = threadIdx.x * 8; k < K; k += blockDim.x * 8) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!zero) v = *reinterpret_cast<const uint4*>(src + k);
    *reinterpret_cast<uint4*>(dst + k) = v;
  }
}

// one CTA per output channel: w_std = (w - mean) / sqrt(var_biased + eps) * gain
__global__ void __launch_bounds__(256)
weight_std_kernel(const float* __restrict__ w, const float* __restrict__ gain,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ mean_invstd, int fan,
                  float eps) {
  __shared__ float sh[2][32];
  const long o = blockIdx.x;
  float s = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < fan; i += blockDim.x) {
    const float v = w[o * fan + i];
    s += v;
    s2 += v * v;
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  s = 0.f; s2 = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) { s += sh[0][i]; s2 += sh[1][i]; }
  const float mean = s / fan;
  const float var = fmaxf(s2 / fan - mean * mean, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float gn = gain ? gain[o] : 1.f;
  if (threadIdx.x == 0 && mean_invstd) { mean_invstd[2 * o] = mean; mean_invstd[2 * o + 1] = invstd; }
  for (int i = threadIdx.x; i < fan; i += blockDim.x)
    out[o * fan + i] = __float2bfloat16_rn((w[o * fan + i] - mean) * invstd * gn);
}

// backward of weight standardisation: dw = invstd*gain * (g - mean(g) - what * mean(g*what))
__global__ void __launch_bounds__(256)
weight_std_bwd_kernel(const float* __restrict__ w, const float* __restrict__ gain,
                      const float* __restrict__ mean_invstd, const float* __restrict__ g,
                      float* __restrict__ dw, int fan) {
  __shared__ float sh[2][32];
  const long o = blockIdx.x;
  const float mean = mean_invstd[2 * o], invstd = mean_invstd[2 * o + 1];
  float s = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < fan; i += blockDim.x) {
    const float gv = g[o * fan + i];
    const float wh = (w[o * fan + i] - mean) * invstd;
    s += gv;
    s2 += gv * wh;
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  s = 0.f; s2 = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) { s += sh[0][i]; s2 += sh[1][i]; }
  const float gn = gain ? gain[o] : 1.f;
  for (int i = threadIdx.x; i < fan; i += blockDim.x) {
    const float wh = (w[o * fan + i] - mean) * invstd;
    dw[o * fan + i] = gn * invstd * (g[o * fan + i] - s / fan - wh * s2 / fan);
  }
}

// All weight-standardised filters of an arena in ONE launch (forward and backward): a table maps
// block ranges to tensors (one block per output channel, as in the per-tensor kernels above).
struct WsEntry {
  long off;          // element offset of the tensor inside the flat parameter / shadow / gradient arenas
  long mi_off;       // offset (in channels) into the flat mean/invstd buffer
  int K, fan;
  int block_base;    // first block of this tensor
  int pad;
};

__device__ __forceinline__ int ws_find(const WsEntry* __restrict__ tab, int n, int block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].block_base <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
weight_std_batch_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                        float* __restrict__ mean_invstd, const WsEntry* __restrict__ tab, int n,
                        float eps) {
  __shared__ float sh[2][32];
  const WsEntry e = tab[ws_find(tab, n, blockIdx.x)];
  const long o = blockIdx.x - e.block_base;
  const float* wp = w + e.off + o * e.fan;
  float s = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < e.fan; i += blockDim.x) {
    const float v = wp[i];
    s += v;
    s2 += v * v;
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  s = 0.f; s2 = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) { s += sh[0][i]; s2 += sh[1][i]; }
  const float mean = s / e.fan;
  const float var = fmaxf(s2 / e.fan - mean * mean, 0.f);
  const float invstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) { mean_invstd[2 * (e.mi_off + o)] = mean; mean_invstd[2 * (e.mi_off + o) + 1] = invstd; }
  __nv_bfloat16* op = out + e.off + o * e.fan;
  for (int i = threadIdx.x; i < e.fan; i += blockDim.x) op[i] = __float2bfloat16_rn((wp[i] - mean) * invstd);
}

__global__ void __launch_bounds__(256)
weight_std_bwd_batch_kernel(const float* __restrict__ w, const float* __restrict__ mean_invstd,
                            float* __restrict__ g, const WsEntry* __restrict__ tab, int n) {
  __shared__ float sh[2][32];
  const WsEntry e = tab[ws_find(tab, n, blockIdx.x)];
  const long o = blockIdx.x - e.block_base;
  const float mean = mean_invstd[2 * (e.mi_off + o)], invstd = mean_invstd[2 * (e.mi_off + o) + 1];
  const float* wp = w + e.off + o * e.fan;
  float* gp = g + e.off + o * e.fan;
  float s = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < e.fan; i += blockDim.x) {
    const float gv = gp[i];
    s += gv;
    s2 += gv * ((wp[i] - mean) * invstd);
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  s = 0.f; s2 = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) { s += sh[0][i]; s2 += sh[1][i]; }
  for (int i = threadIdx.x; i < e.fan; i += blockDim.x) {
    const float wh = (wp[i] - mean) * invstd;
    gp[i] = invstd * (gp[i] - s / e.fan - wh * s2 / e.fan);
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

// segs_dev: device array of nseg records {long end; float lr, wd, momentum, dampening; int nesterov, pad}
extern "C" int sib_sgd_step(float* params, const float* grads, float* momentum_buf,
                            void* params_bf16, float* ema, float ema_decay, const void* segs_dev,
                            int nseg, long n, int first_step, void* stream) {
  SIB_CHECK(n % 4 == 0, "sgd: arena length must be a multiple of 4 (pad the arena)");
  SIB_CHECK(nseg >= 1, "sgd: need at least one segment");
  const long n4 = n / 4;
  sgd_kernel<<<ew_grid(n4, 256), 256, 0, ST(stream)>>>(
      params, grads, momentum_buf, static_cast<__nv_bfloat16*>(params_bf16), ema, ema_decay,
      static_cast<const SgdSeg*>(segs_dev), nseg, n4, first_step);
  SIB_LAUNCH_CHECK();
  return 0;
}

// table_dev: ntensors x {long begin, end; int unit_len, ngroups, norm_base; float lr, decay,
// beta1, 1-beta1, beta2, 1-beta2; int pad}.  group_sumsq must be zero on the first call (the
// norm kernel re-zeroes it); group_ema_norm holds ema_norm_init before the first step.
extern "C" int sib_novograd_step(float* params, const float* grads, float* ema_grad,
                                 void* params_bf16, float* ema, float ema_decay,
                                 const void* table_dev, int ntensors, float* group_sumsq,
                                 float* group_ema_norm, float* group_denom, long n, float eps,
                                 int unitwise, void* stream) {
  SIB_CHECK(n % 4 == 0, "novograd: arena length must be a multiple of 4 (pad the arena)");
  SIB_CHECK(ntensors >= 1, "novograd: need at least one tensor record");
  const long n4 = n / 4;
  const NovoTensor* tab = static_cast<const NovoTensor*>(table_dev);
  const long per_block = 8L * 32 * NOVO_ROUNDS;      // float4s one 256-thread block walks
  novograd_sumsq_kernel<<<(unsigned)((n4 + per_block - 1) / per_block), 256, 0, ST(stream)>>>(
      params, tab, ntensors, n4, group_sumsq);
  SIB_LAUNCH_CHECK();
  novograd_norm_kernel<<<ntensors, 128, 0, ST(stream)>>>(tab, group_sumsq, group_ema_norm,
                                                         group_denom, eps, unitwise);
  SIB_LAUNCH_CHECK();
  novograd_kernel<<<ew_grid(n4, 256), 256, 0, ST(stream)>>>(
      params, grads, ema_grad, static_cast<__nv_bfloat16*>(params_bf16), ema, ema_decay, tab,
      ntensors, group_denom, n4);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_cast_bf16(const float* src, void* dst, long n, void* stream) {
  SIB_CHECK(n % 4 == 0, "cast: length must be a multiple of 4");
  cast_bf16_kernel<<<ew_grid(n / 4, 256), 256, 0, ST(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n / 4);
  SIB_LAUNCH_CHECK();
  return 0;
}

// table_dev: device array of nent records {long src_off, dst_off; int K, RS, C, block_begin}
extern "C" int sib_pack_dgrad_weights(const void* w_bf16, void* w_dgrad, const void* table_dev,
                                      int nent, int total_blocks, void* stream) {
  if (nent == 0 || total_blocks == 0) return 0;
  pack_dgrad_kernel<<<total_blocks, 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(w_bf16), static_cast<__nv_bfloat16*>(w_dgrad),
      static_cast<const PackEntry*>(table_dev), nent);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_pack_dgrad_s2(const void* w_dgrad, void* w_sub0, void* w_sub1, int C, int K,
                                 void* stream) {
  SIB_CHECK(K % 8 == 0, "pack_dgrad_s2: K must be a multiple of 8");
  pack_dgrad_s2_kernel<<<2 * C * 6, 128, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(w_dgrad), static_cast<__nv_bfloat16*>(w_sub0),
      static_cast<__nv_bfloat16*>(w_sub1), C, K);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_weight_standardize(const float* w, const float* gain, void* out_bf16,
                                      float* mean_invstd, int out_channels, int fan, float eps,
                                      void* stream) {
  weight_std_kernel<<<out_channels, 256, 0, ST(stream)>>>(
      w, gain, static_cast<__nv_bfloat16*>(out_bf16), mean_invstd, fan, eps);
  SIB_LAUNCH_CHECK();
  return 0;
}

// every standardised filter of an arena in one launch; table_dev: WsEntry[n] (see arena.py)
extern "C" int sib_weight_standardize_batch(const float* params, void* shadow_bf16, float* mean_invstd,
                                            const void* table_dev, int n, int total_blocks, float eps,
                                            void* stream) {
  if (n <= 0 || total_blocks <= 0) return 0;
  weight_std_batch_kernel<<<total_blocks, 256, 0, ST(stream)>>>(
      params, static_cast<__nv_bfloat16*>(shadow_bf16), mean_invstd, static_cast<const WsEntry*>(table_dev), n, eps);
  SIB_LAUNCH_CHECK();
  return 0;
}

// gradient w.r.t. the raw filters from the gradient w.r.t. the standardised ones, in place in the
// flat gradient arena
extern "C" int sib_weight_standardize_bwd_batch(const float* params, const float* mean_invstd, float* grads,
                                                const void* table_dev, int n, int total_blocks,
                                                void* stream) {
  if (n <= 0 || total_blocks <= 0) return 0;
  weight_std_bwd_batch_kernel<<<total_blocks, 256, 0, ST(stream)>>>(params, mean_invstd, grads,
                                                                    static_cast<const WsEntry*>(table_dev), n);
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_weight_standardize_bwd(const float* w, const float* gain,
                                          const float* mean_invstd, const float* g, float* dw,
                                          int out_channels, int fan, void* stream) {
  weight_std_bwd_kernel<<<out_channels, 256, 0, ST(stream)>>>(w, gain, mean_invstd, g, dw, fan);
  SIB_LAUNCH_CHECK();
  return 0;
}
