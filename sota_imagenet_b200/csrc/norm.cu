// Memory-bound NHWC bf16 kernels around the convolutions: BatchNorm statistics,
// finalize (+running stats), apply (+activation, +residual, +second BN), backward
// reduce / apply, 3x3/2 max-pool and global average pool.  All global traffic is
// 128-bit; per-channel sums use register accumulation -> smem -> fp32 atomics.
//
// Arithmetic follows nn.BatchNorm2d in training mode, which is what
// pytorch_tools' ABN / the reference's `patch_bn_mom` path run (reference train.py:76).
#include <stdlib.h>

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
// Thread layout shared by all BatchNorm kernels: a thread OWNS one 8-channel vector column
// (tx) and walks rows (ty + k * rows-in-flight), so every per-channel coefficient is loaded
// once into registers and all global traffic is 128-bit and fully coalesced (adjacent tx =
// adjacent 16 bytes; adjacent ty = adjacent rows = adjacent memory).  U rows are in flight per
// thread to cover HBM latency.  Requires C % 8 == 0 and C / 8 <= kRedThreads.
// ---------------------------------------------------------------------------
struct ColOwner {
  int tx, ty, rpb;
  long row0, stride;
  bool active;
  __device__ __forceinline__ ColOwner(int C) {
    const int cvec = C >> 3;
    rpb = kRedThreads / cvec;
    tx = threadIdx.x % cvec;
    ty = threadIdx.x / cvec;
    active = ty < rpb;
    row0 = (long)blockIdx.x * rpb + ty;
    stride = (long)gridDim.x * rpb;
  }
};

// SIB_DETERMINISTIC=1: the cross-block combination of the reduction kernels does not use
// floating-point atomics (whose arrival order changes the last bits from run to run); every block
// stores its partial sums to a scratch buffer and the LAST block to finish (ticket counter) adds
// them up in block order.  One scratch buffer per process: all reductions run on one stream.
struct DetScratch {
  float* partials;          // [max blocks][4][kRedThreads * 8]
  unsigned* counter;
};

// cross-thread reduction of per-thread column accumulators -> fp32 atomics on out[NACC][C]
template <int NACC>
__device__ __forceinline__ void column_reduce_store(const ColOwner& co, int C,
                                                    float (&acc)[NACC][8], float* __restrict__ out,
                                                    const DetScratch det = DetScratch{nullptr, nullptr}) {
  __shared__ float red[NACC][kRedThreads][8];
  __shared__ unsigned s_ticket;
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[a][threadIdx.x][j] = co.active ? acc[a][j] : 0.f;
  __syncthreads();
  const int cvec = C >> 3;
  // one 4-channel vector atomic per thread and accumulator (C % 8 == 0): the grids of the
  // reduction kernels are capped at two blocks per SM (bn_reduce_grid) because same-address
  // atomics from ~1200 blocks finishing together cost ~50 us at the end of the kernel
  for (int idx = threadIdx.x; idx < NACC * (C >> 2); idx += kRedThreads) {
    const int a = idx / (C >> 2);
    const int c = (idx - a * (C >> 2)) * 4;
    const int v = c >> 3, j = c & 7;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int y = 0; y < co.rpb; ++y) {
      const float* r4 = &red[a][y * cvec + v][j];
      s0 += r4[0]; s1 += r4[1]; s2 += r4[2]; s3 += r4[3];
    }
    if (det.partials != nullptr) {
      *reinterpret_cast<float4*>(det.partials + ((long)blockIdx.x * NACC + a) * C + c) = make_float4(s0, s1, s2, s3);
    } else {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (long)a * C + c),
                   "f"(s0), "f"(s1), "f"(s2), "f"(s3)
                   : "memory");
    }
  }
  if (det.partials != nullptr) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(det.counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {           // last block: fixed-order sum over the blocks
      __threadfence();
      for (int idx = threadIdx.x; idx < NACC * C; idx += kRedThreads) {
        float t = 0.f;
        for (unsigned b = 0; b < gridDim.x; ++b)
          t += __ldcg(det.partials + (long)b * NACC * C + idx);
        out[idx] = t;
      }
      if (threadIdx.x == 0) *det.counter = 0u;  // ready for the next launch (same stream)
    }
  }
}

constexpr int kU = 4;   // rows in flight per thread

__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long M, int C, float* __restrict__ stats,
                const DetScratch det) {
  griddep_launch();
  griddep_wait();
  const ColOwner co(C);
  float acc[2][8] = {};
  if (co.active) {
    const __nv_bfloat16* px = x + co.tx * 8;
    for (long r = co.row0; r < M; r += kU * co.stride) {
      uint4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (r + u * co.stride < M) v[u] = ldg_stream(px + (r + u * co.stride) * C);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (r + u * co.stride < M) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acc[1][j] += f[j] * f[j]; }
        }
      }
    }
  }
  column_reduce_store<2>(co, C, acc, stats, det);
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
  const BnCoeffs k = bn_coeffs(stats[c], stats[C + c], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f,
                               count, eps);
  mean_invstd[c] = k.mean;
  mean_invstd[C + c] = k.invstd;
  scale_shift[c] = k.scale;
  scale_shift[C + c] = k.shift;
  if (running_mean) {
    running_mean[c] = bn_running(running_mean[c], k.mean, momentum);
    running_var[c] = bn_running(running_var[c], bn_unbiased(k.var, count), momentum);
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
template <int MODE>   // 0: plain, 1: + res, 2: + res*scale2 + shift2
__global__ void __launch_bounds__(kRedThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ ss,
                const __nv_bfloat16* __restrict__ res, const float* __restrict__ ss2,
                __nv_bfloat16* __restrict__ y, long M, int C, int act, float slope) {
  griddep_launch();
  griddep_wait();
  const ColOwner co(C);
  if (!co.active) return;
  float sc[8], sh[8], sc2[8], sh2[8];
  load8f(ss + co.tx * 8, sc);
  load8f(ss + C + co.tx * 8, sh);
  if (MODE == 2) {
    load8f(ss2 + co.tx * 8, sc2);
    load8f(ss2 + C + co.tx * 8, sh2);
  }
  const long col = co.tx * 8;
  for (long r = co.row0; r < M; r += kU * co.stride) {
    uint4 vx[kU], vr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        vx[u] = ldg_stream(x + rr * C + col);
        if (MODE != 0) vr[u] = ldg_stream(res + rr * C + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        float f[8], o[8];
        unpack8(vx[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], sc[j], sh[j]);
        if (MODE != 0) {
          float q[8];
          unpack8(vr[u], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += (MODE == 2) ? fmaf(q[j], sc2[j], sh2[j]) : q[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = act_fwd(o[j], act, slope);
        stg_stream(y + rr * C + col, pack8(o));
      }
    }
  }
}

// ---------------------------------------------------------------------------
// finalize + apply in one launch: every thread derives scale/shift of its 8 channels from the
// raw statistics; the first row-lane of block 0 also publishes mean/invstd, scale/shift (needed
// by the backward pass) and updates the running statistics.  Saves one tiny launch per BN.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bn_derive(const float* __restrict__ stats,
                                          const float* __restrict__ gamma,
                                          const float* __restrict__ beta,
                                          float* __restrict__ running_mean,
                                          float* __restrict__ running_var,
                                          float* __restrict__ mean_invstd,
                                          float* __restrict__ scale_shift, int C, int c0,
                                          float count, float eps, float momentum, bool publish,
                                          float* sc, float* sh) {
  float s1[8], s2[8], g[8], b[8];
  load8f(stats + c0, s1);
  load8f(stats + C + c0, s2);
  if (gamma) load8f(gamma + c0, g);
  if (beta) load8f(beta + c0, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const BnCoeffs k = bn_coeffs(s1[j], s2[j], gamma ? g[j] : 1.f, beta ? b[j] : 0.f, count, eps);
    sc[j] = k.scale;
    sh[j] = k.shift;
    if (publish) {
      const int c = c0 + j;
      mean_invstd[c] = k.mean;
      mean_invstd[C + c] = k.invstd;
      scale_shift[c] = sc[j];
      scale_shift[C + c] = sh[j];
      if (running_mean) {
        running_mean[c] = bn_running(running_mean[c], k.mean, momentum);
        running_var[c] = bn_running(running_var[c], bn_unbiased(k.var, count), momentum);
      }
    }
  }
}

struct BnFinalizeArgs {
  const float* stats; const float* gamma; const float* beta;
  float* running_mean; float* running_var; float* mean_invstd; float* scale_shift;
};

template <int MODE>   // 0: plain, 1: + res, 2: + BN2(res)
__global__ void __launch_bounds__(kRedThreads)
bn_finalize_apply_kernel(const __nv_bfloat16* __restrict__ x, BnFinalizeArgs f1,
                         const __nv_bfloat16* __restrict__ res, BnFinalizeArgs f2,
                         __nv_bfloat16* __restrict__ y, long M, int C, float count, float eps,
                         float momentum, int act, float slope) {
  griddep_launch();
  griddep_wait();
  const ColOwner co(C);
  if (!co.active) return;
  const bool publish = blockIdx.x == 0 && co.ty == 0;
  float sc[8], sh[8], sc2[8], sh2[8];
  bn_derive(f1.stats, f1.gamma, f1.beta, f1.running_mean, f1.running_var, f1.mean_invstd,
            f1.scale_shift, C, co.tx * 8, count, eps, momentum, publish, sc, sh);
  if (MODE == 2)
    bn_derive(f2.stats, f2.gamma, f2.beta, f2.running_mean, f2.running_var, f2.mean_invstd,
              f2.scale_shift, C, co.tx * 8, count, eps, momentum, publish, sc2, sh2);
  const long col = co.tx * 8;
  for (long r = co.row0; r < M; r += kU * co.stride) {
    uint4 vx[kU], vr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        vx[u] = ldg_stream(x + rr * C + col);
        if (MODE != 0) vr[u] = ldg_stream(res + rr * C + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        float f[8], o[8];
        unpack8(vx[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], sc[j], sh[j]);
        if (MODE != 0) {
          float q[8];
          unpack8(vr[u], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += (MODE == 2) ? fmaf(q[j], sc2[j], sh2[j]) : q[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = act_fwd(o[j], act, slope);
        stg_stream(y + rr * C + col, pack8(o));
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward.  g = dy * act'(.) where the activation mask comes either from the stored output
// (`out`, needed when a residual was added) or is recomputed from x with the forward's own
// scale/shift (`mask_ss`): fmaf(x, scale, shift) > 0 is bit-identical to what the forward
// evaluated, so no mask tensor and no read of `out` is needed for plain BN+act.
//   reduce: sums[0] = sum g, sums[1] = sum g*xhat (+ sums[2..3] for a second BN sharing g)
//   apply : dx = A*g + B*x + Cc with per-channel A = gamma*invstd, B = -gamma*invstd^2*sgx/m,
//           Cc = -gamma*invstd*sg/m + gamma*invstd^2*mean*sgx/m
// ---------------------------------------------------------------------------
template <bool SECOND, bool USE_OUT>
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ out,
                     const float* __restrict__ mask_ss, const __nv_bfloat16* __restrict__ x,
                     const float* __restrict__ mi, const __nv_bfloat16* __restrict__ x2,
                     const float* __restrict__ mi2, long M, int C, int act, float slope,
                     float* __restrict__ sums, const DetScratch det) {
  griddep_launch();
  griddep_wait();
  const ColOwner co(C);
  constexpr int NACC = SECOND ? 4 : 2;
  constexpr int U = kU;
  float acc[NACC][8] = {};
  if (co.active) {
    float mean[8], invstd[8], mean2[8], invstd2[8], sc[8], sh[8];
    load8f(mi + co.tx * 8, mean);
    load8f(mi + C + co.tx * 8, invstd);
    if (SECOND) {
      load8f(mi2 + co.tx * 8, mean2);
      load8f(mi2 + C + co.tx * 8, invstd2);
    }
    const bool remask = !USE_OUT && act != SIB_ACT_NONE;
    if (remask) {
      load8f(mask_ss + co.tx * 8, sc);
      load8f(mask_ss + C + co.tx * 8, sh);
    }
    const long col = co.tx * 8;
    for (long r = co.row0; r < M; r += U * co.stride) {
      uint4 vg[U], vx[U], vo[U], vw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long rr = r + u * co.stride;
        if (rr < M) {
          vg[u] = ldg_stream(dy + rr * C + col);
          vx[u] = ldg_stream(x + rr * C + col);
          if (USE_OUT) vo[u] = ldg_stream(out + rr * C + col);
          if (SECOND) vw[u] = ldg_stream(x2 + rr * C + col);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long rr = r + u * co.stride;
        if (rr < M) {
          float g[8], xv[8];
          unpack8(vg[u], g);
          unpack8(vx[u], xv);
          if (USE_OUT) {
            float o[8];
            unpack8(vo[u], o);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
          } else if (remask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= act_grad(fmaf(xv[j], sc[j], sh[j]), act, slope);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0][j] += g[j];
            acc[1][j] += g[j] * ((xv[j] - mean[j]) * invstd[j]);
          }
          if (SECOND) {
            float xw[8];
            unpack8(vw[u], xw);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[2][j] += g[j];
              acc[3][j] += g[j] * ((xw[j] - mean2[j]) * invstd2[j]);
            }
          }
        }
      }
    }
  }
  column_reduce_store<NACC>(co, C, acc, sums, det);
}

//   EMIT_A: also rewrite the forward activation a = fwd_act(fmaf(x, scale, shift)) of this BatchNorm
//   (x is in registers anyway): the fused conv prologue never stored it and the weight gradient
//   of the consumer conv needs it.  +2 B/element written, no extra read.
//   DYT: the incoming gradient is dy * dy_mul[n][c] + dy_add[n][c] (n = row / HW): the ECA gate
//   (* drop-connect keep) and the gate's pooled-path gradient of the BResNet block tail, folded
//   into this pass instead of a separate scale pass over the widest tensor of the block.
template <bool SECOND, bool USE_OUT, bool WRITE_G, bool EMIT_A = false, bool DYT = false>
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ out,
                    const float* __restrict__ mask_ss, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ mi, const float* __restrict__ gamma,
                    const float* __restrict__ sums, const __nv_bfloat16* __restrict__ x2,
                    const float* __restrict__ mi2, const float* __restrict__ gamma2,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx2,
                    __nv_bfloat16* __restrict__ gout, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ dgamma2,
                    float* __restrict__ dbeta2, long M, int C, float inv_count, int act,
                    float slope, float pg_scale, const float* __restrict__ emit_ss = nullptr,
                    int emit_act = 0, float emit_slope = 0.f,
                    __nv_bfloat16* __restrict__ a_out = nullptr,
                    const float* __restrict__ dy_mul = nullptr,
                    const float* __restrict__ dy_add = nullptr, int HW = 1) {
  griddep_launch();
  griddep_wait();
  const ColOwner co(C);
  if (!co.active) return;
  if (blockIdx.x == 0 && co.ty == 0) {
    // parameter gradients ride along: dbeta += sum g, dgamma += sum g*xhat (one writer per channel).
    // pg_scale = 1 / world under SyncBN: `sums` are then totals over ALL ranks, while the
    // data-parallel wrapper averages parameter gradients over the ranks afterwards.
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = co.tx * 8 + j;
      if (dbeta) dbeta[c] += sums[c] * pg_scale;
      if (dgamma) dgamma[c] += sums[C + c] * pg_scale;
      if (SECOND) {
        if (dbeta2) dbeta2[c] += sums[2 * C + c] * pg_scale;
        if (dgamma2) dgamma2[c] += sums[3 * C + c] * pg_scale;
      }
    }
  }
  float cA[8], cB[8], cC[8], dA[8], dB[8], dC[8], sc[8], sh[8];
  {
    float mean[8], invstd[8], gm[8], sg[8], sgx[8];
    load8f(mi + co.tx * 8, mean);
    load8f(mi + C + co.tx * 8, invstd);
    load8f(sums + co.tx * 8, sg);
    load8f(sums + C + co.tx * 8, sgx);
    if (gamma) load8f(gamma + co.tx * 8, gm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = (gamma ? gm[j] : 1.f) * invstd[j];
      cA[j] = a;
      cB[j] = -a * invstd[j] * sgx[j] * inv_count;
      cC[j] = -a * sg[j] * inv_count - cB[j] * mean[j];
    }
    if (SECOND) {
      load8f(mi2 + co.tx * 8, mean);
      load8f(mi2 + C + co.tx * 8, invstd);
      load8f(sums + 2 * C + co.tx * 8, sg);
      load8f(sums + 3 * C + co.tx * 8, sgx);
      if (gamma2) load8f(gamma2 + co.tx * 8, gm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = (gamma2 ? gm[j] : 1.f) * invstd[j];
        dA[j] = a;
        dB[j] = -a * invstd[j] * sgx[j] * inv_count;
        dC[j] = -a * sg[j] * inv_count - dB[j] * mean[j];
      }
    }
  }
  const bool remask = !USE_OUT && act != SIB_ACT_NONE;
  if (remask) {
    load8f(mask_ss + co.tx * 8, sc);
    load8f(mask_ss + C + co.tx * 8, sh);
  }
  float esc[8], esh[8];
  if (EMIT_A) {
    load8f(emit_ss + co.tx * 8, esc);
    load8f(emit_ss + C + co.tx * 8, esh);
  }
  const long col = co.tx * 8;
  for (long r = co.row0; r < M; r += kU * co.stride) {
    uint4 vg[kU], vx[kU], vo[kU], vw[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        vg[u] = ldg_stream(dy + rr * C + col);
        vx[u] = ldg_stream(x + rr * C + col);
        if (USE_OUT) vo[u] = ldg_stream(out + rr * C + col);
        if (SECOND) vw[u] = ldg_stream(x2 + rr * C + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long rr = r + u * co.stride;
      if (rr < M) {
        float g[8], xv[8], d[8];
        unpack8(vg[u], g);
        unpack8(vx[u], xv);
        if (DYT) {
          const long nn = (long)((unsigned)rr / (unsigned)HW);     // (M < 2^31: 32-bit division)
          float m[8], ad[8];
          load8f(dy_mul + nn * C + col, m);
          load8f(dy_add + nn * C + col, ad);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = fmaf(g[j], m[j], ad[j]);
        }
        if (USE_OUT) {
          float o[8];
          unpack8(vo[u], o);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
        } else if (remask) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] *= act_grad(fmaf(xv[j], sc[j], sh[j]), act, slope);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fmaf(cA[j], g[j], fmaf(cB[j], xv[j], cC[j]));
        stg_stream(dx + rr * C + col, pack8(d));
        if (SECOND) {
          float xw[8];
          unpack8(vw[u], xw);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = fmaf(dA[j], g[j], fmaf(dB[j], xw[j], dC[j]));
          stg_stream(dx2 + rr * C + col, pack8(d));
        }
        if (WRITE_G) stg_stream(gout + rr * C + col, pack8(g));
        if (EMIT_A) {
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = act_fwd(fmaf(xv[j], esc[j], esh[j]), emit_act, emit_slope);
          stg_stream(a_out + rr * C + col, pack8(a));
        }
      }
    }
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
  griddep_launch();
  griddep_wait();
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

// One block per PAIR of input rows (n, 2a / 2a+1), one thread per 2x2 input quad and 8-channel vector.
// The quad (2a + dy, 2b + dx) is fed by exactly the four windows (a | a+1, b | b+1): window (a, b) through
// its taps r, s in {1, 2}, (a, b+1) through s = 0, (a+1, b) through r = 0, (a+1, b+1) through (0, 0) -- nine
// (window, tap) pairs per channel as before, but four (index, gradient) loads per quad instead of nine and one
// set of index arithmetic for four outputs (32-bit only: the 64-bit divisions of a flat grid-stride loop once
// made this kernel instruction-bound, 340 us against an 87 us HBM time; per-pixel gathers left it at 255 us).
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int OH, int OW,
                   int cshift) {
  griddep_launch();
  griddep_wait();
  const int cvec = 1 << cshift;
  const int HP = (H + 1) >> 1, WP = (W + 1) >> 1;
  for (int row = blockIdx.x; row < N * HP; row += gridDim.x) {
    const int n = row / HP;
    const int a = row - n * HP;
    const bool pok[2] = {a < OH, a + 1 < OH};
    const size_t prow[2] = {((size_t)n * OH + a) * OW, ((size_t)n * OH + a + 1) * OW};
    const bool hok1 = 2 * a + 1 < H;
    __nv_bfloat16* drow0 = dx + ((size_t)n * H + 2 * a) * W * C;
    __nv_bfloat16* drow1 = drow0 + (size_t)W * C;
    for (int item = threadIdx.x; item < (WP << cshift); item += blockDim.x) {
      const int b = item >> cshift;
      const int v = item & (cvec - 1);
      const bool qok[2] = {b < OW, b + 1 < OW};
      uint2 pk[4];
      uint4 gv[4];
#pragma unroll
      for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
        for (int qi = 0; qi < 2; ++qi) {
          pk[pi * 2 + qi] = make_uint2(0xffffffffu, 0xffffffffu);      // no tap matches 0xff
          gv[pi * 2 + qi] = make_uint4(0, 0, 0, 0);
          if (pok[pi] && qok[qi]) {
            const size_t o = (((prow[pi] + b + qi) << cshift) + v) * 8;
            pk[pi * 2 + qi] = __ldg(reinterpret_cast<const uint2*>(idx + o));
            gv[pi * 2 + qi] = __ldg(reinterpret_cast<const uint4*>(dy + o));
          }
        }
      }
      float acc[4][8];                      // [dy * 2 + dx][channel]
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll
      for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
        for (int qi = 0; qi < 2; ++qi) {
          const int k = pi * 2 + qi;
          float g[8];
          unpack8(gv[k], g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t t = ((j < 4 ? (pk[k].x >> (8 * j)) : (pk[k].y >> (8 * (j - 4)))) & 0xffu);
            // pixel (2a + d, 2b + e) <- window (a + pi, b + qi) tap (r, s) = (d + 1 - 2 pi, e + 1 - 2 qi)
#pragma unroll
            for (int d = pi; d < 2; ++d) {          // pi = 1 only reaches d = 1 (r = 0)
#pragma unroll
              for (int e = qi; e < 2; ++e) {
                const uint32_t tap = (uint32_t)((d + 1 - 2 * pi) * 3 + (e + 1 - 2 * qi));
                acc[d * 2 + e][j] += t == tap ? g[j] : 0.f;
              }
            }
          }
        }
      }
      const bool wok1 = 2 * b + 1 < W;
      const size_t o0 = ((size_t)(2 * b) << cshift) * 8 + (size_t)v * 8;
      stg_stream(drow0 + o0, pack8(acc[0]));
      if (wok1) stg_stream(drow0 + o0 + C, pack8(acc[1]));
      if (hok1) {
        stg_stream(drow1 + o0, pack8(acc[2]));
        if (wok1) stg_stream(drow1 + o0 + C, pack8(acc[3]));
      }
    }
  }
}

// BatchNorm finalize + apply + activation + 3x3/s2 max pool in one pass (the stem): the
// normalised activation is never written.  Values are rounded to bf16 before the comparison, so
// y and idx are bit-identical to bn_finalize_apply followed by maxpool_fwd.
__global__ void __launch_bounds__(256)
bn_act_maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, BnFinalizeArgs f,
                          __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int N, int H,
                          int W, int C, int OH, int OW, int cshift, float count, float eps,
                          float momentum, int act, float slope) {
  griddep_launch();
  griddep_wait();
  const int cvec = 1 << cshift;
  // blockDim is a multiple of cvec: a thread keeps its 8 channels across items
  const int v = threadIdx.x & (cvec - 1);
  float sc[8], sh[8];
  bn_derive(f.stats, f.gamma, f.beta, f.running_mean, f.running_var, f.mean_invstd, f.scale_shift,
            C, v * 8, count, eps, momentum, blockIdx.x == 0 && threadIdx.x < cvec, sc, sh);
  // one block per output row (n, p): 32-bit index math only
  for (int row = blockIdx.x; row < N * OH; row += gridDim.x) {
    const int n = row / OH;
    const int p = row - n * OH;
    const __nv_bfloat16* xin[3];
    bool hok[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = p * 2 - 1 + r;
      hok[r] = h >= 0 && h < H;
      xin[r] = x + ((size_t)n * H + (hok[r] ? h : 0)) * W * C + v * 8;
    }
    for (int item = threadIdx.x; item < (OW << cshift); item += blockDim.x) {
      const int q = item >> cshift;
      uint4 raw[9];
      bool ok[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int w = q * 2 - 1 + s;
          ok[r * 3 + s] = hok[r] && w >= 0 && w < W;
          if (ok[r * 3 + s]) raw[r * 3 + s] = __ldg(reinterpret_cast<const uint4*>(xin[r] + (size_t)w * C));
        }
      }
      const size_t o8p = ((size_t)row * OW << cshift) * 8 + (size_t)item * 8;
      if (act == SIB_ACT_RELU) {
        // ReLU (every ResNet stem): the whole comparison runs on PACKED bf16 pairs.  fma.f32x2 is the same
        // round-to-nearest fused multiply-add as fmaf, and rounding commutes with the clamp at zero, so
        // max(round(fma), 0) are bit for bit the values of the generic path; the running maximum is __hmax2,
        // the winning tap a masked select driven by __hgt2_mask (strict >: the first maximum keeps the tap).
        // 4.5 instructions per element and tap instead of 8 (the kernel is instruction-bound).
        __nv_bfloat162 best2[4];
        uint32_t bi2[4];
        const __nv_bfloat162 ninf = __halves2bfloat162(__ushort_as_bfloat16((unsigned short)0xFF80), __ushort_as_bfloat16((unsigned short)0xFF80));
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { best2[j] = ninf; bi2[j] = 0; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          if (!ok[k]) continue;
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw[k]);
          const uint32_t kk = (uint32_t)k * 0x00010001u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x2 = make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xffff0000u));
            const float2 z = __ffma2_rn(x2, make_float2(sc[2 * j], sc[2 * j + 1]), make_float2(sh[2 * j], sh[2 * j + 1]));
            const __nv_bfloat162 o2 = __hmax2(__floats2bfloat162_rn(z.x, z.y), zero2);
            const uint32_t gt = __hgt2_mask(o2, best2[j]);
            best2[j] = __hmax2(best2[j], o2);
            bi2[j] = (bi2[j] & ~gt) | (kk & gt);
          }
        }
        uint4 yo;
        yo.x = *reinterpret_cast<uint32_t*>(&best2[0]);
        yo.y = *reinterpret_cast<uint32_t*>(&best2[1]);
        yo.z = *reinterpret_cast<uint32_t*>(&best2[2]);
        yo.w = *reinterpret_cast<uint32_t*>(&best2[3]);
        stg_stream(y + o8p, yo);
        if (idx != nullptr) {
          uint2 pk;
          pk.x = (bi2[0] & 0xffu) | ((bi2[0] >> 16) << 8) | ((bi2[1] & 0xffu) << 16) | ((bi2[1] >> 16) << 24);
          pk.y = (bi2[2] & 0xffu) | ((bi2[2] >> 16) << 8) | ((bi2[3] & 0xffu) << 16) | ((bi2[3] >> 16) << 24);
          *reinterpret_cast<uint2*>(idx + o8p) = pk;
        }
        continue;
      }
      float best[8];
      int bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (!ok[k]) continue;
        float fv[8];
        unpack8(raw[k], fv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float o = __bfloat162float(__float2bfloat16(act_fwd(fmaf(fv[j], sc[j], sh[j]), act, slope)));
          if (o > best[j]) { best[j] = o; bi[j] = k; }
        }
      }
      const size_t o8 = ((size_t)row * OW << cshift) * 8 + (size_t)item * 8;
      stg_stream(y + o8, pack8(best));
      if (idx != nullptr) {
        uint2 pk;
        pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        *reinterpret_cast<uint2*>(idx + o8) = pk;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// global average pool [N][HW][C] -> [N][C] and its backward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int HW,
               int C) {
  griddep_launch();
  griddep_wait();
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
  griddep_launch();
  griddep_wait();
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
  griddep_launch();
  griddep_wait();
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

// dst[n][p*stride][q*stride][:] += src[n][p][q][:]   (strided 1x1 dgrad)
__global__ void __launch_bounds__(256)
scatter_add_strided_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                           int N, int OH, int OW, int C, int H, int W, int stride) {
  griddep_launch();
  griddep_wait();
  const int cvec = C >> 3;
  const long total = (long)N * OH * OW * cvec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    long t = i / cvec;
    const int q = (int)(t % OW); t /= OW;
    const int pp = (int)(t % OH);
    const int n = (int)(t / OH);
    float a[8], b[8];
    unpack8(ldg_stream(src + i * 8), a);
    __nv_bfloat16* d = dst + ((((long)n * H + (long)pp * stride) * W + (long)q * stride) * cvec + v) * 8;
    unpack8(*reinterpret_cast<const uint4*>(d), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    *reinterpret_cast<uint4*>(d) = pack8(a);
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

static int bn_grid(long M, int C) {
  const int rpb = kRedThreads / (C / 8);
  long blocks = (M + (long)rpb * kU - 1) / ((long)rpb * kU);
  const long cap = (long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// SIB_DETERMINISTIC=1 (read once): scratch for the fixed-order cross-block sums
static DetScratch det_scratch() {
  static DetScratch d{nullptr, nullptr};
  static bool init = false;
  if (!init) {
    init = true;
    const char* e = getenv("SIB_DETERMINISTIC");
    if (e != nullptr && e[0] == '1') {
      void* p = nullptr;
      const size_t bytes = (size_t)2 * sm_count() * 4 * kRedThreads * 8 * sizeof(float) + 256;
      if (cudaMalloc(&p, bytes) == cudaSuccess && cudaMemset(p, 0, bytes) == cudaSuccess) {
        d.counter = static_cast<unsigned*>(p);
        d.partials = reinterpret_cast<float*>(static_cast<char*>(p) + 256);
      }
    }
  }
  return d;
}

// statistics / backward-sum kernels: every block ends in 2C..4C atomics on the same addresses
static int bn_reduce_grid(long M, int C) {
  const int g = bn_grid(M, C);
  const int cap = sm_count() * 2;
  return g < cap ? g : cap;
}

extern "C" int sib_bn_stats(const void* x, long M, int C, float* stats, void* stream) {
  if (int rc = check_c(C)) return rc;
  SIB_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * C, ST(stream)));
  SIB_CUDA(launch_pdl(bn_stats_kernel, dim3(bn_reduce_grid(M, C)), dim3(kRedThreads), 0, ST(stream), static_cast<const __nv_bfloat16*>(x), M, C, stats, det_scratch()));
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
  if (int rc = check_c(C)) return rc;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(res);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  const int grid = bn_grid(M, C);
  if (res == nullptr)
    SIB_CUDA(launch_pdl(bn_apply_kernel<0>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, scale_shift, rp, scale_shift2, yp,
                                                            M, C, act, slope));
  else if (scale_shift2 == nullptr)
    SIB_CUDA(launch_pdl(bn_apply_kernel<1>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, scale_shift, rp, scale_shift2, yp,
                                                            M, C, act, slope));
  else
    SIB_CUDA(launch_pdl(bn_apply_kernel<2>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, scale_shift, rp, scale_shift2, yp,
                                                            M, C, act, slope));
  SIB_LAUNCH_CHECK();
  return 0;
}

// y = act(BN1(x) [+ res | + BN2(res)]) with both finalizations fused (training mode).
extern "C" int sib_bn_finalize_apply(const void* x, const float* stats, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var,
                                     float* mean_invstd, float* scale_shift, const void* res,
                                     const float* stats2, const float* gamma2, const float* beta2,
                                     float* running_mean2, float* running_var2,
                                     float* mean_invstd2, float* scale_shift2, void* y, long M,
                                     int C, double count, float eps, float momentum, int act,
                                     float slope, void* stream) {
  if (int rc = check_c(C)) return rc;
  const BnFinalizeArgs f1{stats, gamma, beta, running_mean, running_var, mean_invstd, scale_shift};
  const BnFinalizeArgs f2{stats2, gamma2, beta2, running_mean2, running_var2, mean_invstd2,
                          scale_shift2};
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(res);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  const int grid = bn_grid(M, C);
  if (res == nullptr)
    SIB_CUDA(launch_pdl(bn_finalize_apply_kernel<0>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, f1, rp, f2, yp, M, C, (float)count, eps, momentum, act, slope));
  else if (stats2 == nullptr)
    SIB_CUDA(launch_pdl(bn_finalize_apply_kernel<1>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, f1, rp, f2, yp, M, C, (float)count, eps, momentum, act, slope));
  else
    SIB_CUDA(launch_pdl(bn_finalize_apply_kernel<2>, dim3(grid), dim3(kRedThreads), 0, ST(stream), xp, f1, rp, f2, yp, M, C, (float)count, eps, momentum, act, slope));
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_bwd_reduce(const void* dy, const void* out, const float* mask_ss,
                                 const void* x, const float* mean_invstd, const void* x2,
                                 const float* mean_invstd2, long M, int C, int act, float slope,
                                 float* sums, void* stream) {
  if (int rc = check_c(C)) return rc;
  const bool prezeroed = (act & SIB_ACT_FLAG_PREZEROED) != 0;
  act &= ~SIB_ACT_FLAG_PREZEROED;
  SIB_CHECK(act == SIB_ACT_NONE || out != nullptr || mask_ss != nullptr,
            "bn_bwd_reduce: activation mask needs `out` or `mask_ss`");
  const int nacc = x2 ? 4 : 2;
  if (!prezeroed) SIB_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * nacc * C, ST(stream)));
  const int grid = bn_reduce_grid(M, C);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* o = static_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* xq = static_cast<const __nv_bfloat16*>(x2);
#define SIB_RED(S, O)                                                                         \
  SIB_CUDA(launch_pdl(bn_bwd_reduce_kernel<S, O>, dim3(grid), dim3(kRedThreads), 0, ST(stream), a, o, mask_ss, xp, mean_invstd, \
                                                                 xq, mean_invstd2, M, C, act,  \
                                                                 slope, sums, det_scratch()))
  const bool use_out = out != nullptr && act != SIB_ACT_NONE;
  if (x2) { if (use_out) SIB_RED(true, true); else SIB_RED(true, false); }
  else    { if (use_out) SIB_RED(false, true); else SIB_RED(false, false); }
#undef SIB_RED
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_bn_bwd_apply(const void* dy, const void* out, const float* mask_ss,
                                const void* x, const float* mean_invstd, const float* gamma,
                                const float* sums, const void* x2, const float* mean_invstd2,
                                const float* gamma2, void* dx, void* dx2, void* gout,
                                float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, long M,
                                int C, double count, int act, float slope, float pgrad_scale,
                                void* stream) {
  if (int rc = check_c(C)) return rc;
  SIB_CHECK(act == SIB_ACT_NONE || out != nullptr || mask_ss != nullptr,
            "bn_bwd_apply: activation mask needs `out` or `mask_ss`");
  const int grid = bn_grid(M, C);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* o = static_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* xq = static_cast<const __nv_bfloat16*>(x2);
  __nv_bfloat16* d1 = static_cast<__nv_bfloat16*>(dx);
  __nv_bfloat16* d2 = static_cast<__nv_bfloat16*>(dx2);
  __nv_bfloat16* gg = static_cast<__nv_bfloat16*>(gout);
  const float ic = (float)(1.0 / count);
#define SIB_APP(S, O, G)                                                                       \
  SIB_CUDA(launch_pdl(bn_bwd_apply_kernel<S, O, G>, dim3(grid), dim3(kRedThreads), 0, ST(stream), \
      a, o, mask_ss, xp, mean_invstd, gamma, sums, xq, mean_invstd2, gamma2, d1, d2, gg, dgamma,   \
      dbeta, dgamma2, dbeta2, M, C, ic, act, slope, pgrad_scale, static_cast<const float*>(nullptr), 0, 0.f, \
      static_cast<__nv_bfloat16*>(nullptr), static_cast<const float*>(nullptr),                     \
      static_cast<const float*>(nullptr), 1))
  const bool use_out = out != nullptr && act != SIB_ACT_NONE;
  const bool wg = gout != nullptr;
  if (x2) {
    if (use_out) { if (wg) SIB_APP(true, true, true); else SIB_APP(true, true, false); }
    else         { if (wg) SIB_APP(true, false, true); else SIB_APP(true, false, false); }
  } else {
    if (use_out) { if (wg) SIB_APP(false, true, true); else SIB_APP(false, true, false); }
    else         { if (wg) SIB_APP(false, false, true); else SIB_APP(false, false, false); }
  }
#undef SIB_APP
  SIB_LAUNCH_CHECK();
  return 0;
}

// bn_bwd_apply of an activation-free BatchNorm whose incoming gradient is dy * dy_mul[n][c] +
// dy_add[n][c] (BResNet block tail: ECA gate x keep mask and the gate's pooled-path gradient).
extern "C" int sib_bn_bwd_apply_scaled(const void* dy, const float* dy_mul, const float* dy_add,
                                       const void* x, const float* mean_invstd, const float* gamma,
                                       const float* sums, void* dx, float* dgamma, float* dbeta,
                                       int N, int HW, int C, double count, float pgrad_scale,
                                       void* stream) {
  if (int rc = check_c(C)) return rc;
  SIB_CHECK(dy_mul != nullptr && dy_add != nullptr, "bn_bwd_apply_scaled: dy_mul and dy_add are required");
  const long M = (long)N * HW;
  const int grid = bn_grid(M, C);
  SIB_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false, false, false, true>, dim3(grid), dim3(kRedThreads), 0,
                      ST(stream), static_cast<const __nv_bfloat16*>(dy),
                      static_cast<const __nv_bfloat16*>(nullptr), static_cast<const float*>(nullptr),
                      static_cast<const __nv_bfloat16*>(x), mean_invstd, gamma, sums,
                      static_cast<const __nv_bfloat16*>(nullptr), static_cast<const float*>(nullptr),
                      static_cast<const float*>(nullptr), static_cast<__nv_bfloat16*>(dx),
                      static_cast<__nv_bfloat16*>(nullptr), static_cast<__nv_bfloat16*>(nullptr), dgamma,
                      dbeta, static_cast<float*>(nullptr), static_cast<float*>(nullptr), M, C,
                      (float)(1.0 / count), (int)SIB_ACT_NONE, 0.f, pgrad_scale,
                      static_cast<const float*>(nullptr), 0, 0.f, static_cast<__nv_bfloat16*>(nullptr),
                      dy_mul, dy_add, HW));
  SIB_LAUNCH_CHECK();
  return 0;
}

// bn_bwd_apply of a plain BatchNorm (+ activation) that also re-materialises the forward
// activation a = fwd_act(fmaf(x, act_ss)) for the consumer conv's weight gradient.
extern "C" int sib_bn_bwd_apply_remat(const void* dy, const float* mask_ss, const void* x,
                                      const float* mean_invstd, const float* gamma,
                                      const float* sums, void* dx, float* dgamma, float* dbeta,
                                      const float* act_ss, int fwd_act, float fwd_slope, void* a_out,
                                      long M, int C, double count, int act, float slope,
                                      float pgrad_scale, void* stream) {
  if (int rc = check_c(C)) return rc;
  SIB_CHECK(act == SIB_ACT_NONE || mask_ss != nullptr, "bn_bwd_apply_remat: activation mask needs `mask_ss`");
  SIB_CHECK(act_ss != nullptr && a_out != nullptr, "bn_bwd_apply_remat: act_ss and a_out are required");
  const int grid = bn_grid(M, C);
  SIB_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false, false, true>, dim3(grid), dim3(kRedThreads), 0,
                      ST(stream), static_cast<const __nv_bfloat16*>(dy),
                      static_cast<const __nv_bfloat16*>(nullptr), mask_ss,
                      static_cast<const __nv_bfloat16*>(x), mean_invstd, gamma, sums,
                      static_cast<const __nv_bfloat16*>(nullptr), static_cast<const float*>(nullptr),
                      static_cast<const float*>(nullptr), static_cast<__nv_bfloat16*>(dx),
                      static_cast<__nv_bfloat16*>(nullptr), static_cast<__nv_bfloat16*>(nullptr), dgamma,
                      dbeta, static_cast<float*>(nullptr), static_cast<float*>(nullptr), M, C,
                      (float)(1.0 / count), act, slope, pgrad_scale, act_ss, fwd_act, fwd_slope,
                      static_cast<__nv_bfloat16*>(a_out), static_cast<const float*>(nullptr),
                      static_cast<const float*>(nullptr), 1));
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
  SIB_CUDA(launch_pdl(maxpool_fwd_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
      static_cast<uint8_t*>(idx), N, H, W, C, OH, OW));
  SIB_LAUNCH_CHECK();
  return 0;
}

static int log2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

extern "C" int sib_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W,
                                    int C, void* stream) {
  const int cshift = C % 8 == 0 ? log2_exact(C / 8) : -1;
  SIB_CHECK(cshift >= 0, "maxpool_bwd: C/8 must be a power of two (got C=%d)", C);
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  int grid = N * ((H + 1) / 2);
  if (grid > sm_count() * 64) grid = sm_count() * 64;
  SIB_CUDA(launch_pdl(maxpool_bwd_kernel, dim3(grid), dim3(256), 0, ST(stream),
                      static_cast<const __nv_bfloat16*>(dy), static_cast<const uint8_t*>(idx),
                      static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH, OW, cshift));
  return 0;
}

extern "C" int sib_bn_act_maxpool3x3s2_fwd(const void* x, const float* stats, const float* gamma,
                                           const float* beta, float* running_mean,
                                           float* running_var, float* mean_invstd,
                                           float* scale_shift, void* y, void* idx, int N, int H,
                                           int W, int C, double count, float eps, float momentum,
                                           int act, float slope, void* stream) {
  const int cshift = C % 8 == 0 ? log2_exact(C / 8) : -1;
  SIB_CHECK(cshift >= 0 && cshift <= 8, "bn_act_maxpool: C/8 must be a power of two <= 256 (got C=%d)", C);
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  int grid = N * OH;
  if (grid > sm_count() * 64) grid = sm_count() * 64;
  BnFinalizeArgs f{stats, gamma, beta, running_mean, running_var, mean_invstd, scale_shift};
  SIB_CUDA(launch_pdl(bn_act_maxpool_fwd_kernel, dim3(grid), dim3(256), 0, ST(stream),
                      static_cast<const __nv_bfloat16*>(x), f, static_cast<__nv_bfloat16*>(y),
                      static_cast<uint8_t*>(idx), N, H, W, C, OH, OW, cshift, (float)count, eps,
                      momentum, act, slope));
  return 0;
}

extern "C" int sib_gap_fwd(const void* x, void* y, int N, int HW, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "gap: C %% 8 != 0");
  SIB_CUDA(launch_pdl(gap_fwd_kernel, dim3(ew_grid((long)N * (C / 8), 256)), dim3(256), 0, ST(stream), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, HW, C));
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_gap_bwd(const void* dy, void* dx, int N, int HW, int C, void* stream) {
  SIB_CHECK(C % 8 == 0, "gap: C %% 8 != 0");
  SIB_CUDA(launch_pdl(gap_bwd_kernel, dim3(ew_grid((long)N * HW * (C / 8), 256)), dim3(256), 0, ST(stream), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, HW, C));
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_upsample_zero(const void* dy, void* up, int N, int OH, int OW, int C, int UH,
                                 int UW, int stride, void* stream) {
  SIB_CHECK(C % 8 == 0, "upsample_zero: C %% 8 != 0");
  const long total = (long)N * UH * UW * (C / 8);
  SIB_CUDA(launch_pdl(upsample_zero_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(up), N, OH, OW, C, UH, UW,
      stride));
  SIB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sib_scatter_add_strided(const void* src, void* dst, int N, int OH, int OW, int C,
                                       int H, int W, int stride, void* stream) {
  SIB_CHECK(C % 8 == 0, "scatter_add: C %% 8 != 0");
  const long total = (long)N * OH * OW * (C / 8);
  SIB_CUDA(launch_pdl(scatter_add_strided_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), N, OH, OW, C, H, W,
      stride));
  SIB_LAUNCH_CHECK();
  return 0;
}
