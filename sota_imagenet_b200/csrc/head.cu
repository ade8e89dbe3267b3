// Fused classification heads.
//   * label-smoothing cross entropy (pytorch_tools.losses.smooth.CrossEntropyLoss as used by
//     reference arg_parser.py:140-142 / configs/hydra_exp/1.r50_baseline.yaml:34-35), forward
//     and d(loss)/d(logits) in one pass, index or dense (one-hot / soft) targets, temperature;
//   * ArcFace (reference angular_losses.py:128-146) and CosFace (angular_losses.py:186-198,
//     332-333) margins applied to cosine logits inside the same kernel;
//   * SphereLinearLayer (angular_losses.py:212-214): cos = normalize(x) . normalize(W)^T with
//     its backward through both normalisations.  fp32 throughout (tiny problem: B x 1000 x 512).
#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

template <class T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// clamp range of the arccos heads: EPS = 1e-7 (angular_losses.py:324,574) in fp32
#define ARCCOS_HI (1.f - 1e-7f)
#define ARCCOS_LO (-1.f + 1e-7f)

struct MarginParams {
  int kind;        // SIB_MARGIN_*
  float s;         // logit scale
  float cos_m, sin_m, th, mm;   // ArcFace constants (angular_losses.py:122-125)
  float m;         // CosFace margin
};

__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

// one CTA per sample row
template <class T>
__global__ void __launch_bounds__(256)
ce_kernel(const T* __restrict__ logits, const long* __restrict__ labels,
          const float* __restrict__ dense_t, int B, int C, int ld, float smoothing,
          float inv_temp, MarginParams mp, float* __restrict__ loss_rows,
          T* __restrict__ dlogits, float grad_scale) {
  extern __shared__ float z[];          // margin-modified, scaled logits
  __shared__ float sh[32];
  const int row = blockIdx.x;
  const T* x = logits + (long)row * ld;
  const long y = labels ? labels[row] : -1;
  float dphi = 1.f;                     // d(target logit)/d(input) before scaling
  float lmax = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = to_f<T>(x[c]);
    if (mp.kind == SIB_MARGIN_ARCCOS) {
      // z = -(acos(clamp(x)) + m [target]) * s   (ArcCosSoftmax :572-576; AdaCos arc_logits :326-329)
      float th = acosf(fminf(fmaxf(v, ARCCOS_LO), ARCCOS_HI));
      const bool tgt = labels ? (c == y) : (dense_t[(long)row * C + c] != 0.f);
      if (tgt) th += mp.m;
      v = -th * mp.s;
    } else if (mp.kind != SIB_MARGIN_NONE) {
      if (c == y) {
        if (mp.kind == SIB_MARGIN_ARC || mp.kind == SIB_MARGIN_ARC_PURE) {
          const float sine = sqrtf(fmaxf(1.f - v * v, 0.f));
          const float phi = v * mp.cos_m - sine * mp.sin_m;
          if (v > mp.th || mp.kind == SIB_MARGIN_ARC_PURE) {
            dphi = mp.cos_m + mp.sin_m * v / fmaxf(sine, 1e-6f);
            v = phi;
          } else {
            v = v - mp.mm;
          }
        } else {
          v = v - mp.m;
        }
      } else if (mp.kind == SIB_MARGIN_COS && labels == nullptr && dense_t != nullptr &&
                 dense_t[(long)row * C + c] != 0.f) {
        v = v - mp.m;   // soft / one-hot targets: margin on every column with target mass
      }
      v *= mp.s;
    }
    v *= inv_temp;
    z[c] = v;
    lmax = fmaxf(lmax, v);
  }
  // dphi lives in the thread that owns the target column; broadcast through smem
  __shared__ float dphi_sh;
  if (threadIdx.x == 0) dphi_sh = 1.f;
  __syncthreads();
  if (y >= 0 && (int)(y % blockDim.x) == (int)threadIdx.x) dphi_sh = dphi;
  lmax = block_reduce(lmax, sh, true);
  float se = 0.f, sz = 0.f, stz = 0.f, st = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = z[c];
    se += __expf(v - lmax);
    sz += v;
    if (dense_t) {
      const float t = dense_t[(long)row * C + c];
      st += t;
      stz += t * v;
    }
  }
  se = block_reduce(se, sh, false);
  sz = block_reduce(sz, sh, false);
  if (dense_t) {
    st = block_reduce(st, sh, false);
    stz = block_reduce(stz, sh, false);
  } else {
    st = 1.f;
    stz = z[y];   // visible after the syncs above
  }
  const float lse = lmax + __logf(se);
  // -sum_c t_c logp_c = st*lse - stz ;  -mean_c logp_c = lse - sz/C
  const float loss = (1.f - smoothing) * (st * lse - stz) + smoothing * (lse - sz / C);
  if (threadIdx.x == 0) loss_rows[row] = loss;
  if (dlogits != nullptr) {
    const float wsum = (1.f - smoothing) * st + smoothing;
    const float dphi_t = dphi_sh;
    T* d = dlogits + (long)row * ld;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float pr = __expf(z[c] - lse);
      const float t = dense_t ? dense_t[(long)row * C + c] : (c == y ? 1.f : 0.f);
      float g = pr * wsum - ((1.f - smoothing) * t + smoothing / C);
      g *= grad_scale * inv_temp;
      if (mp.kind == SIB_MARGIN_ARCCOS) {
        // d(-acos(clamp(v)))/dv = 1/sqrt(1-v^2) inside the clamp range, 0 outside (torch clamp)
        const float v0 = to_f<T>(x[c]);
        g *= (v0 >= ARCCOS_LO && v0 <= ARCCOS_HI) ? mp.s * rsqrtf(1.f - v0 * v0) : 0.f;
      } else if (mp.kind != SIB_MARGIN_NONE) {
        g *= mp.s;
        if (c == y) g *= dphi_t;
      }
      d[c] = from_f<T>(g);
    }
  }
}

__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_reduce(s, sh, false);
  if (threadIdx.x == 0) *out = s / n;
}

// ---------------------------------------------------------------------------
// small fp32 GEMM family for the sphere-linear head
//   C[M][N] = sum_k A(m,k) * B(n,k)      (A: [M][K] or [K][M] via strides, same for B)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long a_sm, long a_sk, const float* __restrict__ Bm,
             long b_sn, long b_sk, float* __restrict__ Cm, int M, int N, int K) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int kk = i & 15, mm = i >> 4;
      const int m = m0 + mm, n = n0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[m * a_sm + k * a_sk] : 0.f;
      Bs[kk][mm] = (n < N && k < K) ? Bm[n * b_sn + k * b_sk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) Cm[(long)m * N + n] = acc[i][j];
    }
}

// rows -> unit rows; norm clamp 1e-12 as F.normalize
__global__ void normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ xn,
                                      float* __restrict__ norms, int D) {
  __shared__ float sh[32];
  const long row = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float v = x[row * D + i];
    s += v * v;
  }
  s = block_reduce(s, sh, false);
  const float nrm = fmaxf(sqrtf(s), 1e-12f);
  if (threadIdx.x == 0) norms[row] = nrm;
  for (int i = threadIdx.x; i < D; i += blockDim.x) xn[row * D + i] = x[row * D + i] / nrm;
}

// dx = (dxn - xn * <xn, dxn>) / ||x||
__global__ void normalize_bwd_kernel(const float* __restrict__ dxn, const float* __restrict__ xn,
                                     const float* __restrict__ norms, float* __restrict__ dx,
                                     int D) {
  __shared__ float sh[32];
  const long row = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s += dxn[row * D + i] * xn[row * D + i];
  s = block_reduce(s, sh, false);
  const float inv = 1.f / norms[row];
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    dx[row * D + i] = (dxn[row * D + i] - xn[row * D + i] * s) * inv;
}

}  // namespace sib

using namespace sib;
#define ST(s) static_cast<cudaStream_t>(s)

static MarginParams make_margin(int kind, float s, float m) {
  MarginParams mp{};
  mp.kind = kind;
  mp.s = s;
  mp.m = m;
  if (kind == SIB_MARGIN_ARC || kind == SIB_MARGIN_ARC_PURE) {
    const double pi = 3.14159265358979323846;
    mp.cos_m = (float)cos((double)m);
    mp.sin_m = (float)sin((double)m);
    mp.th = (float)cos(pi - (double)m);
    mp.mm = (float)(sin(pi - (double)m) * (double)m);
  }
  return mp;
}

extern "C" int sib_ce_fwd_bwd(const void* logits, int logits_fp32, const long* labels,
                              const float* dense_targets, int B, int C, int ld, float smoothing,
                              float temperature, int margin_kind, float s, float m,
                              float* loss_rows, float* loss_mean, void* dlogits, float grad_scale,
                              void* stream) {
  SIB_CHECK(labels != nullptr || dense_targets != nullptr, "ce: no targets given");
  SIB_CHECK((margin_kind != SIB_MARGIN_ARC && margin_kind != SIB_MARGIN_ARC_PURE) || labels != nullptr,
            "ce: ArcFace needs index labels (reference angular_losses.py:140 scatter_)");
  SIB_CHECK(C * sizeof(float) <= 96 * 1024, "ce: too many classes for one CTA (%d)", C);
  const MarginParams mp = make_margin(margin_kind, s, m);
  const size_t smem = sizeof(float) * C;
  if (logits_fp32) {
    if (smem > 48 * 1024)
      SIB_CUDA(cudaFuncSetAttribute(ce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    ce_kernel<float><<<B, 256, smem, ST(stream)>>>(
        static_cast<const float*>(logits), labels, dense_targets, B, C, ld, smoothing,
        1.f / temperature, mp, loss_rows, static_cast<float*>(dlogits), grad_scale / B);
  } else {
    if (smem > 48 * 1024)
      SIB_CUDA(cudaFuncSetAttribute(ce_kernel<__nv_bfloat16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ce_kernel<__nv_bfloat16><<<B, 256, smem, ST(stream)>>>(
        static_cast<const __nv_bfloat16*>(logits), labels, dense_targets, B, C, ld, smoothing,
        1.f / temperature, mp, loss_rows, static_cast<__nv_bfloat16*>(dlogits), grad_scale / B);
  }
  SIB_LAUNCH_CHECK();
  if (loss_mean != nullptr) {
    mean_kernel<<<1, 256, 0, ST(stream)>>>(loss_rows, B, loss_mean);
    SIB_LAUNCH_CHECK();
  }
  return 0;
}

static int sgemm(const float* A, long a_sm, long a_sk, const float* Bm, long b_sn, long b_sk,
                 float* Cm, int M, int N, int K, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, st>>>(A, a_sm, a_sk, Bm, b_sn, b_sk, Cm, M, N, K);
  SIB_LAUNCH_CHECK();
  return 0;
}

// cos[B][C] = normalize(x[B][D]) . normalize(w[C][D])^T ; xn/wn/norms are saved for backward
extern "C" int sib_sphere_linear_fwd(const float* x, const float* w, float* cosv, float* xn,
                                     float* wn, float* xnorm, float* wnorm, int B, int C, int D,
                                     int normalize_x, void* stream) {
  if (normalize_x) {
    normalize_rows_kernel<<<B, 128, 0, ST(stream)>>>(x, xn, xnorm, D);
    SIB_LAUNCH_CHECK();
  }
  normalize_rows_kernel<<<C, 128, 0, ST(stream)>>>(w, wn, wnorm, D);
  SIB_LAUNCH_CHECK();
  return sgemm(normalize_x ? xn : x, D, 1, wn, D, 1, cosv, B, C, D, ST(stream));
}

// dx[B][D], dw[C][D] from dcos[B][C]; scratch must hold max(B,C)*D floats
extern "C" int sib_sphere_linear_bwd(const float* dcos, const float* x_or_xn, const float* wn,
                                     const float* xnorm, const float* wnorm, float* dx, float* dw,
                                     float* scratch, int B, int C, int D, int normalize_x,
                                     void* stream) {
  // d(xn) = dcos . wn   -> [B][D]
  if (dx != nullptr) {
    if (normalize_x) {
      if (int rc = sgemm(dcos, C, 1, wn, 1, D, scratch, B, D, C, ST(stream))) return rc;
      normalize_bwd_kernel<<<B, 128, 0, ST(stream)>>>(scratch, x_or_xn, xnorm, dx, D);
      SIB_LAUNCH_CHECK();
    } else {
      if (int rc = sgemm(dcos, C, 1, wn, 1, D, dx, B, D, C, ST(stream))) return rc;
    }
  }
  // d(wn) = dcos^T . xn -> [C][D]
  if (dw != nullptr) {
    if (int rc = sgemm(dcos, 1, C, x_or_xn, 1, D, scratch, C, D, B, ST(stream))) return rc;
    normalize_bwd_kernel<<<C, 128, 0, ST(stream)>>>(scratch, wn, wnorm, dw, D);
    SIB_LAUNCH_CHECK();
  }
  return 0;
}
