// Fused classification heads.
//   * label-smoothing cross entropy (pytorch_tools.losses.smooth.CrossEntropyLoss as used by
//     reference arg_parser.py:140-142 / configs/hydra_exp/1.r50_baseline.yaml:34-35), forward
//     and d(loss)/d(logits) in one pass, index or dense (one-hot / soft) targets, temperature;
//   * ArcFace (reference angular_losses.py:128-146) and CosFace (angular_losses.py:186-198,
//     332-333) margins applied to cosine logits inside the same kernel;
//   * SphereLinearLayer (angular_losses.py:212-214): cos = normalize(x) . normalize(W)^T with
//     its backward through both normalisations.  fp32 operands and results; the three GEMMs
//     run on the tensor cores with every operand split into three bf16 terms (tc_gemm_f32x3_kernel).
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
// fp32-accurate GEMM on the tensor cores for the sphere-linear head
//   C[M][N] = sum_k A(m,k) * B(n,k)      A(m,k) = A[m*a_sm + k*a_sk], B(n,k) = B[n*b_sn + k*b_sk]
// The operands are fp32 in HBM (unit rows: the cosines feed acos / margins, and the reference's
// goldens hold them to 1e-6).  Every 8-element run of K is split on the fly into THREE bf16
// terms x = h + m + l (8 + 8 + 8 significant bits: the split is exact up to 2^-24 |x|), written by
// ordinary stores into the K-major 128B-swizzled layout a UMMA descriptor expects (row r at
// r*128 bytes, 16-byte chunk c at position c ^ (r&7)) and published to the async proxy; a k-block's
// product is the six bf16 MMAs whose weight is >= 2^-16 (l*h, h*l, m*m, m*h, h*m, h*h; tcgen05.mma
// kind::f16, M = 128, N = 64, K = 16), smallest terms first, in a TMEM accumulator that is
// CLEARED per k-block: the running sum over k-blocks lives in registers (round-to-nearest fp32
// adds), because the tensor core's own fp32 accumulation truncates and 100 chained accumulations
// cost ~1e-6 of the result (measured).  Either stride of an operand may be the contiguous one: the
// (row, chunk) -> thread mapping follows it, so global reads are coalesced for both orientations
// (dcos and dcos^T, wn and wn^T) and no transposed copy is ever made.  Two stages and two
// accumulators: the split of k-block i+1 overlaps the MMAs of k-block i.  Up to two independent
// problems per launch (blockIdx.z): the backward computes d(xn) and d(wn) with one kernel.
// ---------------------------------------------------------------------------
struct TcGemm {
  const float* A; long a_sm, a_sk;
  const float* B; long b_sn, b_sk;
  float* C;            // [M][N] row-major
  int M, N, K;
};
struct TcGemmBatch { TcGemm g[2]; };

constexpr int kTcBM = 128, kTcBN = 64, kTcBK = 64, kTcThreads = 256;
constexpr int kTcABytes = kTcBM * kTcBK * 2;               // one bf16 A tile (h, m or l)
constexpr int kTcBBytes = kTcBN * kTcBK * 2;
constexpr int kTcStageBytes = 3 * kTcABytes + 3 * kTcBBytes;   // A h | m | l | B h | m | l
constexpr int kTcSmem = 2 * kTcStageBytes + 1024;
constexpr uint32_t kTcTmemCols = 2 * kTcBN;                // two accumulators

// 8 consecutive k of one operand row (zero past the matrix edge)
__device__ __forceinline__ void tc_load_chunk(const float* __restrict__ src, long s_row, long s_k,
                                              int row, int nrows, int k0, int K, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (row < nrows && k0 < K) {
    const float* p = src + (long)row * s_row + (long)k0 * s_k;
    if (s_k == 1 && k0 + 8 <= K && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      load8f(p, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k0 + j < K) v[j] = __ldg(p + (long)j * s_k);
    }
  }
}
// ... -> the row's 16-byte chunk `c` of the h, m and l tiles
__device__ __forceinline__ void tc_split_chunk(const float* v, uint8_t* tiles, int tile_bytes, int r,
                                               int c) {
  float h[8], m[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
    const float r1 = v[j] - h[j];            // exact in fp32
    m[j] = __bfloat162float(__float2bfloat16_rn(r1));
    l[j] = r1 - m[j];                        // exact; rounded to bf16 by pack8
  }
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(tiles + off) = pack8(h);
  *reinterpret_cast<uint4*>(tiles + tile_bytes + off) = pack8(m);
  *reinterpret_cast<uint4*>(tiles + 2 * tile_bytes + off) = pack8(l);
}

__global__ void __launch_bounds__(kTcThreads)
tc_gemm_f32x3_kernel(const TcGemmBatch batch) {
  const TcGemm& g = batch.g[blockIdx.z];
  const int m0 = blockIdx.y * kTcBM, n0 = blockIdx.x * kTcBN;
  if (m0 >= g.M || n0 >= g.N) return;          // (uniform per CTA; nothing allocated yet)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t done_bar[2];             // the MMAs of the k-block in stage s have retired
  __shared__ uint32_t tmem_base_smem;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&done_bar[0], 1);
      mbar_init(&done_bar[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_smem, kTcTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  constexpr uint32_t idesc = umma_idesc_bf16(kTcBM, kTcBN, 0, 0);
  const bool a_kmajor = g.a_sk == 1, b_kmajor = g.b_sk == 1;
  const int nkb = (g.K + kTcBK - 1) / kTcBK;
  // warp w owns TMEM lane quarter w % 4 (32 output rows) and the 32-column half w / 4
  const int quarter = warp & 3, colh = warp >> 2;
  const uint32_t t_mine = tmem_base + ((uint32_t)(quarter * 32) << 16) + colh * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  // completion j of done_bar[s] belongs to k-block s + 2j; every thread waits for each one once
  auto drain = [&](int kb_done) {
    const int s = kb_done & 1;
    mbar_wait(&done_bar[s], (kb_done >> 1) & 1);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_mine + s * kTcBN, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
    tc_fence_before();
  };
  // software pipeline: the global loads of k-block kb+1 are issued right after the MMAs of kb and
  // stay in flight (48 registers) across the accumulator drain at the top of the next iteration
  constexpr int kAPer = kTcBM * 8 / kTcThreads, kBPer = kTcBN * 8 / kTcThreads;
  float va[kAPer][8], vb[kBPer][8];
  auto load_block = [&](int kb) {
    const int k0 = kb * kTcBK;
#pragma unroll
    for (int i = 0; i < kAPer; ++i) {
      const int q = tid + kTcThreads * i;
      const int r = a_kmajor ? (q >> 3) : (q & (kTcBM - 1));
      const int c = a_kmajor ? (q & 7) : (q / kTcBM);
      tc_load_chunk(g.A, g.a_sm, g.a_sk, m0 + r, g.M, k0 + c * 8, g.K, va[i]);
    }
#pragma unroll
    for (int i = 0; i < kBPer; ++i) {
      const int q = tid + kTcThreads * i;
      const int r = b_kmajor ? (q >> 3) : (q & (kTcBN - 1));
      const int c = b_kmajor ? (q & 7) : (q / kTcBN);
      tc_load_chunk(g.B, g.b_sn, g.b_sk, n0 + r, g.N, k0 + c * 8, g.K, vb[i]);
    }
  };
  load_block(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= 2) drain(kb - 2);                 // frees stage s and accumulator s
    uint8_t* a_t = smem + s * kTcStageBytes;
    uint8_t* b_t = a_t + 3 * kTcABytes;
#pragma unroll
    for (int i = 0; i < kAPer; ++i) {
      const int q = tid + kTcThreads * i;
      const int r = a_kmajor ? (q >> 3) : (q & (kTcBM - 1));
      const int c = a_kmajor ? (q & 7) : (q / kTcBM);
      tc_split_chunk(va[i], a_t, kTcABytes, r, c);
    }
#pragma unroll
    for (int i = 0; i < kBPer; ++i) {
      const int q = tid + kTcThreads * i;
      const int r = b_kmajor ? (q >> 3) : (q & (kTcBN - 1));
      const int c = b_kmajor ? (q & 7) : (q / kTcBN);
      tc_split_chunk(vb[i], b_t, kTcBBytes, r, c);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {                            // converged warp, one elected lane issues
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + s * kTcBN;
      const uint64_t da = umma_smem_desc(smem_u32(a_t), 16, 1024, kSwizzle128B);
      const uint64_t db = umma_smem_desc(smem_u32(b_t), 16, 1024, kSwizzle128B);
      constexpr uint32_t kAT = kTcABytes >> 4, kBT = kTcBBytes >> 4;   // tile strides, 16-byte units
#pragma unroll
      for (int k = 0; k < kTcBK / 16; ++k) {    // +32 bytes along K = +2 address units
        const uint64_t ah = da + 2 * k, am = ah + kAT, al = am + kAT;
        const uint64_t bh = db + 2 * k, bm = bh + kBT, bl = bm + kBT;
        umma_bf16_ss_w(d_tmem, al, bh, idesc, k != 0);
        umma_bf16_ss_w(d_tmem, ah, bl, idesc, 1);
        umma_bf16_ss_w(d_tmem, am, bm, idesc, 1);
        umma_bf16_ss_w(d_tmem, am, bh, idesc, 1);
        umma_bf16_ss_w(d_tmem, ah, bm, idesc, 1);
        umma_bf16_ss_w(d_tmem, ah, bh, idesc, 1);
      }
      umma_commit_w(&done_bar[s]);
    }
    if (kb + 1 < nkb) load_block(kb + 1);
  }
  if (nkb >= 2) drain(nkb - 2);
  drain(nkb - 1);
  {
    const int m = m0 + quarter * 32 + lane;
    const int nb = n0 + colh * 32;
    if (m < g.M) {
      float* dst = g.C + (long)m * g.N + nb;
      if (nb + 32 <= g.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j < g.N) dst[j] = acc[j];
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// rows -> unit rows; norm clamp 1e-12 as F.normalize.  Two row sets per launch (x and W).
struct RowSet {
  const float* src;
  float* dst;
  float* norms;
  int rows;
};
__global__ void normalize_rows_kernel(RowSet s0, RowSet s1, int D) {
  __shared__ float sh[32];
  long row = blockIdx.x;
  const RowSet& rs = row < s0.rows ? s0 : s1;
  if (row >= s0.rows) row -= s0.rows;
  const float* x = rs.src + row * D;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float v = x[i];
    s += v * v;
  }
  s = block_reduce(s, sh, false);
  const float nrm = fmaxf(sqrtf(s), 1e-12f);
  if (threadIdx.x == 0) rs.norms[row] = nrm;
  for (int i = threadIdx.x; i < D; i += blockDim.x) rs.dst[row * D + i] = x[i] / nrm;
}

// dx = (dxn - xn * <xn, dxn>) / ||x||   (rows of x and rows of W in one launch)
struct RowSetBwd {
  const float* dxn;
  const float* xn;
  const float* norms;
  float* dx;
  int rows;
};
__global__ void normalize_bwd_kernel(RowSetBwd s0, RowSetBwd s1, int D) {
  __shared__ float sh[32];
  long row = blockIdx.x;
  const RowSetBwd& rs = row < s0.rows ? s0 : s1;
  if (row >= s0.rows) row -= s0.rows;
  const float* dxn = rs.dxn + row * D;
  const float* xn = rs.xn + row * D;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s += dxn[i] * xn[i];
  s = block_reduce(s, sh, false);
  const float inv = 1.f / rs.norms[row];
  for (int i = threadIdx.x; i < D; i += blockDim.x) rs.dx[row * D + i] = (dxn[i] - xn[i] * s) * inv;
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

static int tc_gemm(const TcGemm* probs, int n, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(tc_gemm_f32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kTcSmem));
    configured = true;
  }
  TcGemmBatch b{};
  int gx = 0, gy = 0;
  for (int i = 0; i < n; ++i) {
    b.g[i] = probs[i];
    gx = max(gx, (probs[i].N + kTcBN - 1) / kTcBN);
    gy = max(gy, (probs[i].M + kTcBM - 1) / kTcBM);
  }
  if (gx == 0 || gy == 0) return 0;
  tc_gemm_f32x3_kernel<<<dim3(gx, gy, n), kTcThreads, kTcSmem, st>>>(b);
  SIB_LAUNCH_CHECK();
  return 0;
}

// cos[B][C] = normalize(x[B][D]) . normalize(w[C][D])^T ; xn/wn/norms are saved for backward
extern "C" int sib_sphere_linear_fwd(const float* x, const float* w, float* cosv, float* xn,
                                     float* wn, float* xnorm, float* wnorm, int B, int C, int D,
                                     int normalize_x, void* stream) {
  const RowSet sx{x, xn, xnorm, normalize_x ? B : 0};
  const RowSet sw{w, wn, wnorm, C};
  normalize_rows_kernel<<<sx.rows + sw.rows, 128, 0, ST(stream)>>>(sx, sw, D);
  SIB_LAUNCH_CHECK();
  const TcGemm g{normalize_x ? xn : x, D, 1, wn, D, 1, cosv, B, C, D};
  return tc_gemm(&g, 1, ST(stream));
}

// dx[B][D], dw[C][D] from dcos[B][C]; scratch must hold (B + C) * D floats
extern "C" int sib_sphere_linear_bwd(const float* dcos, const float* x_or_xn, const float* wn,
                                     const float* xnorm, const float* wnorm, float* dx, float* dw,
                                     float* scratch, int B, int C, int D, int normalize_x,
                                     void* stream) {
  float* dxn = (dx != nullptr && !normalize_x) ? dx : scratch;      // d(xn) = dcos . wn      [B][D]
  float* dwn = scratch + (long)B * D;                               // d(wn) = dcos^T . xn    [C][D]
  TcGemm g[2];
  int n = 0;
  if (dx != nullptr) g[n++] = TcGemm{dcos, C, 1, wn, 1, D, dxn, B, D, C};
  if (dw != nullptr) g[n++] = TcGemm{dcos, 1, C, x_or_xn, 1, D, dwn, C, D, B};
  if (n == 0) return 0;
  if (int rc = tc_gemm(g, n, ST(stream))) return rc;
  const RowSetBwd bx{dxn, x_or_xn, xnorm, dx, (dx != nullptr && normalize_x) ? B : 0};
  const RowSetBwd bw{dwn, wn, wnorm, dw, dw != nullptr ? C : 0};
  if (bx.rows + bw.rows > 0) {
    normalize_bwd_kernel<<<bx.rows + bw.rows, 128, 0, ST(stream)>>>(bx, bw, D);
    SIB_LAUNCH_CHECK();
  }
  return 0;
}
