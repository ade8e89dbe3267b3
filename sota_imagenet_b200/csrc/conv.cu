// NHWC bf16 implicit-GEMM convolution on sm_100a tensor cores.
//
//   fprop / dgrad : D[M = pixels][N = out channels] = sum_k A[pixels][k] * B[out ch][k]
//                   A gathered by TMA im2col loads (one filter tap x 64 channels per k-block),
//                   B = packed weights [Cout][R*S*Cin] (K-major), both 128B-swizzled in smem,
//                   tcgen05.mma (M=128, N=BN, K=16) accumulating fp32 in TMEM.
//   wgrad         : dW[Cout][tap][Cin] = sum_pixels dY[pixel][Cout] * X_im2col[pixel][tap, Cin]
//                   both operands MN-major, split over pixel ranges, fp32 red.add to dW.
//
// Replaces what the reference reaches through cuDNN: F.conv2d / autograd conv backward
// inside pytorch_tools.models.resnet50 (reference train.py:64).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

constexpr int kBM = 128;        // GEMM-M tile (output pixels / wgrad out channels)
constexpr int kBK = 64;         // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int kABytes = kBM * kBK * 2;
constexpr int kThreads = 192;   // wgrad kernel: warp0 TMA, warp1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int kIgemmThreads = 384;  // igemm: warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-11 epilogue
constexpr int kIgemmThreadsPro = 512;   // + warps 12-15: BatchNorm/activation transform of the A operand
constexpr int kProMaxC = 512;           // input channels the fused prologue's scale/shift table holds
// threads of an igemm CTA: 4 control warps + epilogue warps (8, or 16 with CG = 4) + 4 transform warps (PRO)
constexpr int igemm_epi_warps(int BN, int CG) { return BN == 64 ? 8 : 4 * (CG ? CG : 2); }
constexpr int igemm_threads(int BN, bool PRO, int CG) { return 128 + 32 * igemm_epi_warps(BN, CG) + (PRO ? 128 : 0); }
constexpr int kSlabBytes = 32 * 128;   // epilogue staging: 32 rows x 64 bf16, 128B-swizzled

// 4 consecutive fp32 sums in one L2 operation (16-byte aligned address)
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

struct IgemmParams {
  int M_total;       // GEMM M (pixels of the traversal space)
  int Cout;          // GEMM N
  int num_kblocks;   // taps * ceil(Cin/64)
  int cin_blocks;    // ceil(Cin / 64)
  int S;             // filter width (tap = r * S + s)
  int trav_hw, trav_w;  // traversal space: pixels per image, pixels per row
  int stride, pad_h, pad_w;   // base pixel = (p*stride - pad_h, q*stride - pad_w)
  int tiled_a;       // A is a plain [M][K] matrix (1x1 stride-1)
  int num_m_tiles, num_n_tiles;
  int has_residual;  // out = acc + residual (same geometry as out)
  const float* bias; // optional [Cout]
  float* stats;      // optional [2][Cout]: sum, sum of squares of the stored bf16 values
                     // (fused BN backward: sum g, sum g * xhat)
  // ---- fused BatchNorm-backward reduction (dgrad epilogue; kernels with AUX >= 1) ----
  //   g = acc [+ residual], masked by the activation derivative of the BN this gradient flows
  //   into; the mask comes from tile `aux1`: fmaf(aux1, scale, shift) > 0 when mask_ss is given
  //   (aux1 = that BN's input, the forward's own arithmetic), else aux1 > 0 (aux1 = the stored
  //   block output).  xhat is built from aux1 (fuse == 1) or aux2 (fuse == 2).
  int fuse;          // 0 off, 1 = one auxiliary tile, 2 = two
  int fuse_act;      // SIB_ACT_* of the BN's activation
  float fuse_slope;
  const float* mask_ss;      // [2][Cout] scale, shift (or null)
  const float* mean_invstd;  // [2][Cout]
  // ---- strided output view (stride-2 dgrad by row parity, see dgrad_s2_impl) ----
  //   GEMM row m = (image row, q) with q in [0, out_q): out / residual / aux tiles are addressed
  //   through 3-D tensor maps (channel, q, image row); out_q in {8, 16, 32} divides the 32-row
  //   slab, columns q >= the real width are clipped by the TMA unit.  0 = plain [M][Cout].
  int out_q;
  int stat_c;        // real channel count behind the GEMM columns: column j is channel j % stat_c
                     // for `stats`, `mask_ss` and `mean_invstd` (== Cout except for the parity GEMMs)
};

// ---- fused BatchNorm (+ activation) prologue on the A operand (fprop; kernels with PRO) ----
//   The conv consumes a = act(fmaf(c, scale, shift)) where c is the raw output of the previous
//   conv: TMA lands the raw tile in shared memory, four transform warps rewrite it in place
//   (zero-padding taps stay zero), fence.proxy.async, and only then the MMA warp may read it.
//   The normalised activation is never written to HBM.  scale/shift come from the raw batch
//   statistics of the producer conv's epilogue (training: every CTA derives the table, CTA 0
//   publishes mean/invstd, scale/shift and the running statistics, i.e. bn_finalize folded in)
//   or from `scale_shift` itself (eval: stats == nullptr).
struct BnPrologue {
  const float* stats;        // [2][Cin] sum, sum of squares (already all-reduced under SyncBN) or null
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean_invstd;        // out [2][Cin] (training)
  float* scale_shift;        // out [2][Cin] (training) / in (eval)
  float count, eps, momentum;
  int act;
  float slope;
  int IH, IW, Cin;           // input extent: taps outside it are padding and stay zero
};

// setmaxnreg: the PRO kernels run 512 threads (128 registers each at launch); the producer /
// MMA warpgroup and the transform warpgroup give registers back so the two epilogue
// warpgroups keep the 168 they need ((64 + 168 + 168 + 104) * 128 <= 64 Ki registers).
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- epilogue (shared by the 1-CTA and 2-CTA kernels) ----
//   warps 4-11 (8 warps; warp e reads TMEM lane quarter e%4 and the 64-column chunks with
//   chunk%2 == e/4; BN=64 uses the first four):
//   TMEM -> regs -> (+bias, +residual) -> bf16 -> 128B-swizzled smem slab -> TMA store, plus
//   per-channel sum / sum-of-squares of the stored values for the following BatchNorm, or (fused
//   BN backward) activation masking of the gradient and sum g / sum g*xhat.
//   CG (column groups, default 2): warps e / 4 == g take the 64-column chunks with chunk % CG == g.
//   CG = 4 with BN = 256 ("16-warp epilogue"): every warp drains ONE chunk per tile and four warps
//   share a scheduler instead of two -- the BN = 256 kernels with short K (all 1x1 layers) are bound
//   by the latency of this epilogue, not by the tensor pipe or HBM.
template <int BN, int SLABS, int AUX, bool kTwoCta, int CG = 0>
__device__ __forceinline__ void igemm_epilogue(
    const CUtensorMap* tmOut, const CUtensorMap* tmRes, const CUtensorMap* tmAux1,
    const CUtensorMap* tmAux2, const IgemmParams& p, uint8_t* smem_slab, uint8_t* smem_aux,
    uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint64_t* res_bar, float* s_part,
    uint32_t tmem_base, int first_tile, int tile_step, int num_tiles, int cta_rank) {
  constexpr int kChunks = BN / 64;
  constexpr int kHalves = CG ? CG : (kChunks >= 2 ? 2 : 1);
  // BN = 64 has a single 64-column chunk: its eight epilogue warps form TWO groups that take
  // alternate tiles (group g drains accumulator stage g), so two epilogues are in flight and
  // their latency no longer bounds the tile rate of the N = 64 kernels (stem, 64-channel 1x1).
  constexpr int kGroups = (BN == 64 && !kTwoCta) ? 2 : 1;
  constexpr int kEpiWarps = 4 * kHalves * kGroups;
  constexpr int kChunksPerWarp = kChunks / kHalves;
  constexpr int kPartStride = kChunksPerWarp * 64;       // s_part[e][2][kPartStride]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int e = warp - 4;
  const int quarter = e & 3;                // TMEM lane quarter (== warp % 4)
  const int group = kGroups == 2 ? (e >> 2) : 0;
  const int half = kGroups == 2 ? 0 : (e >> 2);   // which interleaved set of 64-column chunks
  uint8_t* slabs = smem_slab + e * SLABS * kSlabBytes;
  uint8_t* aux = smem_aux + e * (AUX > 0 ? AUX : 1) * kSlabBytes;
  const int et = threadIdx.x - 128;         // 0 .. 32*kEpiWarps-1
  // swizzled 16-byte slots of this lane's row inside a slab (row = lane)
  uint32_t row_slot[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) row_slot[g] = lane * 128 + ((g ^ (lane & 7)) << 4);
  // stats pass: lane reads 16 bytes (8 columns, slot lane%8) of rows it*4 + lane/8; the row's
  // swizzle phase (row & 7) alternates between st_row and st_row + 4 with the parity of `it`
  const int st_row = lane >> 3, st_slot = lane & 7;
  const uint32_t st_off0 = st_row * 128 + ((st_slot ^ st_row) << 4);
  const uint32_t st_off1 = st_row * 128 + ((st_slot ^ (st_row + 4)) << 4);
  const bool fused = AUX > 0 && p.fuse != 0;
  const int n_loads = (p.has_residual ? 1 : 0) + (fused ? p.fuse : 0);
  // One n-tile per launch (Cout <= BN, all the 56x56 layers): every tile of this CTA covers the
  // same columns, so the per-channel sums are accumulated in this warp's s_part slots for the
  // whole kernel and published once at the end -- no named barrier and no 2*BN atomics per tile.
  // The same holds with several n-tiles when the launcher made the tile step a multiple of their
  // number (run_igemm): tile % num_n_tiles is then the same for every tile of this CTA.
  const bool cta_sums = p.stats != nullptr &&
                        (p.num_n_tiles == 1 || kGroups == 2 || tile_step % p.num_n_tiles == 0);
  const int n0_cta = (first_tile % p.num_n_tiles) * BN;      // this CTA's fixed column offset (cta_sums)
  float2 acc1[kChunksPerWarp][4], acc2[kChunksPerWarp][4];
#pragma unroll
  for (int ci = 0; ci < kChunksPerWarp; ++ci) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc1[ci][j] = make_float2(0.f, 0.f); acc2[ci][j] = make_float2(0.f, 0.f); }
  }
  int acc = group;
  uint32_t acc_phase = 0;
  uint32_t res_phase = 0;
  int slab_idx = 0;
  int local = 0;
  for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++local) {
    if (kGroups == 2 && (local & 1) != group) continue;
    const int m0 = kTwoCta ? (tile / p.num_n_tiles) * 2 * kBM + cta_rank * kBM
                           : (tile / p.num_n_tiles) * kBM;
    const int n0 = (tile % p.num_n_tiles) * BN;
    const int row0 = m0 + quarter * 32;
    const int rows_valid = p.M_total - row0;   // rows of this 32-row slab that exist
    mbar_wait(&tmem_full_bar[acc], acc_phase);
    tc_fence_after();
    bool released = false;
#pragma unroll
    for (int ci = 0; ci < kChunksPerWarp; ++ci) {
      const int chunk = half + ci * kHalves;
      const int col0 = n0 + chunk * 64;
      const bool live = col0 < p.Cout;           // ragged N (warp-uniform)
      // packed fp32x2 accumulators (add.f32x2 / fma.f32x2 on sm_100); with cta_sums they run on
      // across all tiles of the CTA and are reduced across lanes only once, after the tile loop
      float2 (&cs1)[4] = acc1[ci];
      float2 (&cs2)[4] = acc2[ci];
      if (!cta_sums) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { cs1[j] = make_float2(0.f, 0.f); cs2[j] = make_float2(0.f, 0.f); }
      }
      if (live) {
        uint8_t* slab = slabs + slab_idx * kSlabBytes;
        // the TMA store that last read this slab must have finished reading it
        if (lane == 0) tma_store_wait_read<SLABS - 1>();
        __syncwarp();
        if (n_loads != 0 && lane == 0) {
          mbar_arrive_expect_tx(&res_bar[e], n_loads * kSlabBytes);
          if (p.out_q == 0) {
            if (p.has_residual) tma_load_2d(slab, tmRes, &res_bar[e], col0, row0);
            if (AUX > 0 && fused) {
              tma_load_2d(aux, tmAux1, &res_bar[e], col0, row0);
              if (AUX > 1 && p.fuse == 2) tma_load_2d(aux + kSlabBytes, tmAux2, &res_bar[e], col0, row0);
            }
          } else {
            const int r3 = row0 / p.out_q;
            if (p.has_residual) tma_load_3d(slab, tmRes, &res_bar[e], col0, 0, r3);
            if (AUX > 0 && fused) {
              tma_load_3d(aux, tmAux1, &res_bar[e], col0, 0, r3);
              if (AUX > 1 && p.fuse == 2) tma_load_3d(aux + kSlabBytes, tmAux2, &res_bar[e], col0, 0, r3);
            }
          }
        }
        uint32_t r[64];
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chunk * 64;
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_32x32b_x32(taddr + 32, r + 32);
        tmem_ld_wait();
        if (ci == kChunksPerWarp - 1 || col0 + 64 * kHalves >= p.Cout) {
          // accumulator fully read by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kTwoCta) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
          released = true;
        }
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (col0 + j < p.Cout) v[j] += __ldg(p.bias + col0 + j);
        }
        if (n_loads != 0) {
          mbar_wait(&res_bar[e], res_phase);
          res_phase ^= 1;
        }
        if (p.has_residual) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint4 q = *reinterpret_cast<const uint4*>(slab + row_slot[g]);
            float prev[8];
            unpack8(q, prev);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] += prev[j];
          }
        }
        // bf16 pack into the 128B-swizzled slab.  Rows past M_total (last m-tile only; the TMA
        // store clips them) are written as zeros so that the statistics pass needs no row mask.
        if (rows_valid < 32 && lane >= rows_valid) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(slab + row_slot[g]) = pack8(&v[g * 8]);
        if (AUX > 0 && fused) {
          __syncwarp();
          // column-owner pass: this lane owns 8 channels; mask the gradient in place and
          // accumulate sum g, sum g * xhat
          const int cbase = col0 + st_slot * 8;
          float sc[8], sh[8], mu[8], is[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; mu[j] = 0.f; is[j] = 0.f; }
          if (cbase < p.Cout) {
            const int ch = cbase % p.stat_c;
            if (p.mask_ss != nullptr) {
              load8f(p.mask_ss + ch, sc);
              load8f(p.mask_ss + p.stat_c + ch, sh);
            }
            load8f(p.mean_invstd + ch, mu);
            load8f(p.mean_invstd + p.stat_c + ch, is);
          }
          const bool do_mask = p.fuse_act != SIB_ACT_NONE;
          const float neg = p.fuse_act == SIB_ACT_LEAKY ? p.fuse_slope : 0.f;
          const uint8_t* xsl = (AUX > 1 && p.fuse == 2) ? aux + kSlabBytes : aux;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const uint32_t off = it * 512 + ((it & 1) ? st_off1 : st_off0);
            const uint4 q = *reinterpret_cast<const uint4*>(slab + off);   // (rows past M_total hold zeros)
            float g[8], mv[8], xv[8];
            unpack8(q, g);
            unpack8(*reinterpret_cast<const uint4*>(aux + off), mv);
            unpack8(*reinterpret_cast<const uint4*>(xsl + off), xv);
            if (do_mask) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float z = fmaf(mv[j], sc[j], sh[j]);
                g[j] = z > 0.f ? g[j] : g[j] * neg;
              }
              *reinterpret_cast<uint4*>(slab + off) = pack8(g);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 gg = make_float2(g[2 * j], g[2 * j + 1]);
              const float2 xh = make_float2((xv[2 * j] - mu[2 * j]) * is[2 * j],
                                            (xv[2 * j + 1] - mu[2 * j + 1]) * is[2 * j + 1]);
              cs1[j] = __fadd2_rn(cs1[j], gg);
              cs2[j] = __ffma2_rn(gg, xh, cs2[j]);
            }
          }
        } else if (p.stats != nullptr) {
          __syncwarp();
          // column sums of the bf16 values as stored: 8 x LDS.128 cover the 32 x 64 slab
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const uint4 q = *reinterpret_cast<const uint4*>(slab + it * 512 + ((it & 1) ? st_off1 : st_off0));
            const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(hq[j]);
              cs1[j] = __fadd2_rn(cs1[j], f);
              cs2[j] = __ffma2_rn(f, f, cs2[j]);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.out_q == 0) tma_store_2d(tmOut, slab, col0, row0);
          else tma_store_3d(tmOut, slab, col0, 0, row0 / p.out_q);
          tma_store_commit();
        }
        if (SLABS > 1) slab_idx ^= 1;
      }
      if (p.stats != nullptr && !cta_sums) {
        // lanes with equal lane%8 hold partial sums of the same 8 columns (different rows)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            float2 t1, t2;
            t1.x = __shfl_xor_sync(0xffffffffu, cs1[j].x, o);
            t1.y = __shfl_xor_sync(0xffffffffu, cs1[j].y, o);
            t2.x = __shfl_xor_sync(0xffffffffu, cs2[j].x, o);
            t2.y = __shfl_xor_sync(0xffffffffu, cs2[j].y, o);
            cs1[j] = __fadd2_rn(cs1[j], t1);
            cs2[j] = __fadd2_rn(cs2[j], t2);
          }
        }
        if (lane < 8) {
          float4* d1 = reinterpret_cast<float4*>(&s_part[(e * 2 + 0) * kPartStride + ci * 64 + lane * 8]);
          float4* d2 = reinterpret_cast<float4*>(&s_part[(e * 2 + 1) * kPartStride + ci * 64 + lane * 8]);
          float4 a0 = make_float4(cs1[0].x, cs1[0].y, cs1[1].x, cs1[1].y);
          float4 a1 = make_float4(cs1[2].x, cs1[2].y, cs1[3].x, cs1[3].y);
          float4 b0 = make_float4(cs2[0].x, cs2[0].y, cs2[1].x, cs2[1].y);
          float4 b1 = make_float4(cs2[2].x, cs2[2].y, cs2[3].x, cs2[3].y);
          d1[0] = a0; d1[1] = a1; d2[0] = b0; d2[1] = b1;
        }
      }
    }
    if (!released) {     // every chunk of this warp was past Cout: still release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kTwoCta) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    if (p.stats != nullptr && !cta_sums) {
      // combine the four row-quarters of each column and publish; s_part is reused next tile
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      // 2 * BN / 4 vector items: (sum | sum of squares) x groups of 4 columns
      for (int idx = et; idx < BN / 2; idx += 32 * kEpiWarps) {
        const int kind = idx / (BN / 4), c = (idx - kind * (BN / 4)) * 4;
        if (n0 + c < p.Cout) {
          const int chunk = c >> 6;
          const int h = chunk % kHalves, ci = chunk / kHalves, lc = ci * 64 + (c & 63);
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(&s_part[((h * 4 + q) * 2 + kind) * kPartStride + lc]);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
          }
          red_add_v4(p.stats + kind * p.stat_c + (n0 + c) % p.stat_c, t.x, t.y, t.z, t.w);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    }
    if (kGroups == 2) {
      acc_phase ^= 1;            // this group always drains accumulator stage `group`
    } else if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
  if (cta_sums) {
    // one cross-lane reduction per CTA: this warp's totals -> its s_part slots
#pragma unroll
    for (int ci = 0; ci < kChunksPerWarp; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          float2 t1, t2;
          t1.x = __shfl_xor_sync(0xffffffffu, acc1[ci][j].x, o);
          t1.y = __shfl_xor_sync(0xffffffffu, acc1[ci][j].y, o);
          t2.x = __shfl_xor_sync(0xffffffffu, acc2[ci][j].x, o);
          t2.y = __shfl_xor_sync(0xffffffffu, acc2[ci][j].y, o);
          acc1[ci][j] = __fadd2_rn(acc1[ci][j], t1);
          acc2[ci][j] = __fadd2_rn(acc2[ci][j], t2);
        }
      }
      if (lane < 8) {
        float4* d1 = reinterpret_cast<float4*>(&s_part[(e * 2 + 0) * kPartStride + ci * 64 + lane * 8]);
        float4* d2 = reinterpret_cast<float4*>(&s_part[(e * 2 + 1) * kPartStride + ci * 64 + lane * 8]);
        d1[0] = make_float4(acc1[ci][0].x, acc1[ci][0].y, acc1[ci][1].x, acc1[ci][1].y);
        d1[1] = make_float4(acc1[ci][2].x, acc1[ci][2].y, acc1[ci][3].x, acc1[ci][3].y);
        d2[0] = make_float4(acc2[ci][0].x, acc2[ci][0].y, acc2[ci][1].x, acc2[ci][1].y);
        d2[1] = make_float4(acc2[ci][2].x, acc2[ci][2].y, acc2[ci][3].x, acc2[ci][3].y);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    for (int idx = et; idx < BN / 2; idx += 32 * kEpiWarps) {
      const int kind = idx / (BN / 4), c = (idx - kind * (BN / 4)) * 4;
      if (n0_cta + c < p.Cout) {
        const int chunk = c >> 6;
        const int h = chunk % kHalves, ci = chunk / kHalves, lc = ci * 64 + (c & 63);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4 * kGroups; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&s_part[((h * 4 + q) * 2 + kind) * kPartStride + lc]);
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        red_add_v4(p.stats + kind * p.stat_c + (n0_cta + c) % p.stat_c, t.x, t.y, t.z, t.w);
      }
    }
  }
  if (lane == 0) tma_store_wait_read<0>();
}

// Persistent, warp-specialised implicit GEMM.
//   grid  = min(#tiles, #SMs) CTAs, each walking tiles t = blockIdx.x, +gridDim.x, ...
//           (n-tile fastest so concurrently running CTAs share the A tile through L2)
//   warp0 : TMA producer            smem ring of STAGES x (A 128x64 + B BNx64), 128B swizzle
//   warp1 : tcgen05.mma issuer      2 TMEM accumulator stages of BN fp32 columns
//   warp2 : TMEM alloc / dealloc
//   warp4-11: epilogue (igemm_epilogue above)
// The epilogue of tile i overlaps the main loop of tile i+1.
//   PRO: warps 12-15 apply the producer BatchNorm (+ activation) to every A stage in place
//        between the TMA load and the MMA (BnPrologue above).
template <int BN, int STAGES, int SLABS, int AUX, bool PRO, int CG = 0>
__global__ void __launch_bounds__(igemm_threads(BN, PRO, CG), 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
             const __grid_constant__ CUtensorMap tmAux1, const __grid_constant__ CUtensorMap tmAux2,
             const IgemmParams p, const BnPrologue pro) {
  constexpr int kBBytes = BN * kBK * 2;
  constexpr uint32_t kTmemCols = 2 * BN;           // two accumulator stages (power of two)
  constexpr int kChunks = BN / 64;
  constexpr int kHalves = CG ? CG : (kChunks >= 2 ? 2 : 1);    // epilogue warp groups splitting the columns
  constexpr int kWarpsPerAcc = 4 * kHalves;        // warps that drain one accumulator stage
  constexpr int kEpiWarps = BN == 64 ? 8 : kWarpsPerAcc;   // BN = 64: two groups on alternate tiles
  constexpr int kChunksPerWarp = kChunks / kHalves;
  constexpr int kThreads_ = igemm_threads(BN, PRO, CG);
  // setmaxnreg rebalancing.  setmaxnreg.inc can only draw what setmaxnreg.dec released inside the
  // SAME CTA (the pool is the launch allocation, registers/thread x threads): the sum of the
  // increments must not exceed the sum of the decrements or the last warpgroup blocks forever.
  constexpr bool kRebalance = PRO || CG == 4;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t res_bar[16];
  __shared__ uint64_t ready_bar[PRO ? STAGES : 1];   // PRO: A stage transformed (4 warp arrivals)
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_part[kEpiWarps * 2 * kChunksPerWarp * 64];   // per-warp column sums of a tile
  __shared__ __align__(16) float s_ss[PRO ? 2 * kProMaxC : 4];   // PRO: scale | shift per input channel

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint8_t* smem_slab = smem_b + STAGES * kBBytes;   // [epilogue warps][SLABS][kSlabBytes]
  uint8_t* smem_aux = smem_slab + kEpiWarps * SLABS * kSlabBytes;   // [warps][AUX][kSlabBytes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if (PRO) mbar_init(&ready_bar[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kWarpsPerAcc);   // one arrival per epilogue warp of the stage
    }
    for (int s = 0; s < 16; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above touched no global data: overlap it with the previous kernel's tail (PDL)
  griddep_launch();
  griddep_wait();

  if (PRO) {
    // BatchNorm finalize folded in: every CTA derives the scale/shift table of the producer BN
    // from its raw batch sums; CTA 0 publishes what backward needs and the running statistics
    const int C = pro.Cin;
    for (int c = threadIdx.x; c < C; c += kThreads_) {
      float sc, sh;
      if (pro.stats != nullptr) {
        const BnCoeffs k = bn_coeffs(pro.stats[c], pro.stats[C + c], pro.gamma ? pro.gamma[c] : 1.f,
                                     pro.beta ? pro.beta[c] : 0.f, pro.count, pro.eps);
        sc = k.scale;
        sh = k.shift;
        if (blockIdx.x == 0) {
          pro.mean_invstd[c] = k.mean;
          pro.mean_invstd[C + c] = k.invstd;
          pro.scale_shift[c] = sc;
          pro.scale_shift[C + c] = sh;
          if (pro.running_mean != nullptr) {
            pro.running_mean[c] = bn_running(pro.running_mean[c], k.mean, pro.momentum);
            pro.running_var[c] = bn_running(pro.running_var[c], bn_unbiased(k.var, pro.count), pro.momentum);
          }
        }
      } else {
        sc = pro.scale_shift[c];
        sh = pro.scale_shift[C + c];
      }
      s_ss[c] = sc;
      s_ss[kProMaxC + c] = sh;
    }
    __syncthreads();
  }

  // (setmaxnreg sits at the top of each role branch so that it dominates the role's code)
  // register budgets (x 128 threads per warpgroup, <= 64 Ki):
  //   PRO, 8 epilogue warps  (512 thr, 128 at launch): control 64, epilogue 2 x 168, transform 104
  //                           released 8192 + 3072 >= drawn 2 x 5120
  //   16 epilogue warps      (640 thr,  96 at launch): control 32, epilogue 4 x 112
  //                           released 8192 >= drawn 4 x 2048
  //   (PRO + 16 epilogue warps would need 768 threads at 80 registers: not instantiated)
  if (PRO && warp >= 4 + kEpiWarps) {
    if (CG != 4) reg_dec<104>();
    // ---- A-operand transform: a = act(fmaf(c, scale, shift)), padding taps stay zero ----
    // thread t owns the logical 16-byte chunk (8 channels) t % 8 of rows t / 8 + 16 i: its
    // coefficients stay in registers for a whole k-block and a warp touches four whole
    // 128-byte rows per access (conflict-free under the 128B swizzle)
    const int tt = threadIdx.x - (128 + 32 * kEpiWarps);
    const int lc = tt & 7, r0 = tt >> 3;
    const uint32_t off0 = (uint32_t)r0 * 128u + (uint32_t)((lc ^ (r0 & 7)) << 4);
    const int act = pro.act;
    const float slope = pro.slope;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * kBM;
      int bh[8], bw[8];         // base pixel (tap 0) of the thread's eight rows
      if (!p.tiled_a) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + r0 + 16 * i;
          const int n_img = m / p.trav_hw;
          const int rem = m - n_img * p.trav_hw;
          const int pp = rem / p.trav_w;
          bh[i] = pp * p.stride - p.pad_h;
          bw[i] = (rem - pp * p.trav_w) * p.stride - p.pad_w;
        }
      }
      int tap = 0, cb = 0;
      for (int kb = 0; kb < p.num_kblocks; ++kb) {
        float sc[8], sh[8];
        {
          const float4* t4 = reinterpret_cast<const float4*>(&s_ss[cb * kBK + lc * 8]);
          const float4* u4 = reinterpret_cast<const float4*>(&s_ss[kProMaxC + cb * kBK + lc * 8]);
          const float4 a0 = t4[0], a1 = t4[1], b0 = u4[0], b1 = u4[1];
          sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
          sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
        }
        const int tr = tap / p.S;
        const int ts = tap - tr * p.S;
        uint32_t inside = 0xffu;      // bit i: row i of this thread is a real pixel for this tap
        if (!p.tiled_a) {
          inside = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            inside |= ((unsigned)(bh[i] + tr) < (unsigned)pro.IH && (unsigned)(bw[i] + ts) < (unsigned)pro.IW)
                          ? (1u << i) : 0u;
        }
        mbar_wait(&full_bar[stage], phase);
        uint8_t* base = smem_a + stage * kABytes + off0;
        // all eight loads first (independent), then the arithmetic: one exposed LDS latency
        uint4 raw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = *reinterpret_cast<const uint4*>(base + i * 2048);
        if (act == SIB_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw[i]);
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x2 = make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xffff0000u));
              const float2 v = __ffma2_rn(x2, make_float2(sc[2 * j], sc[2 * j + 1]),
                                          make_float2(sh[2 * j], sh[2 * j + 1]));
              ow[j] = pack2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
            }
            if (!((inside >> i) & 1u)) o = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(base + i * 2048) = o;
          }
        } else {
          const float neg = act == SIB_ACT_LEAKY ? slope : 1.f;     // identity: v * 1
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw[i]);
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x2 = make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xffff0000u));
              const float2 v = __ffma2_rn(x2, make_float2(sc[2 * j], sc[2 * j + 1]),
                                          make_float2(sh[2 * j], sh[2 * j + 1]));
              ow[j] = pack2(v.x > 0.f ? v.x : v.x * neg, v.y > 0.f ? v.y : v.y * neg);
            }
            if (!((inside >> i) & 1u)) o = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(base + i * 2048) = o;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[stage]);
        if (++cb == p.cin_blocks) { cb = 0; ++tap; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    if (kRebalance) { if (CG == 4) reg_dec<32>(); else reg_dec<64>(); }
    if (lane == 0) {
      // ---- TMA producer ----
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * kBM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        int n_img = 0, w_base = 0, h_base = 0;
        if (!p.tiled_a) {
          n_img = m0 / p.trav_hw;
          const int rem = m0 - n_img * p.trav_hw;
          const int pp = rem / p.trav_w;
          const int qq = rem - pp * p.trav_w;
          w_base = qq * p.stride - p.pad_w;
          h_base = pp * p.stride - p.pad_h;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
          if (p.tiled_a) {
            tma_load_2d(smem_a + stage * kABytes, &tmA, &full_bar[stage], kb * kBK, m0);
          } else {
            const int r = tap / p.S;
            const int s = tap - r * p.S;
            tma_load_im2col_4d(smem_a + stage * kABytes, &tmA, &full_bar[stage], cb * kBK,
                               w_base, h_base, n_img, (uint16_t)s, (uint16_t)r);
          }
          tma_load_2d(smem_b + stage * kBBytes, &tmB, &full_bar[stage], kb * kBK, n0);
          if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the whole warp runs the loop converged, one elected lane issues ----
    // (Two issuing warps on alternate tiles, as in the halo kernel, were tried for BN = 64 and
    //  REMOVED: with a shared k-block ring a warp can wait on lap L of a stage whose lap L-1 --
    //  owned by the other warp -- has not completed yet; the parity test then passes at once and
    //  the MMAs read data still in flight (launch failure at full size).  The halo kernel's ring
    //  holds one whole tile per stage and both of its issuing warps observe every tile's barrier.)
    if (kRebalance) { if (CG == 4) reg_dec<32>(); else reg_dec<64>(); }
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), 16, 1024, kSwizzle128B);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), 16, 1024, kSwizzle128B);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_w(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kblocks; ++kb) {
        mbar_wait_w(&full_bar[stage], phase);
        if (PRO) mbar_wait_w(&ready_bar[stage], phase);   // A stage rewritten by the transform warps
        tc_fence_after();
        const uint64_t a_desc = a_desc0 + (uint32_t)(stage * (kABytes >> 4));
        const uint64_t b_desc = b_desc0 + (uint32_t)(stage * (kBBytes >> 4));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // +32 bytes along K inside the 128B swizzle row = +2 in 16-byte address units
          umma_bf16_ss_w(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit_w(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_w(&tmem_full_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    static_assert(!(PRO && CG == 4), "PRO + 16-warp epilogue is not instantiated (register budget)");
    if (kRebalance) { if (CG == 4) reg_inc<112>(); else reg_inc<168>(); }
    igemm_epilogue<BN, SLABS, AUX, false, CG>(&tmOut, &tmRes, &tmAux1, &tmAux2, p, smem_slab, smem_aux,
                                              tmem_full_bar, tmem_empty_bar, res_bar, s_part, tmem_base,
                                              blockIdx.x, gridDim.x, num_tiles, 0);
  } else if (kRebalance) {
    // warps 2 and 3: the rest of warpgroup 0 must execute the same setmaxnreg
    if (CG == 4) reg_dec<32>(); else reg_dec<64>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// 2-CTA variant of the persistent implicit GEMM (cta_group::2): a cluster of two CTAs owns a
// 256 x 256 output tile.  Each CTA loads its own 128 activation rows and HALF of the weight tile
// (128 of the 256 output channels); the leader CTA (cluster rank 0) issues tcgen05.mma with
// M = 256, which reads both CTAs' shared memory and writes 128 accumulator rows into each CTA's
// TMEM.  Per-SM operand traffic per FLOP drops by 1/3 against the 1-CTA 128 x 256 tile.
// TMA loads of both CTAs complete on the leader's `full` barriers; tcgen05.commit multicasts to
// both CTAs' `empty` / `tmem_full` barriers; both epilogues release the accumulator on the
// leader's `tmem_empty` barrier.  Everything else is the 1-CTA kernel.
template <int BN, int STAGES, int SLABS, int AUX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kIgemmThreads, 1)
igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
              const __grid_constant__ CUtensorMap tmAux1, const __grid_constant__ CUtensorMap tmAux2,
              const IgemmParams p) {
  constexpr int kBBytes = (BN / 2) * kBK * 2;      // this CTA's half of the weight tile
  constexpr uint32_t kTmemCols = 2 * BN;           // two accumulator stages (power of two)
  constexpr int kEpiWarps = 8;
  constexpr int kChunksPerWarp = BN / 128;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t res_bar[8];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_part[kEpiWarps * 2 * kChunksPerWarp * 64];   // per-warp column sums of a tile

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint8_t* smem_slab = smem_b + STAGES * kBBytes;   // [epilogue warps][SLABS][kSlabBytes]
  uint8_t* smem_aux = smem_slab + kEpiWarps * SLABS * kSlabBytes;   // [warps][AUX][kSlabBytes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = (p.num_m_tiles >> 1) * p.num_n_tiles;     // 256-row pair tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps);   // epilogue warps of BOTH CTAs (leader's copy is used)
    }
    for (int s = 0; s < 8; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(&tmem_base_smem, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / p.num_n_tiles) * 2 * kBM + (int)cta_rank * kBM;
        const int n0 = (tile % p.num_n_tiles) * BN + (int)cta_rank * (BN / 2);   // this CTA's weight half
        int n_img = 0, w_base = 0, h_base = 0;
        if (!p.tiled_a) {
          n_img = m0 / p.trav_hw;
          const int rem = m0 - n_img * p.trav_hw;
          const int pp = rem / p.trav_w;
          const int qq = rem - pp * p.trav_w;
          w_base = qq * p.stride - p.pad_w;
          h_base = pp * p.stride - p.pad_h;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
          if (p.tiled_a) {
            tma2_load_2d(smem_a + stage * kABytes, &tmA, &full_bar[stage], kb * kBK, m0);
          } else {
            const int r = tap / p.S;
            const int s = tap - r * p.S;
            tma2_load_im2col_4d(smem_a + stage * kABytes, &tmA, &full_bar[stage], cb * kBK,
                                w_base, h_base, n_img, (uint16_t)s, (uint16_t)r);
          }
          tma2_load_2d(smem_b + stage * kBBytes, &tmB, &full_bar[stage], kb * kBK, n0);
          if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---- MMA issuer (leader CTA only): whole warp converged, one elected lane issues ----
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, BN, 0, 0);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), 16, 1024, kSwizzle128B);
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), 16, 1024, kSwizzle128B);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait_w(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait_w(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint32_t)(stage * (kABytes >> 4));
          const uint64_t b_desc = b_desc0 + (uint32_t)(stage * (kBBytes >> 4));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes along K inside the 128B swizzle row = +2 in 16-byte address units
            umma2_bf16_ss_w(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma2_commit_multicast_w(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma2_commit_multicast_w(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    igemm_epilogue<BN, SLABS, AUX, true>(&tmOut, &tmRes, &tmAux1, &tmAux2, p, smem_slab, smem_aux,
                                         tmem_full_bar, tmem_empty_bar, res_bar, s_part, tmem_base,
                                         cluster_id, num_clusters, num_tiles, (int)cta_rank);
  }
  tc_fence_before();
  cluster_sync_all();          // the leader's MMAs read the peer's smem: nobody leaves early
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 with shared-memory halo reuse
// ----------------------------------------------------------------------------
// The im2col kernel above fetches every input pixel nine times from L2 (once per filter tap), and
// the 64- and 128-channel 3x3 layers at 56x56 / 28x28 are bound by exactly that L2 -> SM traffic
// (profiles/r01_ncu_conv_layers_metrics.txt: 8-10 TB/s against a ~10.5 TB/s fabric ceiling, tensor
// pipe 17-32 %).  Here a CTA loads, per 64 input channels, ONE zero-padded pixel region
// [TH+2 rows][W+2 columns] (two im2col TMA loads over a bounding box one pixel larger than the
// image on every side; the border is zero-filled), and every tap is the same region shifted by (r * (W+2) + s) rows of 128 bytes: the 128B swizzle
// is a function of the absolute shared-memory address, so a K-major UMMA descriptor may start at
// any 128-byte row (scripts/probes/umma_shift_probe.cu, measured exact for arbitrary shifts).
// GEMM row j of a tile is the padded position (j / (W+2), j % (W+2)); positions with column >= W
// are junk rows that the epilogue drops (W = 56: 112 of 128 rows useful).  With 64 -> 64 channels
// the nine weight k-blocks (72 KB) stay resident in shared memory for the whole kernel.
struct HaloParams {
  int N, H, W, Cin, Cout;
  int pitch;          // W + 2
  int tile_h;         // output rows per tile: floor(128 / pitch)
  int num_h_tiles;    // ceil(H / tile_h)
  int num_n_tiles;
  int cin_blocks;
  int b_stationary;   // all weight k-blocks fit the ring: load them once
  float* stats;       // [2][Cout] (fprop: sum, sumsq; fused BN backward: sum g, sum g*xhat) or null
  int fuse;           // 1: fused BN-backward reduction (see IgemmParams)
  int fuse_act;
  float fuse_slope;
  const __nv_bfloat16* aux;    // [N][H][W][Cout] mask / xhat source
  const float* mask_ss;
  const float* mean_invstd;
};

// A stage = kALoads im2col loads of 128 padded positions x 128 bytes.  The nine taps read 128 rows
// starting at up to 2 * pitch + 2, i.e. 2 * pitch + 130 rows in all: two loads (256 rows) serve
// pitch <= 63 (ResNet's 56x56 / 28x28 layers), three (384 rows) pitch <= 127 (the 112x112 layers of
// the BResNet deep stem, one output row per tile).
template <int BN, int BSTAGES, int kHaloAStages, int kALoads = 2>
__global__ void __launch_bounds__(kIgemmThreads, 1)
halo3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               __nv_bfloat16* __restrict__ out, const HaloParams p) {
  constexpr int kBBytes = BN * kBK * 2;
  constexpr int kHaloAStage = kALoads * kABytes;
  constexpr uint32_t kTmemCols = 2 * BN;
  constexpr int kGroups = BN == 64 ? 2 : 1;     // epilogue warp groups working on alternate tiles
  constexpr int kEpiWarps = 8;
  constexpr int kWarpsPerAcc = 8 / kGroups;     // warps that drain one accumulator stage
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[kHaloAStages], a_empty[kHaloAStages];
  __shared__ uint64_t b_full[BSTAGES], b_empty[BSTAGES];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kHaloAStages * kHaloAStage;
  uint8_t* smem_slab = smem_b + BSTAGES * kBBytes;    // [epilogue warps][kSlabBytes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.N * p.num_h_tiles * p.num_n_tiles;
  const int num_kblocks = 9 * p.cin_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kHaloAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], kWarpsPerAcc); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: one region per (tile, 64 input channels), one weight k-block per tap ----
      int a_stage = 0, b_stage = 0;
      uint32_t a_phase = 0, b_phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles;
        const int n0 = (tile - mt * p.num_n_tiles) * BN;
        const int n_img = mt / p.num_h_tiles;
        const int h0 = (mt - n_img * p.num_h_tiles) * p.tile_h;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait(&a_empty[a_stage], a_phase ^ 1);
          // The padded region is fetched as kALoads im2col loads of 128 consecutive padded
          // positions (bounding box = image + 1 pixel of zero-filled border on every side, so the
          // traversal pitch is W + 2).  The last load may run past the region; those rows only
          // feed junk outputs.  (A tiled 4-D box of the same region measured 7x slower per row.)
          int hj[kALoads], wj[kALoads], nj[kALoads], n_ok = 1;
#pragma unroll
          for (int j = 1; j < kALoads; ++j) {
            hj[j] = h0 - 1 + (128 * j) / p.pitch;
            wj[j] = -1 + (128 * j) % p.pitch;
            nj[j] = n_img;
            if (hj[j] > p.H) { hj[j] -= p.H + 2; ++nj[j]; }
            if (nj[j] < p.N) ++n_ok;
          }
          mbar_arrive_expect_tx(&a_full[a_stage], n_ok * kABytes);
          uint8_t* dst = smem_a + a_stage * kHaloAStage;
          tma_load_im2col_4d(dst, &tmA, &a_full[a_stage], cb * kBK, -1, h0 - 1, n_img, 0, 0);
#pragma unroll
          for (int j = 1; j < kALoads; ++j)
            if (nj[j] < p.N)
              tma_load_im2col_4d(dst + j * kABytes, &tmA, &a_full[a_stage], cb * kBK, wj[j], hj[j], nj[j], 0, 0);
          if (++a_stage == kHaloAStages) { a_stage = 0; a_phase ^= 1; }
          if (!p.b_stationary || first) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[b_stage], b_phase ^ 1);
              mbar_arrive_expect_tx(&b_full[b_stage], kBBytes);
              tma_load_2d(smem_b + b_stage * kBBytes, &tmB, &b_full[b_stage],
                          tap * p.Cin + cb * kBK, n0);
              if (++b_stage == BSTAGES) { b_stage = 0; b_phase ^= 1; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1 || (warp == 3 && p.b_stationary && p.cin_blocks == 1)) {
    // ---- MMA issuer: the whole warp runs the loop converged, one elected lane issues ----
    // With resident weights and one channel block per tile the issue loop itself bounds the tile
    // rate (36 MMAs of 32 clk each), so TWO warps (1 and 3) issue alternate tiles, each into its
    // own TMEM accumulator stage.
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
    // tap (r, s) = the region shifted by r * pitch + s pixel rows of 128 bytes (16-byte units)
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = ((tap / 3) * p.pitch + (tap % 3)) * 8;
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), 16, 1024, kSwizzle128B);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), 16, 1024, kSwizzle128B);
    if (p.b_stationary && p.cin_blocks == 1) {
      const int g = warp == 3 ? 1 : 0;
      // the weight k-blocks are loaded once and never released
      for (int kb = 0; kb < 9; ++kb) mbar_wait_w(&b_full[kb], 0);
      tc_fence_after();
      int i = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
        const int st = i % kHaloAStages;
        const uint32_t ph = (uint32_t)(i / kHaloAStages) & 1u;
        if ((i & 1) != g) {
          // not this warp's tile, but OBSERVE its barrier: a parity wait can only be trusted by
          // a waiter that has seen every earlier lap of the stage complete (otherwise "previous
          // lap still in flight" looks like "this lap done")
          mbar_wait_w(&a_full[st], ph);
          continue;
        }
        const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
        mbar_wait_w(&tmem_empty_bar[g], acc_ph ^ 1);
        tc_fence_after();
        mbar_wait_w(&a_full[st], ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + g * BN;
        const uint64_t a_desc = a_desc0 + (uint32_t)(st * (kHaloAStage >> 4));
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t ad = a_desc + tap_off[tap];
          const uint64_t bd = b_desc0 + (uint32_t)(tap * (kBBytes >> 4));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_ss_w(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0);
        }
        umma_commit_w(&a_empty[st]);
        umma_commit_w(&tmem_full_bar[g]);
      }
    } else {
      int a_stage = 0, b_stage = 0;
      uint32_t a_phase = 0, b_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.b_stationary) {
        for (int kb = 0; kb < num_kblocks; ++kb) mbar_wait_w(&b_full[kb], 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait_w(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait_w(&a_full[a_stage], a_phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint32_t)(a_stage * (kHaloAStage >> 4));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (!p.b_stationary) {
              mbar_wait_w(&b_full[b_stage], b_phase);
              tc_fence_after();
            }
            const uint64_t ad = a_desc + tap_off[tap];
            const uint64_t bd =
                b_desc0 + (uint32_t)((p.b_stationary ? cb * 9 + tap : b_stage) * (kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_ss_w(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (cb | tap | k) != 0);
            if (!p.b_stationary) {
              umma_commit_w(&b_empty[b_stage]);
              if (++b_stage == BSTAGES) { b_stage = 0; b_phase ^= 1; }
            }
          }
          umma_commit_w(&a_empty[a_stage]);
          if (++a_stage == kHaloAStages) { a_stage = 0; a_phase ^= 1; }
        }
        umma_commit_w(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ---- epilogue: TMEM -> bf16 slab -> column-owner pass (statistics / fused BN backward,
    //      junk-row removal) -> coalesced 128-bit global stores.
    // BN = 64: two groups of four warps take alternate tiles (group g owns accumulator stage g),
    // so two epilogues are in flight and their latency no longer bounds the tile rate.
    // BN = 128: one group of eight warps, warp e handles 64-column chunk e / 4.
    // The per-channel sums stay in registers across all tiles of the CTA (one n-tile per launch)
    // and are published once at the end.
    const int e = warp - 4;
    const int quarter = e & 3;
    const int group = kGroups == 2 ? (e >> 2) : 0;
    const int chunk = kGroups == 2 ? 0 : (e >> 2);
    uint8_t* slab = smem_slab + e * kSlabBytes;
    uint32_t row_slot[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) row_slot[g] = lane * 128 + ((g ^ (lane & 7)) << 4);
    const int st_row = lane >> 3, st_slot = lane & 7;
    const uint32_t st_off0 = st_row * 128 + ((st_slot ^ st_row) << 4);
    const uint32_t st_off1 = st_row * 128 + ((st_slot ^ (st_row + 4)) << 4);
    // padded position of the 8 rows this lane handles in the column-owner pass
    int rel_m[8], rel_h[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int j = quarter * 32 + it * 4 + st_row;
      const int hh = j / p.pitch;
      const int ww = j - hh * p.pitch;
      rel_h[it] = hh;
      rel_m[it] = (ww < p.W && hh < p.tile_h) ? hh * p.W + ww : -1;
    }
    const bool fused = p.fuse != 0;
    const int cbase = chunk * 64 + st_slot * 8;       // (num_n_tiles == 1: n0 = 0)
    const bool col_ok = cbase < p.Cout;
    float sc[8], sh[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; mu[j] = 0.f; is[j] = 0.f; }
    if (fused && col_ok) {
      if (p.mask_ss != nullptr) {
        load8f(p.mask_ss + cbase, sc);
        load8f(p.mask_ss + p.Cout + cbase, sh);
      }
      load8f(p.mean_invstd + cbase, mu);
      load8f(p.mean_invstd + p.Cout + cbase, is);
    }
    const bool do_mask = p.fuse_act != SIB_ACT_NONE;
    const float neg = p.fuse_act == SIB_ACT_LEAKY ? p.fuse_slope : 0.f;
    float2 cs1[4], cs2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { cs1[j] = make_float2(0.f, 0.f); cs2[j] = make_float2(0.f, 0.f); }
    const int acc = group;                    // kGroups == 1: alternates below
    int acc1 = 0;
    uint32_t acc_phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      if (kGroups == 2 && (local & 1) != group) continue;
      const int a = kGroups == 2 ? acc : acc1;
      const int mt = tile / p.num_n_tiles;
      const int n_img = mt / p.num_h_tiles;
      const int h0 = (mt - n_img * p.num_h_tiles) * p.tile_h;
      const int rows_here = p.H - h0;                   // rows of this tile that exist
      const size_t m_base = ((size_t)n_img * p.H + h0) * p.W;
      mbar_wait(&tmem_full_bar[a], acc_phase);
      tc_fence_after();
      uint32_t r[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + chunk * 64;
      tmem_ld_32x32b_x32(taddr, r);
      tmem_ld_32x32b_x32(taddr + 32, r + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[a]);
      {
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(slab + row_slot[g]) = pack8(&v[g * 8]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const bool ok = rel_m[it] >= 0 && rel_h[it] < rows_here && col_ok;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (ok) q = *reinterpret_cast<const uint4*>(slab + it * 512 + ((it & 1) ? st_off1 : st_off0));
        const size_t o = (m_base + (ok ? rel_m[it] : 0)) * p.Cout + cbase;
        if (fused) {
          float g[8], xv[8];
          unpack8(q, g);
          uint4 xa = make_uint4(0, 0, 0, 0);
          if (ok) xa = __ldg(reinterpret_cast<const uint4*>(p.aux + o));
          unpack8(xa, xv);
          if (do_mask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float z = fmaf(xv[j], sc[j], sh[j]);
              g[j] = z > 0.f ? g[j] : g[j] * neg;
            }
            q = pack8(g);
          }
          if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 gg = make_float2(g[2 * j], g[2 * j + 1]);
              const float2 xh = make_float2((xv[2 * j] - mu[2 * j]) * is[2 * j],
                                            (xv[2 * j + 1] - mu[2 * j + 1]) * is[2 * j + 1]);
              cs1[j] = __fadd2_rn(cs1[j], gg);
              cs2[j] = __ffma2_rn(gg, xh, cs2[j]);
            }
          }
        } else if (p.stats != nullptr) {
          const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(hq[j]);
            cs1[j] = __fadd2_rn(cs1[j], f);
            cs2[j] = __ffma2_rn(f, f, cs2[j]);
          }
        }
        if (ok) *reinterpret_cast<uint4*>(out + o) = q;
      }
      __syncwarp();     // the slab is rewritten by this warp's next tile
      if (kGroups == 2) {
        acc_phase ^= 1;
      } else if (++acc1 == 2) {
        acc1 = 0;
        acc_phase ^= 1;
      }
    }
    if (p.stats != nullptr) {
      // lanes with equal lane%8 hold partial sums of the same 8 columns (different rows)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          float2 t1, t2;
          t1.x = __shfl_xor_sync(0xffffffffu, cs1[j].x, o);
          t1.y = __shfl_xor_sync(0xffffffffu, cs1[j].y, o);
          t2.x = __shfl_xor_sync(0xffffffffu, cs2[j].x, o);
          t2.y = __shfl_xor_sync(0xffffffffu, cs2[j].y, o);
          cs1[j] = __fadd2_rn(cs1[j], t1);
          cs2[j] = __fadd2_rn(cs2[j], t2);
        }
      }
      // combine the eight warps in shared memory (the drained slabs), then ONE atomic per column
      // and CTA: same-address atomics from all CTAs at once otherwise pile up at the kernel's end
      float* part = reinterpret_cast<float*>(slab);     // this warp's own (drained) slab: [2][64]
      if (lane < 8) {
        float4* d1 = reinterpret_cast<float4*>(&part[lane * 8]);
        float4* d2 = reinterpret_cast<float4*>(&part[64 + lane * 8]);
        d1[0] = make_float4(cs1[0].x, cs1[0].y, cs1[1].x, cs1[1].y);
        d1[1] = make_float4(cs1[2].x, cs1[2].y, cs1[3].x, cs1[3].y);
        d2[0] = make_float4(cs2[0].x, cs2[0].y, cs2[1].x, cs2[1].y);
        d2[1] = make_float4(cs2[2].x, cs2[2].y, cs2[3].x, cs2[3].y);
      }
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 128;
      for (int idx = et; idx < BN / 2; idx += 256) {
        const int kind = idx / (BN / 4), c = (idx - kind * (BN / 4)) * 4;
        const int ch = c >> 6, lc = c & 63;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w)
          if (kGroups == 2 || (w >> 2) == ch) {
            const float4 v = *reinterpret_cast<const float4*>(
                reinterpret_cast<const float*>(smem_slab + w * kSlabBytes) + kind * 64 + lc);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
          }
        if (c < p.Cout) red_add_v4(p.stats + kind * p.Cout + c, t.x, t.y, t.z, t.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// wgrad
// ----------------------------------------------------------------------------
struct WgradParams {
  int M_total;       // pixels (reduction length)
  int Cout, Cin;
  int S;
  int trav_hw, trav_w;
  int stride, pad_h, pad_w;
  int kblocks_per_split;   // 64-pixel blocks handled by one CTA
  int total_kblocks;
  int ldw;           // dW row pitch = R*S*Cin
  int cin_blocks;    // Cin / 64
  int total_units;   // R*S*Cin/64 : 64-column units of the [Cout][R*S*Cin] gradient matrix
};

constexpr int kWgPix = 64;   // pixels per k-block
constexpr int kWgUnits = 4;  // 64-column units per CTA -> 128 x 256 output tile

// dW tile [128 out channels][4 units x 64 columns]; a unit is one (filter tap, 64 input
// channels) pair, i.e. 64 consecutive columns of the [Cout][R*S*Cin] matrix.  All units of a
// CTA share the dY operand; each unit's X operand is its own im2col TMA load.
// Epilogue (kTmaReduce): TMEM -> registers -> 128B-swizzled fp32 slabs (32 rows x 32 columns, carved
// out of the drained operand ring) -> TMA add-reduction into dW, so the split-K accumulation is
// done by the TMA unit / L2 instead of 8192 per-lane RED.v4 per tile.
// CL > 1: the CL CTAs of a cluster (consecutive out-channel tiles, same units, same pixel
// range) need the same X tiles; each loads 1/CL of the units and multicasts them to all, so the
// L2 -> SM traffic of the X operand drops by CL (the wgrad kernels sit at the L2 -> SM fabric
// ceiling).  A stage is refilled only when the MMAs of ALL CTAs of the cluster have retired it.
template <int STAGES, bool kTmaReduce, int CL>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ CUtensorMap tmDw, float* __restrict__ dw,
             const WgradParams p) {
  constexpr int BNC = 64 * kWgUnits;
  constexpr int kUnitBytes = kWgPix * 128;     // 64 pixels x 64 channels bf16
  constexpr int kABytesW = 2 * kUnitBytes;     // dY tile: 64 pixels x 128 out channels
  constexpr int kBBytesW = kWgUnits * kUnitBytes;
  constexpr uint32_t kTmemCols = BNC;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytesW;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kBM;                 // out-channel tile
  const int u0 = blockIdx.y * kWgUnits;            // first 64-column unit
  int nunits = p.total_units - u0;
  if (nunits > kWgUnits) nunits = kWgUnits;
  const int kb_begin = blockIdx.z * p.kblocks_per_split;
  int kb_end = kb_begin + p.kblocks_per_split;
  if (kb_end > p.total_kblocks) kb_end = p.total_kblocks;
  const int nkb = kb_end - kb_begin;
  const int a_boxes = (p.Cout - k0 > 64) ? 2 : 1;  // second 64-channel half may not exist

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    if (kTmaReduce) tma_prefetch_desc(&tmDw);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);       // one commit per CTA of the cluster
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int cl_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1);
  // everything above touched no global data: overlap it with the previous kernel's tail (PDL)
  griddep_launch();
  griddep_wait();
  if (nkb <= 0) {
    // nothing to do for this split (uniform across the CTA); fall through to teardown
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int m0 = kb * kWgPix;
        const int n_img = m0 / p.trav_hw;
        const int rem = m0 - n_img * p.trav_hw;
        const int pp = rem / p.trav_w;
        const int qq = rem - pp * p.trav_w;
        const int wb = qq * p.stride - p.pad_w, hb = pp * p.stride - p.pad_h;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], (a_boxes + nunits) * kUnitBytes);
        uint8_t* a_dst = smem_a + stage * kABytesW;
        for (int h = 0; h < a_boxes; ++h)
          tma_load_2d(a_dst + h * kUnitBytes, &tmDy, &full_bar[stage], k0 + h * 64, m0);
        uint8_t* b_dst = smem_b + stage * kBBytesW;
        for (int h = 0; h < nunits; ++h) {
          const int u = u0 + h;
          const int tap = u / p.cin_blocks;
          const int c0 = (u - tap * p.cin_blocks) * 64;
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          if (CL == 1)
            tma_load_im2col_4d(b_dst + h * kUnitBytes, &tmX, &full_bar[stage], c0, wb, hb, n_img,
                               (uint16_t)s, (uint16_t)r);
          else if (h % CL == cl_rank)
            tma_load_im2col_4d_mc(b_dst + h * kUnitBytes, &tmX, &full_bar[stage], c0, wb, hb, n_img,
                                  (uint16_t)s, (uint16_t)r, kClMask);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // both operands MN-major: 64-element groups along M/N are kUnitBytes apart (LBO),
    // 8-pixel groups along K are 1024 bytes apart (SBO).  Whole warp converged, one elected
    // lane issues.
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BNC, 1, 1);
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), kUnitBytes, 1024, kSwizzle128B);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), kUnitBytes, 1024, kSwizzle128B);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait_w(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t a_desc = a_desc0 + (uint32_t)(stage * (kABytesW >> 4));
      const uint64_t b_desc = b_desc0 + (uint32_t)(stage * (kBBytesW >> 4));
#pragma unroll
      for (int k = 0; k < kWgPix / 16; ++k) {
        // 16 pixels along K = two 1024-byte groups = +128 in 16-byte address units
        umma_bf16_ss_w(tmem_base, a_desc + 128 * k, b_desc + 128 * k, idesc, (i | k) != 0);
      }
      if (CL == 1) umma_commit_w(&empty_bar[stage]);
      else umma_commit_mc_w(&empty_bar[stage], kClMask);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_w(&tmem_full_bar);
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;   // out channel within the tile
    const int k = k0 + row;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int ncols = nunits * 64;
    if (kTmaReduce) {
      // all MMAs have completed (tmem_full), hence the operand ring is free: 4 slabs per warp
      constexpr int kRedSlabs = 4;
      static_assert(STAGES * (kABytesW + kBBytesW) >= 4 * kRedSlabs * kSlabBytes, "ring too small");
      uint8_t* slabs = smem + (warp - 2) * kRedSlabs * kSlabBytes;
      if (k0 + quarter * 32 < p.Cout) {       // warp-uniform: rows past Cout would be clipped anyway
        // (chunk order rotated by the split index: the splits of one gradient tile finish together and
        //  would otherwise add into the same 4 KB at the same time)
        const int nchunks = ncols / 32;
#pragma unroll 1
        for (int ci = 0; ci < nchunks; ++ci) {
          const int chunk = (ci + (int)blockIdx.z) % nchunks;
          uint8_t* slab = slabs + (ci % kRedSlabs) * kSlabBytes;
          if (lane == 0) tma_store_wait_read<kRedSlabs - 1>();
          __syncwarp();
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + chunk * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(slab + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmDw, slab, u0 * 64 + chunk * 32, k0 + quarter * 32);
            tma_store_commit();
          }
        }
        if (lane == 0) tma_store_wait<0>();
      }
    } else {
      float* drow = dw + (long)k * p.ldw + (long)u0 * 64;
#pragma unroll 1
      for (int chunk = 0; chunk < BNC / 32; ++chunk) {
        if (chunk * 32 >= ncols) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + chunk * 32, r);
        tmem_ld_wait();
        if (k < p.Cout) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float* d = drow + chunk * 32 + g * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d),
                         "f"(__uint_as_float(r[g * 4])), "f"(__uint_as_float(r[g * 4 + 1])),
                         "f"(__uint_as_float(r[g * 4 + 2])), "f"(__uint_as_float(r[g * 4 + 3]))
                         : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  // (cluster: peers still arrive on this CTA's barriers until their last commit has landed)
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// wgrad of 3x3 / stride 1 / pad 1 filters with 64 output channels: halo reuse on BOTH operands
// ----------------------------------------------------------------------------
// The generic kernel above reads X once per filter tap (nine im2col loads per pixel block) and, with
// 64 output channels, fills only half of its M = 128 tile: ncu shows the 64 -> 64 3x3 layers pinned
// at the L2 -> SM fabric (1.2 GB moved for 0.2 GB of operands, 0.147 ms against 0.043 ms ideal).
// Here the reduction runs over PADDED flat positions J of the (H+2) x (W+2) grid of every image
// (pitch = W+2; both operands are zero on the border, fetched by im2col TMA loads over a bounding box
// one pixel larger than the image, exactly like the halo fprop kernel):
//     dW[k][(r,s)][c] = sum_J  X_pad[J + (r-1)*pitch][c] * dY_pad[J + 1 - s][k]
// so every tap is a pair of ROW-SHIFTED VIEWS of two shared-memory regions that are loaded once per
// 128 positions.  Both operands are MN-major (a pixel is a 128-byte row, the channels are the M / N
// index, the pixels the K index), and for MN-major 128B-swizzled descriptors the start address may
// be advanced by any number of rows AND the leading byte offset (distance between the 64-channel
// groups of an operand) may be any multiple of 128 bytes (scripts/probes/umma_mn_major_shift_probe.cu:
// exact for all shifts / offsets) -- the 64-channel groups of one MMA operand can be shifted views:
//     B (N = 192) = X region at rows +0, +pitch, +2*pitch        (r = 0, 1, 2; LBO = pitch rows)
//     A (M = 128) = dY region at rows +0, +1                      (s = 1, 0;    LBO = 1 row)
// One MMA (128 x 192 x 16) therefore produces SIX taps, a second one with A = dY at row -1 (s = 2; its
// upper half is a don't-care duplicate) the remaining three: two MMAs per 16 positions instead of nine
// N = 64 ones.  The kernel is bound by the L2 -> SM fabric, so the tile is 240 positions (15 k-steps):
// the regions (240 + 2 * pitch rows of X, 242 rows of dY) then cost 5 loads of 128 positions = 341
// bytes per position (pitch <= 72), against 512 with 128-position tiles and 1.5 KB in the generic
// kernel.  Split over position ranges (one CTA per SM), TMA add-reduction of the 9 x 64 x 64 partials.
struct WgradHaloParams {
  int N, Cin;
  int pitch;          // W + 2
  int P;              // (H + 2) * pitch: padded positions per image
  long total_pos;     // N * P
  int total_tiles;    // tiles of kWhTile positions; tile t covers J in [pitch + kWhTile t, +kWhTile)
  int tiles_per_cta;
};

constexpr int kWhTile = 240;                 // positions per tile (15 MMA k-steps)
constexpr int kWhLoad = 128;                 // positions per im2col TMA load
constexpr int kWhChunk = kWhLoad * 128;      // bytes of one load (64 channels)

// kXLoads: 128-position loads of the X region (needs 240 + 2 * pitch rows: 3 for pitch <= 72, 4 for
// pitch <= 127); the dY region (242 rows) always takes two.
template <int STAGES, int kXLoads>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmDw, const WgradHaloParams p) {
  constexpr int kXBytes = kXLoads * kWhChunk;
  constexpr int kDyBytes = 2 * kWhChunk;
  constexpr int kStageBytes = kXBytes + kDyBytes;
  constexpr uint32_t kTmemCols = 512;          // 2 x 192 accumulator columns
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cb = blockIdx.y;                       // 64-channel block of X / of the dW columns
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  int t_end = t_begin + p.tiles_per_cta;
  if (t_end > p.total_tiles) t_end = p.total_tiles;
  const int ntiles = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDw);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_launch();
  griddep_wait();
  if (ntiles <= 0) {
    // nothing to do for this split (uniform across the CTA); fall through to teardown
  } else if (warp == 0) {
    // ---- producer: the whole warp walks the tiles; lane 0 issues the TMA loads, a 128-position
    //      load that would START past the last image is replaced by zeros written by all lanes ----
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* xs = smem + stage * kStageBytes;
      uint8_t* ds = xs + kXBytes;
      // X region starts at J0 - pitch = kWhTile t; the dY region at J0 - 1 = pitch - 1 + kWhTile t
      long pos[kXLoads + 2];
#pragma unroll
      for (int j = 0; j < kXLoads; ++j) pos[j] = (long)t * kWhTile + j * kWhLoad;
      pos[kXLoads] = (long)t * kWhTile + p.pitch - 1;
      pos[kXLoads + 1] = pos[kXLoads] + kWhLoad;
      int n_tma = 0;
#pragma unroll
      for (int j = 0; j < kXLoads + 2; ++j) n_tma += pos[j] < p.total_pos ? 1 : 0;
#pragma unroll
      for (int j = 0; j < kXLoads + 2; ++j) {
        if (pos[j] >= p.total_pos) {
          uint4* z = reinterpret_cast<uint4*>(j < kXLoads ? xs + j * kWhChunk : ds + (j - kXLoads) * kWhChunk);
          for (int i = lane; i < kWhChunk / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
      }
      if (n_tma != kXLoads + 2) {
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], n_tma * kWhChunk);
#pragma unroll
        for (int j = 0; j < kXLoads + 2; ++j) {
          if (pos[j] < p.total_pos) {
            const int n_img = (int)(pos[j] / p.P);
            const int rem = (int)(pos[j] - (long)n_img * p.P);
            const int hp = rem / p.pitch;
            const int wp = rem - hp * p.pitch;
            if (j < kXLoads)
              tma_load_im2col_4d(xs + j * kWhChunk, &tmX, &full_bar[stage], cb * 64, wp - 1, hp - 1, n_img, 0, 0);
            else
              tma_load_im2col_4d(ds + (j - kXLoads) * kWhChunk, &tmDy, &full_bar[stage], 0, wp - 1, hp - 1, n_img, 0, 0);
          }
        }
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (whole warp converged, one elected lane issues) ----
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, 192, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      mbar_wait_w(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t xs = smem_u32(smem + stage * kStageBytes);
      const uint32_t ds = xs + kXBytes;
      // B: X rows +0 / +pitch / +2 pitch  (r = 0, 1, 2);  A1: dY rows +1 / +2 of the region that starts
      // at J0 - 1, i.e. J + 0 (s = 1) and J + 1 (s = 0);  A2: region row 0 = J - 1 (s = 2) [+ duplicate]
      const uint64_t b_desc = umma_smem_desc(xs, p.pitch * 128, 1024, kSwizzle128B);
      const uint64_t a1_desc = umma_smem_desc(ds + 128, 128, 1024, kSwizzle128B);
      const uint64_t a2_desc = umma_smem_desc(ds, 128, 1024, kSwizzle128B);
#pragma unroll
      for (int k = 0; k < kWhTile / 16; ++k) {
        // 16 positions along K = two 1024-byte groups = +128 in 16-byte address units
        umma_bf16_ss_w(tmem_base, a1_desc + 128 * k, b_desc + 128 * k, idesc, (i | k) != 0);
        umma_bf16_ss_w(tmem_base + 192, a2_desc + 128 * k, b_desc + 128 * k, idesc, (i | k) != 0);
      }
      umma_commit_w(&empty_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_w(&tmem_full_bar);
  } else {
    // ---- epilogue: TMEM -> 128B-swizzled fp32 slabs (the drained operand ring) -> TMA add-reduction ----
    const int quarter = warp & 3;                   // TMEM lane quarter
    const int half = quarter >> 1;                  // 0: lanes 0-63, 1: lanes 64-127
    const int k0 = (quarter & 1) * 32;              // out-channel offset of this warp's 32 rows
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    constexpr int kRedSlabs = 4;
    static_assert(STAGES * kStageBytes >= 4 * kRedSlabs * kSlabBytes, "ring too small");
    uint8_t* slabs = smem + (warp - 2) * kRedSlabs * kSlabBytes;
    // work list of this warp: (accumulator g, r, 32-column half); the upper lane half of the second
    // accumulator is a duplicate.  Every CTA starts at a different item: all splits finish their main
    // loops together, and walking the list in the same order would pile ~150 reductions onto the same
    // 4 KB of dW at the same time (same-address L2 atomics serialise).
    const int n_items = half == 0 ? 12 : 6;
#pragma unroll 1
    for (int it = 0; it < n_items; ++it) {
      const int item = (it + (int)blockIdx.x) % n_items;
      const int g = item / 6, chunk = item - g * 6;
      const int s_tap = g == 0 ? (half == 0 ? 1 : 0) : 2;
      const int r_tap = chunk >> 1;
      uint8_t* slab = slabs + (it % kRedSlabs) * kSlabBytes;
      if (lane == 0) tma_store_wait_read<kRedSlabs - 1>();
      __syncwarp();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + g * 192 + chunk * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(slab + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&tmDw, slab, (r_tap * 3 + s_tap) * p.Cin + cb * 64 + (chunk & 1) * 32, k0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------
struct IgemmMaps {
  CUtensorMap a, b, out, res, aux1, aux2;
};

template <int BN, int STAGES, int SLABS, int AUX, bool PRO = false, int CG = 0>
static int launch_igemm(const IgemmMaps& tm, const IgemmParams& p, cudaStream_t stream,
                        const BnPrologue* pro = nullptr) {
  constexpr int kEpiWarps = igemm_epi_warps(BN, CG);
  constexpr int smem =
      STAGES * (kABytes + BN * kBK * 2) + kEpiWarps * (SLABS + AUX) * kSlabBytes + 1024;
  static_assert(smem + 8192 + 512 + (PRO ? 8 * kProMaxC + 64 : 0) <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(igemm_kernel<BN, STAGES, SLABS, AUX, PRO, CG>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int grid = p.num_m_tiles * p.num_n_tiles;
  if (grid > sm_count()) grid = sm_count();
  // a tile step that is a multiple of the n-tile count pins every CTA to ONE n-tile: its per-channel
  // sums then live in registers for the whole kernel (no per-tile barrier + atomics, igemm_epilogue)
  if (p.stats != nullptr && p.num_n_tiles > 1 && grid >= p.num_n_tiles) grid -= grid % p.num_n_tiles;
  SIB_CUDA(launch_pdl(igemm_kernel<BN, STAGES, SLABS, AUX, PRO, CG>, dim3(grid),
                      dim3(igemm_threads(BN, PRO, CG)), smem, stream, tm.a, tm.b, tm.out,
                      tm.res, tm.aux1, tm.aux2, p, pro != nullptr ? *pro : BnPrologue{}));
  return 0;
}

template <int BN, int STAGES, int SLABS, int AUX>
static int launch_igemm2(const IgemmMaps& tm, const IgemmParams& p, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + (BN / 2) * kBK * 2) + 8 * (SLABS + AUX) * kSlabBytes + 1024;
  static_assert(smem + 8192 + 512 <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(igemm2_kernel<BN, STAGES, SLABS, AUX>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int pairs = (p.num_m_tiles / 2) * p.num_n_tiles;
  if (pairs > sm_count() / 2) pairs = sm_count() / 2;
  if (p.stats != nullptr && p.num_n_tiles > 1 && pairs >= p.num_n_tiles) pairs -= pairs % p.num_n_tiles;   // (see launch_igemm)
  SIB_CUDA(launch_pdl(igemm2_kernel<BN, STAGES, SLABS, AUX>, dim3(2 * pairs), dim3(kIgemmThreads), smem,
                      stream, tm.a, tm.b, tm.out, tm.res, tm.aux1, tm.aux2, p));
  return 0;
}

template <int BN, int BSTAGES, int kHaloAStages, int kALoads = 2>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, void* out, const HaloParams& p,
                       cudaStream_t stream) {
  constexpr int smem = kHaloAStages * kALoads * kABytes + BSTAGES * BN * kBK * 2 + 8 * kSlabBytes + 1024;
  static_assert(smem + 512 <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(halo3x3_kernel<BN, BSTAGES, kHaloAStages, kALoads>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int grid = p.N * p.num_h_tiles * p.num_n_tiles;
  if (grid > sm_count()) grid = sm_count();
  SIB_CUDA(launch_pdl(halo3x3_kernel<BN, BSTAGES, kHaloAStages, kALoads>, dim3(grid), dim3(kIgemmThreads), smem, stream, tmA,
                      tmB, static_cast<__nv_bfloat16*>(out), p));
  return 0;
}

// Fused BatchNorm-backward reduction riding on a dgrad (see IgemmParams::fuse).
struct BnBwdFuse {
  const void* aux1 = nullptr;        // mask source (and xhat source unless aux2 is given)
  const void* aux2 = nullptr;        // xhat source when the mask comes from a different tensor
  const float* mask_ss = nullptr;    // [2][C] forward scale/shift: mask = fmaf(aux1, s, t) > 0
  const float* mean_invstd = nullptr;
  int act = SIB_ACT_NONE;
  float slope = 0.f;
  float* sums = nullptr;             // [2][C]: sum g, sum g * xhat
};

// Output addressed as (channel, q, image row) with independent strides: the stride-2 dgrad writes
// one row parity of dx per launch (dgrad_s2_impl).
struct StridedOut {
  int q_pad;             // traversal width of the GEMM (8, 16 or 32; >= q_valid)
  int q_valid;           // real pixels per image row
  long rows;             // image rows
  long stride_q;         // elements between consecutive q
  long stride_row;       // elements between consecutive image rows
  int stat_c;            // real channels behind the GEMM columns (column j = channel j % stat_c)
};

// Shared by fprop and dgrad.  `in` is [N][IH][IW][Cin] NHWC bf16, `w` is [Cout][R][S][Cin].
// The traversal space (GEMM rows) is [N][TH][TW] and equals the output tensor's pixel space;
// row (n,p,q) reads taps starting at (p*stride - pad_h, q*stride - pad_w).
// out = conv [+ bias] [+ residual]; residual has the geometry of out (may alias it).
static int run_igemm(const void* in, const void* w, void* out, const void* residual, int N,
                     int IH, int IW, int Cin, int Cout, int R, int S, int stride, int pad_h,
                     int pad_w, int TH, int TW, const float* bias, float* stats, int flags,
                     cudaStream_t stream, const BnBwdFuse* fuse = nullptr,
                     const StridedOut* so = nullptr, const BnPrologue* pro = nullptr) {
  // 1x1 filters may have a ragged K: TMA zero-fills both operands past Cin
  SIB_CHECK(Cin % 64 == 0 || (R == 1 && S == 1 && Cin % 8 == 0),
            "igemm: Cin must be a multiple of 64 (or of 8 for 1x1 filters), got %d", Cin);
  SIB_CHECK(Cout % 8 == 0, "igemm: Cout must be a multiple of 8 (got %d)", Cout);
  SIB_CHECK((long)N * TH * TW < (1l << 31), "igemm: too many pixels");
  // 3x3 / stride 1 / pad 1 on 64 or 128 output channels at >= 28 pixel rows: halo-reuse kernel
  {
    const bool geometry = pro == nullptr && so == nullptr && R == 3 && S == 3 && stride == 1 && pad_h == 1 && pad_w == 1 && TH == IH &&
                          TW == IW && Cin % 64 == 0 && residual == nullptr && bias == nullptr &&
                          (IW + 2 <= 63 ? (Cout == 64 || Cout == 128)
                                        : (IW + 2 <= 127 && Cin == 64 && Cout == 64)) &&    // wide rows: weight-stationary variant only
                          (fuse == nullptr || fuse->aux2 == nullptr);
    // measured (B200, batch 256): 64 -> 64 @ 56x56 0.166 -> 0.12 ms; 128 -> 128 @ 28x28 is slower
    // than the im2col kernel (0.104 vs 0.092 ms: the streamed weight k-blocks dominate), so only
    // the weight-stationary shape is picked automatically
    const bool wanted = (flags & SIB_FLAG_FORCE_HALO) ||
                        (IW >= 28 && Cin == 64 && Cout == 64 && !(flags & SIB_FLAG_NO_HALO));
    if (geometry && wanted) {
      HaloParams h{};
      h.N = N; h.H = IH; h.W = IW; h.Cin = Cin; h.Cout = Cout;
      h.pitch = IW + 2;
      h.tile_h = 128 / h.pitch;
      if (h.tile_h > IH) h.tile_h = IH;
      h.num_h_tiles = (IH + h.tile_h - 1) / h.tile_h;
      h.num_n_tiles = 1;
      h.cin_blocks = Cin / 64;
      h.stats = stats;
      if (fuse != nullptr) {
        SIB_CHECK(stats == nullptr && fuse->aux1 && fuse->mean_invstd && fuse->sums,
                  "halo: fused BN backward needs aux1, mean_invstd and sums");
        h.fuse = 1;
        h.fuse_act = fuse->act;
        h.fuse_slope = fuse->slope;
        h.aux = static_cast<const __nv_bfloat16*>(fuse->aux1);
        h.mask_ss = fuse->mask_ss;
        h.mean_invstd = fuse->mean_invstd;
        h.stats = fuse->sums;
      }
      CUtensorMap tmA, tmB;
      if (int rc = make_tmap_im2col_bf16(&tmA, in, N, IH, IW, Cin, -1, -1, 1, 1, 1, 1, kBK, kBM, true)) return rc;
      if (int rc = make_tmap_2d_bf16(&tmB, w, Cout, (uint64_t)9 * Cin, (uint64_t)9 * Cin, Cout, kBK, true))
        return rc;
      if (h.stats != nullptr && !(flags & SIB_FLAG_STATS_ZEROED))
        SIB_CUDA(cudaMemsetAsync(h.stats, 0, sizeof(float) * 2 * Cout, stream));
      if (Cout == 64) {
        h.b_stationary = 9 * h.cin_blocks <= 9;
        if (h.pitch > 63) return launch_halo<64, 9, 2, 3>(tmA, tmB, out, h, stream);   // 384-row regions, 2 stages
        return launch_halo<64, 9, 3>(tmA, tmB, out, h, stream);
      }
      h.b_stationary = 0;
      return launch_halo<128, 6, 2>(tmA, tmB, out, h, stream);
    }
  }
  IgemmParams p{};
  p.M_total = N * TH * TW;
  p.Cout = Cout;
  p.cin_blocks = (Cin + 63) / 64;
  p.num_kblocks = R * S * p.cin_blocks;
  p.S = S;
  p.trav_hw = TH * TW;
  p.trav_w = TW;
  p.stride = stride;
  p.pad_h = pad_h;
  p.pad_w = pad_w;
  p.bias = bias;
  p.stats = stats;
  p.stat_c = so != nullptr ? so->stat_c : Cout;
  p.has_residual = residual != nullptr;
  int aux = 0;
  if (fuse != nullptr) {
    SIB_CHECK(stats == nullptr && fuse->aux1 != nullptr && fuse->mean_invstd != nullptr &&
                  fuse->sums != nullptr,
              "igemm: fused BN backward needs aux1, mean_invstd and sums (and no fprop stats)");
    aux = fuse->aux2 != nullptr ? 2 : 1;
    p.fuse = aux;
    p.fuse_act = fuse->act;
    p.fuse_slope = fuse->slope;
    p.mask_ss = fuse->mask_ss;
    p.mean_invstd = fuse->mean_invstd;
    p.stats = fuse->sums;
  }
  const bool plain =
      (R == 1 && S == 1 && stride == 1 && pad_h == 0 && pad_w == 0 && TH == IH && TW == IW);
  p.tiled_a = (plain && !(flags & SIB_FLAG_FORCE_IM2COL)) ? 1 : 0;

  int BN = 64;
  if (Cout > 64) BN = 128;
  if (Cout >= 256 && Cout % 256 == 0) BN = 256;
  if (flags & SIB_FLAG_TILE_N128 && BN == 256) BN = 128;
  p.num_m_tiles = (p.M_total + kBM - 1) / kBM;
  p.num_n_tiles = (Cout + BN - 1) / BN;

  IgemmMaps tm;
  int rc;
  if (p.tiled_a) {
    rc = make_tmap_2d_bf16(&tm.a, in, (uint64_t)p.M_total, Cin, Cin, kBM, kBK, true);
  } else {
    // base pixel range: [-pad, -pad + (T-1)*stride]  =>  upper corner = that max - (I-1)
    const int up_w = -pad_w + (TW - 1) * stride - (IW - 1);
    const int up_h = -pad_h + (TH - 1) * stride - (IH - 1);
    rc = make_tmap_im2col_bf16(&tm.a, in, N, IH, IW, Cin, -pad_w, -pad_h, up_w, up_h, stride,
                               stride, kBK, kBM, true);
  }
  if (rc) return rc;
  // 2-CTA pairs (M = 256 per cluster) for the tensor-bound shapes
  // (a 2-CTA variant with a 128-column tile measured no faster than the 1-CTA kernel:
  //  128 -> 128 3x3 at 28x28 0.094 vs 0.092 ms)
  const bool two_cta = BN == 256 && p.num_m_tiles % 2 == 0 && p.num_kblocks >= 4 &&
                       !(flags & SIB_FLAG_NO_2CTA) && pro == nullptr;
  if (pro != nullptr)
    SIB_CHECK(Cin % 64 == 0 && Cin <= kProMaxC && aux == 0 && so == nullptr && residual == nullptr,
              "igemm: the fused BatchNorm prologue needs Cin %% 64 == 0, Cin <= %d (got %d) and a plain fprop",
              kProMaxC, Cin);
  rc = make_tmap_2d_bf16(&tm.b, w, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin,
                         two_cta ? BN / 2 : BN, kBK, true);
  if (rc) return rc;
  const void* res_p = residual ? residual : out;
  const void* aux1_p = aux >= 1 ? fuse->aux1 : out;
  const void* aux2_p = aux >= 2 ? fuse->aux2 : out;
  if (so == nullptr) {
    rc = make_tmap_2d_bf16(&tm.out, out, (uint64_t)p.M_total, Cout, Cout, 32, 64, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tm.res, res_p, (uint64_t)p.M_total, Cout, Cout, 32, 64, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tm.aux1, aux1_p, (uint64_t)p.M_total, Cout, Cout, 32, 64, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tm.aux2, aux2_p, (uint64_t)p.M_total, Cout, Cout, 32, 64, true);
    if (rc) return rc;
  } else {
    SIB_CHECK((so->q_pad == 8 || so->q_pad == 16 || so->q_pad == 32) && so->q_valid <= so->q_pad &&
                  TW == so->q_pad && (long)N * TH == so->rows,
              "igemm: bad strided output view (q_pad %d, q_valid %d)", so->q_pad, so->q_valid);
    p.out_q = so->q_pad;
    const uint32_t bq = so->q_pad, br = 32 / so->q_pad;
    const void* ptrs[4] = {out, res_p, aux1_p, aux2_p};
    CUtensorMap* maps[4] = {&tm.out, &tm.res, &tm.aux1, &tm.aux2};
    for (int i = 0; i < 4; ++i) {
      rc = make_tmap_3d_bf16(maps[i], ptrs[i], Cout, so->q_valid, so->rows, so->stride_q,
                             so->stride_row, 64, bq, br);
      if (rc) return rc;
    }
  }
  if (p.stats != nullptr && !(flags & SIB_FLAG_STATS_ZEROED))
    SIB_CUDA(cudaMemsetAsync(p.stats, 0, sizeof(float) * 2 * p.stat_c, stream));
  // 16-warp epilogue for the 256-column 1-CTA tiles: opt-in (SIB_EPI16=1) until measured faster
  static const bool epi16 = [] { const char* e = getenv("SIB_EPI16"); return e && e[0] == '1'; }();
  if (pro != nullptr) {
    if (BN == 64) return launch_igemm<64, 6, 1, 0, true>(tm, p, stream, pro);
    if (BN == 128) return launch_igemm<128, 5, 1, 0, true>(tm, p, stream, pro);
    return launch_igemm<256, 3, 2, 0, true>(tm, p, stream, pro);
  }
  if (aux == 0) {
    if (two_cta) return launch_igemm2<256, 5, 1, 0>(tm, p, stream);
    if (BN == 64) return launch_igemm<64, 6, 1, 0>(tm, p, stream);
    if (BN == 128) return launch_igemm<128, 5, 1, 0>(tm, p, stream);
    if (epi16) return launch_igemm<256, 3, 1, 0, false, 4>(tm, p, stream);
    return launch_igemm<256, 3, 2, 0>(tm, p, stream);
  }
  // fused variants trade pipeline stages for the auxiliary slabs (227 KB of smem per CTA)
  if (aux == 1) {
    if (two_cta) return launch_igemm2<256, 4, 1, 1>(tm, p, stream);
    if (BN == 64) return launch_igemm<64, 6, 1, 1>(tm, p, stream);
    if (BN == 128) return launch_igemm<128, 4, 1, 1>(tm, p, stream);
    return launch_igemm<256, 3, 1, 1>(tm, p, stream);
  }
  if (two_cta) return launch_igemm2<256, 3, 1, 2>(tm, p, stream);
  if (BN == 64) return launch_igemm<64, 5, 1, 2>(tm, p, stream);
  if (BN == 128) return launch_igemm<128, 3, 1, 2>(tm, p, stream);
  return launch_igemm<256, 2, 1, 2>(tm, p, stream);
}

template <int STAGES, bool kTmaReduce, int CL>
static int launch_wgrad(const CUtensorMap& tmDy, const CUtensorMap& tmX, const CUtensorMap& tmDw,
                        float* dw, const WgradParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (2 + kWgUnits) * kWgPix * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(wgrad_kernel<STAGES, kTmaReduce, CL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  SIB_CUDA(launch_cluster_pdl(wgrad_kernel<STAGES, kTmaReduce, CL>, grid, dim3(kThreads), smem,
                              stream, (unsigned)CL, tmDy, tmX, tmDw, dw, p));
  return 0;
}

template <int STAGES, int kXLoads>
static int launch_wgrad_halo(const CUtensorMap& tmDy, const CUtensorMap& tmX, const CUtensorMap& tmDw,
                             const WgradHaloParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (kXLoads + 2) * kWhChunk + 1024;
  static_assert(smem + 512 <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    SIB_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<STAGES, kXLoads>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  SIB_CUDA(launch_pdl(wgrad_halo_kernel<STAGES, kXLoads>, grid, dim3(kThreads), smem, stream, tmDy, tmX,
                      tmDw, p));
  return 0;
}

}  // namespace sib

using namespace sib;

extern "C" int sib_conv2d_fprop(const void* x, const void* w, void* y, int N, int H, int W,
                                int C, int K, int R, int S, int stride, int pad_h, int pad_w,
                                int OH, int OW, const float* bias, float* stats, int flags,
                                void* stream) {
  SIB_CHECK(OH >= 1 && OW >= 1 && (OH - 1) * stride - pad_h < H && (OW - 1) * stride - pad_w < W,
            "fprop: output extent %dx%d inconsistent with input %dx%d", OH, OW, H, W);
  return run_igemm(x, w, y, nullptr, N, H, W, C, K, R, S, stride, pad_h, pad_w, OH, OW, bias,
                   stats, flags, static_cast<cudaStream_t>(stream));
}

// y = conv(act(BN(x)), w): the BatchNorm (+ activation) of the producer layer is applied to the
// A operand inside the kernel (BnPrologue); training mode (bn_stats != null) also finalises that
// BatchNorm (mean/invstd, scale/shift, running statistics).
extern "C" int sib_conv2d_fprop_bnact(const void* x, const void* w, void* y, int N, int H, int W,
                                      int C, int K, int R, int S, int stride, int pad_h, int pad_w,
                                      int OH, int OW, float* stats, int flags, const float* bn_stats,
                                      const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, float* mean_invstd, float* scale_shift,
                                      double count, float eps, float momentum, int act, float slope,
                                      void* stream) {
  SIB_CHECK(OH >= 1 && OW >= 1 && (OH - 1) * stride - pad_h < H && (OW - 1) * stride - pad_w < W,
            "fprop_bnact: output extent %dx%d inconsistent with input %dx%d", OH, OW, H, W);
  SIB_CHECK(scale_shift != nullptr && (bn_stats == nullptr || mean_invstd != nullptr),
            "fprop_bnact: scale_shift (and mean_invstd in training mode) are required");
  BnPrologue pro{};
  pro.stats = bn_stats; pro.gamma = gamma; pro.beta = beta;
  pro.running_mean = running_mean; pro.running_var = running_var;
  pro.mean_invstd = mean_invstd; pro.scale_shift = scale_shift;
  pro.count = (float)count; pro.eps = eps; pro.momentum = momentum;
  pro.act = act; pro.slope = slope;
  pro.IH = H; pro.IW = W; pro.Cin = C;
  return run_igemm(x, w, y, nullptr, N, H, W, C, K, R, S, stride, pad_h, pad_w, OH, OW, nullptr,
                   stats, flags, static_cast<cudaStream_t>(stream), nullptr, nullptr, &pro);
}

extern "C" int sib_upsample_zero(const void* dy, void* up, int N, int OH, int OW, int C, int UH,
                                 int UW, int stride, void* stream);
extern "C" int sib_scatter_add_strided(const void* src, void* dst, int N, int OH, int OW, int C,
                                       int H, int W, int stride, void* stream);

// dx = dgrad(dy) [+ residual].  stride 1: full correlation of dY with the tap-flipped,
// transposed filter.  stride > 1 needs `workspace`:
//   1x1 : compact GEMM into workspace [N][OH][OW][C], then dx[n][p*s][q*s][:] += it
//         (dx must already hold the other gradient contribution; residual must be null)
//   RxS : zero-insert dY into workspace [N][(OH-1)s+1][(OW-1)s+1][K], stride-1 conv over it
static int dgrad_impl(const void* dy, const void* w_dgrad, void* dx, const void* residual,
                      void* workspace, int N, int H, int W, int C, int K, int R, int S, int stride,
                      int pad, int flags, void* stream, const BnBwdFuse* fuse) {
  const int OH = (H + 2 * pad - R) / stride + 1;
  const int OW = (W + 2 * pad - S) / stride + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stride == 1)
    return run_igemm(dy, w_dgrad, dx, residual, N, OH, OW, K, C, R, S, 1, R - 1 - pad,
                     S - 1 - pad, H, W, nullptr, nullptr, flags, st, fuse);
  SIB_CHECK(workspace != nullptr, "strided dgrad needs a workspace");
  if (R == 1 && S == 1 && pad == 0) {
    SIB_CHECK(residual == nullptr, "strided 1x1 dgrad accumulates into dx; residual must be null");
    SIB_CHECK(fuse == nullptr, "strided 1x1 dgrad cannot carry the fused BN backward reduction");
    if (int rc = run_igemm(dy, w_dgrad, workspace, nullptr, N, OH, OW, K, C, 1, 1, 1, 0, 0, OH, OW,
                           nullptr, nullptr, flags, st))
      return rc;
    return sib_scatter_add_strided(workspace, dx, N, OH, OW, C, H, W, stride, stream);
  }
  const int UH = (OH - 1) * stride + 1, UW = (OW - 1) * stride + 1;
  if (int rc = sib_upsample_zero(dy, workspace, N, OH, OW, K, UH, UW, stride, stream)) return rc;
  return run_igemm(workspace, w_dgrad, dx, residual, N, UH, UW, K, C, R, S, 1, R - 1 - pad,
                   S - 1 - pad, H, W, nullptr, nullptr, flags, st, fuse);
}

extern "C" int sib_conv2d_dgrad(const void* dy, const void* w_dgrad, void* dx,
                                const void* residual, void* workspace, int N, int H, int W,
                                int C, int K, int R, int S, int stride, int pad, int flags,
                                void* stream) {
  return dgrad_impl(dy, w_dgrad, dx, residual, workspace, N, H, W, C, K, R, S, stride, pad, flags,
                    stream, nullptr);
}

extern "C" int sib_conv2d_dgrad_bnbwd(const void* dy, const void* w_dgrad, void* dx,
                                      const void* residual, void* workspace, int N, int H, int W,
                                      int C, int K, int R, int S, int stride, int pad, int flags,
                                      const void* mask_src, const float* mask_ss,
                                      const void* xhat_src, const float* mean_invstd, int act,
                                      float slope, float* sums, void* stream) {
  SIB_CHECK(mask_src != nullptr && mean_invstd != nullptr && sums != nullptr,
            "dgrad_bnbwd: mask_src, mean_invstd and sums are required");
  BnBwdFuse f;
  f.aux1 = mask_src;
  f.aux2 = (xhat_src != nullptr && xhat_src != mask_src) ? xhat_src : nullptr;
  f.mask_ss = mask_ss;
  f.mean_invstd = mean_invstd;
  f.act = act;
  f.slope = slope;
  f.sums = sums;
  return dgrad_impl(dy, w_dgrad, dx, residual, workspace, N, H, W, C, K, R, S, stride, pad, flags,
                    stream, &f);
}

// 3x3 / stride 2 / pad 1 dgrad without zero insertion.  dx row 2p+a only receives filter rows
// r with (2p + a + 1 - r) even, so each row parity a is a small stride-1 convolution over dy:
//   a = 0 : r = 1            -> 1 x 2 taps
//   a = 1 : r in {2, 0}      -> 2 x 2 taps           (tap (dp, dq) reads dy[p + dp][q + dq])
// with 2C output columns (b, c): column parity b = 1 uses s = 2 - 2 dq, b = 0 uses s = 1 at dq = 0
// (zero weights elsewhere; packed by sib_pack_dgrad_s2).  The output row (n, p) of parity a IS the
// dx row (n, 2p + a) verbatim ((q, b, c) = (2q + b) C + c), so the GEMM stores straight into dx
// through a 3-D tensor map (channel pair block, q, image row).  The traversal width is padded to
// 8 / 16 / 32 so that a 32-row epilogue slab is a whole number of image rows; the padding columns
// read zeros (im2col bounding box widened to the right) and are clipped on store.
// 12 C K MACs per output pixel pair instead of 36 for the zero-inserted form.
static int dgrad_s2_impl(const void* dy, const void* sub0, const void* sub1, void* dx, int N, int H,
                         int W, int C, int K, int flags, void* stream, const BnBwdFuse* fuse) {
  SIB_CHECK(H % 2 == 0 && W % 2 == 0 && W / 2 <= 32, "dgrad_s2: needs even H, W and W/2 <= 32 (got %dx%d)", H, W);
  SIB_CHECK((2 * C) % 64 == 0 && K % 64 == 0, "dgrad_s2: needs 2C and K multiples of 64");
  const int OH = H / 2, OW = W / 2;
  StridedOut so{};
  so.q_pad = OW <= 8 ? 8 : (OW <= 16 ? 16 : 32);
  so.q_valid = OW;
  so.rows = (long)N * OH;
  so.stride_q = 2 * C;
  so.stride_row = 2l * W * C;
  so.stat_c = C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int a = 0; a < 2; ++a) {
    const long off = (long)a * W * C;
    BnBwdFuse f;
    if (fuse != nullptr) {
      f = *fuse;
      f.aux1 = static_cast<const __nv_bfloat16*>(fuse->aux1) + off;
      if (fuse->aux2 != nullptr) f.aux2 = static_cast<const __nv_bfloat16*>(fuse->aux2) + off;
    }
    int fl = flags;
    if (a == 1 && fuse != nullptr) fl |= SIB_FLAG_STATS_ZEROED;   // both parities add into one buffer
    if (int rc = run_igemm(dy, a == 0 ? sub0 : sub1, static_cast<__nv_bfloat16*>(dx) + off, nullptr,
                           N, OH, OW, K, 2 * C, a == 0 ? 1 : 2, 2, 1, 0, 0, OH, so.q_pad, nullptr,
                           nullptr, fl, st, fuse != nullptr ? &f : nullptr, &so))
      return rc;
  }
  return 0;
}

extern "C" int sib_conv2d_dgrad_s2(const void* dy, const void* w_sub0, const void* w_sub1, void* dx,
                                   int N, int H, int W, int C, int K, int flags,
                                   const void* mask_src, const float* mask_ss,
                                   const float* mean_invstd, int act, float slope, float* sums,
                                   void* stream) {
  if (mask_src == nullptr)
    return dgrad_s2_impl(dy, w_sub0, w_sub1, dx, N, H, W, C, K, flags, stream, nullptr);
  SIB_CHECK(mean_invstd != nullptr && sums != nullptr, "dgrad_s2: mean_invstd and sums are required with mask_src");
  BnBwdFuse f;
  f.aux1 = mask_src;
  f.mask_ss = mask_ss;
  f.mean_invstd = mean_invstd;
  f.act = act;
  f.slope = slope;
  f.sums = sums;
  return dgrad_s2_impl(dy, w_sub0, w_sub1, dx, N, H, W, C, K, flags, stream, &f);
}

extern "C" int sib_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W,
                                int C, int K, int R, int S, int stride, int pad_h, int pad_w,
                                int OH, int OW, int flags, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SIB_CHECK(C % 64 == 0, "wgrad: Cin must be a multiple of 64 (got %d)", C);
  SIB_CHECK(K % 8 == 0, "wgrad: Cout must be a multiple of 8 (got %d)", K);
  static const bool det0 = [] { const char* e = getenv("SIB_DETERMINISTIC"); return e && e[0] == '1'; }();
  // 3x3 / stride 1 / pad 1 with 64 output channels: halo-reuse kernel (both operands loaded once per
  // 128 padded positions, six + three taps per MMA pair).  Not in deterministic mode (split reduction).
  if (R == 3 && S == 3 && stride == 1 && pad_h == 1 && pad_w == 1 && OH == H && OW == W && K == 64 &&
      W + 2 <= 127 && !det0 && !(flags & SIB_FLAG_NO_HALO)) {
    WgradHaloParams hp{};
    hp.N = N;
    hp.Cin = C;
    hp.pitch = W + 2;
    hp.P = (H + 2) * hp.pitch;
    hp.total_pos = (long)N * hp.P;
    SIB_CHECK(hp.total_pos < (1l << 31), "wgrad: too many pixels");
    hp.total_tiles = (int)((hp.total_pos - hp.pitch + kWhTile - 1) / kWhTile);
    const int cbs = C / 64;
    int splits = sm_count() / cbs;
    if (splits < 1) splits = 1;
    if (splits > hp.total_tiles) splits = hp.total_tiles;
    hp.tiles_per_cta = (hp.total_tiles + splits - 1) / splits;
    splits = (hp.total_tiles + hp.tiles_per_cta - 1) / hp.tiles_per_cta;
    CUtensorMap tmDy, tmX, tmDw;
    if (int rc = make_tmap_im2col_bf16(&tmDy, dy, N, H, W, K, -1, -1, 1, 1, 1, 1, 64, kWhLoad, true)) return rc;
    if (int rc = make_tmap_im2col_bf16(&tmX, x, N, H, W, C, -1, -1, 1, 1, 1, 1, 64, kWhLoad, true)) return rc;
    if (int rc = make_tmap_2d_f32(&tmDw, dw, K, (uint64_t)9 * C, (uint64_t)9 * C, 32, 32)) return rc;
    const dim3 grid(splits, cbs, 1);
    if (hp.pitch <= 72) return launch_wgrad_halo<2, 3>(tmDy, tmX, tmDw, hp, grid, st);
    return launch_wgrad_halo<2, 4>(tmDy, tmX, tmDw, hp, grid, st);
  }
  WgradParams p{};
  p.M_total = N * OH * OW;
  p.Cout = K;
  p.Cin = C;
  p.S = S;
  p.trav_hw = OH * OW;
  p.trav_w = OW;
  p.stride = stride;
  p.pad_h = pad_h;
  p.pad_w = pad_w;
  p.total_kblocks = (p.M_total + kWgPix - 1) / kWgPix;
  p.ldw = R * S * C;
  p.cin_blocks = C / 64;
  p.total_units = R * S * p.cin_blocks;
  const int groups = (p.total_units + kWgUnits - 1) / kWgUnits;
  const int tiles = ((K + kBM - 1) / kBM) * groups;
  // split the pixel reduction so that the CTAs fill `waves` rounds of the 2-per-SM slots without
  // spilling into a nearly empty extra round (floor, not ceil), >= 8 k-blocks each
  static const int waves = [] { const char* e = getenv("SIB_WGRAD_WAVES"); return e ? atoi(e) : 1; }();
  // SIB_DETERMINISTIC=1: no split over the pixel reduction (the TMA add-reduction of several splits
  // into one gradient tile lands in a varying order); one CTA per gradient tile, much slower
  static const bool det = [] { const char* e = getenv("SIB_DETERMINISTIC"); return e && e[0] == '1'; }();
  int splits = det ? 1 : (waves * 2 * sm_count()) / tiles;
  int max_splits = (p.total_kblocks + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kblocks_per_split = (p.total_kblocks + splits - 1) / splits;
  splits = (p.total_kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;

  CUtensorMap tmDy, tmX;
  int rc = make_tmap_2d_bf16(&tmDy, dy, (uint64_t)p.M_total, K, K, kWgPix, 64, true);
  if (rc) return rc;
  const int up_w = -pad_w + (OW - 1) * stride - (W - 1);
  const int up_h = -pad_h + (OH - 1) * stride - (H - 1);
  rc = make_tmap_im2col_bf16(&tmX, x, N, H, W, C, -pad_w, -pad_h, up_w, up_h, stride, stride, 64,
                             kWgPix, true);
  if (rc) return rc;
  dim3 grid((K + kBM - 1) / kBM, groups, splits);
  static const bool lane_red = [] { const char* e = getenv("SIB_WGRAD_LANE_RED"); return e && atoi(e); }();
  CUtensorMap tmDw;
  rc = make_tmap_2d_f32(&tmDw, dw, K, p.ldw, p.ldw, 32, 32);
  if (rc) return rc;
  // 97 KB smem -> two CTAs per SM
  if (lane_red) return launch_wgrad<2, false, 1>(tmDy, tmX, tmDw, dw, p, grid, st);
  // clusters of 2 out-channel tiles share the X operand by TMA multicast (measured: -5..-9 % on the
  // layers with >= 256 output channels; clusters of 4 are much slower, 0.090 vs 0.053 ms)
  static const int max_cl = [] { const char* e = getenv("SIB_WGRAD_CLUSTER"); return e ? atoi(e) : 2; }();
  if (max_cl >= 4 && grid.x % 4 == 0) return launch_wgrad<2, true, 4>(tmDy, tmX, tmDw, dw, p, grid, st);
  if (max_cl >= 2 && grid.x % 2 == 0) return launch_wgrad<2, true, 2>(tmDy, tmX, tmDw, dw, p, grid, st);
  return launch_wgrad<2, true, 1>(tmDy, tmX, tmDw, dw, p, grid, st);
}
