// Probe for the halo-reuse weight-gradient kernel: MN-major 128B-swizzled UMMA operands (rows of 128
// bytes = one pixel x 64 channels; the channels are the M/N index, the pixels the K index).
//   (1) may the start address be advanced by whole pixel rows that are not a multiple of the 8-row
//       swizzle atom (a shift along K)?
//   (2) may the leading byte offset (distance between the two 64-channel halves of an M = 128 operand)
//       be an arbitrary multiple of 128 bytes, i.e. may the two halves be two SHIFTED VIEWS of one
//       region (two filter taps in one MMA)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I sota_imagenet_b200/csrc \
//        scripts/probes/umma_mn_major_shift_probe.cu sota_imagenet_b200/csrc/host.cu -o gpurun_out/umma_mn_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"
#include "host.h"
using namespace sib;

constexpr int kRows = 512, kC = 64, kKpix = 64;   // region rows; channels; pixels reduced per test

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmB,
                                                   float* __restrict__ d, int shift, int lbo_rows) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sa = smem;                       // 512 pixel rows x 128 B
  uint8_t* sb = smem + kRows * 128;         // 64 pixel rows x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full_bar, kRows * 128 + kKpix * 128);
    tma_load_2d(sa, &tmA, &full_bar, 0, 0);
    tma_load_2d(sa + 256 * 128, &tmA, &full_bar, 0, 256);
    tma_load_2d(sb, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint64_t a_desc = umma_smem_desc(smem_u32(sa) + shift * 128, lbo_rows * 128, 1024, kSwizzle128B);
    const uint64_t b_desc = umma_smem_desc(smem_u32(sb), 8192, 1024, kSwizzle128B);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    for (int k = 0; k < kKpix / 16; ++k) umma_bf16_ss(tmem_base, a_desc + 128 * k, b_desc + 128 * k, idesc, k != 0);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  uint32_t r[64];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_32x32b_x32(taddr + 32, r + 32);
  tmem_ld_wait();
  const int row = warp * 32 + lane;
  for (int j = 0; j < 64; ++j) d[row * 64 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int main() {
  std::vector<__nv_bfloat16> ha(kRows * kC), hb(kKpix * kC);
  std::vector<float> fa(kRows * kC), fb(kKpix * kC);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 4.f); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 13 - 6) / 8.f); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (make_tmap_2d_bf16(&tmA, da, kRows, kC, kC, 256, kC, true)) { printf("tmap A failed\n"); return 1; }
  if (make_tmap_2d_bf16(&tmB, db, kKpix, kC, kC, kKpix, kC, true)) { printf("tmap B failed\n"); return 1; }
  const int smem = (kRows + kKpix) * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 3, 8, 9, 59, 117};
  const int lbos[] = {64, 1, 2, 7, 8, 57, 58, 59, 114, 116, 230};
  std::vector<float> hd(128 * 64);
  int bad = 0;
  for (int lbo : lbos) {
    for (int shift : shifts) {
      cudaMemset(dd, 0, 128 * 64 * 4);
      probe_kernel<<<1, 128, smem>>>(tmA, tmB, dd, shift, lbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("lbo %d shift %d: CUDA error %s\n", lbo, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      double err_lo = 0, err_hi = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          const int base = shift + (m >= 64 ? lbo : 0);
          double ref = 0;
          for (int k = 0; k < kKpix; ++k) ref += (double)fa[(base + k) * kC + (m & 63)] * fb[k * kC + n];
          const double err = fabs(ref - hd[m * 64 + n]);
          if (m < 64) { if (err > err_lo) err_lo = err; } else { if (err > err_hi) err_hi = err; }
        }
      const bool ok = err_lo < 1e-3 && err_hi < 1e-3;
      bad += !ok;
      printf("LBO %3d rows  K-shift %3d rows: max abs err lower half %.4g, upper half %.4g %s\n", lbo, shift, err_lo,
             err_hi, ok ? "OK" : "MISMATCH");
    }
  }
  printf("%s\n", bad ? "SOME MISMATCH" : "ALL OK");
  return 0;
}
