// Probe: may the start address of a 128B-swizzled K-major UMMA operand be advanced by whole rows
// (128 bytes each) that are not a multiple of the 8-row swizzle atom?  This is what smem halo reuse
// across the taps of a 3x3 filter needs (tap (r,s) = the same tile shifted by r*pitch + s rows).
// Two descriptor variants are tried: base_offset = 0 and base_offset = (start >> 7) & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I sota_imagenet_b200/csrc \
//        scripts/probes/umma_shift_probe.cu sota_imagenet_b200/csrc/host.cu -o gpurun_out/umma_shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"
#include "host.h"
using namespace sib;

constexpr int kRows = 256, kK = 64, kN = 64;

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmB,
                                                   float* __restrict__ d, int shift, int variant) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sa = smem;                       // 256 rows x 128 B
  uint8_t* sb = smem + kRows * 128;         // 64 rows x 128 B
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
    mbar_arrive_expect_tx(&full_bar, kRows * 128 + kN * 128);
    tma_load_2d(sa, &tmA, &full_bar, 0, 0);
    tma_load_2d(sb, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    uint64_t a_desc = umma_smem_desc(a_addr, 16, 1024, kSwizzle128B);
    if (variant == 1) a_desc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t b_desc = umma_smem_desc(smem_u32(sb), 16, 1024, kSwizzle128B);
    constexpr uint32_t idesc = umma_idesc_bf16(128, kN, 0, 0);
    for (int k = 0; k < kK / 16; ++k) umma_bf16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
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
  for (int j = 0; j < 64; ++j) d[row * kN + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int main() {
  std::vector<__nv_bfloat16> ha(kRows * kK), hb(kN * kK);
  std::vector<float> fa(kRows * kK), fb(kN * kK);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 4.f); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 13 - 6) / 8.f); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * kN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (make_tmap_2d_bf16(&tmA, da, kRows, kK, kK, kRows, kK, true)) { printf("tmap A failed\n"); return 1; }
  if (make_tmap_2d_bf16(&tmB, db, kN, kK, kK, kN, kK, true)) { printf("tmap B failed\n"); return 1; }
  const int smem = (kRows + kN) * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 16, 58, 59, 117, 118};
  std::vector<float> hd(128 * kN);
  for (int variant = 0; variant < 2; ++variant) {
    for (int shift : shifts) {
      cudaMemset(dd, 0, 128 * kN * 4);
      probe_kernel<<<1, 128, smem>>>(tmA, tmB, dd, shift, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d shift %d: CUDA error %s\n", variant, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fa[(shift + i) * kK + k] * fb[n * kK + k];
          double err = fabs(ref - hd[i * kN + n]);
          if (err > maxerr) maxerr = err;
        }
      printf("base_offset %s  shift %3d rows: max abs err %.4g %s\n", variant ? "=(addr>>7)&7" : "=0           ", shift, maxerr,
             maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
