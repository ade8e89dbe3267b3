// Host-side plumbing shared by the C-ABI entry points: error reporting and
// CUtensorMap construction through the driver entry points (no -lcuda link).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sib {

// Records a message retrievable through sib_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);

#define SIB_CHECK(cond, ...)                      \
  do {                                            \
    if (!(cond)) return ::sib::fail(1, __VA_ARGS__); \
  } while (0)

#define SIB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return ::sib::fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                       \
  } while (0)

#define SIB_LAUNCH_CHECK() SIB_CUDA(cudaGetLastError())

int sm_count();

// Programmatic dependent launch (PDL): the kernel may start (prologue: barrier init, TMEM
// allocation, descriptor prefetch) while the previous kernel of the stream is still draining.
// Every kernel launched through here MUST execute griddep_wait() (common.cuh) before it touches
// global memory.  Opt-in with SIB_PDL=1; otherwise they are ordinary stream-ordered launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t stream, unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {      // thread-block cluster along x (runtime cluster dimension)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  return launch_cluster_pdl(kernel, grid, block, smem, stream, 1u, args...);
}

// 2-D row-major bf16 matrix [rows][cols] (cols contiguous); box = box_rows x box_cols.
// swizzle128: inner box extent must be 64 elements (128 bytes).
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols,
                      bool swizzle128);

// 2-D row-major fp32 matrix, box = 32 rows x 32 columns (128 bytes), 128B swizzle; used for
// TMA add-reductions of fp32 accumulator tiles
int make_tmap_2d_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                     uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

// 3-D bf16 tensor (d0 contiguous, strides in elements for d1 and d2), box = box0 x box1 x box2,
// 128B swizzle (box0 must be 64).  Out-of-bounds elements are dropped on stores, zero on loads.
int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1,
                      uint32_t box2);

// Tiled 4-D map over an NHWC bf16 tensor, dims (C, W, H, N); box = box_c channels x box_w x box_h
// pixels of one image, 128B swizzle, out-of-bounds pixels (halo) read as zeros.
int make_tmap_nhwc_tile_bf16(CUtensorMap* tm, const void* base, int N, int H, int W, int C,
                             uint32_t box_c, uint32_t box_w, uint32_t box_h);

// 3-D view used for the no-swizzle K-major "core matrix" layout:
// dims (inner=8 elems, rows, kchunks) with strides (1, row_stride, 8) elements.
int make_tmap_kchunk_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_kchunks);

// im2col map over an NHWC bf16 tensor.  Base-pixel bounding box:
//   lower = {lo_w, lo_h}, upper = {up_w, up_h}; traversal strides {sw, sh};
// each load gathers `pixels` pixels x `channels` channels.
int make_tmap_im2col_bf16(CUtensorMap* tm, const void* base, int N, int H, int W, int C,
                          int lo_w, int lo_h, int up_w, int up_h, int sw, int sh,
                          uint32_t channels, uint32_t pixels, bool swizzle128);

}  // namespace sib
