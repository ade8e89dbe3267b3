// Error reporting + CUtensorMap builders.  The driver's cuTensorMapEncode* are
// resolved at run time with cudaGetDriverEntryPoint so the library does not link
// libcuda (the build box has no driver).
#include <stdlib.h>
#include "host.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/sib200.h"

namespace sib {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = [] {
    // measured on B200 (ResNet-50 step inside one CUDA graph): 20.55 ms with PDL edges vs 20.26 ms
    // without -- graph kernel nodes already launch back to back, so it stays opt-in
    const char* e = getenv("SIB_PDL");
    return e && e[0] == '1';
  }();
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;
static int g_driver_version = 0;
static std::once_flag g_once;

static void resolve() {
  cudaDriverEntryPointQueryResult q;
  void* p = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
          cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_tiled = reinterpret_cast<EncodeTiledFn>(p);
  p = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) ==
          cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_im2col = reinterpret_cast<EncodeIm2colFn>(p);
  cudaDriverGetVersion(&g_driver_version);
}

int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols,
                      bool swizzle128) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_tiled != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  SIB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SIB_CHECK((row_stride_elems * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled(2d) failed: %d rows=%llu cols=%llu box=%ux%u", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
  return 0;
}

int make_tmap_2d_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                     uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_tiled != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  SIB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SIB_CHECK((row_stride_elems * 4) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d f32) failed: %d rows=%llu cols=%llu",
            (int)r, (unsigned long long)rows, (unsigned long long)cols);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1,
                      uint32_t box2) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_tiled != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  SIB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SIB_CHECK((stride1_elems * 2) % 16 == 0 && (stride2_elems * 2) % 16 == 0,
            "TMA strides must be multiples of 16 bytes");
  SIB_CHECK(box0 == 64 && box1 <= 256 && box2 <= 256, "3d box %ux%ux%u unsupported", box0, box1, box2);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d dims=%llu,%llu,%llu", (int)r,
            (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return 0;
}

int make_tmap_nhwc_tile_bf16(CUtensorMap* tm, const void* base, int N, int H, int W, int C,
                             uint32_t box_c, uint32_t box_w, uint32_t box_h) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_tiled != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  SIB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SIB_CHECK((C * 2) % 16 == 0, "NHWC channel count must be a multiple of 8 for TMA");
  SIB_CHECK(box_c == 64 && box_w <= 256 && box_h <= 256, "nhwc tile box %ux%ux%u unsupported", box_c,
            box_w, box_h);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc tile) failed: %d", (int)r);
  return 0;
}

int make_tmap_kchunk_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_kchunks) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_tiled != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  SIB_CHECK(cols % 8 == 0, "kchunk map needs cols %% 8 == 0");
  cuuint64_t dims[3] = {8, rows, cols / 8};
  cuuint64_t strides[2] = {row_stride_elems * 2, 16};
  cuuint32_t box[3] = {8, box_rows, box_kchunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(kchunk) failed: %d", (int)r);
  return 0;
}

int make_tmap_im2col_bf16(CUtensorMap* tm, const void* base, int N, int H, int W, int C,
                          int lo_w, int lo_h, int up_w, int up_h, int sw, int sh,
                          uint32_t channels, uint32_t pixels, bool swizzle128) {
  std::call_once(g_once, resolve);
  SIB_CHECK(g_im2col != nullptr, "cuTensorMapEncodeIm2col unavailable (no CUDA driver?)");
  SIB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SIB_CHECK((C * 2) % 16 == 0, "NHWC channel count must be a multiple of 8 for TMA");
  SIB_CHECK(lo_w >= -128 && lo_w <= 127 && lo_h >= -128 && lo_h <= 127 && up_w >= -128 &&
                up_w <= 127 && up_h >= -128 && up_h <= 127,
            "im2col box corner out of range");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  int lower[2] = {lo_w, lo_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
  CUresult r = g_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                        strides, lower, upper, channels, pixels, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIB_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeIm2col failed: %d NHWC=%d,%d,%d,%d lo=(%d,%d) up=(%d,%d) "
            "stride=(%d,%d) ch=%u px=%u",
            (int)r, N, H, W, C, lo_w, lo_h, up_w, up_h, sw, sh, channels, pixels);
  // Known driver issue (<= 13.1): im2col descriptors of tensors smaller than 128 KiB get a
  // flag that makes the hardware mis-handle them; clear it (same fix CUTLASS applies).
  if (g_driver_version <= 13010) {
    uint64_t bytes = (uint64_t)N * H * W * C * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  }
  return 0;
}

}  // namespace sib

extern "C" const char* sib_last_error(void) { return sib::g_err; }

extern "C" int sib_abi_version(void) { return SIB_ABI_VERSION; }

extern "C" int sib_device_check(void) {
  int dev = 0;
  SIB_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  SIB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SIB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  SIB_CHECK(major == 10, "sota_imagenet_b200 needs an sm_100a GPU (B200); found sm_%d%d", major,
            minor);
  return 0;
}


// CRC-32C (Castagnoli), slicing-by-8, on HOST memory: the record framing check of the TFRecord reader
// (records.py; reference create_records.py:84-106 writes the shards through TensorFlow, DALI verifies
// them when reading, dali_dataloader.py:47-65).  ~1 GB/s per core instead of a per-byte Python loop.
extern "C" int sib_crc32c_host(const void* data_host, unsigned long long n, unsigned int* out_host) {
  static uint32_t tab[8][256];
  static std::once_flag once;
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xFF];
  });
  SIB_CHECK(out_host != nullptr && (data_host != nullptr || n == 0), "crc32c: null pointer");
  const unsigned char* p = static_cast<const unsigned char*>(data_host);
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tab[7][lo & 0xFF] ^ tab[6][(lo >> 8) & 0xFF] ^ tab[5][(lo >> 16) & 0xFF] ^ tab[4][lo >> 24] ^
        tab[3][hi & 0xFF] ^ tab[2][(hi >> 8) & 0xFF] ^ tab[1][(hi >> 16) & 0xFF] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  *out_host = c ^ 0xFFFFFFFFu;
  return 0;
}
