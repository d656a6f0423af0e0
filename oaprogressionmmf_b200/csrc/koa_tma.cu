// koa_tma.cu — TMA tensor-map encoding + library-wide error state.
#include "koa_tma.h"

#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "koa_common.cuh"

static thread_local char t_koa_error[1024] = "";

void koa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_koa_error, sizeof(t_koa_error), fmt, ap);
  va_end(ap);
}

extern "C" const char* koa_last_error(void) { return t_koa_error; }

// Number of kernels this library has launched in this process (every launcher counts itself).
static std::atomic<long long> s_launches{0};
void koa_count_launch() { s_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long koa_launch_count(void) { return s_launches.load(std::memory_order_relaxed); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled s_encode_tiled = nullptr;
static PFN_encodeIm2col s_encode_im2col = nullptr;
static int s_driver_version = 0;
static std::once_flag s_once;

static void resolve_entry_points() {
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    s_encode_tiled = (PFN_encodeTiled)fn;
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    s_encode_im2col = (PFN_encodeIm2col)fn;
  cudaDriverGetVersion(&s_driver_version);
}

static int tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                         uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle);

int koa_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  return tmap_2d_16bit(out, base, inner, outer, pitch_bytes, box_inner, box_outer, CU_TENSOR_MAP_SWIZZLE_128B);
}

int koa_tmap_2d_sw64(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes) {
  return tmap_2d_16bit(out, base, inner, outer, pitch_bytes, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
}

static int tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                         uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  std::call_once(s_once, resolve_entry_points);
  KOA_REQUIRE(s_encode_tiled != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  KOA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA base address must be 16-byte aligned");
  KOA_REQUIRE((pitch_bytes & 15) == 0, "TMA row pitch (%llu B) must be a multiple of 16", (unsigned long long)pitch_bytes);
  KOA_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256, "TMA box too large");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = s_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    koa_set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner,
                  box_outer);
    return KOA_ERR_CUDA;
  }
  return 0;
}

int koa_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n_img, int h, int w, int c, int filt_r, int filt_s,
                         int stride, int pad, uint32_t pixels) {
  std::call_once(s_once, resolve_entry_points);
  KOA_REQUIRE(s_encode_im2col != nullptr, "cuTensorMapEncodeIm2col not available from the driver");
  KOA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA base address must be 16-byte aligned");
  KOA_REQUIRE(c % 8 == 0, "im2col TMA needs C %% 8 == 0 (got %d)", c);
  KOA_REQUIRE(pixels <= 256, "im2col TMA pixelsPerColumn <= 256");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_img};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * w, (cuuint64_t)c * 2 * w * h};
  // Bounding box of base pixels: [lower, dim + upper - 1]; the filter tap is added as an offset.
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (filt_s - 1), pad - (filt_r - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  const cuuint32_t ch_per_pixel = c < 64 ? (cuuint32_t)c : 64u;
  CUresult r = s_encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, ch_per_pixel, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    koa_set_error("cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d R=%d S=%d stride=%d pad=%d", (int)r, n_img,
                  h, w, c, filt_r, filt_s, stride, pad);
    return KOA_ERR_CUDA;
  }
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (the same
  // adjustment CUTLASS applies in make_im2col_tma_copy_desc).
  const uint64_t total_bytes = (uint64_t)n_img * h * w * c * 2;
  if (s_driver_version <= 13010 && total_bytes < 131072) {
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  }
  return 0;
}

int koa_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    // experiments: persistent GEMM grids smaller than the machine leave SMs to the kernels of concurrent streams
    const char* e = getenv("KOA_NUM_SMS");
    if (e != nullptr && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
  }
  return n;
}
