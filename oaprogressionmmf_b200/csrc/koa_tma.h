// koa_tma.h — host-side construction of TMA tensor maps (bf16, 128-byte swizzle).
// The driver entry points are resolved at run time through cudaGetDriverEntryPoint,
// so the library has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <stdint.h>

// 2-D row-major bf16 matrix [outer][inner] with `pitch_bytes` between rows.
// Box = {box_inner (<= 64 elements), box_outer (<= 256 rows)}.
int koa_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer);

// im2col-mode map over an NHWC bf16 activation tensor (dims {C, W, H, N}).
// A load fetches `pixels` consecutive output pixels x 64 channels for one filter tap.
int koa_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n_img, int h, int w, int c, int filt_r, int filt_s,
                         int stride, int pad, uint32_t pixels);

// 2-D row-major 16-bit matrix viewed in {32 columns x 32 rows} boxes with the 64-byte swizzle: the staging-tile layout of
// the convolution epilogue (gemm_conv.cuh), used for its operand loads and its output store.
int koa_tmap_2d_sw64(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes);
