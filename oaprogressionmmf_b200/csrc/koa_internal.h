// koa_internal.h — launchers shared between translation units of libkoa_b200.so.
// (The public C ABI is include/koa_b200.h; these take a typed cudaStream_t.)
#pragma once
#include <cuda_runtime.h>

#include "../../include/koa_b200.h"

int koa_num_sms();
// algorithmic share of the FLOPs of this thread's next GEMM launches in the per-launch profile (returns the previous value)
double koa_profile_flop_scale(double scale);
struct KoaFlopScale {  // RAII form
  double old;
  explicit KoaFlopScale(double s) : old(koa_profile_flop_scale(s)) {}
  ~KoaFlopScale() { koa_profile_flop_scale(old); }
};

int koa_gemm_launch(const void* a, const void* b, int m, int n, int k, const koa_epilogue_t* ep, cudaStream_t st);
int koa_gemm_kcat_launch(const void* a1, int k1, const void* a2, int k2, const void* b, int m, int n,
                         const koa_epilogue_t* ep, cudaStream_t st);
int koa_conv_fprop_launch(const void* x, const void* w, int n_img, int h, int w_in, int cin, int cout, int filt_r,
                          int filt_s, int stride, int pad, const koa_epilogue_t* ep, cudaStream_t st);
int koa_gemm_wgrad_launch(const void* dy, const void* x, float* dw, int pixels, int cout, int cin, int x_f16,
                          cudaStream_t st);
int koa_conv_wgrad_launch(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int cin, int cout,
                          int filt_r, int filt_s, int stride, int pad, int x_f16, cudaStream_t st);
int koa_conv_grouped_launch(const void* x, const void* w, int n_img, int h, int w_in, int c, int stride,
                            const koa_epilogue_t* ep, cudaStream_t st);
int koa_conv_grouped_wgrad_launch(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int c,
                                  int stride, int x_f16, cudaStream_t st);
