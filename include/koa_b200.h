/* koa_b200.h — C-ABI of the B200-native koafusion hot path (libkoa_b200.so).
 *
 * The reference (imedslab/OAProgressionMMF, package `koafusion`) has no FFI layer: its hot path is
 * the `koafusion.models` nn.Modules calling PyTorch eager ops (cuDNN / cuBLAS). Each entry point
 * below replaces one of those implicit library calls; the reference call site it stands in for is
 * cited as file:line relative to the reference root.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure; koa_last_error() gives
 *     the message for the calling thread. No exceptions, no torch types cross this boundary.
 *   - all pointers are device pointers owned by the caller (PyTorch caching allocator in the host
 *     mirror); `stream` is a cudaStream_t passed as void*. Nothing is allocated behind the caller's
 *     back except per-process immutable state (kernel attributes, driver entry points).
 *   - activations are NHWC bf16 ("pixel-major": a 1x1 convolution is a plain GEMM), parameters
 *     arrive as fp32 master copies in PyTorch layout and are packed to bf16 by koa_pack_*.
 *   - the library is sm_100a only and has no CPU path; loading it without a B200 works (symbol
 *     checks), calling compute entry points without one fails with KOA_ERR_CUDA.
 */
#ifndef KOA_B200_H
#define KOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOA_OK 0
#define KOA_ERR_ARG (-1)
#define KOA_ERR_CUDA (-2)
#define KOA_ERR_UNSUPPORTED (-3)
#define KOA_ERR_DEVICE (-4)

/* ---- library state ------------------------------------------------------------------------- */
const char* koa_last_error(void);
int koa_version(void);
/* Reads and clears the device-side diagnostic word (non-zero: a pipeline barrier timed out). */
int koa_debug_flag(unsigned int* out);

/* ---- fused GEMM epilogue --------------------------------------------------------------------- */
enum { KOA_ACT_NONE = 0, KOA_ACT_RELU = 1, KOA_ACT_GELU = 2, KOA_ACT_GELU_GRAD = 3 };

typedef struct koa_epilogue {
  void* out;                 /* [M, ldo] bf16, or fp32 when out_fp32 != 0 */
  int ldo;                   /* leading dimension of every [M, *] epilogue tensor, in elements */
  int out_fp32;
  int act;                   /* KOA_ACT_* applied after the bias */
  const float* bias;         /* [N] or NULL */
  void* pre_out_bf16;        /* optional bf16 copy of (acc + bias) before the activation */
  const void* aux_bf16;      /* KOA_ACT_GELU_GRAD: pre-activation h; value *= gelu'(h) */
  const float* residual_f32; /* optional fp32 addend */
  const void* add_bf16;      /* optional bf16 addend ... */
  const void* mask_bf16;     /* ... gated by mask > 0 when non-NULL (ReLU'd residual gradient) */
  void* out_bf16_copy;       /* optional bf16 copy of the final value when out is fp32 */
  float* col_sum;            /* optional per-column sum / sum of squares (BatchNorm batch statistics), */
  float* col_sumsq;          /* accumulated with atomics: caller zeroes them first */
} koa_epilogue_t;

/* out[M,N] = epilogue(A[M,K] . B[N,K]^T); A, B bf16 row-major. tcgen05/TMEM tiles fed by TMA.
 * Replaces nn.Linear / F.linear (koafusion/models/_core_trf.py:104,112,145,148,161,163) and the 1x1
 * stride-1 nn.Conv2d of the bottlenecks (koafusion/models/_torchvision.py:29-31,108,112) in forward
 * and in data-gradient form. Requires K % 8 == 0 and N % 32 == 0. */
int koa_gemm_bf16(const void* a, const void* b, int m, int n, int k, const koa_epilogue_t* ep, void* stream);

/* y[N,Ho,Wo,Cout] = epilogue(conv(x[N,H,W,Cin], w[Cout,R,S,Cin])): implicit GEMM, the activation
 * operand is fetched by im2col-mode TMA. Replaces nn.Conv2d 3x3 (stride 1/2) and 1x1 stride 2
 * (koafusion/models/_torchvision.py:23-31,110,211-213). Requires Cin % 64 == 0, Cout % 32 == 0.
 * The data gradient of a stride-1 kxk convolution is the same call on dY with flipped weights. */
int koa_conv_fprop_bf16(const void* x, const void* w, int n_img, int h, int w_in, int cin, int cout, int filt_r,
                        int filt_s, int stride, int pad, const koa_epilogue_t* ep, void* stream);

/* dw[Cout,Cin] += dy[P,Cout]^T . x[P,Cin] (fp32 accumulate with atomics; caller zeroes dw).
 * Weight gradient of nn.Linear and of 1x1 stride-1 convolutions (autograd of the call sites above). */
int koa_gemm_wgrad_bf16(const void* dy, const void* x, float* dw, int pixels, int cout, int cin, void* stream);

/* dw[Cout,R,S,Cin] += conv weight gradient for dy[N,Ho,Wo,Cout], x[N,H,W,Cin]. */
int koa_conv_wgrad_bf16(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int cin, int cout,
                        int filt_r, int filt_s, int stride, int pad, void* stream);

/* Test/debug knob: MN-major shared-memory descriptor strides used by the wgrad kernels. */
int koa_debug_set_wgrad_desc(unsigned int lbo_bytes, unsigned int sbo_bytes, unsigned int k_adv_bytes);

#ifdef __cplusplus
}
#endif
#endif /* KOA_B200_H */
