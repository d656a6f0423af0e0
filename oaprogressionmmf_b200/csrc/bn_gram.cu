// bn_gram.cu — the small kernels of the y-free bottleneck tail.
//
// The last 1x1 convolution of a bottleneck (koafusion/models/_torchvision.py:130-136: conv3 -> bn3 -> += identity -> relu)
// is the widest tensor of the block (4x the channels of conv1 / conv2). In train mode its BatchNorm needs the batch
// statistics of y3 = a2 . W3^T before a single normalised value can be produced, which used to cost one write and two
// reads of y3 plus a separate BatchNorm pass in each direction. y3 is linear in the narrow a2, so everything BatchNorm
// needs from it follows from three small quantities of a2 (N pixels, w channels; C = 4w output channels):
//     s = colsum(a2) [w],   Gram = a2^T a2 [w, w]  (one weight-gradient-shaped tensor-core GEMM),   Q = W3 . Gram / N [C, w]
//   forward :  mean_c = W3[c] . s / N,   E[y3_c^2] = Q[c] . W3[c]      -> scale / shift; the conv3 GEMM applies them (+ residual,
//              ReLU) in its epilogue (gemm_conv.cuh MODE 2) and y3 never exists in HBM;
//   backward:  with G = dL/d(pre-ReLU sum) and T = G^T a2 [C, w] (the ordinary weight-gradient GEMM, on G):
//              sum(G * y3)_c = T[c] . W3[c]           -> dgamma, dbeta and the coefficients of dy3 = k0 G - k1 - k2 y3
//              dW3   = k0 * T - k1 (x) s - k2 * N * Q
//              d(a2) = [G | a2] . [k0 * W3 ; -W3^T diag(k2) W3] - k1 . W3     (ONE K-concatenated GEMM + column bias)
// tests/test_gram_bn_algebra.py pins these formulas against autograd in fp64.
#include "koa_common.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"

using namespace koa;

namespace {

// Gram / N as an fp16 head + an fp16 tail (22 significant bits): the operands of Q = W3 . (head + tail) on the tensor cores
__global__ void gram_split_kernel(const float* __restrict__ gram, __half* __restrict__ hi, __half* __restrict__ lo, int n,
                                  float inv_count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = gram[i] * inv_count;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

// One warp per output channel c: batch statistics of y3[:, c] from s and Q, then the BatchNorm coefficients
// (as bn_finalize_kernel: running statistics with momentum 0.1 and the unbiased variance).
__global__ void bn_gram_stats_kernel(const __half* __restrict__ w3h, const float* __restrict__ sa2, const float* __restrict__ q,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ scale,
                                     float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                     int c_out, int w, double count) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= c_out) return;
  const __half* wr = w3h + (size_t)c * w;
  const float* qr = q + (size_t)c * w;
  double m = 0.0, e2 = 0.0;
  for (int j = lane; j < w; j += 32) {
    const double wv = (double)__half2float(wr[j]);
    m += wv * (double)sa2[j];
    e2 += wv * (double)qr[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m += __shfl_xor_sync(0xffffffffu, m, o);
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  }
  if (lane != 0) return;
  const double mean = m / count;
  double var = e2 - mean * mean;  // Q was built from Gram / N: e2 = E[y^2]
  if (var < 0.0) var = 0.0;
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  run_mean[c] = (1.0f - 0.1f) * run_mean[c] + 0.1f * (float)mean;
  run_var[c] = (1.0f - 0.1f) * run_var[c] + 0.1f * (float)unbiased;
  const float invstd = rsqrtf((float)var + 1e-5f);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
}

// Backward coefficients. One block = 8 warps = 8 consecutive output channels c0 .. c0 + 7; warp i owns channel c0 + i:
//   sdzy = T[c] . W3h[c];  dgamma = invstd (sdzy - mean * sdz);  dbeta = sdz;  k0, k1, k2;  dW3[c] += k0 T[c] - k1 s - k2 N Q[c]
// and the two operands derived from W3 that the data-gradient GEMM and the M GEMM read with the OUTPUT channel innermost
// (wext[j][c] = bf16(k0_c W3[c][j]), k2w[j][c] = bf16(-k2_c W3[c][j])): staged as [w][8 channels] in shared memory and
// written as 16-byte pieces.
constexpr int kCoefWarps = 8;
__global__ void __launch_bounds__(kCoefWarps * 32)
bn_gram_bwd_kernel(const float* __restrict__ t, const __half* __restrict__ w3h, const float* __restrict__ w3m,
                   const float* __restrict__ sa2, const float* __restrict__ q, const float* __restrict__ sdz,
                   const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw, bf16* __restrict__ wext,
                   bf16* __restrict__ k2w, float* __restrict__ k0_out, float* __restrict__ k1_out, float* __restrict__ k2_out,
                   int c_out, int w, int ld_ext, double count) {
  extern __shared__ uint4 s_tile[];  // [2][w] pieces of 8 bf16 (one per channel of the block)
  bf16* s_ext = reinterpret_cast<bf16*>(s_tile);
  bf16* s_k2w = s_ext + (size_t)w * 8;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCoefWarps;
  const int c = c0 + wi;  // c_out is a multiple of 8 (host check)
  const float* tr = t + (size_t)c * w;
  const __half* wr = w3h + (size_t)c * w;
  double acc = 0.0;
  for (int j = lane; j < w; j += 32) acc += (double)tr[j] * (double)__half2float(wr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float is = invstd[c], mu = mean[c], g = gamma[c], sd = sdz[c];
  const float dga = is * (float)(acc - (double)mu * (double)sd);
  const float inv_n = (float)(1.0 / count);
  const float k0 = g * is;
  const float k2 = k0 * is * dga * inv_n;
  const float k1 = k0 * sd * inv_n - k2 * mu;
  if (lane == 0) {
    if (dgamma != nullptr) dgamma[c] += dga;
    if (dbeta != nullptr) dbeta[c] += sd;
    k0_out[c] = k0; k1_out[c] = k1; k2_out[c] = k2;
  }
  const float k2n = k2 * (float)count;
  const float* wm = w3m + (size_t)c * w;
  const float* qr = q + (size_t)c * w;
  for (int j = lane; j < w; j += 32) {
    const float wv = wm[j];
    if (dw != nullptr) dw[(size_t)c * w + j] += k0 * tr[j] - k1 * sa2[j] - k2n * qr[j];
    s_ext[j * 8 + wi] = __float2bfloat16_rn(k0 * wv);
    s_k2w[j * 8 + wi] = __float2bfloat16_rn(-k2 * wv);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < w; j += kCoefWarps * 32) {
    *reinterpret_cast<uint4*>(wext + (size_t)j * ld_ext + c0) = s_tile[j];
    *reinterpret_cast<uint4*>(k2w + (size_t)j * c_out + c0) = s_tile[w + j];
  }
}

// bias[j] = -(k1 . W3)[j] from the data-gradient form of the weights (W3^T, bf16 [w][C]); one warp per j
__global__ void bn_gram_bias_kernel(const bf16* __restrict__ w3t, const float* __restrict__ k1, float* __restrict__ bias, int w,
                                    int c_out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= w) return;
  const bf16* row = w3t + (size_t)j * c_out;
  float acc = 0.f;
  for (int c = lane * 2; c < c_out; c += 64) {
    const float2 wv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + c));
    acc = fmaf(k1[c], wv.x, acc);
    acc = fmaf(k1[c + 1], wv.y, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) bias[j] = -acc;
}

}  // namespace

int koa_k_gram_split(const float* gram, void* hi, void* lo, int w, double count, cudaStream_t st) {
  const int n = w * w;
  gram_split_kernel<<<koa_cdiv(n, 256) > 592 ? 592 : koa_cdiv(n, 256), 256, 0, st>>>(gram, (__half*)hi, (__half*)lo, n,
                                                                                      (float)(1.0 / count));
  KOA_LAUNCH_CHECK();
  return 0;
}

int koa_k_bn_gram_stats(const void* w3h, const float* sa2, const float* q, const float* gamma, const float* beta,
                        float* run_mean, float* run_var, float* scale, float* shift, float* mean, float* invstd, int c_out,
                        int w, double count, cudaStream_t st) {
  bn_gram_stats_kernel<<<koa_cdiv(c_out, 8), 256, 0, st>>>((const __half*)w3h, sa2, q, gamma, beta, run_mean, run_var, scale,
                                                           shift, mean, invstd, c_out, w, count);
  KOA_LAUNCH_CHECK();
  return 0;
}

int koa_k_bn_gram_bwd(const float* t, const void* w3h, const float* w3m, const float* sa2, const float* q, const float* sdz,
                      const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta, float* dw,
                      void* wext, void* k2w, float* k0, float* k1, float* k2, int c_out, int w, int ld_ext, double count,
                      cudaStream_t st) {
  KOA_REQUIRE(c_out % 8 == 0 && ld_ext % 8 == 0 && w <= 1536, "bn_gram_bwd: C %% 8 == 0, ld %% 8 == 0, w <= 1536 (got %d, %d, %d)",
              c_out, ld_ext, w);
  bn_gram_bwd_kernel<<<c_out / kCoefWarps, kCoefWarps * 32, (size_t)w * 32, st>>>(
      t, (const __half*)w3h, w3m, sa2, q, sdz, gamma, mean, invstd, dgamma, dbeta, dw, (bf16*)wext, (bf16*)k2w, k0, k1, k2,
      c_out, w, ld_ext, count);
  KOA_LAUNCH_CHECK();
  return 0;
}

int koa_k_bn_gram_bias(const void* w3t, const float* k1, float* bias, int w, int c_out, cudaStream_t st) {
  KOA_REQUIRE(c_out % 2 == 0, "bn_gram_bias: even channel count");
  bn_gram_bias_kernel<<<koa_cdiv(w, 8), 256, 0, st>>>((const bf16*)w3t, k1, bias, w, c_out);
  KOA_LAUNCH_CHECK();
  return 0;
}
