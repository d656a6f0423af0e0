// stem.cu — the 7x7 stride-2 stem convolution of the feature extractors, specialised for the koafusion
// input contract: every image is ONE grey channel repeated three times (einops `repeat k=3`,
// koafusion/models/_xrNmrMcP.py:211-213) and MRI volumes arrive slice-innermost (B,1,R,C,S).
//   - koa_k_stem_pack    : (B,1,R,C,S) fp32 -> [B*S][R][C] fp32 (the `rearrange "b ch r c s -> (b s) ch r c"`)
//   - koa_k_stem_fold_w  : W[64][3][7][7] -> Wfold[49][64] = sum over the three identical channels (K = 49, not 147)
//   - koa_k_stem_conv_fwd: direct fp32 convolution on CUDA cores (1 % of the network's FLOPs, K too ragged for
//                          TMA), bf16 NHWC output + BatchNorm batch statistics
//   - koa_k_stem_wgrad   : weight gradient w.r.t. the folded filter; unfold replicates it over the 3 channels
// No data gradient: the network input does not require grad.
#include "koa_common.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"

using namespace koa;

namespace {

constexpr int TILE = 16;             // 16x16 output pixels per tile
constexpr int PATCH = TILE * 2 + 5;  // 37 input rows/cols per tile
constexpr int PATCH_LD = PATCH + 1;

__global__ void stem_pack_kernel(const float* __restrict__ vol, float* __restrict__ img, int rc, int slices) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  const float* src = vol + (long long)b * rc * slices;
  float* dst = img + (long long)b * slices * rc;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < rc && s < slices) ? src[(long long)p * slices + s] : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j, p = p0 + threadIdx.x;
    if (p < rc && s < slices) dst[(long long)s * rc + p] = tile[threadIdx.x][j];
  }
}

__global__ void stem_fold_w_kernel(const float* __restrict__ w, float* __restrict__ wfold) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 49 * 64) return;
  const int tap = i / 64, co = i % 64;
  wfold[i] = w[(co * 3 + 0) * 49 + tap] + w[(co * 3 + 1) * 49 + tap] + w[(co * 3 + 2) * 49 + tap];
}

__global__ void stem_unfold_dw_kernel(const float* __restrict__ dwfold, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 3 * 49) return;
  const int tap = i % 49, co = i / (3 * 49);
  dw[i] += dwfold[tap * 64 + co];
}

// Tensor-core form of the stem: img fp32 [n][h][w] -> A fp16 [n*ho*wo][64], column k = r*7 + s holds
// img[2*oh - 3 + r][2*ow - 3 + s] (zero outside the image and for the 15 padding columns k >= 49). The convolution
// is then the plain GEMM A . Wb^T with Wb[64 co][64 k] (koa_k_stem_pack_wb), its weight gradient dy^T . A.
template <bool F16>
__global__ void stem_im2col_kernel(const float* __restrict__ img, bf16* __restrict__ a, long long total_ll, int h, int w,
                                   int ho, int wo) {
  const long long total = total_ll;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < total; i += ((long long)gridDim.x * blockDim.x)) {
    const int kg = (int)(i & 7);
    long long pix = i >> 3;
    const int ow = (int)(pix % wo); pix /= wo;
    const int oh = (int)(pix % ho);
    const long long ni = (long long)(pix / ho);
    const float* src = img + ni * h * w;
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = kg * 8 + u;
      const int r = k / 7, s = k - r * 7;
      const int iy = oh * 2 - 3 + r, ix = ow * 2 - 3 + s;
      f[u] = (k < 49 && iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(src + (long long)iy * w + ix) : 0.0f;
    }
    uint4 q;  // fp16 as the forward GEMM operand, bf16 when rebuilt for the weight gradient (pairs with bf16 dy)
    q.x = pack16x2(f[0], f[1], F16); q.y = pack16x2(f[2], f[3], F16); q.z = pack16x2(f[4], f[5], F16); q.w = pack16x2(f[6], f[7], F16);
    *reinterpret_cast<uint4*>(a + (long long)i * 8) = q;
  }
}

// W[64][3][7][7] fp32 -> Wb[64 co][64 k] fp16 = sum over the three identical input channels, zero for k >= 49
__global__ void stem_pack_wb_kernel(const float* __restrict__ w, bf16* __restrict__ wb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64) return;
  const int co = i / 64, k = i % 64;
  float v = 0.0f;
  if (k < 49) v = w[(co * 3 + 0) * 49 + k] + w[(co * 3 + 1) * 49 + k] + w[(co * 3 + 2) * 49 + k];
  reinterpret_cast<__half*>(wb)[i] = __float2half_rn(v);
}

// dWb[64 co][64 k] fp32 -> dW[64][3][7][7] += (the folded gradient reaches each of the three channels)
__global__ void stem_unfold_dwb_kernel(const float* __restrict__ dwb, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 3 * 49) return;
  const int tap = i % 49, co = i / (3 * 49);
  dw[i] += dwb[co * 64 + tap];
}

__device__ __forceinline__ void load_patch(const float* __restrict__ img, float* sp, int h, int w, int oy0, int ox0) {
  // input rows 2*oy0-3 .. 2*oy0-3+36
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = threadIdx.x; i < PATCH * PATCH; i += blockDim.x) {
    const int py = i / PATCH, px = i % PATCH;
    const int iy = iy0 + py, ix = ix0 + px;
    sp[py * PATCH_LD + px] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? img[(long long)iy * w + ix] : 0.0f;
  }
}

__global__ void __launch_bounds__(256)
stem_conv_fwd_kernel(const float* __restrict__ img, const float* __restrict__ wfold, bf16* __restrict__ y,
                     float* __restrict__ sum, float* __restrict__ sumsq, int h, int w, int ho, int wo, int tiles_x,
                     int tiles_y) {
  __shared__ float sp[PATCH * PATCH_LD];
  __shared__ __align__(16) float sw[49 * 64];
  __shared__ float sstat[128];
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int oy0 = (tile / tiles_x) * TILE, ox0 = (tile % tiles_x) * TILE;
  for (int i = threadIdx.x; i < 49 * 64; i += 256) sw[i] = wfold[i];
  if (threadIdx.x < 128) sstat[threadIdx.x] = 0.0f;
  load_patch(img + (long long)n * h * w, sp, h, w, oy0, ox0);
  __syncthreads();

  const int ty = threadIdx.x / TILE, tx = threadIdx.x % TILE;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = 0.0f;
#pragma unroll 1
  for (int r = 0; r < 7; ++r) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const float x = sp[(ty * 2 + r) * PATCH_LD + tx * 2 + s];
      const float4* wr = reinterpret_cast<const float4*>(sw + (r * 7 + s) * 64);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 wv = wr[c4];
        acc[c4 * 4 + 0] = fmaf(x, wv.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(x, wv.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(x, wv.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(x, wv.w, acc[c4 * 4 + 3]);
      }
    }
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool ok = oy < ho && ox < wo;
  if (ok) {
    uint4* dst = reinterpret_cast<uint4*>(y + (((long long)n * ho + oy) * wo + ox) * 64);
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      uint4 q;
      q.x = pack_bf16x2(acc[c], acc[c + 1]); q.y = pack_bf16x2(acc[c + 2], acc[c + 3]);
      q.z = pack_bf16x2(acc[c + 4], acc[c + 5]); q.w = pack_bf16x2(acc[c + 6], acc[c + 7]);
      dst[c / 8] = q;
    }
  }
  if (sum != nullptr) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float s[32], q[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = ok ? bf16_round(acc[half * 32 + j]) : 0.0f;
        s[j] = v;
        q[j] = v * v;
      }
      warp_transpose_reduce32(s, lane);
      warp_transpose_reduce32(q, lane);
      atomicAdd(&sstat[half * 32 + lane], s[0]);
      atomicAdd(&sstat[64 + half * 32 + lane], q[0]);
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(&sum[threadIdx.x], sstat[threadIdx.x]);
    else if (threadIdx.x < 128) atomicAdd(&sumsq[threadIdx.x - 64], sstat[threadIdx.x]);
  }
}

// Each block walks tiles with a grid stride, keeping its slice of dWfold in registers:
// thread = (channel pair cop = t % 32, tap group tg = t / 32 in [0,8)); taps tg, tg+8, ... (<= 7 of them).
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ img, const bf16* __restrict__ dy, float* __restrict__ dwfold, int n_img, int h,
                  int w, int ho, int wo, int tiles_x, int tiles_y) {
  __shared__ float sp[PATCH * PATCH_LD];
  __shared__ __align__(16) bf16 sdy[TILE * TILE * 64];
  const int cop = threadIdx.x & 31, tg = threadIdx.x >> 5;
  float acc[7][2];
  int toff[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    acc[i][0] = acc[i][1] = 0.0f;
    const int tap = tg + i * 8;
    toff[i] = tap < 49 ? (tap / 7) * PATCH_LD + (tap % 7) : -1;
  }
  const long long total_tiles = (long long)n_img * tiles_x * tiles_y;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int tile = (int)(t % (tiles_x * tiles_y));
    const int n = (int)(t / (tiles_x * tiles_y));
    const int oy0 = (tile / tiles_x) * TILE, ox0 = (tile % tiles_x) * TILE;
    __syncthreads();
    load_patch(img + (long long)n * h * w, sp, h, w, oy0, ox0);
    for (int i = threadIdx.x; i < TILE * TILE * 8; i += 256) {
      const int p = i / 8, v = i % 8;
      const int oy = oy0 + p / TILE, ox = ox0 + p % TILE;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (oy < ho && ox < wo) q = *reinterpret_cast<const uint4*>(dy + (((long long)n * ho + oy) * wo + ox) * 64 + v * 8);
      *reinterpret_cast<uint4*>(sdy + p * 64 + v * 8) = q;
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < TILE * TILE; ++p) {
      const float2 d = __bfloat1622float2(*reinterpret_cast<const bf162*>(sdy + p * 64 + cop * 2));
      const float* px = sp + ((p / TILE) * 2) * PATCH_LD + (p % TILE) * 2;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        if (toff[i] >= 0) {
          const float x = px[toff[i]];
          acc[i][0] = fmaf(x, d.x, acc[i][0]);
          acc[i][1] = fmaf(x, d.y, acc[i][1]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int tap = tg + i * 8;
    if (tap < 49) {
      atomicAdd(&dwfold[tap * 64 + cop * 2], acc[i][0]);
      atomicAdd(&dwfold[tap * 64 + cop * 2 + 1], acc[i][1]);
    }
  }
}

}  // namespace

int koa_k_stem_pack(const float* vol, float* img, int batch, int rc, int slices, cudaStream_t st) {
  dim3 grid(koa_cdiv(rc, 32), koa_cdiv(slices, 32), batch), block(32, 8);
  stem_pack_kernel<<<grid, block, 0, st>>>(vol, img, rc, slices);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_fold_w(const float* w, float* wfold, cudaStream_t st) {
  stem_fold_w_kernel<<<koa_cdiv(49 * 64, 256), 256, 0, st>>>(w, wfold);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_unfold_dw(const float* dwfold, float* dw, cudaStream_t st) {
  stem_unfold_dw_kernel<<<koa_cdiv(64 * 3 * 49, 256), 256, 0, st>>>(dwfold, dw);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_im2col(const float* img, void* a, int n, int h, int w, int f16, cudaStream_t st) {
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const long long total = (long long)n * ho * wo * 8;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (f16) stem_im2col_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(img, (bf16*)a, total, h, w, ho, wo);
  else stem_im2col_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(img, (bf16*)a, total, h, w, ho, wo);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_pack_wb(const float* w, void* wb, cudaStream_t st) {
  stem_pack_wb_kernel<<<koa_cdiv(64 * 64, 256), 256, 0, st>>>(w, (bf16*)wb);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_unfold_dwb(const float* dwb, float* dw, cudaStream_t st) {
  stem_unfold_dwb_kernel<<<koa_cdiv(64 * 3 * 49, 256), 256, 0, st>>>(dwb, dw);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_conv_fwd(const float* img, const float* wfold, void* y, float* sum, float* sumsq, int n, int h, int w,
                        cudaStream_t st) {
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const int tx = koa_cdiv(wo, TILE), ty = koa_cdiv(ho, TILE);
  const long long blocks = (long long)n * tx * ty;
  KOA_REQUIRE(blocks > 0 && blocks < 2147483647LL, "bad stem grid");
  stem_conv_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, wfold, (bf16*)y, sum, sumsq, h, w, ho, wo, tx, ty);
  KOA_LAUNCH_CHECK();
  return 0;
}
int koa_k_stem_wgrad(const float* img, const void* dy, float* dwfold, int n, int h, int w, cudaStream_t st) {
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const int tx = koa_cdiv(wo, TILE), ty = koa_cdiv(ho, TILE);
  long long blocks = (long long)n * tx * ty;
  if (blocks > 148 * 4) blocks = 148 * 4;
  stem_wgrad_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, (const bf16*)dy, dwfold, n, h, w, ho, wo, tx, ty);
  KOA_LAUNCH_CHECK();
  return 0;
}
