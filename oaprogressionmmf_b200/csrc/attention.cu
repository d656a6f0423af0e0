// attention.cu — multi-head self-attention over the token sequences of koafusion's FeaT
// (koafusion/models/_core_trf.py:167-182). Sequences are short (n <= 128 tokens: 64/32/25 slices, 92-124
// fused tokens) with head_dim 256 (D=2048, 8 heads), so one CTA owns a whole (batch, head) pair: K and V
// live in shared memory, each warp walks query rows. Attention is < 1.5 % of the transformer FLOPs
// (SURVEY.md §8 a12); the QKV / output projections around it are tcgen05 GEMMs.
//
// qkv  : bf16 [B*n][3*D], feature index = (qkv, head, d)  (the reference's `(qkv h d)` split)
// out  : bf16 [B*n][D],   feature index = (head, d)       (`b h n d -> b n (h d)`)
// probs: fp32 [B][H][n][n] softmax(QK^T * scale), returned as the reference returns `attn`
#include "koa_common.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"

using namespace koa;

namespace {

constexpr int kWarps = 32;  // one CTA per SM (K, V and dS fill the shared memory): 32 warps hide the FMA / LDS latency
constexpr int kMaxN = 128;

// FA / FB: the operand holds fp16 (forward activations of the transformer) instead of bf16 (gradients)
template <bool F>
__device__ __forceinline__ float2 up2(uint32_t v) { return F ? unpack_f16x2(v) : unpack_bf16x2(v); }
template <bool FA, bool FB>
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const float2 a0 = up2<FA>(a.x), a1 = up2<FA>(a.y), a2 = up2<FA>(a.z), a3 = up2<FA>(a.w);
  const float2 b0 = up2<FB>(b.x), b1 = up2<FB>(b.y), b2 = up2<FB>(b.z), b3 = up2<FB>(b.w);
  float s = a0.x * b0.x;
  s = fmaf(a0.y, b0.y, s); s = fmaf(a1.x, b1.x, s); s = fmaf(a1.y, b1.y, s);
  s = fmaf(a2.x, b2.x, s); s = fmaf(a2.y, b2.y, s); s = fmaf(a3.x, b3.x, s); s = fmaf(a3.y, b3.y, s);
  return s;
}
template <bool F>
__device__ __forceinline__ void fma8(float (&acc)[8], float p, const uint4& v) {
  const float2 v0 = up2<F>(v.x), v1 = up2<F>(v.y), v2 = up2<F>(v.z), v3 = up2<F>(v.w);
  acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]);
  acc[2] = fmaf(p, v1.x, acc[2]); acc[3] = fmaf(p, v1.y, acc[3]);
  acc[4] = fmaf(p, v2.x, acc[4]); acc[5] = fmaf(p, v2.y, acc[5]);
  acc[6] = fmaf(p, v3.x, acc[6]); acc[7] = fmaf(p, v3.y, acc[7]);
}
template <bool F>
__device__ __forceinline__ uint4 pack8s(const float (&f)[8], float s) {
  uint4 q;
  if (F) {
    q.x = pack_f16x2(f[0] * s, f[1] * s); q.y = pack_f16x2(f[2] * s, f[3] * s);
    q.z = pack_f16x2(f[4] * s, f[5] * s); q.w = pack_f16x2(f[6] * s, f[7] * s);
  } else {
    q.x = pack_bf16x2(f[0] * s, f[1] * s); q.y = pack_bf16x2(f[2] * s, f[3] * s);
    q.z = pack_bf16x2(f[4] * s, f[5] * s); q.w = pack_bf16x2(f[6] * s, f[7] * s);
  }
  return q;
}

// Copies the [n][hd] bf16 slab of one (batch, head, which) from qkv-like global memory into shared rows of
// `ld` elements (ld = hd + 8 gives conflict-free 16-byte row-strided reads).
__device__ __forceinline__ void stage_rows(const bf16* __restrict__ g, long long g_row_stride, bf16* s, int ld, int n,
                                           int hd) {
  const int chunks = hd / 8;
  for (int i = threadIdx.x; i < n * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i % chunks;
    *reinterpret_cast<uint4*>(s + r * ld + c * 8) = *reinterpret_cast<const uint4*>(g + r * g_row_stride + c * 8);
  }
}

template <bool F16>
__global__ void __launch_bounds__(kWarps * 32)
attention_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ probs, int n, int heads,
                     int hd, float scale) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int ldk = hd + 8;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + n * ldk;
  bf16* sQ = sV + n * hd;                                   // [kWarps][hd]
  float* sP = reinterpret_cast<float*>(sQ + kWarps * hd);   // [kWarps][kMaxN]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dmodel = heads * hd;
  const long long row_stride = 3LL * dmodel;
  const bf16* base = qkv + (long long)b * n * row_stride + h * hd;
  stage_rows(base + dmodel, row_stride, sK, ldk, n, hd);
  stage_rows(base + 2 * dmodel, row_stride, sV, hd, n, hd);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* q = sQ + warp * hd;
  float* p = sP + warp * kMaxN;
  const int chunks = hd / 8;
  for (int i = warp; i < n; i += kWarps) {
    for (int c = lane; c < chunks; c += 32)
      *reinterpret_cast<uint4*>(q + c * 8) = *reinterpret_cast<const uint4*>(base + i * row_stride + c * 8);
    __syncwarp();
    float s[kMaxN / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kMaxN / 32; ++jj) {
      const int j = jj * 32 + lane;
      float acc = 0.0f;
      if (j < n) {
        for (int c = 0; c < chunks; ++c)
          acc += dot8<F16, F16>(*reinterpret_cast<const uint4*>(q + c * 8), *reinterpret_cast<const uint4*>(sK + j * ldk + c * 8));
        acc *= scale;
        mx = fmaxf(mx, acc);
      }
      s[jj] = acc;
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int jj = 0; jj < kMaxN / 32; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = j < n ? __expf(s[jj] - mx) : 0.0f;
      sum += s[jj];
    }
    const float inv = 1.0f / warp_sum(sum);
    float* prow = probs + (((long long)b * heads + h) * n + i) * n;
#pragma unroll
    for (int jj = 0; jj < kMaxN / 32; ++jj) {
      const int j = jj * 32 + lane;
      if (j < n) {
        const float pv = s[jj] * inv;
        p[j] = pv;
        prow[j] = pv;
      }
    }
    __syncwarp();
    for (int c = lane; c < chunks; c += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < n; ++j) fma8<F16>(acc, p[j], *reinterpret_cast<const uint4*>(sV + j * hd + c * 8));
      *reinterpret_cast<uint4*>(out + ((long long)b * n + i) * dmodel + h * hd + c * 8) = pack8s<F16>(acc, 1.0f);
    }
    __syncwarp();
  }
}

// Backward: dS = P o (dP - rowsum(P o dP)), dP = dO V^T; dV = P^T dO; dQ = scale dS K; dK = scale dS^T Q.
template <bool F16>
__global__ void __launch_bounds__(kWarps * 32)
attention_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ probs, const bf16* __restrict__ dout,
                     bf16* __restrict__ dqkv, int n, int heads, int hd, float scale) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int ldk = hd + 8;
  bf16* bufA = reinterpret_cast<bf16*>(smem);               // [n][hd+8]
  bf16* bufB = bufA + n * ldk;                               // [n][hd]
  bf16* sRow = bufB + n * hd;                                // [kWarps][hd]
  float* sDS = reinterpret_cast<float*>(sRow + kWarps * hd); // [n][n+1]
  const int ldd = n + 1;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dmodel = heads * hd;
  const long long row_stride = 3LL * dmodel;
  const bf16* base = qkv + (long long)b * n * row_stride + h * hd;
  const bf16* dobase = dout + (long long)b * n * dmodel + h * hd;
  bf16* dbase = dqkv + (long long)b * n * row_stride + h * hd;
  const float* pbase = probs + ((long long)b * heads + h) * n * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = hd / 8;

  // Phase 1: dS (needs V rows for the dot products, dO row broadcast per warp)
  stage_rows(base + 2 * dmodel, row_stride, bufA, ldk, n, hd);
  __syncthreads();
  {
    bf16* r = sRow + warp * hd;
    for (int i = warp; i < n; i += kWarps) {
      for (int c = lane; c < chunks; c += 32)
        *reinterpret_cast<uint4*>(r + c * 8) = *reinterpret_cast<const uint4*>(dobase + (long long)i * dmodel + c * 8);
      __syncwarp();
      float dp[kMaxN / 32], pv[kMaxN / 32];
      float delta = 0.0f;
#pragma unroll
      for (int jj = 0; jj < kMaxN / 32; ++jj) {
        const int j = jj * 32 + lane;
        float acc = 0.0f, pij = 0.0f;
        if (j < n) {
          for (int c = 0; c < chunks; ++c)
            acc += dot8<false, F16>(*reinterpret_cast<const uint4*>(r + c * 8), *reinterpret_cast<const uint4*>(bufA + j * ldk + c * 8));
          pij = pbase[(long long)i * n + j];
          delta += pij * acc;
        }
        dp[jj] = acc;
        pv[jj] = pij;
      }
      delta = warp_sum(delta);
#pragma unroll
      for (int jj = 0; jj < kMaxN / 32; ++jj) {
        const int j = jj * 32 + lane;
        if (j < n) sDS[i * ldd + j] = pv[jj] * (dp[jj] - delta);
      }
      __syncwarp();
    }
  }
  // Phase 2: dV_j = sum_i P_ij dO_i
  stage_rows(dobase, dmodel, bufB, hd, n, hd);
  __syncthreads();
  for (int j = warp; j < n; j += kWarps) {
    for (int c = lane; c < chunks; c += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < n; ++i) fma8<false>(acc, pbase[(long long)i * n + j], *reinterpret_cast<const uint4*>(bufB + i * hd + c * 8));
      *reinterpret_cast<uint4*>(dbase + j * row_stride + 2 * dmodel + c * 8) = pack8s<false>(acc, 1.0f);
    }
  }
  __syncthreads();
  // Phase 3: dQ_i = scale * sum_j dS_ij K_j
  stage_rows(base + dmodel, row_stride, bufB, hd, n, hd);
  __syncthreads();
  for (int i = warp; i < n; i += kWarps) {
    for (int c = lane; c < chunks; c += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < n; ++j) fma8<F16>(acc, sDS[i * ldd + j], *reinterpret_cast<const uint4*>(bufB + j * hd + c * 8));
      *reinterpret_cast<uint4*>(dbase + i * row_stride + c * 8) = pack8s<false>(acc, scale);
    }
  }
  __syncthreads();
  // Phase 4: dK_j = scale * sum_i dS_ij Q_i
  stage_rows(base, row_stride, bufB, hd, n, hd);
  __syncthreads();
  for (int j = warp; j < n; j += kWarps) {
    for (int c = lane; c < chunks; c += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < n; ++i) fma8<F16>(acc, sDS[i * ldd + j], *reinterpret_cast<const uint4*>(bufB + i * hd + c * 8));
      *reinterpret_cast<uint4*>(dbase + j * row_stride + dmodel + c * 8) = pack8s<false>(acc, scale);
    }
  }
}

size_t fwd_smem(int n, int hd) {
  return (size_t)n * (hd + 8) * 2 + (size_t)n * hd * 2 + (size_t)kWarps * hd * 2 + (size_t)kWarps * kMaxN * 4;
}
size_t bwd_smem(int n, int hd) {
  return (size_t)n * (hd + 8) * 2 + (size_t)n * hd * 2 + (size_t)kWarps * hd * 2 + (size_t)n * (n + 1) * 4;
}

// KOA_ATTN_TC (default 1): the tcgen05 kernels of attention_tc.cu wherever the head shape allows; 0 = the CUDA-core kernels
// of this file everywhere (they remain the path for head dimensions that are not a multiple of 64)
bool attn_tc_enabled() {
  static const int v = [] {
    const char* e = getenv("KOA_ATTN_TC");
    return e == nullptr ? 1 : atoi(e);
  }();
  return v != 0;
}

int check(int n, int hd) {
  KOA_REQUIRE(n >= 1 && n <= kMaxN, "attention supports 1..%d tokens per sequence (got %d)", kMaxN, n);
  KOA_REQUIRE(hd % 8 == 0 && hd >= 8 && hd <= 256, "attention head_dim must be a multiple of 8, <= 256 (got %d)", hd);
  return 0;
}

}  // namespace

int koa_k_attention_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                        cudaStream_t st, int f16) {
  int rc = check(n, head_dim);
  if (rc) return rc;
  if (attn_tc_enabled() && koa_attention_tc_ok(n, head_dim))
    return koa_k_attention_tc_fwd(qkv, out, probs, batch, n, heads, head_dim, scale, st, f16);
  static std::atomic<unsigned long long> attr_done{0}, attr_done_h{0};
  if (f16) {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_fwd_kernel<true>, (int)fwd_smem(kMaxN, 256), attr_done_h));
    attention_fwd_kernel<true><<<batch * heads, kWarps * 32, fwd_smem(n, head_dim), st>>>((const bf16*)qkv, (bf16*)out, probs,
                                                                                          n, heads, head_dim, scale);
  } else {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_fwd_kernel<false>, (int)fwd_smem(kMaxN, 256), attr_done));
    attention_fwd_kernel<false><<<batch * heads, kWarps * 32, fwd_smem(n, head_dim), st>>>((const bf16*)qkv, (bf16*)out, probs,
                                                                                           n, heads, head_dim, scale);
  }
  KOA_LAUNCH_CHECK();
  return 0;
}

int koa_k_attention_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                        int head_dim, float scale, cudaStream_t st, int f16) {
  int rc = check(n, head_dim);
  if (rc) return rc;
  if (attn_tc_enabled() && koa_attention_tc_ok(n, head_dim))
    return koa_k_attention_tc_bwd(qkv, probs, dout, dqkv, batch, n, heads, head_dim, scale, st, f16);
  static std::atomic<unsigned long long> attr_done{0}, attr_done_h{0};
  if (f16) {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_bwd_kernel<true>, (int)bwd_smem(kMaxN, 256), attr_done_h));
    attention_bwd_kernel<true><<<batch * heads, kWarps * 32, bwd_smem(n, head_dim), st>>>(
        (const bf16*)qkv, probs, (const bf16*)dout, (bf16*)dqkv, n, heads, head_dim, scale);
  } else {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_bwd_kernel<false>, (int)bwd_smem(kMaxN, 256), attr_done));
    attention_bwd_kernel<false><<<batch * heads, kWarps * 32, bwd_smem(n, head_dim), st>>>(
        (const bf16*)qkv, probs, (const bf16*)dout, (bf16*)dqkv, n, heads, head_dim, scale);
  }
  KOA_LAUNCH_CHECK();
  return 0;
}
