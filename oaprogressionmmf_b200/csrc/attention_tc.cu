// attention_tc.cu — multi-head self-attention of koafusion's FeaT (koafusion/models/_core_trf.py:167-182) on tcgen05.
//
// One CTA owns one (batch, head): n <= 128 tokens, head_dim hd in {64, 128, 192, 256} (D = 2048, 8 heads: hd = 256), so
// every contraction is ONE UMMA tile:
//   forward :  S = Q K^T  (128 x 128 x hd, K-major operands)        -> TMEM columns [0, 128)
//              softmax over the row held by one thread (tcgen05.ld), probabilities to global memory in fp32 (the
//              reference returns them) and, as 16-bit, to shared memory in the 128-byte-swizzle layout of an A operand
//              O = P V    (128 x hd x 128, V as an MN-major B operand)   -> TMEM columns [128, 128 + hd)
//   backward:  dP = dO V^T                                          -> TMEM [0, 128)
//              dS = P o (dP - rowsum(P o dP)) per row; P and dS (bf16) to shared memory
//              dV = P^T dO   (MN-major A and B: a weight-gradient-shaped contraction over the query rows) -> TMEM [256, 512)
//              dQ = scale dS K  (K as MN-major B)                     -> TMEM [0, 256)
//              dK = scale dS^T Q                                      -> TMEM [256, 512)
// Operand tiles are fetched by TMA straight from the packed qkv / dO matrices as [128 rows][64 columns] boxes in the
// 128-byte swizzle; ONE copy of a tile serves as a K-major operand (contraction over its columns) and as an MN-major operand
// (contraction over its rows): only the descriptor differs. The forward operands are fp16 (KOA_FEAT_F16) or bf16; gradients
// are bf16, so the backward converts V, K and Q to bf16 in shared memory (a tcgen05 kind::f16 MMA takes one format).
//
// qkv  : 16-bit [B*n][3*D], feature index = (qkv, head, d);  out: 16-bit [B*n][D];  probs: fp32 [B][H][n][n]
// dout : bf16 [B*n][D];  dqkv: bf16 [B*n][3*D]
// Rows past n inside a 128-row box belong to the next sequence (or are zero-filled past the end of the matrix): their
// scores are masked, their probabilities written as exact zeros, so they never contribute.
#include "gemm_tc.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"
#include "koa_tma.h"

using namespace koa;

namespace {

constexpr int kAttThreads = 192;       // warp 0: TMA + MMA issue (one thread); warp 1: TMEM; warps 2..5: one thread per row
constexpr uint32_t kTile = 128 * 128;  // bytes of one [128 rows][64 x 16-bit] tile
constexpr uint32_t kRegion = 4 * kTile;  // one operand at hd = 256

struct AttDesc {
  uint32_t lbo_mn;  // MN-major descriptors: byte stride between 64-element groups along MN (tiles are kTile apart)
};

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) { return umma_desc_sw128(addr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) { return umma_desc_sw128(addr, lbo, 1024); }

// 16-byte piece `ch` of row `row` inside a [128][128 B] swizzled tile
__device__ __forceinline__ uint32_t sw_off(int row, int ch) { return (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4)); }

template <bool F16>
__device__ __forceinline__ uint32_t pk(float a, float b) { return F16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); }

// fp16 -> bf16 in place over `bytes` of shared memory, `nthreads` cooperating threads with index t
__device__ __forceinline__ void cvt_region(uint32_t base, uint32_t bytes, int t, int nthreads) {
  for (uint32_t o = (uint32_t)t * 16; o < bytes; o += (uint32_t)nthreads * 16) {
    uint4 v = lds128(base + o);
    v.x = f16x2_to_bf16x2(v.x); v.y = f16x2_to_bf16x2(v.y); v.z = f16x2_to_bf16x2(v.z); v.w = f16x2_to_bf16x2(v.w);
    sts128(base + o, v);
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ probs, int n,
                        int heads, int hd, float scale, AttDesc ad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base0 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base0 & 1023u)) & 1023u);
  uint8_t* sQ = smem;                  // [hd / 64] tiles; later P: [2] tiles
  uint8_t* sK = smem + kRegion;
  uint8_t* sV = smem + 2 * kRegion;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kRegion);  // qk, v, s, p, o
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dmodel = heads * hd;
  const int kt = hd / 64;              // 64-column tiles per operand
  const int nkb = (n + 63) / 64;       // 64-token blocks that hold a real token

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 4); mbar_init(&bars[4], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = b * n;
      mbar_arrive_expect_tx(&bars[0], 2u * kt * kTile);
      for (int c = 0; c < kt; ++c) {
        tma_load_2d(sQ + c * kTile, &tmQKV, &bars[0], h * hd + c * 64, row0);
        tma_load_2d(sK + c * kTile, &tmQKV, &bars[0], dmodel + h * hd + c * 64, row0);
      }
      mbar_arrive_expect_tx(&bars[1], (uint32_t)kt * kTile);
      for (int c = 0; c < kt; ++c) tma_load_2d(sV + c * kTile, &tmQKV, &bars[1], 2 * dmodel + h * hd + c * 64, row0);
      // S = Q K^T
      mbar_wait(&bars[0], 0, 0xe00);
      tc_fence_after();
      const uint32_t id_s = umma_idesc_16(128, 128, 0, 0, F16, F16);
      for (int c = 0; c < kt; ++c) {
        const uint64_t ad_ = desc_k(smem_u32(sQ + c * kTile)), bd_ = desc_k(smem_u32(sK + c * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, ad_ + (uint64_t)(k * 2), bd_ + (uint64_t)(k * 2), id_s, (c | k) != 0);
      }
      umma_commit(&bars[2]);
      // O = P V: contraction over the tokens (rows of V): V is an MN-major B operand
      mbar_wait(&bars[1], 0, 0xe01);
      mbar_wait(&bars[3], 0, 0xe03);
      tc_fence_after();
      const uint32_t id_o = umma_idesc_16(128, hd, 0, 1, F16, F16);
      for (int t = 0; t < nkb; ++t) {
        const uint64_t ad_ = desc_k(smem_u32(sQ + t * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bd_ = desc_mn(smem_u32(sV) + (uint32_t)t * 8192u + (uint32_t)k * 2048u, ad.lbo_mn);
          umma_bf16_ss(tmem + 128, ad_ + (uint64_t)(k * 2), bd_, id_o, (t | k) != 0);
        }
      }
      umma_commit(&bars[4]);
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int i = q * 32 + lane;  // query row of this thread
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    mbar_wait(&bars[2], 0, 0xe02);
    tc_fence_after();
    float s[128];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + (uint32_t)c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) s[c * 32 + j] = __uint_as_float(r[j]);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      s[j] = j < n ? s[j] * scale : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      s[j] = j < n ? __expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    const float inv = i < n ? 1.f / sum : 0.f;  // rows past the sequence: exact zeros
#pragma unroll
    for (int j = 0; j < 128; ++j) s[j] *= inv;
    if (i < n) {
      float* prow = probs + (((long long)b * heads + h) * n + i) * n;
      if ((n & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 128; j += 4)
          if (j < n) *reinterpret_cast<float4*>(prow + j) = make_float4(s[j], s[j + 1], s[j + 2], s[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (j < n) prow[j] = s[j];
      }
    }
    // P as the A operand of the second MMA (Q is dead: S is complete)
    const uint32_t pbase = smem_u32(sQ);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const int j0 = t * 64 + ch * 8;
        sts128(pbase + (uint32_t)t * kTile + sw_off(i, ch),
               make_uint4(pk<F16>(s[j0], s[j0 + 1]), pk<F16>(s[j0 + 2], s[j0 + 3]), pk<F16>(s[j0 + 4], s[j0 + 5]),
                          pk<F16>(s[j0 + 6], s[j0 + 7])));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[3]);
    mbar_wait(&bars[4], 0, 0xe04);
    tc_fence_after();
    bf16* orow = out + ((long long)b * n + i) * dmodel + h * hd;
    for (int c = 0; c < hd / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + 128u + (uint32_t)c * 32, r);
      tmem_ld_wait();
      if (i < n) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) =
              make_uint4(pk<F16>(__uint_as_float(r[u * 8 + 0]), __uint_as_float(r[u * 8 + 1])),
                         pk<F16>(__uint_as_float(r[u * 8 + 2]), __uint_as_float(r[u * 8 + 3])),
                         pk<F16>(__uint_as_float(r[u * 8 + 4]), __uint_as_float(r[u * 8 + 5])),
                         pk<F16>(__uint_as_float(r[u * 8 + 6]), __uint_as_float(r[u * 8 + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---- backward --------------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const float* __restrict__ probs, bf16* __restrict__ dqkv, int n, int heads, int hd, float scale,
                        AttDesc ad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base0 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base0 & 1023u)) & 1023u);
  uint8_t* r0 = smem;                 // dO, later Q
  uint8_t* r1 = smem + kRegion;       // V, later K
  uint8_t* r2 = smem + 2 * kRegion;   // P [2 tiles] + dS [2 tiles]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kRegion);
  enum { B_LD0 = 0, B_CV0, B_DP, B_LD1, B_PS, B_DV, B_DQ, B_DVDONE, B_LD2, B_CV2, B_DK, B_COUNT };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dmodel = heads * hd;
  const int kt = hd / 64;
  const int nkb = (n + 63) / 64;
  const uint32_t op_bytes = (uint32_t)kt * kTile;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < B_COUNT; ++i) {
      const bool workers = i == B_CV0 || i == B_PS || i == B_DVDONE || i == B_CV2;
      mbar_init(&bars[i], workers ? 4 : 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int row0 = b * n;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[B_LD0], 2 * op_bytes);
      for (int c = 0; c < kt; ++c) {
        tma_load_2d(r0 + c * kTile, &tmDO, &bars[B_LD0], h * hd + c * 64, row0);
        tma_load_2d(r1 + c * kTile, &tmQKV, &bars[B_LD0], 2 * dmodel + h * hd + c * 64, row0);
      }
      // dP = dO V^T (bf16 x bf16: the workers convert V first when the forward format is fp16)
      mbar_wait(&bars[B_CV0], 0, 0xe10);
      tc_fence_after();
      const uint32_t id_kk = umma_idesc_16(128, 128, 0, 0, 0, 0);
      for (int c = 0; c < kt; ++c) {
        const uint64_t ad_ = desc_k(smem_u32(r0 + c * kTile)), bd_ = desc_k(smem_u32(r1 + c * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, ad_ + (uint64_t)(k * 2), bd_ + (uint64_t)(k * 2), id_kk, (c | k) != 0);
      }
      umma_commit(&bars[B_DP]);
      mbar_wait(&bars[B_DP], 0, 0xe11);  // V is dead: K takes its place
      mbar_arrive_expect_tx(&bars[B_LD1], op_bytes);
      for (int c = 0; c < kt; ++c) tma_load_2d(r1 + c * kTile, &tmQKV, &bars[B_LD1], dmodel + h * hd + c * 64, row0);
      mbar_wait(&bars[B_PS], 0, 0xe12);  // P, dS written, K converted
      tc_fence_after();
      // dV = P^T dO: contraction over the query rows i; A = P (MN-major: M = j), B = dO (MN-major: N = d)
      const uint32_t id_mm = umma_idesc_16(128, hd, 1, 1, 0, 0);
      for (int t = 0; t < nkb; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t adv = (uint32_t)t * 8192u + (uint32_t)k * 2048u;
          umma_bf16_ss(tmem + 256, desc_mn(smem_u32(r2) + adv, ad.lbo_mn), desc_mn(smem_u32(r0) + adv, ad.lbo_mn), id_mm,
                       (t | k) != 0);
        }
      }
      umma_commit(&bars[B_DV]);
      // dQ = dS K: contraction over the key rows j; A = dS (K-major), B = K (MN-major)
      const uint32_t id_km = umma_idesc_16(128, hd, 0, 1, 0, 0);
      for (int t = 0; t < nkb; ++t) {
        const uint64_t ad_ = desc_k(smem_u32(r2 + 2 * kTile + t * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem, ad_ + (uint64_t)(k * 2),
                       desc_mn(smem_u32(r1) + (uint32_t)t * 8192u + (uint32_t)k * 2048u, ad.lbo_mn), id_km, (t | k) != 0);
      }
      umma_commit(&bars[B_DQ]);
      mbar_wait(&bars[B_DVDONE], 0, 0xe13);  // dV drained (so the dV MMAs have read dO): Q takes dO's place
      mbar_arrive_expect_tx(&bars[B_LD2], op_bytes);
      for (int c = 0; c < kt; ++c) tma_load_2d(r0 + c * kTile, &tmQKV, &bars[B_LD2], h * hd + c * 64, row0);
      mbar_wait(&bars[B_CV2], 0, 0xe14);
      tc_fence_after();
      // dK = dS^T Q: contraction over i; A = dS (MN-major: M = j), B = Q (MN-major)
      for (int t = 0; t < nkb; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t adv = (uint32_t)t * 8192u + (uint32_t)k * 2048u;
          umma_bf16_ss(tmem + 256, desc_mn(smem_u32(r2 + 2 * kTile) + adv, ad.lbo_mn), desc_mn(smem_u32(r0) + adv, ad.lbo_mn),
                       id_mm, (t | k) != 0);
        }
      }
      umma_commit(&bars[B_DK]);
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int i = q * 32 + lane;
    const int wt = (warp - 2) * 32 + lane;  // 0..127
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    auto arrive = [&](int bar) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[bar]);
    };
    mbar_wait(&bars[B_LD0], 0, 0xe20);
    if (F16) cvt_region(smem_u32(r1), op_bytes, wt, 128);
    arrive(B_CV0);
    // dS = P o (dP - rowsum(P o dP)): two passes over the row (dP from TMEM, P from global / L1), 32 columns at a time
    const float* prow = probs + (((long long)b * heads + h) * n + (i < n ? i : 0)) * n;
    mbar_wait(&bars[B_DP], 0, 0xe21);
    tc_fence_after();
    float delta = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + (uint32_t)c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pj = (i < n && c * 32 + j < n) ? __ldg(prow + c * 32 + j) : 0.f;
        delta = fmaf(pj, __uint_as_float(r[j]), delta);
      }
    }
    const uint32_t pb = smem_u32(r2);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + (uint32_t)c * 32, r);
      tmem_ld_wait();
      float p[32], ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        p[j] = (i < n && c * 32 + j < n) ? __ldg(prow + c * 32 + j) : 0.f;  // masked rows / columns: exact zeros
        ds[j] = p[j] * (__uint_as_float(r[j]) - delta);
      }
      const int t = c >> 1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ch = (c & 1) * 4 + u;
        sts128(pb + (uint32_t)t * kTile + sw_off(i, ch),
               make_uint4(pack_bf16x2(p[u * 8 + 0], p[u * 8 + 1]), pack_bf16x2(p[u * 8 + 2], p[u * 8 + 3]),
                          pack_bf16x2(p[u * 8 + 4], p[u * 8 + 5]), pack_bf16x2(p[u * 8 + 6], p[u * 8 + 7])));
        sts128(pb + (uint32_t)(2 + t) * kTile + sw_off(i, ch),
               make_uint4(pack_bf16x2(ds[u * 8 + 0], ds[u * 8 + 1]), pack_bf16x2(ds[u * 8 + 2], ds[u * 8 + 3]),
                          pack_bf16x2(ds[u * 8 + 4], ds[u * 8 + 5]), pack_bf16x2(ds[u * 8 + 6], ds[u * 8 + 7])));
      }
    }
    mbar_wait(&bars[B_LD1], 0, 0xe22);
    if (F16) cvt_region(smem_u32(r1), op_bytes, wt, 128);
    arrive(B_PS);
    // drain one [128][hd] accumulator into rows of dqkv (this thread: row i of the sequence), scaled
    auto drain = [&](uint32_t col0, int which, float mul) {
      bf16* drow = dqkv + ((long long)b * n + i) * (3LL * dmodel) + (long long)which * dmodel + h * hd;
      for (int c = 0; c < hd / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(lane_addr + col0 + (uint32_t)c * 32, r);
        tmem_ld_wait();
        if (i < n) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(drow + c * 32 + u * 8) =
                make_uint4(pack_bf16x2(__uint_as_float(r[u * 8 + 0]) * mul, __uint_as_float(r[u * 8 + 1]) * mul),
                           pack_bf16x2(__uint_as_float(r[u * 8 + 2]) * mul, __uint_as_float(r[u * 8 + 3]) * mul),
                           pack_bf16x2(__uint_as_float(r[u * 8 + 4]) * mul, __uint_as_float(r[u * 8 + 5]) * mul),
                           pack_bf16x2(__uint_as_float(r[u * 8 + 6]) * mul, __uint_as_float(r[u * 8 + 7]) * mul));
        }
      }
    };
    mbar_wait(&bars[B_DV], 0, 0xe23);
    tc_fence_after();
    drain(256, 2, 1.f);   // dV
    arrive(B_DVDONE);
    mbar_wait(&bars[B_DQ], 0, 0xe24);
    tc_fence_after();
    drain(0, 0, scale);   // dQ
    mbar_wait(&bars[B_LD2], 0, 0xe25);
    if (F16) cvt_region(smem_u32(r0), op_bytes, wt, 128);
    arrive(B_CV2);
    mbar_wait(&bars[B_DK], 0, 0xe26);
    tc_fence_after();
    drain(256, 1, scale);  // dK
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

constexpr size_t kAttSmem = 1024 + 3 * (size_t)kRegion + 256;

uint32_t att_lbo() {
  static const uint32_t v = [] {
    const char* e = getenv("KOA_ATTN_LBO");  // debug knob (bytes); the tiles of one operand are kTile apart
    return e != nullptr && atoi(e) > 0 ? (uint32_t)atoi(e) : kTile;
  }();
  return v;
}

}  // namespace

bool koa_attention_tc_ok(int n, int head_dim) { return n >= 1 && n <= 128 && head_dim % 64 == 0 && head_dim >= 64 && head_dim <= 256; }

int koa_k_attention_tc_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                           cudaStream_t st, int f16) {
  KOA_REQUIRE(koa_attention_tc_ok(n, head_dim), "tcgen05 attention: n <= 128, head_dim in {64, 128, 192, 256}");
  const int dmodel = heads * head_dim;
  CUtensorMap tm;
  int rc = koa_tmap_2d_bf16(&tm, qkv, (uint64_t)3 * dmodel, (uint64_t)batch * n, (uint64_t)3 * dmodel * 2, 64, 128);
  if (rc) return rc;
  const AttDesc ad{att_lbo()};
  static std::atomic<unsigned long long> d0{0}, d1{0};
  if (f16) {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_tc_fwd_kernel<true>, (int)kAttSmem, d1));
    attention_tc_fwd_kernel<true><<<batch * heads, kAttThreads, kAttSmem, st>>>(tm, (bf16*)out, probs, n, heads, head_dim, scale, ad);
  } else {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_tc_fwd_kernel<false>, (int)kAttSmem, d0));
    attention_tc_fwd_kernel<false><<<batch * heads, kAttThreads, kAttSmem, st>>>(tm, (bf16*)out, probs, n, heads, head_dim, scale, ad);
  }
  KOA_LAUNCH_CHECK();
  return 0;
}

int koa_k_attention_tc_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                           int head_dim, float scale, cudaStream_t st, int f16) {
  KOA_REQUIRE(koa_attention_tc_ok(n, head_dim), "tcgen05 attention: n <= 128, head_dim in {64, 128, 192, 256}");
  const int dmodel = heads * head_dim;
  CUtensorMap tq, td;
  int rc = koa_tmap_2d_bf16(&tq, qkv, (uint64_t)3 * dmodel, (uint64_t)batch * n, (uint64_t)3 * dmodel * 2, 64, 128);
  if (rc) return rc;
  rc = koa_tmap_2d_bf16(&td, dout, (uint64_t)dmodel, (uint64_t)batch * n, (uint64_t)dmodel * 2, 64, 128);
  if (rc) return rc;
  const AttDesc ad{att_lbo()};
  static std::atomic<unsigned long long> d0{0}, d1{0};
  if (f16) {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_tc_bwd_kernel<true>, (int)kAttSmem, d1));
    attention_tc_bwd_kernel<true><<<batch * heads, kAttThreads, kAttSmem, st>>>(tq, td, probs, (bf16*)dqkv, n, heads, head_dim, scale, ad);
  } else {
    KOA_CHECK_CUDA(koa_ensure_dyn_smem(attention_tc_bwd_kernel<false>, (int)kAttSmem, d0));
    attention_tc_bwd_kernel<false><<<batch * heads, kAttThreads, kAttSmem, st>>>(tq, td, probs, (bf16*)dqkv, n, heads, head_dim, scale, ad);
  }
  KOA_LAUNCH_CHECK();
  return 0;
}
