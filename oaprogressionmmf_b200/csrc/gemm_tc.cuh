// gemm_tc.cuh — tcgen05/TMEM GEMM core shared by the 1x1/kxk convolutions and the
// transformer linears of the koafusion hot path.
//
// One CTA computes a 128 x BN fp32 accumulator tile held in TMEM:
//   warp 0     : TMA producer (tiled 2-D loads, or im2col-mode loads of an NHWC tensor)
//   warp 1     : TMEM allocation + single-thread tcgen05.mma issue
//   warps 2..5 : epilogue (tcgen05.ld -> registers -> fused epilogue -> global)
// Operands are bf16 staged in 128-byte-swizzled shared memory, STAGES deep.
#pragma once
#include "koa_common.cuh"

namespace koa {

struct ConvGeom {
  int hout, wout;   // output spatial size; GEMM rows enumerate (n, oh, ow)
  int stride, pad;  // convolution stride / padding
  int filt_s;       // filter width (taps are enumerated r-major: tap = r * filt_s + s)
  int cin_blocks;   // Cin / 64
  int grouped;      // 1: block-diagonal (grouped) 3x3 conv, output chunk n_t only reads input chunk n_t
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_GELU_GRAD = 3 };

struct EpiParams {
  void* out;              // [M, ldo] bf16 (out_fp32 == 0) or fp32
  int ldo;
  int out_fp32;
  int act;
  const float* bias;      // [N]
  bf16* pre_out;          // optional bf16 copy of (acc + bias) before the activation
  const bf16* aux;        // ACT_GELU_GRAD: pre-activation h; result = v * gelu'(h)
  const float* res_f32;   // optional fp32 addend [M, ldo]
  const bf16* add_bf16;   // optional bf16 addend [M, ldo]
  const bf16* mask_bf16;  // optional: add_bf16 contributes only where mask > 0
  bf16* out_bf16_copy;    // optional bf16 copy of the final value (when out is fp32)
  float* col_sum;         // optional per-column sum / sum of squares of the bf16-rounded output
  float* col_sumsq;
};

constexpr int kGemmThreads = 192;
constexpr int BM = 128;
constexpr int BK = 64;

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return 1024 /*align slack*/ + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 +
         2 * BN * sizeof(float);
}

__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], const EpiParams& ep, long long row_off, int n,
                                               bool row_ok, int lane, float* s_sum, float* s_sumsq, int c_local) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

  if (ep.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(ep.bias + n + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (ep.pre_out != nullptr && row_ok) {
    uint4* dst = reinterpret_cast<uint4*>(ep.pre_out + row_off + n);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 q;
      q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
      q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
      dst[j / 8] = q;
    }
  }
  if (ep.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (ep.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (ep.act == ACT_GELU_GRAD) {
    if (row_ok) {
      const uint4* src = reinterpret_cast<const uint4*>(ep.aux + row_off + n);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 q = src[j / 8];
        const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
        v[j] *= gelu_erf_grad(a.x); v[j + 1] *= gelu_erf_grad(a.y);
        v[j + 2] *= gelu_erf_grad(b.x); v[j + 3] *= gelu_erf_grad(b.y);
        v[j + 4] *= gelu_erf_grad(c.x); v[j + 5] *= gelu_erf_grad(c.y);
        v[j + 6] *= gelu_erf_grad(d.x); v[j + 7] *= gelu_erf_grad(d.y);
      }
    }
  }
  if (ep.res_f32 != nullptr && row_ok) {
    const float4* src = reinterpret_cast<const float4*>(ep.res_f32 + row_off + n);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = src[j / 4];
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (ep.add_bf16 != nullptr && row_ok) {
    const uint4* src = reinterpret_cast<const uint4*>(ep.add_bf16 + row_off + n);
    const uint4* msk = ep.mask_bf16 ? reinterpret_cast<const uint4*>(ep.mask_bf16 + row_off + n) : nullptr;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const uint4 q = src[j / 8];
      float a[8];
      float2 t;
      t = unpack_bf16x2(q.x); a[0] = t.x; a[1] = t.y;
      t = unpack_bf16x2(q.y); a[2] = t.x; a[3] = t.y;
      t = unpack_bf16x2(q.z); a[4] = t.x; a[5] = t.y;
      t = unpack_bf16x2(q.w); a[6] = t.x; a[7] = t.y;
      if (msk != nullptr) {
        const uint4 mq = msk[j / 8];
        float m[8];
        t = unpack_bf16x2(mq.x); m[0] = t.x; m[1] = t.y;
        t = unpack_bf16x2(mq.y); m[2] = t.x; m[3] = t.y;
        t = unpack_bf16x2(mq.z); m[4] = t.x; m[5] = t.y;
        t = unpack_bf16x2(mq.w); m[6] = t.x; m[7] = t.y;
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = m[u] > 0.0f ? a[u] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v[j + u] += a[u];
    }
  }
  if (row_ok) {
    if (ep.out_fp32) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + row_off + n);
#pragma unroll
      for (int j = 0; j < 32; j += 4) dst[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      if (ep.out_bf16_copy != nullptr) {
        uint4* d2 = reinterpret_cast<uint4*>(ep.out_bf16_copy + row_off + n);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 q;
          q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
          q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
          d2[j / 8] = q;
        }
      }
    } else {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + row_off + n);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
        q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
        dst[j / 8] = q;
      }
    }
  }
  if (ep.col_sum != nullptr) {
    // Statistics of the values as stored (bf16-rounded); rows past M hold exact zeros.
    float s[32], q[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = row_ok ? bf16_round(v[j]) : 0.0f;
      s[j] = x;
      q[j] = x * x;
    }
    warp_transpose_reduce32(s, lane);
    warp_transpose_reduce32(q, lane);
    atomicAdd(&s_sum[c_local + lane], s[0]);
    atomicAdd(&s_sumsq[c_local + lane], q[0]);
  }
}

// out[M, N] = A[M, K] * B[N, K]^T with the fused epilogue. A_IM2COL: A rows are the output
// pixels of a convolution over an NHWC tensor and K enumerates (r, s, cin).
template <int BN, int STAGES, bool A_IM2COL>
__global__ void __launch_bounds__(kGemmThreads)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                   int K, ConvGeom g, EpiParams ep) {
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_sumsq = s_sum + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN;
  const int n_t = blockIdx.x % n_tiles;
  const int m_t = blockIdx.x / n_tiles;
  const int m0 = m_t * BM;
  const int n0 = n_t * BN;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int pw = 0, ph = 0, pn = 0;
      if (A_IM2COL) {
        const int hw = g.hout * g.wout;
        pn = m0 / hw;
        const int rem = m0 - pn * hw;
        const int oh = rem / g.wout;
        const int ow = rem - oh * g.wout;
        pw = ow * g.stride - g.pad;
        ph = oh * g.stride - g.pad;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1, 0x100 + s);
        mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        if (A_IM2COL) {
          const int tap = kb / g.cin_blocks;
          const int cb = kb - tap * g.cin_blocks;
          const int fr = tap / g.filt_s;
          const int fs = tap - fr * g.filt_s;
          const int c0 = g.grouped ? n_t * 64 : cb * 64;
          tma_load_im2col_4d(sA + s * A_BYTES, &tmA, &full_bar[s], c0, pw, ph, pn, (uint16_t)fs, (uint16_t)fr);
        } else {
          tma_load_2d(sA + s * A_BYTES, &tmA, &full_bar[s], kb * BK, m0);
        }
        tma_load_2d(sB + s * B_BYTES, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], phase, 0x200 + s);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 (32 bytes) along K inside the 128-byte swizzle row: +2 in (addr >> 4) units
          umma_bf16_ss(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // Epilogue warps 2..5; TMEM lane quarter is fixed by warp id modulo 4.
    const int q = warp & 3;
    const int ep_tid = (warp - 2) * 32 + lane;
    if (ep.col_sum != nullptr) {
      for (int i = ep_tid; i < 2 * BN; i += 128) s_sum[i] = 0.0f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0, 0x300);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < M;
    const long long row_off = (long long)row * ep.ldo;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      epilogue_chunk(r, ep, row_off, n0 + c0, row_ok, lane, s_sum, s_sumsq, c0);
    }
    if (ep.col_sum != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = ep_tid; i < BN; i += 128) {
        if (n0 + i < N) {
          atomicAdd(ep.col_sum + n0 + i, s_sum[i]);
          atomicAdd(ep.col_sumsq + n0 + i, s_sumsq[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// Weight-gradient GEMM: dW[Cout, tap, Cin] += sum over pixels p of dY[p, Cout] * X[p (shifted by tap), Cin].
// Both operands have the reduction dimension (pixels) outermost in memory, i.e. they are MN-major
// UMMA operands. Split-K over gridDim.y with fp32 vector atomics into dW.
struct WgradDesc {
  uint32_t lbo, sbo;     // MN-major smem descriptor strides (bytes)
  uint32_t k_adv;        // descriptor start-address advance per 16-pixel MMA step (bytes)
};

template <int BN, int STAGES>
constexpr size_t wgrad_smem_bytes() {
  return 1024 + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16;
}

template <int BN, int STAGES, bool B_IM2COL>
__global__ void __launch_bounds__(kGemmThreads)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int cout, int cin,
                  int pixels, int taps, ConvGeom g, float* __restrict__ dw, int kb_per_split, WgradDesc wd) {
  // g.grouped: dw is [C][taps][64] (per-64-channel-chunk dense blocks); tile n_t pairs input chunk n_t with
  // output chunk n_t only (BN must be 64; the upper 64 accumulator rows are discarded).
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t BOX_BYTES = 64 * BK * 2;  // one {64 channels x 64 pixels} TMA box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (cin + BN - 1) / BN;
  const int m_tiles = g.grouped ? 1 : (cout + BM - 1) / BM;
  int t = blockIdx.x;
  const int n_t = t % n_tiles; t /= n_tiles;
  const int m_t = t % m_tiles; t /= m_tiles;
  const int tap = t;
  const int m0 = g.grouped ? n_t * 64 : m_t * BM;
  const int n0 = n_t * BN;
  const int num_kb_total = (pixels + BK - 1) / BK;
  const int kb_begin = blockIdx.y * kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + kb_per_split);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int fr = tap / g.filt_s;
      const int fs = tap - fr * g.filt_s;
      const int hw = g.hout * g.wout;
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb_begin + i;
        const int s = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1, 0x400 + s);
        mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        const int p0 = kb * BK;
#pragma unroll
        for (int j = 0; j < BM / 64; ++j)
          tma_load_2d(sA + s * A_BYTES + j * BOX_BYTES, &tmA, &full_bar[s], m0 + j * 64, p0);
        if (B_IM2COL) {
          const int pn = p0 / hw;
          const int rem = p0 - pn * hw;
          const int oh = rem / g.wout;
          const int ow = rem - oh * g.wout;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_im2col_4d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64,
                               ow * g.stride - g.pad, oh * g.stride - g.pad, pn, (uint16_t)fs, (uint16_t)fr);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64, p0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 1, 1);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], phase, 0x500 + s);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), wd.lbo, wd.sbo);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), wd.lbo, wd.sbo);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * wd.k_adv) >> 4);
          umma_bf16_ss(tmem_base, adesc + adv, bdesc + adv, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0, 0x600);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;  // output channel
    const bool row_ok = g.grouped ? (q * 32 + lane) < 64 : row < cout;
    const int ldw = g.grouped ? 64 : cin;
    const int col0 = g.grouped ? 0 : n0;
    float* dst_row = dw + ((long long)row * taps + tap) * ldw;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= cin) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (row_ok) {
        float* dst = dst_row + col0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                       "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                       "f"(__uint_as_float(r[j + 3]))
                       : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

}  // namespace koa
