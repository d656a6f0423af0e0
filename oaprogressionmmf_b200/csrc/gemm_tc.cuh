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

constexpr int kGemmThreads = 192;     // wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kEpiWarps = 8;          // fprop/dgrad kernel: 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kKmajorThreads = 64 + 32 * kEpiWarps;
constexpr int BM = 128;
constexpr int BK = 64;

constexpr int kMaxStatCols = 2048;  // widest BatchNorm in the networks (layer4 output)
template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return 1024 /*align slack*/ + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 4) * 8 + 16 +
         2 * kMaxStatCols * sizeof(float) + kEpiWarps * 2048 /*epilogue staging*/ + 16;
}

// ---- coalesced epilogue stores -------------------------------------------------------------------
// After tcgen05.ld a thread owns 32 consecutive columns of ONE row, so a direct store would write 16-byte
// pieces of 32 different rows per instruction (half-filled sectors). Each epilogue warp therefore stages its
// 32-row sub-tile in a private 2 KB shared-memory buffer ([32 rows][64 bytes], 16-byte pieces XOR-swizzled by
// (row >> 1) & 3 so both the row-wise writes and the transposed reads are bank-conflict free) and writes it
// back 8 rows x 64 contiguous bytes per instruction.
constexpr int kStageBytesPerWarp = 32 * 64;

__device__ __forceinline__ uint32_t stage_off(int row, int piece) {
  return (uint32_t)(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
}

// Row-major store of a [32 rows][64 bytes] warp tile: `gbase` points at (first row of the warp, first byte of the
// 64-byte segment); rows are `pitch_bytes` apart; rows >= rows_valid are skipped.
__device__ __forceinline__ void stage_flush(const uint8_t* stage, uint8_t* gbase, long long pitch_bytes, int rows_valid,
                                            int lane) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2);
    const int piece = lane & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + stage_off(row, piece));
    if (row < rows_valid) *reinterpret_cast<uint4*>(gbase + row * pitch_bytes + piece * 16) = v;
  }
  __syncwarp();
}

// 32 fp32 values of this thread's row -> bf16 -> staged -> coalesced store at column n of a bf16 [M][ld] tensor.
__device__ __forceinline__ void store_row_bf16(const float (&v)[32], uint8_t* stage, bf16* g, long long ld, long long row0,
                                               int n, int rows_valid, int lane) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint4 q;
    q.x = pack_bf16x2(v[p * 8 + 0], v[p * 8 + 1]); q.y = pack_bf16x2(v[p * 8 + 2], v[p * 8 + 3]);
    q.z = pack_bf16x2(v[p * 8 + 4], v[p * 8 + 5]); q.w = pack_bf16x2(v[p * 8 + 6], v[p * 8 + 7]);
    *reinterpret_cast<uint4*>(stage + stage_off(lane, p)) = q;
  }
  stage_flush(stage, reinterpret_cast<uint8_t*>(g + row0 * ld + n), ld * 2, rows_valid, lane);
}

__device__ __forceinline__ void store_row_f32(const float (&v)[32], uint8_t* stage, float* g, long long ld, long long row0,
                                              int n, int rows_valid, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // two 16-column halves of 64 bytes each
#pragma unroll
    for (int p = 0; p < 4; ++p)
      *reinterpret_cast<float4*>(stage + stage_off(lane, p)) =
          make_float4(v[h * 16 + p * 4 + 0], v[h * 16 + p * 4 + 1], v[h * 16 + p * 4 + 2], v[h * 16 + p * 4 + 3]);
    stage_flush(stage, reinterpret_cast<uint8_t*>(g + row0 * ld + n + h * 16), ld * 4, rows_valid, lane);
  }
}

// Epilogue of one 32-row x 32-column chunk. `row0` = first row of this warp's 32-row group, the thread owns row
// row0 + lane. `stage` = this warp's 2 KB staging buffer.
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], const EpiParams& ep, long long row0, int rows_valid,
                                               int n, int lane, uint8_t* stage, float* s_sum, float* s_sumsq,
                                               int c_local) {
  const bool row_ok = lane < rows_valid;
  const long long row_off = (row0 + lane) * ep.ldo;
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
  if (ep.pre_out != nullptr) store_row_bf16(v, stage, ep.pre_out, ep.ldo, row0, n, rows_valid, lane);
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
  if (ep.out_fp32) {
    store_row_f32(v, stage, reinterpret_cast<float*>(ep.out), ep.ldo, row0, n, rows_valid, lane);
    if (ep.out_bf16_copy != nullptr) store_row_bf16(v, stage, ep.out_bf16_copy, ep.ldo, row0, n, rows_valid, lane);
  } else {
    // bf16 output: stage, (statistics from the staged, i.e. rounded, values), coalesced write-back
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint4 q;
      q.x = pack_bf16x2(v[p * 8 + 0], v[p * 8 + 1]); q.y = pack_bf16x2(v[p * 8 + 2], v[p * 8 + 3]);
      q.z = pack_bf16x2(v[p * 8 + 4], v[p * 8 + 5]); q.w = pack_bf16x2(v[p * 8 + 6], v[p * 8 + 7]);
      if (!row_ok) q = make_uint4(0, 0, 0, 0);  // rows past M contribute exact zeros to the statistics
      *reinterpret_cast<uint4*>(stage + stage_off(lane, p)) = q;
    }
    if (ep.col_sum != nullptr) {
      __syncwarp();
      // lane = (h, p): column pair (2p, 2p+1) over the 16 rows 2i + h; the two lane halves read rows in different
      // bank halves, each half reads one contiguous 64-byte row per step (conflict free)
      const int h = lane >> 4, p = lane & 15;
      float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + stage_off(2 * i + h, p >> 2) + (p & 3) * 4);
        const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
        s0 += x0; s1 += x1;
        q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
      if (h == 0) {
        atomicAdd(&s_sum[c_local + 2 * p], s0); atomicAdd(&s_sum[c_local + 2 * p + 1], s1);
        atomicAdd(&s_sumsq[c_local + 2 * p], q0); atomicAdd(&s_sumsq[c_local + 2 * p + 1], q1);
      }
    }
    stage_flush(stage, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(ep.out) + row0 * ep.ldo + n), (long long)ep.ldo * 2,
                rows_valid, lane);
  }
}

// out[M, N] = A[M, K] * B[N, K]^T with the fused epilogue. A_IM2COL: A rows are the output
// pixels of a convolution over an NHWC tensor and K enumerates (r, s, cin).
//
// Persistent: gridDim.x = min(#tiles, #SMs) CTAs walk the tile list (n-tile fastest, so CTAs running at the
// same time share their A rows through L2). Three pipelines overlap across tiles:
//   TMA -> smem ring (STAGES deep, full/empty mbarriers, runs ahead into the next tile),
//   MMA -> TMEM accumulator ring (2 x BN columns, tmem_full/tmem_empty mbarriers),
//   epilogue warps drain accumulator i while the MMA of tile i+1 is issued.
template <int BN, int STAGES, bool A_IM2COL>
__global__ void __launch_bounds__(kKmajorThreads, 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                   int K, ConvGeom g, EpiParams ep) {
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr int ACC = 2;  // TMEM accumulator stages
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC; // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);
  // BatchNorm statistics are accumulated per output column in shared memory over ALL tiles of this CTA and
  // flushed once at the end: 148 same-address global atomics per column instead of one per tile.
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_sumsq = s_sum + kMaxStatCols;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(((uintptr_t)(s_sumsq + kMaxStatCols) + 15) & ~(uintptr_t)15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + BM - 1) / BM;
  const int num_tiles = n_tiles * m_tiles;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, ACC * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // k-block counter across all tiles of this CTA
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_t = tile % n_tiles;
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = n_t * BN;
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
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
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
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      uint32_t it = 0, lt = 0;  // k-block counter, local tile counter
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt % ACC;
        const uint32_t acc_phase = (lt / ACC) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, 0x700 + acc);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], phase, 0x200 + s);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 (32 bytes) along K inside the 128-byte swizzle row: +2 in (addr >> 4) units
            umma_bf16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // Epilogue warps 2..9: the TMEM lane quarter is fixed by warp id modulo 4; the two warps of a quarter take
    // alternate 32-column chunks of the tile.
    const int q = warp & 3;
    const int e = warp - 2;
    const int chunk_par = e >> 2;
    const int ep_tid = e * 32 + lane;
    constexpr int kEpiThreads = kEpiWarps * 32;
    uint8_t* stage = s_stage + e * kStageBytesPerWarp;
    const bool stats = ep.col_sum != nullptr;
    if (stats) {
      for (int i = ep_tid; i < 2 * kMaxStatCols; i += kEpiThreads) s_sum[i] = 0.0f;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int n0 = (tile % n_tiles) * BN;
      const int m0 = (tile / n_tiles) * BM;
      const uint32_t acc = lt % ACC;
      const uint32_t acc_phase = (lt / ACC) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase, 0x300 + acc);
      tc_fence_after();
      const long long row0 = (long long)m0 + q * 32;
      const int rows_valid = max(0, min(32, M - (int)row0));
      bool released = false;
#pragma unroll 1
      for (int c0 = chunk_par * 32; c0 < BN; c0 += 64) {
        if (n0 + c0 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 64 >= BN || n0 + c0 + 64 >= N) {
          // last TMEM read of this warp for this tile: hand the accumulator back before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          released = true;
        }
        epilogue_chunk(r, ep, row0, rows_valid, n0 + c0, lane, stage, s_sum, s_sumsq, n0 + c0);
      }
      if (!released) {  // this warp had no chunk inside N
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    if (stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int i = ep_tid; i < N; i += kEpiThreads) {
        atomicAdd(ep.col_sum + i, s_sum[i]);
        atomicAdd(ep.col_sumsq + i, s_sumsq[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ACC * BN);
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
