// gemm_conv.cuh — the convolution flavour of the tcgen05 GEMM (16-bit output, optional addend / ReLU-backward gate /
// BatchNorm statistics), built for the layers whose epilogue is the bottleneck: the 1x1 convolutions of layer1/layer2
// move 4-16 output bytes per MMA byte and are HBM-bound, so the kernel is organised around draining TMEM fast.
//
// Same TMA -> smem ring -> tcgen05.mma -> double-buffered TMEM pipeline as gemm_kmajor_kernel (gemm_tc.cuh), but
//   * 16 epilogue warps (4 per TMEM lane quarter): with BN = 128 every warp owns one fixed 32-column chunk of every
//     tile (two chunks with BN = 256), with BN = 64 the warps split into two sets that take alternate tiles (both
//     accumulator stages drain at the same time). Four warps per scheduler hide the tcgen05.ld / LDS / LDG latencies
//     that two could not.
//   * a CTA keeps ONE n-tile for its whole life (tile = (m, n_t) with n_t = blockIdx % n_tiles): the columns of a
//     warp never change, so the BatchNorm statistics live in four registers per lane for the whole kernel: no
//     shared-memory accumulation, no per-tile barrier; one exchange through shared memory + one global atomic per
//     column and CTA at the end.
//   * formats are template parameters (no per-element format selects), the statistics use packed fp32x2 math
//     (FADD2 / FFMA2), the ReLU gate is one HSET2 + AND per element pair on the packed output,
//     sum(dz * xhat) is accumulated as sum(dz * (y - mean)) (the mean of the warp's fixed column pair sits in two
//     registers; subtracting it per element instead of correcting the total once avoids the cancellation
//     sum(dz*y) - mean*sum(dz), which turned the summation-order noise of the atomics into 1e-3 of dgamma for channels
//     with |mean| >> std) and scaled by invstd once per CTA.
//   * MODE 1 (backward) moves its epilogue operands and its output with the TMA unit ([32 x 32] boxes in the 64-byte
//     swizzle, which IS the staging layout): no operand registers, no address arithmetic, no row predicates.
//   * MODE 2 (forward with the BatchNorm already known: eval mode, and the y-free last convolution of a bottleneck in
//     train mode, fe_engine.cu): value = acc * scale[col] + shift[col] (+ residual [* rscale[col] + rshift[col]]), ReLU,
//     stored as fp16 and optionally once more as bf16 (the weight-gradient operand of the backward pass): the raw
//     convolution output never reaches HBM and no BatchNorm-apply pass follows. The residual tile arrives by TMA like the
//     MODE 1 operands; the per-column coefficients sit in shared memory and are read as warp-wide broadcasts.
//   * K concatenation (MODE 1, 1x1 only): k-blocks >= kb_split come from a second A tensor (tmA2): the data gradient of the
//     y-free BatchNorm is ONE GEMM over [G | a2] (fe_engine.cu); `col_bias` is added per column before the gate.
//   * CTA2: the compute-bound layers run as CTA pairs (cluster of 2 along M, tcgen05.mma.cta_group::2, M = 256): each
//     CTA stages its own 128 A rows and HALF of the B tile, the pair's tensor cores read both halves, so the
//     shared-memory fill per FLOP drops from (128 + BN) to (128 + BN / 2) rows per k-block. Only the rank-0 CTA
//     issues MMAs; its commits arrive on the barriers of both CTAs; both CTAs' TMA loads complete on rank 0's barrier.
#pragma once
#include "gemm_tc.cuh"

namespace koa {

constexpr int kConvEpiWarps = 16;
constexpr int kConvThreads = 64 + 32 * kConvEpiWarps;  // TMA warp, MMA warp, 16 epilogue warps

// MODE 0: forward (statistics optional). MODE 1: backward (addend / gate / BatchNorm-backward statistics / column bias).
// MODE 2: forward with BatchNorm apply (+ residual) + ReLU in the epilogue, fp16 output + optional bf16 copy.
template <int BN, int MODE>
constexpr int conv_coef_floats() {
  // MODE 0: end-of-kernel statistics exchange; MODE 2: scale, shift, residual scale, residual shift
  return MODE == 0 ? kConvEpiWarps * 128 : (MODE == 1 ? 0 : 4 * BN);
}
template <int MODE>
constexpr int conv_stage_bufs() { return MODE == 0 ? 1 : (MODE == 1 ? 3 : 2); }
template <int BN, int STAGES, int MODE, bool CTA2>
constexpr size_t conv_smem_bytes() {
  // MODE 0: one staging buffer per epilogue warp + the end-of-kernel statistics exchange; MODE 1: three staging
  // buffers per warp (addend -> output, gate, y), the statistics exchange re-uses them; MODE 2: two (residual -> fp16
  // output, bf16 copy)
  return 1024 /*align slack*/ + (size_t)STAGES * (BM * BK * 2 + (CTA2 ? BN / 2 : BN) * BK * 2) + (2 * STAGES + 4) * 8 + 16 +
         kConvEpiWarps * 8 /*operand mbarriers*/ + conv_coef_floats<BN, MODE>() * sizeof(float) + 1024 +
         (size_t)kConvEpiWarps * conv_stage_bufs<MODE>() * kStageBytesPerWarp;
}

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return r;
}
template <bool F16>
__device__ __forceinline__ float2 unpack16(uint32_t v) { return F16 ? unpack_f16x2(v) : unpack_bf16x2(v); }
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
// 0xffff per 16-bit half that is > 0
template <bool F16>
__device__ __forceinline__ uint32_t gt0_mask(uint32_t v) {
  if (F16) return __hgt2_mask(*reinterpret_cast<__half2*>(&v), __half2(__ushort_as_half(0), __ushort_as_half(0)));
  return __hgt2_mask(*reinterpret_cast<bf162*>(&v), bf162(__ushort_as_bfloat16(0), __ushort_as_bfloat16(0)));
}

template <int BN, int STAGES, bool A_IM2COL, int MODE, bool OF16, bool AF16, bool CTA2>
__global__ void __launch_bounds__(kConvThreads, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAdd,
                 const __grid_constant__ CUtensorMap tmGate, const __grid_constant__ CUtensorMap tmY,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmOut2, int M, int N, int K,
                 int kb_split, ConvGeom g, EpiParams ep) {
  static_assert(BN == 64 || BN == 128 || BN == 256, "tile widths");
  static_assert(BN != 256 || CTA2, "BN = 256 needs the CTA pair (shared-memory budget)");
  constexpr uint32_t B_ROWS = CTA2 ? BN / 2 : BN;    // B rows staged by this CTA
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = B_ROWS * BK * 2;
  constexpr int ACC = 2;
  constexpr int NSTG = conv_stage_bufs<MODE>();      // staging buffers per epilogue warp
  constexpr int CH = BN == 256 ? 2 : 1;              // 32-column chunks per epilogue warp and tile
  constexpr int kDrainWarps = BN == 64 ? 8 : 16;     // warps (of one CTA) that read one accumulator stage
  constexpr int kEpiThreads = kConvEpiWarps * 32;
  constexpr int BMT = CTA2 ? 2 * BM : BM;            // rows of the tile of a CTA (pair)
#ifdef KOA_CONV_FWD_STG
  constexpr bool TMA_OUT = false;                    // MODE 0 output through LDS + STG.128 (first version)
#else
  constexpr bool TMA_OUT = true;                     // MODE 0 output through TMA stores as well
#endif
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;    // [ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;  // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);
  uint64_t* op_bar = reinterpret_cast<uint64_t*>(tmem_slot + 4);  // [16] operand tiles of an epilogue warp have landed
  float* s_fin0 = reinterpret_cast<float*>(op_bar + kConvEpiWarps);
  // staging buffers: 1024-byte aligned (the 64-byte TMA swizzle pattern is a function of the address bits 4..8)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(
      ((uintptr_t)(s_fin0 + conv_coef_floats<BN, MODE>()) + 1023) & ~(uintptr_t)1023);
  float* s_coef = s_fin0;  // MODE 2: [4][BN] scale, shift, residual scale, residual shift
  // end-of-kernel statistics exchange [16 warps][CH][16 lanes][4]; MODE 1 re-uses the first staging buffer of each warp
  float* s_fin = MODE == 0 ? s_fin0 : reinterpret_cast<float*>(s_stage);
  constexpr int kFinStride = MODE == 0 ? 128 : NSTG * kStageBytesPerWarp / 4;  // floats between two warps' slots

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + BMT - 1) / BMT;
  const int num_kb = (K + BK - 1) / BK;
  // this CTA (pair): n-tile n_t, m-tiles m_start, m_start + m_step, ... (CTAs that run together share their A rows in L2)
  const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_t = unit % n_tiles;
  const int m_start = unit / n_tiles;
  const int m_step = n_units / n_tiles;
  const int my_tiles = m_start < m_tiles ? (m_tiles - m_start + m_step - 1) / m_step : 0;
  const int n0 = n_t * BN;
  const int m_off = (int)rank * BM;  // rows of this CTA inside the pair's tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE == 1) {
      tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmAdd); tma_prefetch_desc(&tmGate); tma_prefetch_desc(&tmY);
      tma_prefetch_desc(&tmA2);
    } else if (MODE == 2) {
      tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmAdd); tma_prefetch_desc(&tmOut2);
    } else if (TMA_OUT) {
      tma_prefetch_desc(&tmOut);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kDrainWarps * (CTA2 ? 2 : 1));  // pair: rank 0 collects both CTAs' epilogue warps
    }
#pragma unroll
    for (int w = 0; w < kConvEpiWarps; ++w) mbar_init(&op_bar[w], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc2(tmem_slot, ACC * BN);
    else tmem_alloc(tmem_slot, ACC * BN);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int m0 = (m_start + lt * m_step) * BMT + m_off;
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
          mbar_wait(&empty_bar[s], phase ^ 1, 0x900 + s);
          // pair: rank 0 announces the bytes of both CTAs; the peer's loads complete on rank 0's barrier
          if (!CTA2) mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          else if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_BYTES + B_BYTES));
          int ca = kb * BK;
          uint16_t fs = 0, fr = 0;
          if (A_IM2COL) {
            const int tap = kb / g.cin_blocks;
            const int cb = kb - tap * g.cin_blocks;
            fr = (uint16_t)(tap / g.filt_s);
            fs = (uint16_t)(tap - fr * g.filt_s);
            ca = g.grouped ? n_t * 64 : cb * 64;
          }
          // K concatenation: the k-blocks from kb_split on read the second A tensor (same rows)
          const bool second = !A_IM2COL && MODE == 1 && kb >= kb_split;
          const CUtensorMap* ta = second ? &tmA2 : &tmA;
          if (second) ca = (kb - kb_split) * BK;
          if (!CTA2) {
            if (A_IM2COL) tma_load_im2col_4d(sA + s * A_BYTES, &tmA, &full_bar[s], ca, pw, ph, pn, fs, fr);
            else tma_load_2d(sA + s * A_BYTES, ta, &full_bar[s], ca, m0);
            tma_load_2d(sB + s * B_BYTES, &tmB, &full_bar[s], kb * BK, n0);
          } else {
            if (A_IM2COL) tma2_load_im2col_4d(sA + s * A_BYTES, &tmA, &full_bar[s], ca, pw, ph, pn, fs, fr);
            else tma2_load_2d(sA + s * A_BYTES, ta, &full_bar[s], ca, m0);
            tma2_load_2d(sB + s * B_BYTES, &tmB, &full_bar[s], kb * BK, n0 + (int)(rank * B_ROWS));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_16(BMT, BN, 0, 0, ep.a_f16, ep.b_f16);
      uint32_t it = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const uint32_t acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, 0xa00 + acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], phase, 0xb00 + s);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (CTA2) umma2_bf16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else umma_bf16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          if (CTA2) umma2_commit_both(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        if (CTA2) umma2_commit_both(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int cg = e >> 2;
    const int c0 = BN == 64 ? (cg & 1) * 32 : cg * 32;  // first chunk of this warp (BN = 256: second chunk at + 128)
    const int lt_first = BN == 64 ? (cg >> 1) : 0;
    constexpr int lt_step = BN == 64 ? 2 : 1;
    // MODE 0: `stage` = the output staging buffer. MODE 1: stage = addend, then output; stage_g = gate; stage_y = y.
    const uint32_t stage = smem_u32(s_stage) + e * (NSTG * kStageBytesPerWarp);
    const uint32_t stage_g = stage + (NSTG > 1 ? 1 : 0) * kStageBytesPerWarp;
    const uint32_t stage_y = stage + (NSTG > 2 ? 2 : 0) * kStageBytesPerWarp;
    const LaneMap lm = make_lane_map(lane);
    const bool stats = ep.col_sum != nullptr;
    const bool has_add = MODE >= 1 && ep.add_bf16 != nullptr, has_gate = MODE == 1 && ep.gate_bf16 != nullptr;
    const bool bwd = MODE == 1 && ep.stat_y != nullptr;
    const bool has_bias = MODE == 1 && ep.col_bias != nullptr;
    const bool has_out2 = MODE == 2 && ep.out_bf16_copy != nullptr;
    const bool res_affine = MODE == 2 && ep.res_scale != nullptr;
    const bool relu = MODE == 2 && ep.act == ACT_RELU;
    if (MODE == 2) {  // per-column coefficients of this CTA's n-tile -> shared memory (broadcast reads later)
      for (int i = e * 32 + lane; i < BN; i += kEpiThreads) {
        const int col = n0 + i;
        const bool ok = col < N;
        s_coef[i] = ok ? __ldg(ep.bn_scale + col) : 0.f;
        s_coef[BN + i] = ok ? __ldg(ep.bn_shift + col) : 0.f;
        s_coef[2 * BN + i] = (ok && res_affine) ? __ldg(ep.res_scale + col) : 1.f;
        s_coef[3 * BN + i] = (ok && res_affine) ? __ldg(ep.res_shift + col) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    const uint32_t coef_s = smem_u32(s_coef);
    const uint32_t op_bytes = (uint32_t)((has_add ? 1 : 0) + (has_gate ? 1 : 0) + (bwd ? 1 : 0)) * kStageBytesPerWarp;
    uint32_t op_phase = 0;
    // statistics lane mapping: lane = (h, p): column pair (2p, 2p + 1) of the chunk over the 16 rows 2i + h
    const int h = lane >> 4, p = lane & 15;
    const uint32_t st_base = (uint32_t)(h * 64 + (p & 3) * 4);
    float2 s_a[CH], s_b[CH], q_a[CH], q_b[CH];  // two chains per chunk: sums over even / odd steps
    float2 neg_mu[CH];                          // -mean of this lane's column pair (backward statistics)
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      s_a[j] = s_b[j] = q_a[j] = q_b[j] = make_float2(0.f, 0.f);
      neg_mu[j] = make_float2(0.f, 0.f);
      const int n = n0 + c0 + j * 128;
      if (bwd && n < N) {
        const float2 mu = __ldg(reinterpret_cast<const float2*>(ep.stat_mean + n + 2 * p));
        neg_mu[j] = make_float2(-mu.x, -mu.y);
      }
    }

    for (int lt = lt_first; lt < my_tiles; lt += lt_step) {
      const int m0 = (m_start + lt * m_step) * BMT + m_off;
      const uint32_t acc = lt & 1;
      const uint32_t acc_phase = (lt >> 1) & 1;
      const int row0 = m0 + q * 32;
      const int rows_valid = max(0, min(32, M - row0));
      bool waited = false;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int cj = c0 + j * 128;
        const int n = n0 + cj;
        const bool col_ok = n < N;
        const bool last = j == CH - 1;
        if (MODE == 0 && TMA_OUT && col_ok && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (MODE >= 1 && col_ok && lane == 0) {
          // The TMA unit fetches the operand tiles of this chunk ([32 rows][32 columns], 64-byte swizzle = the staging
          // layout; rows past M arrive as zeros) while the accumulator is still being computed. The previous
          // output store must have finished reading the buffer the addend lands in.
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (op_bytes != 0) {
            mbar_arrive_expect_tx(&op_bar[e], op_bytes);
            if (has_add) tma_load_2d(s_stage + (stage - smem_u32(s_stage)), &tmAdd, &op_bar[e], n, row0);
            if (has_gate) tma_load_2d(s_stage + (stage_g - smem_u32(s_stage)), &tmGate, &op_bar[e], n, row0);
            if (bwd) tma_load_2d(s_stage + (stage_y - smem_u32(s_stage)), &tmY, &op_bar[e], n, row0);
          }
        }
        // (lane 0 has waited for the previous TMA store to finish reading the staging buffer: no lane may write it earlier)
        if (MODE >= 1 || TMA_OUT) __syncwarp();
        if (!waited) {
          mbar_wait(&tmem_full_bar[acc], acc_phase, 0xc00 + acc);
          tc_fence_after();
          waited = true;
        }
        uint32_t r[32];
        if (col_ok) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)cj, r);
          tmem_ld_wait();
        }
        if (last) {  // accumulator handed back before the global traffic of the last chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTA2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        if (!col_ok) continue;

        if (MODE == 2) {  // BatchNorm apply: acc * scale + shift (coefficients: warp-wide shared-memory broadcasts)
          const uint32_t cb = coef_s + (uint32_t)cj * 4;
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            const uint4 sc = lds128(cb + t4 * 16), sh = lds128(cb + BN * 4 + t4 * 16);
            r[4 * t4 + 0] = __float_as_uint(fmaf(__uint_as_float(r[4 * t4 + 0]), __uint_as_float(sc.x), __uint_as_float(sh.x)));
            r[4 * t4 + 1] = __float_as_uint(fmaf(__uint_as_float(r[4 * t4 + 1]), __uint_as_float(sc.y), __uint_as_float(sh.y)));
            r[4 * t4 + 2] = __float_as_uint(fmaf(__uint_as_float(r[4 * t4 + 2]), __uint_as_float(sc.z), __uint_as_float(sh.z)));
            r[4 * t4 + 3] = __float_as_uint(fmaf(__uint_as_float(r[4 * t4 + 3]), __uint_as_float(sc.w), __uint_as_float(sh.w)));
          }
        }
        if (has_bias) {  // every lane reads the same 128 bytes: L1 broadcasts (N is a multiple of 32: no column guard)
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.col_bias + n) + t4);
            r[4 * t4 + 0] = __float_as_uint(__uint_as_float(r[4 * t4 + 0]) + b.x);
            r[4 * t4 + 1] = __float_as_uint(__uint_as_float(r[4 * t4 + 1]) + b.y);
            r[4 * t4 + 2] = __float_as_uint(__uint_as_float(r[4 * t4 + 2]) + b.z);
            r[4 * t4 + 3] = __float_as_uint(__uint_as_float(r[4 * t4 + 3]) + b.w);
          }
        }
        if (MODE >= 1 && op_bytes != 0) {
          mbar_wait(&op_bar[e], op_phase, 0xd00 + e);
          op_phase ^= 1;
        }
        if (has_add) {
          uint4 qa[4];
          row_lds(qa, stage, lm);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const uint32_t w[4] = {qa[pc].x, qa[pc].y, qa[pc].z, qa[pc].w};
            float a8[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              // MODE 1: bf16 gradient addend; MODE 2: residual in the forward activation format
              const float2 a = MODE == 2 ? unpack16<AF16>(w[u]) : unpack_bf16x2(w[u]);
              a8[2 * u] = a.x; a8[2 * u + 1] = a.y;
            }
            if (res_affine) {  // the residual is a raw convolution output with its own BatchNorm (downsample branch)
              const uint32_t cb = coef_s + (uint32_t)(2 * BN + cj + pc * 8) * 4;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint4 sc = lds128(cb + hh * 16), sh = lds128(cb + BN * 4 + hh * 16);
                a8[4 * hh + 0] = fmaf(a8[4 * hh + 0], __uint_as_float(sc.x), __uint_as_float(sh.x));
                a8[4 * hh + 1] = fmaf(a8[4 * hh + 1], __uint_as_float(sc.y), __uint_as_float(sh.y));
                a8[4 * hh + 2] = fmaf(a8[4 * hh + 2], __uint_as_float(sc.z), __uint_as_float(sh.z));
                a8[4 * hh + 3] = fmaf(a8[4 * hh + 3], __uint_as_float(sc.w), __uint_as_float(sh.w));
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) r[pc * 8 + u] = __float_as_uint(__uint_as_float(r[pc * 8 + u]) + a8[u]);
          }
        }
        if (relu) {
#pragma unroll
          for (int u = 0; u < 32; ++u) r[u] = __float_as_uint(fmaxf(__uint_as_float(r[u]), 0.f));
        }
        uint4 qv[4];
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          qv[pc].x = pack16<OF16>(__uint_as_float(r[pc * 8 + 0]), __uint_as_float(r[pc * 8 + 1]));
          qv[pc].y = pack16<OF16>(__uint_as_float(r[pc * 8 + 2]), __uint_as_float(r[pc * 8 + 3]));
          qv[pc].z = pack16<OF16>(__uint_as_float(r[pc * 8 + 4]), __uint_as_float(r[pc * 8 + 5]));
          qv[pc].w = pack16<OF16>(__uint_as_float(r[pc * 8 + 6]), __uint_as_float(r[pc * 8 + 7]));
        }
        if (has_gate) {  // ReLU backward: zero where the forward activation is not positive
          uint4 qg[4];
          row_lds(qg, stage_g, lm);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            qv[pc].x &= gt0_mask<AF16>(qg[pc].x); qv[pc].y &= gt0_mask<AF16>(qg[pc].y);
            qv[pc].z &= gt0_mask<AF16>(qg[pc].z); qv[pc].w &= gt0_mask<AF16>(qg[pc].w);
          }
        }
        if (rows_valid < 32 && lane >= rows_valid) {  // rows past M contribute exact zeros to the statistics
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) qv[pc] = make_uint4(0, 0, 0, 0);
        }
        if (has_add) __syncwarp();  // every lane has read its addend row: the buffer becomes the output buffer
        row_sts(stage, qv, lm);
        if (has_out2) {  // bf16 copy of the same values (weight-gradient operand of the backward pass)
          uint4 qb[4];
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            qb[pc].x = pack_bf16x2(__uint_as_float(r[pc * 8 + 0]), __uint_as_float(r[pc * 8 + 1]));
            qb[pc].y = pack_bf16x2(__uint_as_float(r[pc * 8 + 2]), __uint_as_float(r[pc * 8 + 3]));
            qb[pc].z = pack_bf16x2(__uint_as_float(r[pc * 8 + 4]), __uint_as_float(r[pc * 8 + 5]));
            qb[pc].w = pack_bf16x2(__uint_as_float(r[pc * 8 + 6]), __uint_as_float(r[pc * 8 + 7]));
          }
          row_sts(stage_g, qb, lm);
        }
        if (MODE >= 1 || TMA_OUT) {
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {  // rows past M / columns past N are clipped by the TMA unit
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmOut)),
                         "r"(n), "r"(row0), "r"(stage)
                         : "memory");
            if (has_out2)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                               reinterpret_cast<uint64_t>(&tmOut2)),
                           "r"(n), "r"(row0), "r"(stage_g)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          __syncwarp();
        }
        if (stats) {
          // row 2i + h, 4 bytes at column pair p: piece p >> 2 (swizzled by ((2i + h) >> 1) & 3 = i & 3), word p & 3
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            w[i] = lds32(stage + st_base + (uint32_t)(i * 128) + ((uint32_t)((p >> 2) ^ (i & 3)) << 4));
          if (bwd) {
            uint32_t wy[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              wy[i] = lds32(stage_y + st_base + (uint32_t)(i * 128) + ((uint32_t)((p >> 2) ^ (i & 3)) << 4));
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float2 x0 = unpack16<OF16>(w[i]), x1 = unpack16<OF16>(w[i + 1]);
              s_a[j] = add2(s_a[j], x0); s_b[j] = add2(s_b[j], x1);
              q_a[j] = fma2(x0, add2(unpack16<AF16>(wy[i]), neg_mu[j]), q_a[j]);
              q_b[j] = fma2(x1, add2(unpack16<AF16>(wy[i + 1]), neg_mu[j]), q_b[j]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float2 x0 = unpack16<OF16>(w[i]), x1 = unpack16<OF16>(w[i + 1]);
              s_a[j] = add2(s_a[j], x0); s_b[j] = add2(s_b[j], x1);
              q_a[j] = fma2(x0, x0, q_a[j]); q_b[j] = fma2(x1, x1, q_b[j]);
            }
          }
        }
        if (MODE == 0 && !TMA_OUT) {
          // coalesced write-back: 8 rows x 64 contiguous bytes per instruction
          uint4 o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = lds128(stage + lm.co_off + i * 512);
          __syncwarp();  // every lane has read the staging buffer: the next chunk may overwrite it
          uint8_t* gp = reinterpret_cast<uint8_t*>(ep.out) + ((long long)row0 * ep.ldo + n) * 2 +
                        lm.co_row * (long long)ep.ldo * 2 + lm.co_byte;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i * 8 + lm.co_row < rows_valid) stg128(gp + (long long)i * 8 * ep.ldo * 2, o[i]);
        } else {
          __syncwarp();  // every lane has read the staging buffers: the next operand loads may overwrite them
        }
      }
    }
    if ((MODE >= 1 || TMA_OUT) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete
    __syncwarp();

    if (stats) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int n = n0 + c0 + j * 128;
        float2 s = add2(s_a[j], s_b[j]), qq = add2(q_a[j], q_b[j]);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, 16); s.y += __shfl_xor_sync(0xffffffffu, s.y, 16);
        qq.x += __shfl_xor_sync(0xffffffffu, qq.x, 16); qq.y += __shfl_xor_sync(0xffffffffu, qq.y, 16);
        if (bwd && n < N) {  // sum(dz * xhat) = invstd * sum(dz * (y - mean)); linear: per-CTA partials are fine
          const float2 is = __ldg(reinterpret_cast<const float2*>(ep.stat_invstd + n + 2 * p));
          qq.x *= is.x;
          qq.y *= is.y;
        }
        if (h == 0) *reinterpret_cast<float4*>(s_fin + e * kFinStride + j * 64 + p * 4) = make_float4(s.x, qq.x, s.y, qq.y);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const int c = e * 32 + lane;  // column of the CTA's n-tile handled by this thread
      if (c < BN && n0 + c < N) {
        const int chunk = c >> 5, within = c & 31;
        const int jj = chunk >> 2;  // BN = 256: chunks 4..7 are the second chunk of their warps
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int w = 0; w < kConvEpiWarps; ++w) {
          const int w_cg = w >> 2;
          const int w_chunk = BN == 64 ? (w_cg & 1) : w_cg;
          if (w_chunk == (chunk & 3)) {
            const float2 v =
                *reinterpret_cast<const float2*>(s_fin + w * kFinStride + jj * 64 + (within >> 1) * 4 + (within & 1) * 2);
            ts += v.x; tq += v.y;
          }
        }
        atomicAdd(ep.col_sum + n0 + c, ts);
        atomicAdd(ep.col_sumsq + n0 + c, tq);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer may still arrive on this CTA's barriers / read its B half until it is done
  if (warp == 1) {
    if (CTA2) tmem_dealloc2(tmem_base, ACC * BN);
    else tmem_dealloc(tmem_base, ACC * BN);
  }
}

}  // namespace koa
