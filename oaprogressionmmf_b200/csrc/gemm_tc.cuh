// gemm_tc.cuh — tcgen05/TMEM GEMM core shared by the 1x1/kxk convolutions and the
// transformer linears of the koafusion hot path.
//
// One CTA computes a 128 x BN fp32 accumulator tile held in TMEM:
//   warp 0     : TMA producer (tiled 2-D loads, or im2col-mode loads of an NHWC tensor)
//   warp 1     : TMEM allocation + single-thread tcgen05.mma issue
//   warps 2..  : epilogue (tcgen05.ld -> registers -> fused epilogue -> global): 8 warps in gemm_kmajor_kernel (general
//                epilogue: transformer Linears), 4 in gemm_wgrad_kernel; the convolution flavour with 16 epilogue warps,
//                TMA epilogue I/O and CTA pairs is gemm_conv.cuh
// Operands are 16-bit (bf16 or fp16) staged in 128-byte-swizzled shared memory, STAGES deep.
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
  const bf16* gate_bf16;  // optional: the final value is zeroed where gate <= 0 (ReLU backward)
  bf16* out_bf16_copy;    // optional bf16 copy of the final value (when out is fp32)
  float* col_sum;         // optional per-column sum of the bf16-rounded output and, in col_sumsq, either its sum of
  float* col_sumsq;       // squares (forward BatchNorm statistics) or, when stat_y is set, sum of out * xhat(stat_y)
  const bf16* stat_y;     // BatchNorm backward: forward conv output y [M, ldo]; xhat = (y - mean) * invstd
  const float* stat_mean;
  const float* stat_invstd;
  int a_f16, b_f16;       // operand format (0 = bf16, 1 = fp16); the launcher enforces a_f16 == b_f16
  int out_f16;            // out / pre_out (16-bit outputs) are fp16 instead of bf16
  int act_f16;            // gate and stat_y (forward activations) are fp16 instead of bf16
  // convolution flavour only (gemm_conv.cuh):
  const float* bn_scale;  // MODE 2: value = acc * bn_scale[col] + bn_shift[col] (+ add_bf16 as the 16-bit residual in the
  const float* bn_shift;  //   act_f16 format, optionally * res_scale[col] + res_shift[col]); act = ACT_RELU; fp16 output
  const float* res_scale; //   (+ out_bf16_copy)
  const float* res_shift;
  const float* col_bias;  // MODE 1: added per column before the gate
  int drop_on;            // dropout (applied after the activation, before the residual add); element index of the
  int drop_cols;          // mask = row * drop_cols + column
  DropSpec drop;
};

constexpr int kGemmThreads = 192;     // wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kEpiWarps = 8;          // fprop/dgrad kernel: 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kKmajorThreads = 64 + 32 * kEpiWarps;
constexpr int BM = 128;
constexpr int BK = 64;

constexpr int kMaxStatCols = 2048;  // widest BatchNorm in the networks (layer4 output)

// ---- coalesced epilogue traffic --------------------------------------------------------------------
// After tcgen05.ld a thread owns 32 consecutive columns of ONE row, so direct global accesses would touch 16-byte
// pieces of 32 different rows per instruction (half-filled sectors). Every [32 rows][64 bytes] sub-tile therefore
// moves between global memory and the row-per-thread register layout through a per-warp shared-memory staging
// buffer (16-byte pieces XOR-swizzled by (row >> 1) & 3: the row-wise accesses and the transposed ones are both
// bank-conflict free); the global side is always 8 rows x 64 contiguous bytes per instruction.
constexpr int kStageBytesPerWarp = 32 * 64;
constexpr int kStagesPerWarp = 2;

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return 1024 /*align slack*/ + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 4) * 8 + 16 +
         2 * kMaxStatCols * sizeof(float) + 2 * BN * 4 * 2 * sizeof(float) /*per-tile partial statistics*/ +
         kEpiWarps * kStagesPerWarp * kStageBytesPerWarp + 32;
}

// ---- shared-memory accessors in the shared state space (generic LD/ST on staging pointers costs a long-scoreboard
// round trip per access; these compile to LDS / STS) ------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64f(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint4 ldg128_nc(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(uint8_t* p, const uint4& v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t stage_off(int row, int piece) {
  return (uint32_t)(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
}

// Per-lane constants of the two access patterns of a [32 rows][64 bytes] staging tile.
struct LaneMap {
  uint32_t row_off[4];  // row-per-thread pattern: byte offset of piece p of row `lane`
  uint32_t co_off;      // coalesced pattern: byte offset of (row lane >> 2, piece lane & 3); rows 8 apart are +512
  int co_row;           // lane >> 2
  int co_byte;          // (lane & 3) * 16
};
__device__ __forceinline__ LaneMap make_lane_map(int lane) {
  LaneMap m;
#pragma unroll
  for (int p = 0; p < 4; ++p) m.row_off[p] = stage_off(lane, p);
  m.co_off = stage_off(lane >> 2, lane & 3);  // ((8i + r) >> 1) & 3 == (r >> 1) & 3: the swizzle does not depend on i
  m.co_row = lane >> 2;
  m.co_byte = (lane & 3) * 16;
  return m;
}

// global -> registers, coalesced (8 rows x 64 contiguous bytes per instruction); rows >= rows_valid read 0.
__device__ __forceinline__ void tile_ldg(uint4 (&t)[4], const uint8_t* gbase, long long pitch_bytes, int rows_valid,
                                         const LaneMap& lm) {
  const uint8_t* p = gbase + lm.co_row * pitch_bytes + lm.co_byte;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t[i] = make_uint4(0, 0, 0, 0);
    if (i * 8 + lm.co_row < rows_valid) t[i] = ldg128_nc(p + i * 8 * pitch_bytes);
  }
}
__device__ __forceinline__ void tile_sts(uint32_t stage, const uint4 (&t)[4], const LaneMap& lm) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts128(stage + lm.co_off + i * 512, t[i]);
}
__device__ __forceinline__ void row_lds(uint4 (&q)[4], uint32_t stage, const LaneMap& lm) {
#pragma unroll
  for (int p = 0; p < 4; ++p) q[p] = lds128(stage + lm.row_off[p]);
}
__device__ __forceinline__ void row_sts(uint32_t stage, const uint4 (&q)[4], const LaneMap& lm) {
#pragma unroll
  for (int p = 0; p < 4; ++p) sts128(stage + lm.row_off[p], q[p]);
}
__device__ __forceinline__ void unpack_row_bf16(float (&f)[32], const uint4 (&q)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float2 a;
    a = unpack_bf16x2(q[p].x); f[p * 8 + 0] = a.x; f[p * 8 + 1] = a.y;
    a = unpack_bf16x2(q[p].y); f[p * 8 + 2] = a.x; f[p * 8 + 3] = a.y;
    a = unpack_bf16x2(q[p].z); f[p * 8 + 4] = a.x; f[p * 8 + 5] = a.y;
    a = unpack_bf16x2(q[p].w); f[p * 8 + 6] = a.x; f[p * 8 + 7] = a.y;
  }
}
__device__ __forceinline__ void pack_row_16(uint4 (&q)[4], const float (&v)[32], bool f16) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    q[p].x = pack16x2(v[p * 8 + 0], v[p * 8 + 1], f16); q[p].y = pack16x2(v[p * 8 + 2], v[p * 8 + 3], f16);
    q[p].z = pack16x2(v[p * 8 + 4], v[p * 8 + 5], f16); q[p].w = pack16x2(v[p * 8 + 6], v[p * 8 + 7], f16);
  }
}
// staged coalesced tile -> this thread's row as 32 floats (whole warp; syncs on both sides)
__device__ __forceinline__ void tile_to_row_bf16(float (&f)[32], uint32_t stage, const uint4 (&t)[4], const LaneMap& lm,
                                                 bool f16 = false) {
  tile_sts(stage, t, lm);
  __syncwarp();
  uint4 q[4];
  row_lds(q, stage, lm);
  __syncwarp();
  if (!f16) {
    unpack_row_bf16(f, q);
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float2 a;
      a = unpack_f16x2(q[p].x); f[p * 8 + 0] = a.x; f[p * 8 + 1] = a.y;
      a = unpack_f16x2(q[p].y); f[p * 8 + 2] = a.x; f[p * 8 + 3] = a.y;
      a = unpack_f16x2(q[p].z); f[p * 8 + 4] = a.x; f[p * 8 + 5] = a.y;
      a = unpack_f16x2(q[p].w); f[p * 8 + 6] = a.x; f[p * 8 + 7] = a.y;
    }
  }
}

// Coalesced write-back of a staged [32 rows][64 bytes] tile; rows >= rows_valid are skipped.
__device__ __forceinline__ void stage_flush(uint32_t stage, uint8_t* gbase, long long pitch_bytes, int rows_valid,
                                            const LaneMap& lm) {
  __syncwarp();
  uint8_t* p = gbase + lm.co_row * pitch_bytes + lm.co_byte;
  uint4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = lds128(stage + lm.co_off + i * 512);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i * 8 + lm.co_row < rows_valid) stg128(p + i * 8 * pitch_bytes, v[i]);
  __syncwarp();
}

__device__ __forceinline__ void store_row_16(const float (&v)[32], uint32_t stage, bf16* g, long long ld, int rows_valid,
                                             const LaneMap& lm, bool f16) {
  uint4 q[4];
  pack_row_16(q, v, f16);
  row_sts(stage, q, lm);
  stage_flush(stage, reinterpret_cast<uint8_t*>(g), ld * 2, rows_valid, lm);
}

__device__ __forceinline__ void store_row_f32(const float (&v)[32], uint32_t stage, float* g, long long ld, int rows_valid,
                                              const LaneMap& lm) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // two 16-column halves of 64 bytes each
#pragma unroll
    for (int p = 0; p < 4; ++p)
      sts128(stage + lm.row_off[p], make_uint4(__float_as_uint(v[h * 16 + p * 4 + 0]), __float_as_uint(v[h * 16 + p * 4 + 1]),
                                               __float_as_uint(v[h * 16 + p * 4 + 2]), __float_as_uint(v[h * 16 + p * 4 + 3])));
    stage_flush(stage, reinterpret_cast<uint8_t*>(g + h * 16), ld * 4, rows_valid, lm);
  }
}

// Epilogue of one 32-row x 32-column chunk. The thread owns row row0 + lane of the warp's 32-row group; `n` = first
// global column; `taddr` = TMEM address of the chunk. `release_bar` (or NULL) is arrived on as soon as the
// accumulator has been read (last chunk of this warp for the tile). `stage` = shared-space address of this warp's
// two 2 KB staging buffers; `part` = shared-space address of this tile's partial-statistics slots
// [BN columns][4 row quarters][2].
// CONV = true compiles the convolution flavour only (bf16 output, optional addend / gate / statistics; no bias,
// activation, fp32 paths): fewer live registers and branches for the epilogue-bound small-K layers.
template <bool CONV>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, uint64_t* release_bar, const EpiParams& ep, long long row0,
                                               int rows_valid, int n, int lane, const LaneMap& lm, uint32_t stage,
                                               uint32_t part, int c_local, int quarter) {
  const uint32_t stage1 = stage + kStageBytesPerWarp;
  const long long tile_off = row0 * ep.ldo + n;  // element offset of this warp tile in every [M, ldo] tensor
  const long long pitch2 = (long long)ep.ldo * 2;
  const bool has_add = ep.add_bf16 != nullptr, has_gate = ep.gate_bf16 != nullptr;
  const bool stats = ep.col_sum != nullptr, bwd = ep.stat_y != nullptr;
  // issue the global reads of the epilogue operands first: they overlap the TMEM load
  uint4 t_add[4], t_gate[4], t_y[4];  // t_y doubles as the GELU' operand (never used together with stat_y)
  if (!CONV && ep.act == ACT_GELU_GRAD) tile_ldg(t_y, reinterpret_cast<const uint8_t*>(ep.aux + tile_off), pitch2, rows_valid, lm);
  else if (bwd) tile_ldg(t_y, reinterpret_cast<const uint8_t*>(ep.stat_y + tile_off), pitch2, rows_valid, lm);
  if (has_add) tile_ldg(t_add, reinterpret_cast<const uint8_t*>(ep.add_bf16 + tile_off), pitch2, rows_valid, lm);
  if (has_gate) tile_ldg(t_gate, reinterpret_cast<const uint8_t*>(ep.gate_bf16 + tile_off), pitch2, rows_valid, lm);

  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  if (release_bar != nullptr) {  // hand the accumulator back before the global traffic of this chunk
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(release_bar);
  }
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

  if (!CONV) {
    if (ep.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (ep.pre_out != nullptr) store_row_16(v, stage, ep.pre_out + tile_off, ep.ldo, rows_valid, lm, ep.out_f16 != 0);
    if (ep.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (ep.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (ep.act == ACT_GELU_GRAD) {
      float h[32];
      tile_to_row_bf16(h, stage, t_y, lm, ep.act_f16 != 0);  // the saved pre-activation is a forward tensor
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(h[j]);
    }
    if (ep.drop_on) {
      const unsigned long long quad0 = ((unsigned long long)(row0 + lane) * (unsigned long long)ep.drop_cols + n) >> 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s[4];
        drop_scales4(ep.drop, quad0 + j, s);
        v[4 * j] *= s[0]; v[4 * j + 1] *= s[1]; v[4 * j + 2] *= s[2]; v[4 * j + 3] *= s[3];
      }
    }
    if (ep.res_f32 != nullptr) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // two 16-column halves of 64 bytes each
        uint4 t[4], q[4];
        tile_ldg(t, reinterpret_cast<const uint8_t*>(ep.res_f32 + tile_off + hh * 16), (long long)ep.ldo * 4, rows_valid, lm);
        tile_sts(stage, t, lm);
        __syncwarp();
        row_lds(q, stage, lm);
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          v[hh * 16 + p * 4 + 0] += __uint_as_float(q[p].x); v[hh * 16 + p * 4 + 1] += __uint_as_float(q[p].y);
          v[hh * 16 + p * 4 + 2] += __uint_as_float(q[p].z); v[hh * 16 + p * 4 + 3] += __uint_as_float(q[p].w);
        }
      }
    }
  }
  if (has_add || has_gate) {  // both operand tiles take one trip through the two staging buffers
    if (has_add) tile_sts(stage, t_add, lm);
    if (has_gate) tile_sts(stage1, t_gate, lm);
    __syncwarp();
    uint4 qa[4], qg[4];
    if (has_add) row_lds(qa, stage, lm);
    if (has_gate) row_lds(qg, stage1, lm);
    __syncwarp();
    if (has_add) {
      float a[32];
      unpack_row_bf16(a, qa);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += a[j];
    }
    if (has_gate) {  // format-agnostic: the sign / zero test reads the 16-bit patterns directly
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint32_t w[4] = {qg[p].x, qg[p].y, qg[p].z, qg[p].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[p * 8 + 2 * u] = pos16(w[u] & 0xffffu) ? v[p * 8 + 2 * u] : 0.0f;
          v[p * 8 + 2 * u + 1] = pos16(w[u] >> 16) ? v[p * 8 + 2 * u + 1] : 0.0f;
        }
      }
    }
  }
  if (!CONV && ep.out_fp32) {
    store_row_f32(v, stage, reinterpret_cast<float*>(ep.out) + tile_off, ep.ldo, rows_valid, lm);
    if (ep.out_bf16_copy != nullptr) store_row_16(v, stage, ep.out_bf16_copy + tile_off, ep.ldo, rows_valid, lm, false);
    return;
  }
  // bf16 output: stage, (statistics from the staged, i.e. rounded, values), coalesced write-back
  const bool of16 = ep.out_f16 != 0, af16 = ep.act_f16 != 0;
  uint4 q[4];
  pack_row_16(q, v, of16);
  if (stats && lane >= rows_valid) {  // rows past M contribute exact zeros to the statistics
#pragma unroll
    for (int p = 0; p < 4; ++p) q[p] = make_uint4(0, 0, 0, 0);
  }
  row_sts(stage, q, lm);
#ifdef KOA_EXP_NO_STATS_LOOP
  if (false) {
#else
  if (stats) {
#endif
    if (bwd) tile_sts(stage1, t_y, lm);
    __syncwarp();
    // lane = (h, p): column pair (2p, 2p+1) over the 16 rows 2i + h; the two lane halves read rows in different
    // bank halves, each half reads one contiguous 64-byte row per step (conflict free)
    const int h = lane >> 4, p = lane & 15;
    float mu0 = 0.0f, mu1 = 0.0f, is0 = 1.0f, is1 = 1.0f;
    if (bwd) {
      const float2 mu = __ldg(reinterpret_cast<const float2*>(ep.stat_mean + n + 2 * p));
      const float2 is = __ldg(reinterpret_cast<const float2*>(ep.stat_invstd + n + 2 * p));
      mu0 = mu.x; mu1 = mu.y; is0 = is.x; is1 = is.y;
    }
    // row 2i + h, 4 bytes at column pair p: piece p >> 2 (swizzled by ((2i + h) >> 1) & 3 = i & 3), word p & 3.
    // All shared-memory reads are issued before the first use (a load-use pair per iteration serialises on the
    // LDS latency), the two flavours are separate straight-line loops, and each sum has two independent chains.
    const uint32_t base = stage + (uint32_t)(h * 64 + (p & 3) * 4);
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = lds32(base + (uint32_t)(i * 128) + ((uint32_t)((p >> 2) ^ (i & 3)) << 4));
    float sa[2] = {0.0f, 0.0f}, sb[2] = {0.0f, 0.0f}, qa[2] = {0.0f, 0.0f}, qb[2] = {0.0f, 0.0f};
    if (bwd) {
      uint32_t wy[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        wy[i] = lds32(base + kStageBytesPerWarp + (uint32_t)(i * 128) + ((uint32_t)((p >> 2) ^ (i & 3)) << 4));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 x = unpack16x2(w[i], of16), y = unpack16x2(wy[i], af16);
        sa[i & 1] += x.x; sb[i & 1] += x.y;
        qa[i & 1] = fmaf(x.x, (y.x - mu0) * is0, qa[i & 1]); qb[i & 1] = fmaf(x.y, (y.y - mu1) * is1, qb[i & 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 x = unpack16x2(w[i], of16);
        sa[i & 1] += x.x; sb[i & 1] += x.y;
        qa[i & 1] = fmaf(x.x, x.x, qa[i & 1]); qb[i & 1] = fmaf(x.y, x.y, qb[i & 1]);
      }
    }
    float s0 = sa[0] + sa[1], s1 = sb[0] + sb[1], q0 = qa[0] + qa[1], q1 = qb[0] + qb[1];
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
    if (h == 0) {  // this warp's 32-row partial of columns (c_local + 2p, + 2p + 1): private slot, no atomics
      const uint32_t dst = part + (uint32_t)(((c_local + 2 * p) * 4 + quarter) * 8);
      sts64f(dst, s0, q0);
      sts64f(dst + 32, s1, q1);
    }
  }
  stage_flush(stage, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(ep.out) + tile_off), pitch2, rows_valid, lm);
}

// out[M, N] = A[M, K] * B[N, K]^T with the fused epilogue. A_IM2COL: A rows are the output
// pixels of a convolution over an NHWC tensor and K enumerates (r, s, cin).
//
// Persistent: gridDim.x = min(#tiles, #SMs) CTAs walk the tile list (n-tile fastest, so CTAs running at the
// same time share their A rows through L2). Three pipelines overlap across tiles:
//   TMA -> smem ring (STAGES deep, full/empty mbarriers, runs ahead into the next tile),
//   MMA -> TMEM accumulator ring (2 x BN columns, tmem_full/tmem_empty mbarriers),
//   epilogue warps drain accumulator i while the MMA of tile i+1 is issued.
template <int BN, int STAGES, bool A_IM2COL, bool CONV_EPI>
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
  float* s_part = s_sumsq + kMaxStatCols;  // [2 tiles in flight][BN][4 row quarters][2]
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(((uintptr_t)(s_part + 2 * BN * 8) + 15) & ~(uintptr_t)15);

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
      const uint32_t idesc = umma_idesc_16(BM, BN, 0, 0, ep.a_f16, ep.b_f16);
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
    const uint32_t stage = smem_u32(s_stage) + e * (kStagesPerWarp * kStageBytesPerWarp);
    const LaneMap lm = make_lane_map(lane);
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
      float* part = s_part + (lt & 1) * (BN * 8);
      const uint32_t part_s = smem_u32(part);
      bool released = false;
#pragma unroll 1
      for (int c0 = chunk_par * 32; c0 < BN; c0 += 64) {
        if (n0 + c0 >= N) break;
        // last TMEM read of this warp for this tile: the accumulator is handed back inside epilogue_chunk
        const bool last = (c0 + 64 >= BN || n0 + c0 + 64 >= N);
        epilogue_chunk<CONV_EPI>(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0,
                                 last ? &tmem_empty_bar[acc] : nullptr, ep, row0, rows_valid, n0 + c0, lane, lm, stage,
                                 part_s, c0, q);
        released = released || last;
      }
      if (!released) {  // this warp had no chunk inside N
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (stats) {
        // Combine the four row-quarter partials of every column of this tile. Column n0 + c is always owned by
        // epilogue thread c, so the running per-CTA totals need no atomics; `part` is double buffered, which makes
        // one barrier per tile sufficient.
#ifndef KOA_EXP_NO_TILE_BARRIER
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
#endif
        if (ep_tid < BN && n0 + ep_tid < N) {
          const float4 a = *reinterpret_cast<const float4*>(part + ep_tid * 8);
          const float4 b = *reinterpret_cast<const float4*>(part + ep_tid * 8 + 4);
          s_sum[n0 + ep_tid] += (a.x + a.z) + (b.x + b.z);
          s_sumsq[n0 + ep_tid] += (a.y + a.w) + (b.y + b.w);
        }
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
  return 1024 + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (3 * STAGES + 1) * 8 + 16;
}

constexpr int kWgradCvtWarps = 8;  // XCVT: warps 2..9 convert the activation tile (fp16 -> bf16) in shared memory
constexpr int kWgradCvtThreads = 64 + 32 * kWgradCvtWarps;

__device__ __forceinline__ uint32_t f16x2_to_bf16x2(uint32_t v) {
  const float2 f = unpack_f16x2(v);
  return pack_bf16x2(f.x, f.y);
}

// XCVT: the activation operand X arrives in fp16 (the forward storage format) and is rewritten as bf16 in shared memory
// between the TMA load and the MMA (one tcgen05 kind::f16 MMA takes ONE 16-bit format and dY is bf16): the forward pass
// does not have to write a second, bf16 copy of every activation for the weight gradients.
// CTA2: CTA pairs (cluster of 2 along Cout, tcgen05 cta_group::2, M = 256): each CTA stages its own 128 dY columns and
// HALF of the BN activation columns, like the convolution kernel's pairs (gemm_conv.cuh); rank 0 issues the MMAs.
template <int BN, int STAGES, bool B_IM2COL, bool XCVT, bool CTA2 = false>
__global__ void __launch_bounds__(XCVT ? kWgradCvtThreads : kGemmThreads)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int cout, int cin,
                  int pixels, int taps, ConvGeom g, float* __restrict__ dw, int kb_per_split, WgradDesc wd, int a_f16,
                  int b_f16) {
  // g.grouped: dw is [C][taps][64] (per-64-channel-chunk dense blocks); tile n_t pairs input chunk n_t with
  // output chunk n_t only (BN must be 64; the upper 64 accumulator rows are discarded).
  static_assert(!(CTA2 && XCVT), "the in-kernel conversion is not combined with CTA pairs");
  constexpr int BN_LOCAL = CTA2 ? BN / 2 : BN;  // activation columns staged by this CTA
  constexpr int BMT = CTA2 ? 2 * BM : BM;       // Cout rows of the tile of a CTA (pair)
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN_LOCAL * BK * 2;
  constexpr uint32_t BOX_BYTES = 64 * BK * 2;  // one {64 channels x 64 pixels} TMA box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* xf_bar = tmem_full_bar + 1;  // [STAGES] XCVT: the converted tile is visible to the tensor core
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_bar + STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int n_tiles = (cin + BN - 1) / BN;
  const int m_tiles = g.grouped ? 1 : (cout + BMT - 1) / BMT;
  int t = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_t = t % n_tiles; t /= n_tiles;
  const int m_t = t % m_tiles; t /= m_tiles;
  const int tap = t;
  const int m0 = g.grouped ? n_t * 64 : m_t * BMT + (int)rank * BM;  // first Cout row of this CTA
  const int n0 = n_t * BN + (int)rank * BN_LOCAL;                     // first activation column staged by this CTA
  const int n0_tile = n_t * BN;
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
      mbar_init(&xf_bar[s], kWgradCvtWarps);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc2(tmem_slot, BN);
    else tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
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
        // pair: rank 0 announces the bytes of both CTAs; the peer's loads complete on rank 0's barrier
        if (!CTA2) mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        else if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_BYTES + B_BYTES));
        const int p0 = kb * BK;
#pragma unroll
        for (int j = 0; j < BM / 64; ++j) {
          if (CTA2) tma2_load_2d(sA + s * A_BYTES + j * BOX_BYTES, &tmA, &full_bar[s], m0 + j * 64, p0);
          else tma_load_2d(sA + s * A_BYTES + j * BOX_BYTES, &tmA, &full_bar[s], m0 + j * 64, p0);
        }
        if (B_IM2COL) {
          const int pn = p0 / hw;
          const int rem = p0 - pn * hw;
          const int oh = rem / g.wout;
          const int ow = rem - oh * g.wout;
#pragma unroll
          for (int j = 0; j < BN_LOCAL / 64; ++j) {
            if (CTA2)
              tma2_load_im2col_4d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64,
                                  ow * g.stride - g.pad, oh * g.stride - g.pad, pn, (uint16_t)fs, (uint16_t)fr);
            else
              tma_load_im2col_4d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64,
                                 ow * g.stride - g.pad, oh * g.stride - g.pad, pn, (uint16_t)fs, (uint16_t)fr);
          }
        } else {
#pragma unroll
          for (int j = 0; j < BN_LOCAL / 64; ++j) {
            if (CTA2) tma2_load_2d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64, p0);
            else tma_load_2d(sB + s * B_BYTES + j * BOX_BYTES, &tmB, &full_bar[s], n0 + j * 64, p0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_16(BMT, BN, 1, 1, a_f16, b_f16);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(XCVT ? &xf_bar[s] : &full_bar[s], phase, 0x500 + s);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), wd.lbo, wd.sbo);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), wd.lbo, wd.sbo);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * wd.k_adv) >> 4);
          if (CTA2) umma2_bf16_ss(tmem_base, adesc + adv, bdesc + adv, idesc, (i | k) != 0);
          else umma_bf16_ss(tmem_base, adesc + adv, bdesc + adv, idesc, (i | k) != 0);
        }
        if (CTA2) umma2_commit_both(&empty_bar[s]);
        else umma_commit(&empty_bar[s]);
      }
      if (CTA2) umma2_commit_both(tmem_full_bar);
      else umma_commit(tmem_full_bar);
    }
  } else {
    if (XCVT) {
      // fp16 -> bf16 in place, 16 bytes per thread and step; the swizzled layout is untouched (element-wise)
      const int t = threadIdx.x - 64;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], phase, 0x800 + s);
        const uint32_t tile = smem_u32(sB + s * B_BYTES);
        constexpr int kVec = B_BYTES / 16;
        constexpr int kPer = kVec / (kWgradCvtWarps * 32);
        static_assert(kVec % (kWgradCvtWarps * 32) == 0, "tile divides over the converting threads");
        uint4 v[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) v[j] = lds128(tile + (uint32_t)(t + j * kWgradCvtWarps * 32) * 16);
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          v[j].x = f16x2_to_bf16x2(v[j].x); v[j].y = f16x2_to_bf16x2(v[j].y);
          v[j].z = f16x2_to_bf16x2(v[j].z); v[j].w = f16x2_to_bf16x2(v[j].w);
          sts128(tile + (uint32_t)(t + j * kWgradCvtWarps * 32) * 16, v[j]);
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[s]);
      }
      if (warp >= 6) goto done;  // warps 6..9 only convert
    }
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0, 0x600);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;  // output channel
    const bool row_ok = g.grouped ? (q * 32 + lane) < 64 : row < cout;
    const int ldw = g.grouped ? 64 : cin;
    const int col0 = g.grouped ? 0 : n0_tile;
    float* dst_row = dw + ((long long)row * taps + tap) * ldw;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0_tile + c0 >= cin) break;
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
done:
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's MMAs read this CTA's tiles and arrive on its barriers until it is done
  if (warp == 1) {
    if (CTA2) tmem_dealloc2(tmem_base, BN);
    else tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace koa
