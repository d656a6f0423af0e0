// step_ops.cu — the kernels on either side of the hot path: the optimiser step that follows backward
// (torch.optim.Adam / AdamW as koafusion/run/train_prog_fus.py:88-91,166 uses them), the "last-chance preprocessing"
// in front of the models (PTToUnitRange + PTNormalize + PTInterpolate, koafusion/preproc/_pt.py:75-124,175-200 and
// run/train_prog_fus.py:111-116,143-146) on integer volumes as they sit on disk, and the prediction / fold-ensemble
// arithmetic behind them (run/eval_prog_fus.py:300-304,330-336). All of it is HBM-bound streaming work:
// coalesced 16-byte accesses, grid-stride loops over a few CTAs per SM, no tensor cores.
#include <math.h>

#include "koa_common.cuh"
#include "koa_internal.h"
#include "step_arith.cuh"

namespace {

constexpr int kThreads = 256;

inline int capped_grid(long long blocks, int per_sm) {
  const long long cap = (long long)koa_num_sms() * per_sm;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

// ======================================================================================================
// Adam / AdamW over many tensors per launch
// ======================================================================================================
constexpr int kAdamSlots = 64;     // tensors per launch (the table travels as a kernel parameter: no device table,
                                   // no host->device copy, nothing to keep alive)
constexpr int kAdamChunk = 8192;   // elements per CTA visit: 8 float4 per thread and tensor stream

struct AdamTable {
  float* p[kAdamSlots];
  const float* g[kAdamSlots];
  float* m[kAdamSlots];
  float* v[kAdamSlots];
  long long n[kAdamSlots];
  int chunk_begin[kAdamSlots + 1];  // strictly increasing: every slot has at least one chunk
  int count;
  unsigned long long vec_mask;      // bit i: the four pointers of slot i are 16-byte aligned
};

using koa_arith::AdamCoef;
using koa_arith::adam_update;

__global__ void __launch_bounds__(kThreads) adam_kernel(const __grid_constant__ AdamTable tab,
                                                         const __grid_constant__ AdamCoef c) {
  const int total = tab.chunk_begin[tab.count];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    int lo = 0, hi = tab.count - 1;
    while (lo < hi) {  // last slot whose first chunk <= chunk
      const int mid = (lo + hi + 1) >> 1;
      if (tab.chunk_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const long long base = (long long)(chunk - tab.chunk_begin[lo]) * kAdamChunk;
    const long long left = tab.n[lo] - base;
    const int len = left < kAdamChunk ? (int)left : kAdamChunk;
    float* __restrict__ p = tab.p[lo] + base;
    const float* __restrict__ g = tab.g[lo] + base;
    float* __restrict__ m = tab.m[lo] + base;
    float* __restrict__ v = tab.v[lo] + base;
    int done = 0;
    if ((tab.vec_mask >> lo) & 1ull) {  // base is a multiple of 4 elements: the alignment of the slot carries over
      const int nvec = len >> 2;
#pragma unroll 2
      for (int i = threadIdx.x; i < nvec; i += kThreads) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_update(pp.x, gg.x, mm.x, vv.x, c);
        adam_update(pp.y, gg.y, mm.y, vv.y, c);
        adam_update(pp.z, gg.z, mm.z, vv.z, c);
        adam_update(pp.w, gg.w, mm.w, vv.w, c);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
      }
      done = nvec << 2;
    }
    for (int i = done + threadIdx.x; i < len; i += kThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(pp, g[i], mm, vv, c);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

// ======================================================================================================
// linear resampling of (B, D0, D1, D2) volumes, D2 innermost (F.interpolate, align_corners=False, scales recomputed
// from the sizes) with an optional per-volume affine map (unit range + z-score) folded in
// ======================================================================================================
using koa_arith::Dims3;
using koa_arith::ld_f;

// the per-element arithmetic lives in step_arith.cuh (shared with the CPU check of tests/host_emul)
template <typename T>
__global__ void __launch_bounds__(kThreads) resample_linear_kernel(const T* __restrict__ in, float* __restrict__ out,
                                                                    long long total, Dims3 di, Dims3 dout, float rs0,
                                                                    float rs1, float rs2, const float* __restrict__ scale,
                                                                    const float* __restrict__ shift) {
  const long long in_per = (long long)di.d0 * di.d1 * di.d2;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    int x0, x1, x2;
    const long long b = koa_arith::split_index(i, dout, x0, x1, x2);
    float val = koa_arith::resample_element<T>(in + b * in_per, di, x0, x1, x2, rs0, rs1, rs2);
    if (scale != nullptr) val = fmaf(val, __ldg(scale + b), __ldg(shift + b));
    out[i] = val;
  }
}

// ======================================================================================================
// per-volume minimum / maximum -> the affine map of PTToUnitRange followed by PTNormalize
// ======================================================================================================
__device__ __forceinline__ unsigned ordered_u32(float f) {  // monotone float -> unsigned map
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(unsigned* mm, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_kernel(const T* __restrict__ in, long long n_per, int vec_ok,
                                                           unsigned* __restrict__ mm) {
  constexpr int V = 16 / (int)sizeof(T);
  const T* base = in + (long long)blockIdx.y * n_per;
  float lo = INFINITY, hi = -INFINITY;
  long long done = 0;
  if (vec_ok) {
    const long long nvec = n_per / V;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < nvec; i += (long long)gridDim.x * kThreads) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(base) + i);
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int u = 0; u < V; ++u) {
        const float f = (float)e[u];
        lo = fminf(lo, f); hi = fmaxf(hi, f);
      }
    }
    done = nvec * V;
  }
  for (long long i = done + blockIdx.x * (long long)kThreads + threadIdx.x; i < n_per; i += (long long)gridDim.x * kThreads) {
    const float f = ld_f(base + i);
    lo = fminf(lo, f); hi = fmaxf(hi, f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[kThreads / 32], s_hi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kThreads / 32; ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
    if (lo <= hi) {  // a CTA that saw no element keeps its +-inf out of the result
      atomicMin(mm + 2 * blockIdx.y, ordered_u32(lo));
      atomicMax(mm + 2 * blockIdx.y + 1, ordered_u32(hi));
    }
  }
}

__global__ void unit_range_affine_kernel(const unsigned* __restrict__ mm, float mean, float stdev, float* scale, float* shift,
                                         float* minmax_out, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const float lo = from_ordered_u32(mm[2 * i]), hi = from_ordered_u32(mm[2 * i + 1]);
  koa_arith::unit_range_coef(lo, hi, mean, stdev, scale + i, shift + i);
  if (minmax_out != nullptr) { minmax_out[2 * i] = lo; minmax_out[2 * i + 1] = hi; }
}

// ======================================================================================================
// predictions
// ======================================================================================================
// proba = softmax(logits, dim=1), pred = argmax(logits, dim=1) (first maximum); one thread per knee
__global__ void predict_kernel(const float* __restrict__ logits, float* __restrict__ proba, long long* __restrict__ pred,
                               int batch, int classes) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  koa_arith::predict_row(logits + (long long)b * classes, classes, proba ? proba + (long long)b * classes : nullptr,
                         pred ? pred + b : nullptr);
}

// the fold ensemble of the reference (step_arith.cuh); proba is [folds][batch][classes]
__global__ void ensemble_kernel(const float* __restrict__ proba, float* __restrict__ out, long long* __restrict__ pred,
                                int folds, int batch, int classes) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  koa_arith::ensemble_row(proba, folds, batch, classes, b, out ? out + (long long)b * classes : nullptr,
                          pred ? pred + b : nullptr);
}

// ======================================================================================================
// training-time augmentation fused with the resampling: crop -> unit range -> in-slice rotation (bilinear, zero
// padding) -> gamma -> z-score -> linear resampling, one pass over the stored integers
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads) augment_resample_kernel(const T* __restrict__ in, float* __restrict__ out,
                                                                     const koa_augment_t* __restrict__ params,
                                                                     long long total, Dims3 src, Dims3 crop, Dims3 dout,
                                                                     float rs0, float rs1, float rs2, float mean,
                                                                     float stdev) {
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    int x0, x1, x2;
    const long long b = koa_arith::split_index(i, dout, x0, x1, x2);
    const koa_augment_t a = params[b];
    out[i] = koa_arith::augment_element<T>(koa_arith::crop_view(in, b, a, src), a, crop, x0, x1, x2, rs0, rs1, rs2, mean,
                                           stdev);
  }
}

// minimum / maximum over the crop of every stored volume (strided), into the ordered-integer workspace
template <typename T>
__global__ void __launch_bounds__(kThreads) crop_minmax_kernel(const T* __restrict__ in,
                                                                const koa_augment_t* __restrict__ params, int s0, int s1,
                                                                int s2, int d0, int d1, int d2, unsigned* __restrict__ mm) {
  const koa_augment_t a = params[blockIdx.y];
  const koa_arith::CropView<T> v = koa_arith::crop_view(in, (long long)blockIdx.y, a, Dims3{s0, s1, s2});
  const long long n = (long long)d0 * d1 * d2;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    long long t = i;
    const int s = (int)(t % d2); t /= d2;
    const int c = (int)(t % d1);
    const int r = (int)(t / d1);
    const float f = ld_f(v.p + r * v.st_r + c * v.st_c + s * v.st_s);
    lo = fminf(lo, f); hi = fmaxf(hi, f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[kThreads / 32], s_hi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kThreads / 32; ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
    if (lo <= hi) {
      atomicMin(mm + 2 * blockIdx.y, ordered_u32(lo));
      atomicMax(mm + 2 * blockIdx.y + 1, ordered_u32(hi));
    }
  }
}

__global__ void crop_minmax_finish_kernel(const unsigned* __restrict__ mm, koa_augment_t* params, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const float lo = from_ordered_u32(mm[2 * i]), hi = from_ordered_u32(mm[2 * i + 1]);
  params[i].lo = lo;
  params[i].range = hi - lo;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int launch_resample(const void* in, float* out, long long total, const int* di, const int* dout, const float* scale,
                    const float* shift, cudaStream_t st) {
  // the scale PyTorch derives when recompute_scale_factor=True: input size / output size, in float
  const float rs0 = (float)di[0] / (float)dout[0], rs1 = (float)di[1] / (float)dout[1], rs2 = (float)di[2] / (float)dout[2];
  resample_linear_kernel<T><<<capped_grid((total + kThreads - 1) / kThreads, 16), kThreads, 0, st>>>(
      static_cast<const T*>(in), out, total, Dims3{di[0], di[1], di[2]}, Dims3{dout[0], dout[1], dout[2]}, rs0, rs1, rs2, scale,
      shift);
  KOA_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int launch_minmax(const void* in, int batch, long long n_per, unsigned* mm, cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T);
  const int vec_ok = aligned16(in) && n_per % V == 0;
  // enough CTAs over all volumes to fill the machine a few times; at least one per volume
  long long per = (n_per / V + kThreads - 1) / kThreads;
  const long long want = ((long long)koa_num_sms() * 8 + batch - 1) / batch;
  if (per > want) per = want;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)batch);
  minmax_kernel<T><<<grid, kThreads, 0, st>>>(static_cast<const T*>(in), n_per, vec_ok, mm);
  KOA_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int launch_augment(const void* in, float* out, const koa_augment_t* params, int batch, const int* src, const int* crop,
                   const int* dout, float mean, float stdev, unsigned* ws, cudaStream_t st) {
  minmax_init_kernel<<<koa_cdiv(batch, 128), 128, 0, st>>>(ws, batch);
  KOA_LAUNCH_CHECK();
  const long long n = (long long)crop[0] * crop[1] * crop[2];
  long long per = (n + kThreads * 8 - 1) / (kThreads * 8);  // ~8 elements per thread
  const long long want = ((long long)koa_num_sms() * 8 + batch - 1) / batch;
  if (per > want) per = want;
  if (per < 1) per = 1;
  crop_minmax_kernel<T><<<dim3((unsigned)per, (unsigned)batch), kThreads, 0, st>>>(static_cast<const T*>(in), params, src[0],
                                                                                  src[1], src[2], crop[0], crop[1], crop[2], ws);
  KOA_LAUNCH_CHECK();
  crop_minmax_finish_kernel<<<koa_cdiv(batch, 128), 128, 0, st>>>(ws, const_cast<koa_augment_t*>(params), batch);
  KOA_LAUNCH_CHECK();
  const long long total = (long long)batch * dout[0] * dout[1] * dout[2];
  const float rs0 = (float)crop[0] / (float)dout[0], rs1 = (float)crop[1] / (float)dout[1], rs2 = (float)crop[2] / (float)dout[2];
  augment_resample_kernel<T><<<capped_grid((total + kThreads - 1) / kThreads, 16), kThreads, 0, st>>>(
      static_cast<const T*>(in), out, params, total, Dims3{src[0], src[1], src[2]}, Dims3{crop[0], crop[1], crop[2]},
      Dims3{dout[0], dout[1], dout[2]}, rs0, rs1, rs2, mean, stdev);
  KOA_LAUNCH_CHECK();
  return 0;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int koa_adam_step(const koa_adam_tensor_t* tensors, int n_tensors, const koa_adam_hyper_t* h, void* stream) {
  KOA_REQUIRE(h != nullptr && n_tensors >= 0 && (tensors != nullptr || n_tensors == 0), "null argument");
  KOA_REQUIRE(h->step >= 1, "Adam step counts from 1 (got %d)", h->step);
  KOA_REQUIRE(h->lr >= 0.0 && h->eps >= 0.0 && h->weight_decay >= 0.0 && h->beta1 >= 0.0 && h->beta1 < 1.0 &&
                  h->beta2 >= 0.0 && h->beta2 < 1.0,
              "invalid Adam hyper-parameters");
  const AdamCoef c = koa_arith::make_adam_coef(*h);  // step_arith.cuh (shared with the CPU check)

  AdamTable tab;
  int i = 0;
  while (i < n_tensors) {
    tab.count = 0;
    tab.vec_mask = 0ull;
    tab.chunk_begin[0] = 0;
    long long chunks = 0;
    for (; i < n_tensors && tab.count < kAdamSlots; ++i) {
      const koa_adam_tensor_t& t = tensors[i];
      if (t.numel <= 0 || t.grad == nullptr) continue;  // torch skips parameters without a gradient
      KOA_REQUIRE(t.param && t.exp_avg && t.exp_avg_sq, "tensor %d: null param / state pointer", i);
      const long long nc = (t.numel + kAdamChunk - 1) / kAdamChunk;
      if (tab.count > 0 && chunks + nc > 0x3fffffff) break;  // chunk ids are ints: start another launch
      KOA_REQUIRE(nc <= 0x3fffffff, "tensor %d is too large for one launch", i);
      const int s = tab.count++;
      tab.p[s] = t.param; tab.g[s] = t.grad; tab.m[s] = t.exp_avg; tab.v[s] = t.exp_avg_sq; tab.n[s] = t.numel;
      if (aligned16(t.param) && aligned16(t.grad) && aligned16(t.exp_avg) && aligned16(t.exp_avg_sq))
        tab.vec_mask |= 1ull << s;
      chunks += nc;
      tab.chunk_begin[s + 1] = (int)chunks;
    }
    if (tab.count == 0) continue;
    for (int s = tab.count; s < kAdamSlots; ++s) {  // defined values in the unused slots
      tab.p[s] = nullptr; tab.g[s] = nullptr; tab.m[s] = nullptr; tab.v[s] = nullptr; tab.n[s] = 0;
      tab.chunk_begin[s + 1] = (int)chunks;
    }
    adam_kernel<<<capped_grid(chunks, 8), kThreads, 0, ST>>>(tab, c);
    KOA_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int koa_resample_linear(const void* in, int in_dtype, float* out, int batch, const int* in_dims,
                                   const int* out_dims, const float* scale, const float* shift, void* stream) {
  KOA_REQUIRE(in && out && in_dims && out_dims && batch > 0, "null / empty argument");
  KOA_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift come together");
  long long total = batch;
  for (int d = 0; d < 3; ++d) {
    KOA_REQUIRE(in_dims[d] > 0 && out_dims[d] > 0, "sizes must be positive");
    total *= out_dims[d];
  }
  switch (in_dtype) {
    case KOA_DT_F32: return launch_resample<float>(in, out, total, in_dims, out_dims, scale, shift, ST);
    case KOA_DT_U8: return launch_resample<uint8_t>(in, out, total, in_dims, out_dims, scale, shift, ST);
    case KOA_DT_U16: return launch_resample<uint16_t>(in, out, total, in_dims, out_dims, scale, shift, ST);
    case KOA_DT_I16: return launch_resample<int16_t>(in, out, total, in_dims, out_dims, scale, shift, ST);
    default: koa_set_error("koa_resample_linear: unknown input type %d", in_dtype); return KOA_ERR_ARG;
  }
}

extern "C" int koa_unit_range_affine(const void* in, int in_dtype, int batch, long long n_per, float mean, float stdev,
                                     unsigned int* workspace, float* scale, float* shift, float* minmax, void* stream) {
  KOA_REQUIRE(in && workspace && scale && shift && batch > 0 && n_per > 0, "null / empty argument");
  KOA_REQUIRE(batch <= 65535, "at most 65535 volumes per call");
  minmax_init_kernel<<<koa_cdiv(batch, 128), 128, 0, ST>>>(workspace, batch);
  KOA_LAUNCH_CHECK();
  int rc;
  switch (in_dtype) {
    case KOA_DT_F32: rc = launch_minmax<float>(in, batch, n_per, workspace, ST); break;
    case KOA_DT_U8: rc = launch_minmax<uint8_t>(in, batch, n_per, workspace, ST); break;
    case KOA_DT_U16: rc = launch_minmax<uint16_t>(in, batch, n_per, workspace, ST); break;
    case KOA_DT_I16: rc = launch_minmax<int16_t>(in, batch, n_per, workspace, ST); break;
    default: koa_set_error("koa_unit_range_affine: unknown input type %d", in_dtype); return KOA_ERR_ARG;
  }
  if (rc != 0) return rc;
  unit_range_affine_kernel<<<koa_cdiv(batch, 128), 128, 0, ST>>>(workspace, mean, stdev, scale, shift, minmax, batch);
  KOA_LAUNCH_CHECK();
  return 0;
}

extern "C" int koa_augment_resample(const void* in, int in_dtype, float* out, koa_augment_t* params, int batch,
                                    const int* src_dims, const int* crop_dims, const int* out_dims, float mean, float stdev,
                                    unsigned int* workspace, void* stream) {
  KOA_REQUIRE(in && out && params && src_dims && crop_dims && out_dims && workspace && batch > 0, "null / empty argument");
  KOA_REQUIRE(batch <= 65535, "at most 65535 volumes per call");
  for (int d = 0; d < 3; ++d)
    KOA_REQUIRE(crop_dims[d] > 0 && out_dims[d] > 0 && src_dims[d] >= crop_dims[d], "crop larger than the stored volume");
  switch (in_dtype) {
    case KOA_DT_F32: return launch_augment<float>(in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev, workspace, ST);
    case KOA_DT_U8: return launch_augment<uint8_t>(in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev, workspace, ST);
    case KOA_DT_U16: return launch_augment<uint16_t>(in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev, workspace, ST);
    case KOA_DT_I16: return launch_augment<int16_t>(in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev, workspace, ST);
    default: koa_set_error("koa_augment_resample: unknown input type %d", in_dtype); return KOA_ERR_ARG;
  }
}

extern "C" int koa_predict(const float* logits, float* proba, long long* pred, int batch, int classes, void* stream) {
  KOA_REQUIRE(logits && batch > 0 && classes > 0, "null / empty argument");
  predict_kernel<<<koa_cdiv(batch, 128), 128, 0, ST>>>(logits, proba, pred, batch, classes);
  KOA_LAUNCH_CHECK();
  return 0;
}

extern "C" int koa_ensemble_proba(const float* proba, float* out, long long* pred, int folds, int batch, int classes,
                                  void* stream) {
  KOA_REQUIRE(proba && folds > 0 && batch > 0 && classes > 0, "null / empty argument");
  ensemble_kernel<<<koa_cdiv(batch, 128), 128, 0, ST>>>(proba, out, pred, folds, batch, classes);
  KOA_LAUNCH_CHECK();
  return 0;
}
