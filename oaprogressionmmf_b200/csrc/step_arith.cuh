// step_arith.cuh — per-element arithmetic of the Adam, resampling and augmentation kernels of step_ops.cu.
//
// The functions are __host__ __device__ so that the *same source* the kernels run is also compiled by g++ into a
// test-only harness (tests/host_emul/step_emul.cpp) and checked on the CPU against the golden vectors made from the
// unmodified reference: index arithmetic, tap order and weights are verified without a GPU. The product only ever calls
// them from the kernels; nothing in the Python package loads the harness.
#pragma once
#include <math.h>

#include "../../include/koa_b200.h"

#if defined(__CUDACC__)
#define KOA_HD __host__ __device__ __forceinline__
#else
#define KOA_HD inline
#endif

namespace koa_arith {

template <typename T> KOA_HD float ld_f(const T* p) {
#if defined(__CUDA_ARCH__)
  return (float)__ldg(p);
#else
  return (float)*p;
#endif
}

KOA_HD float div_rn(float a, float b) {  // IEEE division whatever the compiler flags
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

KOA_HD float sqrt_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// ---- Adam / AdamW ------------------------------------------------------------------------------------------------------
struct AdamCoef {
  float beta2, one_m_beta1, one_m_beta2, eps, step_size, bc2_sqrt, l2, decay_mul, grad_scale;
};

// scalars exactly as torch.optim.Adam forms them: in double on the host, rounded to fp32 where the tensor op takes them
inline AdamCoef make_adam_coef(const koa_adam_hyper_t& h) {
  const double bc1 = 1.0 - pow(h.beta1, (double)h.step), bc2 = 1.0 - pow(h.beta2, (double)h.step);
  AdamCoef c;
  c.beta2 = (float)h.beta2;
  c.one_m_beta1 = (float)(1.0 - h.beta1);
  c.one_m_beta2 = (float)(1.0 - h.beta2);
  c.eps = (float)h.eps;
  c.step_size = (float)(h.lr / bc1);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.l2 = h.decoupled_weight_decay ? 0.f : (float)h.weight_decay;
  c.decay_mul = h.decoupled_weight_decay ? (float)(1.0 - h.lr * h.weight_decay) : 1.f;
  c.grad_scale = (float)(h.grad_scale == 0.0 ? 1.0 : h.grad_scale);
  return c;
}

KOA_HD void adam_update(float& p, float g, float& m, float& v, const AdamCoef& c) {
  // the order of torch.optim.adam._single_tensor_adam: L2 term into the gradient (Adam) or decay of the parameter
  // (AdamW), exp_avg.lerp_(grad, 1 - beta1), exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2),
  // denom = sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps, param.addcdiv_(exp_avg, denom, -lr / bias_correction1)
  g = fmaf(c.l2, p, g * c.grad_scale);
  p *= c.decay_mul;
  m = fmaf(g - m, c.one_m_beta1, m);
  v = fmaf(c.one_m_beta2 * g, g, v * c.beta2);
  const float denom = div_rn(sqrt_rn(v), c.bc2_sqrt) + c.eps;
  p = fmaf(-c.step_size, div_rn(m, denom), p);
}

// ---- predictions ---------------------------------------------------------------------------------------------------------
// softmax + argmax of one row of logits (koafusion/run/eval_prog_fus.py:300-304); first maximum wins, as torch / numpy
KOA_HD void predict_row(const float* x, int classes, float* proba, long long* pred) {
  float mx = x[0];
  int arg = 0;
  for (int c = 1; c < classes; ++c)
    if (x[c] > mx) { mx = x[c]; arg = c; }
  float sum = 0.f;
  for (int c = 0; c < classes; ++c) sum += expf(x[c] - mx);
  if (proba != nullptr)
    for (int c = 0; c < classes; ++c) proba[c] = expf(x[c] - mx) / sum;
  if (pred != nullptr) *pred = arg;
}

// the fold ensemble of the reference for knee b: softmax over classes of the fold-mean of the per-fold probabilities
// (sic, eval_prog_fus.py:330-336), then argmax; proba is [folds][batch][classes]
KOA_HD void ensemble_row(const float* proba, int folds, int batch, int classes, int b, float* out, long long* pred) {
  const float inv = 1.f / (float)folds;
  float mx = 0.f, sum = 0.f;
  int arg = 0;
  for (int pass = 0; pass < 3; ++pass) {  // maximum, normaliser, output: the fold means are recomputed, not stored
    for (int c = 0; c < classes; ++c) {
      float s = 0.f;
      for (int f = 0; f < folds; ++f) s += proba[((long long)f * batch + b) * classes + c];
      s *= inv;
      if (pass == 0) { if (c == 0 || s > mx) { mx = s; arg = c; } }
      else if (pass == 1) sum += expf(s - mx);
      else if (out != nullptr) out[c] = expf(s - mx) / sum;
    }
  }
  if (pred != nullptr) *pred = arg;
}

// ---- unit range + z-score as one affine map --------------------------------------------------------------------------
// z = ((x - lo) / (hi - lo) - mean) / std = x * scale + shift (PTToUnitRange + PTNormalize, koafusion/preproc/_pt.py:75-124);
// a constant volume gives inf / nan like the reference
KOA_HD void unit_range_coef(float lo, float hi, float mean, float stdev, float* scale, float* shift) {
  const float range = hi - lo;
  *scale = 1.f / (range * stdev);
  *shift = (-lo / range - mean) / stdev;
}

struct Tap { int i0, step; float l0, l1; };
KOA_HD Tap make_tap(int dst, float rscale, int n_in) {
  // area_pixel_compute_source_index: src = scale * (dst + 0.5) - 0.5, clamped at 0 (linear modes)
  float src = rscale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Tap t;
  t.i0 = (int)src;
  if (t.i0 > n_in - 1) t.i0 = n_in - 1;
  t.step = t.i0 < n_in - 1 ? 1 : 0;
  t.l1 = src - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

struct Dims3 { int d0, d1, d2; };

// output element i of (B, D0o, D1o, D2o) -> batch index and the three output coordinates
KOA_HD long long split_index(long long i, const Dims3& o, int& x0, int& x1, int& x2) {
  long long t = i;
  x2 = (int)(t % o.d2); t /= o.d2;
  x1 = (int)(t % o.d1); t /= o.d1;
  x0 = (int)(t % o.d0);
  return t / o.d0;
}

// one output element of the linear resampling (F.interpolate, align_corners=False, scales recomputed from the sizes)
template <typename T>
KOA_HD float resample_element(const T* vol, const Dims3& in, int x0, int x1, int x2, float rs0, float rs1, float rs2) {
  const Tap a = make_tap(x0, rs0, in.d0), h = make_tap(x1, rs1, in.d1), w = make_tap(x2, rs2, in.d2);
  const T* p00 = vol + ((long long)a.i0 * in.d1 + h.i0) * in.d2 + w.i0;
  const T* p01 = p00 + (long long)h.step * in.d2;
  const T* p10 = p00 + (long long)a.step * in.d1 * in.d2;
  const T* p11 = p10 + (long long)h.step * in.d2;
  // nesting of upsample_trilinear3d: depth outermost, width innermost
  const float v00 = w.l0 * ld_f(p00) + w.l1 * ld_f(p00 + w.step);
  const float v01 = w.l0 * ld_f(p01) + w.l1 * ld_f(p01 + w.step);
  const float v10 = w.l0 * ld_f(p10) + w.l1 * ld_f(p10 + w.step);
  const float v11 = w.l0 * ld_f(p11) + w.l1 * ld_f(p11 + w.step);
  return a.l0 * (h.l0 * v00 + h.l1 * v01) + a.l1 * (h.l0 * v10 + h.l1 * v11);
}

// The crop of volume b as a strided view of the stored batch: p is crop voxel (0, 0, 0), st_* the element strides along
// rows, columns and slices. A mirrored volume (koa_augment_t::flip, RIGHT knees) is the same view with a negative stride:
// mirrored column off1 + c is stored column C - 1 - off1 - c.
template <typename T> struct CropView { const T* p; long long st_r, st_c, st_s; };
template <typename T>
KOA_HD CropView<T> crop_view(const T* in, long long b, const koa_augment_t& a, const Dims3& src) {
  const long long st_r = (long long)src.d1 * src.d2, st_c = src.d2;
  CropView<T> v;
  v.p = in + b * (long long)src.d0 * st_r + a.off0 * st_r;
  v.st_r = st_r;
  if (a.flip == 1) { v.p += (long long)(src.d1 - 1 - a.off1) * st_c; v.st_c = -st_c; }
  else { v.p += a.off1 * st_c; v.st_c = st_c; }
  if (a.flip == 2) { v.p += src.d2 - 1 - a.off2; v.st_s = -1; }
  else { v.p += a.off2; v.st_s = 1; }
  return v;
}

// unit-range value of the crop at (yy, xx, s); zero padding of F.grid_sample outside the crop
template <typename T>
KOA_HD float unit_tap(const CropView<T>& v, const koa_augment_t& a, int yy, int xx, int s, int R, int C) {
  if (yy < 0 || yy >= R || xx < 0 || xx >= C) return 0.f;
  return div_rn(ld_f(v.p + yy * v.st_r + xx * v.st_c + s * v.st_s) - a.lo, a.range);
}

// value of the augmented full-resolution crop at (r, c, s), crop coordinates
template <typename T>
KOA_HD float aug_voxel(const CropView<T>& v, const koa_augment_t& a, int r, int c, int s, int R, int C, float mean,
                       float stdev) {
  float u;
  if (a.rotate) {
    // F.affine_grid([[cos, -sin, 0], [sin, cos, 0]]) + F.grid_sample on the (S, CH, R, C) view, align_corners=False:
    // H = R, W = C (koafusion/preproc/_pt.py:257-358)
    const float x = (2.f * (float)c + 1.f) / (float)C - 1.f;
    const float y = (2.f * (float)r + 1.f) / (float)R - 1.f;
    const float gx = a.cos_t * x - a.sin_t * y;
    const float gy = a.sin_t * x + a.cos_t * y;
    const float ix = ((gx + 1.f) * (float)C - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * (float)R - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    u = unit_tap(v, a, y0, x0, s, R, C) * (wx0 * wy0) + unit_tap(v, a, y0, x0 + 1, s, R, C) * (wx1 * wy0) +
        unit_tap(v, a, y0 + 1, x0, s, R, C) * (wx0 * wy1) + unit_tap(v, a, y0 + 1, x0 + 1, s, R, C) * (wx1 * wy1);
  } else {
    u = div_rn(ld_f(v.p + r * v.st_r + c * v.st_c + s * v.st_s) - a.lo, a.range);
  }
  if (a.inv_gamma != 0.f) u = powf(u, a.inv_gamma);
  return div_rn(u - mean, stdev);
}

// one output element of mirror -> crop -> unit range -> rotation -> gamma -> z-score -> linear resampling
template <typename T>
KOA_HD float augment_element(const CropView<T>& v, const koa_augment_t& a, const Dims3& crop, int x0, int x1, int x2,
                             float rs0, float rs1, float rs2, float mean, float stdev) {
  const Tap ta = make_tap(x0, rs0, crop.d0), th = make_tap(x1, rs1, crop.d1), tw = make_tap(x2, rs2, crop.d2);
  float acc = 0.f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 8; ++k) {
    const float wgt = ((k & 4) ? ta.l1 : ta.l0) * ((k & 2) ? th.l1 : th.l0) * ((k & 1) ? tw.l1 : tw.l0);
    if (wgt == 0.f) continue;  // factor 1 along an axis, last index of an axis
    const int r = ta.i0 + ((k & 4) ? ta.step : 0), c = th.i0 + ((k & 2) ? th.step : 0), s = tw.i0 + ((k & 1) ? tw.step : 0);
    acc = fmaf(wgt, aug_voxel<T>(v, a, r, c, s, crop.d0, crop.d1, mean, stdev), acc);
  }
  return acc;
}

}  // namespace koa_arith
