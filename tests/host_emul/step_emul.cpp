// step_emul.cpp — TEST-ONLY host build of the per-element arithmetic of step_ops.cu (csrc/step_arith.cuh).
//
// g++ compiles the very functions the CUDA kernels call (resample_element, augment_element, crop_view, split_index)
// and loops over the output elements the way the grid-stride loops do; tests/test_kernel_arith_host.py checks the
// result against the golden vectors of the unmodified reference. It verifies indexing, tap order and weights of the
// kernels without a GPU. It is not part of the product: the Python package never loads it and libkoa_b200.so does not
// contain it.
#include <stdint.h>

#include <algorithm>
#include <limits>

#include "../../oaprogressionmmf_b200/csrc/step_arith.cuh"

using namespace koa_arith;

namespace {

template <typename T>
void resample(const T* in, float* out, int batch, const int* di, const int* dout, const float* scale, const float* shift) {
  const Dims3 d_in{di[0], di[1], di[2]}, d_out{dout[0], dout[1], dout[2]};
  const long long in_per = (long long)di[0] * di[1] * di[2];
  const long long total = (long long)batch * dout[0] * dout[1] * dout[2];
  const float rs0 = (float)di[0] / (float)dout[0], rs1 = (float)di[1] / (float)dout[1], rs2 = (float)di[2] / (float)dout[2];
  for (long long i = 0; i < total; ++i) {
    int x0, x1, x2;
    const long long b = split_index(i, d_out, x0, x1, x2);
    float val = resample_element<T>(in + b * in_per, d_in, x0, x1, x2, rs0, rs1, rs2);
    if (scale != nullptr) val = fmaf(val, scale[b], shift[b]);
    out[i] = val;
  }
}

template <typename T>
void augment(const T* in, float* out, koa_augment_t* params, int batch, const int* s, const int* c, const int* o, float mean,
             float stdev) {
  const Dims3 src{s[0], s[1], s[2]}, crop{c[0], c[1], c[2]}, dout{o[0], o[1], o[2]};
  for (int b = 0; b < batch; ++b) {  // crop_minmax_kernel + crop_minmax_finish_kernel
    const CropView<T> v = crop_view(in, (long long)b, params[b], src);
    float lo = std::numeric_limits<float>::infinity(), hi = -lo;
    for (int r = 0; r < c[0]; ++r)
      for (int cc = 0; cc < c[1]; ++cc)
        for (int z = 0; z < c[2]; ++z) {
          const float f = ld_f(v.p + r * v.st_r + cc * v.st_c + z * v.st_s);
          lo = std::min(lo, f);
          hi = std::max(hi, f);
        }
    params[b].lo = lo;
    params[b].range = hi - lo;
  }
  const long long total = (long long)batch * o[0] * o[1] * o[2];
  const float rs0 = (float)c[0] / (float)o[0], rs1 = (float)c[1] / (float)o[1], rs2 = (float)c[2] / (float)o[2];
  for (long long i = 0; i < total; ++i) {
    int x0, x1, x2;
    const long long b = split_index(i, dout, x0, x1, x2);
    const koa_augment_t a = params[b];
    out[i] = augment_element<T>(crop_view(in, b, a, src), a, crop, x0, x1, x2, rs0, rs1, rs2, mean, stdev);
  }
}

}  // namespace

extern "C" int emul_resample_linear(const void* in, int in_dtype, float* out, int batch, const int* in_dims,
                                    const int* out_dims, const float* scale, const float* shift) {
  switch (in_dtype) {
    case KOA_DT_F32: resample((const float*)in, out, batch, in_dims, out_dims, scale, shift); return 0;
    case KOA_DT_U8: resample((const uint8_t*)in, out, batch, in_dims, out_dims, scale, shift); return 0;
    case KOA_DT_U16: resample((const uint16_t*)in, out, batch, in_dims, out_dims, scale, shift); return 0;
    case KOA_DT_I16: resample((const int16_t*)in, out, batch, in_dims, out_dims, scale, shift); return 0;
  }
  return -1;
}

extern "C" int emul_augment_resample(const void* in, int in_dtype, float* out, koa_augment_t* params, int batch,
                                     const int* src_dims, const int* crop_dims, const int* out_dims, float mean,
                                     float stdev) {
  switch (in_dtype) {
    case KOA_DT_F32: augment((const float*)in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev); return 0;
    case KOA_DT_U8: augment((const uint8_t*)in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev); return 0;
    case KOA_DT_U16: augment((const uint16_t*)in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev); return 0;
    case KOA_DT_I16: augment((const int16_t*)in, out, params, batch, src_dims, crop_dims, out_dims, mean, stdev); return 0;
  }
  return -1;
}

// koa_adam_step on host pointers: the same coefficient set-up and per-element update the kernel runs
extern "C" int emul_adam_step(const koa_adam_tensor_t* tensors, int n_tensors, const koa_adam_hyper_t* h) {
  const AdamCoef c = make_adam_coef(*h);
  for (int t = 0; t < n_tensors; ++t) {
    const koa_adam_tensor_t& a = tensors[t];
    if (a.numel <= 0 || a.grad == nullptr) continue;  // torch skips parameters without a gradient
    for (long long i = 0; i < a.numel; ++i) adam_update(a.param[i], a.grad[i], a.exp_avg[i], a.exp_avg_sq[i], c);
  }
  return 0;
}

extern "C" int emul_predict(const float* logits, float* proba, long long* pred, int batch, int classes) {
  for (int b = 0; b < batch; ++b)
    predict_row(logits + (long long)b * classes, classes, proba ? proba + (long long)b * classes : nullptr, pred ? pred + b : nullptr);
  return 0;
}

extern "C" int emul_ensemble_proba(const float* proba, float* out, long long* pred, int folds, int batch, int classes) {
  for (int b = 0; b < batch; ++b)
    ensemble_row(proba, folds, batch, classes, b, out ? out + (long long)b * classes : nullptr, pred ? pred + b : nullptr);
  return 0;
}

// koa_unit_range_affine on host pointers: per-volume minimum / maximum, then the kernel's coefficient formula
template <typename T>
static void unit_range(const T* in, int batch, long long n_per, float mean, float stdev, float* scale, float* shift, float* minmax) {
  for (int b = 0; b < batch; ++b) {
    float lo = std::numeric_limits<float>::infinity(), hi = -lo;
    for (long long i = 0; i < n_per; ++i) {
      const float f = ld_f(in + b * n_per + i);
      lo = std::min(lo, f);
      hi = std::max(hi, f);
    }
    unit_range_coef(lo, hi, mean, stdev, scale + b, shift + b);
    if (minmax != nullptr) { minmax[2 * b] = lo; minmax[2 * b + 1] = hi; }
  }
}

extern "C" int emul_unit_range_affine(const void* in, int in_dtype, int batch, long long n_per, float mean, float stdev,
                                      float* scale, float* shift, float* minmax) {
  switch (in_dtype) {
    case KOA_DT_F32: unit_range((const float*)in, batch, n_per, mean, stdev, scale, shift, minmax); return 0;
    case KOA_DT_U8: unit_range((const uint8_t*)in, batch, n_per, mean, stdev, scale, shift, minmax); return 0;
    case KOA_DT_U16: unit_range((const uint16_t*)in, batch, n_per, mean, stdev, scale, shift, minmax); return 0;
    case KOA_DT_I16: unit_range((const int16_t*)in, batch, n_per, mean, stdev, scale, shift, minmax); return 0;
  }
  return -1;
}
