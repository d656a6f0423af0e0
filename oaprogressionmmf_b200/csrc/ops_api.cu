// ops_api.cu — per-operator C-ABI wrappers (include/koa_b200.h) around the kernels of elementwise.cu,
// stem.cu and attention.cu. The engines (fe_engine.cu / feat_engine.cu) call the typed launchers directly;
// these entry points exist for the small layers the host mirror runs outside the engines (clinical
// embedding, XR1Cnn head, focal loss) and so that every kernel can be parity-tested on its own.
#include "../../include/koa_b200.h"
#include "koa_common.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"

#define ST ((cudaStream_t)stream)

extern "C" int koa_linear_small_fwd(const float* x, const float* w, const float* b, float* y, float* pre, int m, int n,
                                    int k, int act, void* stream) {
  KOA_REQUIRE(x && w && y, "null pointer argument");
  return koa_k_linear_small_fwd(x, w, b, y, pre, m, n, k, k, act, ST);
}
extern "C" int koa_linear_small_bwd(const float* dy, const float* pre, const float* x, const float* w, float* scratch,
                                    float* dx, float* dw, float* db, int m, int n, int k, int act, void* stream) {
  KOA_REQUIRE(dy && x && w && scratch, "null pointer argument");
  return koa_k_linear_small_bwd(dy, pre, x, w, scratch, dx, dw, db, m, n, k, k, k, act, 0, ST);
}
extern "C" int koa_focal_loss(const float* logits, const long long* target, float* loss, float* dlogits, int batch,
                              int classes, float gamma, void* stream) {
  KOA_REQUIRE(logits && target && loss, "null pointer argument");
  return koa_k_focal_loss(logits, target, loss, dlogits, batch, classes, gamma, ST);
}
extern "C" int koa_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32,
                                 float* mean, float* rstd, int rows, int d, void* stream) {
  return koa_k_layernorm_fwd(x, gamma, beta, out_bf16, out_f32, mean, rstd, rows, d, d, ST);
}
extern "C" int koa_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                 float* dx, float* dgamma, float* dbeta, int rows, int d, void* stream) {
  return koa_k_layernorm_bwd(dy, x, gamma, mean, rstd, nullptr, dx, nullptr, dgamma, dbeta, rows, d, d, d, ST);
}
extern "C" int koa_attention_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim,
                                 float scale, void* stream) {
  return koa_k_attention_fwd(qkv, out, probs, batch, n, heads, head_dim, scale, ST);
}
extern "C" int koa_attention_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n,
                                 int heads, int head_dim, float scale, void* stream) {
  return koa_k_attention_bwd(qkv, probs, dout, dqkv, batch, n, heads, head_dim, scale, ST);
}
extern "C" int koa_attention_fwd_fmt(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim,
                                     float scale, int f16, void* stream) {
  return koa_k_attention_fwd(qkv, out, probs, batch, n, heads, head_dim, scale, ST, f16);
}
extern "C" int koa_attention_bwd_fmt(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n,
                                     int heads, int head_dim, float scale, int f16, void* stream) {
  return koa_k_attention_bwd(qkv, probs, dout, dqkv, batch, n, heads, head_dim, scale, ST, f16);
}
extern "C" int koa_stem_pack(const float* vol, float* img, int batch, int rc, int slices, void* stream) {
  return koa_k_stem_pack(vol, img, batch, rc, slices, ST);
}
extern "C" int koa_maxpool_fwd(const void* x, void* out, void* idx, int n, int h, int w, int c, void* stream) {
  return koa_k_maxpool_fwd(x, out, nullptr, idx, n, h, w, c, ST);
}
extern "C" int koa_maxpool_bwd(const void* dout, const void* idx, void* dx, int n, int h, int w, int c, void* stream) {
  return koa_k_maxpool_bwd(dout, idx, dx, n, h, w, c, ST);
}
extern "C" int koa_col_stats(const void* y, float* sum, float* sumsq, long long rows, int c, void* stream) {
  return koa_k_col_stats(y, sum, sumsq, rows, c, ST);
}
extern "C" int koa_channel_dropout(const float* x, float* out, long long n_img, int positions, int c,
                                   unsigned long long seed, unsigned int site, float p, void* stream) {
  KOA_REQUIRE(x != nullptr && out != nullptr, "null pointer argument");
  return koa_k_channel_dropout(x, out, n_img, positions, c, seed, site, p, ST);
}
extern "C" int koa_dropout_mask(unsigned long long seed, unsigned int site, long long rows, int cols, float p, float* out,
                                void* stream) {
  KOA_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "dropout mask needs cols %% 4 == 0");
  return koa_k_dropout_mask(seed, site, rows * cols, p, out, ST);
}
