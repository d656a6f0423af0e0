// koa_kernels.h — typed launchers of the HBM-bound / CUDA-core kernels (elementwise.cu, stem.cu,
// attention.cu). Internal to libkoa_b200.so; the engines (fe_engine.cu, feat_engine.cu) and the
// per-op C ABI wrappers (ops_api.cu) are built from these.
#pragma once
#include <cuda_runtime.h>

// ---- parameter packing --------------------------------------------------------------------------
int koa_k_pack_conv_w(const float* src, void* dst, int cout, int cin, int fr, int fs, int dgrad_form, cudaStream_t st);
int koa_k_pack_grouped_w(const float* src, void* dst, int c, int cg, int dgrad_form, cudaStream_t st);
int koa_k_unpack_grouped_dw(const float* dense, float* grad, int c, int cg, cudaStream_t st);
int koa_k_unpack_conv_dw(const float* src, float* dst, int cout, int cin, int fr, int fs, cudaStream_t st);
// All convolution weights of one extractor in ONE launch: job j converts src (fp32 [Cout][Cin/g][k][k]) into the
// forward ([Cout][k][k][Cin], or the per-64-channel block-diagonal form of a grouped conv) and, when dgrad != NULL,
// the data-gradient ([Cin][k'][k'][Cout], taps flipped) bf16 operands.
struct KoaPackJob {
  const float* src;
  void* fwd;
  void* dgrad;
  int cout, cin, k, cg;  // cg > 0: grouped 3x3 convolution with cg input channels per group (C = cout = cin)
};
constexpr int kKoaMaxPackJobs = 64;
int koa_k_pack_fe_weights(const KoaPackJob* jobs, int n_jobs, cudaStream_t st);
// dst: 16-bit copy (fp16 when dst_f16, else bf16), dst_t: bf16 transpose (data-gradient operand); either may be NULL
int koa_k_pack_matrix(const float* src, void* dst, void* dst_t, int rows, int cols, cudaStream_t st, int dst_f16 = 0);
int koa_k_cast_bf16(const float* src, void* dst, long long n, cudaStream_t st, int f16 = 0);

// ---- BatchNorm ------------------------------------------------------------------------------------
int koa_k_bn_finalize(const float* sum, const float* sumsq, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, float* scale, float* shift, float* mean, float* invstd, int c, double count,
                      int training, cudaStream_t st);
int koa_k_col_stats(const void* y, float* sum, float* sumsq, long long rows, int c, cudaStream_t st);
// y / res / y2 / out: fp16 forward activations; out_bf16 (optional): bf16 copy of out (the weight-gradient GEMMs of
// the backward pass pair it with bf16 gradients: tcgen05 kind::f16 wants both operands in one format)
int koa_k_bn_act(const void* y, const float* scale, const float* shift, const void* res, const void* y2,
                 const float* scale2, const float* shift2, void* out, void* out_bf16, long long rows, int c, int relu,
                 cudaStream_t st);
// Finalize-in-consumer forms (one launch instead of finalize + apply; C/8 must divide 256, see koa_k_bn_fused_ok):
// the apply kernels derive their per-channel coefficients from the raw sums themselves; the threads that own each
// channel group first also write what later kernels read (mean, invstd, scale, shift / k0..k2), update the running
// statistics (forward, training) and accumulate dgamma / dbeta (backward).
struct KoaBnFwdFin {
  const float* sum; const float* sumsq; const float* gamma; const float* beta;
  float* run_mean; float* run_var; float* scale; float* shift; float* mean; float* invstd;
};
struct KoaBnBwdFin {
  const float* sum_dz; const float* sum_dzx; const float* gamma; const float* mean; const float* invstd;
  float* dgamma; float* dbeta; float* k0; float* k1; float* k2;
};
int koa_k_bn_fused_ok(int c);
// out = act(bn_a(y) [+ res | + bn_b(y2)]); b may be NULL. out_colsum (optional, fp32 [c], caller zeroes): per-channel sums
// of the stored fp16 values of `out` (the `s` of the y-free bottleneck tail, bn_gram.cu)
int koa_k_bn_act_fin(const void* y, const KoaBnFwdFin* a, const void* res, const void* y2, const KoaBnFwdFin* b, void* out,
                     void* out_bf16, float* out_colsum, long long rows, int c, int relu, double count, int training,
                     cudaStream_t st);
// dy = BatchNorm-backward of a (and dy2 of b, sharing dz = dout * (act > 0)); b may be NULL
int koa_k_bn_bwd_apply_fin(const void* dout, const void* act, const void* y, const KoaBnBwdFin* a, void* dy, const void* y2,
                           const KoaBnBwdFin* b, void* dy2, long long rows, int c, double count, int training,
                           cudaStream_t st);
int koa_k_bn_bwd_reduce(const void* dout, const void* act, const void* y, const float* mean, const float* invstd,
                        const void* y2, const float* mean2, const float* invstd2, float* sum_dz, float* sum_dzx,
                        float* sum_dzx2, long long rows, int c, cudaStream_t st);
int koa_k_bn_bwd_finalize(const float* sum_dz, const float* sum_dzx, const float* gamma, const float* mean,
                          const float* invstd, float* dgamma, float* dbeta, float* k0, float* k1, float* k2, int c,
                          double count, int training, cudaStream_t st);
int koa_k_bn_bwd_apply(const void* dout, const void* act, const void* y, const float* k0, const float* k1,
                       const float* k2, void* dy, const void* y2, const float* k0b, const float* k1b, const float* k2b,
                       void* dy2, long long rows, int c, cudaStream_t st);

// ---- y-free bottleneck tail (bn_gram.cu; formulas in its header) -------------------------------------------------
// hi + lo = gram / count as two fp16 [w][w] matrices
int koa_k_gram_split(const float* gram, void* hi, void* lo, int w, double count, cudaStream_t st);
// statistics of y3 = a2 . W3^T from s = colsum(a2) and Q = W3 . Gram / N; w3h: the fp16 forward operand [C][w]
int koa_k_bn_gram_stats(const void* w3h, const float* sa2, const float* q, const float* gamma, const float* beta,
                        float* run_mean, float* run_var, float* scale, float* shift, float* mean, float* invstd, int c_out,
                        int w, double count, cudaStream_t st);
// t = G^T a2 [C][w]; w3m: fp32 master weights [C][w]; wext: bf16 [w][ld_ext] (columns 0..C-1 written), k2w: bf16 [w][C]
int koa_k_bn_gram_bwd(const float* t, const void* w3h, const float* w3m, const float* sa2, const float* q, const float* sdz,
                      const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta, float* dw,
                      void* wext, void* k2w, float* k0, float* k1, float* k2, int c_out, int w, int ld_ext, double count,
                      cudaStream_t st);
// bias[j] = -(k1 . W3)[j]; w3t: bf16 [w][C]
int koa_k_bn_gram_bias(const void* w3t, const float* k1, float* bias, int w, int c_out, cudaStream_t st);

// ---- pooling / resampling -------------------------------------------------------------------------
int koa_k_maxpool_fwd(const void* x, void* out, void* out_bf16, void* idx, int n, int h, int w, int c, cudaStream_t st);
int koa_k_maxpool_bwd(const void* dout, const void* idx, void* dx, int n, int h, int w, int c, cudaStream_t st);
int koa_k_gap_fwd(const void* x, float* feat, int n, int hw, int c, cudaStream_t st);
int koa_k_gap_bwd(const float* dfeat, const void* gate, void* dx, int n, int hw, int c, cudaStream_t st);
int koa_k_zero_insert2(const void* src, void* dst, int n, int h, int w, int c, int ho, int wo, cudaStream_t st);
int koa_k_scatter_add2(const void* src, const void* gate, void* dx, int n, int h, int w, int c, int ho, int wo,
                       cudaStream_t st);

// ---- transformer pieces ---------------------------------------------------------------------------
int koa_k_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32,
                        float* mean, float* rstd, int rows, int d, long long x_row_stride, cudaStream_t st, int out_f16 = 0);
int koa_k_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                        const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int rows, int d,
                        long long x_row_stride, long long dx_row_stride, cudaStream_t st);
int koa_k_token_assemble(const float* emb, const float* cls, const float* pos, float* x, int batch, int n_tok, int n_cls,
                         int d, cudaStream_t st);
int koa_k_token_assemble_bwd(const float* dx, float* dpos, float* dcls, void* demb, int batch, int n_tok, int n_cls, int d,
                             cudaStream_t st);
int koa_k_col_sum(const void* x, int is_bf16, float* out, long long rows, int c, long long ld, cudaStream_t st);
int koa_k_linear_small_fwd(const float* x, const float* w, const float* b, float* y, float* pre, int m, int n, int k,
                           long long x_row_stride, int act, cudaStream_t st);
int koa_k_linear_small_bwd(const float* dy, const float* pre, const float* x, const float* w, float* dpre_scratch,
                           float* dx, float* dw, float* db, int m, int n, int k, long long x_row_stride,
                           long long dx_row_stride, int act, int accumulate_dx, cudaStream_t st);
// ---- dropout (counter-based Philox masks, koa_common.cuh) -------------------------------------------------
// mask (0 or 1/(1-p)) of (seed, site) over n elements, element index = flat index
int koa_k_dropout_mask(unsigned long long seed, unsigned int site, long long n, float p, float* out, cudaStream_t st);
// x *= mask in place (fp32) and/or out_bf16 = bf16(x * mask); n % 4 == 0
int koa_k_dropout_apply(float* x_inplace, const float* x, void* out_bf16, unsigned long long seed, unsigned int site,
                        long long n, float p, cudaStream_t st);
// out = x * mask[img][ch] for tokens x [n_img][positions][c] (nn.Dropout2d: one draw per image and channel)
int koa_k_channel_dropout(const float* x, float* out, long long n_img, int positions, int c, unsigned long long seed,
                          unsigned int site, float p, cudaStream_t st);
int koa_k_focal_loss(const float* logits, const long long* target, float* loss, float* dlogits, int batch, int classes,
                     float gamma, cudaStream_t st);

// ---- attention (attention.cu) -----------------------------------------------------------------------
// qkv: bf16 [B*n][3*D] with feature index = (qkv, head, d); out: bf16 [B*n][D]; probs: fp32 [B][H][n][n].
// f16: qkv and out are fp16 (the transformer's forward format); gradients (dout, dqkv) are always bf16
int koa_k_attention_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                        cudaStream_t st, int f16 = 0);
int koa_k_attention_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                        int head_dim, float scale, cudaStream_t st, int f16 = 0);

// tcgen05 form (attention_tc.cu): one UMMA tile per contraction; n <= 128, head_dim a multiple of 64 up to 256
bool koa_attention_tc_ok(int n, int head_dim);
int koa_k_attention_tc_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                           cudaStream_t st, int f16);
int koa_k_attention_tc_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                           int head_dim, float scale, cudaStream_t st, int f16);

// ---- stem (stem.cu) -----------------------------------------------------------------------------------
// vol (B,1,R,C,S) fp32 slice-innermost -> img [B*S][R][C] fp32
int koa_k_stem_pack(const float* vol, float* img, int batch, int rc, int slices, cudaStream_t st);
// w [64][3][7][7] fp32 -> wfold [49][64] fp32 (sum over the 3 identical input channels)
int koa_k_stem_fold_w(const float* w, float* wfold, cudaStream_t st);
int koa_k_stem_conv_fwd(const float* img, const float* wfold, void* y, float* sum, float* sumsq, int n, int h, int w,
                        cudaStream_t st);
// dwfold [49][64] accumulated with atomics (caller zeroes); then expanded to dw [64][3][7][7]
int koa_k_stem_wgrad(const float* img, const void* dy, float* dwfold, int n, int h, int w, cudaStream_t st);
int koa_k_stem_unfold_dw(const float* dwfold, float* dw, cudaStream_t st);
// tensor-core form: im2col operand A [n*ho*wo][64] bf16 (K = 49 padded to 64), weights Wb [64][64] bf16,
// folded gradient dWb [64][64] fp32 -> dw [64][3][7][7] +=
int koa_k_stem_im2col(const float* img, void* a, int n, int h, int w, int f16, cudaStream_t st);
int koa_k_stem_pack_wb(const float* w, void* wb, cudaStream_t st);
int koa_k_stem_unfold_dwb(const float* dwb, float* dw, cudaStream_t st);
