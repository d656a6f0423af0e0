// feat_engine.cu — whole forward / backward of koafusion's token transformer `FeaT`
// (koafusion/models/_core_trf.py:74-205): patch embedding, CLS + positional embedding, `depth` pre-norm
// blocks (LayerNorm -> attention -> +res, LayerNorm -> Linear/GELU/Linear -> +res, no final norm) and the
// LayerNorm/Linear/GELU/Linear classification head on token 0.
//
// Every Linear is a tcgen05 GEMM (gemm_api.cu) with its bias / GELU / residual fused in the epilogue;
// the residual stream stays fp32. Forward GEMM operands (weights, LayerNorm outputs, q/k/v, attention output, GELU
// output) are fp16 (KOA_FEAT_F16, default 1): everything here sits behind a LayerNorm or a softmax, so the fp16 range is
// not an issue, and the 11-bit significand brings the logits within BASELINE.json's 1e-2 of the fp32 reference where
// bf16 operands left them at 2e-2 (the head's hidden activations alone carried 2^-9 per element). Gradients are bf16;
// the weight-gradient GEMMs convert their fp16 activation operand in shared memory (gemm_wgrad_kernel XCVT).
// LayerNorm, attention (n <= 128 tokens), token assembly and bias gradients are the kernels of elementwise.cu /
// attention.cu.
//
// Parameter / gradient tables (fp32 device pointers, reference state_dict order, _core_trf.py:99-116,189-193):
//   [0] cls_token (NULL when with_cls == 0)  [1] pos_embedding  [2] patch_to_embedding.weight  [3] .bias
//   per layer l, base 4 + 11*l: prenorm_0.{w,b}, attn.to_qkv.weight, attn.to_out.0.{w,b}, prenorm_1.{w,b},
//                               ff.net.0.{w,b}, ff.net.3.{w,b}
//   head, base 4 + 11*depth:    mlp_head0.0.{w,b} (LayerNorm), mlp_head0.1.{w,b}, mlp_head0.4.{w,b}
#include <vector>

#include "../../include/koa_b200.h"
#include "koa_common.cuh"
#include "koa_internal.h"
#include "koa_kernels.h"

#define KOA_TRY(x)          \
  do {                      \
    int rc_ = (x);          \
    if (rc_) return rc_;    \
  } while (0)

namespace {

struct LayerBuf {
  size_t w_qkv, w_qkv_t, w_out, w_out_t, w_ff0, w_ff0_t, w_ff3, w_ff3_t;  // bf16 weights (+ transposes)
  size_t x_in, x_mid;                                                      // fp32 residual stream [M][D]
  size_t ln0, ln1, attn_out, qkv, h_pre, g;                                // bf16 activations
  size_t probs;                                                            // fp32 [B][H][n][n]
  size_t st0, st1;                                                         // LayerNorm mean/rstd [2][M]
};
struct Plan {
  int B, n_p, n, D, depth, heads, mlp, classes, n_cls;
  long long M, Mp;
  size_t tok_bf16, emb, w_pe, w_pe_t, x_final;
  std::vector<LayerBuf> L;
  size_t w_h1, w_h1_t, cls_ln, st_h, hh, hh_pre;
  // backward scratch
  size_t dxa, dxb, dx_bf16, d_ln, dh, dattn, dqkv, d_emb, d_hh, d_scr, d_hpre, d_clsln;
  size_t total;
};
struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  }
};

// forward operand format of the transformer: 1 = fp16 (default), 0 = bf16 (the round-1 arithmetic)
int feat_f16() {
  static const int v = [] {
    const char* e = getenv("KOA_FEAT_F16");
    return e == nullptr ? 1 : (atoi(e) != 0 ? 1 : 0);
  }();
  return v;
}

int build_plan(const koa_feat_desc_t* d, Plan& p) {
  KOA_REQUIRE(d != nullptr, "null descriptor");
  KOA_REQUIRE(d->batch > 0 && d->n_patches > 0 && d->depth >= 0, "bad FeaT geometry");
  KOA_REQUIRE(d->dim % 256 == 0 && d->dim <= 2048, "FeaT width %d must be a multiple of 256 and <= 2048", d->dim);
  KOA_REQUIRE(d->mlp_dim % 64 == 0, "FeaT mlp_dim %d must be a multiple of 64", d->mlp_dim);
  KOA_REQUIRE(d->heads > 0 && d->dim % d->heads == 0 && (d->dim / d->heads) % 8 == 0 && d->dim / d->heads <= 256,
              "unsupported head split %d / %d", d->dim, d->heads);
  KOA_REQUIRE(d->emb_dropout >= 0.0f && d->emb_dropout < 1.0f && d->mlp_dropout >= 0.0f && d->mlp_dropout < 1.0f,
              "dropout probabilities must be in [0, 1)");
  p.B = d->batch; p.n_p = d->n_patches; p.n_cls = d->with_cls ? 1 : 0; p.n = p.n_p + p.n_cls;
  KOA_REQUIRE(p.n <= 128, "FeaT supports at most 128 tokens (got %d)", p.n);
  p.D = d->dim; p.depth = d->depth; p.heads = d->heads; p.mlp = d->mlp_dim; p.classes = d->num_classes;
  p.M = (long long)p.B * p.n; p.Mp = (long long)p.B * p.n_p;
  const size_t D = p.D, mlp = p.mlp, M = p.M, Mp = p.Mp;
  Bump ws;
  p.tok_bf16 = ws.take(Mp * D * 2);
  p.emb = ws.take(Mp * D * 4);
  p.w_pe = ws.take(D * D * 2);
  p.w_pe_t = ws.take(D * D * 2);
  p.L.resize(p.depth);
  for (LayerBuf& l : p.L) {
    l.w_qkv = ws.take(3 * D * D * 2); l.w_qkv_t = ws.take(3 * D * D * 2);
    l.w_out = ws.take(D * D * 2); l.w_out_t = ws.take(D * D * 2);
    l.w_ff0 = ws.take(mlp * D * 2); l.w_ff0_t = ws.take(mlp * D * 2);
    l.w_ff3 = ws.take(mlp * D * 2); l.w_ff3_t = ws.take(mlp * D * 2);
    l.x_in = ws.take(M * D * 4); l.x_mid = ws.take(M * D * 4);
    l.ln0 = ws.take(M * D * 2); l.ln1 = ws.take(M * D * 2); l.attn_out = ws.take(M * D * 2);
    l.qkv = ws.take(M * 3 * D * 2);
    l.h_pre = ws.take(M * mlp * 2); l.g = ws.take(M * mlp * 2);
    l.probs = ws.take((size_t)p.B * p.heads * p.n * p.n * 4);
    l.st0 = ws.take(2 * M * 4); l.st1 = ws.take(2 * M * 4);
  }
  p.x_final = ws.take(M * D * 4);
  p.w_h1 = ws.take(mlp * D * 2); p.w_h1_t = ws.take(mlp * D * 2);
  p.cls_ln = ws.take((size_t)p.B * D * 2);
  p.st_h = ws.take(2 * (size_t)p.B * 4);
  p.hh = ws.take((size_t)p.B * mlp * 4);
  p.hh_pre = ws.take((size_t)p.B * mlp * 2);
  p.dxa = ws.take(M * D * 4); p.dxb = ws.take(M * D * 4);
  p.dx_bf16 = ws.take(M * D * 2);
  p.d_ln = ws.take(M * D * 4);
  p.dh = ws.take(M * mlp * 2);
  p.dattn = ws.take(M * D * 2);
  p.dqkv = ws.take(M * 3 * D * 2);
  p.d_emb = ws.take(Mp * D * 2);
  p.d_hh = ws.take((size_t)p.B * mlp * 4);
  p.d_scr = ws.take((size_t)p.B * std::max<size_t>(mlp, 64) * 4);
  p.d_hpre = ws.take((size_t)p.B * mlp * 2);
  p.d_clsln = ws.take((size_t)p.B * D * 4);
  p.total = ws.off;
  return 0;
}

inline uint8_t* at(void* ws, size_t off) { return reinterpret_cast<uint8_t*>(ws) + off; }
inline float* atf(void* ws, size_t off) { return reinterpret_cast<float*>(at(ws, off)); }

struct Params {
  const void* const* t;
  int depth;
  const float* cls() const { return (const float*)t[0]; }
  const float* pos() const { return (const float*)t[1]; }
  const float* pe_w() const { return (const float*)t[2]; }
  const float* pe_b() const { return (const float*)t[3]; }
  const float* layer(int l, int i) const { return (const float*)t[4 + 11 * l + i]; }
  const float* head(int i) const { return (const float*)t[4 + 11 * depth + i]; }
};
struct Grads {
  void* const* t;
  int depth;
  float* at(int i) const { return (float*)t[i]; }
  float* layer(int l, int i) const { return (float*)t[4 + 11 * l + i]; }
  float* head(int i) const { return (float*)t[4 + 11 * depth + i]; }
};
enum { P_LN0_W = 0, P_LN0_B, P_QKV_W, P_OUT_W, P_OUT_B, P_LN1_W, P_LN1_B, P_FF0_W, P_FF0_B, P_FF3_W, P_FF3_B };
enum { H_LN_W = 0, H_LN_B, H_1_W, H_1_B, H_4_W, H_4_B };

// out = A . W^T (+ fused epilogue)
int linear(const void* a, const void* w, long long m, int n, int k, koa_epilogue_t* ep, cudaStream_t st) {
  ep->ldo = n;
  return koa_gemm_launch(a, w, (int)m, n, k, ep, st);
}

constexpr unsigned kSiteEmb = 0xE000u, kSiteHead = 0xF000u;
// dropout layers inside a transformer block (nn.Dropout of to_out / FeedForward, _core_trf.py:146-149,164)
inline unsigned site_attn_out(int l) { return 4u * l + 0u; }
inline unsigned site_ff_act(int l) { return 4u * l + 1u; }
inline unsigned site_ff_out(int l) { return 4u * l + 2u; }

struct Drop {
  bool emb, mlp;
  float p_emb, p_mlp;
  unsigned long long seed;
  explicit Drop(const koa_feat_desc_t* d)
      : emb(d->training && d->emb_dropout > 0.0f), mlp(d->training && d->mlp_dropout > 0.0f), p_emb(d->emb_dropout),
        p_mlp(d->mlp_dropout), seed(d->seed) {}
  void set(koa_epilogue_t& ep, unsigned site) const {
    if (!mlp) return;
    ep.drop_p = p_mlp; ep.drop_site = site; ep.drop_seed = seed;
  }
};

__global__ void gelu_bwd_rows_kernel(const float* __restrict__ d, const bf16* __restrict__ pre, bf16* __restrict__ out,
                                     long long total, int pre_f16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float h = pre_f16 ? __half2float(reinterpret_cast<const __half*>(pre)[i]) : __bfloat162float(pre[i]);
    out[i] = __float2bfloat16_rn(d[i] * koa::gelu_erf_grad(h));
  }
}

}  // namespace

extern "C" size_t koa_feat_workspace_bytes(const koa_feat_desc_t* d) {
  Plan p;
  if (build_plan(d, p)) return 0;
  return p.total;
}
extern "C" int koa_feat_num_params(const koa_feat_desc_t* d) { return d ? 4 + 11 * d->depth + 6 : -1; }

// Offset of the fp32 attention probabilities [B][H][n][n] of layer `layer` inside the workspace.
extern "C" int koa_feat_probs_offset(const koa_feat_desc_t* d, int layer, size_t* offset, size_t* bytes) {
  Plan p;
  KOA_TRY(build_plan(d, p));
  KOA_REQUIRE(layer >= 0 && layer < p.depth, "layer out of range");
  *offset = p.L[layer].probs;
  *bytes = (size_t)p.B * p.heads * p.n * p.n * 4;
  return 0;
}

extern "C" int koa_feat_forward(const koa_feat_desc_t* d, const void* const* params, const float* tokens, void* ws,
                                float* states_out, float* logits_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Plan p;
  KOA_TRY(build_plan(d, p));
  KOA_REQUIRE(params != nullptr && tokens != nullptr && ws != nullptr, "null pointer argument");
  const Params pr{params, p.depth};
  const int D = p.D, mlp = p.mlp;
  const bool bw = d->need_backward != 0;
  const float scale = 1.0f / sqrtf((float)D);  // reference quirk: model dim, not head dim (_core_trf.py:160)
  const Drop drop(d);
  const int hf = feat_f16();  // forward operands in fp16
  auto fmt = [hf](koa_epilogue_t& e) { e.a_f16 = e.b_f16 = e.out_f16 = hf; };

  KOA_TRY(koa_k_cast_bf16(tokens, at(ws, p.tok_bf16), p.Mp * D, st, hf));
  KOA_TRY(koa_k_pack_matrix(pr.pe_w(), at(ws, p.w_pe), bw ? at(ws, p.w_pe_t) : nullptr, D, D, st, hf));
  {
    koa_epilogue_t ep{};
    fmt(ep);
    ep.out = at(ws, p.emb); ep.out_fp32 = 1; ep.bias = pr.pe_b();
    KOA_TRY(linear(at(ws, p.tok_bf16), at(ws, p.w_pe), p.Mp, D, D, &ep, st));
  }
  float* x0 = atf(ws, p.depth > 0 ? p.L[0].x_in : p.x_final);
  KOA_TRY(koa_k_token_assemble(atf(ws, p.emb), pr.cls(), pr.pos(), x0, p.B, p.n, p.n_cls, D, st));
  if (drop.emb) KOA_TRY(koa_k_dropout_apply(x0, nullptr, nullptr, drop.seed, kSiteEmb, p.M * D, drop.p_emb, st));

  for (int l = 0; l < p.depth; ++l) {
    const LayerBuf& L = p.L[l];
    float* x_in = atf(ws, L.x_in);
    float* x_mid = atf(ws, L.x_mid);
    float* x_next = atf(ws, l + 1 < p.depth ? p.L[l + 1].x_in : p.x_final);
    KOA_TRY(koa_k_pack_matrix(pr.layer(l, P_QKV_W), at(ws, L.w_qkv), bw ? at(ws, L.w_qkv_t) : nullptr, 3 * D, D, st, hf));
    KOA_TRY(koa_k_pack_matrix(pr.layer(l, P_OUT_W), at(ws, L.w_out), bw ? at(ws, L.w_out_t) : nullptr, D, D, st, hf));
    KOA_TRY(koa_k_pack_matrix(pr.layer(l, P_FF0_W), at(ws, L.w_ff0), bw ? at(ws, L.w_ff0_t) : nullptr, mlp, D, st, hf));
    KOA_TRY(koa_k_pack_matrix(pr.layer(l, P_FF3_W), at(ws, L.w_ff3), bw ? at(ws, L.w_ff3_t) : nullptr, D, mlp, st, hf));
    KOA_TRY(koa_k_layernorm_fwd(x_in, pr.layer(l, P_LN0_W), pr.layer(l, P_LN0_B), at(ws, L.ln0), nullptr, atf(ws, L.st0),
                                atf(ws, L.st0) + p.M, (int)p.M, D, D, st, hf));
    {
      koa_epilogue_t ep{};
      fmt(ep);
      ep.out = at(ws, L.qkv);
      KOA_TRY(linear(at(ws, L.ln0), at(ws, L.w_qkv), p.M, 3 * D, D, &ep, st));
    }
    KOA_TRY(koa_k_attention_fwd(at(ws, L.qkv), at(ws, L.attn_out), atf(ws, L.probs), p.B, p.n, p.heads, D / p.heads,
                                scale, st, hf));
    {
      koa_epilogue_t ep{};
      fmt(ep);
      ep.out = x_mid; ep.out_fp32 = 1; ep.bias = pr.layer(l, P_OUT_B); ep.residual_f32 = x_in;
      drop.set(ep, site_attn_out(l));
      KOA_TRY(linear(at(ws, L.attn_out), at(ws, L.w_out), p.M, D, D, &ep, st));
    }
    KOA_TRY(koa_k_layernorm_fwd(x_mid, pr.layer(l, P_LN1_W), pr.layer(l, P_LN1_B), at(ws, L.ln1), nullptr, atf(ws, L.st1),
                                atf(ws, L.st1) + p.M, (int)p.M, D, D, st, hf));
    {
      koa_epilogue_t ep{};
      fmt(ep);
      ep.out = at(ws, L.g); ep.bias = pr.layer(l, P_FF0_B); ep.act = KOA_ACT_GELU; ep.pre_out_bf16 = at(ws, L.h_pre);
      drop.set(ep, site_ff_act(l));
      KOA_TRY(linear(at(ws, L.ln1), at(ws, L.w_ff0), p.M, mlp, D, &ep, st));
    }
    {
      koa_epilogue_t ep{};
      fmt(ep);
      ep.out = x_next; ep.out_fp32 = 1; ep.bias = pr.layer(l, P_FF3_B); ep.residual_f32 = x_mid;
      drop.set(ep, site_ff_out(l));
      KOA_TRY(linear(at(ws, L.g), at(ws, L.w_ff3), p.M, D, mlp, &ep, st));
    }
  }
  if (states_out != nullptr)
    KOA_CHECK_CUDA(cudaMemcpyAsync(states_out, at(ws, p.x_final), (size_t)p.M * D * 4, cudaMemcpyDeviceToDevice, st));

  if (d->compute_head) {
    KOA_REQUIRE(logits_out != nullptr, "compute_head needs logits_out");
    const float* xf = atf(ws, p.x_final);
    KOA_TRY(koa_k_pack_matrix(pr.head(H_1_W), at(ws, p.w_h1), bw ? at(ws, p.w_h1_t) : nullptr, mlp, D, st, hf));
    KOA_TRY(koa_k_layernorm_fwd(xf, pr.head(H_LN_W), pr.head(H_LN_B), at(ws, p.cls_ln), nullptr, atf(ws, p.st_h),
                                atf(ws, p.st_h) + p.B, p.B, D, (long long)p.n * D, st, hf));
    koa_epilogue_t ep{};
    fmt(ep);
    ep.out = at(ws, p.hh); ep.out_fp32 = 1; ep.bias = pr.head(H_1_B); ep.act = KOA_ACT_GELU;
    ep.pre_out_bf16 = at(ws, p.hh_pre);
    drop.set(ep, kSiteHead);
    KOA_TRY(linear(at(ws, p.cls_ln), at(ws, p.w_h1), p.B, mlp, D, &ep, st));
    KOA_TRY(koa_k_linear_small_fwd(atf(ws, p.hh), pr.head(H_4_W), pr.head(H_4_B), logits_out, nullptr, p.B, p.classes, mlp,
                                   mlp, KOA_ACT_NONE, st));
  }
  return 0;
}

extern "C" int koa_feat_backward(const koa_feat_desc_t* d, const void* const* params, void* const* grads, void* ws,
                                 const float* d_states, const float* d_logits, float* d_tokens, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Plan p;
  KOA_TRY(build_plan(d, p));
  KOA_REQUIRE(params != nullptr && grads != nullptr && ws != nullptr, "null pointer argument");
  KOA_REQUIRE(d->need_backward, "forward was not run with need_backward");
  const Params pr{params, p.depth};
  const Grads gr{grads, p.depth};
  const int D = p.D, mlp = p.mlp;
  const long long M = p.M;
  const float scale = 1.0f / sqrtf((float)D);
  const Drop drop(d);
  const int hf = feat_f16();
  const int xf = hf ? 2 : 0;  // weight gradients: the saved forward activation is fp16, converted to bf16 inside the kernel

  float* dx = atf(ws, p.dxa);
  float* dx_other = atf(ws, p.dxb);
  if (d_states != nullptr) KOA_CHECK_CUDA(cudaMemcpyAsync(dx, d_states, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, st));
  else KOA_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)M * D * 4, st));

  if (d->compute_head && d_logits != nullptr) {
    // logits = W4 . gelu(W1 . LN(x_final[:, 0]) + b1) + b4
    KOA_TRY(koa_k_linear_small_bwd(d_logits, nullptr, atf(ws, p.hh), pr.head(H_4_W), atf(ws, p.d_scr), atf(ws, p.d_hh),
                                   gr.head(H_4_W), gr.head(H_4_B), p.B, p.classes, mlp, mlp, mlp, KOA_ACT_NONE, 0, st));
    const long long tot = (long long)p.B * mlp;
    if (drop.mlp) KOA_TRY(koa_k_dropout_apply(atf(ws, p.d_hh), nullptr, nullptr, drop.seed, kSiteHead, tot, drop.p_mlp, st));
    gelu_bwd_rows_kernel<<<koa_cdiv(tot, 256), 256, 0, st>>>(atf(ws, p.d_hh), (const bf16*)at(ws, p.hh_pre),
                                                             (bf16*)at(ws, p.d_hpre), tot, hf);
    KOA_LAUNCH_CHECK();
    KOA_TRY(koa_k_col_sum(at(ws, p.d_hpre), 1, gr.head(H_1_B), p.B, mlp, mlp, st));
    KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.d_hpre), at(ws, p.cls_ln), gr.head(H_1_W), p.B, mlp, D, xf, st));
    koa_epilogue_t ep{};
    ep.out = at(ws, p.d_clsln); ep.out_fp32 = 1;
    KOA_TRY(linear(at(ws, p.d_hpre), at(ws, p.w_h1_t), p.B, D, mlp, &ep, st));
    // LayerNorm backward on token 0 of every sequence, accumulated into dx rows b*n
    KOA_TRY(koa_k_layernorm_bwd(atf(ws, p.d_clsln), atf(ws, p.x_final), pr.head(H_LN_W), atf(ws, p.st_h),
                                atf(ws, p.st_h) + p.B, dx, dx, nullptr, gr.head(H_LN_W), gr.head(H_LN_B), p.B, D,
                                (long long)p.n * D, (long long)p.n * D, st));
  }
  KOA_TRY(koa_k_cast_bf16(dx, at(ws, p.dx_bf16), M * D, st));

  for (int l = p.depth - 1; l >= 0; --l) {
    const LayerBuf& L = p.L[l];
    // ---- x_next = ff3(g) + b3 + x_mid ------------------------------------------------------------
    // p.dx_bf16 = bf16 copy of dx: the gradient w.r.t. (ff3(g) + b3) once the dropout mask of that branch is applied
    if (drop.mlp) {
      KOA_TRY(koa_k_dropout_apply(nullptr, dx, at(ws, p.dx_bf16), drop.seed, site_ff_out(l), M * D, drop.p_mlp, st));
      KOA_TRY(koa_k_col_sum(at(ws, p.dx_bf16), 1, gr.layer(l, P_FF3_B), M, D, D, st));
    } else {
      KOA_TRY(koa_k_col_sum(dx, 0, gr.layer(l, P_FF3_B), M, D, D, st));
    }
    KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.dx_bf16), at(ws, L.g), gr.layer(l, P_FF3_W), (int)M, D, mlp, xf, st));
    {
      koa_epilogue_t ep{};
      ep.out = at(ws, p.dh); ep.act = KOA_ACT_GELU_GRAD; ep.aux_bf16 = at(ws, L.h_pre); ep.act_f16 = hf;
      drop.set(ep, site_ff_act(l));
      KOA_TRY(linear(at(ws, p.dx_bf16), at(ws, L.w_ff3_t), M, mlp, D, &ep, st));
    }
    KOA_TRY(koa_k_col_sum(at(ws, p.dh), 1, gr.layer(l, P_FF0_B), M, mlp, mlp, st));
    KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.dh), at(ws, L.ln1), gr.layer(l, P_FF0_W), (int)M, mlp, D, xf, st));
    {
      koa_epilogue_t ep{};
      ep.out = at(ws, p.d_ln); ep.out_fp32 = 1;
      KOA_TRY(linear(at(ws, p.dh), at(ws, L.w_ff0_t), M, D, mlp, &ep, st));
    }
    KOA_TRY(koa_k_layernorm_bwd(atf(ws, p.d_ln), atf(ws, L.x_mid), pr.layer(l, P_LN1_W), atf(ws, L.st1),
                                atf(ws, L.st1) + M, dx, dx_other, at(ws, p.dx_bf16), gr.layer(l, P_LN1_W),
                                gr.layer(l, P_LN1_B), (int)M, D, D, D, st));
    // ---- x_mid = to_out(attn) + bo + x_in ----------------------------------------------------------
    if (drop.mlp) {
      KOA_TRY(koa_k_dropout_apply(nullptr, dx_other, at(ws, p.dx_bf16), drop.seed, site_attn_out(l), M * D, drop.p_mlp, st));
      KOA_TRY(koa_k_col_sum(at(ws, p.dx_bf16), 1, gr.layer(l, P_OUT_B), M, D, D, st));
    } else {
      KOA_TRY(koa_k_col_sum(dx_other, 0, gr.layer(l, P_OUT_B), M, D, D, st));
    }
    KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.dx_bf16), at(ws, L.attn_out), gr.layer(l, P_OUT_W), (int)M, D, D, xf, st));
    {
      koa_epilogue_t ep{};
      ep.out = at(ws, p.dattn);
      KOA_TRY(linear(at(ws, p.dx_bf16), at(ws, L.w_out_t), M, D, D, &ep, st));
    }
    KOA_TRY(koa_k_attention_bwd(at(ws, L.qkv), atf(ws, L.probs), at(ws, p.dattn), at(ws, p.dqkv), p.B, p.n, p.heads,
                                D / p.heads, scale, st, hf));
    KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.dqkv), at(ws, L.ln0), gr.layer(l, P_QKV_W), (int)M, 3 * D, D, xf, st));
    {
      koa_epilogue_t ep{};
      ep.out = at(ws, p.d_ln); ep.out_fp32 = 1;
      KOA_TRY(linear(at(ws, p.dqkv), at(ws, L.w_qkv_t), M, D, 3 * D, &ep, st));
    }
    KOA_TRY(koa_k_layernorm_bwd(atf(ws, p.d_ln), atf(ws, L.x_in), pr.layer(l, P_LN0_W), atf(ws, L.st0),
                                atf(ws, L.st0) + M, dx_other, dx, at(ws, p.dx_bf16), gr.layer(l, P_LN0_W),
                                gr.layer(l, P_LN0_B), (int)M, D, D, D, st));
  }
  // ---- token assembly + patch embedding ---------------------------------------------------------------
  if (drop.emb) KOA_TRY(koa_k_dropout_apply(dx, nullptr, nullptr, drop.seed, kSiteEmb, M * D, drop.p_emb, st));
  float* dcls = p.n_cls ? gr.at(0) : nullptr;
  KOA_TRY(koa_k_token_assemble_bwd(dx, gr.at(1), dcls, at(ws, p.d_emb), p.B, p.n, p.n_cls, D, st));
  KOA_TRY(koa_k_col_sum(at(ws, p.d_emb), 1, gr.at(3), p.Mp, D, D, st));
  KOA_TRY(koa_gemm_wgrad_launch(at(ws, p.d_emb), at(ws, p.tok_bf16), gr.at(2), (int)p.Mp, D, D, xf, st));
  if (d_tokens != nullptr) {
    koa_epilogue_t ep{};
    ep.out = d_tokens; ep.out_fp32 = 1;
    KOA_TRY(linear(at(ws, p.d_emb), at(ws, p.w_pe_t), p.Mp, D, D, &ep, st));
  }
  return 0;
}
