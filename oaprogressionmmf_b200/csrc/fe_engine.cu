// fe_engine.cu — whole-network forward / backward of the per-slice CNN feature extractors
// (ResNet-18/34/50, ResNeXt-50-32x4d without `fc`; koafusion/models/_torchvision.py:227-239 and the
// torchvision twins picked by koafusion/models/_core_fes.py:6-15).
//
// One C call runs every layer of the extractor on one stream:
//   stem pack -> 7x7 stem (im2col + tensor-core GEMM, K = 49 folded taps) -> BN/ReLU -> max-pool -> [bottleneck | basic] blocks -> GAP
// Convolutions are tcgen05 implicit GEMMs (gemm_api.cu) that also emit the BatchNorm batch statistics;
// BN-apply/ReLU/residual, pooling and the BN backward are the HBM-bound kernels of elementwise.cu.
// Activations are NHWC bf16; everything needed by backward stays in the caller-provided workspace whose
// layout is a pure function of the descriptor (koa_fe_workspace_bytes / koa_fe_debug_offset).
#include <algorithm>
#include <map>
#include <mutex>
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
// Operands of the weight gradients. dY is bf16 and one tcgen05 MMA takes one 16-bit format, so the fp16 forward
// activations either get a bf16 copy (written by the BatchNorm-apply pass: 2 of its 6-8 bytes per element) or are
// converted to bf16 in shared memory by the weight-gradient kernel (gemm_wgrad_kernel XCVT, x_f16 = 2), which doubles the
// shared-memory traffic of a kernel whose 128 x 128 tiles are already shared-memory bound (494 -> 425 TFLOP/s).
// KOA_WGRAD_XCVT: 0 = copies everywhere; 1 = no copies (-25 % workspace, same step time as 0); 2 (default) = no copies
// where the consuming weight gradient is HBM-bound anyway and the conversion is free: the pooled stem output and the
// outputs of the bottleneck blocks with at most 512 channels (layer1 / layer2: consumed by 1x1 convolutions with
// 50-170 FLOP per byte), copies for the rest.
int wgrad_xcvt_mode() {
  static const int v = [] {
    const char* e = getenv("KOA_WGRAD_XCVT");
    return e == nullptr ? 2 : atoi(e);
  }();
  return v;
}
// KOA_BN_GRAM (default 1): train-mode bottlenecks never store the output of their last 1x1 convolution; its BatchNorm runs
// from the Gram matrix of the convolution's input (bn_gram.cu). 0: the round-1 dataflow (y3 stored, separate
// BatchNorm apply / backward-apply passes over the widest tensor of the block).
bool bn_gram_enabled() {
  static const int v = [] {
    const char* e = getenv("KOA_BN_GRAM");
    return e == nullptr ? 1 : atoi(e);
  }();
  return v != 0;
}
// KOA_EVAL_FUSED (default 1): inference (eval mode, no backward) applies every BatchNorm (+ residual + ReLU) in the epilogue
// of its convolution (gemm_conv.cuh MODE 2): no raw convolution output is stored and no BatchNorm pass runs.
bool eval_fused_enabled() {
  static const int v = [] {
    const char* e = getenv("KOA_EVAL_FUSED");
    return e == nullptr ? 1 : atoi(e);
  }();
  return v != 0;
}
}  // namespace

namespace {

struct Unit {
  int cin, cout, k, stride, pad, groups;
  int hin, win, hout, wout;
  long long rows_in, rows_out;
  size_t w_fwd, w_dgrad, dw_scratch, y;
  size_t fstat;  // forward batch statistics: sum, sumsq            (zeroed by forward)
  size_t bstat;  // backward reductions: sum_dz, sum_dz_xhat         (zeroed by backward)
  size_t coef;   // scale, shift, mean, invstd, k0, k1, k2
  int idx;  // position in the params (x5) / grads (x3) tables
};
struct Block {
  int kind;  // 0 bottleneck, 1 basic
  int u1, u2, u3, ud;
  size_t a1, a2, out;              // fp16 ReLU outputs (operands of the next forward GEMM, gates, mask sources)
  size_t a1_bf, a2_bf, out_bf;     // bf16 copies, kept only when backward will run: operands of the weight gradients
  size_t in, in_bf;                // activation feeding the block
  int stride;
  // y-free tail (Plan::gram): s = colsum(a2) [w], Gram = a2^T a2 [w][w] (both fp32, zeroed with the forward statistics),
  // fp16 head / tail of Gram / N, Q = W3 . Gram / N [C][w] fp32 (kept for backward), T = G^T a2 [C][w] fp32 (zeroed with the
  // backward statistics), wext = [k0 * W3^T | -W3^T diag(k2) W3] bf16 [w][C + w], k2w = -k2 * W3^T bf16 [w][C], bias fp32 [w]
  size_t sa2, gram, gram_hi, gram_lo, q, tq, wext, k2w, cbias;
};
struct Plan {
  int n_img, h, w, out_c, out_h, out_w;
  std::vector<Unit> units;
  std::vector<Block> blocks;
  size_t img, wstem, dwstem, a_stem, a0, p0, p0_bf, idx0;
  size_t fstat_begin, fstat_end;   // contiguous fp32 regions zeroed at the start of forward / backward
  size_t bstat_begin, bstat_end;
  size_t g[2], t[5];               // backward scratch (block gradients ping-pong + temporaries)
  size_t total;
  bool gram;                       // train-mode bottlenecks without y3 (bn_gram.cu)
  bool fused_eval;                 // inference: BatchNorm in the convolution epilogues, no y at all
};

enum { S_SUM = 0, S_SUMSQ, S_SDZ, S_SDZX, S_SCALE, S_SHIFT, S_MEAN, S_INVSTD, S_K0, S_K1, S_K2 };
constexpr int COEF_SLOTS = 7;

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  }
};

int build_plan(const koa_fe_desc_t* d, Plan& p) {
  KOA_REQUIRE(d != nullptr, "null descriptor");
  KOA_REQUIRE(d->arch >= 0 && d->arch <= 3, "unknown feature-extractor arch %d", d->arch);
  KOA_REQUIRE(d->n_img > 0 && d->h >= 32 && d->w >= 32, "bad input geometry n=%d h=%d w=%d", d->n_img, d->h, d->w);
  const bool bottleneck = d->arch >= 2;
  const int layers_r18[4] = {2, 2, 2, 2}, layers_r50[4] = {3, 4, 6, 3};
  const int* layers = d->arch == 0 ? layers_r18 : layers_r50;
  const int groups = d->arch == 3 ? 32 : 1;
  const int width_per_group = d->arch == 3 ? 4 : 64;
  const int expansion = bottleneck ? 4 : 1;
  p.n_img = d->n_img; p.h = d->h; p.w = d->w;
  p.units.clear(); p.blocks.clear();
  p.gram = bottleneck && d->training != 0 && bn_gram_enabled();
  p.fused_eval = d->training == 0 && d->need_backward == 0 && eval_fused_enabled();
  Bump ws;
  const long long n = d->n_img;
  auto conv_out = [](int x, int k, int s, int pad) { return (x + 2 * pad - k) / s + 1; };

  auto add_unit = [&](int cin, int cout, int k, int stride, int pad, int grp, int hin, int win) {
    Unit u{};
    u.cin = cin; u.cout = cout; u.k = k; u.stride = stride; u.pad = pad; u.groups = grp;
    u.hin = hin; u.win = win; u.hout = conv_out(hin, k, stride, pad); u.wout = conv_out(win, k, stride, pad);
    u.rows_in = n * hin * win; u.rows_out = n * u.hout * u.wout;
    u.idx = (int)p.units.size();
    p.units.push_back(u);
    return u.idx;
  };

  // ---- topology -------------------------------------------------------------------------------
  const int stem = add_unit(3, 64, 7, 2, 3, 1, d->h, d->w);
  int ch = p.units[stem].hout, cw = p.units[stem].wout;
  const int ph = conv_out(ch, 3, 2, 1), pw = conv_out(cw, 3, 2, 1);
  int inplanes = 64, hh = ph, ww = pw;
  for (int li = 0; li < 4; ++li) {
    const int planes = 64 << li;
    for (int bi = 0; bi < layers[li]; ++bi) {
      const int stride = (li > 0 && bi == 0) ? 2 : 1;
      Block b{};
      b.kind = bottleneck ? 0 : 1;
      b.stride = stride;
      b.u3 = -1; b.ud = -1;
      const int outc = planes * expansion;
      if (bottleneck) {
        const int width = (planes * width_per_group / 64) * groups;
        b.u1 = add_unit(inplanes, width, 1, 1, 0, 1, hh, ww);
        b.u2 = add_unit(width, width, 3, stride, 1, groups, hh, ww);
        b.u3 = add_unit(width, outc, 1, 1, 0, 1, p.units[b.u2].hout, p.units[b.u2].wout);
      } else {
        b.u1 = add_unit(inplanes, planes, 3, stride, 1, 1, hh, ww);
        b.u2 = add_unit(planes, planes, 3, 1, 1, 1, p.units[b.u1].hout, p.units[b.u1].wout);
      }
      if (bi == 0 && (stride != 1 || inplanes != outc)) b.ud = add_unit(inplanes, outc, 1, stride, 0, 1, hh, ww);
      p.blocks.push_back(b);
      inplanes = outc;
      hh = p.units[b.u2].hout; ww = p.units[b.u2].wout;
    }
  }
  p.out_c = inplanes; p.out_h = hh; p.out_w = ww;
  for (const Unit& u : p.units) {
    if (u.idx == stem) continue;
    KOA_REQUIRE(u.cin % 64 == 0 && u.cout % 64 == 0, "unit %d: channels must be multiples of 64", u.idx);
    if (u.groups > 1) KOA_REQUIRE(u.cin == u.cout && 64 % (u.cin / u.groups) == 0, "unit %d: unsupported grouping", u.idx);
    if (u.stride == 2 && u.k == 3) KOA_REQUIRE(u.hin % 2 == 0 && u.win % 2 == 0, "stride-2 3x3 conv needs even input size (got %dx%d)", u.hin, u.win);
  }

  // ---- workspace --------------------------------------------------------------------------------
  p.img = ws.take((size_t)n * d->h * d->w * 4);
  p.wstem = ws.take(64 * 64 * 2);
  for (Unit& u : p.units) {
    if (u.idx == stem) continue;
    const size_t wcount = u.groups > 1 ? (size_t)u.cout * 9 * 64 : (size_t)u.cout * u.k * u.k * u.cin;
    u.w_fwd = ws.take(wcount * 2);
    u.w_dgrad = ws.take(wcount * 2);
    u.dw_scratch = (u.k > 1) ? ws.take(wcount * 4) : 0;
  }
  for (const Block& b : p.blocks)  // the coefficient kernels of the y-free tail want power-of-two widths <= 1024
    if (p.gram && !(koa_k_bn_fused_ok(p.units[b.u2].cout) && p.units[b.u2].cout <= 1024)) p.gram = false;
  {
    std::vector<char> no_y(p.units.size(), p.fused_eval ? 1 : 0);
    for (const Block& b : p.blocks) {
      if (p.gram) no_y[b.u3] = 1;
      if (p.fused_eval && b.ud >= 0) no_y[b.ud] = 0;  // holds bn_d(conv_d(x)), the residual of the block's last convolution
    }
    for (Unit& u : p.units) u.y = no_y[u.idx] ? 0 : ws.take((size_t)u.rows_out * u.cout * 2);
  }
  const Unit& us = p.units[stem];
  p.a_stem = ws.take((size_t)us.rows_out * 64 * 2);  // im2col operand of the stem GEMM (kept for its weight gradient)
  p.a0 = ws.take((size_t)us.rows_out * 64 * 2);
  const bool gram_bw = p.gram && d->need_backward != 0;
  const bool bw = d->need_backward != 0;
  // bf16 copies of the activations (operands of the weight gradients): see wgrad_xcvt_mode()
  const int xm = wgrad_xcvt_mode();
  const bool bwc = bw && xm != 1;                 // copies of the narrow a1 / a2 tensors
  const bool bwc_p0 = bwc && !(xm == 2 && bottleneck);
  p.p0 = ws.take((size_t)n * ph * pw * 64 * 2);
  p.p0_bf = bwc_p0 ? ws.take((size_t)n * ph * pw * 64 * 2) : 0;
  p.idx0 = ws.take((size_t)n * ph * pw * 64);
  size_t prev = p.p0, prev_bf = p.p0_bf;
  size_t max_act = (size_t)us.rows_out * 64;
  for (Block& b : p.blocks) {
    const Unit& u1 = p.units[b.u1];
    const Unit& u2 = p.units[b.u2];
    b.in = prev;
    b.in_bf = prev_bf;
    b.a1 = ws.take((size_t)u1.rows_out * u1.cout * 2);
    b.a1_bf = bwc ? ws.take((size_t)u1.rows_out * u1.cout * 2) : 0;
    if (b.kind == 0) {
      b.a2 = ws.take((size_t)u2.rows_out * u2.cout * 2);
      // (the y-free tail reads a2 in bf16 as the second K segment of its data-gradient GEMM: always a copy)
      b.a2_bf = (bwc || gram_bw) ? ws.take((size_t)u2.rows_out * u2.cout * 2) : 0;
      const Unit& u3 = p.units[b.u3];
      b.out = ws.take((size_t)u3.rows_out * u3.cout * 2);
      const bool wide_hbm_bound = xm == 2 && u3.cout <= 512;  // consumers: 1x1 conv1 / downsample of the next block
      b.out_bf = (bwc && !wide_hbm_bound) ? ws.take((size_t)u3.rows_out * u3.cout * 2) : 0;
      max_act = std::max(max_act, (size_t)u3.rows_out * u3.cout);
    } else {
      b.a2 = 0; b.a2_bf = 0;
      b.out = ws.take((size_t)u2.rows_out * u2.cout * 2);
      b.out_bf = bwc ? ws.take((size_t)u2.rows_out * u2.cout * 2) : 0;
    }
    max_act = std::max(max_act, (size_t)u1.rows_in * u1.cin);
    max_act = std::max(max_act, (size_t)u1.rows_in * u1.cout);  // zero-inserted / pre-stride tensors
    max_act = std::max(max_act, (size_t)u2.rows_in * u2.cout);
    prev = b.out;
    prev_bf = b.out_bf;
  }
  p.fstat_begin = ws.off;
  for (Unit& u : p.units) u.fstat = ws.take((size_t)2 * u.cout * 4);
  if (p.gram)
    for (Block& b : p.blocks) {
      const size_t w = (size_t)p.units[b.u2].cout;
      b.sa2 = ws.take(w * 4);
      b.gram = ws.take(w * w * 4);
    }
  p.fstat_end = ws.off;
  p.bstat_begin = ws.off;
  p.dwstem = ws.take(64 * 64 * 4);
  for (Unit& u : p.units) u.bstat = ws.take((size_t)2 * u.cout * 4);
  if (gram_bw)
    for (Block& b : p.blocks) b.tq = ws.take((size_t)p.units[b.u3].cout * p.units[b.u2].cout * 4);
  p.bstat_end = ws.off;
  if (p.gram)
    for (Block& b : p.blocks) {
      const size_t w = (size_t)p.units[b.u2].cout, c = (size_t)p.units[b.u3].cout;
      b.gram_hi = ws.take(w * w * 2);
      b.gram_lo = ws.take(w * w * 2);
      b.q = ws.take(c * w * 4);
      if (gram_bw) {
        b.wext = ws.take(w * (c + w) * 2);
        b.k2w = ws.take(w * c * 2);
        b.cbias = ws.take(w * 4);
      }
    }
  for (Unit& u : p.units) u.coef = ws.take((size_t)COEF_SLOTS * u.cout * 4);
  for (int i = 0; i < 2; ++i) p.g[i] = ws.take(max_act * 2);
  for (int i = 0; i < 5; ++i) p.t[i] = ws.take(max_act * 2);
  p.total = ws.off;
  return 0;
}

inline uint8_t* at(void* ws, size_t off) { return reinterpret_cast<uint8_t*>(ws) + off; }
inline float* bn_slot(void* ws, const Unit& u, int slot) {
  if (slot <= S_SUMSQ) return reinterpret_cast<float*>(at(ws, u.fstat)) + (size_t)slot * u.cout;
  if (slot <= S_SDZX) return reinterpret_cast<float*>(at(ws, u.bstat)) + (size_t)(slot - S_SDZ) * u.cout;
  return reinterpret_cast<float*>(at(ws, u.coef)) + (size_t)(slot - S_SCALE) * u.cout;
}

struct ParamView {
  const void* const* params;
  const float* w(const Unit& u) const { return (const float*)params[u.idx * 5 + 0]; }
  const float* gamma(const Unit& u) const { return (const float*)params[u.idx * 5 + 1]; }
  const float* beta(const Unit& u) const { return (const float*)params[u.idx * 5 + 2]; }
  float* run_mean(const Unit& u) const { return (float*)params[u.idx * 5 + 3]; }
  float* run_var(const Unit& u) const { return (float*)params[u.idx * 5 + 4]; }
};

// bf16 GEMM operands of every convolution but the stem, forward and (when backward will run) data-gradient form,
// packed from the fp32 master weights by one launch (ResNet-50: 52 jobs fit one kernel-parameter table).
int pack_all_weights(const Plan& p, const ParamView& pv, void* ws, bool need_dgrad, cudaStream_t st) {
  std::vector<KoaPackJob> jobs;
  for (const Unit& u : p.units) {
    if (u.idx == 0) continue;  // stem: folded separately
    KoaPackJob j{};
    j.src = pv.w(u);
    j.fwd = at(ws, u.w_fwd);
    j.dgrad = need_dgrad ? at(ws, u.w_dgrad) : nullptr;
    j.cout = u.cout; j.cin = u.cin; j.k = u.k;
    j.cg = u.groups > 1 ? u.cin / u.groups : 0;
    jobs.push_back(j);
    if ((int)jobs.size() == kKoaMaxPackJobs) {
      KOA_TRY(koa_k_pack_fe_weights(jobs.data(), (int)jobs.size(), st));
      jobs.clear();
    }
  }
  if (!jobs.empty()) KOA_TRY(koa_k_pack_fe_weights(jobs.data(), (int)jobs.size(), st));
  return 0;
}

// y = conv(x) (+ batch statistics when training)
int conv_forward(const Plan& p, const Unit& u, const void* x, void* ws, int training, cudaStream_t st) {
  koa_epilogue_t ep{};
  ep.out = at(ws, u.y);
  ep.ldo = u.cout;
  ep.a_f16 = ep.b_f16 = ep.out_f16 = 1;  // forward activations and weights of the CNN are fp16
  if (training) {
    ep.col_sum = bn_slot(ws, u, S_SUM);
    ep.col_sumsq = bn_slot(ws, u, S_SUMSQ);
  }
  if (u.groups > 1) {
    const KoaFlopScale fs((double)(u.cin / u.groups) / 64.0);  // block-diagonal chunks: cin / groups of 64 columns are real
    return koa_conv_grouped_launch(x, at(ws, u.w_fwd), p.n_img, u.hin, u.win, u.cin, u.stride, &ep, st);
  }
  if (u.k == 1 && u.stride == 1) return koa_gemm_launch(x, at(ws, u.w_fwd), (int)u.rows_out, u.cout, u.cin, &ep, st);
  return koa_conv_fprop_launch(x, at(ws, u.w_fwd), p.n_img, u.hin, u.win, u.cin, u.cout, u.k, u.k, u.stride, u.pad, &ep, st);
}

// out = act(bn(conv(x)) [+ res [* rscale + rshift]]) in ONE kernel: the BatchNorm coefficients are already known (eval mode, or
// the y-free tail in train mode) and applied by the GEMM epilogue (gemm_conv.cuh MODE 2); y never exists in HBM.
int conv_bn_forward(const Plan& p, const Unit& u, const void* x, void* ws, const void* res, const Unit* res_bn, int relu,
                    void* out, void* out_bf, cudaStream_t st) {
  koa_epilogue_t ep{};
  ep.out = out;
  ep.ldo = u.cout;
  ep.a_f16 = ep.b_f16 = ep.out_f16 = ep.act_f16 = 1;
  ep.bn_scale = bn_slot(ws, u, S_SCALE);
  ep.bn_shift = bn_slot(ws, u, S_SHIFT);
  ep.add_bf16 = res;
  if (res_bn != nullptr) {
    ep.res_scale = bn_slot(ws, *res_bn, S_SCALE);
    ep.res_shift = bn_slot(ws, *res_bn, S_SHIFT);
  }
  ep.act = relu ? KOA_ACT_RELU : KOA_ACT_NONE;
  ep.out_bf16_copy = out_bf;
  if (u.groups > 1) {
    const KoaFlopScale fs((double)(u.cin / u.groups) / 64.0);  // block-diagonal chunks: cin / groups of 64 columns are real
    return koa_conv_grouped_launch(x, at(ws, u.w_fwd), p.n_img, u.hin, u.win, u.cin, u.stride, &ep, st);
  }
  if (u.k == 1 && u.stride == 1) return koa_gemm_launch(x, at(ws, u.w_fwd), (int)u.rows_out, u.cout, u.cin, &ep, st);
  return koa_conv_fprop_launch(x, at(ws, u.w_fwd), p.n_img, u.hin, u.win, u.cin, u.cout, u.k, u.k, u.stride, u.pad, &ep, st);
}

int bn_finalize(const Unit& u, const ParamView& pv, void* ws, int training, cudaStream_t st) {
  return koa_k_bn_finalize(bn_slot(ws, u, S_SUM), bn_slot(ws, u, S_SUMSQ), pv.gamma(u), pv.beta(u), pv.run_mean(u),
                           pv.run_var(u), bn_slot(ws, u, S_SCALE), bn_slot(ws, u, S_SHIFT), bn_slot(ws, u, S_MEAN),
                           bn_slot(ws, u, S_INVSTD), u.cout, (double)u.rows_out, training, st);
}

KoaBnFwdFin fwd_fin(const Unit& u, const ParamView& pv, void* ws) {
  return KoaBnFwdFin{bn_slot(ws, u, S_SUM), bn_slot(ws, u, S_SUMSQ), pv.gamma(u), pv.beta(u), pv.run_mean(u), pv.run_var(u),
                     bn_slot(ws, u, S_SCALE), bn_slot(ws, u, S_SHIFT), bn_slot(ws, u, S_MEAN), bn_slot(ws, u, S_INVSTD)};
}

// out = relu(bn(y) [+ res | + bn_d(y_d)]) with the BatchNorm coefficients derived inside the apply kernel (one launch
// per BatchNorm instead of finalize + apply) whenever the channel count allows it.
int bn_apply(const Unit& u, const Unit* ud, const ParamView& pv, void* ws, const void* res, void* out, void* out_bf,
             int training, cudaStream_t st, float* out_colsum = nullptr) {
  if (koa_k_bn_fused_ok(u.cout)) {
    const KoaBnFwdFin fa = fwd_fin(u, pv, ws);
    KoaBnFwdFin fb{};
    if (ud) fb = fwd_fin(*ud, pv, ws);
    return koa_k_bn_act_fin(at(ws, u.y), &fa, res, ud ? at(ws, ud->y) : nullptr, ud ? &fb : nullptr, out, out_bf, out_colsum,
                            u.rows_out, u.cout, 1, (double)u.rows_out, training, st);
  }
  KOA_REQUIRE(out_colsum == nullptr, "column sums need the fused BatchNorm-apply kernel (C = %d)", u.cout);
  KOA_TRY(bn_finalize(u, pv, ws, training, st));
  if (ud) KOA_TRY(bn_finalize(*ud, pv, ws, training, st));
  return koa_k_bn_act(at(ws, u.y), bn_slot(ws, u, S_SCALE), bn_slot(ws, u, S_SHIFT), res, ud ? at(ws, ud->y) : nullptr,
                      ud ? bn_slot(ws, *ud, S_SCALE) : nullptr, ud ? bn_slot(ws, *ud, S_SHIFT) : nullptr, out, out_bf,
                      u.rows_out, u.cout, 1, st);
}

// Forward of the y-free bottleneck tail (bn_gram.cu): batch statistics of conv3's output from the Gram matrix of its input a2
// (column sums `sa2` come from the BatchNorm-apply pass that wrote a2), then conv3 with BatchNorm + residual + ReLU in its
// epilogue. `res`: the identity x, or the raw output of the downsample convolution (whose BatchNorm the epilogue applies).
int gram_tail_forward(const Plan& p, const Block& b, const ParamView& pv, void* ws, const void* x, cudaStream_t st) {
  const Unit& u2 = p.units[b.u2];
  const Unit& u3 = p.units[b.u3];
  const int w = u2.cout, c = u3.cout;
  const double count = (double)u3.rows_out;
  float* gram = (float*)at(ws, b.gram);
  float* q = (float*)at(ws, b.q);
  // Gram = a2^T a2: a weight-gradient-shaped GEMM over the pixels, both operands the fp16 tensor itself (products exact in fp32)
  const KoaFlopScale bookkeeping(0.0);  // Gram and Q are not part of the reference's arithmetic: 0 algorithmic FLOPs
  KOA_TRY(koa_gemm_wgrad_launch(at(ws, b.a2), at(ws, b.a2), gram, (int)u2.rows_out, w, w, 1, st));
  KOA_TRY(koa_k_gram_split(gram, at(ws, b.gram_hi), at(ws, b.gram_lo), w, count, st));
  // Q = W3 . (hi + lo): two small GEMMs with fp16 operands and an fp32 result (the second accumulates onto the first)
  for (int part = 0; part < 2; ++part) {
    koa_epilogue_t ep{};
    ep.out = q;
    ep.ldo = w;
    ep.out_fp32 = 1;
    ep.a_f16 = ep.b_f16 = 1;
    if (part == 1) ep.residual_f32 = q;
    KOA_TRY(koa_gemm_launch(at(ws, u3.w_fwd), at(ws, part == 0 ? b.gram_hi : b.gram_lo), c, w, w, &ep, st));
  }
  KOA_TRY(koa_k_bn_gram_stats(at(ws, u3.w_fwd), (const float*)at(ws, b.sa2), q, pv.gamma(u3), pv.beta(u3), pv.run_mean(u3),
                              pv.run_var(u3), bn_slot(ws, u3, S_SCALE), bn_slot(ws, u3, S_SHIFT), bn_slot(ws, u3, S_MEAN),
                              bn_slot(ws, u3, S_INVSTD), c, w, count, st));
  koa_profile_flop_scale(1.0);  // (the scope object restores the caller's value at return)
  const void* res = x;
  const Unit* res_bn = nullptr;
  if (b.ud >= 0) {
    const Unit& ud = p.units[b.ud];
    KOA_TRY(conv_forward(p, ud, x, ws, 1, st));
    KOA_TRY(bn_finalize(ud, pv, ws, 1, st));
    res = at(ws, ud.y);
    res_bn = &ud;
  }
  return conv_bn_forward(p, u3, at(ws, b.a2), ws, res, res_bn, 1, at(ws, b.out), b.out_bf ? at(ws, b.out_bf) : nullptr, st);
}

// Inference forward of one block: every BatchNorm in the epilogue of its convolution.
int fused_eval_block(const Plan& p, const Block& b, const ParamView& pv, void* ws, const void* x, cudaStream_t st) {
  const Unit& u1 = p.units[b.u1];
  const Unit& u2 = p.units[b.u2];
  const Unit* last = b.kind == 0 ? &p.units[b.u3] : &u2;
  KOA_TRY(bn_finalize(u1, pv, ws, 0, st));
  KOA_TRY(conv_bn_forward(p, u1, x, ws, nullptr, nullptr, 1, at(ws, b.a1), nullptr, st));
  const void* last_in = at(ws, b.a1);
  if (b.kind == 0) {
    KOA_TRY(bn_finalize(u2, pv, ws, 0, st));
    KOA_TRY(conv_bn_forward(p, u2, at(ws, b.a1), ws, nullptr, nullptr, 1, at(ws, b.a2), nullptr, st));
    last_in = at(ws, b.a2);
  }
  const void* res = x;
  if (b.ud >= 0) {  // r = bn_d(conv_d(x)) (no activation), stored where the raw output would have gone
    const Unit& ud = p.units[b.ud];
    KOA_TRY(bn_finalize(ud, pv, ws, 0, st));
    KOA_TRY(conv_bn_forward(p, ud, x, ws, nullptr, nullptr, 0, at(ws, ud.y), nullptr, st));
    res = at(ws, ud.y);
  }
  KOA_TRY(bn_finalize(*last, pv, ws, 0, st));
  return conv_bn_forward(p, *last, last_in, ws, res, nullptr, 1, at(ws, b.out), nullptr, st);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" size_t koa_fe_workspace_bytes(const koa_fe_desc_t* d) {
  Plan p;
  if (build_plan(d, p)) return 0;
  return p.total;
}

extern "C" int koa_fe_out_shape(const koa_fe_desc_t* d, int* channels, int* h, int* w) {
  Plan p;
  KOA_TRY(build_plan(d, p));
  *channels = p.out_c; *h = p.out_h; *w = p.out_w;
  return 0;
}

extern "C" int koa_fe_num_units(const koa_fe_desc_t* d) {
  Plan p;
  if (build_plan(d, p)) return -1;
  return (int)p.units.size();
}

// what: 0 raw conv output y of unit `index`; 1 block output of block `index`; 2 a1 of block; 3 a2 of block;
// 4 stem activation a0; 5 pooled p0; 7 max-pool argmax bytes; 8/9/10 bf16 copies of block out / a1 / a2; 11 bf16 pooled;
// 6 BN coefficients (scale, shift, mean, invstd, k0, k1, k2) of unit `index`.
extern "C" int koa_fe_debug_offset(const koa_fe_desc_t* d, int what, int index, size_t* offset, size_t* bytes) {
  Plan p;
  KOA_TRY(build_plan(d, p));
  if (what == 0 || what == 6) {
    KOA_REQUIRE(index >= 0 && index < (int)p.units.size(), "unit index out of range");
    const Unit& u = p.units[index];
    // (no y: inference with fused BatchNorm epilogues, and the last convolution of a train-mode bottleneck)
    if (what == 0) { *offset = u.y; *bytes = (u.y == 0 && u.idx != 0) ? 0 : (size_t)u.rows_out * u.cout * 2; }
    else { *offset = u.coef; *bytes = (size_t)COEF_SLOTS * u.cout * 4; }
    return 0;
  }
  if (what >= 1 && what <= 3) {
    KOA_REQUIRE(index >= 0 && index < (int)p.blocks.size(), "block index out of range");
    const Block& b = p.blocks[index];
    const Unit& last = p.units[b.kind == 0 ? b.u3 : b.u2];
    const Unit& u1 = p.units[b.u1];
    const Unit& u2 = p.units[b.u2];
    if (what == 1) { *offset = b.out; *bytes = (size_t)last.rows_out * last.cout * 2; }
    if (what == 2) { *offset = b.a1; *bytes = (size_t)u1.rows_out * u1.cout * 2; }
    if (what == 3) { *offset = b.a2; *bytes = (size_t)u2.rows_out * u2.cout * 2; }
    return 0;
  }
  if (what >= 8 && what <= 10) {  // bf16 copies of block out / a1 / a2 (only with need_backward)
    KOA_REQUIRE(index >= 0 && index < (int)p.blocks.size(), "block index out of range");
    const Block& b = p.blocks[index];
    const Unit& last = p.units[b.kind == 0 ? b.u3 : b.u2];
    const Unit& u1 = p.units[b.u1];
    const Unit& u2 = p.units[b.u2];
    if (what == 8) { *offset = b.out_bf; *bytes = b.out_bf ? (size_t)last.rows_out * last.cout * 2 : 0; }
    if (what == 9) { *offset = b.a1_bf; *bytes = b.a1_bf ? (size_t)u1.rows_out * u1.cout * 2 : 0; }
    if (what == 10) { *offset = b.a2_bf; *bytes = b.a2_bf ? (size_t)u2.rows_out * u2.cout * 2 : 0; }
    return 0;
  }
  if (what == 12 || what == 13) {  // y-free tail: colsum(a2) [w] fp32 / Q = W3 . Gram / N [C][w] fp32 of block `index`
    KOA_REQUIRE(index >= 0 && index < (int)p.blocks.size(), "block index out of range");
    const Block& b = p.blocks[index];
    if (!p.gram) { *offset = 0; *bytes = 0; return 0; }
    const size_t w = (size_t)p.units[b.u2].cout, c = (size_t)p.units[b.u3].cout;
    if (what == 12) { *offset = b.sa2; *bytes = w * 4; }
    else { *offset = b.q; *bytes = c * w * 4; }
    return 0;
  }
  if (what == 11) { *offset = p.p0_bf; *bytes = !p.p0_bf ? 0 : (size_t)p.n_img * p.units[p.blocks[0].u1].hin * p.units[p.blocks[0].u1].win * 64 * 2; return 0; }
  if (what == 4) { *offset = p.a0; *bytes = (size_t)p.units[0].rows_out * 64 * 2; return 0; }
  if (what == 5) { *offset = p.p0; *bytes = (size_t)p.n_img * p.units[p.blocks[0].u1].hin * p.units[p.blocks[0].u1].win * 64 * 2; return 0; }
  if (what == 7) { *offset = p.idx0; *bytes = (size_t)p.n_img * p.units[p.blocks[0].u1].hin * p.units[p.blocks[0].u1].win * 64; return 0; }
  koa_set_error("unknown debug selector %d", what);
  return KOA_ERR_ARG;
}

extern "C" int koa_fe_forward(const koa_fe_desc_t* d, const void* const* params, const float* input, void* ws,
                              float* feat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Plan p;
  KOA_TRY(build_plan(d, p));
  KOA_REQUIRE(params != nullptr && input != nullptr && ws != nullptr && feat != nullptr, "null pointer argument");
  const ParamView pv{params};
  const int training = d->training;
  const bool need_dgrad = d->need_backward != 0;
  KOA_CHECK_CUDA(cudaMemsetAsync(at(ws, p.fstat_begin), 0, p.fstat_end - p.fstat_begin, st));

  KOA_TRY(pack_all_weights(p, pv, ws, need_dgrad, st));
  // ---- stem ---------------------------------------------------------------------------------------
  const Unit& us = p.units[0];
  const float* img = input;
  if (d->slices > 0) {  // (B,1,R,C,S) slice-innermost volume
    KOA_REQUIRE(d->n_img % d->slices == 0, "n_img must be batch * slices");
    KOA_TRY(koa_k_stem_pack(input, (float*)at(ws, p.img), d->n_img / d->slices, d->h * d->w, d->slices, st));
    img = (const float*)at(ws, p.img);
  }
  // stem as a tensor-core GEMM: [pixels, 64 (49 taps of the folded grey channel)] x [64, 64]^T
  KOA_TRY(koa_k_stem_pack_wb(pv.w(us), at(ws, p.wstem), st));
  KOA_TRY(koa_k_stem_im2col(img, at(ws, p.a_stem), p.n_img, d->h, d->w, 1, st));
  {
    koa_epilogue_t ep{};
    ep.out = at(ws, us.y);
    ep.ldo = 64;
    ep.a_f16 = ep.b_f16 = ep.out_f16 = 1;
    if (training) {
      ep.col_sum = bn_slot(ws, us, S_SUM);
      ep.col_sumsq = bn_slot(ws, us, S_SUMSQ);
    }
    if (p.fused_eval) {  // BatchNorm (running statistics) + ReLU in the epilogue: a0 directly
      KOA_TRY(bn_finalize(us, pv, ws, 0, st));
      ep.out = at(ws, p.a0);
      ep.act_f16 = 1;
      ep.bn_scale = bn_slot(ws, us, S_SCALE);
      ep.bn_shift = bn_slot(ws, us, S_SHIFT);
      ep.act = KOA_ACT_RELU;
    }
    const KoaFlopScale fs(49.0 / 64.0);  // 49 taps of the folded grey channel, padded to K = 64
    KOA_TRY(koa_gemm_launch(at(ws, p.a_stem), at(ws, p.wstem), (int)us.rows_out, 64, 64, &ep, st));
  }
  if (!p.fused_eval) KOA_TRY(bn_apply(us, nullptr, pv, ws, nullptr, at(ws, p.a0), nullptr, training, st));
  KOA_TRY(koa_k_maxpool_fwd(at(ws, p.a0), at(ws, p.p0), p.p0_bf ? at(ws, p.p0_bf) : nullptr, at(ws, p.idx0), p.n_img, us.hout,
                            us.wout, 64, st));

  // ---- residual blocks --------------------------------------------------------------------------
  for (const Block& b : p.blocks) {
    const Unit& u1 = p.units[b.u1];
    const Unit& u2 = p.units[b.u2];
    const void* x = at(ws, b.in);
    if (p.fused_eval) {
      KOA_TRY(fused_eval_block(p, b, pv, ws, x, st));
      continue;
    }
    KOA_TRY(conv_forward(p, u1, x, ws, training, st));
    KOA_TRY(bn_apply(u1, nullptr, pv, ws, nullptr, at(ws, b.a1), b.a1_bf ? at(ws, b.a1_bf) : nullptr, training, st));
    KOA_TRY(conv_forward(p, u2, at(ws, b.a1), ws, training, st));
    const Unit* last = &u2;
    if (b.kind == 0) {
      const Unit& u3 = p.units[b.u3];
      KOA_TRY(bn_apply(u2, nullptr, pv, ws, nullptr, at(ws, b.a2), b.a2_bf ? at(ws, b.a2_bf) : nullptr, training, st,
                       p.gram ? (float*)at(ws, b.sa2) : nullptr));
      if (p.gram) {
        KOA_TRY(gram_tail_forward(p, b, pv, ws, x, st));
        continue;
      }
      KOA_TRY(conv_forward(p, u3, at(ws, b.a2), ws, training, st));
      last = &u3;
    }
    void* out_bf = b.out_bf ? at(ws, b.out_bf) : nullptr;
    if (b.ud >= 0) {
      const Unit& ud = p.units[b.ud];
      KOA_TRY(conv_forward(p, ud, x, ws, training, st));
      KOA_TRY(bn_apply(*last, &ud, pv, ws, nullptr, at(ws, b.out), out_bf, training, st));
    } else {
      KOA_TRY(bn_apply(*last, nullptr, pv, ws, x, at(ws, b.out), out_bf, training, st));
    }
  }
  // ---- head: global average pool (with_gap) or the raw NHWC map as tokens -------------------------
  const Block& lb = p.blocks.back();
  const int hw = p.out_h * p.out_w;
  if (d->with_gap) KOA_TRY(koa_k_gap_fwd(at(ws, lb.out), feat, p.n_img, hw, p.out_c, st));
  else KOA_TRY(koa_k_gap_fwd(at(ws, lb.out), feat, p.n_img * hw, 1, p.out_c, st));
  return 0;
}

namespace {

// BatchNorm backward of one unit (optionally two units sharing the incoming gradient):
// reduce -> finalize (dgamma/dbeta + coefficients) -> apply.
// `stats_done`: sum_dz / sum_dz_xhat of `u` were already accumulated by the epilogue of the GEMM that produced `dout`.
int bn_backward(const Unit& u, const Unit* u_b, const ParamView& pv, void* const* grads, void* ws, const void* dout,
                const void* act_mask, void* dy, void* dy_b, int training, bool stats_done, cudaStream_t st) {
  if (!stats_done)
    KOA_TRY(koa_k_bn_bwd_reduce(dout, act_mask, at(ws, u.y), bn_slot(ws, u, S_MEAN), bn_slot(ws, u, S_INVSTD),
                                u_b ? at(ws, u_b->y) : nullptr, u_b ? bn_slot(ws, *u_b, S_MEAN) : nullptr,
                                u_b ? bn_slot(ws, *u_b, S_INVSTD) : nullptr, bn_slot(ws, u, S_SDZ), bn_slot(ws, u, S_SDZX),
                                u_b ? bn_slot(ws, *u_b, S_SDZX) : nullptr, u.rows_out, u.cout, st));
  if (koa_k_bn_fused_ok(u.cout)) {  // finalize (dgamma / dbeta + coefficients) inside the apply kernel
    const KoaBnBwdFin fa{bn_slot(ws, u, S_SDZ), bn_slot(ws, u, S_SDZX), pv.gamma(u), bn_slot(ws, u, S_MEAN),
                         bn_slot(ws, u, S_INVSTD), (float*)grads[u.idx * 3 + 1], (float*)grads[u.idx * 3 + 2],
                         bn_slot(ws, u, S_K0), bn_slot(ws, u, S_K1), bn_slot(ws, u, S_K2)};
    KoaBnBwdFin fb{};
    if (u_b)
      fb = KoaBnBwdFin{bn_slot(ws, u, S_SDZ), bn_slot(ws, *u_b, S_SDZX), pv.gamma(*u_b), bn_slot(ws, *u_b, S_MEAN),
                       bn_slot(ws, *u_b, S_INVSTD), (float*)grads[u_b->idx * 3 + 1], (float*)grads[u_b->idx * 3 + 2],
                       bn_slot(ws, *u_b, S_K0), bn_slot(ws, *u_b, S_K1), bn_slot(ws, *u_b, S_K2)};
    return koa_k_bn_bwd_apply_fin(dout, act_mask, at(ws, u.y), &fa, dy, u_b ? at(ws, u_b->y) : nullptr, u_b ? &fb : nullptr,
                                  dy_b, u.rows_out, u.cout, (double)u.rows_out, training, st);
  }
  KOA_TRY(koa_k_bn_bwd_finalize(bn_slot(ws, u, S_SDZ), bn_slot(ws, u, S_SDZX), pv.gamma(u), bn_slot(ws, u, S_MEAN),
                                bn_slot(ws, u, S_INVSTD), (float*)grads[u.idx * 3 + 1], (float*)grads[u.idx * 3 + 2],
                                bn_slot(ws, u, S_K0), bn_slot(ws, u, S_K1), bn_slot(ws, u, S_K2), u.cout,
                                (double)u.rows_out, training, st));
  if (u_b) {
    KOA_TRY(koa_k_bn_bwd_finalize(bn_slot(ws, u, S_SDZ), bn_slot(ws, *u_b, S_SDZX), pv.gamma(*u_b),
                                  bn_slot(ws, *u_b, S_MEAN), bn_slot(ws, *u_b, S_INVSTD), (float*)grads[u_b->idx * 3 + 1],
                                  (float*)grads[u_b->idx * 3 + 2], bn_slot(ws, *u_b, S_K0), bn_slot(ws, *u_b, S_K1),
                                  bn_slot(ws, *u_b, S_K2), u_b->cout, (double)u_b->rows_out, training, st));
  }
  return koa_k_bn_bwd_apply(dout, act_mask, at(ws, u.y), bn_slot(ws, u, S_K0), bn_slot(ws, u, S_K1), bn_slot(ws, u, S_K2),
                            dy, u_b ? at(ws, u_b->y) : nullptr, u_b ? bn_slot(ws, *u_b, S_K0) : nullptr,
                            u_b ? bn_slot(ws, *u_b, S_K1) : nullptr, u_b ? bn_slot(ws, *u_b, S_K2) : nullptr, dy_b,
                            u.rows_out, u.cout, st);
}

// dW of one unit into the fp32 gradient tensor (PyTorch layout [Cout][Cin/g][k][k]). `x`: bf16 copy of the unit's input.
// `x_bf`: offset of the bf16 copy of the unit's input (0 = none: the fp16 activation at `x_f16` is converted in-kernel).
int conv_wgrad(const Plan& p, const Unit& u, size_t x_bf, size_t x_f16, const void* dy, void* const* grads, void* ws,
               cudaStream_t st) {
  float* gw = (float*)grads[u.idx * 3 + 0];
  if (gw == nullptr) return 0;
  const void* x = at(ws, x_bf ? x_bf : x_f16);
  const int xf = x_bf ? 0 : 2;  // 2: x is the fp16 activation, converted to bf16 inside the kernel
  if (u.k == 1 && u.stride == 1) return koa_gemm_wgrad_launch(dy, x, gw, (int)u.rows_out, u.cout, u.cin, xf, st);
  if (u.k == 1) return koa_conv_wgrad_launch(dy, x, gw, p.n_img, u.hin, u.win, u.cin, u.cout, 1, 1, u.stride, 0, xf, st);
  float* scratch = (float*)at(ws, u.dw_scratch);
  if (u.groups > 1) {
    KOA_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)u.cout * 9 * 64 * 4, st));
    {
      const KoaFlopScale fs((double)(u.cin / u.groups) / 64.0);
      KOA_TRY(koa_conv_grouped_wgrad_launch(dy, x, scratch, p.n_img, u.hin, u.win, u.cin, u.stride, xf, st));
    }
    return koa_k_unpack_grouped_dw(scratch, gw, u.cout, u.cin / u.groups, st);
  }
  KOA_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)u.cout * u.k * u.k * u.cin * 4, st));
  KOA_TRY(koa_conv_wgrad_launch(dy, x, scratch, p.n_img, u.hin, u.win, u.cin, u.cout, u.k, u.k, u.stride, u.pad, xf, st));
  return koa_k_unpack_conv_dw(scratch, gw, u.cout, u.cin, u.k, u.k, st);
}

// dx = data gradient of unit u given dy. `ep` carries the output pointer (+ optional fused addends).
// tmp: scratch for the zero-inserted dy of stride-2 3x3 convolutions.
int conv_dgrad(const Plan& p, const Unit& u, const void* dy, void* ws, koa_epilogue_t* ep, void* tmp, cudaStream_t st) {
  ep->ldo = u.cin;
  ep->a_f16 = ep->b_f16 = 0;  // dy and the data-gradient form of the weights: bf16
  ep->out_f16 = 0;            // dx: bf16 gradient
  ep->act_f16 = 1;            // gate / stat_y are fp16 forward activations
  if (u.k == 1) {
    // 1x1: dx[rows_out, cin] = dy[rows_out, cout] . W[cout, cin]; B operand = W^T stored [cin][cout]
    return koa_gemm_launch(dy, at(ws, u.w_dgrad), (int)u.rows_out, u.cin, u.cout, ep, st);
  }
  const void* src = dy;
  double alg = 1.0;
  if (u.stride == 2) {  // three of four rows of the zero-inserted gradient are structural zeros
    KOA_TRY(koa_k_zero_insert2(dy, tmp, p.n_img, u.hin, u.win, u.cout, u.hout, u.wout, st));
    src = tmp;
    alg = 0.25;
  }
  if (u.groups > 1) {
    const KoaFlopScale fs(alg * (double)(u.cin / u.groups) / 64.0);
    return koa_conv_grouped_launch(src, at(ws, u.w_dgrad), p.n_img, u.hin, u.win, u.cout, 1, ep, st);
  }
  const KoaFlopScale fs(alg);
  return koa_conv_fprop_launch(src, at(ws, u.w_dgrad), p.n_img, u.hin, u.win, u.cout, u.cin, u.k, u.k, 1, u.pad, ep, st);
}

// Gate an epilogue by the forward activation `act_off` (ReLU backward) and, when `fuse`, let it accumulate the
// BatchNorm-backward reductions of unit `u` (whose ReLU output `act_off` is) on the values it stores.
void gate_and_stats(koa_epilogue_t& ep, void* ws, size_t act_off, const Unit* u, bool fuse) {
  ep.gate_bf16 = at(ws, act_off);
  if (fuse && u != nullptr) {
    ep.col_sum = bn_slot(ws, *u, S_SDZ);
    ep.col_sumsq = bn_slot(ws, *u, S_SDZX);
    ep.stat_y = at(ws, u->y);
    ep.stat_mean = bn_slot(ws, *u, S_MEAN);
    ep.stat_invstd = bn_slot(ws, *u, S_INVSTD);
  }
}

// Producer-side statistics of the y-free tail: only sum(G) is needed from the epilogue that produces G (sum(G * y3) comes
// out of the weight-gradient GEMM, bn_gram.cu); the sum-of-squares slot the kernel also fills is scratch.
void gate_and_colsum(koa_epilogue_t& ep, void* ws, size_t act_off, const Unit* u, bool fuse) {
  ep.gate_bf16 = at(ws, act_off);
  if (fuse && u != nullptr) {
    ep.col_sum = bn_slot(ws, *u, S_SDZ);
    ep.col_sumsq = bn_slot(ws, *u, S_SDZX);
  }
}

bool fuse_bn_bwd_stats() {
  static const int v = [] {
    const char* e = getenv("KOA_FUSE_BN_BWD");
    return e == nullptr ? 1 : atoi(e);
  }();
  return v != 0;
}

// ---- weight gradients on a side stream ------------------------------------------------------------------------------
// Nothing downstream in the backward chain reads a weight gradient, so the wgrad GEMMs (compute-bound, 17 % of a step)
// CAN run on a second stream next to the BatchNorm-backward passes (HBM-bound) of the main chain. OFF by default
// (KOA_WGRAD_STREAM=1 enables it): measured on B200 it is neutral when the modality branches already run on concurrent
// streams (165.0 vs 164.3 knees/s) and a loss with serial branches (123.8 vs 107.1 ms per step: the wgrad CTAs and the
// persistent data-gradient CTAs evict each other's shared memory residency instead of overlapping). Dependencies: a wgrad
// starts after its dy is ready (event on the main stream) and the main chain waits for it before it overwrites the dy
// scratch buffer the wgrad reads (one event per scratch buffer).
struct WgradSide {
  cudaStream_t side = nullptr;
  cudaEvent_t ready = nullptr, all_done = nullptr;
  cudaEvent_t done[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool pending[5] = {false, false, false, false, false};
  cudaStream_t main = nullptr;
  bool on = false;

  // the main chain is about to overwrite scratch buffer k
  int before_write(int k) {
    if (on && pending[k]) {
      KOA_CHECK_CUDA(cudaStreamWaitEvent(main, done[k], 0));
      pending[k] = false;
    }
    return 0;
  }
  // stream for a wgrad whose operands the main chain has just produced
  int begin(cudaStream_t* out) {
    *out = main;
    if (!on) return 0;
    KOA_CHECK_CUDA(cudaEventRecord(ready, main));
    KOA_CHECK_CUDA(cudaStreamWaitEvent(side, ready, 0));
    *out = side;
    return 0;
  }
  // the wgrad just launched reads scratch buffer k
  int reads(int k) {
    if (!on) return 0;
    KOA_CHECK_CUDA(cudaEventRecord(done[k], side));
    pending[k] = true;
    return 0;
  }
  int join() {
    if (!on) return 0;
    KOA_CHECK_CUDA(cudaEventRecord(all_done, side));
    KOA_CHECK_CUDA(cudaStreamWaitEvent(main, all_done, 0));
    for (bool& p : pending) p = false;
    return 0;
  }
};

bool wgrad_side_enabled() {
  static const int v = [] {
    const char* e = getenv("KOA_WGRAD_STREAM");
    return e == nullptr ? 0 : atoi(e);
  }();
  return v != 0;
}

// one side stream (+ events) per calling stream, created on first use and kept for the life of the process
int get_wgrad_side(cudaStream_t main, WgradSide* out) {
  static std::mutex mu;
  static std::map<cudaStream_t, WgradSide> pool;
  *out = WgradSide{};
  out->main = main;
  if (!wgrad_side_enabled()) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(main, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return 0;
  std::lock_guard<std::mutex> lk(mu);
  auto it = pool.find(main);
  if (it == pool.end()) {
    WgradSide w;
    w.main = main;
    KOA_CHECK_CUDA(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking));
    KOA_CHECK_CUDA(cudaEventCreateWithFlags(&w.ready, cudaEventDisableTiming));
    KOA_CHECK_CUDA(cudaEventCreateWithFlags(&w.all_done, cudaEventDisableTiming));
    for (auto& e : w.done) KOA_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    it = pool.emplace(main, w).first;
  }
  *out = it->second;
  out->on = true;
  for (bool& p : out->pending) p = false;
  return 0;
}

}  // namespace

// Backward. The gradient handed from block to block is G_b = dL/d(pre-ReLU residual sum of block b), i.e. the
// gradient w.r.t. the block output already multiplied by (out_b > 0): the ReLU mask is applied once, by the epilogue
// that produces the gradient (gate = the forward activation), never re-read by the BatchNorm kernels. The same
// epilogues accumulate sum(dz) and sum(dz * xhat) of the BatchNorm that follows in backward order, so the separate
// reduction pass only remains for the BatchNorm pairs that share a gradient (blocks with a downsample branch), for
// gradients assembled by more than one kernel (stride-2 downsample) and for the stem.
extern "C" int koa_fe_backward(const koa_fe_desc_t* d, const void* const* params, void* const* grads, void* ws,
                               const float* dfeat, void* stream) {
  return koa_fe_backward_range(d, params, grads, ws, dfeat, 0, -1, 1, stream);
}

extern "C" int koa_fe_num_blocks(const koa_fe_desc_t* d) {
  Plan p;
  if (build_plan(d, p)) return -1;
  return (int)p.blocks.size();
}

namespace {
// does the epilogue that produces G for block `bi` (the conv1 / downsample data gradient of block bi + 1) also reduce it
// for the last BatchNorm of block `bi`? A pure function of the plan, so that a backward pass split over several calls
// (koa_fe_backward_range) needs no state besides the workspace.
bool producer_reduces_g(const Plan& p, int bi, bool fuse) {
  if (bi + 1 >= (int)p.blocks.size()) return false;  // G of the last block comes from the pooling backward
  const Block& nb = p.blocks[bi + 1];
  const Unit* ud = nb.ud >= 0 ? &p.units[nb.ud] : nullptr;
  const bool single_producer = !(ud && ud->stride != 1);
  return fuse && (p.gram || p.blocks[bi].ud < 0) && single_producer;
}
}  // namespace

// Blocks [block_begin, block_end) of the backward pass, from the last to the first; block_end < 0 or == the block count
// starts the pass (statistics reset + pooling backward), with_stem != 0 ends it (max-pool / stem backward after block 0;
// needs block_begin == 0). Successive calls on the same stream with adjacent ranges compute exactly what one call does:
// the data-parallel wrapper issues the gradient all-reduce of a stage between two calls (dataparallel.py).
extern "C" int koa_fe_backward_range(const koa_fe_desc_t* d, const void* const* params, void* const* grads, void* ws,
                                     const float* dfeat, int block_begin, int block_end, int with_stem, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Plan p;
  KOA_TRY(build_plan(d, p));
  KOA_REQUIRE(params != nullptr && grads != nullptr && ws != nullptr && dfeat != nullptr, "null pointer argument");
  KOA_REQUIRE(d->need_backward, "forward was not run with need_backward");
  const int n_blocks = (int)p.blocks.size();
  if (block_end < 0) block_end = n_blocks;
  KOA_REQUIRE(block_begin >= 0 && block_begin <= block_end && block_end <= n_blocks, "bad block range [%d, %d) of %d",
              block_begin, block_end, n_blocks);
  KOA_REQUIRE(!with_stem || block_begin == 0, "the stem follows block 0");
  const ParamView pv{params};
  const int training = d->training;
  const int hw = p.out_h * p.out_w;
  const bool fuse = fuse_bn_bwd_stats();
  WgradSide side;
  KOA_TRY(get_wgrad_side(st, &side));
  cudaStream_t sw = st;  // stream of the next weight-gradient GEMM
  int cur = (n_blocks - block_end) & 1;  // p.g[cur] holds G of the current block (flips once per block)
  if (block_end == n_blocks) {
    KOA_CHECK_CUDA(cudaMemsetAsync(at(ws, p.bstat_begin), 0, p.bstat_end - p.bstat_begin, st));
    const Block& lb = p.blocks.back();
    if (d->with_gap) KOA_TRY(koa_k_gap_bwd(dfeat, at(ws, lb.out), at(ws, p.g[cur]), p.n_img, hw, p.out_c, st));
    else KOA_TRY(koa_k_gap_bwd(dfeat, at(ws, lb.out), at(ws, p.g[cur]), p.n_img * hw, 1, p.out_c, st));
  }
  // the producer of G already reduced it for the last BatchNorm of the current block
  bool g_stats_done = block_end > 0 && producer_reduces_g(p, block_end - 1, fuse);

  for (int bi = block_end - 1; bi >= block_begin; --bi) {
    const Block& b = p.blocks[bi];
    const Unit& u1 = p.units[b.u1];
    const Unit& u2 = p.units[b.u2];
    const Unit* u3 = b.kind == 0 ? &p.units[b.u3] : nullptr;
    const Unit* ud = b.ud >= 0 ? &p.units[b.ud] : nullptr;
    const Unit& last = u3 ? *u3 : u2;
    void* g_out = at(ws, p.g[cur]);
    void* g_in = at(ws, p.g[cur ^ 1]);
    void* x = at(ws, b.in);
    void* dy_last = at(ws, p.t[0]);
    void* dy_down = at(ws, p.t[1]);
    KOA_TRY(side.before_write(0));
    KOA_TRY(side.before_write(1));
    void* d_a1 = at(ws, p.t[3]);
    if (p.gram) {
      // y-free tail (bn_gram.cu): no dy3 tensor. T = G^T a2 -> coefficients, dW3, the extended data-gradient operand ->
      // dz2 = ([G | a2] . wext^T + bias) * (a2 > 0) in ONE GEMM.
      const int w = u2.cout, c = u3->cout;
      const double count = (double)u3->rows_out;
      // sum(G) per channel: from the producer of G when it reduced it (into the downsample unit's slots when the block has
      // one: that BatchNorm needs sum(G * xhat_d) from the same epilogue, and sum(G) is the same for both), else by a pass
      const float* sdz = bn_slot(ws, (ud && g_stats_done) ? *ud : *u3, S_SDZ);
      if (!g_stats_done)
        KOA_TRY(koa_k_col_stats(g_out, bn_slot(ws, *u3, S_SDZ), bn_slot(ws, *u3, S_SDZX), u3->rows_out, c, st));
      if (ud) KOA_TRY(bn_backward(*ud, nullptr, pv, grads, ws, g_out, nullptr, dy_down, nullptr, training, g_stats_done, st));
      float* tq = (float*)at(ws, b.tq);
      KOA_TRY(koa_gemm_wgrad_launch(g_out, at(ws, b.a2_bf), tq, (int)u3->rows_out, c, w, 0, st));
      KOA_TRY(koa_k_bn_gram_bwd(tq, at(ws, u3->w_fwd), pv.w(*u3), (const float*)at(ws, b.sa2), (const float*)at(ws, b.q),
                                sdz, pv.gamma(*u3), bn_slot(ws, *u3, S_MEAN), bn_slot(ws, *u3, S_INVSTD),
                                (float*)grads[u3->idx * 3 + 1], (float*)grads[u3->idx * 3 + 2], (float*)grads[u3->idx * 3 + 0],
                                at(ws, b.wext), at(ws, b.k2w), bn_slot(ws, *u3, S_K0), bn_slot(ws, *u3, S_K1),
                                bn_slot(ws, *u3, S_K2), c, w, c + w, count, st));
      KOA_TRY(koa_k_bn_gram_bias(at(ws, u3->w_dgrad), bn_slot(ws, *u3, S_K1), (float*)at(ws, b.cbias), w, c, st));
      {  // -M = W3^T . (-k2 * W3) [w][w], stored as the last w columns of every wext row
        const KoaFlopScale bookkeeping(0.0);
        koa_epilogue_t em{};
        em.out = (uint8_t*)at(ws, b.wext) + (size_t)c * 2;
        em.ldo = c + w;
        KOA_TRY(koa_gemm_launch(at(ws, u3->w_dgrad), at(ws, b.k2w), w, w, c, &em, st));
      }
      void* d_a2 = at(ws, p.t[2]);
      koa_epilogue_t ep{};
      ep.out = d_a2;
      ep.ldo = w;
      ep.act_f16 = 1;
      ep.col_bias = (const float*)at(ws, b.cbias);
      gate_and_stats(ep, ws, b.a2, &u2, fuse);  // dz2 = (...) * (a2 > 0) + sums for bn2
      KOA_TRY(side.before_write(2));
      {
        const KoaFlopScale fs((double)c / (double)(c + w));  // the reference's data gradient has K = C; the a2 segment is ours
        KOA_TRY(koa_gemm_kcat_launch(g_out, c, at(ws, b.a2_bf), w, at(ws, b.wext), (int)u3->rows_out, w, &ep, st));
      }
      KOA_TRY(bn_backward(u2, nullptr, pv, grads, ws, d_a2, nullptr, d_a2, nullptr, training, fuse, st));  // in place -> dy2
      KOA_TRY(side.begin(&sw));
      KOA_TRY(conv_wgrad(p, u2, b.a1_bf, b.a1, d_a2, grads, ws, sw));
      KOA_TRY(side.reads(2));
      koa_epilogue_t ep2{};
      ep2.out = d_a1;
      gate_and_stats(ep2, ws, b.a1, &u1, fuse);
      KOA_TRY(side.before_write(3));
      KOA_TRY(conv_dgrad(p, u2, d_a2, ws, &ep2, at(ws, p.t[4]), st));
    } else {
    KOA_TRY(bn_backward(last, ud, pv, grads, ws, g_out, nullptr, dy_last, ud ? dy_down : nullptr, training, g_stats_done, st));
    if (u3) {
      KOA_TRY(side.begin(&sw));
      KOA_TRY(conv_wgrad(p, *u3, b.a2_bf, b.a2, dy_last, grads, ws, sw));
      KOA_TRY(side.reads(0));
      void* d_a2 = at(ws, p.t[2]);
      koa_epilogue_t ep{};
      ep.out = d_a2;
      gate_and_stats(ep, ws, b.a2, &u2, fuse);  // dz2 = dgrad * (a2 > 0)
      KOA_TRY(side.before_write(2));
      KOA_TRY(conv_dgrad(p, *u3, dy_last, ws, &ep, nullptr, st));
      KOA_TRY(bn_backward(u2, nullptr, pv, grads, ws, d_a2, nullptr, d_a2, nullptr, training, fuse, st));  // in place -> dy2
      KOA_TRY(side.begin(&sw));
      KOA_TRY(conv_wgrad(p, u2, b.a1_bf, b.a1, d_a2, grads, ws, sw));
      KOA_TRY(side.reads(2));
      koa_epilogue_t ep2{};
      ep2.out = d_a1;
      gate_and_stats(ep2, ws, b.a1, &u1, fuse);
      KOA_TRY(side.before_write(3));
      KOA_TRY(conv_dgrad(p, u2, d_a2, ws, &ep2, at(ws, p.t[4]), st));
    } else {
      KOA_TRY(side.begin(&sw));
      KOA_TRY(conv_wgrad(p, u2, b.a1_bf, b.a1, dy_last, grads, ws, sw));
      KOA_TRY(side.reads(0));
      koa_epilogue_t ep2{};
      ep2.out = d_a1;
      gate_and_stats(ep2, ws, b.a1, &u1, fuse);
      KOA_TRY(side.before_write(3));
      KOA_TRY(conv_dgrad(p, u2, dy_last, ws, &ep2, at(ws, p.t[4]), st));
    }
    }
    KOA_TRY(bn_backward(u1, nullptr, pv, grads, ws, d_a1, nullptr, d_a1, nullptr, training, fuse, st));  // in place -> dy1
    KOA_TRY(side.begin(&sw));
    KOA_TRY(conv_wgrad(p, u1, b.in_bf, b.in, d_a1, grads, ws, sw));
    KOA_TRY(side.reads(3));
    // G of the previous block = (dgrad(conv1) + identity path) * (x > 0); x is that block's output (or the pooled
    // stem activation, where the gate is a no-op for the gradient that survives the stem's own ReLU mask).
    const Block* prev = bi > 0 ? &p.blocks[bi - 1] : nullptr;
    const Unit* prev_last = prev ? &p.units[prev->kind == 0 ? prev->u3 : prev->u2] : nullptr;
    const bool single_producer = !(ud && ud->stride != 1);
    // (y-free tail: the previous block only needs sum(G) from this epilogue, also when it has a downsample branch)
    const bool fuse_prev = prev != nullptr && producer_reduces_g(p, bi - 1, fuse);
    (void)single_producer;
    auto gate_prev = [&](koa_epilogue_t& e) {
      if (!p.gram) gate_and_stats(e, ws, b.in, prev_last, fuse_prev);
      // y-free tail: the last BatchNorm of the previous block only needs sum(G); the statistics slot of the epilogue is
      // free for the BatchNorm of its downsample branch (sum(G) and sum(G * xhat_d), both into that unit's slots)
      else if (prev != nullptr && prev->ud >= 0) gate_and_stats(e, ws, b.in, &p.units[prev->ud], fuse_prev);
      else gate_and_colsum(e, ws, b.in, prev_last, fuse_prev);
    };
    if (!ud) {
      koa_epilogue_t ep{};
      ep.out = g_in;
      ep.add_bf16 = g_out;  // identity path: G of this block
      gate_prev(ep);
      KOA_TRY(conv_dgrad(p, u1, d_a1, ws, &ep, at(ws, p.t[4]), st));
    } else {
      KOA_TRY(side.begin(&sw));
      KOA_TRY(conv_wgrad(p, *ud, b.in_bf, b.in, dy_down, grads, ws, sw));
      KOA_TRY(side.reads(1));
      if (ud->stride == 1) {
        koa_epilogue_t ep{};
        ep.out = g_in;
        KOA_TRY(conv_dgrad(p, u1, d_a1, ws, &ep, at(ws, p.t[4]), st));
        koa_epilogue_t ep2{};
        ep2.out = g_in;
        ep2.add_bf16 = g_in;  // accumulate in place, then gate the sum
        gate_prev(ep2);
        KOA_TRY(conv_dgrad(p, *ud, dy_down, ws, &ep2, nullptr, st));
      } else {
        koa_epilogue_t ep{};
        ep.out = g_in;
        gate_and_stats(ep, ws, b.in, nullptr, false);
        KOA_TRY(conv_dgrad(p, u1, d_a1, ws, &ep, at(ws, p.t[4]), st));
        koa_epilogue_t ep2{};
        ep2.out = at(ws, p.t[2]);
        KOA_TRY(side.before_write(2));
        KOA_TRY(conv_dgrad(p, *ud, dy_down, ws, &ep2, nullptr, st));
        KOA_TRY(koa_k_scatter_add2(at(ws, p.t[2]), x, g_in, p.n_img, ud->hin, ud->win, ud->cin, ud->hout, ud->wout, st));
      }
    }
    g_stats_done = fuse_prev;
    cur ^= 1;
  }
  if (!with_stem) return side.join();
  // ---- stem -----------------------------------------------------------------------------------------
  const Unit& us = p.units[0];
  void* d_a0 = at(ws, p.t[0]);
  KOA_TRY(side.before_write(0));
  KOA_TRY(koa_k_maxpool_bwd(at(ws, p.g[cur]), at(ws, p.idx0), d_a0, p.n_img, us.hout, us.wout, 64, st));
  KOA_TRY(bn_backward(us, nullptr, pv, grads, ws, d_a0, at(ws, p.a0), d_a0, nullptr, training, false, st));
  if (grads[0] != nullptr) {
    // the forward pass kept its fp16 im2col operand; the weight-gradient kernel converts it to bf16 in shared memory
    // (x_f16 = 2) to pair it with the bf16 dy. This GEMM is HBM-bound (64 x 64 outputs over millions of pixels), so the
    // conversion is free and the second im2col pass (1 ms per step) is gone.
    const KoaFlopScale fs(49.0 / 64.0);
    KOA_TRY(koa_gemm_wgrad_launch(d_a0, at(ws, p.a_stem), (float*)at(ws, p.dwstem), (int)us.rows_out, 64, 64, 2, st));
    KOA_TRY(koa_k_stem_unfold_dwb((const float*)at(ws, p.dwstem), (float*)grads[0], st));
  }
  KOA_TRY(side.join());  // every weight gradient is complete in the order of the caller's stream
  return 0;
}
