/* koa_b200.h — C-ABI of the B200-native koafusion hot path (libkoa_b200.so).
 *
 * The reference (imedslab/OAProgressionMMF, package `koafusion`) has no FFI layer: its hot path is
 * the `koafusion.models` nn.Modules calling PyTorch eager ops (cuDNN / cuBLAS). Each entry point
 * below replaces one of those implicit library calls; the reference call site it stands in for is
 * cited as file:line relative to the reference root.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure; koa_last_error() gives
 *     the message for the calling thread. No exceptions, no torch types cross this boundary.
 *   - all pointers are device pointers owned by the caller (PyTorch caching allocator in the host
 *     mirror); `stream` is a cudaStream_t passed as void*. Nothing is allocated behind the caller's
 *     back except per-process immutable state (kernel attributes, driver entry points).
 *   - activations are NHWC 16-bit ("pixel-major": a 1x1 convolution is a plain GEMM): the CNN's forward
 *     activations and packed weights are fp16, every gradient and the transformer's operands are bf16; the
 *     per-operator GEMM / conv entry points take the operand formats as flags (default bf16). Parameters
 *     arrive as fp32 master copies in PyTorch layout and are packed inside the engines.
 *   - the library is sm_100a only and has no CPU path; loading it without a B200 works (symbol
 *     checks), calling compute entry points without one fails with KOA_ERR_CUDA.
 */
#ifndef KOA_B200_H
#define KOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOA_OK 0
#define KOA_ERR_ARG (-1)
#define KOA_ERR_CUDA (-2)
#define KOA_ERR_UNSUPPORTED (-3)
#define KOA_ERR_DEVICE (-4)

/* ---- library state ------------------------------------------------------------------------- */
const char* koa_last_error(void);
int koa_version(void);
/* Reads and clears the device-side diagnostic word (non-zero: a pipeline barrier timed out). */
int koa_debug_flag(unsigned int* out);
/* The same word (and the code of the FIRST barrier that timed out since the last koa_debug_flag) read without waiting for
 * the work in flight and without clearing: for a watchdog that wants to know why a stream does not drain. */
int koa_debug_flag_peek(unsigned int* latest, unsigned int* first);
/* Kernels launched by this library in this process so far. */
long long koa_launch_count(void);
/* Optional per-launch CUDA-event timing of the tcgen05 kernel family (used by bench.py for the roofline):
 * out[cls*3 + {0,1,2}] = device ms, algorithmic FLOPs (2*M*N*K), launches; cls 0 = forward/data-gradient
 * GEMMs and implicit-GEMM convolutions, cls 1 = weight-gradient GEMMs. Reading synchronises and clears. */
int koa_profile_enable(int on);
int koa_profile_read(double* out);
/* Writes a per-shape breakdown of the recorded launches to a text file (does not clear the records). */
int koa_profile_dump(const char* path);
/* Recorded launches that have started but not finished, as text lines "cls tag m n k" (never waits: for a watchdog that
 * wants to name the kernel a stream is stuck in). Returns their number. */
int koa_profile_pending(char* buf, int cap);

/* ---- fused GEMM epilogue --------------------------------------------------------------------- */
enum { KOA_ACT_NONE = 0, KOA_ACT_RELU = 1, KOA_ACT_GELU = 2, KOA_ACT_GELU_GRAD = 3 };

typedef struct koa_epilogue {
  void* out;                 /* [M, ldo] bf16, or fp32 when out_fp32 != 0 */
  int ldo;                   /* leading dimension of every [M, *] epilogue tensor, in elements */
  int out_fp32;
  int act;                   /* KOA_ACT_* applied after the bias */
  const float* bias;         /* [N] or NULL */
  void* pre_out_bf16;        /* optional bf16 copy of (acc + bias) before the activation */
  const void* aux_bf16;      /* KOA_ACT_GELU_GRAD: pre-activation h; value *= gelu'(h) */
  const float* residual_f32; /* optional fp32 addend */
  const void* add_bf16;      /* optional bf16 addend */
  const void* gate_bf16;     /* optional: the final value is zeroed where gate <= 0 (ReLU backward by the forward
                                activation, applied after every addend) */
  void* out_bf16_copy;       /* optional bf16 copy of the final value when out is fp32 */
  float* col_sum;            /* optional per-column sum / sum of squares of the stored bf16 output (BatchNorm batch */
  float* col_sumsq;          /* statistics), accumulated with atomics: caller zeroes them first */
  const void* stat_y;        /* BatchNorm backward form of the statistics: when non-NULL, col_sumsq receives */
  const float* stat_mean;    /* sum(out * xhat) with xhat = (stat_y - stat_mean) * stat_invstd (stat_y: bf16 */
  const float* stat_invstd;  /* [M, ldo], the forward conv output), i.e. col_sum / col_sumsq = dbeta / dgamma */
  int a_f16, b_f16;          /* operand format: 0 = bf16, 1 = fp16; must be equal (one format per tcgen05 kind::f16 MMA) */
  int out_f16;               /* out / pre_out_bf16 hold fp16 instead of bf16 */
  int act_f16;               /* gate_bf16 / stat_y / aux_bf16 (forward activations) hold fp16 instead of bf16 */
  float drop_p;              /* > 0: dropout after the activation, before the residual add: value *= mask / (1 - p), */
  unsigned int drop_site;    /* mask = koa_dropout_mask(drop_seed, drop_site, M, N, drop_p) (counter-based Philox, */
  unsigned long long drop_seed; /* regenerated in backward instead of stored) */
  /* BatchNorm folded into the convolution (koafusion/models/_torchvision.py:118-138 conv -> bn -> [+ identity] -> relu as
   * ONE kernel; eval mode, or train mode once the batch statistics are known): value = acc * bn_scale[col] + bn_shift[col],
   * + add_bf16 read as a 16-bit residual in the act_f16 format (itself * res_scale[col] + res_shift[col] when res_scale is
   * set: the downsample branch's raw convolution output), act = KOA_ACT_NONE / KOA_ACT_RELU; needs out_f16 = 1;
   * out_bf16_copy receives a bf16 copy of the same values. */
  const float* bn_scale;
  const float* bn_shift;
  const float* res_scale;
  const float* res_shift;
  const float* col_bias;     /* backward flavour: fp32 [N] added per column before the gate */
} koa_epilogue_t;

/* out[M,N] = epilogue(A[M,K] . B[N,K]^T); A, B bf16 row-major. tcgen05/TMEM tiles fed by TMA.
 * Replaces nn.Linear / F.linear (koafusion/models/_core_trf.py:104,112,145,148,161,163) and the 1x1
 * stride-1 nn.Conv2d of the bottlenecks (koafusion/models/_torchvision.py:29-31,108,112) in forward
 * and in data-gradient form. Requires K % 8 == 0 and N % 32 == 0. */
int koa_gemm_bf16(const void* a, const void* b, int m, int n, int k, const koa_epilogue_t* ep, void* stream);
/* Same with the reduction dimension split over two A tensors: out = epilogue([A1[M,K1] | A2[M,K2]] . B[N,K1+K2]^T),
 * K1 and K2 multiples of 64. The data gradient through a train-mode BatchNorm whose input was never stored is one such
 * GEMM over [G | a2] (autograd of koafusion/models/_torchvision.py:130-131). Convolution-flavour epilogues
 * only (16-bit output; addend / gate / statistics / col_bias). */
int koa_gemm_kcat_bf16(const void* a1, int k1, const void* a2, int k2, const void* b, int m, int n,
                       const koa_epilogue_t* ep, void* stream);

/* y[N,Ho,Wo,Cout] = epilogue(conv(x[N,H,W,Cin], w[Cout,R,S,Cin])): implicit GEMM, the activation
 * operand is fetched by im2col-mode TMA. Replaces nn.Conv2d 3x3 (stride 1/2) and 1x1 stride 2
 * (koafusion/models/_torchvision.py:23-31,110,211-213). Requires Cin % 64 == 0, Cout % 32 == 0.
 * The data gradient of a stride-1 kxk convolution is the same call on dY with flipped weights. */
int koa_conv_fprop_bf16(const void* x, const void* w, int n_img, int h, int w_in, int cin, int cout, int filt_r,
                        int filt_s, int stride, int pad, const koa_epilogue_t* ep, void* stream);

/* dw[Cout,Cin] += dy[P,Cout]^T . x[P,Cin] (fp32 accumulate with atomics; caller zeroes dw). Both operands bf16, or
 * both fp16 with x_f16 == 1 (tcgen05 kind::f16 takes one 16-bit format per instruction); x_f16 == 2: dy bf16 and x fp16,
 * x is converted to bf16 in shared memory inside the kernel (no bf16 copy of the forward activations in HBM).
 * Weight gradient of nn.Linear and of 1x1 stride-1 convolutions (autograd of the call sites above). */
int koa_gemm_wgrad_bf16(const void* dy, const void* x, float* dw, int pixels, int cout, int cin, int x_f16, void* stream);

/* dw[Cout,R,S,Cin] += conv weight gradient for dy[N,Ho,Wo,Cout], x[N,H,W,Cin]. */
int koa_conv_wgrad_bf16(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int cin, int cout,
                        int filt_r, int filt_s, int stride, int pad, int x_f16, void* stream);

/* Test/debug knob: MN-major shared-memory descriptor strides used by the wgrad kernels. */
int koa_debug_set_wgrad_desc(unsigned int lbo_bytes, unsigned int sbo_bytes, unsigned int k_adv_bytes);


/* ---- whole feature extractor (per-slice CNN) ---------------------------------------------------- */
/* Replaces `self._feK(t_inK)` of every model class, i.e. nn.Sequential(*resnet.children()[:-1])
 * (koafusion/models/_xrNmrMcP.py:40-59,218-220; _mrN_cnn_trf.py:20-28,121; _xr1mrN.py:19-33,132-133;
 * _xr1_cnn.py:15-19,66) together with the einops rearrange/repeat in front of it (_xrNmrMcP.py:209-213)
 * and its autograd backward (koafusion/run/train_prog_fus.py:165). */
enum { KOA_ARCH_RESNET18 = 0, KOA_ARCH_RESNET34 = 1, KOA_ARCH_RESNET50 = 2, KOA_ARCH_RESNEXT50_32X4D = 3 };

typedef struct koa_fe_desc {
  int arch;          /* KOA_ARCH_* */
  int n_img;         /* number of 2-D images = batch * slices */
  int h, w;          /* image size */
  int slices;        /* > 0: input is a (B,1,R,C,S) slice-innermost volume; 0: input is [n_img][h][w] */
  int with_gap;      /* 1: features [n_img][C]; 0: [n_img][h_out*w_out][C] (tokens per position) */
  int training;      /* BatchNorm batch statistics + running-stat update (1) or running statistics (0) */
  int need_backward; /* keep what koa_fe_backward needs and pack data-gradient weights */
  const float* input_for_backward; /* slices == 0 only: the [n_img][h][w] input again, for the stem wgrad */
} koa_fe_desc_t;

size_t koa_fe_workspace_bytes(const koa_fe_desc_t* d);
int koa_fe_out_shape(const koa_fe_desc_t* d, int* channels, int* h, int* w);
int koa_fe_num_units(const koa_fe_desc_t* d); /* conv+BN units; params has 5, grads 3 entries per unit */
/* params[5*u + {0..4}] = conv weight, bn weight, bn bias, running_mean, running_var of unit u in
 * state_dict order (fp32, PyTorch layout). running_* are updated in place when training. */
int koa_fe_forward(const koa_fe_desc_t* d, const void* const* params, const float* input, void* workspace,
                   float* feat, void* stream);
/* grads[3*u + {0,1,2}] = d conv weight, d bn weight, d bn bias (fp32, accumulated into; NULL skips). */
int koa_fe_backward(const koa_fe_desc_t* d, const void* const* params, void* const* grads, void* workspace,
                    const float* dfeat, void* stream);
/* The same backward pass in pieces: residual blocks [block_begin, block_end) from the last to the first. block_end < 0 (or
 * the block count, koa_fe_num_blocks) starts the pass; with_stem != 0 (needs block_begin == 0) ends it with the max-pool /
 * stem backward. Adjacent ranges issued in order on one stream compute exactly what koa_fe_backward computes; the
 * data-parallel wrapper puts the gradient all-reduce of a finished stage between two calls, so that it overlaps the rest of
 * the backward pass (nn.DataParallel reduces on the way back through the replicas: koafusion/run/train_prog_fus.py:84). */
int koa_fe_backward_range(const koa_fe_desc_t* d, const void* const* params, void* const* grads, void* workspace,
                          const float* dfeat, int block_begin, int block_end, int with_stem, void* stream);
int koa_fe_num_blocks(const koa_fe_desc_t* d);
int koa_fe_debug_offset(const koa_fe_desc_t* d, int what, int index, size_t* offset, size_t* bytes);

/* ---- whole token transformer (FeaT) --------------------------------------------------------------- */
/* Replaces FeaT.forward / Transformer / Attention / FeedForward (koafusion/models/_core_trf.py:118-205)
 * and their autograd backward. Tables follow the reference state_dict order (see feat_engine.cu). */
typedef struct koa_feat_desc {
  int batch, n_patches, dim, depth, heads, mlp_dim, num_classes;
  int with_cls;      /* prepend the learned CLS token */
  int compute_head;  /* run mlp_head0 on token 0 (dead compute for the per-sequence transformers) */
  int training;
  int need_backward;
  float emb_dropout, mlp_dropout; /* nn.Dropout probabilities (_core_trf.py:105,127,146-149,164); active when training */
  unsigned long long seed;   /* Philox seed of this call's dropout masks; backward must get the same descriptor */
} koa_feat_desc_t;

size_t koa_feat_workspace_bytes(const koa_feat_desc_t* d);
int koa_feat_num_params(const koa_feat_desc_t* d);
int koa_feat_probs_offset(const koa_feat_desc_t* d, int layer, size_t* offset, size_t* bytes);
int koa_feat_forward(const koa_feat_desc_t* d, const void* const* params, const float* tokens, void* workspace,
                     float* states_out, float* logits_out, void* stream);
int koa_feat_backward(const koa_feat_desc_t* d, const void* const* params, void* const* grads, void* workspace,
                      const float* d_states, const float* d_logits, float* d_tokens, void* stream);

/* ---- single operators ------------------------------------------------------------------------------ */
/* Small dense layers on CUDA cores, fp32: y = act(x[m,k] . w[n,k]^T + b). Used where N or K is too small
 * for tcgen05 tiles: FeatC1 Linear(9->2048)+GELU (koafusion/models/_xrNmrMcP.py:15-19; the input is three
 * z-scores and three one-hot pairs, so the product is a gather/scale of weight columns), the XR1Cnn head
 * (koafusion/models/_xr1_cnn.py:31-39) and mlp_head0.4 (koafusion/models/_core_trf.py:115). `pre` (optional)
 * receives the pre-activation for backward. */
int koa_linear_small_fwd(const float* x, const float* w, const float* b, float* y, float* pre, int m, int n, int k,
                         int act, void* stream);
/* dx (may be NULL) is overwritten; dw / db (may be NULL) are accumulated into. scratch: m*n floats. */
int koa_linear_small_bwd(const float* dy, const float* pre, const float* x, const float* w, float* scratch, float* dx,
                         float* dw, float* db, int m, int n, int k, int act, void* stream);
/* FocalLoss(gamma, reduction="mean") and its gradient w.r.t. the logits (koafusion/various/_losses.py:89-108). */
int koa_focal_loss(const float* logits, const long long* target, float* loss, float* dlogits, int batch, int classes,
                   float gamma, void* stream);
/* nn.LayerNorm over the last dim, eps 1e-5 (koafusion/models/_core_trf.py:111,190,192). */
int koa_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32, float* mean,
                      float* rstd, int rows, int d, void* stream);
int koa_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx,
                      float* dgamma, float* dbeta, int rows, int d, void* stream);
/* softmax(Q K^T * scale) V per (batch, head) on a packed bf16 qkv buffer (koafusion/models/_core_trf.py:170-180). */
int koa_attention_fwd(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                      void* stream);
int koa_attention_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                      int head_dim, float scale, void* stream);
/* Same with the format of the forward tensors as an argument: f16 != 0: qkv and out hold fp16 (the transformer's forward
 * format, KOA_FEAT_F16); dout / dqkv are bf16 either way. n <= 128 tokens with head_dim a multiple of 64 (<= 256) runs
 * on tcgen05 (one UMMA tile per contraction, attention_tc.cu), other head sizes on CUDA cores. */
int koa_attention_fwd_fmt(const void* qkv, void* out, float* probs, int batch, int n, int heads, int head_dim, float scale,
                          int f16, void* stream);
int koa_attention_bwd_fmt(const void* qkv, const float* probs, const void* dout, void* dqkv, int batch, int n, int heads,
                          int head_dim, float scale, int f16, void* stream);
/* (B,1,R,C,S) -> [B*S][R*C]: einops "b ch r c s -> (b s) ch r c" (koafusion/models/_xrNmrMcP.py:209-210). */
int koa_stem_pack(const float* vol, float* img, int batch, int rc, int slices, void* stream);
/* nn.MaxPool2d(3, 2, 1) on NHWC (koafusion/models/_torchvision.py:174); idx keeps the winning tap. Forward: fp16
 * activations in and out; backward: bf16 gradients in and out. */
int koa_maxpool_fwd(const void* x, void* out, void* idx, int n, int h, int w, int c, void* stream);
int koa_maxpool_bwd(const void* dout, const void* idx, void* dx, int n, int h, int w, int c, void* stream);
/* Dropout scale factors (0 or 1/(1-p)) of site `site` as an fp32 [rows][cols] tensor: exactly the mask the engines
 * and koa_gemm_bf16 apply for (seed, site) (nn.Dropout of koafusion/models/_core_trf.py:105,146-149,164 with a
 * counter-based generator). FeaT sites: 4*layer + {0: to_out, 1: ff GELU, 2: ff out}, 0xE000: embedding, 0xF000: head. */
int koa_dropout_mask(unsigned long long seed, unsigned int site, long long rows, int cols, float p, float* out,
                     void* stream);
/* nn.Dropout2d on the extractor output (koafusion/models/_xrNmrMcP.py:62-74,226-229: applied to (B*S, C, h, w)), on the
 * token layout x [n_img][positions][c] fp32: out = x * m[img][ch] with ONE draw per (image, channel) shared by all
 * positions; m = koa_dropout_mask(seed, site, n_img, c, p). The backward pass is the same call on the gradient.
 * out may alias x. Site 0xD000 + extractor index in the model classes. */
int koa_channel_dropout(const float* x, float* out, long long n_img, int positions, int c, unsigned long long seed,
                        unsigned int site, float p, void* stream);
/* per-channel sum / sum of squares of a bf16 [rows][c] tensor (stand-alone BatchNorm statistics). */
int koa_col_stats(const void* y, float* sum, float* sumsq, long long rows, int c, void* stream);

/* ---- either side of the path: optimiser step, input resampling, predictions --------- */
/* One tensor of an Adam step: fp32 parameter, its gradient and the two moment buffers (device pointers). */
typedef struct koa_adam_tensor {
  float* param;
  const float* grad;         /* NULL: the parameter is skipped, as torch.optim skips `p.grad is None` */
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
} koa_adam_tensor_t;

typedef struct koa_adam_hyper {
  double lr, beta1, beta2, eps, weight_decay;
  double grad_scale;         /* gradients are multiplied by this first; 0 means 1 */
  int step;                  /* 1 for the first update (bias corrections 1 - beta^step) */
  int decoupled_weight_decay; /* 0: Adam (weight_decay * param added to the gradient), 1: AdamW */
} koa_adam_hyper_t;

/* torch.optim.Adam / AdamW (amsgrad=False, maximize=False) on `n_tensors` tensors: the optimiser the reference builds
 * from dict_optimizers (koafusion/various/_optimizers.py:49-54, koafusion/run/train_prog_fus.py:88-91) and steps after
 * every backward (train_prog_fus.py:166). `tensors` is a HOST array of descriptors holding device pointers; the library
 * passes them to the kernel 64 at a time as launch parameters (no device-side table, no copy). 28 bytes of HBM traffic
 * per parameter element. */
int koa_adam_step(const koa_adam_tensor_t* tensors, int n_tensors, const koa_adam_hyper_t* h, void* stream);

enum { KOA_DT_F32 = 0, KOA_DT_U8 = 1, KOA_DT_U16 = 2, KOA_DT_I16 = 3 };

/* out[b] = scale[b] * interpolate(in[b]) + shift[b] for `batch` volumes of in_dims[3] -> out_dims[3] elements (last
 * dimension innermost; 2-D images pass a leading 1): torch.nn.functional.interpolate(mode="linear" / "bilinear" /
 * "trilinear", align_corners=False, recompute_scale_factor=True) as PTInterpolate applies it to every modality right
 * before the model (koafusion/preproc/_pt.py:175-200, koafusion/run/train_prog_fus.py:111-116,143-146). For the factors
 * of the recipes (0.5 / 1.0 on even sizes) this is a 2x2(x2) box mean. The input may be fp32 or the integer type the
 * volumes have on disk (KOA_DT_*); scale / shift (fp32 [batch], both or neither) fold an affine intensity map in front
 * of the model into the same pass (the interpolation is linear, so the two commute). */
int koa_resample_linear(const void* in, int in_dtype, float* out, int batch, const int* in_dims, const int* out_dims,
                        const float* scale, const float* shift, void* stream);

/* Per-volume coefficients of PTToUnitRange followed by PTNormalize(mean, std) (koafusion/preproc/_pt.py:75-124;
 * koafusion/datasets/_data_provider.py:297-334): z = ((x - min) / (max - min) - mean) / std = x * scale + shift with the
 * minimum / maximum taken over each of the `batch` volumes of n_per elements. workspace: 2 * batch unsigned ints;
 * minmax (optional): fp32 [batch][2]. */
int koa_unit_range_affine(const void* in, int in_dtype, int batch, long long n_per, float mean, float stdev,
                          unsigned int* workspace, float* scale, float* shift, float* minmax, void* stream);

/* One volume of a training batch: where its crop starts in the stored volume and the state of the random transforms. */
typedef struct koa_augment {
  int off0, off1, off2;      /* first voxel of the crop (RandomCrop / CenterCrop, koafusion/preproc/_np_nd.py:62-140) */
  int rotate;                /* 1: in-slice rotation by theta (PTRotate3DInSlice / PTRotate2D, _pt.py:257-358) */
  float cos_t, sin_t;
  float inv_gamma;           /* 0: no gamma correction; else u -> u^(1/gamma) (PTGammaCorrection, _pt.py:203-232) */
  float lo, range;           /* minimum and max - min of the crop: written by the call, the caller leaves them alone */
  int flip;                  /* RIGHT knees are mirrored to the LEFT orientation before the crop (koafusion/datasets/oai/
                                _dataset.py:303-316): 0 none, 1 along the columns (COR IW TSE, XR), 2 along the slices
                                (SAG 3D DESS, SAG T2 map); the crop offsets count in the mirrored volume */
} koa_augment_t;

/* The per-sample transform chain of the training loader (koafusion/datasets/_data_provider.py:297-334) followed by the
 * on-GPU downscale, on the volumes as stored: mirror (RIGHT knees) -> crop -> PTToUnitRange (min / max of the crop) -> rotation about the slice
 * axis (F.affine_grid + F.grid_sample, bilinear, zero padding, align_corners=False) -> gamma -> PTNormalize(mean, std)
 * -> PTInterpolate to out_dims. in: `batch` volumes of src_dims[3] elements (rows, columns, slices; slices innermost; a
 * 2-D image has 1 slice); params: DEVICE array [batch]; workspace: 2 * batch unsigned ints. The validation / test
 * chain is the same call with rotate = 0, inv_gamma = 0 and the centre-crop offsets. */
int koa_augment_resample(const void* in, int in_dtype, float* out, koa_augment_t* params, int batch, const int* src_dims,
                         const int* crop_dims, const int* out_dims, float mean, float stdev, unsigned int* workspace,
                         void* stream);

/* proba = softmax(logits, dim=1), pred = argmax(logits, dim=1) for logits [batch][classes]
 * (koafusion/run/eval_prog_fus.py:300-304). proba / pred may be NULL. */
int koa_predict(const float* logits, float* proba, long long* pred, int batch, int classes, void* stream);
/* Fold ensemble of the reference: out = softmax(mean over folds of proba[folds][batch][classes]), pred = argmax(out)
 * (koafusion/run/eval_prog_fus.py:330-336; the softmax is applied to the averaged probabilities, as the reference does). */
int koa_ensemble_proba(const float* proba, float* out, long long* pred, int folds, int batch, int classes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KOA_B200_H */
