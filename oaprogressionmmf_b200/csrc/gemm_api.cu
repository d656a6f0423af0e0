// gemm_api.cu — host launchers + C-ABI entry points for the tcgen05 GEMM / implicit-GEMM kernels.
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/koa_b200.h"
#include "gemm_conv.cuh"
#include "gemm_tc.cuh"
#include "koa_internal.h"
#include "koa_tma.h"

using namespace koa;

static WgradDesc s_wgrad_desc = {8192u, 1024u, 2048u};

// ---- optional profiling of the tcgen05 kernel family -------------------------------------------------
// When enabled every launch is bracketed by CUDA events on its own stream; koa_profile_read() synchronises
// and returns device time, algorithmic FLOPs (2*M*N*K) and launch counts per kernel class
// (0 = fprop/dgrad/linear "kmajor" kernels, 1 = weight-gradient kernels).
namespace {
struct ProfRec { cudaEvent_t a, b; int cls; double flops; int m, n, k, tag; };
bool s_prof_on = false;
std::vector<ProfRec> s_prof;
std::mutex s_prof_mu;

// Algorithmic share of the FLOPs of the next launches of this thread (koa_profile_flop_scale): the engines set it where a
// launch multiplies structural zeros or does bookkeeping work (zero-inserted stride-2 data gradients: 1/4; grouped 3x3 as
// block-diagonal 64-channel chunks: channels per group / 64; stem K = 49 of 64; the Gram / coefficient GEMMs of the y-free
// BatchNorm: 0), so that the roofline counts what the reference computes, not what was launched.
thread_local double t_flop_scale = 1.0;

struct ProfScope {
  cudaStream_t st; int cls; double flops; int m, n, k, tag; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(cudaStream_t st_, int cls_, double flops_, int m_ = 0, int n_ = 0, int k_ = 0, int tag_ = 0)
      : st(st_), cls(cls_), flops(flops_ * t_flop_scale), m(m_), n(n_), k(k_), tag(tag_) {
    if (!s_prof_on) return;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(s_prof_mu);
    s_prof.push_back({a, b, cls, flops, m, n, k, tag});
  }
};
}  // namespace

double koa_profile_flop_scale(double scale) {
  const double old = t_flop_scale;
  t_flop_scale = scale;
  return old;
}

extern "C" int koa_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(s_prof_mu);
  s_prof_on = on != 0;
  for (auto& r : s_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  s_prof.clear();
  return 0;
}
// The recorded launches that have STARTED but not finished, as text lines "cls tag m n k" (cudaEventQuery only: never
// waits). With profiling enabled, this names the kernel(s) a stream is stuck in. Returns the number of such launches.
extern "C" int koa_profile_pending(char* buf, int cap) {
  std::lock_guard<std::mutex> lk(s_prof_mu);
  int n = 0, used = 0;
  if (buf != nullptr && cap > 0) buf[0] = 0;
  for (auto& r : s_prof) {
    if (cudaEventQuery(r.a) != cudaSuccess) continue;
    if (cudaEventQuery(r.b) == cudaSuccess) continue;
    ++n;
    if (buf != nullptr && used < cap - 80) used += snprintf(buf + used, cap - used, "%d %d %d %d %d\n", r.cls, r.tag, r.m, r.n, r.k);
  }
  (void)cudaGetLastError();  // cudaErrorNotReady is not an error here
  return n;
}
// Per-shape breakdown of the recorded launches as text lines "cls tag m n k launches ms tflops" (does not clear).
extern "C" int koa_profile_dump(const char* path) {
  KOA_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(s_prof_mu);
  struct Acc { double ms = 0, flops = 0; long n = 0; };
  std::map<std::tuple<int, int, int, int, int>, Acc> acc;
  for (auto& r : s_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Acc& a = acc[std::make_tuple(r.cls, r.tag, r.m, r.n, r.k)];
    a.ms += ms; a.flops += r.flops; a.n += 1;
  }
  FILE* f = fopen(path, "w");
  KOA_REQUIRE(f != nullptr, "cannot open %s", path);
  fprintf(f, "# cls(0=kmajor,1=wgrad) tag(bit0 im2col, 1 stats, 2 add, 3 gate, 4 res_f32, 5 fp32 out, 6 act, 7 bn-bwd stats) m n k launches total_ms tflops\n");
  for (auto& kv : acc)
    fprintf(f, "%d %d %d %d %d %ld %.4f %.1f\n", std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first),
            std::get<3>(kv.first), std::get<4>(kv.first), kv.second.n, kv.second.ms,
            kv.second.ms > 0 ? kv.second.flops / (kv.second.ms * 1e-3) / 1e12 : 0.0);
  fclose(f);
  return 0;
}
// out[cls*3 + {0,1,2}] = device milliseconds, algorithmic FLOPs, launches for cls in {0,1}; clears the records.
extern "C" int koa_profile_read(double* out) {
  KOA_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(s_prof_mu);
  for (int i = 0; i < 6; ++i) out[i] = 0.0;
  for (auto& r : s_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    out[r.cls * 3 + 0] += ms;
    out[r.cls * 3 + 1] += r.flops;
    out[r.cls * 3 + 2] += 1.0;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  s_prof.clear();
  return 0;
}

extern "C" int koa_version(void) { return 1; }

namespace {
struct DebugWords { const void* flag; const void* first; };
std::vector<DebugWords>& debug_words() {  // function-local: constructed before the first registrar runs
  static std::vector<DebugWords> v;
  return v;
}
}  // namespace
void koa_register_debug_words(const void* flag_symbol, const void* first_symbol) {
  debug_words().push_back({flag_symbol, first_symbol});
}

extern "C" int koa_debug_flag(unsigned int* out) {
  const unsigned int zero = 0;
  *out = 0;
  for (const DebugWords& w : debug_words()) {
    unsigned int v = 0;
    KOA_CHECK_CUDA(cudaMemcpyFromSymbol(&v, w.flag, sizeof(unsigned int)));
    KOA_CHECK_CUDA(cudaMemcpyToSymbol(w.flag, &zero, sizeof(unsigned int)));
    KOA_CHECK_CUDA(cudaMemcpyToSymbol(w.first, &zero, sizeof(unsigned int)));
    *out |= v;
  }
  return 0;
}

// The same words read WITHOUT waiting for the work in flight (a copy on a non-blocking stream of its own) and without
// clearing them: what a watchdog calls while a kernel is still spinning on a barrier.
extern "C" int koa_debug_flag_peek(unsigned int* latest, unsigned int* first) {
  int dev = 0;
  KOA_CHECK_CUDA(cudaGetDevice(&dev));
  // the stream and the pinned landing buffer are created by the FIRST call on a device (make it while the device is idle:
  // creating them can itself wait for running kernels)
  struct Peek { cudaStream_t st; unsigned int* host; };
  static std::mutex mu;
  static std::map<int, Peek> peeks;
  std::lock_guard<std::mutex> lock(mu);
  auto it = peeks.find(dev);
  if (it == peeks.end()) {
    Peek pk{};
    KOA_CHECK_CUDA(cudaStreamCreateWithFlags(&pk.st, cudaStreamNonBlocking));
    KOA_CHECK_CUDA(cudaMallocHost(&pk.host, 2 * sizeof(unsigned int)));
    it = peeks.emplace(dev, pk).first;
  }
  cudaStream_t st = it->second.st;
  unsigned int* host = it->second.host;
  *latest = 0;
  *first = 0;
  int rc = 0;
  for (const DebugWords& w : debug_words()) {
    void *df = nullptr, *d1 = nullptr;
    if (cudaGetSymbolAddress(&df, w.flag) != cudaSuccess || cudaGetSymbolAddress(&d1, w.first) != cudaSuccess ||
        cudaMemcpyAsync(host, df, sizeof(unsigned int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(host + 1, d1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      rc = KOA_ERR_CUDA;
      break;
    }
    *latest |= host[0];
    if (*first == 0) *first = host[1];
  }
  if (rc) koa_set_error("koa_debug_flag_peek: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

extern "C" int koa_debug_set_wgrad_desc(unsigned int lbo, unsigned int sbo, unsigned int k_adv) {
  s_wgrad_desc.lbo = lbo;
  s_wgrad_desc.sbo = sbo;
  s_wgrad_desc.k_adv = k_adv;
  return 0;
}

static EpiParams to_epi(const koa_epilogue_t* e, int n) {
  EpiParams p;
  p.out = e->out;
  p.ldo = e->ldo;
  p.out_fp32 = e->out_fp32;
  p.act = e->act;
  p.bias = e->bias;
  p.pre_out = (bf16*)e->pre_out_bf16;
  p.aux = (const bf16*)e->aux_bf16;
  p.res_f32 = e->residual_f32;
  p.add_bf16 = (const bf16*)e->add_bf16;
  p.gate_bf16 = (const bf16*)e->gate_bf16;
  p.out_bf16_copy = (bf16*)e->out_bf16_copy;
  p.col_sum = e->col_sum;
  p.col_sumsq = e->col_sumsq;
  p.stat_y = (const bf16*)e->stat_y;
  p.stat_mean = e->stat_mean;
  p.stat_invstd = e->stat_invstd;
  p.a_f16 = e->a_f16; p.b_f16 = e->b_f16; p.out_f16 = e->out_f16; p.act_f16 = e->act_f16;
  p.bn_scale = e->bn_scale; p.bn_shift = e->bn_shift; p.res_scale = e->res_scale; p.res_shift = e->res_shift;
  p.col_bias = e->col_bias;
  p.drop_on = e->drop_p > 0.0f ? 1 : 0;
  p.drop_cols = n;
  p.drop = make_drop_spec(e->drop_seed, e->drop_site, e->drop_p);
  return p;
}

static int check_epi(const koa_epilogue_t* ep, int n) {
  KOA_REQUIRE(ep != nullptr && ep->out != nullptr, "epilogue/out must not be NULL");
  KOA_REQUIRE(n % 32 == 0, "tcgen05 GEMM path needs N %% 32 == 0 (got %d)", n);
  KOA_REQUIRE(ep->ldo >= n && ep->ldo % 8 == 0, "ldo (%d) must be >= N and a multiple of 8", ep->ldo);
  KOA_REQUIRE(ep->act != KOA_ACT_GELU_GRAD || ep->aux_bf16 != nullptr, "KOA_ACT_GELU_GRAD needs aux_bf16");
  KOA_REQUIRE((ep->col_sum == nullptr) == (ep->col_sumsq == nullptr), "col_sum and col_sumsq go together");
  KOA_REQUIRE(ep->col_sum == nullptr || n <= kMaxStatCols, "column statistics support N <= %d (got %d)", kMaxStatCols, n);
  KOA_REQUIRE(ep->col_sum == nullptr || !ep->out_fp32, "column statistics need a bf16 output");
  KOA_REQUIRE((ep->a_f16 != 0) == (ep->b_f16 != 0), "tcgen05 kind::f16 needs both operands in the same 16-bit format");
  KOA_REQUIRE(ep->drop_p >= 0.0f && ep->drop_p < 1.0f, "dropout probability %f out of range", (double)ep->drop_p);
  KOA_REQUIRE(ep->stat_y == nullptr || (ep->col_sum != nullptr && ep->stat_mean != nullptr && ep->stat_invstd != nullptr),
              "stat_y needs col_sum/col_sumsq and stat_mean/stat_invstd");
  KOA_REQUIRE((ep->bn_scale == nullptr) == (ep->bn_shift == nullptr), "bn_scale and bn_shift go together");
  KOA_REQUIRE((ep->res_scale == nullptr) == (ep->res_shift == nullptr), "res_scale and res_shift go together");
  if (ep->bn_scale != nullptr) {
    KOA_REQUIRE(ep->out_f16 && !ep->out_fp32, "the BatchNorm-apply epilogue stores fp16 (out_f16 = 1)");
    KOA_REQUIRE(ep->act == KOA_ACT_NONE || ep->act == KOA_ACT_RELU, "the BatchNorm-apply epilogue takes no activation but ReLU");
    KOA_REQUIRE(ep->add_bf16 == nullptr || ep->act_f16, "the residual of the BatchNorm-apply epilogue is an fp16 activation");
    KOA_REQUIRE(ep->res_scale == nullptr || ep->add_bf16 != nullptr, "res_scale needs the residual (add_bf16)");
    KOA_REQUIRE(ep->gate_bf16 == nullptr && ep->col_sum == nullptr && ep->col_bias == nullptr && ep->bias == nullptr &&
                    ep->pre_out_bf16 == nullptr && ep->residual_f32 == nullptr && ep->drop_p == 0.0f,
                "the BatchNorm-apply epilogue combines with add_bf16 / act / out_bf16_copy only");
  } else {
    KOA_REQUIRE(ep->res_scale == nullptr, "res_scale needs bn_scale");
  }
  return 0;
}

// The convolution flavour of the epilogue (bf16 output, optional addend / gate / statistics) is a separate, leaner
// instantiation; everything else (bias, activations, fp32 residual stream) takes the full one.
static bool conv_epilogue(const EpiParams& ep) {
  if (ep.bn_scale != nullptr) return true;  // BatchNorm-apply flavour (MODE 2): act / out_bf16_copy belong to it (check_epi)
  return !ep.out_fp32 && ep.act == ACT_NONE && ep.bias == nullptr && ep.pre_out == nullptr && ep.res_f32 == nullptr &&
         ep.out_bf16_copy == nullptr && !ep.drop_on;
}
// features only gemm_conv_kernel implements: there is no fallback kernel for them
static bool conv_only(const EpiParams& ep) { return ep.bn_scale != nullptr || ep.col_bias != nullptr; }

// second A tensor of a K-concatenated GEMM (k-blocks >= kb_split read it); kb_split < 0: none
struct AExt {
  const CUtensorMap* ta2 = nullptr;
  int kb_split = -1;
};

static int prof_flavor(bool im2col, const EpiParams& ep) {
  return (im2col ? 1 : 0) | (ep.col_sum ? 2 : 0) | (ep.add_bf16 ? 4 : 0) | (ep.gate_bf16 ? 8 : 0) | (ep.res_f32 ? 16 : 0) |
         (ep.out_fp32 ? 32 : 0) | (ep.act ? 64 : 0) | (ep.stat_y ? 128 : 0) | (ep.bn_scale ? 512 : 0) |
         (ep.col_bias ? 1024 : 0);
}

template <int BN, int STAGES, bool IM2COL, bool CONV>
static int launch_kmajor_t(const CUtensorMap& ta, const CUtensorMap& tb, int m, int n, int k, const ConvGeom& g,
                           const EpiParams& ep, cudaStream_t st) {
  constexpr size_t smem = gemm_smem_bytes<BN, STAGES>();
  static std::atomic<unsigned long long> attr_done{0};
  KOA_CHECK_CUDA(koa_ensure_dyn_smem(gemm_kmajor_kernel<BN, STAGES, IM2COL, CONV>, (int)smem, attr_done));
  const long long tiles = (long long)koa_cdiv(m, BM) * koa_cdiv(n, BN);
  KOA_REQUIRE(tiles > 0 && tiles < 2147483647LL, "bad tile count");
  const unsigned grid = (unsigned)(tiles < koa_num_sms() ? tiles : koa_num_sms());  // persistent: one CTA per SM
  {
    ProfScope prof(st, 0, 2.0 * (double)m * (double)n * (double)k, m, n, k, prof_flavor(IM2COL, ep));
    gemm_kmajor_kernel<BN, STAGES, IM2COL, CONV><<<grid, kKmajorThreads, smem, st>>>(ta, tb, m, n, k, g, ep);
  }
  KOA_LAUNCH_CHECK();
  return 0;
}

// Convolution flavour: 16 epilogue warps, one n-tile per CTA (gemm_conv.cuh). grid = a multiple of the n-tile count.
// CTA2: CTA pairs (cluster of 2, tcgen05 cta_group::2) on 256-row tiles; `tb` then has a box of BN / 2 rows.
template <int BN, int STAGES, bool IM2COL, int MODE, bool OF16, bool AF16, bool CTA2>
static int launch_conv_t(const CUtensorMap& ta, const CUtensorMap& tb, int m, int n, int k, const ConvGeom& g,
                         const EpiParams& ep, cudaStream_t st, const AExt& ax) {
  constexpr size_t smem = conv_smem_bytes<BN, STAGES, MODE, CTA2>();
  static_assert(smem <= 232448, "shared memory budget of one CTA per SM");
  auto kern = gemm_conv_kernel<BN, STAGES, IM2COL, MODE, OF16, AF16, CTA2>;
  static std::atomic<unsigned long long> attr_done{0};
  KOA_CHECK_CUDA(koa_ensure_dyn_smem(kern, (int)smem, attr_done));
  const int n_tiles = koa_cdiv(n, BN);
  const long long tiles = (long long)koa_cdiv(m, CTA2 ? 2 * BM : BM) * n_tiles;
  KOA_REQUIRE(tiles > 0 && tiles < 2147483647LL, "bad tile count");
  const long long units_max = CTA2 ? koa_num_sms() / 2 : koa_num_sms();
  const long long cap = tiles < units_max ? tiles : units_max;
  const unsigned units = (unsigned)(cap / n_tiles * n_tiles);
  KOA_REQUIRE(units > 0, "more n-tiles than SMs");
  // MODE 1 moves its epilogue operands and its output with the TMA unit: [M, N] views with row pitch ldo
  CUtensorMap t_out = ta, t_add = ta, t_gate = ta, t_y = ta, t_out2 = ta;
  {
    const uint64_t pitch = (uint64_t)ep.ldo * 2;
    int rc = koa_tmap_2d_sw64(&t_out, ep.out, (uint64_t)n, (uint64_t)m, pitch);
    if (rc) return rc;
  }
  if (MODE >= 1) {
    const uint64_t pitch = (uint64_t)ep.ldo * 2;
    int rc;
    t_add = t_gate = t_y = t_out2 = t_out;
    if (ep.add_bf16 && (rc = koa_tmap_2d_sw64(&t_add, ep.add_bf16, (uint64_t)n, (uint64_t)m, pitch))) return rc;
    if (MODE == 1 && ep.gate_bf16 && (rc = koa_tmap_2d_sw64(&t_gate, ep.gate_bf16, (uint64_t)n, (uint64_t)m, pitch))) return rc;
    if (MODE == 1 && ep.stat_y && (rc = koa_tmap_2d_sw64(&t_y, ep.stat_y, (uint64_t)n, (uint64_t)m, pitch))) return rc;
    if (MODE == 2 && ep.out_bf16_copy && (rc = koa_tmap_2d_sw64(&t_out2, ep.out_bf16_copy, (uint64_t)n, (uint64_t)m, pitch)))
      return rc;
  }
  KOA_REQUIRE(ax.kb_split < 0 || (MODE == 1 && !IM2COL), "K concatenation: backward flavour of a plain GEMM only");
  const CUtensorMap& ta2 = ax.kb_split >= 0 ? *ax.ta2 : ta;
  const int kb_split = ax.kb_split >= 0 ? ax.kb_split : 0x7fffffff;
  {
    ProfScope prof(st, 0, 2.0 * (double)m * (double)n * (double)k, m, n, k, prof_flavor(IM2COL, ep) | (CTA2 ? 256 : 0));
    if (CTA2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * units, 1, 1);
      cfg.blockDim = dim3(kConvThreads, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      KOA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, t_out, t_add, t_gate, t_y, ta2, t_out2, m, n, k, kb_split, g, ep));
    } else {
      kern<<<units, kConvThreads, smem, st>>>(ta, tb, t_out, t_add, t_gate, t_y, ta2, t_out2, m, n, k, kb_split, g, ep);
    }
  }
  KOA_LAUNCH_CHECK();
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int conv_mode1_enabled() { static const int v = env_int("KOA_CONV_MODE1", 1); return v; }
static int conv_max_ntiles() { static const int v = env_int("KOA_CONV_MAX_NT", 16); return v; }
// CTA pairs for the compute-bound layers: K at least this (0 disables)
static int conv_cta2_min_k() { static const int v = env_int("KOA_CONV_CTA2_MINK", 512); return v; }

// mode / format dispatch of the convolution flavour; returns 1 when no instantiation matches (caller falls back)
template <int BN, int ST0, int ST1, bool IM2COL, bool CTA2>
static int launch_conv_fmt(const CUtensorMap& ta, const CUtensorMap& tb, int m, int n, int k, const ConvGeom& g,
                           const EpiParams& ep, cudaStream_t st, const AExt& ax) {
  if (ep.bn_scale != nullptr)  // BatchNorm apply (+ residual) + ReLU, fp16 activations (check_epi)
    return launch_conv_t<BN, ST1, IM2COL, 2, true, true, CTA2>(ta, tb, m, n, k, g, ep, st, ax);
  const bool mode1 = ep.add_bf16 != nullptr || ep.gate_bf16 != nullptr || ep.stat_y != nullptr || ep.col_bias != nullptr ||
                     ax.kb_split >= 0;
  if (!mode1) {
    if (ep.out_f16) return launch_conv_t<BN, ST0, IM2COL, 0, true, false, CTA2>(ta, tb, m, n, k, g, ep, st, ax);
    return launch_conv_t<BN, ST0, IM2COL, 0, false, false, CTA2>(ta, tb, m, n, k, g, ep, st, ax);
  }
  if (!ep.out_f16 && conv_mode1_enabled()) {
    if (ep.act_f16) return launch_conv_t<BN, ST1, IM2COL, 1, false, true, CTA2>(ta, tb, m, n, k, g, ep, st, ax);
    return launch_conv_t<BN, ST1, IM2COL, 1, false, false, CTA2>(ta, tb, m, n, k, g, ep, st, ax);
  }
  return 1;
}

template <int BN, int STAGES, bool IM2COL>
static int launch_kmajor(const CUtensorMap& ta, const CUtensorMap& tb, int m, int n, int k, const ConvGeom& g,
                         const EpiParams& ep, cudaStream_t st, const AExt& ax) {
  const bool only = conv_only(ep) || ax.kb_split >= 0;
  if (conv_epilogue(ep)) {
    const int n_tiles = koa_cdiv(n, BN);
    if ((n_tiles <= conv_max_ntiles() || only) && n_tiles <= koa_num_sms()) {
      const int rc = launch_conv_fmt<BN, (BN == 128 ? 5 : 6), (BN == 128 ? 4 : 5), IM2COL, false>(ta, tb, m, n, k, g, ep, st, ax);
      if (rc != 1) return rc;
    }
    KOA_REQUIRE(!only, "BatchNorm-apply / col_bias / K-concatenated epilogues: no kernel for %d x %d x %d", m, n, k);
    return launch_kmajor_t<BN, STAGES, IM2COL, true>(ta, tb, m, n, k, g, ep, st);
  }
  KOA_REQUIRE(!only, "col_bias / K concatenation need a convolution-flavour epilogue");
  return launch_kmajor_t<BN, STAGES, IM2COL, false>(ta, tb, m, n, k, g, ep, st);
}

template <bool IM2COL>
static int dispatch_kmajor(const CUtensorMap& ta, const void* b, int m, int n, int k, const ConvGeom& g,
                           const EpiParams& ep, cudaStream_t st, const AExt& ax = AExt{}) {
  const bool bn128 = (n % 128 == 0);
  CUtensorMap tb;
  int rc;
  // compute-bound convolution layers: CTA pairs on 256 x BN tiles, each CTA stages BN / 2 rows of B
  if (bn128 && !g.grouped && conv_epilogue(ep) && conv_cta2_min_k() > 0 && k >= conv_cta2_min_k() && koa_num_sms() >= 2) {
    // (pairs on 256 x 128 tiles, N = 128, measured slower than single CTAs: 430 vs 480 TFLOP/s at K = 512)
    if (n % 256 == 0 && n / 256 <= koa_num_sms() / 2) {
      rc = koa_tmap_2d_bf16(&tb, b, (uint64_t)k, (uint64_t)n, (uint64_t)k * 2, 64, 128);
      if (rc) return rc;
      rc = launch_conv_fmt<256, 5, 4, IM2COL, true>(ta, tb, m, n, k, g, ep, st, ax);
      if (rc != 1) return rc;
    }
  }
  rc = koa_tmap_2d_bf16(&tb, b, (uint64_t)k, (uint64_t)n, (uint64_t)k * 2, 64, bn128 ? 128 : 64);
  if (rc) return rc;
  // persistent kernel, one CTA per SM: a deep smem ring lets the TMA producer run ahead across tiles
  if (bn128) return launch_kmajor<128, 5, IM2COL>(ta, tb, m, n, k, g, ep, st, ax);
  return launch_kmajor<64, 6, IM2COL>(ta, tb, m, n, k, g, ep, st, ax);
}

int koa_gemm_launch(const void* a, const void* b, int m, int n, int k, const koa_epilogue_t* ep, cudaStream_t st) {
  int rc = check_epi(ep, n);
  if (rc) return rc;
  KOA_REQUIRE(m > 0 && n > 0 && k > 0, "empty GEMM %dx%dx%d", m, n, k);
  KOA_REQUIRE(k % 8 == 0, "GEMM K (%d) must be a multiple of 8 (16-byte TMA pitch)", k);
  CUtensorMap ta;
  rc = koa_tmap_2d_bf16(&ta, a, (uint64_t)k, (uint64_t)m, (uint64_t)k * 2, 64, 128);
  if (rc) return rc;
  ConvGeom g = {1, 1, 1, 0, 1, 1, 0};
  return dispatch_kmajor<false>(ta, b, m, n, k, g, to_epi(ep, n), st);
}

// out = epilogue([A1 | A2] . B^T): the k-blocks of A2 follow those of A1 (B has K = k1 + k2 columns)
int koa_gemm_kcat_launch(const void* a1, int k1, const void* a2, int k2, const void* b, int m, int n,
                         const koa_epilogue_t* ep, cudaStream_t st) {
  int rc = check_epi(ep, n);
  if (rc) return rc;
  KOA_REQUIRE(m > 0 && n > 0 && k1 > 0 && k2 > 0, "empty GEMM %dx%dx(%d+%d)", m, n, k1, k2);
  KOA_REQUIRE(k1 % BK == 0 && k2 % BK == 0, "K concatenation needs K1, K2 multiples of %d (got %d, %d)", BK, k1, k2);
  CUtensorMap ta, ta2;
  rc = koa_tmap_2d_bf16(&ta, a1, (uint64_t)k1, (uint64_t)m, (uint64_t)k1 * 2, 64, 128);
  if (rc) return rc;
  rc = koa_tmap_2d_bf16(&ta2, a2, (uint64_t)k2, (uint64_t)m, (uint64_t)k2 * 2, 64, 128);
  if (rc) return rc;
  ConvGeom g = {1, 1, 1, 0, 1, 1, 0};
  AExt ax;
  ax.ta2 = &ta2;
  ax.kb_split = k1 / BK;
  return dispatch_kmajor<false>(ta, b, m, n, k1 + k2, g, to_epi(ep, n), st, ax);
}

int koa_conv_fprop_launch(const void* x, const void* w, int n_img, int h, int w_in, int cin, int cout, int filt_r,
                          int filt_s, int stride, int pad, const koa_epilogue_t* ep, cudaStream_t st) {
  int rc = check_epi(ep, cout);
  if (rc) return rc;
  KOA_REQUIRE(cin % 64 == 0, "implicit-GEMM conv needs Cin %% 64 == 0 (got %d)", cin);
  KOA_REQUIRE(stride >= 1 && pad >= 0 && filt_r >= 1 && filt_s >= 1, "bad conv geometry");
  const int hout = (h + 2 * pad - filt_r) / stride + 1;
  const int wout = (w_in + 2 * pad - filt_s) / stride + 1;
  KOA_REQUIRE(hout > 0 && wout > 0, "empty conv output");
  const long long m = (long long)n_img * hout * wout;
  KOA_REQUIRE(m < 2147483647LL, "too many output pixels");
  CUtensorMap ta;
  rc = koa_tmap_im2col_bf16(&ta, x, n_img, h, w_in, cin, filt_r, filt_s, stride, pad, 128);
  if (rc) return rc;
  ConvGeom g = {hout, wout, stride, pad, filt_s, cin / 64, 0};
  return dispatch_kmajor<true>(ta, w, (int)m, cout, filt_r * filt_s * cin, g, to_epi(ep, cout), st);
}

// x_f16: 0 = dY and X bf16; 1 = both fp16; 2 = dY bf16, X fp16 converted to bf16 inside the kernel (XCVT). (A mixed
// bf16 x fp16 MMA, one format field per operand in the instruction descriptor, faults with an illegal instruction on
// B200: measured in round 2, removed.)
// CTA2: CTA pairs on 256 x BN tiles (the tensor map of X then has boxes of 64 columns as always; each CTA loads BN / 2)
template <int BN, int STAGES, bool IM2COL, bool XCVT = false, bool CTA2 = false>
static int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tb, int cout, int cin, int pixels, int taps,
                        const ConvGeom& g, float* dw, int x_f16, cudaStream_t st) {
  if (!XCVT && !CTA2 && x_f16 == 2) return launch_wgrad<BN, STAGES, IM2COL, true, false>(ta, tb, cout, cin, pixels, taps, g, dw, 0, st);
  const int a_f16 = x_f16, b_f16 = x_f16;  // dY, X
  // (a CTA pair allocates tensor memory collectively: it must not share its SMs with other allocating CTAs, see wgrad_cta2)
  constexpr size_t smem = CTA2 ? (size_t)227 * 1024 : wgrad_smem_bytes<BN, STAGES>();
  static_assert(wgrad_smem_bytes<CTA2 ? BN / 2 : BN, STAGES>() <= smem, "shared memory of the weight-gradient kernel");
  auto kern = gemm_wgrad_kernel<BN, STAGES, IM2COL, XCVT, CTA2>;
  static std::atomic<unsigned long long> attr_done{0};
  KOA_CHECK_CUDA(koa_ensure_dyn_smem(kern, (int)smem, attr_done));
  const int tiles = (g.grouped ? 1 : koa_cdiv(cout, CTA2 ? 2 * BM : BM)) * koa_cdiv(cin, BN) * taps;
  const int num_kb = koa_cdiv(pixels, BK);
  // Split the pixel (reduction) range so that the grid covers the machine a few times over.
  static const int waves = [] {
    const char* e = getenv("KOA_WGRAD_WAVES");
    return e ? atoi(e) : 2;  // measured on the full step: 2 waves 103.3 ms, 4 waves 103.7 ms, 6 waves 105.2 ms (serial pass)
  }();
  int splits = koa_cdiv(waves * koa_num_sms(), tiles * (CTA2 ? 2 : 1));
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  int kb_per_split = koa_cdiv(num_kb, splits);
  if (kb_per_split < 4 && num_kb >= 4) kb_per_split = 4;
  splits = koa_cdiv(num_kb, kb_per_split);
  dim3 grid((unsigned)tiles * (CTA2 ? 2 : 1), (unsigned)splits);
  const unsigned threads = XCVT ? kWgradCvtThreads : kGemmThreads;
  {
    // grouped: only the diagonal 64x64 blocks are algorithmic work
    const double n_eff = g.grouped ? 64.0 : (double)cin;
    ProfScope prof(st, 1, 2.0 * (double)pixels * (double)cout * n_eff * (double)taps, cout, cin * taps, pixels,
                   (IM2COL ? 1 : 0) | (CTA2 ? 256 : 0));
    if (CTA2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(threads, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      KOA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, cout, cin, pixels, taps, g, dw, kb_per_split, s_wgrad_desc, a_f16, b_f16));
    } else {
      kern<<<grid, threads, smem, st>>>(ta, tb, cout, cin, pixels, taps, g, dw, kb_per_split, s_wgrad_desc, a_f16, b_f16);
    }
  }
  KOA_LAUNCH_CHECK();
  return 0;
}

// CTA pairs for weight gradients. KOA_WGRAD_CTA2: 0 (default) never, 1 the one shape where a pair measured faster on B200
// (3x3, 256 -> 256 channels: 79.9 -> 68.6 us, tools/prof_wgrad.py), 2 wherever the shape allows (Cout and Cin multiples of
// 256; slower: 3x3 with 512 channels 80.9 -> 85.0 us, 1x1 layers 50 -> 61 us, the transformer Linears 34-46 -> 39-69 us -
// with half as many, twice as large tiles the split-K count doubles and the kernel becomes bound by its fp32 atomics).
// OFF by default since round 2: with 97 KB of shared memory two CTAs of DIFFERENT pairs (and CTAs of the other
// weight-gradient kernels) shared an SM, and about once per 500 training steps a step never finished - GPU at 100 %, no
// mbarrier time-out code, i.e. a warp sitting in a blocking instruction; the collective tcgen05.alloc.cta_group::2 of two
// pairs that each hold one SM's allocation permit is the suspect. 2500 steps with the pair kernel off, and every other
// cta_group::2 kernel of the library (206 KB: alone on its SM), never showed it (tools/r2_hunt.sh). When the
// switch is on the pair kernel now asks for the whole shared memory of its SM, so that it never shares one.
static bool wgrad_cta2(int cout, int cin, int x_f16, bool conv3x3) {
  static const int mode = env_int("KOA_WGRAD_CTA2", 0);
  if (mode == 0 || x_f16 == 2 || koa_num_sms() < 2 || cout % 256 != 0 || cin % 256 != 0) return false;
  return mode >= 2 || (conv3x3 && cout == 256 && cin == 256);
}

int koa_gemm_wgrad_launch(const void* dy, const void* x, float* dw, int pixels, int cout, int cin, int x_f16, cudaStream_t st) {
  KOA_REQUIRE(pixels > 0 && cout > 0 && cin > 0, "empty wgrad");
  KOA_REQUIRE(cout % 8 == 0 && cin % 64 == 0, "wgrad needs Cout %% 8 == 0 and Cin %% 64 == 0 (got %d, %d)", cout, cin);
  CUtensorMap ta, tb;
  int rc = koa_tmap_2d_bf16(&ta, dy, (uint64_t)cout, (uint64_t)pixels, (uint64_t)cout * 2, 64, 64);
  if (rc) return rc;
  rc = koa_tmap_2d_bf16(&tb, x, (uint64_t)cin, (uint64_t)pixels, (uint64_t)cin * 2, 64, 64);
  if (rc) return rc;
  ConvGeom g = {1, 1, 1, 0, 1, 1, 0};
  if (wgrad_cta2(cout, cin, x_f16, false)) return launch_wgrad<256, 3, false, false, true>(ta, tb, cout, cin, pixels, 1, g, dw, x_f16, st);
  if (cin % 128 == 0) return launch_wgrad<128, 3, false>(ta, tb, cout, cin, pixels, 1, g, dw, x_f16, st);
  return launch_wgrad<64, 3, false>(ta, tb, cout, cin, pixels, 1, g, dw, x_f16, st);
}

int koa_conv_wgrad_launch(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int cin, int cout,
                          int filt_r, int filt_s, int stride, int pad, int x_f16, cudaStream_t st) {
  KOA_REQUIRE(cout % 8 == 0 && cin % 64 == 0, "wgrad needs Cout %% 8 == 0 and Cin %% 64 == 0 (got %d, %d)", cout, cin);
  const int hout = (h + 2 * pad - filt_r) / stride + 1;
  const int wout = (w_in + 2 * pad - filt_s) / stride + 1;
  const long long pixels = (long long)n_img * hout * wout;
  KOA_REQUIRE(pixels > 0 && pixels < 2147483647LL, "bad pixel count");
  CUtensorMap ta, tb;
  int rc = koa_tmap_2d_bf16(&ta, dy, (uint64_t)cout, (uint64_t)pixels, (uint64_t)cout * 2, 64, 64);
  if (rc) return rc;
  rc = koa_tmap_im2col_bf16(&tb, x, n_img, h, w_in, cin, filt_r, filt_s, stride, pad, 64);
  if (rc) return rc;
  ConvGeom g = {hout, wout, stride, pad, filt_s, cin / 64, 0};
  const int taps = filt_r * filt_s;
  if (wgrad_cta2(cout, cin, x_f16, taps == 9)) return launch_wgrad<256, 3, true, false, true>(ta, tb, cout, cin, (int)pixels, taps, g, dw, x_f16, st);
  if (cin % 128 == 0) return launch_wgrad<128, 3, true>(ta, tb, cout, cin, (int)pixels, taps, g, dw, x_f16, st);
  return launch_wgrad<64, 3, true>(ta, tb, cout, cin, (int)pixels, taps, g, dw, x_f16, st);
}

// Grouped 3x3 convolution (ResNeXt, koafusion/models/_torchvision.py:110,327-328) as a block-diagonal dense
// convolution per 64-channel chunk: w is [C][3][3][64] (koa_k_pack_grouped_w), K = 576 per output chunk.
int koa_conv_grouped_launch(const void* x, const void* w, int n_img, int h, int w_in, int c, int stride,
                            const koa_epilogue_t* ep, cudaStream_t st) {
  int rc = check_epi(ep, c);
  if (rc) return rc;
  KOA_REQUIRE(c % 64 == 0, "grouped conv needs C %% 64 == 0 (got %d)", c);
  const int hout = (h + 2 - 3) / stride + 1, wout = (w_in + 2 - 3) / stride + 1;
  const long long m = (long long)n_img * hout * wout;
  KOA_REQUIRE(m > 0 && m < 2147483647LL, "bad pixel count");
  CUtensorMap ta, tb;
  rc = koa_tmap_im2col_bf16(&ta, x, n_img, h, w_in, c, 3, 3, stride, 1, 128);
  if (rc) return rc;
  rc = koa_tmap_2d_bf16(&tb, w, 576, (uint64_t)c, 576 * 2, 64, 64);
  if (rc) return rc;
  ConvGeom g = {hout, wout, stride, 1, 3, 1, 1};
  return launch_kmajor<64, 6, true>(ta, tb, (int)m, c, 576, g, to_epi(ep, c), st, AExt{});
}

// dw[C][9][64] += per-chunk dense weight gradient of the grouped convolution.
int koa_conv_grouped_wgrad_launch(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int c,
                                  int stride, int x_f16, cudaStream_t st) {
  KOA_REQUIRE(c % 64 == 0, "grouped conv needs C %% 64 == 0 (got %d)", c);
  const int hout = (h + 2 - 3) / stride + 1, wout = (w_in + 2 - 3) / stride + 1;
  const long long pixels = (long long)n_img * hout * wout;
  KOA_REQUIRE(pixels > 0 && pixels < 2147483647LL, "bad pixel count");
  CUtensorMap ta, tb;
  int rc = koa_tmap_2d_bf16(&ta, dy, (uint64_t)c, (uint64_t)pixels, (uint64_t)c * 2, 64, 64);
  if (rc) return rc;
  rc = koa_tmap_im2col_bf16(&tb, x, n_img, h, w_in, c, 3, 3, stride, 1, 64);
  if (rc) return rc;
  ConvGeom g = {hout, wout, stride, 1, 3, c / 64, 1};
  return launch_wgrad<64, 3, true>(ta, tb, c, c, (int)pixels, 9, g, dw, x_f16, st);
}

// ------------------------------------ C ABI ----------------------------------------------------
extern "C" int koa_gemm_bf16(const void* a, const void* b, int m, int n, int k, const koa_epilogue_t* ep,
                             void* stream) {
  return koa_gemm_launch(a, b, m, n, k, ep, (cudaStream_t)stream);
}
extern "C" int koa_gemm_kcat_bf16(const void* a1, int k1, const void* a2, int k2, const void* b, int m, int n,
                                  const koa_epilogue_t* ep, void* stream) {
  return koa_gemm_kcat_launch(a1, k1, a2, k2, b, m, n, ep, (cudaStream_t)stream);
}
extern "C" int koa_conv_fprop_bf16(const void* x, const void* w, int n_img, int h, int w_in, int cin, int cout,
                                   int filt_r, int filt_s, int stride, int pad, const koa_epilogue_t* ep,
                                   void* stream) {
  return koa_conv_fprop_launch(x, w, n_img, h, w_in, cin, cout, filt_r, filt_s, stride, pad, ep, (cudaStream_t)stream);
}
extern "C" int koa_gemm_wgrad_bf16(const void* dy, const void* x, float* dw, int pixels, int cout, int cin, int x_f16,
                                   void* stream) {
  return koa_gemm_wgrad_launch(dy, x, dw, pixels, cout, cin, x_f16, (cudaStream_t)stream);
}
extern "C" int koa_conv_wgrad_bf16(const void* dy, const void* x, float* dw, int n_img, int h, int w_in, int cin,
                                   int cout, int filt_r, int filt_s, int stride, int pad, int x_f16, void* stream) {
  return koa_conv_wgrad_launch(dy, x, dw, n_img, h, w_in, cin, cout, filt_r, filt_s, stride, pad, x_f16, (cudaStream_t)stream);
}
