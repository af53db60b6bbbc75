// C-ABI entry points of the ADN-SSD mixer (include/adnb200.h).  Validates the shape, carves the caller's
// buffers and enqueues the kernels on the caller's stream.  No allocation, no host sync, no CPU fallback.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "adn_common.cuh"
#include "adnssd_generic.cuh"
#include "adnssd_sm100.cuh"
#include "adnssd_wide.cuh"

namespace adn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = -1;
  if (n < 0) {
    int dev = 0, v = 0;
    n = (cudaGetDevice(&dev) == cudaSuccess &&
         cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
    cudaGetLastError();
    // ADN_SM_RESERVE=k: size every persistent grid for (SMs - k).  Data-parallel runs set 1: the persistent kernels hold one CTA
    // per SM with most of its shared memory, so a concurrently running NCCL all-reduce CTA would otherwise displace one of
    // them and add a whole extra wave to that kernel (measured as the weak-scaling loss of round 1: 0.281 -> 0.337 ms at N = 8).
    const char* e = getenv("ADN_SM_RESERVE");
    if (e) { const int k = atoi(e); if (k > 0 && k < n) n -= k; }
  }
  return n;
}

static EnvCfg& env_mut() {
  static EnvCfg cfg = [] {
    auto flag = [](const char* name, bool dflt) { const char* e = getenv(name); return e ? e[0] != '0' : dflt; };
    auto num = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
    EnvCfg c;
    c.rows_per_cta = num("ADN_ROWS_PER_CTA");
    c.rowconv = flag("ADN_ROWCONV", true);
    c.row_wide = flag("ADN_ROW_WIDE", true);
    c.bwd_ws = flag("ADN_BWD_WS", true);
    c.du_dbg = num("ADN_DU_DBG");
    c.wide = flag("ADN_WIDE", true);
    c.gemm_dbg = num("ADN_GEMM_DBG");
    c.variant = num("ADN_VARIANT");
    return c;
  }();
  return cfg;
}
const EnvCfg& env() { return env_mut(); }

// ---- diagnostics: launch counter + optional per-launch CUDA-event timing
constexpr int PROF_CAP = 4096;
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static ProfRec g_prof[PROF_CAP];
static int g_prof_n = 0, g_prof_events = 0;
static bool g_prof_on = false;
static unsigned long long g_launches = 0;

ProfScope::ProfScope(const char* name, cudaStream_t stream) : slot(-1), st(stream) {
  __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED);
  if (g_prof_on && g_prof_n < PROF_CAP) {
    slot = g_prof_n++;
    if (slot >= g_prof_events) {
      cudaEventCreate(&g_prof[slot].e0);
      cudaEventCreate(&g_prof[slot].e1);
      g_prof_events = slot + 1;
    }
    g_prof[slot].name = name;
    cudaEventRecord(g_prof[slot].e0, st);
  }
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, st);
}

static int validate(const AdnShape* s) {
  ADN_REQUIRE(s != nullptr, ADN_ERR_NULL, "AdnShape is NULL");
  ADN_REQUIRE(s->B > 0 && s->H > 0 && s->W > 0 && s->D > 0, ADN_ERR_SHAPE, "B/H/W/D must be positive (got %d,%d,%d,%d)",
              s->B, s->H, s->W, s->D);
  ADN_REQUIRE(s->G == 2, ADN_ERR_SHAPE, "only ngroups == 2 is supported (the branch models/ADNssd.py:278 takes); got %d",
              s->G);
  ADN_REQUIRE(s->Di > 0 && s->Di % 4 == 0, ADN_ERR_SHAPE, "d_inner must be a positive multiple of 4 (got %d)", s->Di);
  ADN_REQUIRE(s->P > 0 && s->Di % s->P == 0, ADN_ERR_SHAPE, "d_inner %% headdim != 0 (%d, %d)", s->Di, s->P);
  // the even/odd channel split gives each parity Di/2 channels grouped into heads of P (models/ADNssd.py:371-386: the
  // reference's rearrange 'b l (h p)' raises otherwise); an odd head count would index dt_bias / A_log / D out of bounds
  ADN_REQUIRE((s->Di / 2) % s->P == 0 && (s->Di / s->P) % 2 == 0, ADN_ERR_SHAPE,
              "(d_inner / 2) %% headdim != 0: the parity split needs an even number of heads (d_inner %d, headdim %d)", s->Di, s->P);
  ADN_REQUIRE(s->N > 0 && (s->G * s->N) % 4 == 0, ADN_ERR_SHAPE, "ngroups*d_state must be a multiple of 4 (got %d)",
              s->G * s->N);
  ADN_REQUIRE(s->dtype == ADN_F32 || s->dtype == ADN_BF16, ADN_ERR_DTYPE, "unsupported dtype %d", s->dtype);
  ADN_REQUIRE(s->flags == 0, ADN_ERR_SHAPE, "flags must be 0");
  ADN_REQUIRE((long long)s->B * s->H * s->W < (1LL << 31) / 4, ADN_ERR_SHAPE, "too many tokens");
  return ADN_OK;
}

static int check_weights(const AdnWeights* w) {
  ADN_REQUIRE(w != nullptr, ADN_ERR_NULL, "AdnWeights is NULL");
  const void* req[] = {w->dt_bias, w->A_log, w->D, w->alpha1, w->in_proj_w, w->conv_13_x1_w, w->conv_31_x1_w,
                       w->conv_13_x2_w, w->conv_31_x2_w, w->conv_13_bc1_w, w->conv_31_bc1_w, w->conv_13_bc2_w,
                       w->conv_31_bc2_w, w->conv2d_w, w->norm_w, w->norm_b, w->conv2d_z_w, w->out_proj_w};
  for (size_t i = 0; i < sizeof(req) / sizeof(req[0]); ++i)
    ADN_REQUIRE(req[i] != nullptr, ADN_ERR_NULL, "AdnWeights: required tensor #%zu is NULL", i);
  return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" {

const char* adn_last_error(void) { return adn::g_err; }
int adn_abi_version(void) { return ADNB200_ABI_VERSION; }

unsigned long long adn_launch_count(void) { return adn::g_launches; }
int adn_prof_enable(int on) {
  adn::g_prof_on = on != 0;
  adn::g_prof_n = 0;
  return ADN_OK;
}
int adn_prof_count(void) { return adn::g_prof_n; }
int adn_prof_get(int i, const char** name, float* ms) {
  ADN_REQUIRE(i >= 0 && i < adn::g_prof_n && name && ms, ADN_ERR_SHAPE, "adn_prof_get: bad index %d", i);
  ADN_CHECK_CUDA(cudaEventSynchronize(adn::g_prof[i].e1));
  ADN_CHECK_CUDA(cudaEventElapsedTime(ms, adn::g_prof[i].e0, adn::g_prof[i].e1));
  *name = adn::g_prof[i].name;
  return ADN_OK;
}

int adn_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int adnssd_workspace_bytes(const AdnShape* s, size_t* saved_bytes, size_t* fwd_ws, size_t* bwd_ws) {
  int rc = validate(s);
  if (rc) return rc;
  MixerDims d = make_dims(*s);
  size_t sv, fw, bw;
  if (s->dtype == ADN_F32) {
    sv = SavedBufs<float>(d, nullptr).bytes;
    fw = FwdWs<float>(d, nullptr).bytes;
    bw = BwdWs<float>(d, nullptr).bytes;
  } else {
    sv = SavedBufs<bf16>(d, nullptr).bytes;
    fw = FwdWs<bf16>(d, nullptr).bytes;
    bw = BwdWs<bf16>(d, nullptr).bytes;
    size_t fw2 = 0, bw2 = 0;
    sm100_workspace_bytes(d, &fw2, &bw2);
    fw = fw > fw2 ? fw : fw2;
    bw = bw > bw2 ? bw : bw2;
    sv += sm100_saved_extra_bytes(d);
    if (!sm100_supported(d) && wide::supported(d)) wide::workspace_bytes(d, &sv, &fw, &bw);
  }
  if (saved_bytes) *saved_bytes = sv;
  if (fwd_ws) *fwd_ws = fw;
  if (bwd_ws) *bwd_ws = bw;
  return ADN_OK;
}

int adnssd_forward(const AdnShape* s, const AdnWeights* w, const void* u, void* out, void* saved, void* workspace,
                   void* stream) {
  int rc = validate(s);
  if (rc) return rc;
  rc = check_weights(w);
  if (rc) return rc;
  ADN_REQUIRE(u && out && workspace, ADN_ERR_NULL, "u / out / workspace must not be NULL");
  MixerDims d = make_dims(*s);
  cudaStream_t st = (cudaStream_t)stream;
  if (s->dtype == ADN_F32) return generic_forward<float>(d, *w, (const float*)u, (float*)out, saved, workspace, st);
  if (sm100_supported(d)) return sm100_forward(d, *w, (const bf16*)u, (bf16*)out, saved, workspace, st);
  if (wide::supported(d)) return wide::forward(d, *w, (const bf16*)u, (bf16*)out, saved, workspace, st);
  return generic_forward<bf16>(d, *w, (const bf16*)u, (bf16*)out, saved, workspace, st);
}

int adnssd_backward(const AdnShape* s, const AdnWeights* w, const void* u, const void* saved, const void* dout,
                    void* du, const AdnWeightGrads* g, void* workspace, void* stream) {
  int rc = validate(s);
  if (rc) return rc;
  rc = check_weights(w);
  if (rc) return rc;
  ADN_REQUIRE(u && saved && dout && du && g && workspace, ADN_ERR_NULL,
              "u / saved / dout / du / grads / workspace must not be NULL");
  MixerDims d = make_dims(*s);
  cudaStream_t st = (cudaStream_t)stream;
  if (s->dtype == ADN_F32)
    return generic_backward<float>(d, *w, (const float*)u, saved, (const float*)dout, (float*)du, *g, workspace, st);
  if (sm100_supported(d))
    return sm100_backward(d, *w, (const bf16*)u, saved, (const bf16*)dout, (bf16*)du, *g, workspace, st);
  if (wide::supported(d))
    return wide::backward(d, *w, (const bf16*)u, saved, (const bf16*)dout, (bf16*)du, *g, workspace, st);
  return generic_backward<bf16>(d, *w, (const bf16*)u, saved, (const bf16*)dout, (bf16*)du, *g, workspace, st);
}


// Diagnostics: flip a kernel-family switch between WHOLE forward + backward passes (the ADN_* environment variables are
// only read once, at first use).  Not thread-safe; tests and profiling scripts only.
int adn_set_option(const char* name, int value) {
  ADN_REQUIRE(name != nullptr, ADN_ERR_NULL, "adn_set_option: NULL name");
  EnvCfg& c = env_mut();
  if (!strcmp(name, "rowconv")) c.rowconv = value != 0;
  else if (!strcmp(name, "row_wide")) c.row_wide = value != 0;
  else if (!strcmp(name, "bwd_ws")) c.bwd_ws = value != 0;
  else if (!strcmp(name, "wide")) c.wide = value != 0;
  else if (!strcmp(name, "rows_per_cta")) c.rows_per_cta = value;
  else if (!strcmp(name, "du_dbg")) c.du_dbg = value;
  else if (!strcmp(name, "gemm_dbg")) c.gemm_dbg = value;
  else if (!strcmp(name, "variant")) c.variant = value;
  else { set_error("adn_set_option: unknown option '%s'", name); return ADN_ERR_SHAPE; }
  return ADN_OK;
}

// Which kernel family serves a shape: 0 generic CUDA-core (fp32 check mode, odd shapes), 1 tile kernels, 2 row kernels
// (both d_model 32, tcgen05), 3 wide path (tcgen05 GEMMs + bf16 bandwidth kernels).  For the sweep / tests.
int adnssd_kernel_family(const AdnShape* s) {
  if (validate(s)) return -1;
  MixerDims d = make_dims(*s);
  if (s->dtype == ADN_F32) return 0;
  if (sm100_supported(d)) return sm100_rowconv(d) ? 2 : 1;
  return wide::supported(d) ? 3 : 0;
}

// Standalone entry to the general tcgen05 GEMM (tests/test_tcgemm_gpu.py): C[b] = alpha * (A0 B0 [+ A1 B1]).
// a_mn / b_mn: 0 = operand stored [rows][K], 1 = stored [K][cols]; c_mode 0 bf16, 1 fp32, 2 fp32 atomic (C pre-zeroed).
int adn_selftest_gemm(int M, int N, int K0, int K1, int a_mn, int b_mn, const void* A0, long long lda0, long long a_bs0,
                      const void* B0, long long ldb0, long long b_bs0, const void* A1, long long lda1, long long a_bs1,
                      const void* B1, long long ldb1, long long b_bs1, void* C, long long ldc, long long c_bs, int c_mode,
                      int batches, int splitk, const float* alpha, int parity_mask, int* status, void* stream) {
  tcg::Op a0{(const bf16*)A0, lda0, a_bs0, a_mn}, b0{(const bf16*)B0, ldb0, b_bs0, b_mn};
  tcg::Op a1{(const bf16*)A1, lda1, a_bs1, a_mn}, b1{(const bf16*)B1, ldb1, b_bs1, b_mn};
  if (K1 == 0) { a1 = tcg::NOOP; b1 = tcg::NOOP; a1.mn = a_mn; b1.mn = b_mn; }
  int rc = tcg::gemm((cudaStream_t)stream, "tcgemm_selftest", M, N, K0, a0, b0, K1, a1, b1, tcg::Out{C, ldc, c_bs, c_mode}, batches,
                     splitk, alpha, parity_mask, status);
  if (rc) return rc;
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
