// Shared device/host helpers for the adnb200 library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/adnb200.h"

namespace adn {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);

#define ADN_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      adn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ADN_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define ADN_CHECK_LAUNCH() ADN_CHECK_CUDA(cudaGetLastError())

#define ADN_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      adn::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// Diagnostics (include/adnb200.h: adn_prof_*): every kernel launch of the library goes through one ProfScope, which
// counts the launch and, when profiling is enabled, brackets it with CUDA events on the launching stream.
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};
#define ADN_CAT2(a, b) a##b
#define ADN_CAT(a, b) ADN_CAT2(a, b)
#define ADN_KERNEL(name, st) adn::ProfScope ADN_CAT(_adn_prof_, __LINE__)(name, st)

// Number of SMs of the current device, queried once per process (148 on a B200; the fallback when no device is visible,
// e.g. the size queries of the CPU-only ABI tests).  Grid sizes and per-CTA slab buffers are derived from it.
int sm_count();
// Diagnostic switches (ADN_* environment variables), read ONCE at first use: forward and backward always agree on the
// kernel family and the saved layout, and no getenv() sits on the launch path.
struct EnvCfg {
  int rows_per_cta;   // ADN_ROWS_PER_CTA (0 = one CTA per SM)
  bool rowconv;       // ADN_ROWCONV=0    keeps the 128-wide shapes on the halo-tile kernels
  bool row_wide;      // ADN_ROW_WIDE=0   restricts the row kernels to W == 128
  bool bwd_ws;        // ADN_BWD_WS=0     monolithic tile kernels for B1 / B2
  int du_dbg;         // ADN_DU_DBG       knock-out mask of k_bconv_du
  bool wide;          // ADN_WIDE=0       keeps d_model >= 64 on the CUDA-core generic path
  int gemm_dbg;       // ADN_GEMM_DBG     knock-out mask of the tcgen05 GEMM (profiling: results are wrong when set)
  int variant;        // ADN_VARIANT      bit mask of kernel tuning variants under measurement (results identical)
};
const EnvCfg& env();

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar load/store with conversion to fp32 math
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- 4-element vector load/store (16 B for float, 8 B for bf16); pointers must be aligned accordingly
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx silu(x) = s * (1 + x * (1 - s))
__device__ __forceinline__ float silu_gradf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
// torch.nn.functional.softplus(beta=1, threshold=20)  (models/ADNssd.py:318)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(__expf(x)); }

// Storage-type dispatched math: the fp32 instantiation of the generic kernels IS the 1e-4 check mode, so it uses the
// accurate expf (sums that cancel - scalar gates, per-head decay gradients - amplify the ~1e-6 error of __expf a hundred-
// fold); the bf16 instantiations keep the fast intrinsics (their error is far below the bf16 rounding of the tensors).
template <typename T> struct Math {
  static __device__ __forceinline__ float exp(float x) { return __expf(x); }
};
template <> struct Math<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
};
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) { return 1.f / (1.f + Math<T>::exp(-x)); }
template <typename T> __device__ __forceinline__ float silu_t(float x) { return x * sigmoid_t<T>(x); }
template <typename T> __device__ __forceinline__ float silu_grad_t(float x) {
  const float s = sigmoid_t<T>(x);
  return s * (1.f + x * (1.f - s));
}
template <typename T> __device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(Math<T>::exp(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// head index of channel c of the x block: models/ADNssd.py:371-386,397-404 (parity split, then (h p) grouping)
__host__ __device__ __forceinline__ int head_of_channel(int c, int P) { return 2 * ((c >> 1) / P) + (c & 1); }

// Derived sizes of one mixer
struct MixerDims {
  int B, H, W, L, D, Di, P, G, N, GN, nh, Wd, CC, dip, ldr;
  long long T;  // tokens = B * L
};

static inline MixerDims make_dims(const AdnShape& s) {
  MixerDims d;
  d.B = s.B; d.H = s.H; d.W = s.W; d.L = s.H * s.W; d.D = s.D; d.Di = s.Di; d.P = s.P; d.G = s.G; d.N = s.N;
  d.GN = s.G * s.N;
  d.nh = s.P > 0 ? s.Di / s.P : 0;
  d.Wd = d.Di + 2 * d.GN;
  d.CC = 2 * d.Di + 2 * d.GN;
  d.dip = d.CC + d.nh;
  d.ldr = (d.dip + 7) / 8 * 8;  // row stride of the in_proj output (16-byte aligned rows for bf16 and fp32)
  d.T = (long long)d.B * d.L;
  return d;
}

}  // namespace adn
