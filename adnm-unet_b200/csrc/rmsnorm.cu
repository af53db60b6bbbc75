// Standalone RMSNorm of the reference README (README.md:22-30; the Block pre-norms norm1 / norm2, models/ADNMUNet.py:149,155,
// 278) fused with the Block's scalar affine `scale * norm(x) + shift`:
//     h = x * rsqrt(mean(x^2) + eps);   y = scale * (h * weight) + shift
// One pass forward, one pass backward, per-token statistics with warp shuffles, HBM-bound (the eager reference issues six
// elementwise / reduction kernels forward and about a dozen backward over the same (B, L, D) tensor).
//   forward : reads x, writes y (+ rstd, 4 bytes per token)                       2 * D * e bytes per token
//   backward: reads x, dy, writes dx; dweight / dscale / dshift via per-block partial sums + atomics
// D = 8 * G with G in {1, 2, 4, ..., 32} (every width of ADNM-UNet up to 256): G lanes own a token, 8 channels (one 16-byte
// bf16 load) per lane, weight and the dweight partial sums live in registers for the whole pass, statistics by shuffles
// inside the lane group - at D = 32 (the refiner Blocks, 524 288 tokens per step) a warp normalises 8 tokens at a time.
// Other widths: one warp per token, 4 channels per lane and step.  D <= 8192.  fp32 or bf16 activations, fp32 weight /
// statistics / parameter gradients.  Backward optionally adds `dres` - the gradient that reaches x through the Block's
// residual path - so that dx leaves in ONE rounding and one pass (dx = dx_norm + dres).
#include "adn_common.cuh"
#include "sm100_utils.cuh"

namespace adn {

__device__ __forceinline__ void ldg8(const bf16* p, float (&v)[8]) { sm100::unpack8(*reinterpret_cast<const uint4*>(p), v); }
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void stg8(bf16* p, const float (&v)[8]) { *reinterpret_cast<uint4*>(p) = sm100::pack8(v); }
__device__ __forceinline__ void stg8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- lane-group kernels: D == 8 * G
template <typename T, int G>
__global__ void __launch_bounds__(256)
k_rmsnorm_fwd_g(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale_p, const float* __restrict__ shift_p,
                T* __restrict__ y, float* __restrict__ rstd_out, long long Ttok, float eps) {
  constexpr int D = 8 * G, TPW = 32 / G;
  const int lane = threadIdx.x & 31, li = lane % G, c = li * 8;
  const float scale = scale_p ? *scale_p : 1.f, shift = shift_p ? *shift_p : 0.f;
  float wv[8];
  ldg8(w + c, wv);
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] *= scale;
  const long long warps = (long long)gridDim.x * 8, gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (long long t0 = gw * TPW; t0 < Ttok; t0 += warps * TPW) {
    const long long t = t0 + lane / G;
    const bool ok = t < Ttok;
    float v[8] = {};
    if (ok) ldg8(x + t * D + c, v);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
    const float rstd = rsqrtf(group_sum<G>(ss) / D + eps);
    if (ok) {
      if (li == 0 && rstd_out) rstd_out[t] = rstd;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(v[i] * rstd, wv[i], shift);
      stg8(y + t * D + c, o);
    }
  }
}

template <typename T, int G>
__global__ void __launch_bounds__(256)
k_rmsnorm_bwd_g(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale_p, const float* __restrict__ rstd_in,
                const T* __restrict__ dy, const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dweight,
                float* __restrict__ dscale, float* __restrict__ dshift, long long Ttok) {
  constexpr int D = 8 * G, TPW = 32 / G;
  __shared__ float sw[8][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = lane % G, c = li * 8;
  const float scale = scale_p ? *scale_p : 1.f;
  float wv[8], dwv[8] = {};
  ldg8(w + c, wv);
  float a_scale = 0.f, a_shift = 0.f;
  const long long warps = (long long)gridDim.x * 8, gw = (long long)blockIdx.x * 8 + warp;
  for (long long t0 = gw * TPW; t0 < Ttok; t0 += warps * TPW) {
    const long long t = t0 + lane / G;
    const bool ok = t < Ttok;
    float v[8] = {}, d[8] = {};
    float rstd = 0.f;
    if (ok) { ldg8(x + t * D + c, v); ldg8(dy + t * D + c, d); rstd = rstd_in[t]; }
    float m = 0.f;      // sum_c dh * h,  dh = scale * dy * w
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] *= rstd; m = fmaf(scale * d[i] * wv[i], v[i], m); }
    m = group_sum<G>(m) / D;
    if (ok) {
      float o[8];
      if (dres) ldg8(dres + t * D + c, o);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] += rstd * (scale * d[i] * wv[i] - v[i] * m);
        dwv[i] = fmaf(scale * d[i], v[i], dwv[i]);
        a_scale = fmaf(d[i], v[i] * wv[i], a_scale);
        a_shift += d[i];
      }
      stg8(dx + t * D + c, o);
    }
  }
  // fold the TPW lane groups of the warp, then the 8 warps, then one atomic per channel and block
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = G; o < 32; o <<= 1) dwv[i] += __shfl_xor_sync(0xffffffffu, dwv[i], o);
  }
  if (lane < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sw[warp][c + i] = dwv[i];
  }
  a_scale = warp_sum(a_scale);
  a_shift = warp_sum(a_shift);
  if (lane == 0) {
    if (dscale && a_scale != 0.f) atomicAdd(dscale, a_scale);
    if (dshift && a_shift != 0.f) atomicAdd(dshift, a_shift);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += sw[q][i];
    if (v != 0.f) atomicAdd(dweight + i, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_rmsnorm_fwd(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale_p,
              const float* __restrict__ shift_p, T* __restrict__ y, float* __restrict__ rstd_out, long long Ttok, int D, float eps) {
  const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= Ttok) return;
  const float scale = scale_p ? *scale_p : 1.f, shift = shift_p ? *shift_p : 0.f;
  const T* xr = x + t * D;
  float ss = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float v[4];
    ld4(xr + c, v);
    ss += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
  if (lane == 0 && rstd_out) rstd_out[t] = rstd;
  for (int c = lane * 4; c < D; c += 128) {
    float v[4], g[4], o[4];
    ld4(xr + c, v);
    ld4(w + c, g);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fmaf(scale, v[i] * rstd * g[i], shift);
    st4(y + t * D + c, o);
  }
}

// dy -> dx, and per-block partial sums of dweight[c] = sum_t scale * dy * h, dscale = sum dy * h * w, dshift = sum dy.
// A warp owns `tpw` consecutive tokens and keeps its dweight partials in registers (D <= 128 * DW_MAX channels per lane
// pass), flushed through shared memory once per block.
constexpr int RMS_DW_MAX = 4;      // register path for D <= 512; wider rows fall back to shared-memory atomics per token
template <typename T>
__global__ void __launch_bounds__(256)
k_rmsnorm_bwd(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale_p,
              const float* __restrict__ rstd_in, const T* __restrict__ dy, const T* __restrict__ dres, T* __restrict__ dx,
              float* __restrict__ dweight, float* __restrict__ dscale, float* __restrict__ dshift, long long Ttok, int tpw, int D) {
  extern __shared__ float sw[];      // [D] block-local dweight
  for (int i = threadIdx.x; i < D; i += blockDim.x) sw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = scale_p ? *scale_p : 1.f;
  const long long t0 = ((long long)blockIdx.x * 8 + warp) * tpw;
  const bool regs = D <= 128 * RMS_DW_MAX;
  float dwr[RMS_DW_MAX][4] = {};
  float a_scale = 0.f, a_shift = 0.f;
  for (long long t = t0; t < min(Ttok, t0 + (long long)tpw); ++t) {
    const T* xr = x + t * D;
    const T* gr = dy + t * D;
    const float rstd = rstd_in[t];
    float m = 0.f;      // sum_c dh * h,  dh = scale * dy * w
    for (int c = lane * 4; c < D; c += 128) {
      float v[4], g[4], d[4];
      ld4(xr + c, v); ld4(w + c, g); ld4(gr + c, d);
#pragma unroll
      for (int i = 0; i < 4; ++i) m = fmaf(scale * d[i] * g[i], v[i] * rstd, m);
    }
    m = warp_sum(m) / D;
    int k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k) {
      float v[4], g[4], d[4], o[4], r[4] = {0.f, 0.f, 0.f, 0.f};
      ld4(xr + c, v); ld4(w + c, g); ld4(gr + c, d);
      if (dres) ld4(dres + t * D + c, r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float h = v[i] * rstd;
        o[i] = r[i] + rstd * (scale * d[i] * g[i] - h * m);
        const float dwv = scale * d[i] * h;
        if (regs) dwr[k < RMS_DW_MAX ? k : 0][i] += dwv; else atomicAdd(&sw[c + i], dwv);
        a_scale = fmaf(d[i], h * g[i], a_scale);
        a_shift += d[i];
      }
      st4(dx + t * D + c, o);
    }
  }
  if (regs) {
    int k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(&sw[c + i], dwr[k < RMS_DW_MAX ? k : 0][i]);
  }
  a_scale = warp_sum(a_scale);
  a_shift = warp_sum(a_shift);
  if (lane == 0) {
    if (dscale && a_scale != 0.f) atomicAdd(dscale, a_scale);
    if (dshift && a_shift != 0.f) atomicAdd(dshift, a_shift);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    if (sw[i] != 0.f) atomicAdd(dweight + i, sw[i]);
}

static int rms_validate(long long T, int D, int dtype, const char* what) {
  ADN_REQUIRE(T > 0 && D > 0 && D % 4 == 0 && D <= 8192, ADN_ERR_SHAPE, "%s: tokens > 0 and D a multiple of 4 in [4, 8192] required (got %lld, %d)",
              what, T, D);
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "%s: unsupported dtype %d", what, dtype);
  return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" {

#define RMS_G_DISPATCH(KERNEL, T, ...)                                                    \
  switch (D / 8) {                                                                        \
    case 1: KERNEL<T, 1><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                       \
    case 2: KERNEL<T, 2><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                       \
    case 4: KERNEL<T, 4><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                       \
    case 8: KERNEL<T, 8><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                       \
    case 16: KERNEL<T, 16><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                     \
    default: KERNEL<T, 32><<<grid, 256, 0, st>>>(__VA_ARGS__); break;                     \
  }
static inline bool rms_group_shape(int D) { return D % 8 == 0 && D <= 256 && ((D / 8) & (D / 8 - 1)) == 0; }
static inline int rms_group_grid(long long tokens, int D) {
  const long long tpw = 32 / (D / 8), want = (tokens + 8 * tpw - 1) / (8 * tpw), cap = (long long)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

int adn_rmsnorm_forward(const void* x, const float* weight, const float* scale, const float* shift, void* y, float* rstd,
                        int64_t tokens, int32_t D, float eps, int32_t dtype, void* stream) {
  int rc = rms_validate(tokens, D, dtype, "adn_rmsnorm_forward");
  if (rc) return rc;
  ADN_REQUIRE(x && weight && y, ADN_ERR_NULL, "adn_rmsnorm_forward: x / weight / y must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (rms_group_shape(D)) {
    const int grid = rms_group_grid(tokens, D);
    ADN_KERNEL("k_rmsnorm_fwd_g", st);
    if (dtype == ADN_F32) { RMS_G_DISPATCH(k_rmsnorm_fwd_g, float, (const float*)x, weight, scale, shift, (float*)y, rstd, tokens, eps) }
    else { RMS_G_DISPATCH(k_rmsnorm_fwd_g, bf16, (const bf16*)x, weight, scale, shift, (bf16*)y, rstd, tokens, eps) }
    ADN_CHECK_LAUNCH();
    return ADN_OK;
  }
  const int grid = cdiv(tokens, 8);
  if (dtype == ADN_F32) { ADN_KERNEL("k_rmsnorm_fwd", st); k_rmsnorm_fwd<float><<<grid, 256, 0, st>>>((const float*)x, weight, scale, shift, (float*)y, rstd, tokens, D, eps); }
  else { ADN_KERNEL("k_rmsnorm_fwd", st); k_rmsnorm_fwd<bf16><<<grid, 256, 0, st>>>((const bf16*)x, weight, scale, shift, (bf16*)y, rstd, tokens, D, eps); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_rmsnorm_backward(const void* x, const float* weight, const float* scale, const float* rstd, const void* dy, const void* dres,
                         void* dx, float* dweight, float* dscale, float* dshift, int64_t tokens, int32_t D, int32_t dtype,
                         void* stream) {
  int rc = rms_validate(tokens, D, dtype, "adn_rmsnorm_backward");
  if (rc) return rc;
  ADN_REQUIRE(x && weight && rstd && dy && dx && dweight, ADN_ERR_NULL, "adn_rmsnorm_backward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  ADN_CHECK_CUDA(cudaMemsetAsync(dweight, 0, (size_t)D * sizeof(float), st));
  if (dscale) ADN_CHECK_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float), st));
  if (dshift) ADN_CHECK_CUDA(cudaMemsetAsync(dshift, 0, sizeof(float), st));
  if (rms_group_shape(D)) {
    long long want = rms_group_grid(tokens, D);
    const int grid = (int)(want > 4LL * sm_count() ? 4LL * sm_count() : want);      // fewer, longer-lived blocks: one flush of the partial sums each
    ADN_KERNEL("k_rmsnorm_bwd_g", st);
    if (dtype == ADN_F32) { RMS_G_DISPATCH(k_rmsnorm_bwd_g, float, (const float*)x, weight, scale, rstd, (const float*)dy, (const float*)dres, (float*)dx, dweight, dscale, dshift, tokens) }
    else { RMS_G_DISPATCH(k_rmsnorm_bwd_g, bf16, (const bf16*)x, weight, scale, rstd, (const bf16*)dy, (const bf16*)dres, (bf16*)dx, dweight, dscale, dshift, tokens) }
    ADN_CHECK_LAUNCH();
    return ADN_OK;
  }
  long long tpw = tokens / (16LL * sm_count());
  tpw = tpw < 1 ? 1 : (tpw > 64 ? 64 : tpw);
  const int grid = cdiv(tokens, 8 * tpw);
  const size_t smem = (size_t)D * sizeof(float);
  if (dtype == ADN_F32) { ADN_KERNEL("k_rmsnorm_bwd", st); k_rmsnorm_bwd<float><<<grid, 256, smem, st>>>((const float*)x, weight, scale, rstd, (const float*)dy, (const float*)dres, (float*)dx, dweight, dscale, dshift, tokens, (int)tpw, D); }
  else { ADN_KERNEL("k_rmsnorm_bwd", st); k_rmsnorm_bwd<bf16><<<grid, 256, smem, st>>>((const bf16*)x, weight, scale, rstd, (const bf16*)dy, (const bf16*)dres, (bf16*)dx, dweight, dscale, dshift, tokens, (int)tpw, D); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
