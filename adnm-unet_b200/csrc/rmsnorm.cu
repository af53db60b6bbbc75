// Standalone RMSNorm of the reference README (README.md:22-30; the Block pre-norms norm1 / norm2, models/ADNMUNet.py:149,155,
// 278) fused with the Block's scalar affine `scale * norm(x) + shift`:
//     h = x * rsqrt(mean(x^2) + eps);   y = scale * (h * weight) + shift
// One pass forward, one pass backward, per-token statistics with warp shuffles, HBM-bound (the eager reference issues six
// elementwise / reduction kernels forward and about a dozen backward over the same (B, L, D) tensor).
//   forward : reads x, writes y (+ rstd, 4 bytes per token)                       2 * D * e bytes per token
//   backward: reads x, dy, writes dx; dweight / dscale / dshift via per-block partial sums + atomics
// One warp per token; D <= 8192.  fp32 or bf16 activations, fp32 weight / statistics / parameter gradients.
#include "adn_common.cuh"

namespace adn {

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
              const float* __restrict__ rstd_in, const T* __restrict__ dy, T* __restrict__ dx, float* __restrict__ dweight,
              float* __restrict__ dscale, float* __restrict__ dshift, long long Ttok, int tpw, int D) {
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
      float v[4], g[4], d[4], o[4];
      ld4(xr + c, v); ld4(w + c, g); ld4(gr + c, d);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float h = v[i] * rstd;
        o[i] = rstd * (scale * d[i] * g[i] - h * m);
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

int adn_rmsnorm_forward(const void* x, const float* weight, const float* scale, const float* shift, void* y, float* rstd,
                        int64_t tokens, int32_t D, float eps, int32_t dtype, void* stream) {
  int rc = rms_validate(tokens, D, dtype, "adn_rmsnorm_forward");
  if (rc) return rc;
  ADN_REQUIRE(x && weight && y, ADN_ERR_NULL, "adn_rmsnorm_forward: x / weight / y must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(tokens, 8);
  if (dtype == ADN_F32) { ADN_KERNEL("k_rmsnorm_fwd", st); k_rmsnorm_fwd<float><<<grid, 256, 0, st>>>((const float*)x, weight, scale, shift, (float*)y, rstd, tokens, D, eps); }
  else { ADN_KERNEL("k_rmsnorm_fwd", st); k_rmsnorm_fwd<bf16><<<grid, 256, 0, st>>>((const bf16*)x, weight, scale, shift, (bf16*)y, rstd, tokens, D, eps); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_rmsnorm_backward(const void* x, const float* weight, const float* scale, const float* rstd, const void* dy, void* dx,
                         float* dweight, float* dscale, float* dshift, int64_t tokens, int32_t D, int32_t dtype, void* stream) {
  int rc = rms_validate(tokens, D, dtype, "adn_rmsnorm_backward");
  if (rc) return rc;
  ADN_REQUIRE(x && weight && rstd && dy && dx && dweight, ADN_ERR_NULL, "adn_rmsnorm_backward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  ADN_CHECK_CUDA(cudaMemsetAsync(dweight, 0, (size_t)D * sizeof(float), st));
  if (dscale) ADN_CHECK_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float), st));
  if (dshift) ADN_CHECK_CUDA(cudaMemsetAsync(dshift, 0, sizeof(float), st));
  long long tpw = tokens / (16LL * sm_count());
  tpw = tpw < 1 ? 1 : (tpw > 64 ? 64 : tpw);
  const int grid = cdiv(tokens, 8 * tpw);
  const size_t smem = (size_t)D * sizeof(float);
  if (dtype == ADN_F32) { ADN_KERNEL("k_rmsnorm_bwd", st); k_rmsnorm_bwd<float><<<grid, 256, smem, st>>>((const float*)x, weight, scale, rstd, (const float*)dy, (float*)dx, dweight, dscale, dshift, tokens, (int)tpw, D); }
  else { ADN_KERNEL("k_rmsnorm_bwd", st); k_rmsnorm_bwd<bf16><<<grid, 256, smem, st>>>((const bf16*)x, weight, scale, rstd, (const bf16*)dy, (bf16*)dx, dweight, dscale, dshift, tokens, (int)tpw, D); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
