// Data-parallel training step tail on flat fp32 buffers: global gradient norm, clip, AdamW, gradient reset.
// Replaces, for the batch-sharded trainer (adnm_unet_b200/trainer.py), the reference's per-tensor sequence
//   torch.nn.utils.clip_grad_norm_(model.parameters(), current_norm)   train.py:140
//   optimizer.step()  (AdamW lr 1e-3, betas .9/.999, eps 1e-9, wd 1e-2) train.py:144, train_untils.py:35-42
//   optimizer.zero_grad()                                              train.py:145
// with two HBM-bound passes over the 71.8 M live elements and no host synchronisation (the reference reads the norm
// back with .item() every step, train.py:141).  Pure bandwidth: sumsq reads 4 B/element, adamw reads 16 and writes 16.
#include "adn_common.cuh"

namespace adn {

// Deterministic two-stage sum of squares: stage 1 one partial per block (fixed grid), stage 2 one block adds them in a
// fixed order.  No atomics -> every rank computes bit-identical norms from bit-identical (all-reduced) gradients, so the
// replicas' weights never drift apart.
__global__ void __launch_bounds__(256)
k_sumsq_partial(const float* __restrict__ x, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
k_sumsq_final(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay;
  float bias1, bias2_sqrt;   // 1 - beta1^t, sqrt(1 - beta2^t)
  float grad_scale;          // 1 / world  (the buckets hold the SUM over ranks)
  float max_norm;            // <= 0: no clipping
};

// g_eff = g * grad_scale * min(1, max_norm / (||g * grad_scale|| + 1e-6))     (torch.nn.utils.clip_grad_norm_)
// p *= 1 - lr*wd;  m = b1 m + (1-b1) g_eff;  v = b2 v + (1-b2) g_eff^2;  p -= lr/bias1 * m / (sqrt(v)/bias2_sqrt + eps)
// g = 0 (optimizer.zero_grad, fused so the gradient buffer is not touched a second time)
__global__ void __launch_bounds__(256)
k_adamw_flat(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             const float* __restrict__ sumsq, float* __restrict__ norm_out, AdamWArgs a) {
  float coef = a.grad_scale;
  if (sumsq != nullptr) {
    const float norm = sqrtf(*sumsq) * a.grad_scale;
    if (a.max_norm > 0.f) coef *= fminf(1.f, a.max_norm / (norm + 1e-6f));
    if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  }
  const float decay = 1.f - a.lr * a.weight_decay, step = a.lr / a.bias1, ib2 = 1.f / a.bias2_sqrt;
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pv.x; float* G = &gv.x; float* M = &mv.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float ge = G[k] * coef;
      M[k] = a.beta1 * M[k] + (1.f - a.beta1) * ge;
      V[k] = a.beta2 * V[k] + (1.f - a.beta2) * ge * ge;
      P[k] = P[k] * decay - step * (M[k] / (sqrtf(V[k]) * ib2 + a.eps));
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float ge = g[i] * coef;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * ge, vi = a.beta2 * v[i] + (1.f - a.beta2) * ge * ge;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] * decay - step * (mi / (sqrtf(vi) * ib2 + a.eps));
    g[i] = 0.f;
  }
}

constexpr int SUMSQ_MAX_PARTS = 2048;

}  // namespace adn

extern "C" {

int adn_sumsq_workspace_floats(void) { return adn::SUMSQ_MAX_PARTS; }

int adn_sumsq_f32(const float* x, int64_t n, float* partial_ws, float* out, void* stream) {
  using namespace adn;
  ADN_REQUIRE(x && partial_ws && out, ADN_ERR_NULL, "adn_sumsq_f32: NULL argument");
  ADN_REQUIRE(n >= 0 && (uintptr_t)x % 16 == 0, ADN_ERR_SHAPE, "adn_sumsq_f32: n >= 0 and 16-byte aligned input required");
  cudaStream_t st = (cudaStream_t)stream;
  long long want = (n / 4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > sm_count() * 8 ? sm_count() * 8 : want));
  if (grid > SUMSQ_MAX_PARTS) grid = SUMSQ_MAX_PARTS;
  { ADN_KERNEL("k_sumsq_partial", st); k_sumsq_partial<<<grid, 256, 0, st>>>(x, n, partial_ws); }
  { ADN_KERNEL("k_sumsq_final", st); k_sumsq_final<<<1, 256, 0, st>>>(partial_ws, grid, out); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_adamw_flat(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float* norm_out, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   float max_norm, void* stream) {
  using namespace adn;
  ADN_REQUIRE(p && g && m && v, ADN_ERR_NULL, "adn_adamw_flat: NULL argument");
  ADN_REQUIRE(n >= 0 && step >= 1, ADN_ERR_SHAPE, "adn_adamw_flat: n >= 0 and step >= 1 required");
  ADN_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, ADN_ERR_SHAPE,
              "adn_adamw_flat: buffers must be 16-byte aligned");
  ADN_REQUIRE(max_norm <= 0.f || sumsq != nullptr, ADN_ERR_NULL, "adn_adamw_flat: clipping needs the sum of squares");
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.grad_scale = grad_scale; a.max_norm = max_norm;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return ADN_OK;
  long long want = (n / 4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > sm_count() * 8 ? sm_count() * 8 : want));
  { ADN_KERNEL("k_adamw_flat", st); k_adamw_flat<<<grid, 256, 0, st>>>(p, g, m, v, n, sumsq, norm_out, a); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
