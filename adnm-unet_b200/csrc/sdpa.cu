// Fused softmax attention for `StandardAttention` (models/ADNssd.py:26-47; the three `Attention` bridges of ADNM-UNet,
// models/ADNMUNet.py:172-238): out = softmax(q k^T * scale) v per (sample, head), straight from the packed to_qkv output
//   qkv [B][L][3 * heads * DH]   ('b n (h d)' order inside each of q | k | v, :41-42)   ->   out [B][L][heads * DH]
// ADNM-UNet builds it with dim_head = headdim = 4 and heads = dim / 4 (models/ADNMUNet.py:181): a 4-deep dot product per
// (query, key) pair.  That is CUDA-core work (a K = 4 contraction would leave a tensor-core tile 75 % empty), so the kernels
// are flash-style CUDA-core kernels: one thread owns one query (forward, dq) or one key (dk, dv), the other side streams
// through shared memory in tiles and is read as warp-wide broadcasts; the B x heads x L x L score tensor the reference
// materialises three times over (dots, attn, dropout(attn): 8.6 GB fp32 at B = 64, 32 heads, L = 1024) never exists.
// Online softmax in the exp2 domain; the log-sum-exp per query is the only thing saved for the backward pass.
#include "adn_common.cuh"

namespace adn {
namespace sdpa {

constexpr int TILE = 128;

template <typename T, int DH>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[DH]) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    float t[4];
    ld4(p + d, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[d + i] = t[i];
  }
}
template <typename T, int DH>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[DH]) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float t[4] = {v[d], v[d + 1], v[d + 2], v[d + 3]};
    st4(p + d, t);
  }
}
template <typename T> __device__ __forceinline__ float exp2_t(float x) { return exp2f(x); }      // accurate in both modes: MUFU.EX2 + range handling

// grid (ceil(L / TILE), heads, B), block TILE: thread = query
template <typename T, int DH>
__global__ void __launch_bounds__(TILE)
k_sdpa_fwd(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int L, int heads, float scale_log2e) {
  __shared__ __align__(16) float Ks[TILE][DH];
  __shared__ __align__(16) float Vs[TILE][DH];
  const int inner = heads * DH, ld = 3 * inner;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * TILE + threadIdx.x;
  const T* base = qkv + (long long)b * L * ld + h * DH;
  float q[DH], acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  if (i < L) load_vec<T, DH>(base + (long long)i * ld, q);
#pragma unroll
  for (int d = 0; d < DH; ++d) q[d] *= scale_log2e;
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < L; j0 += TILE) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    {
      float kv[DH], vv[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) { kv[d] = 0.f; vv[d] = 0.f; }
      if (j < L) { load_vec<T, DH>(base + (long long)j * ld + inner, kv); load_vec<T, DH>(base + (long long)j * ld + 2 * inner, vv); }
#pragma unroll
      for (int d = 0; d < DH; ++d) { Ks[threadIdx.x][d] = kv[d]; Vs[threadIdx.x][d] = vv[d]; }
    }
    __syncthreads();
    const int nj = min(TILE, L - j0);
    for (int c0 = 0; c0 < nj; c0 += 8) {
      float s[8];
      float cm = -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(q[d], Ks[c0 + c][d], a);
        s[c] = (c0 + c < nj) ? a : -INFINITY;
        cm = fmaxf(cm, s[c]);
      }
      const float mn = fmaxf(m, cm);
      const float corr = exp2_t<T>(m - mn);      // m = -inf on the first chunk: exp2(-inf) = 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[d] *= corr;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float p = exp2_t<T>(s[c] - mn);
        l += p;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = fmaf(p, Vs[c0 + c][d], acc[d]);
      }
      m = mn;
    }
  }
  if (i < L) {
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] *= inv;
    store_vec<T, DH>(out + ((long long)b * L + i) * inner + h * DH, acc);
    if (lse) lse[((long long)b * heads + h) * L + i] = m + log2f(l);      // log2 domain (of the scaled scores)
  }
}

// dq: thread = query.  p_ij = exp2(s_ij - lse_i); ds_ij = p_ij (do_i . v_j - D_i), D_i = do_i . out_i;  dq_i = scale sum_j ds_ij k_j
template <typename T, int DH>
__global__ void __launch_bounds__(TILE)
k_sdpa_bwd_q(const T* __restrict__ qkv, const T* __restrict__ out, const float* __restrict__ lse, const T* __restrict__ dout,
             T* __restrict__ dqkv, int L, int heads, float scale, float scale_log2e) {
  __shared__ __align__(16) float Ks[TILE][DH];
  __shared__ __align__(16) float Vs[TILE][DH];
  const int inner = heads * DH, ld = 3 * inner;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * TILE + threadIdx.x;
  const T* base = qkv + (long long)b * L * ld + h * DH;
  float q[DH], dO[DH], o[DH], dq[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; dO[d] = 0.f; o[d] = 0.f; dq[d] = 0.f; }
  float ls = 0.f;
  if (i < L) {
    load_vec<T, DH>(base + (long long)i * ld, q);
    load_vec<T, DH>(dout + ((long long)b * L + i) * inner + h * DH, dO);
    load_vec<T, DH>(out + ((long long)b * L + i) * inner + h * DH, o);
    ls = lse[((long long)b * heads + h) * L + i];
  }
  float Di = 0.f;
#pragma unroll
  for (int d = 0; d < DH; ++d) { Di = fmaf(dO[d], o[d], Di); q[d] *= scale_log2e; }
  for (int j0 = 0; j0 < L; j0 += TILE) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    {
      float kv[DH], vv[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) { kv[d] = 0.f; vv[d] = 0.f; }
      if (j < L) { load_vec<T, DH>(base + (long long)j * ld + inner, kv); load_vec<T, DH>(base + (long long)j * ld + 2 * inner, vv); }
#pragma unroll
      for (int d = 0; d < DH; ++d) { Ks[threadIdx.x][d] = kv[d]; Vs[threadIdx.x][d] = vv[d]; }
    }
    __syncthreads();
    const int nj = min(TILE, L - j0);
#pragma unroll 4
    for (int c = 0; c < nj; ++c) {
      float a = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { a = fmaf(q[d], Ks[c][d], a); dp = fmaf(dO[d], Vs[c][d], dp); }
      const float ds = exp2_t<T>(a - ls) * (dp - Di);
#pragma unroll
      for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, Ks[c][d], dq[d]);
    }
  }
  if (i < L) {
#pragma unroll
    for (int d = 0; d < DH; ++d) dq[d] *= scale;
    store_vec<T, DH>(dqkv + ((long long)b * L + i) * ld + h * DH, dq);
  }
}

// dk, dv: thread = key; queries stream through shared memory with their dout, lse and D.
//   dv_j = sum_i p_ij do_i;   dk_j = scale sum_i ds_ij q_i
template <typename T, int DH>
__global__ void __launch_bounds__(TILE)
k_sdpa_bwd_kv(const T* __restrict__ qkv, const T* __restrict__ out, const float* __restrict__ lse, const T* __restrict__ dout,
              T* __restrict__ dqkv, int L, int heads, float scale, float scale_log2e) {
  __shared__ __align__(16) float Qs[TILE][DH];
  __shared__ __align__(16) float Os[TILE][DH];
  __shared__ float2 LD[TILE];      // (lse_i, D_i)
  const int inner = heads * DH, ld = 3 * inner;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * TILE + threadIdx.x;
  const T* base = qkv + (long long)b * L * ld + h * DH;
  float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  if (j < L) { load_vec<T, DH>(base + (long long)j * ld + inner, k); load_vec<T, DH>(base + (long long)j * ld + 2 * inner, v); }
#pragma unroll
  for (int d = 0; d < DH; ++d) k[d] *= scale_log2e;
  for (int i0 = 0; i0 < L; i0 += TILE) {
    const int i = i0 + threadIdx.x;
    __syncthreads();
    {
      float qv[DH], dO[DH], o[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) { qv[d] = 0.f; dO[d] = 0.f; o[d] = 0.f; }
      float ls = INFINITY;      // rows past L: p = exp2(-inf) = 0
      if (i < L) {
        load_vec<T, DH>(base + (long long)i * ld, qv);
        load_vec<T, DH>(dout + ((long long)b * L + i) * inner + h * DH, dO);
        load_vec<T, DH>(out + ((long long)b * L + i) * inner + h * DH, o);
        ls = lse[((long long)b * heads + h) * L + i];
      }
      float Di = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { Di = fmaf(dO[d], o[d], Di); Qs[threadIdx.x][d] = qv[d]; Os[threadIdx.x][d] = dO[d]; }
      LD[threadIdx.x] = make_float2(ls, Di);
    }
    __syncthreads();
    const int ni = min(TILE, L - i0);
#pragma unroll 4
    for (int c = 0; c < ni; ++c) {
      float a = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { a = fmaf(Qs[c][d], k[d], a); dp = fmaf(Os[c][d], v[d], dp); }
      const float2 t = LD[c];
      const float p = exp2_t<T>(a - t.x);
      const float ds = p * (dp - t.y);
#pragma unroll
      for (int d = 0; d < DH; ++d) { dv[d] = fmaf(p, Os[c][d], dv[d]); dk[d] = fmaf(ds, Qs[c][d], dk[d]); }
    }
  }
  if (j < L) {
#pragma unroll
    for (int d = 0; d < DH; ++d) dk[d] *= scale;
    store_vec<T, DH>(dqkv + ((long long)b * L + j) * ld + inner + h * DH, dk);
    store_vec<T, DH>(dqkv + ((long long)b * L + j) * ld + 2 * inner + h * DH, dv);
  }
}

static int check(int B, int L, int heads, int dh, int dtype, const char* what) {
  ADN_REQUIRE(B > 0 && L > 0 && heads > 0 && B <= 65535 && heads <= 65535, ADN_ERR_SHAPE, "%s: B, L, heads must be positive (B, heads <= 65535)", what);
  ADN_REQUIRE(dh == 4 || dh == 8 || dh == 16, ADN_ERR_SHAPE, "%s: dim_head %d not in {4, 8, 16} (ADNM-UNet uses 4)", what, dh);
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "%s: unsupported dtype %d", what, dtype);
  return ADN_OK;
}

}  // namespace sdpa
}  // namespace adn

using namespace adn;
using namespace adn::sdpa;

extern "C" {

int adn_sdpa_forward(const void* qkv, void* out, float* lse, int32_t B, int32_t L, int32_t heads, int32_t dh, float scale,
                     int32_t dtype, void* stream) {
  int rc = check(B, L, heads, dh, dtype, "adn_sdpa_forward");
  if (rc) return rc;
  ADN_REQUIRE(qkv && out, ADN_ERR_NULL, "adn_sdpa_forward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(L, TILE), heads, B);
  const float sl2 = scale * 1.4426950408889634f;
  ADN_KERNEL("k_sdpa_fwd", st);
#define QKV(T) (const T*)qkv, (T*)out
  if (dtype == ADN_F32) {
    if (dh == 4) k_sdpa_fwd<float, 4><<<grid, TILE, 0, st>>>(QKV(float), lse, L, heads, sl2);
    else if (dh == 8) k_sdpa_fwd<float, 8><<<grid, TILE, 0, st>>>(QKV(float), lse, L, heads, sl2);
    else k_sdpa_fwd<float, 16><<<grid, TILE, 0, st>>>(QKV(float), lse, L, heads, sl2);
  } else {
    if (dh == 4) k_sdpa_fwd<bf16, 4><<<grid, TILE, 0, st>>>(QKV(bf16), lse, L, heads, sl2);
    else if (dh == 8) k_sdpa_fwd<bf16, 8><<<grid, TILE, 0, st>>>(QKV(bf16), lse, L, heads, sl2);
    else k_sdpa_fwd<bf16, 16><<<grid, TILE, 0, st>>>(QKV(bf16), lse, L, heads, sl2);
  }
#undef QKV
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_sdpa_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int32_t B, int32_t L,
                      int32_t heads, int32_t dh, float scale, int32_t dtype, void* stream) {
  int rc = check(B, L, heads, dh, dtype, "adn_sdpa_backward");
  if (rc) return rc;
  ADN_REQUIRE(qkv && out && lse && dout && dqkv, ADN_ERR_NULL, "adn_sdpa_backward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(L, TILE), heads, B);
  const float sl2 = scale * 1.4426950408889634f;
#define ARGS(T) (const T*)qkv, (const T*)out, lse, (const T*)dout, (T*)dqkv, L, heads, scale, sl2
  {
    ADN_KERNEL("k_sdpa_bwd_q", st);
    if (dtype == ADN_F32) {
      if (dh == 4) k_sdpa_bwd_q<float, 4><<<grid, TILE, 0, st>>>(ARGS(float));
      else if (dh == 8) k_sdpa_bwd_q<float, 8><<<grid, TILE, 0, st>>>(ARGS(float));
      else k_sdpa_bwd_q<float, 16><<<grid, TILE, 0, st>>>(ARGS(float));
    } else {
      if (dh == 4) k_sdpa_bwd_q<bf16, 4><<<grid, TILE, 0, st>>>(ARGS(bf16));
      else if (dh == 8) k_sdpa_bwd_q<bf16, 8><<<grid, TILE, 0, st>>>(ARGS(bf16));
      else k_sdpa_bwd_q<bf16, 16><<<grid, TILE, 0, st>>>(ARGS(bf16));
    }
  }
  {
    ADN_KERNEL("k_sdpa_bwd_kv", st);
    if (dtype == ADN_F32) {
      if (dh == 4) k_sdpa_bwd_kv<float, 4><<<grid, TILE, 0, st>>>(ARGS(float));
      else if (dh == 8) k_sdpa_bwd_kv<float, 8><<<grid, TILE, 0, st>>>(ARGS(float));
      else k_sdpa_bwd_kv<float, 16><<<grid, TILE, 0, st>>>(ARGS(float));
    } else {
      if (dh == 4) k_sdpa_bwd_kv<bf16, 4><<<grid, TILE, 0, st>>>(ARGS(bf16));
      else if (dh == 8) k_sdpa_bwd_kv<bf16, 8><<<grid, TILE, 0, st>>>(ARGS(bf16));
      else k_sdpa_bwd_kv<bf16, 16><<<grid, TILE, 0, st>>>(ARGS(bf16));
    }
  }
#undef ARGS
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
