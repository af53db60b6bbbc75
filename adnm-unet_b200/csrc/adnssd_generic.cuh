// Generic (any d_model / headdim / d_state) ADN-SSD mixer path: plain CUDA-core kernels, all math fp32.
// Instantiated for float ("check mode", ADN_F32) and for bf16 storage.  The sm_100a tensor-core path
// (adnssd_sm100.cu) replaces the contractions for the shapes it supports; everything here is also the
// shape-fallback of the bf16 path.  Stage split follows oracle/adnssd_oracle.py (mixer_forward / mixer_backward).
#pragma once
#include "adn_common.cuh"

namespace adn {

// ------------------------------------------------------------------------------------------------
// conv kernel assembly: the ten conv weight tensors -> one per-channel 3x3 kernel Kc[CC][9]
// (models/ADNssd.py:329-364 for xBC, :388-390 for z; the 3x1 o 1x3 pairs are rank-1 3x3 kernels)
// ------------------------------------------------------------------------------------------------
struct ConvWeightPtrs {
  const float *c13x1, *c31x1, *c13x2, *c31x2, *c13bc1, *c31bc1, *c13bc2, *c31bc2, *c2d, *c2dz;
};

// out9[0..8] = the 3x3 kernel of channel cc
__device__ __forceinline__ void assemble_conv_channel(const ConvWeightPtrs& w, float* __restrict__ out9, int Di, int cc) {
  float k[9];
  if (cc < Di) {
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = w.c2dz[cc * 9 + t];
  } else {
    int c = cc - Di;
    if ((c & 1) == 0) {
#pragma unroll
      for (int t = 0; t < 9; ++t) k[t] = w.c2d[(c >> 1) * 9 + t];
    } else {
      int i = c >> 2, nx = Di >> 2;
      bool first = (c & 3) == 1;
      const float *w31, *w13;
      if (i < nx) {
        w31 = (first ? w.c31x1 : w.c31x2) + i * 3;
        w13 = (first ? w.c13x1 : w.c13x2) + i * 3;
      } else {
        w31 = (first ? w.c31bc1 : w.c31bc2) + (i - nx) * 3;
        w13 = (first ? w.c13bc1 : w.c13bc2) + (i - nx) * 3;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) k[a * 3 + b] = w31[a] * w13[b];
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) out9[t] = k[t];
}

static __global__ void k_assemble_conv(ConvWeightPtrs w, float* __restrict__ Kc, int Di, int CC) {
  int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc < CC) assemble_conv_channel(w, Kc + cc * 9, Di, cc);
}

// ------------------------------------------------------------------------------------------------
// C[M,N] (=|+=) alpha * A[M,K] . W  with W either [N,K] (W_NK, "x @ W^T") or [K,N] ("x @ W"); fp32 W.
// grid (ceil(M/64), ceil(N/64), batches)
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TC, bool W_NK>
__global__ void __launch_bounds__(256)
k_gemm(const TA* __restrict__ A, long long lda, long long a_bs, const float* __restrict__ Wt, long long ldw,
       long long w_bs, TC* __restrict__ C, long long ldc, long long c_bs, int M, int N, int K,
       const float* __restrict__ alpha_ptr, int accumulate) {
  __shared__ float As[16][68];
  __shared__ float Ws[16][68];
  A += (long long)blockIdx.z * a_bs;
  Wt += (long long)blockIdx.z * w_bs;
  C += (long long)blockIdx.z * c_bs;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    {
      int r = tid >> 2, kk = (tid & 3) * 4;
      int m = m0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + kk + i;
        As[kk + i][r] = (m < M && k < K) ? ldf(A + (long long)m * lda + k) : 0.f;
      }
    }
    if (W_NK) {
      int r = tid >> 2, kk = (tid & 3) * 4;
      int n = n0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + kk + i;
        Ws[kk + i][r] = (n < N && k < K) ? Wt[(long long)n * ldw + k] : 0.f;
      }
    } else {
      int kk = tid >> 4, nn = (tid & 15) * 4;
      int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int n = n0 + nn + i;
        Ws[kk][nn + i] = (n < N && k < K) ? Wt[(long long)k * ldw + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float alpha = alpha_ptr ? *alpha_ptr : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      TC* p = C + (long long)m * ldc + n;
      float v = alpha * acc[i][j];
      if (accumulate) v += ldf(p);
      stf(p, v);
    }
  }
}

template <typename TA, typename TC, bool W_NK>
static inline void launch_gemm(cudaStream_t st, const TA* A, long long lda, long long a_bs, const float* Wt,
                               long long ldw, long long w_bs, TC* C, long long ldc, long long c_bs, int M, int N,
                               int K, int batches, const float* alpha_ptr, int accumulate) {
  dim3 grid(cdiv(M, 64), cdiv(N, 64), batches);  // M (tokens) on grid.x: no 65535 limit
  { ADN_KERNEL("k_gemm", st); k_gemm<TA, TC, W_NK><<<grid, 256, 0, st>>>(A, lda, a_bs, Wt, ldw, w_bs, C, ldc, c_bs, M, N, K, alpha_ptr,
                                              accumulate); }
}

// ------------------------------------------------------------------------------------------------
// Out[b][n1][n2] += sum_{l in batch b} X[l,n1] * Y[l,n2]    (fp32 atomics; Out must be zeroed)
// optional parity mask keeps only (n1 & 1) == (n2 & 1)  (the even/odd SSD split, models/ADNssd.py:397-404)
// grid (ceil(N2/64), ceil(N1/64), B*splits)
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
k_reduce_gemm(const TX* __restrict__ X, long long ldx, const TY* __restrict__ Y, long long ldy,
              float* __restrict__ Out, long long ldo, long long o_bs, int N1, int N2, int L, int chunk, int splits,
              int parity_mask) {
  __shared__ float Xs[16][68];
  __shared__ float Ys[16][68];
  const int b = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int l0 = sp * chunk, l1 = min(L, l0 + chunk);
  X += (long long)b * L * ldx;
  Y += (long long)b * L * ldy;
  Out += (long long)b * o_bs;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int lb = l0; lb < l1; lb += 16) {
    int kk = tid >> 4, nn = (tid & 15) * 4;
    int l = lb + kk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n1 = i0 + nn + i, n2 = j0 + nn + i;
      Xs[kk][nn + i] = (l < l1 && n1 < N1) ? ldf(X + (long long)l * ldx + n1) : 0.f;
      Ys[kk][nn + i] = (l < l1 && n2 < N2) ? ldf(Y + (long long)l * ldy + n2) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ys[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n1 = i0 + ty * 4 + i;
    if (n1 >= N1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n2 = j0 + tx * 4 + j;
      if (n2 >= N2) continue;
      if (parity_mask && ((n1 ^ n2) & 1)) continue;
      atomicAdd(Out + (long long)n1 * ldo + n2, acc[i][j]);
    }
  }
}

template <typename TX, typename TY>
static inline void launch_reduce_gemm(cudaStream_t st, const TX* X, long long ldx, const TY* Y, long long ldy,
                                      float* Out, long long ldo, long long o_bs, int N1, int N2, int L, int B,
                                      int parity_mask) {
  // enough splits to fill the machine (148 SMs x ~4 CTAs) without shredding the reduction
  int tiles = cdiv(N1, 64) * cdiv(N2, 64) * B;
  int splits = max(1, min(cdiv(L, 64), cdiv(sm_count() * 4, tiles)));
  int chunk = cdiv(cdiv(L, splits), 16) * 16;
  splits = cdiv(L, chunk);
  dim3 grid(cdiv(N2, 64), cdiv(N1, 64), B * splits);
  { ADN_KERNEL("k_reduce_gemm", st); k_reduce_gemm<TX, TY><<<grid, 256, 0, st>>>(X, ldx, Y, ldy, Out, ldo, o_bs, N1, N2, L, chunk, splits, parity_mask); }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 + SiLU, channels-last.  thread = (4 channels, one column x), walks ROWS rows keeping a
// 3x3 register window.  block (8 channel-vectors, 32 columns); grid (ceil(CV/8), ceil(W/32), B*ceil(H/ROWS))
// ------------------------------------------------------------------------------------------------
constexpr int CONV_ROWS = 8;

template <typename T>
__device__ __forceinline__ void load_row3(const T* __restrict__ base, long long ld, int W, int y, int H, int x,
                                          float (&r)[3][4]) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int xx = x + s - 1;
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
      ld4(base + ((long long)y * W + xx) * ld, r[s]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[s][i] = 0.f;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_conv_fwd(const T* __restrict__ raw, long long ldr, const float* __restrict__ Kc, T* __restrict__ pre,
           T* __restrict__ act, int H, int W, int CC) {
  const int CV = CC >> 2;
  const int cv = blockIdx.x * 8 + threadIdx.x;
  const int x = blockIdx.y * 32 + threadIdx.y;
  const int ybl = cdiv(H, CONV_ROWS);
  const int b = blockIdx.z / ybl, y0 = (blockIdx.z % ybl) * CONV_ROWS;
  if (cv >= CV || x >= W) return;
  const int c0 = cv * 4;
  float k[9][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t][i] = Kc[(c0 + i) * 9 + t];
  const T* src = raw + (long long)b * H * W * ldr + c0;
  float win[3][3][4];
  load_row3(src, ldr, W, y0 - 1, H, x, win[0]);
  load_row3(src, ldr, W, y0, H, x, win[1]);
  const int y1 = min(H, y0 + CONV_ROWS);
  for (int y = y0; y < y1; ++y) {
    load_row3(src, ldr, W, y + 1, H, x, win[2]);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = fmaf(k[r * 3 + s][i], win[r][s][i], a[i]);
    long long off = (((long long)b * H + y) * W + x) * CC + c0;
    if (pre) st4(pre + off, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = silu_t<T>(a[i]);
    st4(act + off, o);
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        win[0][s][i] = win[1][s][i];
        win[1][s][i] = win[2][s][i];
      }
  }
}

// conv backward: draw[:, :CC] = convT( dact * silu'(pre) ), dK[c][a][b] += raw[y,x,c] * dpre[y-a+1, x-b+1, c]
template <typename TW, typename T>
__device__ __forceinline__ void load_dpre_row3(const TW* __restrict__ dact, const T* __restrict__ pre, long long ld,
                                               int W, int y, int H, int x, float (&r)[3][4]) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int xx = x + s - 1;
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
      float g[4], p[4];
      long long off = ((long long)y * W + xx) * ld;
      ld4(dact + off, g);
      ld4(pre + off, p);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[s][i] = g[i] * silu_grad_t<T>(p[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[s][i] = 0.f;
    }
  }
}

template <typename T, typename TW>
__global__ void __launch_bounds__(256)
k_conv_bwd(const TW* __restrict__ dact, const T* __restrict__ pre, const T* __restrict__ raw, long long ldr,
           const float* __restrict__ Kc, TW* __restrict__ draw, float* __restrict__ dK, int H, int W, int CC) {
  __shared__ float red[8][36];
  const int CV = CC >> 2;
  const int cv = blockIdx.x * 8 + threadIdx.x;
  const int x = blockIdx.y * 32 + threadIdx.y;
  const int ybl = cdiv(H, CONV_ROWS);
  const int b = blockIdx.z / ybl, y0 = (blockIdx.z % ybl) * CONV_ROWS;
  const int tid = threadIdx.y * 8 + threadIdx.x;
  for (int i = tid; i < 8 * 36; i += 256) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const bool active = (cv < CV && x < W);
  const int c0 = cv * 4;
  float dk[9][4] = {};
  if (active) {
    float k[9][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int t = 0; t < 9; ++t) k[t][i] = Kc[(c0 + i) * 9 + t];
    const long long boff = (long long)b * H * W;
    const TW* g = dact + boff * CC + c0;
    const T* p = pre + boff * CC + c0;
    float win[3][3][4];
    load_dpre_row3(g, p, CC, W, y0 - 1, H, x, win[0]);
    load_dpre_row3(g, p, CC, W, y0, H, x, win[1]);
    const int y1 = min(H, y0 + CONV_ROWS);
    for (int y = y0; y < y1; ++y) {
      load_dpre_row3(g, p, CC, W, y + 1, H, x, win[2]);
      float rc[4], o[4] = {0.f, 0.f, 0.f, 0.f};
      long long tok = boff + (long long)y * W + x;
      ld4(raw + tok * ldr + c0, rc);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float d = win[2 - a][2 - bb][i];
            o[i] = fmaf(k[a * 3 + bb][i], d, o[i]);
            dk[a * 3 + bb][i] = fmaf(rc[i], d, dk[a * 3 + bb][i]);
          }
      st4(draw + tok * ldr + c0, o);
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          win[0][s][i] = win[1][s][i];
          win[1][s][i] = win[2][s][i];
        }
    }
  }
  // reduce over the 32 columns of the block: lanes of a warp = 4 columns x 8 channel-vectors
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = dk[t][i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((tid & 31) < 8) atomicAdd(&red[threadIdx.x][t * 4 + i], v);
    }
  __syncthreads();
  for (int i = tid; i < 8 * 36; i += 256) {
    int v = i / 36, r = i % 36, t = r >> 2, ch = (blockIdx.x * 8 + v) * 4 + (r & 3);
    if (ch < CC) atomicAdd(dK + ch * 9 + t, red[v][r]);
  }
}

// ------------------------------------------------------------------------------------------------
// decay weights: w[t,h] = softplus(raw_dt + dt_bias) * exp(A_log)  (models/ADNssd.py:318,:310,:267-270)
// and wx[t,c] = w[t,hd(c)] * xc[t,c].   thread = (token, head)
// ------------------------------------------------------------------------------------------------
template <typename T, typename TW>
__global__ void k_decay_wx(const T* __restrict__ raw, long long ldr, const T* __restrict__ act,
                           const float* __restrict__ dt_bias, const float* __restrict__ A_log,
                           float* __restrict__ wdec, TW* __restrict__ wx, long long Ttok, int nh, int P, int Di, int CC) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ttok * nh) return;
  long long t = idx / nh;
  int h = (int)(idx % nh);
  float w = softplus_t<T>(ldf(raw + t * ldr + CC + h) + dt_bias[h]) * Math<T>::exp(A_log[h]);
  if (wdec) wdec[idx] = w;
  int cb = 2 * P * (h >> 1) + (h & 1);
  for (int i = 0; i < P; ++i) {
    int c = cb + 2 * i;
    stf(wx + t * Di + c, w * ldf(act + t * CC + Di + c));
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over Di of y = ygemm + D[hd(c)] * xc  (models/ADNssd.py:283,:456), one warp per token
// ------------------------------------------------------------------------------------------------
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
k_ln_fwd(const TW* __restrict__ ygemm, const T* __restrict__ act, const float* __restrict__ Dp,
         const float* __restrict__ gamma, const float* __restrict__ beta, TW* __restrict__ yn, long long Ttok,
         int Di, int P, int CC) {
  long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (t >= Ttok) return;
  const TW* yg = ygemm + t * Di;
  const T* xc = act + t * CC + Di;
  float s = 0.f;
  for (int c = lane; c < Di; c += 32) s += ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c);
  float mu = warp_sum(s) / Di;
  float v = 0.f;
  for (int c = lane; c < Di; c += 32) {
    float d = ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c) - mu;
    v += d * d;
  }
  float rstd = rsqrtf(warp_sum(v) / Di + 1e-5f);
  for (int c = lane; c < Di; c += 32) {
    float y = ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c);
    stf(yn + t * Di + c, (y - mu) * rstd * gamma[c] + beta[c]);
  }
}

// backward of: out = alpha1 * [LN(y) | zc] @ W_out^T.  g = dout @ W_out (no alpha1).  Writes yn (for dW_out),
// dy -> dact[:, Di:2Di], dzc -> dact[:, :Di]; accumulates dgamma, dbeta, dalpha1.
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
k_ln_bwd(const TW* __restrict__ ygemm, const T* __restrict__ act, const TW* __restrict__ g,
         const float* __restrict__ Dp, const float* __restrict__ gamma, const float* __restrict__ beta,
         const float* __restrict__ alpha1p, TW* __restrict__ yn, TW* __restrict__ dact, float* __restrict__ dgamma,
         float* __restrict__ dbeta, float* __restrict__ dalpha1, long long Ttok, int tokens_per_warp, int Di, int P,
         int CC) {
  extern __shared__ float sm[];  // [2*Di] block-local dgamma / dbeta
  float* sg = sm;
  float* sb = sm + Di;
  for (int i = threadIdx.x; i < 2 * Di; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const float a1 = *alpha1p;
  const int lane = threadIdx.x & 31;
  long long t0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * tokens_per_warp;
  // d alpha1 = <g, [yn | zc]> is one scalar summed over every token and channel with heavy cancellation: accumulated in
  // fp64 end to end (the 8-byte accumulator slot is read back as a double by finalize_body)
  double da = 0.0;
  for (long long t = t0; t < min(Ttok, t0 + tokens_per_warp); ++t) {
    const TW* yg = ygemm + t * Di;
    const T* zc = act + t * CC;
    const T* xc = zc + Di;
    const TW* gy = g + t * 2 * Di;
    const TW* gz = gy + Di;
    float s = 0.f;
    for (int c = lane; c < Di; c += 32) s += ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c);
    float mu = warp_sum(s) / Di;
    float v = 0.f;
    for (int c = lane; c < Di; c += 32) {
      float d = ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c) - mu;
      v += d * d;
    }
    float rstd = rsqrtf(warp_sum(v) / Di + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
    for (int c = lane; c < Di; c += 32) {
      float yh = (ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c) - mu) * rstd;
      float gyc = ldf(gy + c), gzc = ldf(gz + c), z = ldf(zc + c);
      float ynv = yh * gamma[c] + beta[c];
      stf(yn + t * Di + c, ynv);
      da += (double)(gyc * ynv + gzc * z);
      float dyn = a1 * gyc;
      atomicAdd(sg + c, dyn * yh);
      atomicAdd(sb + c, dyn);
      float dyh = dyn * gamma[c];
      m1 += dyh;
      m2 += dyh * yh;
      stf(dact + t * CC + c, a1 * gzc);
    }
    m1 = warp_sum(m1) / Di;
    m2 = warp_sum(m2) / Di;
    for (int c = lane; c < Di; c += 32) {
      float yh = (ldf(yg + c) + Dp[head_of_channel(c, P)] * ldf(xc + c) - mu) * rstd;
      float dyh = a1 * ldf(gy + c) * gamma[c];
      stf(dact + t * CC + Di + c, rstd * (dyh - m1 - yh * m2));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
  if (lane == 0 && da != 0.0) atomicAdd(reinterpret_cast<double*>(dalpha1), da);
  __syncthreads();
  for (int i = threadIdx.x; i < Di; i += blockDim.x) {
    atomicAdd(dgamma + i, sg[i]);
    atomicAdd(dbeta + i, sb[i]);
  }
}

// per (token, head): dxc = D*dy + w*G (in place over dact x block), wx = w*xc (in place over G),
// ddt -> draw[:, CC+h]; accumulates dD, dA_log, ddt_bias.   block (32 heads, 8 tokens)
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
k_bwd_heads(const T* __restrict__ raw, long long ldr, const T* __restrict__ act, const float* __restrict__ wdec,
            const float* __restrict__ dt_bias, const float* __restrict__ A_log, const float* __restrict__ Dp,
            TW* __restrict__ dact, TW* __restrict__ Gwx, TW* __restrict__ draw, float* __restrict__ dD,
            float* __restrict__ dAlog, float* __restrict__ ddtb, long long Ttok, int tokens_per_thread, int nh, int P,
            int Di, int CC) {
  __shared__ float red[3][8][32];
  const int h = blockIdx.x * 32 + threadIdx.x;
  float aD = 0.f, aA = 0.f, aB = 0.f;
  if (h < nh) {
    const float Dh = Dp[h], eA = Math<T>::exp(A_log[h]), bias = dt_bias[h];
    const int cb = 2 * P * (h >> 1) + (h & 1);
    long long t0 = ((long long)blockIdx.y * 8 + threadIdx.y) * tokens_per_thread;
    for (long long t = t0; t < min(Ttok, t0 + tokens_per_thread); ++t) {
      float w = wdec[t * nh + h];
      float dw = 0.f;
      for (int i = 0; i < P; ++i) {
        int c = cb + 2 * i;
        float dy = ldf(dact + t * CC + Di + c), G = ldf(Gwx + t * Di + c), x = ldf(act + t * CC + Di + c);
        stf(dact + t * CC + Di + c, Dh * dy + w * G);
        stf(Gwx + t * Di + c, w * x);
        dw += x * G;
        aD += dy * x;
      }
      float ddt = dw * eA * sigmoid_t<T>(ldf(raw + t * ldr + CC + h) + bias);
      stf(draw + t * ldr + CC + h, ddt);
      aA += dw * w;
      aB += ddt;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = aD;
  red[1][threadIdx.y][threadIdx.x] = aA;
  red[2][threadIdx.y][threadIdx.x] = aB;
  __syncthreads();
  if (threadIdx.y < 3 && h < nh) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[threadIdx.y][j][threadIdx.x];
    float* dst = threadIdx.y == 0 ? dD : (threadIdx.y == 1 ? dAlog : ddtb);
    atomicAdd(dst + h, s);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: fp32 accumulators -> caller's gradient tensors (overwrite), incl. the rank-1 chain rule of the
// 3x1 / 1x3 pairs: dw31[a] = sum_b dK[a,b] w13[b], dw13[b] = sum_a dK[a,b] w31[a]
// ------------------------------------------------------------------------------------------------
struct GradAcc {
  float *dWin, *dWout, *dgamma, *dbeta, *dD, *dAlog, *ddtb, *dalpha1, *dK;
  int dalpha1_f64;      // the dalpha1 slot (8 bytes) holds a double (generic k_ln_bwd) instead of a float
  // optional per-CTA partial sums of dW_in (the tcgen05 path writes one slab per CTA instead of contended atomics);
  // finalize adds dWin_parts slabs of dip*D floats starting at dWin_part to dWin
  const float* dWin_part;
  int dWin_parts;
  // optional per-CTA partial sums of dK (k_bconv_wg): dK_parts slabs of CC*9 floats; dK itself is then unused
  const float* dK_part;
  int dK_parts, dK_stride;
  int* sync_counter;   // zeroed with the accumulators; grid-level hand-off inside k_finalize_fast
  // optional per-CTA partial sums of [dD | dA_log | ddt_bias] (k_bwd2_ws): head_parts slabs of 3*nh floats
  const float* head_part;
  int head_parts;
};

__device__ __forceinline__ float dk_at(const GradAcc& a, int idx) {
  if (a.dK_parts == 0) return __ldcg(a.dK + idx);     // may have been reduced by other SMs earlier in the same kernel
  float v = 0.f;
  for (int p = 0; p < a.dK_parts; ++p) v += a.dK_part[(long long)p * a.dK_stride + idx];
  return v;
}

// `proj` = also emit in_proj / out_proj / norm / alpha1 gradients from the accumulators (the tcgen05 path produces those
// in its own fused finalize and passes proj = false)
__device__ __forceinline__ void finalize_body(const GradAcc& a, const AdnWeights& w, const AdnWeightGrads& g, int D, int Di,
                                              int GN, int nh, int dip, long long i0, long long stride, bool proj) {
  const float a1 = *w.alpha1;
  if (proj) {
  if (g.in_proj_w)
    for (long long i = i0; i < (long long)dip * D; i += stride) {
      float v = a.dWin[i];
      for (int p = 0; p < a.dWin_parts; ++p) v += a.dWin_part[(long long)p * dip * D + i];
      g.in_proj_w[i] = v;
    }
  if (g.out_proj_w)
    for (long long i = i0; i < (long long)D * 2 * Di; i += stride) g.out_proj_w[i] = a1 * a.dWout[i];
  for (long long i = i0; i < Di; i += stride) {
    if (g.norm_w) g.norm_w[i] = a.dgamma[i];
    if (g.norm_b) g.norm_b[i] = a.dbeta[i];
  }
  if (i0 == 0 && g.alpha1) g.alpha1[0] = a.dalpha1_f64 ? (float)*reinterpret_cast<const double*>(a.dalpha1) : a.dalpha1[0];
  }
  for (long long i = i0; i < nh; i += stride) {
    float vD = __ldcg(a.dD + i), vA = __ldcg(a.dAlog + i), vB = __ldcg(a.ddtb + i);
    for (int p = 0; p < a.head_parts; ++p) {
      vD += a.head_part[(long long)p * 3 * nh + i];
      vA += a.head_part[(long long)p * 3 * nh + nh + i];
      vB += a.head_part[(long long)p * 3 * nh + 2 * nh + i];
    }
    if (g.D) g.D[i] = vD;
    if (g.A_log) g.A_log[i] = vA;
    if (g.dt_bias) g.dt_bias[i] = vB;
  }
  if (g.conv2d_z_w)
    for (long long i = i0; i < (long long)Di * 9; i += stride) g.conv2d_z_w[i] = dk_at(a, (int)i);
  const int Wd = Di + 2 * GN;
  if (g.conv2d_w)
    for (long long i = i0; i < (long long)(Wd / 2) * 9; i += stride)
      g.conv2d_w[i] = dk_at(a, (int)((Di + 2 * (i / 9)) * 9 + (i % 9)));
  // pairs: index space (tag in {0,1}) x (Wd/4 groups) x 3 taps
  const int ng = Wd / 4, nx = Di / 4;
  for (long long i = i0; i < 2LL * ng * 3; i += stride) {
    int tap = (int)(i % 3), gi = (int)((i / 3) % ng), tag = (int)(i / (3 * ng));
    int ch = Di + 4 * gi + (tag == 0 ? 1 : 3);
    float dk[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) dk[q] = dk_at(a, ch * 9 + q);
    const float *w31, *w13;
    float *g31, *g13;
    if (gi < nx) {
      w31 = (tag == 0 ? w.conv_31_x1_w : w.conv_31_x2_w) + gi * 3;
      w13 = (tag == 0 ? w.conv_13_x1_w : w.conv_13_x2_w) + gi * 3;
      g31 = (tag == 0 ? g.conv_31_x1_w : g.conv_31_x2_w);
      g13 = (tag == 0 ? g.conv_13_x1_w : g.conv_13_x2_w);
      if (g31) g31 += gi * 3;
      if (g13) g13 += gi * 3;
    } else {
      int j = gi - nx;
      w31 = (tag == 0 ? w.conv_31_bc1_w : w.conv_31_bc2_w) + j * 3;
      w13 = (tag == 0 ? w.conv_13_bc1_w : w.conv_13_bc2_w) + j * 3;
      g31 = (tag == 0 ? g.conv_31_bc1_w : g.conv_31_bc2_w);
      g13 = (tag == 0 ? g.conv_13_bc1_w : g.conv_13_bc2_w);
      if (g31) g31 += j * 3;
      if (g13) g13 += j * 3;
    }
    if (g31) g31[tap] = dk[tap * 3 + 0] * w13[0] + dk[tap * 3 + 1] * w13[1] + dk[tap * 3 + 2] * w13[2];
    if (g13) g13[tap] = dk[0 * 3 + tap] * w31[0] + dk[1 * 3 + tap] * w31[1] + dk[2 * 3 + tap] * w31[2];
  }
}

static __global__ void k_finalize(GradAcc a, AdnWeights w, AdnWeightGrads g, int D, int Di, int GN, int nh, int dip) {
  finalize_body(a, w, g, D, Di, GN, nh, dip, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                (long long)gridDim.x * blockDim.x, true);
}

// ------------------------------------------------------------------------------------------------
// buffer carving
// ------------------------------------------------------------------------------------------------
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base((char*)p), off(0) {}
  template <typename U>
  U* take(size_t n) {
    U* r = base ? (U*)(base + off) : nullptr;
    off += align_up(n * sizeof(U), 256);
    return r;
  }
};

template <typename T>
struct SavedBufs {
  T *raw, *pre, *act;
  float *wdec, *S;
  size_t bytes;
  SavedBufs(const MixerDims& d, void* p) {
    Carver c(p);
    raw = c.take<T>((size_t)d.T * d.ldr);
    pre = c.take<T>((size_t)d.T * d.CC);
    act = c.take<T>((size_t)d.T * d.CC);
    wdec = c.take<float>((size_t)d.T * d.nh);
    S = c.take<float>((size_t)d.B * d.GN * d.Di);
    bytes = c.off;
  }
};

// Workspace (never saved) intermediates are fp32 in both modes: bf16 rounding is confined to the I/O and saved tensors.
typedef float TWs;

template <typename T, typename TW = TWs>
struct FwdWs {
  float* Kc;
  TW *bufA, *bufB;  // (T, Di) each: wx then y-gemm ; yn
  SavedBufs<T> tmp;  // used when the caller passes saved == NULL (inference)
  size_t bytes;
  FwdWs(const MixerDims& d, void* p) : tmp(d, nullptr) {
    Carver c(p);
    Kc = c.take<float>((size_t)d.CC * 9);
    bufA = c.take<TW>((size_t)d.T * d.Di);
    bufB = c.take<TW>((size_t)d.T * d.Di);
    size_t here = c.off;
    tmp = SavedBufs<T>(d, p ? (char*)p + here : nullptr);
    bytes = here + tmp.bytes;
  }
};

template <typename T, typename TW = TWs>
struct BwdWs {
  float* Kc;
  float* zero_begin;
  GradAcc acc;
  float* dS;
  int* status;
  size_t zero_bytes;
  TW *g, *ybuf, *ynbuf, *dact, *draw;
  size_t bytes;
  BwdWs(const MixerDims& d, void* p) {
    Carver c(p);
    Kc = c.take<float>((size_t)d.CC * 9);
    size_t z0 = c.off;
    zero_begin = p ? (float*)((char*)p + z0) : nullptr;
    acc.dWin = c.take<float>((size_t)d.dip * d.D);
    acc.dWin_part = nullptr;
    acc.dWin_parts = 0;
    acc.dK_part = nullptr;
    acc.dK_parts = 0;
    acc.dK_stride = 0;
    acc.head_part = nullptr;
    acc.head_parts = 0;
    acc.dWout = c.take<float>((size_t)d.D * 2 * d.Di);
    acc.dgamma = c.take<float>(d.Di);
    acc.dbeta = c.take<float>(d.Di);
    acc.dD = c.take<float>(d.nh);
    acc.dAlog = c.take<float>(d.nh);
    acc.ddtb = c.take<float>(d.nh);
    acc.dalpha1 = c.take<float>(2);
    acc.dalpha1_f64 = 1;
    acc.dK = c.take<float>((size_t)d.CC * 9);
    acc.sync_counter = c.take<int>(64);
    status = c.take<int>(64);      // pipeline-fault word of the backward pass, zeroed with the accumulators
    dS = c.take<float>((size_t)d.B * d.GN * d.Di);
    zero_bytes = c.off - z0;
    g = c.take<TW>((size_t)d.T * 2 * d.Di);
    ybuf = c.take<TW>((size_t)d.T * d.Di);
    ynbuf = c.take<TW>((size_t)d.T * d.Di);
    dact = c.take<TW>((size_t)d.T * d.CC);
    draw = c.take<TW>((size_t)d.T * d.ldr);
    bytes = c.off;
  }
};

static inline ConvWeightPtrs conv_ptrs(const AdnWeights& w) {
  ConvWeightPtrs c;
  c.c13x1 = w.conv_13_x1_w; c.c31x1 = w.conv_31_x1_w; c.c13x2 = w.conv_13_x2_w; c.c31x2 = w.conv_31_x2_w;
  c.c13bc1 = w.conv_13_bc1_w; c.c31bc1 = w.conv_31_bc1_w; c.c13bc2 = w.conv_13_bc2_w; c.c31bc2 = w.conv_31_bc2_w;
  c.c2d = w.conv2d_w; c.c2dz = w.conv2d_z_w;
  return c;
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
template <typename T>
int generic_forward(const MixerDims& d, const AdnWeights& w, const T* u, T* out, void* saved, void* ws,
                    cudaStream_t st) {
  FwdWs<T> W(d, ws);
  SavedBufs<T> S = saved ? SavedBufs<T>(d, saved) : W.tmp;
  const bool training = saved != nullptr;
  const long long Tt = d.T;
  { ADN_KERNEL("k_assemble_conv", st); k_assemble_conv<<<cdiv(d.CC, 128), 128, 0, st>>>(conv_ptrs(w), W.Kc, d.Di, d.CC); }
  // (1) in_proj  (models/ADNssd.py:309)
  launch_gemm<T, T, true>(st, u, d.D, 0, w.in_proj_w, d.D, 0, S.raw, d.ldr, 0, (int)Tt, d.dip, d.D, 1, nullptr, 0);
  // (2) depthwise 3x3 + SiLU over [z | x | B | C]
  {
    dim3 grid(cdiv(d.CC / 4, 8), cdiv(d.W, 32), d.B * cdiv(d.H, CONV_ROWS)), block(8, 32);
    { ADN_KERNEL("k_conv_fwd", st); k_conv_fwd<T><<<grid, block, 0, st>>>(S.raw, d.ldr, W.Kc, training ? S.pre : nullptr, S.act, d.H, d.W, d.CC); }
  }
  // (3) decay weights and w*x
  { ADN_KERNEL("k_decay_wx", st); k_decay_wx<T, TWs><<<cdiv(Tt * d.nh, 256), 256, 0, st>>>(S.raw, d.ldr, S.act, w.dt_bias, w.A_log, S.wdec, W.bufA, Tt,
                                                      d.nh, d.P, d.Di, d.CC); }
  // (4a) state S'[b] = mask . Bc^T (w x)
  ADN_CHECK_CUDA(cudaMemsetAsync(S.S, 0, (size_t)d.B * d.GN * d.Di * sizeof(float), st));
  launch_reduce_gemm<T, TWs>(st, S.act + 2 * d.Di, d.CC, W.bufA, d.Di, S.S, d.Di, (long long)d.GN * d.Di, d.GN, d.Di,
                           d.L, d.B, 1);
  // (4b) readout y = Cc S'  (D-skip is added inside the LayerNorm kernel)
  launch_gemm<T, TWs, false>(st, S.act + 2 * d.Di + d.GN, d.CC, (long long)d.L * d.CC, S.S, d.Di,
                           (long long)d.GN * d.Di, W.bufA, d.Di, (long long)d.L * d.Di, d.L, d.Di, d.GN, d.B, nullptr,
                           0);
  // (5) D-skip + LayerNorm, then out = alpha1 * (yn Wy^T + zc Wz^T)
  { ADN_KERNEL("k_ln_fwd", st); k_ln_fwd<T, TWs><<<cdiv(Tt, 8), 256, 0, st>>>(W.bufA, S.act, w.D, w.norm_w, w.norm_b, W.bufB, Tt, d.Di, d.P, d.CC); }
  launch_gemm<TWs, T, true>(st, W.bufB, d.Di, 0, w.out_proj_w, 2 * d.Di, 0, out, d.D, 0, (int)Tt, d.D, d.Di, 1,
                          w.alpha1, 0);
  launch_gemm<T, T, true>(st, S.act, d.CC, 0, w.out_proj_w + d.Di, 2 * d.Di, 0, out, d.D, 0, (int)Tt, d.D, d.Di, 1,
                          w.alpha1, 1);
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

template <typename T>
int generic_backward(const MixerDims& d, const AdnWeights& w, const T* u, const void* saved, const T* dout, T* du,
                     const AdnWeightGrads& g, void* ws, cudaStream_t st) {
  BwdWs<T> W(d, ws);
  SavedBufs<T> S(d, const_cast<void*>(saved));
  const long long Tt = d.T;
  const long long sS = (long long)d.GN * d.Di;
  ADN_CHECK_CUDA(cudaMemsetAsync(W.zero_begin, 0, W.zero_bytes, st));
  { ADN_KERNEL("k_assemble_conv", st); k_assemble_conv<<<cdiv(d.CC, 128), 128, 0, st>>>(conv_ptrs(w), W.Kc, d.Di, d.CC); }
  // ---- phase B1
  launch_gemm<T, TWs, false>(st, dout, d.D, 0, w.out_proj_w, 2 * d.Di, 0, W.g, 2 * d.Di, 0, (int)Tt, 2 * d.Di, d.D, 1,
                           nullptr, 0);
  launch_gemm<T, TWs, false>(st, S.act + 2 * d.Di + d.GN, d.CC, (long long)d.L * d.CC, S.S, d.Di, sS, W.ybuf, d.Di,
                           (long long)d.L * d.Di, d.L, d.Di, d.GN, d.B, nullptr, 0);
  {
    int tpw = 4;
    { ADN_KERNEL("k_ln_bwd", st); k_ln_bwd<T, TWs><<<cdiv(Tt, 8 * tpw), 256, 2 * d.Di * sizeof(float), st>>>(
        W.ybuf, S.act, W.g, w.D, w.norm_w, w.norm_b, w.alpha1, W.ynbuf, W.dact, W.acc.dgamma, W.acc.dbeta,
        W.acc.dalpha1, Tt, tpw, d.Di, d.P, d.CC); }
  }
  launch_reduce_gemm<T, TWs>(st, dout, d.D, W.ynbuf, d.Di, W.acc.dWout, 2 * d.Di, 0, d.D, d.Di, (int)Tt, 1, 0);
  launch_reduce_gemm<T, T>(st, dout, d.D, S.act, d.CC, W.acc.dWout + d.Di, 2 * d.Di, 0, d.D, d.Di, (int)Tt, 1, 0);
  launch_reduce_gemm<T, TWs>(st, S.act + 2 * d.Di + d.GN, d.CC, W.dact + d.Di, d.CC, W.dS, d.Di, sS, d.GN, d.Di, d.L,
                           d.B, 1);
  launch_gemm<TWs, TWs, true>(st, W.dact + d.Di, d.CC, (long long)d.L * d.CC, S.S, d.Di, sS, W.dact + 2 * d.Di + d.GN,
                          d.CC, (long long)d.L * d.CC, d.L, d.GN, d.Di, d.B, nullptr, 0);
  // ---- phase B2
  launch_gemm<T, TWs, false>(st, S.act + 2 * d.Di, d.CC, (long long)d.L * d.CC, W.dS, d.Di, sS, W.ybuf, d.Di,
                           (long long)d.L * d.Di, d.L, d.Di, d.GN, d.B, nullptr, 0);
  {
    int tpt = 8;
    dim3 grid(cdiv(d.nh, 32), cdiv(Tt, 8 * tpt)), block(32, 8);
    { ADN_KERNEL("k_bwd_heads", st); k_bwd_heads<T, TWs><<<grid, block, 0, st>>>(S.raw, d.ldr, S.act, S.wdec, w.dt_bias, w.A_log, w.D, W.dact, W.ybuf,
                                           W.draw, W.acc.dD, W.acc.dAlog, W.acc.ddtb, Tt, tpt, d.nh, d.P, d.Di, d.CC); }
  }
  launch_gemm<TWs, TWs, true>(st, W.ybuf, d.Di, (long long)d.L * d.Di, W.dS, d.Di, sS, W.dact + 2 * d.Di, d.CC,
                          (long long)d.L * d.CC, d.L, d.GN, d.Di, d.B, nullptr, 0);
  // ---- conv backward
  {
    dim3 grid(cdiv(d.CC / 4, 8), cdiv(d.W, 32), d.B * cdiv(d.H, CONV_ROWS)), block(8, 32);
    { ADN_KERNEL("k_conv_bwd", st); k_conv_bwd<T, TWs><<<grid, block, 0, st>>>(W.dact, S.pre, S.raw, d.ldr, W.Kc, W.draw, W.acc.dK, d.H, d.W, d.CC); }
  }
  // ---- in_proj backward
  launch_gemm<TWs, T, false>(st, W.draw, d.ldr, 0, w.in_proj_w, d.D, 0, du, d.D, 0, (int)Tt, d.D, d.dip, 1, nullptr, 0);
  launch_reduce_gemm<TWs, T>(st, W.draw, d.ldr, u, d.D, W.acc.dWin, d.D, 0, d.dip, d.D, (int)Tt, 1, 0);
  { ADN_KERNEL("k_finalize", st); k_finalize<<<sm_count(), 256, 0, st>>>(W.acc, w, g, d.D, d.Di, d.GN, d.nh, d.dip); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // namespace adn
