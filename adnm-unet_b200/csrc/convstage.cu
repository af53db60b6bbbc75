// The full-resolution conv stages that bracket the Refiner: WTLayer (models/model_untils.py:358-426),
// PatchEmbed (:226-314) and OutProj (:799-892).  Around the native WTConv2d (wtconv.cu) they consist of
//   * a dense 3x3 convolution (Conv2dLayer, :71-93)                                   -> adn_conv3x3_*
//   * InstanceNorm2d * scale + shift [+ GELU], the alpha / beta shortcut mix and the layer-scale gamma
//     (WTConvLayer :96-116, WTLayer :416-421, PatchEmbed :303-307, OutProj :879-883)   -> adn_plane_stats, adn_plane_mix_*
//   * the (B, L, C) <-> (B, C, H, W) layout changes, the gama1 / gama2 skip concat (:404-414) -> adn_nchw_pack_*
//   * GELU / Swish after the convs                                                      -> adn_act_*
// and an Mlp (two Linears, adn_linear_* in block.cu).
//
// Dense 3x3 convolution.  Activations are channels-last = token-major (B, L, C), the layout every neighbour (Block, Mlp,
// DownSample / UpSample inputs) already uses, so the reference's permutes around the conv do not exist.  bf16 with
// Cin % 8 == 0, Cout % 8 == 0 and a power-of-two-ish grid: implicit GEMM on the tensor cores (tcgemm.cuh `Conv`: the image is
// an operand read through a rank-4 TMA map, a tap is a box shifted by (dx, dy), the zero padding is the TMA zero fill) -
//   forward        y[T][Cout]       = sum_tap x[T + tap][Cin]  . Wf[Cout][tap, Cin]^T    (+ bias in the epilogue)
//   data gradient  dx[T][Cin]       = sum_tap dy[T + tap][Cout] . Wd[Cin][tap, Cout]^T   (taps flipped)
//   weight grad.   dWt[Cout][tap, Cin] = sum_T dy[T][Cout]^T . x[T + tap][Cin]           (split-K over tokens, TMA reduce-add)
// The layer scale gamma[Cin] that precedes the conv in WTLayer / OutProj (x.mul(gamma), :420-421,:882-883) is folded into
// the weight images (zero padding commutes with a per-channel scale), so the scaled activation is never written:
//   d gamma[ci] = sum_{co,tap} dWt[co][tap][ci] w[co][ci][tap],   dw = dWt * gamma,   dx comes out of Wd already scaled.
// fp32 activations (1e-4 check mode) and the two thin convs (PatchEmbed 5 -> 32, OutProj 20 -> 20) run CUDA-core kernels.
#include "adn_common.cuh"
#include "tcgemm.cuh"

namespace adn {
namespace cst {

static inline int ew_grid(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
// Storage-type dispatched GELU: fp32 (the 1e-4 check mode) keeps erff / expf; bf16 uses Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7 on erf, far below the bf16 rounding of the stored result) with ONE fast exponential shared between erf and the
// Gaussian of the derivative (exp(-z^2) with z = x / sqrt 2 is exp(-x^2 / 2)): the exact pair kept k_act_* at 88 % issue slots.
template <typename T> struct Act {
  static __device__ __forceinline__ float gelu(float x) { return gelu_f(x); }
  static __device__ __forceinline__ float gelu_grad(float x) { return gelu_grad_f(x); }
};
__device__ __forceinline__ void erf_as(float x, float& erf_z, float& gauss) {      // erf(x / sqrt 2), exp(-x^2 / 2)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  gauss = __expf(-z * z);
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  erf_z = copysignf(1.f - poly * gauss, x);
}
template <> struct Act<bf16> {
  static __device__ __forceinline__ float gelu(float x) { float e, g; erf_as(x, e, g); return 0.5f * x * (1.f + e); }
  static __device__ __forceinline__ float gelu_grad(float x) { float e, g; erf_as(x, e, g); return 0.5f * (1.f + e) + x * 0.3989422804014327f * g; }
};

// ================================================================ dense 3x3 convolution
struct ConvDims {
  int B, H, W, Cin, Cout;
  int ci8, co8;          // channel counts rounded up to 8: the tensor-core path needs 16-byte rows; thin tensors (PatchEmbed 5 -> 32,
                         // OutProj 20 -> 20) are copied into zero-padded rows in the workspace on the way in and compacted on the way out
  int cpi, cpo;          // channels rounded up to 64 (one swizzle group per tap chunk)
  long long T;
  bool tc;
};

static int conv_dims(const AdnConvShape* s, ConvDims* d, const char* what) {
  ADN_REQUIRE(s != nullptr, ADN_ERR_NULL, "%s: NULL shape", what);
  ADN_REQUIRE(s->B > 0 && s->H > 0 && s->W > 0 && s->Cin > 0 && s->Cout > 0, ADN_ERR_SHAPE, "%s: bad extents", what);
  ADN_REQUIRE(s->dtype == ADN_F32 || s->dtype == ADN_BF16, ADN_ERR_DTYPE, "%s: unsupported dtype %d", what, s->dtype);
  d->B = s->B; d->H = s->H; d->W = s->W; d->Cin = s->Cin; d->Cout = s->Cout;
  d->ci8 = (s->Cin + 7) / 8 * 8; d->co8 = (s->Cout + 7) / 8 * 8;
  d->cpi = (s->Cin + 63) / 64 * 64; d->cpo = (s->Cout + 63) / 64 * 64;
  d->T = (long long)s->B * s->H * s->W;
  ADN_REQUIRE(d->T < (1LL << 31) / 16, ADN_ERR_SHAPE, "%s: too many tokens", what);
  d->tc = s->dtype == ADN_BF16 && s->Cin <= 256 && s->Cout <= 256 && tcg::image_tiles(s->H, s->W) && env().wide;
  return ADN_OK;
}

// workspace layout (both passes): [status 256 B][Wf bf16 co8 x 9 cpi][Wd bf16 ci8 x 9 cpo][dWt fp32 co8 x 9 cpi][bias fp32 co8]
//                                 [xp bf16 T x ci8][yp bf16 T x co8][dxp bf16 T x ci8]   (the last three only for padded channel counts)
struct ConvWs {
  int* status; bf16* Wf; bf16* Wd; float* dWt; float* biasp; bf16 *xp, *yp, *dxp;
  size_t bytes;
};
static ConvWs conv_ws(const ConvDims& d, void* base) {
  ConvWs w;
  char* p = (char*)base;
  size_t off = 0;
  w.status = (int*)(p + off); off += 256;
  w.Wf = (bf16*)(p + off); off += align_up((size_t)d.co8 * 9 * d.cpi * sizeof(bf16), 256);
  w.Wd = (bf16*)(p + off); off += align_up((size_t)d.ci8 * 9 * d.cpo * sizeof(bf16), 256);
  w.dWt = (float*)(p + off); off += align_up((size_t)d.co8 * 9 * d.cpi * sizeof(float), 256);
  w.biasp = (float*)(p + off); off += align_up((size_t)d.co8 * sizeof(float), 256);
  w.xp = w.yp = w.dxp = nullptr;
  if (d.tc && d.ci8 != d.Cin) {
    w.xp = (bf16*)(p + off); off += align_up((size_t)d.T * d.ci8 * sizeof(bf16), 256);
    w.dxp = (bf16*)(p + off); off += align_up((size_t)d.T * d.ci8 * sizeof(bf16), 256);
  }
  if (d.tc && d.co8 != d.Cout) { w.yp = (bf16*)(p + off); off += align_up((size_t)d.T * d.co8 * sizeof(bf16), 256); }
  w.bytes = off;
  return w;
}

// dst[t][0 .. Cd) = src[t][0 .. min(Cs, Cd)), zero beyond Cs: pads thin rows to whole 16-byte pieces / compacts them again
template <typename T>
__global__ void __launch_bounds__(256)
k_copy_rows(const T* __restrict__ src, int Cs, T* __restrict__ dst, int Cd, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long t = i / Cd;
    const int c = (int)(i - t * Cd);
    dst[i] = c < Cs ? src[t * Cs + c] : T(0.f);
  }
}

// weight images of the implicit GEMMs from the state_dict layout w[Cout][Cin][3][3] (gamma folded in, pad rows / channels zero)
__global__ void k_conv_wprep(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ bias, bf16* __restrict__ Wf,
                             bf16* __restrict__ Wd, float* __restrict__ biasp, int Cin, int Cout, int ci8, int co8, int cpi, int cpo) {
  const int nf = co8 * 9 * cpi, nd = ci8 * 9 * cpo;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf + nd + co8; i += gridDim.x * blockDim.x) {
    if (i < nf) {
      const int co = i / (9 * cpi), r = i - co * 9 * cpi, t = r / cpi, ci = r - t * cpi;
      const float v = (ci < Cin && co < Cout) ? w[((long long)co * Cin + ci) * 9 + t] * (gamma ? gamma[ci] : 1.f) : 0.f;
      Wf[i] = __float2bfloat16_rn(v);
    } else if (i < nf + nd) {
      const int j = i - nf, ci = j / (9 * cpo), r = j - ci * 9 * cpo, t = r / cpo, co = r - t * cpo;
      const float v = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * 9 + (8 - t)] * (gamma ? gamma[ci] : 1.f) : 0.f;
      Wd[j] = __float2bfloat16_rn(v);
    } else {
      const int co = i - nf - nd;
      biasp[co] = (bias && co < Cout) ? bias[co] : 0.f;
    }
  }
}

// dWt[Cout][9][cpi] (gradient w.r.t. the gamma-scaled weights) -> dw[Cout][Cin][3][3], dgamma[Cin] (zeroed by the caller)
__global__ void k_conv_wfinal(const float* __restrict__ dWt, const float* __restrict__ w, const float* __restrict__ gamma,
                              float* __restrict__ dw, float* __restrict__ dgamma, int Cin, int Cout, int cpi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin) return;
  const int co = i / Cin, ci = i - co * Cin;
  const float g = gamma ? gamma[ci] : 1.f;
  float dg = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float v = dWt[((long long)co * 9 + t) * cpi + ci];
    dw[(long long)i * 9 + t] = v * g;
    dg += v * w[(long long)i * 9 + t];
  }
  if (dgamma) atomicAdd(dgamma + ci, dg);
}

// ---- CUDA-core kernels (fp32 check mode, thin channel counts): one thread per output element, fp32 accumulation
template <typename T>
__global__ void __launch_bounds__(256)
k_conv3_fwd_direct(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ bias,
                   T* __restrict__ y, int H, int W, int Cin, int Cout, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long tok = i / Cout;
    const int co = (int)(i - tok * Cout);
    const int xx = (int)(tok % W), yy = (int)((tok / W) % H);
    float acc = bias ? bias[co] : 0.f;
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      if (yy + dy < 0 || yy + dy >= H || xx + dx < 0 || xx + dx >= W) continue;
      const T* xp = x + (tok + dy * W + dx) * Cin;
      const float* wp = w + (long long)co * Cin * 9 + t;
      if (gamma) { for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ldf(xp + ci) * gamma[ci], wp[ci * 9], acc); }
      else { for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ldf(xp + ci), wp[ci * 9], acc); }
    }
    stf(y + i, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_conv3_dgrad_direct(const T* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ gamma, T* __restrict__ dx,
                     int H, int W, int Cin, int Cout, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long tok = i / Cin;
    const int ci = (int)(i - tok * Cin);
    const int xx = (int)(tok % W), yy = (int)((tok / W) % H);
    float acc = 0.f;
    for (int t = 0; t < 9; ++t) {
      const int dyy = t / 3 - 1, dxx = t % 3 - 1;      // output position p = q + delta_t read input q  ->  q = p: dy at p - delta
      if (yy - dyy < 0 || yy - dyy >= H || xx - dxx < 0 || xx - dxx >= W) continue;
      const T* gp = dy + (tok - dyy * W - dxx) * Cout;
      const float* wp = w + (long long)ci * 9 + t;
      for (int co = 0; co < Cout; ++co) acc = fmaf(ldf(gp + co), wp[(long long)co * Cin * 9], acc);
    }
    stf(dx + i, acc * (gamma ? gamma[ci] : 1.f));
  }
}

// dWt[co][t][ci] += sum_{tokens of this block} dy[p][co] x[p + delta_t][ci]
template <typename T>
__global__ void __launch_bounds__(256)
k_conv3_wgrad_direct(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dWt, int H, int W, int Cin, int Cout, int cpi,
                     long long Ttok, int tpb) {
  const long long t0 = (long long)blockIdx.x * tpb, t1 = min(Ttok, t0 + (long long)tpb);
  const int entries = Cout * 9 * Cin;
  for (int e = threadIdx.x; e < entries; e += blockDim.x) {
    const int ci = e % Cin, r = e / Cin, t = r % 9, co = r / 9;      // ci fastest: neighbouring threads read neighbouring x
    const int dyy = t / 3 - 1, dxx = t % 3 - 1;
    float acc = 0.f;
    for (long long p = t0; p < t1; ++p) {
      const int xx = (int)(p % W), yy = (int)((p / W) % H);
      if (yy + dyy < 0 || yy + dyy >= H || xx + dxx < 0 || xx + dxx >= W) continue;
      acc = fmaf(ldf(dy + p * Cout + co), ldf(x + (p + dyy * W + dxx) * Cin + ci), acc);
    }
    if (acc != 0.f) atomicAdd(dWt + ((long long)co * 9 + t) * cpi + ci, acc);
  }
}

// out[n] += sum_t X[t][n]   (out zeroed by the caller); lanes stride over columns, warps over rows
template <typename T>
__global__ void __launch_bounds__(256)
k_colsum_any(const T* __restrict__ X, int N, long long Ttok, int tpb, float* __restrict__ out) {
  const long long t0 = (long long)blockIdx.x * tpb, t1 = min(Ttok, t0 + (long long)tpb);
  extern __shared__ float cs_acc[];
  for (int c = threadIdx.x; c < N; c += blockDim.x) cs_acc[c] = 0.f;
  __syncthreads();
  const long long n0 = t0 * N, n1 = t1 * N;
  // flat walk: thread i handles elements i, i + 256, ...; when N divides 256 its column never changes -> register sum
  if (256 % N == 0) {
    float v = 0.f;
    for (long long i = n0 + threadIdx.x; i < n1; i += 256) v += ldf(X + i);
    atomicAdd(cs_acc + (int)((n0 + threadIdx.x) % N), v);
  } else {
    for (long long i = n0 + threadIdx.x; i < n1; i += 256) atomicAdd(cs_acc + (int)(i % N), ldf(X + i));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += blockDim.x)
    if (cs_acc[c] != 0.f) atomicAdd(out + c, cs_acc[c]);
}

template <typename T>
__global__ void k_poison_if(const int* __restrict__ status, T* __restrict__ out, int n) {
  if (*status != 0 && threadIdx.x < n) stf(out + threadIdx.x, __int_as_float(0x7fc00000));
}

static inline void launch_wprep(const ConvDims& d, const ConvWs& ws, const float* w, const float* gamma, const float* bias, cudaStream_t st) {
  ADN_KERNEL("k_conv_wprep", st);
  k_conv_wprep<<<ew_grid((long long)d.co8 * 9 * d.cpi + (long long)d.ci8 * 9 * d.cpo + d.co8), 256, 0, st>>>(
      w, gamma, bias, ws.Wf, ws.Wd, ws.biasp, d.Cin, d.Cout, d.ci8, d.co8, d.cpi, d.cpo);
}
template <typename T>
static inline void launch_copy_rows(const T* src, int Cs, T* dst, int Cd, long long Ttok, cudaStream_t st) {
  ADN_KERNEL("k_copy_rows", st);
  k_copy_rows<T><<<ew_grid(Ttok * Cd), 256, 0, st>>>(src, Cs, dst, Cd, Ttok * Cd);
}

template <typename T>
static int conv_forward(const ConvDims& d, const T* x, const float* w, const float* bias, const float* gamma, T* y, void* wsp, cudaStream_t st) {
  ConvWs ws = conv_ws(d, wsp);
  if (d.tc) {
    using namespace tcg;
    ADN_CHECK_CUDA(cudaMemsetAsync(ws.status, 0, 256, st));
    launch_wprep(d, ws, w, gamma, bias, st);
    const bf16* xe = (const bf16*)x;
    bf16* ye = (bf16*)y;
    if (ws.xp) { launch_copy_rows<bf16>((const bf16*)x, d.Cin, ws.xp, d.ci8, d.T, st); xe = ws.xp; }
    if (ws.yp) ye = ws.yp;
    int rc = gemm(st, "conv3x3_fwd", (int)d.T, d.co8, 9 * d.cpi, kmaj(xe, d.ci8), kmaj(ws.Wf, 9 * d.cpi), 0, NOOP, NOOP,
                  Out{ye, d.co8, 0, C_BF16}, 1, 1, nullptr, 0, ws.status, NOAUX, bias ? ws.biasp : nullptr, Conv{1, d.W, d.H, d.cpi, 0},
                  Image{xe, d.B, d.H, d.W, d.ci8});
    if (rc) return rc;
    if (ws.yp) launch_copy_rows<bf16>(ws.yp, d.co8, (bf16*)y, d.Cout, d.T, st);
    { ADN_KERNEL("k_poison_if", st); k_poison_if<T><<<1, 32, 0, st>>>(ws.status, y, 4); }
  } else {
    const long long n = d.T * d.Cout;
    ADN_KERNEL("k_conv3_fwd_direct", st);
    k_conv3_fwd_direct<T><<<ew_grid(n), 256, 0, st>>>(x, w, gamma, bias, y, d.H, d.W, d.Cin, d.Cout, n);
  }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

template <typename T>
static int conv_backward(const ConvDims& d, const T* x, const float* w, const float* gamma, const T* dy, T* dx, float* dw, float* dbias,
                         float* dgamma, void* wsp, cudaStream_t st) {
  ConvWs ws = conv_ws(d, wsp);
  ADN_CHECK_CUDA(cudaMemsetAsync(ws.status, 0, 256, st));
  ADN_CHECK_CUDA(cudaMemsetAsync(ws.dWt, 0, (size_t)d.co8 * 9 * d.cpi * sizeof(float), st));
  if (dgamma) ADN_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)d.Cin * sizeof(float), st));
  if (dbias) ADN_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)d.Cout * sizeof(float), st));
  if (d.tc) {
    using namespace tcg;
    int rc = ADN_OK;
    const bf16 *xe = (const bf16*)x, *dye = (const bf16*)dy;
    if (ws.xp) { launch_copy_rows<bf16>((const bf16*)x, d.Cin, ws.xp, d.ci8, d.T, st); xe = ws.xp; }
    if (ws.yp) { launch_copy_rows<bf16>((const bf16*)dy, d.Cout, ws.yp, d.co8, d.T, st); dye = ws.yp; }
    if (dx) {
      launch_wprep(d, ws, w, gamma, nullptr, st);
      bf16* dxe = ws.dxp ? ws.dxp : (bf16*)dx;
      rc = gemm(st, "conv3x3_dgrad", (int)d.T, d.ci8, 9 * d.cpo, kmaj(dye, d.co8), kmaj(ws.Wd, 9 * d.cpo), 0, NOOP, NOOP,
                Out{dxe, d.ci8, 0, C_BF16}, 1, 1, nullptr, 0, ws.status, NOAUX, nullptr, Conv{1, d.W, d.H, d.cpo, 0},
                Image{dye, d.B, d.H, d.W, d.co8});
      if (rc) return rc;
      if (ws.dxp) launch_copy_rows<bf16>(ws.dxp, d.ci8, (bf16*)dx, d.Cin, d.T, st);
    }
    const int N = 9 * d.cpi, bn = N >= 256 ? 256 : 0;
    const int splitk = pick_splitk(cdiv(d.co8, BM) * cdiv(N, bn ? bn : pick_bn(N, 1)), (int)d.T);
    rc = gemm(st, "conv3x3_wgrad", d.co8, N, (int)d.T, mnmaj(dye, d.co8), mnmaj(xe, d.ci8), 0, NOOP, NOOP,
              Out{ws.dWt, N, 0, C_ATOMIC_F32}, 1, splitk, nullptr, 0, ws.status, NOAUX, nullptr, Conv{2, d.W, d.H, d.cpi, bn},
              Image{xe, d.B, d.H, d.W, d.ci8});
    if (rc) return rc;
  } else {
    if (dx) {
      const long long n = d.T * d.Cin;
      ADN_KERNEL("k_conv3_dgrad_direct", st);
      k_conv3_dgrad_direct<T><<<ew_grid(n), 256, 0, st>>>(dy, w, gamma, dx, d.H, d.W, d.Cin, d.Cout, n);
    }
    int blocks = 4 * sm_count();
    long long tpb = (d.T + blocks - 1) / blocks;
    tpb = tpb < 32 ? 32 : tpb;
    ADN_KERNEL("k_conv3_wgrad_direct", st);
    k_conv3_wgrad_direct<T><<<cdiv(d.T, tpb), 256, 0, st>>>(x, dy, ws.dWt, d.H, d.W, d.Cin, d.Cout, d.cpi, d.T, (int)tpb);
  }
  { ADN_KERNEL("k_conv_wfinal", st); k_conv_wfinal<<<cdiv((long long)d.Cout * d.Cin, 256), 256, 0, st>>>(ws.dWt, w, gamma, dw, dgamma, d.Cin, d.Cout, d.cpi); }
  if (dbias) {
    int blocks = 4 * sm_count();
    long long tpb = (d.T + blocks - 1) / blocks;
    tpb = tpb < 64 ? 64 : tpb;
    ADN_KERNEL("k_colsum_any", st);
    k_colsum_any<T><<<cdiv(d.T, tpb), 256, d.Cout * sizeof(float), st>>>(dy, d.Cout, d.T, (int)tpb, dbias);
  }
  if (d.tc && dx) { ADN_KERNEL("k_poison_if", st); k_poison_if<T><<<1, 32, 0, st>>>(ws.status, dx, 4); }
  if (d.tc) { ADN_KERNEL("k_poison_if", st); k_poison_if<float><<<1, 32, 0, st>>>(ws.status, dw, 4); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

// ================================================================ grouped convolution, 4 channels per group
// The `groups = dim / 4` convolutions of the EncoderToDecoder bridges (models/model_untils.py:621-675: 1x3, 3x1 and 3x3 kernels,
// stride 1, 'same' zero padding, bias) on channels-last activations (B, H*W, C).  cuDNN runs them as one launch PER GROUP plus
// two layout conversions per launch (256 groups at dim 1024: ~1 300 launches and 2 ms per bridge and step); here a conv is a
// 4 x 4 x taps stencil per output element: one launch forward, two backward.  w: (C, 4, kh, kw) fp32, state_dict layout.
// Thread mapping of all three kernels: lane = channel (32 consecutive channels = 8 groups per CTA: a warp reads / writes 64
// contiguous bytes of a bf16 token row), the 8 warps of a CTA stride over the CTA's `tps` tokens; a thread keeps the 4 x taps
// weights that involve its channel in registers.  grid = (ceil(C / 32), ceil(tokens / tps)).
constexpr int GC_MAXK = 9;
template <typename T>
__global__ void __launch_bounds__(256)
k_gconv4_fwd(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ y, int H, int W, int C,
             int kh, int kw, long long Ttok, int tps) {
  const int K = kh * kw, ph = kh / 2, pw = kw / 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = blockIdx.x * 32 + lane, g4 = co & ~3;
  if (co >= C) return;
  float wr[4][GC_MAXK];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) wr[j][t] = t < K ? w[((long long)co * 4 + j) * K + t] : 0.f;
  const float b0 = bias ? bias[co] : 0.f;
  int tdy[GC_MAXK], tdx[GC_MAXK];      // tap offsets: constant over the token loop (no division per tap and token)
#pragma unroll
  for (int t = 0; t < GC_MAXK; ++t) { tdy[t] = t / kw - ph; tdx[t] = t % kw - pw; }
  const int t0 = blockIdx.y * tps, t1 = (int)min(Ttok, (long long)t0 + tps);
  for (int p = t0 + warp; p < t1; p += 8) {
    const int row = p / W, xx = p - row * W, yy = row % H;
    float acc = b0;
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) {
      if (t < K) {
        const int dy = tdy[t], dx = tdx[t];
        if (yy + dy >= 0 && yy + dy < H && xx + dx >= 0 && xx + dx < W) {
          float v[4];
          ld4(x + (long long)(p + dy * W + dx) * C + g4, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc = fmaf(v[j], wr[j][t], acc);
        }
      }
    }
    stf(y + (long long)p * C + co, acc);
  }
}
// dx[p][4g + j] = sum_t sum_{co in group} dy[p - delta_t][4g + co] w[4g + co][j][t]
template <typename T>
__global__ void __launch_bounds__(256)
k_gconv4_dgrad(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, int H, int W, int C, int kh, int kw, long long Ttok,
               int tps) {
  const int K = kh * kw, ph = kh / 2, pw = kw / 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * 32 + lane, g4 = ci & ~3, j = ci & 3;
  if (ci >= C) return;
  float wr[4][GC_MAXK];
#pragma unroll
  for (int co = 0; co < 4; ++co)
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) wr[co][t] = t < K ? w[((long long)(g4 + co) * 4 + j) * K + t] : 0.f;
  int tdy[GC_MAXK], tdx[GC_MAXK];
#pragma unroll
  for (int t = 0; t < GC_MAXK; ++t) { tdy[t] = t / kw - ph; tdx[t] = t % kw - pw; }
  const int t0 = blockIdx.y * tps, t1 = (int)min(Ttok, (long long)t0 + tps);
  for (int p = t0 + warp; p < t1; p += 8) {
    const int row = p / W, xx = p - row * W, yy = row % H;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) {
      if (t < K) {
        const int dyy = tdy[t], dxx = tdx[t];
        if (yy - dyy >= 0 && yy - dyy < H && xx - dxx >= 0 && xx - dxx < W) {
          float v[4];
          ld4(dy + (long long)(p - dyy * W - dxx) * C + g4, v);
#pragma unroll
          for (int co = 0; co < 4; ++co) acc = fmaf(v[co], wr[co][t], acc);
        }
      }
    }
    stf(dx + (long long)p * C + ci, acc);
  }
}
// dw[co][j][t] += sum_p dy[p][co] x[p + delta_t][4g + j],  dbias[co] += sum_p dy[p][co]: 4 x taps + 1 partial sums per thread over
// its tokens, summed over the 8 warps through shared memory, one atomic per entry and CTA (dw / dbias zeroed by the caller)
template <typename T>
__global__ void __launch_bounds__(256)
k_gconv4_wgrad(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ dbias, int H, int W, int C,
               int kh, int kw, long long Ttok, int tps) {
  __shared__ float red[8][4 * GC_MAXK + 1][32];
  const int K = kh * kw, ph = kh / 2, pw = kw / 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = blockIdx.x * 32 + lane, g4 = co & ~3;
  float acc[4][GC_MAXK], ab = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) acc[j][t] = 0.f;
  int tdy[GC_MAXK], tdx[GC_MAXK];
#pragma unroll
  for (int t = 0; t < GC_MAXK; ++t) { tdy[t] = t / kw - ph; tdx[t] = t % kw - pw; }
  const int t0 = blockIdx.y * tps, t1 = (int)min(Ttok, (long long)t0 + tps);
  if (co < C) {
    for (int p = t0 + warp; p < t1; p += 8) {
      const int row = p / W, xx = p - row * W, yy = row % H;
      const float g = ldf(dy + (long long)p * C + co);
      ab += g;
#pragma unroll
      for (int t = 0; t < GC_MAXK; ++t) {
        if (t < K) {
          const int dyy = tdy[t], dxx = tdx[t];
          if (yy + dyy >= 0 && yy + dyy < H && xx + dxx >= 0 && xx + dxx < W) {
            float v[4];
            ld4(x + (long long)(p + dyy * W + dxx) * C + g4, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][t] = fmaf(g, v[j], acc[j][t]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < GC_MAXK; ++t) red[warp][j * GC_MAXK + t][lane] = acc[j][t];
  red[warp][4 * GC_MAXK][lane] = ab;
  __syncthreads();
  for (int e = threadIdx.x; e < (4 * GC_MAXK + 1) * 32; e += 256) {
    const int q = e >> 5, l = e & 31, c = blockIdx.x * 32 + l;
    float v = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) v += red[wq][q][l];
    if (c >= C || v == 0.f) continue;
    if (q == 4 * GC_MAXK) { if (dbias) atomicAdd(dbias + c, v); }
    else {
      const int j = q / GC_MAXK, t = q - j * GC_MAXK;
      if (t < K) atomicAdd(dw + ((long long)c * 4 + j) * K + t, v);
    }
  }
}

// ================================================================ layout changes, InstanceNorm, shortcut mix
// A (32 V) x (32 V) (pixels x channels) tile goes through shared memory: the token-major side is read / written with lanes along
// the channels, the NCHW side with lanes along the pixels, V elements per lane: V = 2 (4-byte bf16x2 / 8-byte float2 accesses,
// 128 / 256 contiguous bytes per warp row) when the channel counts and the plane size are even, V = 1 otherwise (C = 5).
template <int V> __device__ __forceinline__ void ldv(const float* p, float* v) {
  if constexpr (V == 1) { v[0] = *p; } else { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
}
template <int V> __device__ __forceinline__ void ldv(const bf16* p, float* v) {
  if constexpr (V == 1) { v[0] = __bfloat162float(*p); } else { sm100::unpack_bf16(*reinterpret_cast<const uint32_t*>(p), v[0], v[1]); }
}
template <int V> __device__ __forceinline__ void stv(float* p, const float* v) {
  if constexpr (V == 1) { *p = v[0]; } else { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
}
template <int V> __device__ __forceinline__ void stv(bf16* p, const float* v) {
  if constexpr (V == 1) { *p = __float2bfloat16_rn(v[0]); } else { *reinterpret_cast<uint32_t*>(p) = sm100::pack_bf16(v[0], v[1]); }
}
// Thread (tx = lane, ty = warp of 8): token side -> pixel row ty + 8 j, channels V tx ..; plane side -> channel row ty + 8 j, pixels V tx ..
#define TILE_DECL(V)                                                                         \
  constexpr int TSV = 32 * V;                                                                \
  __shared__ float tile[TSV][TSV + 1];                                                       \
  const int b = blockIdx.z, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                    \
  const int c0 = blockIdx.y * TSV;

// out[b][c][p] = g1 x[b][p][c] (c < C1) | g2 res[b][p][c - C1]        (WTLayer :404-414: cat(gama1 x, gama2 residual), then NCHW)
template <typename T, int V>
__global__ void __launch_bounds__(256)
k_pack_fwd(const T* __restrict__ x, int C1, const T* __restrict__ res, int C2, const float* __restrict__ g1p, const float* __restrict__ g2p,
           T* __restrict__ out, long long HW) {
  TILE_DECL(V)
  const int C = C1 + C2;
  const long long p0 = (long long)blockIdx.x * TSV;
  const float g1 = g1p ? *g1p : 1.f, g2 = g2p ? *g2p : 1.f;
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int pl = ty + 8 * j, c = c0 + V * tx;
    const long long p = p0 + pl;
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = 0.f;
    if (p < HW && c < C) {
      if (c < C1) ldv<V>(x + ((long long)b * HW + p) * C1 + c, v);
      else ldv<V>(res + ((long long)b * HW + p) * C2 + (c - C1), v);
      const float g = c < C1 ? g1 : g2;
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] *= g;
    }
#pragma unroll
    for (int e = 0; e < V; ++e) tile[pl][V * tx + e] = v[e];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int cl = ty + 8 * j, c = c0 + cl;
    const long long p = p0 + V * tx;
    if (p < HW && c < C) {
      float v[V];
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = tile[V * tx + e][cl];
      stv<V>(out + ((long long)b * C + c) * HW + p, v);
    }
  }
}

// dx = g1 dout^T, dres = g2 dout^T, acc[0] += <x, dout^T>, acc[1] += <res, dout^T>   (fp64 accumulators, zeroed by the caller)
template <typename T, int V>
__global__ void __launch_bounds__(256)
k_pack_bwd(const T* __restrict__ x, int C1, const T* __restrict__ res, int C2, const float* __restrict__ g1p, const float* __restrict__ g2p,
           const T* __restrict__ dout, T* __restrict__ dx, T* __restrict__ dres, double* __restrict__ acc, long long HW) {
  TILE_DECL(V)
  __shared__ float red[2][8];
  const int C = C1 + C2;
  const long long p0 = (long long)blockIdx.x * TSV;
  const float g1 = g1p ? *g1p : 1.f, g2 = g2p ? *g2p : 1.f;
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int cl = ty + 8 * j, c = c0 + cl;
    const long long p = p0 + V * tx;
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = 0.f;
    if (p < HW && c < C) ldv<V>(dout + ((long long)b * C + c) * HW + p, v);
#pragma unroll
    for (int e = 0; e < V; ++e) tile[V * tx + e][cl] = v[e];
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int pl = ty + 8 * j, c = c0 + V * tx;
    const long long p = p0 + pl;
    if (p < HW && c < C) {
      float g[V], o[V], xv[V];
#pragma unroll
      for (int e = 0; e < V; ++e) g[e] = tile[pl][V * tx + e];
      if (c < C1) {
        const long long off = ((long long)b * HW + p) * C1 + c;
        if (acc && g1p) {
          ldv<V>(x + off, xv);
#pragma unroll
          for (int e = 0; e < V; ++e) s1 = fmaf(g[e], xv[e], s1);
        }
        if (dx) {
#pragma unroll
          for (int e = 0; e < V; ++e) o[e] = g1 * g[e];
          stv<V>(dx + off, o);
        }
      } else {
        const long long off = ((long long)b * HW + p) * C2 + (c - C1);
        if (acc && g2p) {
          ldv<V>(res + off, xv);
#pragma unroll
          for (int e = 0; e < V; ++e) s2 = fmaf(g[e], xv[e], s2);
        }
        if (dres) {
#pragma unroll
          for (int e = 0; e < V; ++e) o[e] = g2 * g[e];
          stv<V>(dres + off, o);
        }
      }
    }
  }
  if (acc) {
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (tx == 0) { red[0][ty] = s1; red[1][ty] = s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float v = 0.f;
      for (int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
      if (v != 0.f) atomicAdd(acc + threadIdx.x, (double)v);
    }
  }
}
__global__ void k_store_acc(const double* __restrict__ acc, float* __restrict__ d1, float* __restrict__ d2) {
  if (threadIdx.x == 0) { if (d1) *d1 = (float)acc[0]; if (d2) *d2 = (float)acc[1]; }
}

__device__ __forceinline__ void ldvec(const float* p, float (&v)[4]) { ld4(p, v); }
__device__ __forceinline__ void ldvec(const bf16* p, float (&v)[8]) { sm100::unpack8(*reinterpret_cast<const uint4*>(p), v); }

// InstanceNorm2d statistics (nn.InstanceNorm2d defaults: eps 1e-5, biased variance, no affine, no running stats):
// stats[plane] = (mean, rstd).  One CTA per plane; sums are taken about the plane's first element (no cancellation).
template <typename T>
__global__ void __launch_bounds__(256)
k_plane_stats(const T* __restrict__ y, float* __restrict__ stats, long long HW, float eps) {
  __shared__ float red[2][8];
  const T* p = y + (long long)blockIdx.x * HW;
  const float k = ldf(p);
  float s1 = 0.f, s2 = 0.f;
  constexpr int VV = 16 / (int)sizeof(T);            // elements per 16-byte load
  if (HW % VV == 0 && ((uintptr_t)y & 15) == 0) {
    for (long long i = (long long)threadIdx.x * VV; i < HW; i += 256 * VV) {
      float v[VV];
      ldvec(p + i, v);
#pragma unroll
      for (int j = 0; j < VV; ++j) { const float d = v[j] - k; s1 += d; s2 = fmaf(d, d, s2); }
    }
  } else {
    for (long long i = threadIdx.x; i < HW; i += 256) {
      const float v = ldf(p + i) - k;
      s1 += v; s2 = fmaf(v, v, s2);
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b2 += red[1][w]; }
    const float m = a / (float)HW;
    const float var = fmaxf(b2 / (float)HW - m * m, 0.f);
    stats[2 * blockIdx.x] = k + m;
    stats[2 * blockIdx.x + 1] = rsqrtf(var + eps);
  }
}

// out[b][p][c] = gamma[c] (alpha act(u) + beta xs[b][c][p]),  u = scale (y - mean) rstd + shift  (stats NULL: u = y)
//   WTLayer    :416  alpha wtconv(x) + beta shortcut            (WTConvLayer :112-113 norm, no act)
//   PatchEmbed :303  alpha1 GELU(wtconv(x)) + beta1 x ;  :305-307 (alpha2 IN(wtconv(s)) + beta2 s) gamma
//   OutProj    :880-883 (alpha GELU(IN(wtconv(x))) + beta shortcut) gamma
struct MixP {
  const float *stats, *scale, *shift, *alpha, *beta, *gamma;
  int act;
};
template <typename T, int V>
__global__ void __launch_bounds__(256)
k_mix_fwd(const T* __restrict__ y, const T* __restrict__ xs, MixP m, T* __restrict__ out, int C, long long HW) {
  TILE_DECL(V)
  const long long p0 = (long long)blockIdx.x * TSV;
  const float sc = m.scale ? *m.scale : 1.f, sh = m.shift ? *m.shift : 0.f, al = *m.alpha, be = *m.beta;
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int cl = ty + 8 * j, c = c0 + cl;
    const long long p = p0 + V * tx;
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = 0.f;
    if (p < HW && c < C) {
      const long long plane = (long long)b * C + c, o = plane * HW + p;
      float u[V], s[V];
      ldv<V>(y + o, u);
      ldv<V>(xs + o, s);
      const float ga = m.gamma ? m.gamma[c] : 1.f;
      float mean = 0.f, rstd = 1.f;
      if (m.stats) { mean = m.stats[2 * plane]; rstd = m.stats[2 * plane + 1]; }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float t = u[e];
        if (m.stats) t = sc * (t - mean) * rstd + sh;
        if (m.act) t = Act<T>::gelu(t);
        v[e] = (al * t + be * s[e]) * ga;
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) tile[V * tx + e][cl] = v[e];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int pl = ty + 8 * j, c = c0 + V * tx;
    const long long p = p0 + pl;
    if (p < HW && c < C) {
      float v[V];
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = tile[pl][V * tx + e];
      stv<V>(out + ((long long)b * HW + p) * C + c, v);
    }
  }
}

// Backward, pass 1: per-plane sums  S1 = sum e, S2 = sum e n, S3 = sum dout act(u), S4 = sum dout xs   with
// n = (y - mean) rstd (or y), u = scale n + shift, e = dout act'(u).  A CTA owns 32 V channels x `nsub` pixel tiles.
template <typename T, int V>
__global__ void __launch_bounds__(256)
k_mix_bwd_sums(const T* __restrict__ y, const T* __restrict__ xs, MixP m, const T* __restrict__ dout, float* __restrict__ sums, int C, long long HW,
               int nsub) {
  TILE_DECL(V)
  const float sc = m.scale ? *m.scale : 1.f, sh = m.shift ? *m.shift : 0.f;
  float acc[4 * V][4];
#pragma unroll
  for (int j = 0; j < 4 * V; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
  for (int sub = 0; sub < nsub; ++sub) {
    const long long p0 = ((long long)blockIdx.x * nsub + sub) * TSV;
    if (p0 >= HW) break;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4 * V; ++j) {
      const int pl = ty + 8 * j, c = c0 + V * tx;
      const long long p = p0 + pl;
      float v[V];
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = 0.f;
      if (p < HW && c < C) ldv<V>(dout + ((long long)b * HW + p) * C + c, v);
#pragma unroll
      for (int e = 0; e < V; ++e) tile[pl][V * tx + e] = v[e];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4 * V; ++j) {
      const int cl = ty + 8 * j, c = c0 + cl;
      const long long p = p0 + V * tx;
      if (p < HW && c < C) {
        const long long plane = (long long)b * C + c, o = plane * HW + p;
        float yv[V], sv[V];
        ldv<V>(y + o, yv);
        ldv<V>(xs + o, sv);
        float mean = 0.f, rstd = 1.f;
        if (m.stats) { mean = m.stats[2 * plane]; rstd = m.stats[2 * plane + 1]; }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float g = tile[V * tx + e][cl];
          const float n = m.stats ? (yv[e] - mean) * rstd : yv[e];
          const float u = m.stats ? sc * n + sh : n;
          const float ee = m.act ? g * Act<T>::gelu_grad(u) : g;
          acc[j][0] += ee;
          acc[j][1] = fmaf(ee, n, acc[j][1]);
          acc[j][2] = fmaf(g, m.act ? Act<T>::gelu(u) : u, acc[j][2]);
          acc[j][3] = fmaf(g, sv[e], acc[j][3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int c = c0 + ty + 8 * j;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = warp_sum(acc[j][q]);
      if (tx == 0 && c < C && v != 0.f) atomicAdd(sums + ((long long)b * C + c) * 4 + q, v);
    }
  }
}

// Backward, scalars: dscal = (dscale, dshift, dalpha, dbeta), dgamma[C]; one block, fp64 reductions over the planes
__global__ void __launch_bounds__(256)
k_mix_bwd_final(const float* __restrict__ sums, MixP m, float* __restrict__ dscal, float* __restrict__ dgamma, int B, int C) {
  __shared__ double red[4][256];
  const float al = *m.alpha, be = *m.beta;
  double a[4] = {0., 0., 0., 0.};
  for (int pl = threadIdx.x; pl < B * C; pl += 256) {
    const double g = m.gamma ? (double)m.gamma[pl % C] : 1.0;
    const float* s = sums + (long long)pl * 4;
    a[0] += g * s[1];      // d scale = alpha sum gamma S2
    a[1] += g * s[0];      // d shift = alpha sum gamma S1
    a[2] += g * s[2];      // d alpha = sum gamma S3
    a[3] += g * s[3];      // d beta  = sum gamma S4
  }
  for (int q = 0; q < 4; ++q) red[q][threadIdx.x] = a[q];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 4; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dscal[0] = (float)(al * red[0][0]);
    dscal[1] = (float)(al * red[1][0]);
    dscal[2] = (float)red[2][0];
    dscal[3] = (float)red[3][0];
  }
  if (dgamma)
    for (int c = threadIdx.x; c < C; c += 256) {
      double v = 0.;
      for (int b = 0; b < B; ++b) v += (double)al * sums[((long long)b * C + c) * 4 + 2] + (double)be * sums[((long long)b * C + c) * 4 + 3];
      dgamma[c] = (float)v;
    }
}

// Backward, pass 2: dy = rstd scale alpha gamma (e - S1 / HW - n S2 / HW)   (no norm: alpha gamma e),  dxs = beta gamma dout
template <typename T, int V>
__global__ void __launch_bounds__(256)
k_mix_bwd_apply(const T* __restrict__ y, MixP m, const T* __restrict__ dout, const float* __restrict__ sums, T* __restrict__ dy, T* __restrict__ dxs,
                int C, long long HW) {
  TILE_DECL(V)
  const long long p0 = (long long)blockIdx.x * TSV;
  const float sc = m.scale ? *m.scale : 1.f, sh = m.shift ? *m.shift : 0.f, al = *m.alpha, be = *m.beta;
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int pl = ty + 8 * j, c = c0 + V * tx;
    const long long p = p0 + pl;
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = 0.f;
    if (p < HW && c < C) ldv<V>(dout + ((long long)b * HW + p) * C + c, v);
#pragma unroll
    for (int e = 0; e < V; ++e) tile[pl][V * tx + e] = v[e];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4 * V; ++j) {
    const int cl = ty + 8 * j, c = c0 + cl;
    const long long p = p0 + V * tx;
    if (p < HW && c < C) {
      const long long plane = (long long)b * C + c, o = plane * HW + p;
      const float ga = m.gamma ? m.gamma[c] : 1.f;
      float g[V], r[V];
#pragma unroll
      for (int e = 0; e < V; ++e) g[e] = tile[V * tx + e][cl];
      if (dxs) {
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = be * ga * g[e];
        stv<V>(dxs + o, r);
      }
      if (dy) {
        float yv[V];
        ldv<V>(y + o, yv);
        if (m.stats) {
          const float mean = m.stats[2 * plane], rstd = m.stats[2 * plane + 1];
          const float k = rstd * sc * al * ga, a1 = inv * sums[plane * 4], a2 = inv * sums[plane * 4 + 1];
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float n = (yv[e] - mean) * rstd;
            const float ee = m.act ? g[e] * Act<T>::gelu_grad(sc * n + sh) : g[e];
            r[e] = k * (ee - a1 - n * a2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) r[e] = al * ga * (m.act ? g[e] * Act<T>::gelu_grad(yv[e]) : g[e]);
        }
        stv<V>(dy + o, r);
      }
    }
  }
}

// ---- activations after the convs: kind 1 GELU (erf form, nn.GELU default), kind 2 Swish x sigmoid(beta x) (model_untils.py:162-169)
// 16-byte accesses (8 bf16 / 4 fp32 per thread and step); the tail of a length that is not a multiple runs element-wise.
__device__ __forceinline__ void stvec(float* p, const float (&v)[4]) { st4(p, v); }
__device__ __forceinline__ void stvec(bf16* p, const float (&v)[8]) { *reinterpret_cast<uint4*>(p) = sm100::pack8(v); }
template <typename T> __device__ __forceinline__ float act_f(float v, int kind, float be) { return kind == 1 ? Act<T>::gelu(v) : v / (1.f + expf(-be * v)); }

template <typename T>
__global__ void __launch_bounds__(256)
k_act_fwd(const T* __restrict__ x, T* __restrict__ y, long long n, int kind, const float* __restrict__ betap) {
  constexpr int VV = 16 / (int)sizeof(T);
  const float be = betap ? *betap : 1.f;
  const long long stride = (long long)gridDim.x * blockDim.x, nv = n / VV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float v[VV];
    ldvec(x + i * VV, v);
#pragma unroll
    for (int j = 0; j < VV; ++j) v[j] = act_f<T>(v[j], kind, be);
    stvec(y + i * VV, v);
  }
  for (long long i = nv * VV + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) stf(y + i, act_f<T>(ldf(x + i), kind, be));
}
template <typename T> __device__ __forceinline__ float act_bwd_f(float v, float g, int kind, float be, float& db) {
  if (kind == 1) return g * Act<T>::gelu_grad(v);
  const float s = 1.f / (1.f + expf(-be * v)), ds = s * (1.f - s);
  db = fmaf(g, v * v * ds, db);
  return g * (s + v * be * ds);
}
template <typename T>
__global__ void __launch_bounds__(256)
k_act_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n, int kind, const float* __restrict__ betap,
          double* __restrict__ dbeta_acc) {
  __shared__ float red[8];
  constexpr int VV = 16 / (int)sizeof(T);
  const float be = betap ? *betap : 1.f;
  const long long stride = (long long)gridDim.x * blockDim.x, nv = n / VV;
  float db = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float v[VV], g[VV];
    ldvec(x + i * VV, v);
    ldvec(dy + i * VV, g);
#pragma unroll
    for (int j = 0; j < VV; ++j) v[j] = act_bwd_f<T>(v[j], g[j], kind, be, db);
    stvec(dx + i * VV, v);
  }
  for (long long i = nv * VV + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    stf(dx + i, act_bwd_f<T>(ldf(x + i), ldf(dy + i), kind, be, db));
  if (dbeta_acc) {
    db = warp_sum(db);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = db;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
      for (int k = 0; k < 8; ++k) v += red[k];
      if (v != 0.f) atomicAdd(dbeta_acc, (double)v);
    }
  }
}

}  // namespace cst
}  // namespace adn

using namespace adn;
using namespace adn::cst;

#define CST_DISPATCH(dtype, CALL_F32, CALL_BF16) ((dtype) == ADN_F32 ? (CALL_F32) : (CALL_BF16))

extern "C" {

int adn_conv3x3_path(const AdnConvShape* s) {
  ConvDims d;
  if (conv_dims(s, &d, "adn_conv3x3_path")) return -1;
  return d.tc ? 1 : 0;
}

int adn_conv3x3_workspace_bytes(const AdnConvShape* s, size_t* bytes) {
  ConvDims d;
  int rc = conv_dims(s, &d, "adn_conv3x3_workspace_bytes");
  if (rc) return rc;
  if (bytes) *bytes = conv_ws(d, nullptr).bytes;
  return ADN_OK;
}

int adn_conv3x3_forward(const AdnConvShape* s, const void* x, const float* w, const float* bias, const float* gamma, void* y, void* ws,
                        void* stream) {
  ConvDims d;
  int rc = conv_dims(s, &d, "adn_conv3x3_forward");
  if (rc) return rc;
  ADN_REQUIRE(x && w && y && ws, ADN_ERR_NULL, "adn_conv3x3_forward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  return s->dtype == ADN_F32 ? conv_forward<float>(d, (const float*)x, w, bias, gamma, (float*)y, ws, st)
                             : conv_forward<bf16>(d, (const bf16*)x, w, bias, gamma, (bf16*)y, ws, st);
}

int adn_conv3x3_backward(const AdnConvShape* s, const void* x, const float* w, const float* gamma, const void* dy, void* dx, float* dw,
                         float* dbias, float* dgamma, void* ws, void* stream) {
  ConvDims d;
  int rc = conv_dims(s, &d, "adn_conv3x3_backward");
  if (rc) return rc;
  ADN_REQUIRE(x && w && dy && dw && ws, ADN_ERR_NULL, "adn_conv3x3_backward: NULL argument");
  ADN_REQUIRE(dgamma == nullptr || gamma != nullptr, ADN_ERR_NULL, "adn_conv3x3_backward: dgamma without gamma");
  cudaStream_t st = (cudaStream_t)stream;
  return s->dtype == ADN_F32 ? conv_backward<float>(d, (const float*)x, w, gamma, (const float*)dy, (float*)dx, dw, dbias, dgamma, ws, st)
                             : conv_backward<bf16>(d, (const bf16*)x, w, gamma, (const bf16*)dy, (bf16*)dx, dw, dbias, dgamma, ws, st);
}

// elements per lane of the tile kernels: 2 when every row of both layouts starts on a 2-element boundary
static inline int tile_v(int C1, int C2, long long HW) { return (C1 % 2 == 0 && C2 % 2 == 0 && HW % 2 == 0) ? 2 : 1; }

static int plane_check(int32_t B, int32_t C, int64_t HW, int32_t dtype, const char* what) {
  ADN_REQUIRE(B > 0 && C > 0 && HW > 0 && B <= 65535 && (long long)B * C * HW < (1LL << 40), ADN_ERR_SHAPE, "%s: bad extents (%d, %d, %lld)", what, B, C, (long long)HW);
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "%s: unsupported dtype %d", what, dtype);
  return ADN_OK;
}

int adn_nchw_pack_forward(const void* x, const void* res, const float* g1, const float* g2, void* out, int32_t B, int64_t HW, int32_t C1,
                          int32_t C2, int32_t dtype, void* stream) {
  int rc = plane_check(B, C1 + C2, HW, dtype, "adn_nchw_pack_forward");
  if (rc) return rc;
  ADN_REQUIRE(x && out && C1 > 0 && C2 >= 0 && (C2 == 0 || res), ADN_ERR_NULL, "adn_nchw_pack_forward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int V = tile_v(C1, C2, HW);
  dim3 grid(cdiv(HW, 32 * V), cdiv(C1 + C2, 32 * V), B);
  ADN_KERNEL("k_pack_fwd", st);
#define PACK_FWD(T, VV) k_pack_fwd<T, VV><<<grid, 256, 0, st>>>((const T*)x, C1, (const T*)res, C2, g1, g2, (T*)out, HW)
  if (dtype == ADN_F32) { if (V == 2) PACK_FWD(float, 2); else PACK_FWD(float, 1); }
  else { if (V == 2) PACK_FWD(bf16, 2); else PACK_FWD(bf16, 1); }
#undef PACK_FWD
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

/* ws: 64 bytes.  dx / dres / dg1 / dg2 may be NULL (not needed). */
int adn_nchw_pack_backward(const void* x, const void* res, const float* g1, const float* g2, const void* dout, void* dx, void* dres, float* dg1,
                           float* dg2, void* ws, int32_t B, int64_t HW, int32_t C1, int32_t C2, int32_t dtype, void* stream) {
  int rc = plane_check(B, C1 + C2, HW, dtype, "adn_nchw_pack_backward");
  if (rc) return rc;
  ADN_REQUIRE(x && dout && ws && C1 > 0 && C2 >= 0 && (C2 == 0 || res), ADN_ERR_NULL, "adn_nchw_pack_backward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (dg1 || dg2) ? (double*)ws : nullptr;
  if (acc) ADN_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  const int V = tile_v(C1, C2, HW);
  dim3 grid(cdiv(HW, 32 * V), cdiv(C1 + C2, 32 * V), B);
  {
    ADN_KERNEL("k_pack_bwd", st);
#define PACK_BWD(T, VV) k_pack_bwd<T, VV><<<grid, 256, 0, st>>>((const T*)x, C1, (const T*)res, C2, g1, g2, (const T*)dout, (T*)dx, (T*)dres, acc, HW)
    if (dtype == ADN_F32) { if (V == 2) PACK_BWD(float, 2); else PACK_BWD(float, 1); }
    else { if (V == 2) PACK_BWD(bf16, 2); else PACK_BWD(bf16, 1); }
#undef PACK_BWD
  }
  if (acc) { ADN_KERNEL("k_store_acc", st); k_store_acc<<<1, 32, 0, st>>>(acc, dg1, dg2); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_plane_stats(const void* y, float* stats, int64_t planes, int64_t HW, float eps, int32_t dtype, void* stream) {
  ADN_REQUIRE(y && stats, ADN_ERR_NULL, "adn_plane_stats: NULL argument");
  ADN_REQUIRE(planes > 0 && planes < (1LL << 31) && HW > 0, ADN_ERR_SHAPE, "adn_plane_stats: bad extents");
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "adn_plane_stats: unsupported dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  ADN_KERNEL("k_plane_stats", st);
  if (dtype == ADN_F32) k_plane_stats<float><<<(unsigned)planes, 256, 0, st>>>((const float*)y, stats, HW, eps);
  else k_plane_stats<bf16><<<(unsigned)planes, 256, 0, st>>>((const bf16*)y, stats, HW, eps);
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_plane_mix_forward(const void* y, const void* xs, const float* stats, const float* scale, const float* shift, const float* alpha,
                          const float* beta, const float* gamma, void* out, int32_t B, int32_t C, int64_t HW, int32_t act, int32_t dtype,
                          void* stream) {
  int rc = plane_check(B, C, HW, dtype, "adn_plane_mix_forward");
  if (rc) return rc;
  ADN_REQUIRE(y && xs && alpha && beta && out, ADN_ERR_NULL, "adn_plane_mix_forward: NULL argument");
  ADN_REQUIRE(act == 0 || act == 1, ADN_ERR_SHAPE, "adn_plane_mix_forward: act must be 0 (none) or 1 (GELU)");
  cudaStream_t st = (cudaStream_t)stream;
  MixP m{stats, scale, shift, alpha, beta, gamma, act};
  const int V = tile_v(C, 0, HW);
  dim3 grid(cdiv(HW, 32 * V), cdiv(C, 32 * V), B);
  ADN_KERNEL("k_mix_fwd", st);
#define MIX_FWD(T, VV) k_mix_fwd<T, VV><<<grid, 256, 0, st>>>((const T*)y, (const T*)xs, m, (T*)out, C, HW)
  if (dtype == ADN_F32) { if (V == 2) MIX_FWD(float, 2); else MIX_FWD(float, 1); }
  else { if (V == 2) MIX_FWD(bf16, 2); else MIX_FWD(bf16, 1); }
#undef MIX_FWD
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_plane_mix_workspace_bytes(int32_t B, int32_t C, size_t* bytes) {
  ADN_REQUIRE(B > 0 && C > 0, ADN_ERR_SHAPE, "adn_plane_mix_workspace_bytes: bad extents");
  if (bytes) *bytes = align_up((size_t)B * C * 4 * sizeof(float), 256);
  return ADN_OK;
}

/* dscal[4] = (dscale, dshift, dalpha, dbeta), dgamma[C] (NULL without gamma); dy / dxs NCHW, either may be NULL. */
int adn_plane_mix_backward(const void* y, const void* xs, const float* stats, const float* scale, const float* shift, const float* alpha,
                           const float* beta, const float* gamma, const void* dout, void* dy, void* dxs, float* dscal, float* dgamma, void* ws,
                           int32_t B, int32_t C, int64_t HW, int32_t act, int32_t dtype, void* stream) {
  int rc = plane_check(B, C, HW, dtype, "adn_plane_mix_backward");
  if (rc) return rc;
  ADN_REQUIRE(y && xs && alpha && beta && dout && dscal && ws, ADN_ERR_NULL, "adn_plane_mix_backward: NULL argument");
  ADN_REQUIRE(act == 0 || act == 1, ADN_ERR_SHAPE, "adn_plane_mix_backward: act must be 0 (none) or 1 (GELU)");
  cudaStream_t st = (cudaStream_t)stream;
  MixP m{stats, scale, shift, alpha, beta, gamma, act};
  float* sums = (float*)ws;
  ADN_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * C * 4 * sizeof(float), st));
  const int nsub = 8, V = tile_v(C, 0, HW);
  dim3 g1(cdiv(HW, (long long)32 * V * nsub), cdiv(C, 32 * V), B), g2(cdiv(HW, 32 * V), cdiv(C, 32 * V), B);
#define MIX_SUMS(T, VV) k_mix_bwd_sums<T, VV><<<g1, 256, 0, st>>>((const T*)y, (const T*)xs, m, (const T*)dout, sums, C, HW, nsub)
#define MIX_APPLY(T, VV) k_mix_bwd_apply<T, VV><<<g2, 256, 0, st>>>((const T*)y, m, (const T*)dout, sums, (T*)dy, (T*)dxs, C, HW)
  {
    ADN_KERNEL("k_mix_bwd_sums", st);
    if (dtype == ADN_F32) { if (V == 2) MIX_SUMS(float, 2); else MIX_SUMS(float, 1); }
    else { if (V == 2) MIX_SUMS(bf16, 2); else MIX_SUMS(bf16, 1); }
  }
  { ADN_KERNEL("k_mix_bwd_final", st); k_mix_bwd_final<<<1, 256, 0, st>>>(sums, m, dscal, gamma ? dgamma : nullptr, B, C); }
  if (dy || dxs) {
    ADN_KERNEL("k_mix_bwd_apply", st);
    if (dtype == ADN_F32) { if (V == 2) MIX_APPLY(float, 2); else MIX_APPLY(float, 1); }
    else { if (V == 2) MIX_APPLY(bf16, 2); else MIX_APPLY(bf16, 1); }
  }
#undef MIX_SUMS
#undef MIX_APPLY
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int adn_act_forward(const void* x, void* y, int64_t n, int32_t kind, const float* beta, int32_t dtype, void* stream) {
  ADN_REQUIRE(x && y, ADN_ERR_NULL, "adn_act_forward: NULL argument");
  ADN_REQUIRE(n > 0 && (kind == 1 || kind == 2), ADN_ERR_SHAPE, "adn_act_forward: kind must be 1 (GELU) or 2 (Swish)");
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "adn_act_forward: unsupported dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  ADN_KERNEL("k_act_fwd", st);
  ADN_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, ADN_ERR_SHAPE, "adn_act_forward: buffers must be 16-byte aligned");
  if (dtype == ADN_F32) k_act_fwd<float><<<ew_grid(n / 4 + 1), 256, 0, st>>>((const float*)x, (float*)y, n, kind, beta);
  else k_act_fwd<bf16><<<ew_grid(n / 8 + 1), 256, 0, st>>>((const bf16*)x, (bf16*)y, n, kind, beta);
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

/* ws: 64 bytes; dbeta (Swish only) may be NULL. */
int adn_act_backward(const void* x, const void* dy, void* dx, int64_t n, int32_t kind, const float* beta, float* dbeta, void* ws, int32_t dtype,
                     void* stream) {
  ADN_REQUIRE(x && dy && dx && ws, ADN_ERR_NULL, "adn_act_backward: NULL argument");
  ADN_REQUIRE(n > 0 && (kind == 1 || kind == 2), ADN_ERR_SHAPE, "adn_act_backward: kind must be 1 (GELU) or 2 (Swish)");
  ADN_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0, ADN_ERR_SHAPE, "adn_act_backward: buffers must be 16-byte aligned");
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "adn_act_backward: unsupported dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (kind == 2 && dbeta) ? (double*)ws : nullptr;
  if (acc) ADN_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  {
    ADN_KERNEL("k_act_bwd", st);
    if (dtype == ADN_F32) k_act_bwd<float><<<ew_grid(n / 4 + 1), 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, n, kind, beta, acc);
    else k_act_bwd<bf16><<<ew_grid(n / 8 + 1), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n, kind, beta, acc);
  }
  if (acc) { ADN_KERNEL("k_store_acc", st); k_store_acc<<<1, 32, 0, st>>>(acc, dbeta, nullptr); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

static int gconv_check(int32_t B, int32_t H, int32_t W, int32_t C, int32_t kh, int32_t kw, int32_t dtype, const char* what) {
  ADN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && (long long)B * H * W < (1LL << 30), ADN_ERR_SHAPE,
              "%s: bad extents (B %d, %d x %d, C %d: C must be a multiple of 4)", what, B, H, W, C);
  ADN_REQUIRE((kh == 1 || kh == 3) && (kw == 1 || kw == 3), ADN_ERR_SHAPE, "%s: kernel %d x %d (1 or 3 per side)", what, kh, kw);
  ADN_REQUIRE(dtype == ADN_F32 || dtype == ADN_BF16, ADN_ERR_DTYPE, "%s: unsupported dtype %d", what, dtype);
  return ADN_OK;
}

int adn_gconv4_forward(const void* x, const float* w, const float* bias, void* y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t kh,
                       int32_t kw, int32_t dtype, void* stream) {
  int rc = gconv_check(B, H, W, C, kh, kw, dtype, "adn_gconv4_forward");
  if (rc) return rc;
  ADN_REQUIRE(x && w && y, ADN_ERR_NULL, "adn_gconv4_forward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long Ttok = (long long)B * H * W;
  const int tps = 128;
  dim3 grid(cdiv(C, 32), cdiv(Ttok, tps));
  ADN_KERNEL("k_gconv4_fwd", st);
  if (dtype == ADN_F32) k_gconv4_fwd<float><<<grid, 256, 0, st>>>((const float*)x, w, bias, (float*)y, H, W, C, kh, kw, Ttok, tps);
  else k_gconv4_fwd<bf16><<<grid, 256, 0, st>>>((const bf16*)x, w, bias, (bf16*)y, H, W, C, kh, kw, Ttok, tps);
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

/* dx may be NULL; dw (C, 4, kh, kw) and dbias (C, may be NULL) are OVERWRITTEN. */
int adn_gconv4_backward(const void* x, const float* w, const void* dy, void* dx, float* dw, float* dbias, int32_t B, int32_t H, int32_t W,
                        int32_t C, int32_t kh, int32_t kw, int32_t dtype, void* stream) {
  int rc = gconv_check(B, H, W, C, kh, kw, dtype, "adn_gconv4_backward");
  if (rc) return rc;
  ADN_REQUIRE(x && w && dy && dw, ADN_ERR_NULL, "adn_gconv4_backward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long Ttok = (long long)B * H * W;
  ADN_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * 4 * kh * kw * sizeof(float), st));
  if (dbias) ADN_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), st));
  const int tps = 128;
  dim3 grid(cdiv(C, 32), cdiv(Ttok, tps));
  if (dtype == ADN_F32) {
    if (dx) { ADN_KERNEL("k_gconv4_dgrad", st); k_gconv4_dgrad<float><<<grid, 256, 0, st>>>((const float*)dy, w, (float*)dx, H, W, C, kh, kw, Ttok, tps); }
    { ADN_KERNEL("k_gconv4_wgrad", st); k_gconv4_wgrad<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, dw, dbias, H, W, C, kh, kw, Ttok, tps); }
  } else {
    if (dx) { ADN_KERNEL("k_gconv4_dgrad", st); k_gconv4_dgrad<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, w, (bf16*)dx, H, W, C, kh, kw, Ttok, tps); }
    { ADN_KERNEL("k_gconv4_wgrad", st); k_gconv4_wgrad<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, dbias, H, W, C, kh, kw, Ttok, tps); }
  }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // extern "C"
