// Haar-wavelet WTConv2d (models/WTConv2d.py:63-153 of the reference) for sm_100a, NCHW, bandwidth-bound.
//
// Algebra used (oracle/wtconv_oracle.py):  with ll_0 = x and ll_{i+1} = LL(DWT(pad_even(ll_i))),
//   nxt_L = 0,   nxt_i = crop( IDWT( scale_i * dwconv_k(DWT(pad(ll_i))) + [nxt_{i+1},0,0,0] ) ),
//   y = base_scale * (dwconv_k(x) + bias) + nxt_0.
// Forward  = LL pyramid (one small kernel per level) + one fused "level" kernel per level, coarsest first; the
//            finest level also applies the base conv and writes y.  The 4C sub-band tensors of the reference
//            (:118-127) are never materialised: each level kernel recomputes the Haar butterflies in shared memory.
// Backward = the same two kernels on dy (the Haar pair is orthonormal, so adjoint(IDWT) = DWT and vice versa) with
//            flipped taps, plus a correlation kernel per level for the weight / scale gradients.
#include "adn_common.cuh"

namespace adn {

constexpr int WT_TH = 16, WT_TW = 32;   // sub-band positions per CTA tile (=> 32 x 64 pixels of that level)
constexpr int WT_MAXK = 7;

struct LevelGeom { int h, w, h2, w2; };  // this level's input plane and its sub-band plane (h2 = ceil(h/2))

// the 2x2 quad (a b; c d) whose top-left pixel is (y0, x0) of a plane with row pitch w (zero beyond the plane);
// when w is even the two pixels of a row are fetched with one 2-element vector load (x0 is always even)
__device__ __forceinline__ void ld2(const float* p, float& a, float& b) { const float2 t = *reinterpret_cast<const float2*>(p); a = t.x; b = t.y; }
__device__ __forceinline__ void ld2(const bf16* p, float& a, float& b) {
  const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(p);
  const float2 f = __bfloat1622float2(t);
  a = f.x; b = f.y;
}
template <typename TI>
__device__ __forceinline__ void load_quad(const TI* __restrict__ src, int h, int w, int y0, int x0, bool weven, float& a, float& b,
                                          float& c, float& d) {
  const bool yb = y0 + 1 < h;
  if (weven) {
    ld2(src + (long long)y0 * w + x0, a, b);
    if (yb) ld2(src + (long long)(y0 + 1) * w + x0, c, d); else { c = 0.f; d = 0.f; }
  } else {
    const bool xb = x0 + 1 < w;
    a = ldf(src + (long long)y0 * w + x0);
    b = xb ? ldf(src + (long long)y0 * w + x0 + 1) : 0.f;
    c = yb ? ldf(src + (long long)(y0 + 1) * w + x0) : 0.f;
    d = (xb && yb) ? ldf(src + (long long)(y0 + 1) * w + x0 + 1) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// LL band only: out[p][y][x] = 0.5 * (a + b + c + d) of the (zero-padded) 2x2 quad.   thread = one output
// ---------------------------------------------------------------------------------------------
template <typename TI>
__global__ void k_haar_ll(const TI* __restrict__ in, float* __restrict__ out, long long planes, int h, int w, int h2,
                          int w2) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = planes * h2 * w2;
  if (idx >= total) return;
  int x = (int)(idx % w2), y = (int)((idx / w2) % h2);
  long long p = idx / ((long long)w2 * h2);
  const TI* src = in + p * h * w;
  int y0 = 2 * y, x0 = 2 * x;
  bool yb = y0 + 1 < h, xb = x0 + 1 < w;
  float a = ldf(src + (long long)y0 * w + x0);
  float b = xb ? ldf(src + (long long)y0 * w + x0 + 1) : 0.f;
  float c = yb ? ldf(src + (long long)(y0 + 1) * w + x0) : 0.f;
  float d = (xb && yb) ? ldf(src + (long long)(y0 + 1) * w + x0 + 1) : 0.f;
  out[idx] = 0.5f * (a + b + c + d);
}

// Same, for even h and w % 4 == 0: one thread = two adjacent outputs from two aligned 4-element row segments
template <typename TI>
__global__ void k_haar_ll_vec(const TI* __restrict__ in, float* __restrict__ out, long long planes, int h, int w) {
  const int w2 = w >> 1, h2 = h >> 1, wq = w >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = planes * h2 * wq;
  if (idx >= total) return;
  const int xq = (int)(idx % wq), y = (int)((idx / wq) % h2);
  const long long p = idx / ((long long)wq * h2);
  const TI* r0 = in + (p * h + 2 * y) * w + 4 * xq;
  float t[4], u[4];
  ld4(r0, t);
  ld4(r0 + w, u);
  float2 o = make_float2(0.5f * (t[0] + t[1] + u[0] + u[1]), 0.5f * (t[2] + t[3] + u[2] + u[3]));
  *reinterpret_cast<float2*>(out + (p * h2 + y) * w2 + 2 * xq) = o;
}

// ---------------------------------------------------------------------------------------------
// Fused level kernel.  One CTA = one (plane, 16x32 sub-band tile); 128 threads, each a 1x4 strip of positions.
//   sub  = DWT(pad(in))                       (in shared memory, with a k/2 halo)
//   t    = scale[ch] * conv_k(sub; Wl[ch])    (taps flipped when FLIP: the transposed conv of the backward pass)
//   t[LL] += coarse                           (nxt_{i+1} forward / dll_{i+1} backward; may be NULL)
//   o    = crop(IDWT(t))
//   if BASE: o += bscale[c] * (conv_k(in; Wb[c]) + bias[c])
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO, int K, bool BASE, bool FLIP>
__global__ void __launch_bounds__(128)
k_wt_level(const TI* __restrict__ in, const float* __restrict__ coarse, TO* __restrict__ out,
           const float* __restrict__ Wl, const float* __restrict__ scale, const float* __restrict__ Wb,
           const float* __restrict__ bias, const float* __restrict__ bscale, int C, LevelGeom g, int tiles_x,
           int tiles_y) {
  constexpr int R = K / 2, SH = WT_TH + 2 * R, SW = WT_TW + 2 * R, SWP = SW + 1;
  __shared__ float S[4][SH][SWP];
  __shared__ float Px[BASE ? 2 * SH : 1][BASE ? 2 * SW + 1 : 1];
  __shared__ float Wk[4][K * K];
  __shared__ float Wbk[K * K];
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const long long plane = bid / tiles_y;
  const int c = (int)(plane % C);
  const int sy0 = ty * WT_TH, sx0 = tx * WT_TW;
  const TI* src = in + plane * (long long)g.h * g.w;
  const int tid = threadIdx.x;
  const bool weven = (g.w & 1) == 0 && ((plane * (long long)g.h * g.w) & 1) == 0;
  for (int i = tid; i < 4 * K * K; i += 128) {
    int band = i / (K * K), t = i % (K * K);
    Wk[band][FLIP ? K * K - 1 - t : t] = Wl[((long long)c * 4 + band) * K * K + t];
  }
  if (BASE)
    for (int i = tid; i < K * K; i += 128) Wbk[FLIP ? K * K - 1 - i : i] = Wb[(long long)c * K * K + i];
  for (int i = tid; i < SH * SW; i += 128) {
    int sy = i / SW, sx = i % SW;
    int gy = sy0 - R + sy, gx = sx0 - R + sx;
    float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
    if (gy >= 0 && gy < g.h2 && gx >= 0 && gx < g.w2) load_quad(src, g.h, g.w, 2 * gy, 2 * gx, weven, a, b, cc, d);
    S[0][sy][sx] = 0.5f * (a + b + cc + d);
    S[1][sy][sx] = 0.5f * (a + b - cc - d);
    S[2][sy][sx] = 0.5f * (a - b + cc - d);
    S[3][sy][sx] = 0.5f * (a - b - cc + d);
    if (BASE) {
      Px[2 * sy][2 * sx] = a; Px[2 * sy][2 * sx + 1] = b;
      Px[2 * sy + 1][2 * sx] = cc; Px[2 * sy + 1][2 * sx + 1] = d;
    }
  }
  __syncthreads();
  const int py = tid >> 3, px = (tid & 7) * 4;  // strip origin inside the tile
  const int gy = sy0 + py;
  if (gy >= g.h2) return;
  float t[4][4];
#pragma unroll
  for (int band = 0; band < 4; ++band) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < K; ++a) {
      float v[4 + K - 1], wv[K];
#pragma unroll
      for (int j = 0; j < 4 + K - 1; ++j) v[j] = S[band][py + a][px + j];
#pragma unroll
      for (int bb = 0; bb < K; ++bb) wv[bb] = Wk[band][a * K + bb];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int bb = 0; bb < K; ++bb) acc[j] = fmaf(wv[bb], v[j + bb], acc[j]);
    }
    const float sc = scale[c * 4 + band];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[band][j] = sc * acc[j];
  }
  float o[2][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int gx = sx0 + px + j;
    float ll = t[0][j];
    if (coarse != nullptr && gx < g.w2) ll += coarse[(plane * g.h2 + gy) * g.w2 + gx];
    float b1 = t[1][j], b2 = t[2][j], b3 = t[3][j];
    o[0][2 * j] = 0.5f * (ll + b1 + b2 + b3);
    o[0][2 * j + 1] = 0.5f * (ll + b1 - b2 - b3);
    o[1][2 * j] = 0.5f * (ll - b1 + b2 - b3);
    o[1][2 * j + 1] = 0.5f * (ll - b1 - b2 + b3);
  }
  if (BASE) {
    const float bs = bscale[c], bi = bias ? bias[c] : 0.f;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bi;
      // pixel (2*py + rr, 2*px + j) of the tile sits at Px[2*(py+R) + rr][2*(px+R) + j]
#pragma unroll
      for (int a = 0; a < K; ++a) {
        float v[8 + K - 1], wv[K];
#pragma unroll
        for (int j = 0; j < 8 + K - 1; ++j) v[j] = Px[2 * (py + R) + rr + a - R][2 * (px + R) - R + j];
#pragma unroll
        for (int bb = 0; bb < K; ++bb) wv[bb] = Wbk[a * K + bb];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int bb = 0; bb < K; ++bb) acc[j] = fmaf(wv[bb], v[j + bb], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[rr][j] += bs * acc[j];
    }
  }
  TO* dst = out + plane * (long long)g.h * g.w;
  const int x0 = 2 * (sx0 + px);                      // multiple of 8
  const bool vec = (g.w & 7) == 0 && x0 + 8 <= g.w;   // the strip's 8 outputs of a row: two 4-element vector stores
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    int y = 2 * gy + rr;
    if (y >= g.h) continue;
    if (vec) {
      float lo[4] = {o[rr][0], o[rr][1], o[rr][2], o[rr][3]}, hi[4] = {o[rr][4], o[rr][5], o[rr][6], o[rr][7]};
      st4(dst + (long long)y * g.w + x0, lo);
      st4(dst + (long long)y * g.w + x0 + 4, hi);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int x = x0 + j;
      if (x < g.w) stf(dst + (long long)y * g.w + x, o[rr][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Correlations for the weight gradients of one level:
//   NB == 4:  Rl[c*4+band][a][b] += sum dsub[band][y][x] * sub[band][y+a-R][x+b-R],  sub = DWT(pad(xin)), dsub = DWT(pad(gin))
//   NB == 1:  Rb[c][a][b] += sum gin[y][x] * xin[y+a-R][x+b-R],  sumg[c] += sum gin            (base conv, pixel units)
// thread = (band, tap row a, group of 2 tile rows): slides a K-wide window along x.
// ---------------------------------------------------------------------------------------------
template <typename TX, typename TG, int K, int NB>
__global__ void k_wt_wgrad(const TX* __restrict__ xin, const TG* __restrict__ gin, float* __restrict__ Rout,
                           float* __restrict__ sumg, int C, LevelGeom g, int tiles_x, int tiles_y) {
  // tile of the correlation domain: 16 x 32 sub-band positions (NB == 4) or 32 x 64 pixels (base conv: less halo, 4x fewer CTAs)
  constexpr int TH = NB == 4 ? WT_TH : 2 * WT_TH, TW = NB == 4 ? WT_TW : 2 * WT_TW;
  constexpr int R = K / 2, SH = TH + 2 * R, SW = TW + 2 * R, SWP = SW + 1;
  __shared__ float S[NB][SH][SWP];
  __shared__ float Dt[NB][TH][TW + 1];
  __shared__ float red[NB * K * K];
  __shared__ float redsum;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const long long plane = bid / tiles_y;
  const int c = (int)(plane % C);
  const int sy0 = ty * TH, sx0 = tx * TW;
  // extent of the tiled domain: sub-band plane (NB==4) or the pixel plane itself (NB==1)
  const int dh = NB == 4 ? g.h2 : g.h, dw = NB == 4 ? g.w2 : g.w;
  const TX* xs = xin + plane * (long long)g.h * g.w;
  const TG* gs = gin + plane * (long long)g.h * g.w;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool weven = (g.w & 1) == 0 && ((plane * (long long)g.h * g.w) & 1) == 0;
  for (int i = tid; i < NB * K * K; i += nt) red[i] = 0.f;
  if (tid == 0) redsum = 0.f;
  for (int i = tid; i < SH * SW; i += nt) {
    int sy = i / SW, sx = i % SW;
    int gy = sy0 - R + sy, gx = sx0 - R + sx;
    bool in = gy >= 0 && gy < dh && gx >= 0 && gx < dw;
    if (NB == 4) {
      float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
      if (in) load_quad(xs, g.h, g.w, 2 * gy, 2 * gx, weven, a, b, cc, d);
      S[0][sy][sx] = 0.5f * (a + b + cc + d);
      S[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + b - cc - d);
      S[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - b + cc - d);
      S[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - b - cc + d);
    } else {
      S[0][sy][sx] = in ? ldf(xs + (long long)gy * g.w + gx) : 0.f;
    }
  }
  float lsum = 0.f;
  for (int i = tid; i < TH * TW; i += nt) {
    int sy = i / TW, sx = i % TW;
    int gy = sy0 + sy, gx = sx0 + sx;
    bool in = gy < dh && gx < dw;
    if (NB == 4) {
      float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
      if (in) load_quad(gs, g.h, g.w, 2 * gy, 2 * gx, weven, a, b, cc, d);
      Dt[0][sy][sx] = 0.5f * (a + b + cc + d);
      Dt[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + b - cc - d);
      Dt[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - b + cc - d);
      Dt[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - b - cc + d);
    } else {
      float v = in ? ldf(gs + (long long)gy * g.w + gx) : 0.f;
      Dt[0][sy][sx] = v;
      lsum += v;
    }
  }
  __syncthreads();
  // roles: tid -> (row group rg, tap row a, band, x segment): slides a K-wide window along x for its tap row.  The window
  // walk is fully unrolled (the shift is register renaming: 2 LDS + K FMA per position).  NB == 4: 8 groups of 2 rows over
  // the whole tile width; NB == 1 (base conv): 32 single rows, so that all K * 32 threads work.
  constexpr int XSEG = 1, XW = TW / XSEG, RPT = NB == 4 ? 2 : 1, NRG = TH / RPT;
  const int role = tid;
  if (role < NB * K * NRG * XSEG) {
    const int rg = role % NRG, a = (role / NRG) % K, band = (role / (NRG * K)) % NB, x0 = (role / (NRG * K * NB)) * XW;
    float acc[K];
#pragma unroll
    for (int bb = 0; bb < K; ++bb) acc[bb] = 0.f;
#pragma unroll
    for (int yy = 0; yy < RPT; ++yy) {
      const int y = rg * RPT + yy;
      const float* srow = &S[band][y + a][x0];
      const float* drow = &Dt[band][y][x0];
      float win[K];
#pragma unroll
      for (int bb = 0; bb < K - 1; ++bb) win[bb + 1] = srow[bb];
#pragma unroll
      for (int x = 0; x < XW; ++x) {
#pragma unroll
        for (int bb = 0; bb < K - 1; ++bb) win[bb] = win[bb + 1];
        win[K - 1] = srow[x + K - 1];
        const float dv = drow[x];
#pragma unroll
        for (int bb = 0; bb < K; ++bb) acc[bb] = fmaf(dv, win[bb], acc[bb]);
      }
    }
    // the NRG row groups of one (x segment, band, tap row) are consecutive lanes: add them up with shuffles (whole warps are
    // inside this branch) so that one lane per tap issues the shared-memory atomic (a CAS loop for fp32: it must not contend)
#pragma unroll
    for (int bb = 0; bb < K; ++bb) {
      float v = acc[bb];
#pragma unroll
      for (int m = 1; m < NRG; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (rg == 0) atomicAdd(&red[(band * K + a) * K + bb], v);
    }
  }
  if (NB == 1) {
    lsum = warp_sum(lsum);
    if ((tid & 31) == 0 && lsum != 0.f) atomicAdd(&redsum, lsum);
  }
  __syncthreads();
  for (int i = tid; i < NB * K * K; i += nt) atomicAdd(Rout + (long long)c * NB * K * K + i, red[i]);
  if (NB == 1 && tid == 0 && sumg) atomicAdd(sumg + c, redsum);
}

// dW_i = scale_i * R_i ; dscale_i[ch] = sum_ab W_i R_i ; dW_b = bs * R_b ; dbs = sum W_b R_b + bias * sumdy ; dbias = bs * sumdy
struct WtFinalize {
  const float* Rl[ADN_WT_MAX_LEVELS];
  const float *Rb, *sumg;
};

__global__ void __launch_bounds__(32)
k_wt_finalize(WtFinalize a, WtWeights w, WtWeightGrads g, int C, int K2, int levels) {
  // blockIdx.x < levels*4C: one warp per (level, sub-band channel); then C warps for the base conv channels
  const int lane = threadIdx.x;
  int id = blockIdx.x;
  if (id < levels * 4 * C) {
    const int l = id / (4 * C), ch = id % (4 * C);
    const float sc = w.wavelet_scale_w[l][ch];
    float ds = 0.f;
    for (int t = lane; t < K2; t += 32) {
      const float r = a.Rl[l][ch * K2 + t];
      if (g.wavelet_conv_w[l]) g.wavelet_conv_w[l][ch * K2 + t] = sc * r;
      ds = fmaf(w.wavelet_conv_w[l][ch * K2 + t], r, ds);
    }
    ds = warp_sum(ds);
    if (lane == 0 && g.wavelet_scale_w[l]) g.wavelet_scale_w[l][ch] = ds;
    return;
  }
  const int c = id - levels * 4 * C;
  if (c >= C) return;
  const float bs = w.base_scale_w[c], sg = a.sumg[c];
  float ds = 0.f;
  for (int t = lane; t < K2; t += 32) {
    const float r = a.Rb[c * K2 + t];
    if (g.base_conv_w) g.base_conv_w[c * K2 + t] = bs * r;
    ds = fmaf(w.base_conv_w[c * K2 + t], r, ds);
  }
  ds = warp_sum(ds);
  if (lane == 0) {
    if (w.base_conv_b) {
      ds += w.base_conv_b[c] * sg;
      if (g.base_conv_b) g.base_conv_b[c] = bs * sg;
    }
    if (g.base_scale_w) g.base_scale_w[c] = ds;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct WtPlan {
  int L;
  long long planes;
  LevelGeom g[ADN_WT_MAX_LEVELS];
  size_t pyr_off[ADN_WT_MAX_LEVELS + 1];  // float offsets of ll_1 .. ll_{L-1} (index i -> ll_i), pyr_off[L] = total
};

static WtPlan make_plan(const WtShape& s) {
  WtPlan p;
  p.L = s.levels;
  p.planes = (long long)s.B * s.C;
  int h = s.H, w = s.W;
  size_t off = 0;
  for (int i = 0; i < p.L; ++i) {
    p.g[i].h = h; p.g[i].w = w;
    p.g[i].h2 = (h + 1) / 2; p.g[i].w2 = (w + 1) / 2;
    p.pyr_off[i] = off;
    if (i >= 1) off += align_up((size_t)p.planes * h * w, 64);
    h = p.g[i].h2; w = p.g[i].w2;
  }
  p.pyr_off[p.L] = off;
  return p;
}

static int wt_validate(const WtShape* s) {
  ADN_REQUIRE(s != nullptr, ADN_ERR_NULL, "WtShape is NULL");
  ADN_REQUIRE(s->B > 0 && s->C > 0 && s->H > 0 && s->W > 0, ADN_ERR_SHAPE, "wtconv: B/C/H/W must be positive");
  ADN_REQUIRE(s->k == 1 || s->k == 3 || s->k == 5 || s->k == 7, ADN_ERR_SHAPE, "wtconv: kernel size %d not in {1,3,5,7}", s->k);
  ADN_REQUIRE(s->levels >= 1 && s->levels <= ADN_WT_MAX_LEVELS, ADN_ERR_SHAPE, "wtconv: levels %d not in 1..%d", s->levels,
              ADN_WT_MAX_LEVELS);
  ADN_REQUIRE(s->dtype == ADN_F32 || s->dtype == ADN_BF16, ADN_ERR_DTYPE, "wtconv: unsupported dtype %d", s->dtype);
  ADN_REQUIRE((long long)s->B * s->C * s->H * s->W < (1LL << 40), ADN_ERR_SHAPE, "wtconv: tensor too large");
  return ADN_OK;
}

// fp32 scratch: [pyramid-shaped buffer A | pyramid-shaped buffer B | accumulators]
struct WtAcc {
  float* Rl[ADN_WT_MAX_LEVELS];
  float *Rb, *sumg;
  size_t floats;
};
static WtAcc carve_acc(float* base, const WtShape& s) {
  WtAcc a;
  size_t off = 0;
  const size_t K2 = (size_t)s.k * s.k;
  for (int l = 0; l < s.levels; ++l) { a.Rl[l] = base ? base + off : nullptr; off += align_up(4 * s.C * K2, 64); }
  a.Rb = base ? base + off : nullptr; off += align_up(s.C * K2, 64);
  a.sumg = base ? base + off : nullptr; off += align_up((size_t)s.C, 64);
  a.floats = off;
  return a;
}

template <typename TI, typename TO, bool BASE, bool FLIP>
static void launch_level(cudaStream_t st, int k, const TI* in, const float* coarse, TO* out, const float* Wl,
                         const float* scale, const float* Wb, const float* bias, const float* bscale, int C,
                         const LevelGeom& g, long long planes) {
  int tx = cdiv(g.w2, WT_TW), ty = cdiv(g.h2, WT_TH);
  long long blocks = planes * tx * ty;
#define ADN_WT_LAUNCH(KK)                                                                                        \
  {                                                                                                              \
    ADN_KERNEL("k_wt_level", st);                                                                                \
    k_wt_level<TI, TO, KK, BASE, FLIP><<<(unsigned)blocks, 128, 0, st>>>(in, coarse, out, Wl, scale, Wb, bias, bscale, C, \
                                                                         g, tx, ty);                             \
  }
  switch (k) {
    case 1: ADN_WT_LAUNCH(1); break;
    case 3: ADN_WT_LAUNCH(3); break;
    case 5: ADN_WT_LAUNCH(5); break;
    default: ADN_WT_LAUNCH(7); break;
  }
#undef ADN_WT_LAUNCH
}

template <typename TX, typename TG, int NB>
static void launch_wgrad(cudaStream_t st, int k, const TX* xin, const TG* gin, float* Rout, float* sumg, int C,
                         const LevelGeom& g, long long planes) {
  int dh = NB == 4 ? g.h2 : g.h, dw = NB == 4 ? g.w2 : g.w;
  int tx = cdiv(dw, NB == 4 ? WT_TW : 2 * WT_TW), ty = cdiv(dh, NB == 4 ? WT_TH : 2 * WT_TH);   // tile of k_wt_wgrad<.., NB>
  long long blocks = planes * tx * ty;
  int threads = (((NB == 4 ? 4 * k * 8 : k * 32) + 31) / 32) * 32;   // one thread per role of k_wt_wgrad
  if (threads < 64) threads = 64;
#define ADN_WT_LAUNCH(KK) \
  {                                                                                                             \
    ADN_KERNEL("k_wt_wgrad", st);                                                                               \
    k_wt_wgrad<TX, TG, KK, NB><<<(unsigned)blocks, threads, 0, st>>>(xin, gin, Rout, sumg, C, g, tx, ty);        \
  }
  switch (k) {
    case 1: ADN_WT_LAUNCH(1); break;
    case 3: ADN_WT_LAUNCH(3); break;
    case 5: ADN_WT_LAUNCH(5); break;
    default: ADN_WT_LAUNCH(7); break;
  }
#undef ADN_WT_LAUNCH
}

// Builds ll_1..ll_{L-1} of `src` into `pyr` and then runs the level kernels coarsest-first.
//   FLIP=false: forward (weights as stored, base bias applied);  FLIP=true: backward-data (dy -> dx)
template <typename T, bool FLIP>
static int run_pass(const WtShape& s, const WtPlan& p, const WtWeights& w, const T* src, T* dst, float* pyr,
                    float* chain, cudaStream_t st) {
  for (int i = 1; i < p.L; ++i) {
    float* o = pyr + p.pyr_off[i];
    long long total = p.planes * p.g[i].h * p.g[i].w;
    const LevelGeom& gi = p.g[i - 1];
    const bool vec = (gi.h % 2 == 0) && (gi.w % 4 == 0);
    if (i == 1) {
      if (vec) { ADN_KERNEL("k_haar_ll", st); k_haar_ll_vec<T><<<cdiv(total / 2, 256), 256, 0, st>>>(src, o, p.planes, gi.h, gi.w); }
      else { ADN_KERNEL("k_haar_ll", st); k_haar_ll<T><<<cdiv(total, 256), 256, 0, st>>>(src, o, p.planes, gi.h, gi.w, gi.h2, gi.w2); }
    } else {
      const float* pin = pyr + p.pyr_off[i - 1];
      if (vec) { ADN_KERNEL("k_haar_ll", st); k_haar_ll_vec<float><<<cdiv(total / 2, 256), 256, 0, st>>>(pin, o, p.planes, gi.h, gi.w); }
      else { ADN_KERNEL("k_haar_ll", st); k_haar_ll<float><<<cdiv(total, 256), 256, 0, st>>>(pin, o, p.planes, gi.h, gi.w, gi.h2, gi.w2); }
    }
  }
  for (int i = p.L - 1; i >= 0; --i) {
    const float* coarse = (i == p.L - 1) ? nullptr : chain + p.pyr_off[i + 1];
    if (i == 0) {
      launch_level<T, T, true, FLIP>(st, s.k, src, coarse, dst, w.wavelet_conv_w[0], w.wavelet_scale_w[0], w.base_conv_w,
                                     FLIP ? nullptr : w.base_conv_b, w.base_scale_w, s.C, p.g[0], p.planes);
    } else {
      launch_level<float, float, false, FLIP>(st, s.k, pyr + p.pyr_off[i], coarse, chain + p.pyr_off[i],
                                              w.wavelet_conv_w[i], w.wavelet_scale_w[i], nullptr, nullptr, nullptr, s.C,
                                              p.g[i], p.planes);
    }
  }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

static int wt_check_weights(const WtShape* s, const WtWeights* w) {
  ADN_REQUIRE(w != nullptr, ADN_ERR_NULL, "WtWeights is NULL");
  ADN_REQUIRE(w->base_conv_w && w->base_scale_w, ADN_ERR_NULL, "wtconv: base_conv / base_scale weights are NULL");
  ADN_REQUIRE(!s->has_bias || w->base_conv_b, ADN_ERR_NULL, "wtconv: has_bias set but base_conv_b is NULL");
  for (int l = 0; l < s->levels; ++l)
    ADN_REQUIRE(w->wavelet_conv_w[l] && w->wavelet_scale_w[l], ADN_ERR_NULL, "wtconv: level %d weights are NULL", l);
  return ADN_OK;
}

template <typename T>
static int wt_forward(const WtShape& s, const WtWeights& w0, const T* x, T* y, void* saved, void* ws, cudaStream_t st) {
  WtPlan p = make_plan(s);
  WtWeights w = w0;
  if (!s.has_bias) w.base_conv_b = nullptr;
  float* scratch = (float*)ws;
  float* pyr = saved ? (float*)saved : scratch;              // ll_1..ll_{L-1} of x (kept for the weight gradients)
  float* chain = scratch + (saved ? 0 : p.pyr_off[p.L]);     // nxt_1..nxt_{L-1}
  return run_pass<T, false>(s, p, w, x, y, pyr, chain, st);
}

template <typename T>
static int wt_backward(const WtShape& s, const WtWeights& w0, const T* x, const void* saved, const T* dy, T* dx,
                       const WtWeightGrads& g, void* ws, cudaStream_t st) {
  WtPlan p = make_plan(s);
  WtWeights w = w0;
  if (!s.has_bias) w.base_conv_b = nullptr;
  const float* xpyr = (const float*)saved;
  float* gpyr = (float*)ws;                       // dn_1..dn_{L-1}: LL pyramid of dy
  float* chain = gpyr + p.pyr_off[p.L];           // dll_1..dll_{L-1}
  float* accb = chain + p.pyr_off[p.L];
  WtAcc acc = carve_acc(accb, s);
  ADN_CHECK_CUDA(cudaMemsetAsync(accb, 0, acc.floats * sizeof(float), st));
  int rc = run_pass<T, true>(s, p, w, dy, dx, gpyr, chain, st);
  if (rc) return rc;
  // weight-gradient correlations
  launch_wgrad<T, T, 1>(st, s.k, x, dy, acc.Rb, acc.sumg, s.C, p.g[0], p.planes);
  launch_wgrad<T, T, 4>(st, s.k, x, dy, acc.Rl[0], nullptr, s.C, p.g[0], p.planes);
  for (int i = 1; i < p.L; ++i)
    launch_wgrad<float, float, 4>(st, s.k, xpyr + p.pyr_off[i], gpyr + p.pyr_off[i], acc.Rl[i], nullptr, s.C, p.g[i],
                                  p.planes);
  WtFinalize f;
  for (int l = 0; l < ADN_WT_MAX_LEVELS; ++l) f.Rl[l] = l < s.levels ? acc.Rl[l] : nullptr;
  f.Rb = acc.Rb; f.sumg = acc.sumg;
  { ADN_KERNEL("k_wt_finalize", st); k_wt_finalize<<<s.levels * 4 * s.C + s.C, 32, 0, st>>>(f, w, g, s.C, s.k * s.k, s.levels); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" {

int wtconv_workspace_bytes(const WtShape* s, size_t* saved_bytes, size_t* fwd_ws, size_t* bwd_ws) {
  int rc = wt_validate(s);
  if (rc) return rc;
  WtPlan p = make_plan(*s);
  size_t pyr = p.pyr_off[p.L] * sizeof(float);
  if (saved_bytes) *saved_bytes = pyr;
  if (fwd_ws) *fwd_ws = 2 * pyr + 256;
  if (bwd_ws) *bwd_ws = 2 * pyr + carve_acc(nullptr, *s).floats * sizeof(float) + 256;
  return ADN_OK;
}

int wtconv_forward(const WtShape* s, const WtWeights* w, const void* x, void* y, void* saved, void* workspace,
                   void* stream) {
  int rc = wt_validate(s);
  if (rc) return rc;
  rc = wt_check_weights(s, w);
  if (rc) return rc;
  ADN_REQUIRE(x && y && workspace, ADN_ERR_NULL, "wtconv_forward: x / y / workspace must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->dtype == ADN_F32) return wt_forward<float>(*s, *w, (const float*)x, (float*)y, saved, workspace, st);
  return wt_forward<bf16>(*s, *w, (const bf16*)x, (bf16*)y, saved, workspace, st);
}

int wtconv_backward(const WtShape* s, const WtWeights* w, const void* x, const void* saved, const void* dy, void* dx,
                    const WtWeightGrads* g, void* workspace, void* stream) {
  int rc = wt_validate(s);
  if (rc) return rc;
  rc = wt_check_weights(s, w);
  if (rc) return rc;
  ADN_REQUIRE(x && dy && dx && g && workspace && (saved || s->levels == 1), ADN_ERR_NULL,
              "wtconv_backward: x / saved / dy / dx / grads / workspace must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->dtype == ADN_F32)
    return wt_backward<float>(*s, *w, (const float*)x, saved, (const float*)dy, (float*)dx, *g, workspace, st);
  return wt_backward<bf16>(*s, *w, (const bf16*)x, saved, (const bf16*)dy, (bf16*)dx, *g, workspace, st);
}

}  // extern "C"
