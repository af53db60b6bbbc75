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
#include <string.h>

#include "adn_common.cuh"
#include "tma_utils.cuh"

namespace adn {
using namespace adn::sm100;


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
// Fused level kernel.  One CTA = one (plane, 32 x 32 sub-band tile) = 64 x 64 pixels of that level; 64 threads, each a
// 4 x 4 block of sub-band positions (8 x 8 pixels).
//   sub  = DWT(pad(in))                       (in shared memory, with a k/2 halo)
//   t    = scale[ch] * conv_k(sub; Wl[ch])    (taps flipped when FLIP: the transposed conv of the backward pass)
//   t[LL] += coarse                           (nxt_{i+1} forward / dll_{i+1} backward; may be NULL)
//   o    = crop(IDWT(t))
//   if BASE: o += bscale[c] * (conv_k(in; Wb[c]) + bias[c])
// The k x k taps make this stage FMA-bound, and the CUDA cores only reach their FMA rate when shared memory delivers less
// than one word per four FMAs (128 FMA lanes, 32 banks per SM).  Round 1's 1 x 4 strips needed 13 words per 20 FMAs and ran
// 3x off the FMA rate; a 2 x 4 block reads (2 + k - 1) rows of (4 + k - 1) values as aligned 16-byte vectors once per band
// for 8 k^2 FMAs (k = 5: 48 words per 200 FMAs), with the k^2 taps of the band in registers.  (A 4 x 4 block has the better
// ratio but halves the warps an SM can hold - 512 bytes of tile per thread - and measured slower: the tile staging, one
// exposed HBM latency per CTA, then has too few other warps to hide behind.)
// ---------------------------------------------------------------------------------------------
constexpr int WT_BH = 2;                                    // a thread owns WT_BH x 4 positions
// sub-band tile edge TT: 32 (128 threads) for planes with more than 16 x 16 sub-band positions, else 16 (32 threads) - a
// 32 x 32 plane of the model's deeper stages (16 x 16 positions per band) filled a quarter of the large tile
__host__ __device__ constexpr int wt_threads(int TT) { return (TT / 4) * (TT / WT_BH); }
template <int K, int TT> struct WtTile {
  static constexpr int R = K / 2, SH = TT + 2 * R;
  static constexpr int VW = ((4 + K - 1) + 3) / 4 * 4;          // values a thread reads per row: whole 16-byte vectors
  static constexpr int SP = ((TT - 4 + VW) + 3) / 4 * 4;      // row pitch of S (the last block's vectors stay inside the row)
  static constexpr int POFF = (4 - R % 4) % 4;                  // column shift of the pixel tile: a block's first column is 16-byte aligned
  static constexpr int PVW = ((8 + K - 1) + 3) / 4 * 4;
  static constexpr int PW_WRITE = 2 * (TT + 2 * R) + POFF, PW_READ = 2 * (TT - 4) + R + POFF + PVW;
  static constexpr int PH = 2 * SH, PP = ((PW_WRITE > PW_READ ? PW_WRITE : PW_READ) + 3) / 4 * 4;
};

// Halo tile staging.  TMA == true: the raw input tile (pixels of this level, storage dtype) arrives by ONE bulk tensor copy
// (cp.async.bulk.tensor, zero fill outside the plane = the conv's zero padding AND the odd-size padding of the DWT) issued by
// one thread and tracked by an mbarrier; the CTA is persistent and the copy of its NEXT tile is issued as soon as the current
// raw tile has been transformed, so it lands while the FMAs of the current tile run.  TMA == false (row pitch or plane size
// not a multiple of 16 bytes): the same tile is gathered with per-thread loads.
template <typename TI, int K, int TT> struct WtRaw {
  static constexpr int PER = 4 / (int)sizeof(TI);      // pixels per 32-bit TMA element
  // The box must START on a 16-byte boundary of the row (measured: a bf16 tile whose first pixel is 8 bytes into a 16-byte
  // unit faults with "illegal instruction" on the first copy; the fp32 tile, 16 bytes in, copies fine), so the tile is widened
  // to the left by XOFF pixels, and its rows are whole 32-byte units.
  static constexpr int APX = 16 / (int)sizeof(TI), XOFF = (APX - (2 * WtTile<K, TT>::R) % APX) % APX;
  static constexpr int PXQ = 32 / (int)sizeof(TI);
  static constexpr int BH = 2 * WtTile<K, TT>::SH, BW = (2 * (TT + 2 * WtTile<K, TT>::R) + XOFF + PXQ - 1) / PXQ * PXQ;
  static constexpr size_t BYTES = ((size_t)BH * BW * sizeof(TI) + 127) / 128 * 128;
};

template <typename TI, typename TO, int K, bool BASE, bool FLIP, bool TMA, int TT>
__global__ void __launch_bounds__(wt_threads(TT), 4)
k_wt_level(const __grid_constant__ CUtensorMap map, const TI* __restrict__ in, const float* __restrict__ coarse, TO* __restrict__ out,
           const float* __restrict__ Wl, const float* __restrict__ scale, const float* __restrict__ Wb,
           const float* __restrict__ bias, const float* __restrict__ bscale, int C, LevelGeom g, int tiles_x,
           int tiles_y, int total) {
  using G = WtTile<K, TT>;
  using RW = WtRaw<TI, K, TT>;
  constexpr int WT_T = TT, WT_THREADS = wt_threads(TT);
  constexpr int R = G::R, SH = G::SH, SW = WT_T + 2 * R, SP = G::SP;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);      // TMA destinations: 128-byte aligned
  TI* raw = reinterpret_cast<TI*>(smem_raw);                                          // [BH][BW]   (TMA only)
  float* smem = reinterpret_cast<float*>(smem_raw + (TMA ? RW::BYTES : 0));
  float (*S)[SH][SP] = reinterpret_cast<float (*)[SH][SP]>(smem);                    // [4][SH][SP]
  float (*Px)[G::PP] = reinterpret_cast<float (*)[G::PP]>(smem + 4 * SH * SP);       // [PH][PP]   (BASE only)
  __shared__ float Wk[4][K * K];
  __shared__ float Wbk[K * K];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int per_plane = tiles_x * tiles_y;
  if (TMA) {
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); tma::tma_prefetch_desc(&map); }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total) {
      const int t0 = blockIdx.x, pl = t0 / per_plane, rem = t0 - pl * per_plane, tyy = rem / tiles_x, txx = rem - tyy * tiles_x;
      mbar_expect_tx(&bar, (uint32_t)((size_t)RW::BH * RW::BW * sizeof(TI)));
      tma::tma_load_3d(smem_u32(raw), &map, (2 * (txx * WT_T - R) - RW::XOFF) / RW::PER, 2 * (tyy * WT_T - R), pl, &bar);
    }
  }
  uint32_t phase = 0;
  const int py = (tid / (WT_T / 4)) * WT_BH, px = (tid % (WT_T / 4)) * 4;  // block origin inside the tile
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int pl = tile / per_plane, rem = tile - pl * per_plane, ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const long long plane = pl;
    const int c = (int)(plane % C);
    const int sy0 = ty * WT_T, sx0 = tx * WT_T;
    for (int i = tid; i < 4 * K * K; i += WT_THREADS) {
      int band = i / (K * K), t = i % (K * K);
      Wk[band][FLIP ? K * K - 1 - t : t] = Wl[((long long)c * 4 + band) * K * K + t];
    }
    if (BASE)
      for (int i = tid; i < K * K; i += WT_THREADS) Wbk[FLIP ? K * K - 1 - i : i] = Wb[(long long)c * K * K + i];
    constexpr int NQ = (SH * SW + WT_THREADS - 1) / WT_THREADS;
    if (TMA) {
      if (!mbar_wait(&bar, phase)) __trap();      // a fault in the async pipe must be loud, not a hang
      phase ^= 1;
#pragma unroll 2
      for (int u = 0; u < NQ; ++u) {
        const int i = u * WT_THREADS + tid;
        if (i >= SH * SW) break;
        const int sy = i / SW, sx = i - sy * SW;
        float a, b, cc, d;
        load_quad(raw, RW::BH, RW::BW, 2 * sy, 2 * sx + RW::XOFF, true, a, b, cc, d);
        S[0][sy][sx] = 0.5f * (a + b + cc + d);
        S[1][sy][sx] = 0.5f * (a + b - cc - d);
        S[2][sy][sx] = 0.5f * (a - b + cc - d);
        S[3][sy][sx] = 0.5f * (a - b - cc + d);
        if (BASE) {
          Px[2 * sy][2 * sx + G::POFF] = a; Px[2 * sy][2 * sx + 1 + G::POFF] = b;
          Px[2 * sy + 1][2 * sx + G::POFF] = cc; Px[2 * sy + 1][2 * sx + 1 + G::POFF] = d;
        }
      }
    } else {
      const TI* src = in + plane * (long long)g.h * g.w;
      const bool weven = (g.w & 1) == 0 && ((plane * (long long)g.h * g.w) & 1) == 0;
      // ALL of a thread's quads are loaded before the first is used (one exposed memory latency per tile)
      float qa[NQ], qb[NQ], qc[NQ], qd[NQ];
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        const int i = u * WT_THREADS + tid;
        const int sy = i / SW, sx = i - sy * SW;
        const int gy = sy0 - R + sy, gx = sx0 - R + sx;
        qa[u] = qb[u] = qc[u] = qd[u] = 0.f;
        if (i < SH * SW && gy >= 0 && gy < g.h2 && gx >= 0 && gx < g.w2) load_quad(src, g.h, g.w, 2 * gy, 2 * gx, weven, qa[u], qb[u], qc[u], qd[u]);
      }
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        const int i = u * WT_THREADS + tid;
        if (i >= SH * SW) continue;
        const int sy = i / SW, sx = i - sy * SW;
        const float a = qa[u], b = qb[u], cc = qc[u], d = qd[u];
        S[0][sy][sx] = 0.5f * (a + b + cc + d);
        S[1][sy][sx] = 0.5f * (a + b - cc - d);
        S[2][sy][sx] = 0.5f * (a - b + cc - d);
        S[3][sy][sx] = 0.5f * (a - b - cc + d);
        if (BASE) {
          Px[2 * sy][2 * sx + G::POFF] = a; Px[2 * sy][2 * sx + 1 + G::POFF] = b;
          Px[2 * sy + 1][2 * sx + G::POFF] = cc; Px[2 * sy + 1][2 * sx + 1 + G::POFF] = d;
        }
      }
    }
    __syncthreads();                                        // S / Px / weights ready; the raw tile is consumed
    if (TMA && tid == 0 && tile + (int)gridDim.x < total) {
      const int t1 = tile + gridDim.x, pl1 = t1 / per_plane, rem1 = t1 - pl1 * per_plane, ty1 = rem1 / tiles_x, tx1 = rem1 - ty1 * tiles_x;
      mbar_expect_tx(&bar, (uint32_t)((size_t)RW::BH * RW::BW * sizeof(TI)));
      tma::tma_load_3d(smem_u32(raw), &map, (2 * (tx1 * WT_T - R) - RW::XOFF) / RW::PER, 2 * (ty1 * WT_T - R), pl1, &bar);
    }
    if (sy0 + py < g.h2 && sx0 + px < g.w2) {              // else: the whole block lies outside the plane
      float t[4][WT_BH][4];                                  // [band][row][col]
#pragma unroll
      for (int band = 0; band < 4; ++band) {
        float wv[K * K];
#pragma unroll
        for (int i = 0; i < K * K; ++i) wv[i] = Wk[band][i];
        float acc[WT_BH][4];
#pragma unroll
        for (int i = 0; i < WT_BH; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int r = 0; r < WT_BH + K - 1; ++r) {
          float v[G::VW];
#pragma unroll
          for (int q = 0; q < G::VW; q += 4) {
            const float4 f = *reinterpret_cast<const float4*>(&S[band][py + r][px + q]);
            v[q] = f.x; v[q + 1] = f.y; v[q + 2] = f.z; v[q + 3] = f.w;
          }
#pragma unroll
          for (int i = 0; i < WT_BH; ++i) {
            const int a = r - i;                               // tap row that maps input row r to output row i
            if (a < 0 || a >= K) continue;
#pragma unroll
            for (int bb = 0; bb < K; ++bb)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[a * K + bb], v[j + bb], acc[i][j]);
          }
        }
        const float sc = scale[c * 4 + band];
#pragma unroll
        for (int i = 0; i < WT_BH; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) t[band][i][j] = sc * acc[i][j];
      }
      TO* dst = out + plane * (long long)g.h * g.w;
      const float bs = BASE ? bscale[c] : 0.f, bi = (BASE && bias) ? bias[c] : 0.f;
      const int x0 = 2 * (sx0 + px);                      // multiple of 8
      const bool vec = (g.w & 7) == 0 && x0 + 8 <= g.w;   // the block's 8 outputs of a row: two 4-element vector stores
      // two pixel rows (one sub-band row) at a time: IDWT, base conv, store
#pragma unroll
      for (int i = 0; i < WT_BH; ++i) {
        const int gy = sy0 + py + i;
        if (gy >= g.h2) break;
        float o[2][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gx = sx0 + px + j;
          float ll = t[0][i][j];
          if (coarse != nullptr && gx < g.w2) ll += coarse[(plane * g.h2 + gy) * g.w2 + gx];
          const float b1 = t[1][i][j], b2 = t[2][i][j], b3 = t[3][i][j];
          o[0][2 * j] = 0.5f * (ll + b1 + b2 + b3);
          o[0][2 * j + 1] = 0.5f * (ll + b1 - b2 - b3);
          o[1][2 * j] = 0.5f * (ll - b1 + b2 - b3);
          o[1][2 * j + 1] = 0.5f * (ll - b1 - b2 + b3);
        }
        if (BASE) {
          float wb[K * K];
#pragma unroll
          for (int q = 0; q < K * K; ++q) wb[q] = Wbk[q];
          float acc[2][8];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[rr][j] = bi;
          // pixel (2 * (py + i) + rr, 2 * px + j) of the tile sits at Px[2 * (py + i + R) + rr][2 * (px + R) + j + POFF]; tap (a, bb)
          // reads row + a - R, column + bb - R: rows 2 * (py + i) + R + rr + a, columns 2 * px + R + POFF + j + bb
#pragma unroll
          for (int r = 0; r < 2 + K - 1; ++r) {
            float v[G::PVW];
#pragma unroll
            for (int q = 0; q < G::PVW; q += 4) {
              const float4 f = *reinterpret_cast<const float4*>(&Px[2 * (py + i) + R + r][2 * px + R + G::POFF + q]);
              v[q] = f.x; v[q + 1] = f.y; v[q + 2] = f.z; v[q + 3] = f.w;
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int a = r - rr;
              if (a < 0 || a >= K) continue;
#pragma unroll
              for (int bb = 0; bb < K; ++bb)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[rr][j] = fmaf(wb[a * K + bb], v[j + bb], acc[rr][j]);
            }
          }
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[rr][j] = fmaf(bs, acc[rr][j], o[rr][j]);
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int y = 2 * gy + rr;
          if (y >= g.h) continue;
          if (vec) {
            float lo[4] = {o[rr][0], o[rr][1], o[rr][2], o[rr][3]}, hi[4] = {o[rr][4], o[rr][5], o[rr][6], o[rr][7]};
            st4(dst + (long long)y * g.w + x0, lo);
            st4(dst + (long long)y * g.w + x0 + 4, hi);
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int x = x0 + j;
            if (x < g.w) stf(dst + (long long)y * g.w + x, o[rr][j]);
          }
        }
      }
    }
    __syncthreads();                                        // S / Px / weights may be overwritten by the next tile
  }
}
template <typename TI, int K, bool BASE, bool TMA, int TT> static constexpr size_t wt_level_smem() {
  using G = WtTile<K, TT>;
  return 128 + (TMA ? WtRaw<TI, K, TT>::BYTES : 0) + sizeof(float) * (4 * (size_t)G::SH * G::SP + (BASE ? (size_t)G::PH * G::PP : 0));
}

// ---------------------------------------------------------------------------------------------
// Correlations for the weight gradients of one level:
//   NB == 4:  Rl[c*4+band][a][b] += sum dsub[band][y][x] * sub[band][y+a-R][x+b-R],  sub = DWT(pad(xin)), dsub = DWT(pad(gin))
//   NB == 1:  Rb[c][a][b] += sum gin[y][x] * xin[y+a-R][x+b-R],  sumg[c] += sum gin            (base conv, pixel units)
// One CTA = one channel and a strided share of its (sample, 32 x 32 tile) work items; 128 threads, each a 2 x 4 block of
// positions with ALL k^2 correlation sums of a band in registers (8 k^2 FMAs per 8 + (2 + k - 1)(4 + k - 1) shared-memory words),
// carried across the work items and flushed once per CTA: shuffles, then one global atomic per (band, tap) and CTA.
// (Round 1: one thread per tap row sliding along x - 2 shared-memory words per 5 FMAs - and a shared-memory atomic flush
// per tile: 365 us of the 0.98 ms forward + backward at C = 32, 128^2, B = 64.)
// ---------------------------------------------------------------------------------------------
// raw tiles of the correlation kernel: `x` with the k/2 halo (pixels of the level for NB == 4, the correlation domain itself
// for NB == 1), `g` without; same 16-byte origin / 32-byte row rules as WtRaw
template <typename T, int K, int NB, int TT> struct WtRawX {
  static constexpr int PER = 4 / (int)sizeof(T), APX = 16 / (int)sizeof(T), PXQ = 32 / (int)sizeof(T);
  static constexpr int R = WtTile<K, TT>::R, SW = TT + 2 * R;
  static constexpr int LEFT = NB == 4 ? 2 * R : R;                               // halo pixels left of the tile origin
  static constexpr int XOFF = (APX - LEFT % APX) % APX;
  static constexpr int BH = (NB == 4 ? 2 : 1) * WtTile<K, TT>::SH, BW = ((NB == 4 ? 2 : 1) * SW + XOFF + PXQ - 1) / PXQ * PXQ;
  static constexpr size_t BYTES = ((size_t)BH * BW * sizeof(T) + 127) / 128 * 128;
};
template <typename T, int NB, int TT> struct WtRawG {
  static constexpr int PER = 4 / (int)sizeof(T);
  static constexpr int BH = (NB == 4 ? 2 : 1) * TT, BW = BH;
  static constexpr size_t BYTES = ((size_t)BH * BW * sizeof(T) + 127) / 128 * 128;
};

template <typename TX, typename TG, int K, int NB, bool TMA, int TT>
__global__ void __launch_bounds__(wt_threads(TT), 4)
k_wt_wgrad(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapg, const TX* __restrict__ xin,
           const TG* __restrict__ gin, float* __restrict__ Rout, float* __restrict__ sumg, int C, int Bn, LevelGeom g, int tiles_x,
           int tiles_y) {
  using G = WtTile<K, TT>;
  using RX = WtRawX<TX, K, NB, TT>;
  using RG = WtRawG<TG, NB, TT>;
  constexpr int WT_T = TT, WT_THREADS = wt_threads(TT);
  constexpr int R = G::R, SH = G::SH, SW = WT_T + 2 * R, SP = G::SP, DP = WT_T + 4;
  constexpr int MUL = NB == 4 ? 2 : 1;                   // raw pixels per correlation position along each axis
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  TX* rawx = reinterpret_cast<TX*>(smem_raw);
  TG* rawg = reinterpret_cast<TG*>(smem_raw + (TMA ? RX::BYTES : 0));
  float* smem = reinterpret_cast<float*>(smem_raw + (TMA ? RX::BYTES + RG::BYTES : 0));
  float (*S)[SH][SP] = reinterpret_cast<float (*)[SH][SP]>(smem);                       // [NB][SH][SP]
  float (*Dt)[WT_T][DP] = reinterpret_cast<float (*)[WT_T][DP]>(smem + NB * SH * SP);   // [NB][WT_T][DP]
  __shared__ float red[(WT_THREADS + 31) / 32][NB * K * K + 1];
  __shared__ uint64_t bar;
  const int c = blockIdx.x;
  const int dh = NB == 4 ? g.h2 : g.h, dw = NB == 4 ? g.w2 : g.w;      // extent of the correlation domain
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int py = (tid / (WT_T / 4)) * WT_BH, px = (tid % (WT_T / 4)) * 4;
  float acc[NB][K * K];
#pragma unroll
  for (int band = 0; band < NB; ++band)
#pragma unroll
    for (int q = 0; q < K * K; ++q) acc[band][q] = 0.f;
  float lsum = 0.f;
  const int items = Bn * tiles_y * tiles_x;
  constexpr uint32_t TX_BYTES = (uint32_t)((size_t)RX::BH * RX::BW * sizeof(TX) + (size_t)RG::BH * RG::BW * sizeof(TG));
  auto issue = [&](int it) {
    const int txi = it % tiles_x, tyi = (it / tiles_x) % tiles_y, b = it / (tiles_x * tiles_y);
    const int plane = b * C + c;
    mbar_expect_tx(&bar, TX_BYTES);
    tma::tma_load_3d(smem_u32(rawx), &mapx, (MUL * (txi * WT_T - R) - RX::XOFF) / RX::PER, MUL * (tyi * WT_T - R), plane, &bar);
    tma::tma_load_3d(smem_u32(rawg), &mapg, (MUL * txi * WT_T) / RG::PER, MUL * tyi * WT_T, plane, &bar);
  };
  if (TMA) {
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); tma::tma_prefetch_desc(&mapx); tma::tma_prefetch_desc(&mapg); }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.y < items) issue(blockIdx.y);
  }
  uint32_t phase = 0;
  constexpr int NQ = (SH * SW + WT_THREADS - 1) / WT_THREADS, NQD = WT_T * WT_T / WT_THREADS;
  for (int it = blockIdx.y; it < items; it += gridDim.y) {
    const int txi = it % tiles_x, tyi = (it / tiles_x) % tiles_y, b = it / (tiles_x * tiles_y);
    const long long plane = (long long)b * C + c;
    const int sy0 = tyi * WT_T, sx0 = txi * WT_T;
    if (TMA) {
      if (!mbar_wait(&bar, phase)) __trap();           // a fault in the async pipe must be loud, not a hang
      phase ^= 1;
#pragma unroll 2
      for (int u = 0; u < NQ; ++u) {
        const int i = u * WT_THREADS + tid;
        if (i >= SH * SW) break;
        const int sy = i / SW, sx = i - sy * SW;
        if (NB == 4) {
          float a, bq, cc, d;
          load_quad(rawx, RX::BH, RX::BW, 2 * sy, 2 * sx + RX::XOFF, true, a, bq, cc, d);
          S[0][sy][sx] = 0.5f * (a + bq + cc + d);
          S[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + bq - cc - d);
          S[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - bq + cc - d);
          S[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - bq - cc + d);
        } else {
          S[0][sy][sx] = ldf(rawx + sy * RX::BW + sx + RX::XOFF);
        }
      }
#pragma unroll 2
      for (int u = 0; u < NQD; ++u) {
        const int i = u * WT_THREADS + tid;
        const int sy = i / WT_T, sx = i - sy * WT_T;
        if (NB == 4) {
          float a, bq, cc, d;
          load_quad(rawg, RG::BH, RG::BW, 2 * sy, 2 * sx, true, a, bq, cc, d);
          Dt[0][sy][sx] = 0.5f * (a + bq + cc + d);
          Dt[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + bq - cc - d);
          Dt[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - bq + cc - d);
          Dt[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - bq - cc + d);
        } else {
          const float v = ldf(rawg + sy * RG::BW + sx);
          Dt[0][sy][sx] = v;
          lsum += v;
        }
      }
    } else {
      const TX* xs = xin + plane * (long long)g.h * g.w;
      const TG* gs = gin + plane * (long long)g.h * g.w;
      const bool weven = (g.w & 1) == 0 && ((plane * (long long)g.h * g.w) & 1) == 0;
      {
        float qa[NQ], qb[NQ], qc[NQ], qd[NQ];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          const int i = u * WT_THREADS + tid;
          const int sy = i / SW, sx = i - sy * SW;
          const int gy = sy0 - R + sy, gx = sx0 - R + sx;
          const bool in = i < SH * SW && gy >= 0 && gy < dh && gx >= 0 && gx < dw;
          qa[u] = qb[u] = qc[u] = qd[u] = 0.f;
          if (in) {
            if (NB == 4) load_quad(xs, g.h, g.w, 2 * gy, 2 * gx, weven, qa[u], qb[u], qc[u], qd[u]);
            else qa[u] = ldf(xs + (long long)gy * g.w + gx);
          }
        }
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          const int i = u * WT_THREADS + tid;
          if (i >= SH * SW) continue;
          const int sy = i / SW, sx = i - sy * SW;
          if (NB == 4) {
            const float a = qa[u], bq = qb[u], cc = qc[u], d = qd[u];
            S[0][sy][sx] = 0.5f * (a + bq + cc + d);
            S[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + bq - cc - d);
            S[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - bq + cc - d);
            S[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - bq - cc + d);
          } else {
            S[0][sy][sx] = qa[u];
          }
        }
      }
      {
        float qa[NQD], qb[NQD], qc[NQD], qd[NQD];
#pragma unroll
        for (int u = 0; u < NQD; ++u) {
          const int i = u * WT_THREADS + tid;
          const int sy = i / WT_T, sx = i - sy * WT_T;
          const int gy = sy0 + sy, gx = sx0 + sx;
          qa[u] = qb[u] = qc[u] = qd[u] = 0.f;
          if (gy < dh && gx < dw) {
            if (NB == 4) load_quad(gs, g.h, g.w, 2 * gy, 2 * gx, weven, qa[u], qb[u], qc[u], qd[u]);
            else qa[u] = ldf(gs + (long long)gy * g.w + gx);
          }
        }
#pragma unroll
        for (int u = 0; u < NQD; ++u) {
          const int i = u * WT_THREADS + tid;
          const int sy = i / WT_T, sx = i - sy * WT_T;
          if (NB == 4) {
            const float a = qa[u], bq = qb[u], cc = qc[u], d = qd[u];
            Dt[0][sy][sx] = 0.5f * (a + bq + cc + d);
            Dt[NB > 1 ? 1 : 0][sy][sx] = 0.5f * (a + bq - cc - d);
            Dt[NB > 2 ? 2 : 0][sy][sx] = 0.5f * (a - bq + cc - d);
            Dt[NB > 3 ? 3 : 0][sy][sx] = 0.5f * (a - bq - cc + d);
          } else {
            Dt[0][sy][sx] = qa[u];
            lsum += qa[u];
          }
        }
      }
    }
    __syncthreads();                                    // S / Dt ready; the raw tiles are consumed
    if (TMA && tid == 0 && it + (int)gridDim.y < items) issue(it + gridDim.y);
#pragma unroll
    for (int band = 0; band < NB; ++band) {
      float dv[WT_BH][4];
#pragma unroll
      for (int i = 0; i < WT_BH; ++i) {
        const float4 f = *reinterpret_cast<const float4*>(&Dt[band][py + i][px]);
        dv[i][0] = f.x; dv[i][1] = f.y; dv[i][2] = f.z; dv[i][3] = f.w;
      }
#pragma unroll
      for (int r = 0; r < WT_BH + K - 1; ++r) {
        float v[G::VW];
#pragma unroll
        for (int q = 0; q < G::VW; q += 4) {
          const float4 f = *reinterpret_cast<const float4*>(&S[band][py + r][px + q]);
          v[q] = f.x; v[q + 1] = f.y; v[q + 2] = f.z; v[q + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < WT_BH; ++i) {
          const int a = r - i;
          if (a < 0 || a >= K) continue;
#pragma unroll
          for (int bb = 0; bb < K; ++bb)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[band][a * K + bb] = fmaf(dv[i][j], v[j + bb], acc[band][a * K + bb]);
        }
      }
    }
    __syncthreads();                                    // S / Dt may be overwritten by the next item
  }
  // flush: warp shuffles, the warps through shared memory, one atomic per (band, tap)
#pragma unroll
  for (int band = 0; band < NB; ++band)
#pragma unroll
    for (int q = 0; q < K * K; ++q) {
      const float v = warp_sum(acc[band][q]);
      if (lane == 0) red[warp][band * K * K + q] = v;
    }
  if (NB == 1) {
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp][NB * K * K] = lsum;
  }
  __syncthreads();
  for (int i = tid; i < NB * K * K + 1; i += WT_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < (WT_THREADS + 31) / 32; ++q) v += red[q][i];
    if (i < NB * K * K) { if (v != 0.f) atomicAdd(Rout + (long long)c * NB * K * K + i, v); }
    else if (NB == 1 && sumg) atomicAdd(sumg + c, v);
  }
}
template <typename TX, typename TG, int K, int NB, bool TMA, int TT> static constexpr size_t wt_wgrad_smem() {
  using G = WtTile<K, TT>;
  return 128 + (TMA ? WtRawX<TX, K, NB, TT>::BYTES + WtRawG<TG, NB, TT>::BYTES : 0) +
         sizeof(float) * ((size_t)NB * G::SH * G::SP + (size_t)NB * TT * (TT + 4));
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

// Rank-3 tensor map [planes][h][w] of a contiguous NCHW tensor viewed as planes, box (box_w, box_h, 1), no swizzle, zero fill
// outside the tensor.  Usable when row pitch and plane size are whole 16-byte units (TMA granularity).
template <typename T> static inline bool wt_tma_ok(const T* base, int h, int w) {
  return ((size_t)w * sizeof(T)) % 16 == 0 && ((size_t)w * h * sizeof(T)) % 16 == 0 && ((uintptr_t)base % 16) == 0;
}
template <typename T>
static int wt_make_map(CUtensorMap* map, const T* base, long long planes, int h, int w, int box_w, int box_h) {
  tma::EncodeTiledFn enc = tma::encode_tiled_fn();
  ADN_REQUIRE(enc != nullptr, ADN_ERR_CUDA, "wtconv: cuTensorMapEncodeTiled is not available from this driver");
  // bf16 planes are described as 32-bit words (two pixels per element; w, the box width and the box origin are even)
  const int per = 4 / (int)sizeof(T);
  cuuint64_t dims[3] = {(cuuint64_t)(w / per), (cuuint64_t)h, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)w * sizeof(T), (cuuint64_t)w * h * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)(box_w / per), (cuuint32_t)box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADN_REQUIRE(r == CUDA_SUCCESS, ADN_ERR_CUDA, "wtconv: cuTensorMapEncodeTiled failed (%d) for a %lld x %d x %d tensor, box %d x %d", (int)r,
              planes, h, w, box_w, box_h);
  return ADN_OK;
}
static inline int wt_persistent_grid(long long total, size_t smem, int max_per_sm) {
  long long per_sm = (long long)(227 * 1024) / (long long)(smem + 1024);
  per_sm = per_sm < 1 ? 1 : (per_sm > max_per_sm ? max_per_sm : per_sm);
  const long long cap = per_sm * sm_count();
  return (int)(total < cap ? total : cap);
}

template <typename TI, typename TO, bool BASE, bool FLIP>
static int launch_level(cudaStream_t st, int k, const TI* in, const float* coarse, TO* out, const float* Wl,
                        const float* scale, const float* Wb, const float* bias, const float* bscale, int C,
                        const LevelGeom& g, long long planes) {
  const bool small = g.w2 <= 16 && g.h2 <= 16;
  const int TTv = small ? 16 : 32;
  int tx = cdiv(g.w2, TTv), ty = cdiv(g.h2, TTv);
  long long blocks = planes * tx * ty;
  ADN_REQUIRE(blocks < (1LL << 31) && planes < (1LL << 31), ADN_ERR_SHAPE, "wtconv: too many tiles");
  const bool use_tma = wt_tma_ok(in, g.h, g.w) && !(env().variant & 2);
#define ADN_WT_LAUNCH_T(KK, TMAF, TTC)                                                                           \
  {                                                                                                              \
    constexpr size_t smem = wt_level_smem<TI, KK, BASE, TMAF, TTC>();                                            \
    static bool attr = false;                                                                                    \
    if (!attr) { ADN_CHECK_CUDA(cudaFuncSetAttribute(k_wt_level<TI, TO, KK, BASE, FLIP, TMAF, TTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
    CUtensorMap map;                                                                                             \
    memset(&map, 0, sizeof(map));                                                                                \
    if (TMAF) { int rc = wt_make_map<TI>(&map, in, planes, g.h, g.w, WtRaw<TI, KK, TTC>::BW, WtRaw<TI, KK, TTC>::BH); if (rc) return rc; } \
    ADN_KERNEL("k_wt_level", st);                                                                                \
    k_wt_level<TI, TO, KK, BASE, FLIP, TMAF, TTC><<<wt_persistent_grid(blocks, smem, TTC == 16 ? 16 : 4), wt_threads(TTC), smem, st>>>( \
        map, in, coarse, out, Wl, scale, Wb, bias, bscale, C, g, tx, ty, (int)blocks);                           \
  }
#define ADN_WT_LAUNCH_S(KK, TTC) if (use_tma) ADN_WT_LAUNCH_T(KK, true, TTC) else ADN_WT_LAUNCH_T(KK, false, TTC)
#define ADN_WT_LAUNCH(KK) if (small) { ADN_WT_LAUNCH_S(KK, 16) } else { ADN_WT_LAUNCH_S(KK, 32) }
  switch (k) {
    case 1: ADN_WT_LAUNCH(1); break;
    case 3: ADN_WT_LAUNCH(3); break;
    case 5: ADN_WT_LAUNCH(5); break;
    default: ADN_WT_LAUNCH(7); break;
  }
#undef ADN_WT_LAUNCH
#undef ADN_WT_LAUNCH_S
#undef ADN_WT_LAUNCH_T
  return ADN_OK;
}

template <typename TX, typename TG, int NB>
static int launch_wgrad(cudaStream_t st, int k, const TX* xin, const TG* gin, float* Rout, float* sumg, int C, int Bn,
                        const LevelGeom& g) {
  int dh = NB == 4 ? g.h2 : g.h, dw = NB == 4 ? g.w2 : g.w;
  const bool small = dw <= 16 && dh <= 16;
  const int TTv = small ? 16 : 32;
  int tx = cdiv(dw, TTv), ty = cdiv(dh, TTv);
  const long long items = (long long)Bn * tx * ty, planes = (long long)Bn * C;
  ADN_REQUIRE(items < (1LL << 31) && planes < (1LL << 31), ADN_ERR_SHAPE, "wtconv: too many tiles");
  const bool use_tma = wt_tma_ok(xin, g.h, g.w) && wt_tma_ok(gin, g.h, g.w) && !(env().variant & 2);
#define ADN_WT_LAUNCH_T(KK, TMAF, TTC)                                                                           \
  {                                                                                                              \
    constexpr size_t smem = wt_wgrad_smem<TX, TG, KK, NB, TMAF, TTC>();                                          \
    static bool attr = false;                                                                                    \
    if (!attr) { ADN_CHECK_CUDA(cudaFuncSetAttribute(k_wt_wgrad<TX, TG, KK, NB, TMAF, TTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
    long long per_sm = (long long)(227 * 1024) / (long long)(smem + 1024);                                       \
    const long long cap_sm = TTC == 16 ? 16 : 4;                                                                 \
    per_sm = per_sm < 1 ? 1 : (per_sm > cap_sm ? cap_sm : per_sm);                                               \
    int split = cdiv(per_sm * sm_count(), C);      /* one resident wave of CTAs, each striding over its channel's work items */ \
    split = split < 1 ? 1 : (split > items ? (int)items : split);                                                \
    if (split > 65535) split = 65535;                                                                            \
    dim3 grid(C, split);                                                                                         \
    CUtensorMap mx, mg;                                                                                          \
    memset(&mx, 0, sizeof(mx)); memset(&mg, 0, sizeof(mg));                                                      \
    if (TMAF) {                                                                                                  \
      int rc = wt_make_map<TX>(&mx, xin, planes, g.h, g.w, WtRawX<TX, KK, NB, TTC>::BW, WtRawX<TX, KK, NB, TTC>::BH); \
      if (rc) return rc;                                                                                         \
      rc = wt_make_map<TG>(&mg, gin, planes, g.h, g.w, WtRawG<TG, NB, TTC>::BW, WtRawG<TG, NB, TTC>::BH);        \
      if (rc) return rc;                                                                                         \
    }                                                                                                            \
    ADN_KERNEL("k_wt_wgrad", st);                                                                                \
    k_wt_wgrad<TX, TG, KK, NB, TMAF, TTC><<<grid, wt_threads(TTC), smem, st>>>(mx, mg, xin, gin, Rout, sumg, C, Bn, g, tx, ty); \
  }
#define ADN_WT_LAUNCH_S(KK, TTC) if (use_tma) ADN_WT_LAUNCH_T(KK, true, TTC) else ADN_WT_LAUNCH_T(KK, false, TTC)
#define ADN_WT_LAUNCH(KK) if (small) { ADN_WT_LAUNCH_S(KK, 16) } else { ADN_WT_LAUNCH_S(KK, 32) }
  switch (k) {
    case 1: ADN_WT_LAUNCH(1); break;
    case 3: ADN_WT_LAUNCH(3); break;
    case 5: ADN_WT_LAUNCH(5); break;
    default: ADN_WT_LAUNCH(7); break;
  }
#undef ADN_WT_LAUNCH
#undef ADN_WT_LAUNCH_S
#undef ADN_WT_LAUNCH_T
  return ADN_OK;
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
      int rc = launch_level<T, T, true, FLIP>(st, s.k, src, coarse, dst, w.wavelet_conv_w[0], w.wavelet_scale_w[0], w.base_conv_w,
                                              FLIP ? nullptr : w.base_conv_b, w.base_scale_w, s.C, p.g[0], p.planes);
      if (rc) return rc;
    } else {
      int rc = launch_level<float, float, false, FLIP>(st, s.k, pyr + p.pyr_off[i], coarse, chain + p.pyr_off[i],
                                                       w.wavelet_conv_w[i], w.wavelet_scale_w[i], nullptr, nullptr, nullptr, s.C,
                                                       p.g[i], p.planes);
      if (rc) return rc;
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
  rc = launch_wgrad<T, T, 1>(st, s.k, x, dy, acc.Rb, acc.sumg, s.C, s.B, p.g[0]);
  if (rc) return rc;
  rc = launch_wgrad<T, T, 4>(st, s.k, x, dy, acc.Rl[0], nullptr, s.C, s.B, p.g[0]);
  if (rc) return rc;
  for (int i = 1; i < p.L; ++i) {
    rc = launch_wgrad<float, float, 4>(st, s.k, xpyr + p.pyr_off[i], gpyr + p.pyr_off[i], acc.Rl[i], nullptr, s.C, s.B, p.g[i]);
    if (rc) return rc;
  }
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
