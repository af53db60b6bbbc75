// Conv-as-GEMM kernels of the sm_100a ADN-SSD path ("row" kernels; token grids with W % 128 == 0, d_model 32, d_state 16;
// grids wider than 128 are processed as 128-wide strip images, see struct Strip).
//
// The depthwise 3x3 convolution that follows in_proj (models/ADNssd.py:329-372,388-390) is linear in u:
//     pre[p][c] = sum_t K[c][t] * raw[p + d_t][c] = sum_t  u[p + d_t][:] . Wt[t][c][:],    Wt[t][c][d] = K[c][t] * W_in[c][d]
// so in_proj + conv is ONE K = 9*32 GEMM per 128-token image row whose A operand is the u row shifted by the tap offset
// d_t = (a-1, b-1), t = 3a+b.  The shift costs nothing on tcgen05: the rows of a T8 operand are 16 bytes apart, so a row
// shift is a change of the descriptor start address (tests/test_umma_gpu.py::test_umma_shift).  Image rows live in
// zero-padded "row slots"  [chunk][130 positions][8]  (position 1 + x holds token x; positions 0 and 129 are the zero
// padding of the conv), and a vertical tap selects the slot of the neighbouring image row.  The same identity gives
// the whole backward of in_proj + conv from dpre = dact * SiLU'(pre) with no CUDA-core convolution at all:
//     du[q]      = sum_t dpre[q - d_t][:] . Wt[t]                                       (k_bconv_du)
//     M_t[c][d]  = sum_p dpre[p][c] * u[p + d_t][d]                                     (k_bconv_wg, accumulated in TMEM)
//     dK[c][t]   = sum_d W_in[c][d] * M_t[c][d],      dW_in[c][d] = sum_t K[c][t] * M_t[c][d]
// raw (the in_proj output) is never materialised; only its dt columns are stored (the decay weights need them).
//
// All three kernels are warp-specialised: a producer warp (cp.async / cp.async.bulk into a ring, mbarrier completion),
// one MMA-issuing thread (an elected lane of a converged warp), and epilogue warps that own TMEM lane quarters
// (thread = token row).  k_bconv_du moves the horizontal shift to the output side and k_bconv_wg merges the three
// horizontal taps into one N = 96 operand (see the kernel headers): a 128xNx16 MMA with both operands in shared memory costs
// max(N/2, (4096 + 32 N)/128) cycles, i.e. ~46 cycles for any N <= 64, so few wide MMAs beat many narrow ones.
#pragma once

namespace rowconv {
using namespace adn;
using namespace adn::sm100;

constexpr int RP = 130;                  // positions per chunk of a padded row slot
constexpr int D = 32, DI = 64, GN = 32, CC = 192, NA = 24, DIP = 208, NH = 16;
constexpr int CHB = RP * 16;             // bytes of one padded chunk (2080)
constexpr int USLOT_B = 4 * CHB;         // padded u row slot: 32 channels (8320 bytes)
constexpr int NUS = 6;                   // u ring slots of k_fconv (one copy per row)
constexpr int NUW = 5;                   // u ring slots of k_bconv_wg (three shifted copies per row)
constexpr int WTF_TAP_B = 4 * DIP * 16;  // forward B operand of one tap: [4 chunks][208][8] bf16 (13312 bytes)
constexpr int WTF_B = 9 * WTF_TAP_B;
constexpr int WTB_TG_B = 2048;           // backward B operand of one (tap, 32-channel group): [4 chunks][32][8] bf16
constexpr int WTB_B = 9 * 6 * WTB_TG_B + 1024;   // + W_dt^T image [2 chunks][32][8]

// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tcgen05.mma with a compile-time accumulate flag (the issue loops of the row kernels are fully unrolled: one thread
// issues every MMA of the CTA, so every instruction in its loop counts)
template <bool ACC>
__device__ __forceinline__ void umma_c(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// descriptor + byte offset (the start-address field counts 16-byte units and never carries out of its 14 bits here)
__device__ __forceinline__ uint64_t dadd(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// One tap of the assembled per-channel 3x3 kernel (same rules as assemble_conv_channel).
__device__ __forceinline__ float conv_tap(const ConvWeightPtrs& w, int Di, int cc, int t) {
  if (cc < Di) return w.c2dz[cc * 9 + t];
  const int c = cc - Di;
  if ((c & 1) == 0) return w.c2d[(c >> 1) * 9 + t];
  const int i = c >> 2, nx = Di >> 2;
  const bool first = (c & 3) == 1;
  const float *w31, *w13;
  if (i < nx) {
    w31 = (first ? w.c31x1 : w.c31x2) + i * 3;
    w13 = (first ? w.c13x1 : w.c13x2) + i * 3;
  } else {
    w31 = (first ? w.c31bc1 : w.c31bc2) + (i - nx) * 3;
    w13 = (first ? w.c13bc1 : w.c13bc2) + (i - nx) * 3;
  }
  return w31[t / 3] * w13[t % 3];
}

// Weight images of the conv-as-GEMM kernels (one thread per element, called from k_prep):
//   WtF[t] : K-major B operand [N = 208][K = 32] of the forward GEMM; rows >= 192 (dt) are W_in for the centre tap, else 0
//   WtB[a][g] : K-major B operand [N = 96 (b, d)][K = 32 channels of group g] of the du GEMM (k_bconv_du), then W_dt^T [N = 32][K = 16]
__device__ __forceinline__ void prep_rowconv(const ConvWeightPtrs& cw, const float* __restrict__ win, bf16* __restrict__ wtf,
                                             bf16* __restrict__ wtb, int i) {
  if (i < 9 * DIP * D) {
    const int t = i / (DIP * D), rem = i % (DIP * D);
    const int dchunk = rem / (DIP * 8), n = (rem >> 3) % DIP, d = dchunk * 8 + (rem & 7);
    float v;
    if (n < CC) v = conv_tap(cw, DI, n, t) * win[n * D + d];
    else v = t == 4 ? win[n * D + d] : 0.f;
    wtf[i] = __float2bfloat16_rn(v);
  }
  if (i < 18 * 3072) {        // (a, g) image: [4 chunks of channels][96 rows n = b*32 + d][8]
    const int ag = i / 3072, rem = i % 3072, a = ag / 6, g = ag % 6;
    const int n = (rem >> 3) % 96, b = n >> 5, d = n & 31, c = g * 32 + (rem / 768) * 8 + (rem & 7);
    wtb[i] = __float2bfloat16_rn(conv_tap(cw, DI, c, a * 3 + b) * win[c * D + d]);
  } else if (i < 9 * 6 * 1024 + 512) {
    const int e = i - 9 * 6 * 1024, j = (e >> 8) * 8 + (e & 7), d = (e >> 3) & 31;
    wtb[i] = __float2bfloat16_rn(win[(CC + j) * D + d]);
  }
}

// Token grids wider than 128 (W = 128 * TPR) are processed as B * TPR independent "strip images" of H rows x 128 tokens:
// row index R of the kernels -> strip image R / H (= b * TPR + xt), image row y = R % H -> global 128-token tile
// (b * H + y) * TPR + xt of the TL tensors.  Vertical neighbours are consecutive rows of a strip; the horizontal
// neighbours across a strip edge come from the adjacent tile of the same image row (pad positions of the row slots).
struct Strip {
  int H, TPR;
  __device__ __forceinline__ int tile(int R) const {
    const int img = R / H, y = R - img * H, b = img / TPR, xt = img - b * TPR;
    return (b * H + y) * TPR + xt;
  }
  __device__ __forceinline__ int xt(int R) const { return (R / H) % TPR; }
  __device__ __forceinline__ int sample(int R) const { return (R / H) / TPR; }
};
// (sample, image row, tile) of consecutive strip rows without per-row integer divisions (the epilogue warps are bound by
// instruction issue: four runtime divisions per row and thread cost k_fconv 11 %)
struct StripIter {
  int H, TPR, b, xt, y, tile;
  __device__ __forceinline__ StripIter(Strip sm, int R) : H(sm.H), TPR(sm.TPR) {
    const int img = R / H;
    y = R - img * H;
    b = img / TPR;
    xt = img - b * TPR;
    tile = (b * H + y) * TPR + xt;
  }
  __device__ __forceinline__ void next() {
    ++y;
    tile += TPR;
    if (y == H) {
      y = 0;
      if (++xt == TPR) { xt = 0; ++b; }
      tile = b * H * TPR + xt;
    }
  }
};

// One warp: copy one row-major u tile (128 tokens x 32 channels, 8 KB contiguous) into positions 1..128 of a padded slot and
// its two horizontal neighbour tokens (or zeros at the image border) into the pad positions 0 and 129.
__device__ __forceinline__ void urow_load(uint8_t* slot, const bf16* __restrict__ u, int tile, bool has_left, bool has_right,
                                          int lane) {
  const bf16* urow = u + (long long)tile * 128 * D;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int piece = i * 32 + lane, tok = piece >> 2, ch = piece & 3;
    cp_async16(slot + ch * CHB + (1 + tok) * 16, urow + piece * 8, 16);
  }
  if (lane < 8) {
    const int side = lane >> 2, ch = lane & 3;
    const bool valid = side ? has_right : has_left;
    const bf16* src = valid ? (side ? urow + 128 * D + ch * 8 : urow - D + ch * 8) : urow;
    cp_async16(slot + ch * CHB + (side ? 129 : 0) * 16, src, valid ? 16 : 0);      // src_bytes 0 zero-fills
  }
}

// Producer warp of the u ring: rows [gfirst, glast] -> slot (g - gfirst) % NSLOT; full[] has count 32, empty[] count 1.
// Up to LAG + 1 row loads are in flight (cp.async groups); a row is published once its group has landed.
// u_tl != nullptr: rows [tl_first, tl_last] are also written back to global memory in the TL layout ([tile][4 chunks][128][8]),
// from which the backward kernels fetch them with bulk copies.
template <int NSLOT>
__device__ __forceinline__ void urow_producer(uint8_t* sU, const bf16* __restrict__ u, Strip sm, int gfirst, int glast,
                                              uint64_t* full, uint64_t* empty, int lane, bf16* __restrict__ u_tl = nullptr,
                                              int tl_first = 0, int tl_last = -1) {
  constexpr int LAG = 2;
  const int n = glast - gfirst + 1;
  for (int i = 0; i < n + LAG; ++i) {
    if (i < n) {
      if (i >= NSLOT) mbar_wait(&empty[i % NSLOT], ((i / NSLOT) - 1) & 1);
      const int g = gfirst + i, xt = sm.xt(g);
      urow_load(sU + (i % NSLOT) * USLOT_B, u, sm.tile(g), xt > 0, xt < sm.TPR - 1, lane);
    }
    cp_async_commit();
    if (i >= LAG) {
      cp_async_wait<LAG>();
      fence_async_smem();
      mbar_arrive(&full[(i - LAG) % NSLOT]);
      const int g = gfirst + i - LAG;
      if (u_tl != nullptr && g >= tl_first && g <= tl_last) {
        __syncwarp();     // the row was fetched by all 32 lanes
        const uint8_t* slot = sU + ((i - LAG) % NSLOT) * USLOT_B;
        uint4* dst = reinterpret_cast<uint4*>(u_tl + (long long)sm.tile(g) * 128 * D);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int piece = k * 32 + lane;
          dst[piece] = *reinterpret_cast<const uint4*>(slot + (piece >> 7) * CHB + (1 + (piece & 127)) * 16);
        }
      }
    }
  }
}

// Three-copy u slot of k_bconv_wg: 12 unpadded chunks [copy b][chunk dc][128 rows][8]; copy b holds tokens b-1 .. b+126 of
// the tile, so that chunk (b, dc) of ONE 12-chunk operand (chunk stride 2048) is chunk dc shifted by b-1 tokens: the three
// horizontal taps are one operand with N = 96.  Row 0 of copy 0 / row 127 of copy 2 are the neighbour tokens of the
// adjacent tiles of the same image row, or zeros at the image border (16-byte bulk copies from `zeros16`).
constexpr int USLOT3_B = 12 * 2048;

// Elected-thread producer of the three-copy u ring from the TL copy of u (20 bulk copies per tile).
template <int NSLOT>
__device__ __forceinline__ bool urow3_bulk_producer(uint8_t* sU, const bf16* __restrict__ u_tl, const void* zeros16, Strip sm,
                                                    int gfirst, int glast, uint64_t* full, uint64_t* empty) {
  bool ok = true;
  const int n = glast - gfirst + 1;
  for (int i = 0; i < n; ++i) {
    const int sl = i % NSLOT;
    if (i >= NSLOT) ok = mbar_wait(&empty[sl], ((i / NSLOT) - 1) & 1) && ok;
    mbar_expect_tx(&full[sl], 3 * 4 * 2048);
    const int g = gfirst + i, xt = sm.xt(g);
    const bf16* src = u_tl + (long long)sm.tile(g) * 128 * D;
    uint8_t* slot = sU + sl * USLOT3_B;
#pragma unroll
    for (int dc = 0; dc < 4; ++dc) {
      const bf16* ch = src + dc * 1024;
      bulk_g2s(slot + (0 * 4 + dc) * 2048 + 16, ch, 2032, &full[sl]);                                   // tokens 0..126 -> rows 1..127
      bulk_g2s(slot + (0 * 4 + dc) * 2048, xt > 0 ? (const void*)(ch - 128 * D + 127 * 8) : zeros16, 16, &full[sl]);
      bulk_g2s(slot + (1 * 4 + dc) * 2048, ch, 2048, &full[sl]);                                        // tokens 0..127
      bulk_g2s(slot + (2 * 4 + dc) * 2048, ch + 8, 2032, &full[sl]);                                    // tokens 1..127 -> rows 0..126
      bulk_g2s(slot + (2 * 4 + dc) * 2048 + 2032, xt < sm.TPR - 1 ? (const void*)(ch + 128 * D) : zeros16, 16, &full[sl]);
    }
  }
  return ok;
}

// SiLU and SiLU' of 32 accumulator columns of this thread's token row (two TMEM loads in flight, packed fp32x2 math);
// stores four chunks of act (and of sgrad) in the TL layout.  MODE 0: nothing else; 1: also copy the bf16 act chunks to
// shared memory (Bc operand of the state MMA); 2: write w * act (decay-weighted x; w8 = the eight head weights of these
// four chunks) to shared memory.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__device__ __forceinline__ void silu32(uint32_t taddr, bf16* a_dst, bf16* g_dst, uint8_t* s_dst, const float* w8) {
  float v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_ld16(taddr + 16, v[1]);
  tmem_wait_ld();
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    float2 a[4], gq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = make_float2(v[hh >> 1][(hh & 1) * 8 + 2 * j], v[hh >> 1][(hh & 1) * 8 + 2 * j + 1]);
      // sigmoid = 1 / (1 + 2^(-x log2 e)): two MUFU ops per element.  (A MUFU-free Newton reciprocal was measured slower:
      // the epilogue is bound by instruction issue / latency of 3 warps per scheduler, not by the MUFU pipe.)
      const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
      const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
      const float2 sg = make_float2(rcp_approx(d.x), rcp_approx(d.y));
      a[j] = __fmul2_rn(x, sg);
      const float2 om = __ffma2_rn(sg, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
      gq[j] = __ffma2_rn(a[j], om, sg);
    }
    const uint4 pa = pack8_f2(a);
    *reinterpret_cast<uint4*>(a_dst + hh * 1024) = pa;
    if (g_dst) *reinterpret_cast<uint4*>(g_dst + hh * 1024) = pack8_f2(gq);
    if (MODE == 1) *reinterpret_cast<uint4*>(s_dst + hh * 2048) = pa;
    if (MODE == 2) {
      float2 r[4];
      unpack8_f2(pa, r);
      const float2 w2 = make_float2(w8[hh * 2], w8[hh * 2 + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = __fmul2_rn(r[j], w2);
      *reinterpret_cast<uint4*>(s_dst + hh * 2048) = pack8_f2(r);
    }
  }
}

// One epilogue column group of k_fconv: 48 accumulator columns [48 G, 48 G + 48) = chunks [6 G, 6 G + 6) of [z | x | B | C]
// (three TMEM loads in flight, packed fp32x2 math).  Per chunk: SiLU -> act, SiLU' -> sgrad (TL layout); x chunks also write
// w * act (decay-weighted x, heads 2(c-8), 2(c-8)+1) and B chunks a copy of act to the state operands in shared memory.
//   G 0: z0 z1 z2 | G 1: z3 x0 x1 (heads 0..7) | G 2: x2 x3 B0 (heads 8..15) | G 3: B1 C0 C1 (+ stores the dt columns)
template <int G>
__device__ __forceinline__ void fconv_epi_group(uint32_t ta, bf16* arow, bf16* grow, uint8_t* st, bf16* dtrow,
                                                const float* s_bias, const float* s_eA) {
  float v[3][16];
#pragma unroll
  for (int i = 0; i < 3; ++i) tmem_ld16(ta + 48 * G + 16 * i, v[i]);
  float w[8];
  if constexpr (G == 1 || G == 2) {
    float d8[8];
    tmem_ld8(ta + CC + (G == 1 ? 0 : 8), d8);
    tmem_wait_ld();
    const uint4 p = pack8(d8);      // the decay weights are formed from the stored (bf16) dt so that backward sees the same w
    unpack8(p, d8);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = softplus_fast(d8[j] + s_bias[(G == 1 ? 0 : 8) + j]) * s_eA[(G == 1 ? 0 : 8) + j];
  } else if constexpr (G == 3) {
    float d16[16];
    tmem_ld16(ta + CC, d16);
    tmem_wait_ld();
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = d16[j]; hi[j] = d16[8 + j]; }
    *reinterpret_cast<uint4*>(dtrow) = pack8(lo);
    *reinterpret_cast<uint4*>(dtrow + 1024) = pack8(hi);
  } else {
    tmem_wait_ld();
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    constexpr int c0 = 6 * G;
    const int c = c0 + i;                      // compile-time after unrolling
    float2 a[4], gq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = make_float2(v[i >> 1][(i & 1) * 8 + 2 * j], v[i >> 1][(i & 1) * 8 + 2 * j + 1]);
      // sigmoid(x) = 0.5 + 0.5 tanh(x / 2): ONE MUFU op per element.  The four epilogue warps of a scheduler share a 4-lane
      // MUFU pipe (8 cycles per warp instruction), which bounds this loop; tanh.approx (relative error 2^-11, i.e. <= 2.5e-4
      // absolute on the sigmoid) is well inside the rounding of the bf16 tensors these values are stored in.
      const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
      const float2 th = make_float2(tanh_approx(hx.x), tanh_approx(hx.y));
      const float2 sg = __ffma2_rn(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
      a[j] = __fmul2_rn(x, sg);
      const float2 om = __ffma2_rn(th, make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));
      gq[j] = __ffma2_rn(a[j], om, sg);
    }
    const uint4 pa = pack8_f2(a);
    *reinterpret_cast<uint4*>(arow + c * 1024) = pa;
    if (grow) *reinterpret_cast<uint4*>(grow + c * 1024) = pack8_f2(gq);
    if (c >= 8 && c < 16) {                    // x chunk: decay-weighted copy for the state MMA (A operand chunk c - 8)
      float2 r[4];
      unpack8_f2(pa, r);
      const int hl = (2 * (c - 8)) & 7;        // local index of the chunk's first head in w[8]
      const float2 w2 = make_float2(w[hl], w[hl + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = __fmul2_rn(r[j], w2);
      *reinterpret_cast<uint4*>(st + (c - 8) * 2048) = pack8_f2(r);
    } else if (c >= 16 && c < 20) {            // B chunk: B operand chunk 8 + (c - 16)
      *reinterpret_cast<uint4*>(st + (8 + c - 16) * 2048) = pa;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fconv: forward in_proj + depthwise 3x3 + SiLU (+ SiLU') + decay weights + state accumulation, one image row
// (128 tokens) per step.  Stages (1)-(4a) of the mixer (models/ADNssd.py:309-390, :267-280).
//   warps 0-15 epilogue (lane quarter q = warp & 3, column group = warp >> 2: 48 of the 192 conv columns each)
//   warp  16   u-row producer (cp.async into the padded ring)
//   warp  17   one elected lane issues every tcgen05.mma (and the bulk copies of the weight image)
// TMEM: two 208-column accumulators [z | x | B | C | dt] + 32 columns of the per-sample state S'[c][j].
// ------------------------------------------------------------------------------------------------
constexpr int FC_ST_B = 12 * 2048;   // state operands of one row: wx chunks 0..7, Bc chunks 8..11
constexpr int FC_SMEM = 2 * FC_ST_B + NUS * USLOT_B + WTF_B;
constexpr int FC_COL_S = 2 * DIP;
constexpr int FC_EPI_WARPS = 16, FC_THREADS = (FC_EPI_WARPS + 2) * 32;   // 4 column groups x 4 lane quarters + producer + MMA

__global__ void __launch_bounds__(FC_THREADS, 1)
k_fconv(const bf16* __restrict__ u, const bf16* __restrict__ WtF, const float* __restrict__ dt_bias,
        const float* __restrict__ A_log, bf16* __restrict__ act, bf16* __restrict__ sgrad, bf16* __restrict__ dtraw,
        float* __restrict__ S, int H, int rows_total, int rows_per_cta, int* __restrict__ status, bf16* __restrict__ u_tl,
        int TPR) {
  ADN_CTA_STAMP(0, 0);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[NUS], empty[NUS], acc_full[2], acc_empty[2], st_full[2], st_empty[2], s_done, s_free, w_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[NH], s_eA[NH];
  uint8_t* sSt = smem;
  uint8_t* sU = smem + 2 * FC_ST_B;
  uint8_t* sW = sU + NUS * USLOT_B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R0 = blockIdx.x * rows_per_cta, R1 = min(rows_total, R0 + rows_per_cta);
  for (int i = tid; i < (2 * FC_ST_B + NUS * USLOT_B) / 16; i += FC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < NH) { s_bias[tid] = dt_bias[tid]; s_eA[tid] = __expf(A_log[tid]); }
  if (tid == 0) {
    for (int i = 0; i < NUS; ++i) { mbar_init(&full[i], 32); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], FC_EPI_WARPS); mbar_init(&st_full[i], FC_EPI_WARPS); mbar_init(&st_empty[i], 1); }
    mbar_init(&s_done, 1);
    mbar_init(&s_free, FC_EPI_WARPS);
    mbar_init(&w_full, 1);
    fence_mbar_init();
  }
  if (warp == FC_EPI_WARPS + 1) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  ADN_CTA_STAMP(0, 2);
  const int gfirst = max(R0 - 1, 0), glast = min(R1, rows_total - 1);
  bool ok = true;
  if (R0 < R1) {
    if (warp == FC_EPI_WARPS) {
      urow_producer<NUS>(sU, u, Strip{H, TPR}, gfirst, glast, full, empty, lane, u_tl, R0, R1 - 1);
    } else if (warp == FC_EPI_WARPS + 1) {
      if (elect_one()) {   // one elected lane of the converged warp: tcgen05.mma is emitted without a lane-serialising loop
        // the weight image arrives by nine bulk copies while the first u rows are in flight
        mbar_expect_tx(&w_full, WTF_B);
        for (int t = 0; t < 9; ++t) bulk_g2s(sW + t * WTF_TAP_B, reinterpret_cast<const uint8_t*>(WtF) + t * WTF_TAP_B, WTF_TAP_B, &w_full);
        const uint32_t idesc = make_idesc_rt(128, DIP, false, false), idesc_s = make_idesc_rt(128, GN, true, true);
        const uint32_t ubase = smem_u32(sU), wbase = smem_u32(sW), stbase = smem_u32(sSt);
        const uint64_t dU0 = make_desc(ubase, CHB, 128), dW0 = make_desc(wbase, DIP * 16, 128);
        int next_wait = gfirst, fl = 0;
        PhaseTimer pt(4, true);
        ok = mbar_wait(&w_full, 0) && ok;
        for (int R = R0; R <= R1; ++R) {
          pt.mark(7);
          if (R < R1) {
            const int it = R - R0, acc = it & 1, y = R % H;
            const int need = min(R + 1, glast);
            while (next_wait <= need) {
              const int i = next_wait - gfirst;
              ok = mbar_wait(&full[i % NUS], (i / NUS) & 1) && ok;
              ++next_wait;
            }
            pt.mark(0);
            if (it >= 2) ok = mbar_wait(&acc_empty[acc], ((it >> 1) - 1) & 1) && ok;
            tc_fence_after();
            pt.mark(1);
            const uint32_t tacc = tbase + acc * DIP;
            if (y > 0 && y < H - 1) {      // interior row: all nine taps, fully unrolled, descriptors = base + constant
              uint64_t dA[3];
#pragma unroll
              for (int a = 0; a < 3; ++a) dA[a] = dadd(dU0, (uint32_t)(((R + a - 1) - gfirst) % NUS) * USLOT_B);
#pragma unroll
              for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t da = dadd(dA[a], b * 16 + ks * 2 * CHB);
                    const uint64_t db = dadd(dW0, (a * 3 + b) * WTF_TAP_B + ks * 2 * DIP * 16);
                    if (a == 0 && b == 0 && ks == 0) umma_c<false>(tacc, da, db, idesc);
                    else umma_c<true>(tacc, da, db, idesc);
                  }
            } else {
              bool first = true;
              for (int a = 0; a < 3; ++a) {
                if (y + a - 1 < 0 || y + a - 1 >= H) continue;
                const uint32_t ab = ubase + (uint32_t)(((R + a - 1) - gfirst) % NUS) * USLOT_B;
                for (int b = 0; b < 3; ++b) {
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t da = make_desc(ab + b * 16 + ks * 2 * CHB, CHB, 128);
                    const uint64_t db = make_desc(wbase + (a * 3 + b) * WTF_TAP_B + ks * 2 * DIP * 16, DIP * 16, 128);
                    umma(tacc, da, db, idesc, !first);
                    first = false;
                  }
                }
              }
            }
            umma_commit(&acc_full[acc]);
            if (R - 1 >= gfirst) umma_commit(&empty[((R - 1) - gfirst) % NUS]);
            pt.mark(2);
          }
          if (R > R0) {   // state accumulation of the previous row (its epilogue produced w*x and Bc)
            const int Rp = R - 1, it = Rp - R0, sb = it & 1, yp = Rp % H;
            ok = mbar_wait(&st_full[sb], (it >> 1) & 1) && ok;
            tc_fence_after();
            pt.mark(3);
            const bool first_of_sample = (Rp == R0) || (yp == 0);
            if (first_of_sample && fl > 0) { ok = mbar_wait(&s_free, (fl - 1) & 1) && ok; tc_fence_after(); }
            const uint32_t a0 = stbase + sb * FC_ST_B, b0 = a0 + 8 * 2048;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma(tbase + FC_COL_S, make_desc(a0 + k * 256, 128, 2048), make_desc(b0 + k * 256, 128, 2048), idesc_s,
                   !(first_of_sample && k == 0));
            umma_commit(&st_empty[sb]);
            if ((Rp == R1 - 1) || (yp == H - 1)) { umma_commit(&s_done); ++fl; }
            pt.mark(4);
          }
        }
        ADN_CTA_STAMP_ANY(0, 3);
        if (!ok) atomicExch(status, 20);
      }
    } else {
      const int q = warp & 3, grp = warp >> 2, row = q * 32 + lane;   // column group: 48 columns each (fconv_epi_group)
      int fl = 0;
      PhaseTimer pt(7, tid == 0 || tid == 128 || tid == 256);
      StripIter si(Strip{H, TPR}, R0);
      for (int R = R0; R < R1; ++R, si.next()) {
        const int it = R - R0, acc = it & 1, y = si.y, b = si.b;
        const long long tile = si.tile;
        pt.mark(7);
        ok = mbar_wait(&acc_full[acc], (it >> 1) & 1) && ok;
        pt.mark(0);
        if (it >= 2) ok = mbar_wait(&st_empty[acc], ((it >> 1) - 1) & 1) && ok;
        tc_fence_after();
        pt.mark(1);
        const uint32_t ta = tbase + ((uint32_t)(q * 32) << 16) + acc * DIP;
        bf16* arow = act + (tile * NA * 128 + row) * 8;
        bf16* grow = sgrad ? sgrad + (tile * NA * 128 + row) * 8 : nullptr;
        uint8_t* st = sSt + acc * FC_ST_B + row * 16;
        bf16* dtrow = dtraw + (tile * 2 * 128 + row) * 8;
        if (grp == 0) fconv_epi_group<0>(ta, arow, grow, st, dtrow, s_bias, s_eA);
        else if (grp == 1) fconv_epi_group<1>(ta, arow, grow, st, dtrow, s_bias, s_eA);
        else if (grp == 2) fconv_epi_group<2>(ta, arow, grow, st, dtrow, s_bias, s_eA);
        else fconv_epi_group<3>(ta, arow, grow, st, dtrow, s_bias, s_eA);
        pt.mark(2 + (grp == 1));
        tc_fence_before();
        if (grp != 0) fence_async_smem();      // group 0 (z columns only) wrote no operand of the state MMA
        __syncwarp();
        if (lane == 0) { mbar_arrive(&acc_empty[acc]); mbar_arrive(&st_full[acc]); }
        pt.mark(4);
        if ((R == R1 - 1) || (y == H - 1)) {   // flush the state of sample b accumulated by this CTA
          ok = mbar_wait(&s_done, fl & 1) && ok;
          ++fl;
          tc_fence_after();
          if (q < 2 && grp < 2 && ok) {
            const int half = grp;
            float v[16];
            tmem_ld16(tbase + ((uint32_t)(q * 32) << 16) + FC_COL_S + half * 16, v);
            tmem_wait_ld();
            const int c = q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int jj = half * 16 + j;
              if (((jj ^ c) & 1) == 0) atomicAdd(S + ((long long)b * GN + jj) * DI + c, v[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free);
        }
      }
      if (!ok && lane == 0) atomicExch(status, 21);
    }
  }
  tc_fence_before();
  __syncthreads();
  ADN_CTA_STAMP(0, 1);
  if (warp == FC_EPI_WARPS + 1) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// k_bconv_du: du[q] = sum_t dpre[q - d_t] . Wt[t] + ddt[q] . W_dt     (backward-data of in_proj o conv).
// The horizontal shift is moved to the OUTPUT side so that every MMA has N = 96 and unshifted, unpadded operands:
//     Z_b[y][x'][d] = sum_a sum_c dpre[y - (a-1)][x'][c] * Wt[a,b][c][d]        (N = 96 = (b, d), K = 3 rows x 192 channels)
//     du[y][x]      = Z_0[y][x+1] + Z_1[y][x] + Z_2[y][x-1]                     (zero outside the row)
// (a 128x32x16 MMA costs the same ~46 cycles of shared-memory operand traffic as a 128x96x16 one, profiles/umma_issue_rate.py).
// The CTA streams the dpre rows of its range ONCE, in (row, 32-channel group) units of one 8 KB bulk copy; input row r feeds
// the accumulators of output rows r-1, r, r+1, which live in a ring of four 96-column TMEM slots.  The epilogue adds the
// three shifted pieces with warp shuffles (+ a 256-byte exchange between neighbouring warps) and stores du row-major.
//   warps 0-3 epilogue, warp 4: elected bulk-copy producer, warp 5: elected MMA issuer.
// ------------------------------------------------------------------------------------------------
constexpr int DU_NST = 8;
constexpr int DU_STG_B = 8192 + 4096;             // one (row, group) tile [4 chunks][128][8] + the row's ddt tile (group 0 only)
constexpr int WTZ_AG_B = 4 * 96 * 16;             // B operand of one (vertical tap a, group g): [4 chunks][96 (b,d)][8] bf16
constexpr int DU_SMEM = DU_NST * DU_STG_B + WTB_B + (2 * 4 + 1) * 64 * 4;   // + exchange rows [parity][warp][64] and one zero row
static_assert(18 * WTZ_AG_B + 1024 == WTB_B, "weight image size");

__global__ void __launch_bounds__(192, 1)
k_bconv_du(const bf16* __restrict__ dpre, const bf16* __restrict__ ddt, const bf16* __restrict__ WtZ, bf16* __restrict__ du,
           int H, int rows_total, int rows_per_cta, int* __restrict__ status, int dbg, int TPR) {
  // TPR > 1 (strip images, see Strip): the pieces Z_0[x+1] / Z_2[x-1] that cross a strip edge are added by k_bconv_du_edge
  // dbg (diagnostics, ADN_DU_DBG): 1 = no loads, 2 = no MMAs, 4 = no du stores; results are then meaningless
  ADN_CTA_STAMP(1, 0);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[DU_NST], empty[DU_NST], slot_full[4], slot_empty[4], w_full;
  __shared__ uint32_t tmem_slot;
  uint8_t* sStg = smem;
  uint8_t* sW = smem + DU_NST * DU_STG_B;
  float* sX = reinterpret_cast<float*>(sW + WTB_B);      // [parity][warp][0: lane 0's Z_0, 1: lane 31's Z_2][32 d]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R0 = blockIdx.x * rows_per_cta, R1 = min(rows_total, R0 + rows_per_cta);
  if (tid < 64) sX[8 * 64 + tid] = 0.f;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < DU_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  ADN_CTA_STAMP(1, 2);
  bool ok = true;
  if (R0 < R1) {
    // input rows: the range plus one halo row on each side when that row belongs to the same sample
    const int rin0 = (R0 % H) > 0 ? R0 - 1 : R0, rin1 = ((R1 - 1) % H) < H - 1 ? R1 : R1 - 1;
    if (warp == 4) {
      if (elect_one()) {
        int un = 0;
        PhaseTimer pp(2, true);
        for (int r = rin0; r <= rin1; ++r) {
          for (int g = 0; g < 6; ++g, ++un) {
            const int stg = un % DU_NST;
            pp.mark(1);
            if (un >= DU_NST) ok = mbar_wait(&empty[stg], ((un / DU_NST) - 1) & 1) && ok;
            pp.mark(0);
            uint8_t* sb = sStg + stg * DU_STG_B;
            const bool dt = g == 0 && r >= R0 && r < R1;
            if (dbg & 1) { mbar_arrive(&full[stg]); continue; }
            mbar_expect_tx(&full[stg], 8192u + (dt ? 4096u : 0u));
            const long long tile = Strip{H, TPR}.tile(r);
            bulk_g2s(sb, dpre + (tile * NA + g * 4) * 1024, 8192, &full[stg]);
            if (dt) bulk_g2s(sb + 8192, ddt + tile * 2 * 1024, 4096, &full[stg]);
          }
        }
        if (!ok) atomicExch(status, 22);
      }
    } else if (warp == 5) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_rt(128, 3 * D, false, false), idesc_dt = make_idesc_rt(128, D, false, false);
        const uint32_t sbase = smem_u32(sStg), wbase = smem_u32(sW);
        const uint64_t dS0 = make_desc(sbase, 2048, 128), dW0 = make_desc(wbase, 96 * 16, 128);
        const uint64_t dWdt = make_desc(wbase + 18 * WTZ_AG_B, 512, 128);
        PhaseTimer pt(5, true);
        // the weight image arrives by bulk copies while the first dpre tiles are in flight
        mbar_expect_tx(&w_full, WTB_B);
        for (int i = 0; i < 18; ++i) bulk_g2s(sW + i * WTZ_AG_B, reinterpret_cast<const uint8_t*>(WtZ) + i * WTZ_AG_B, WTZ_AG_B, &w_full);
        bulk_g2s(sW + 18 * WTZ_AG_B, reinterpret_cast<const uint8_t*>(WtZ) + 18 * WTZ_AG_B, 1024, &w_full);
        ok = mbar_wait(&w_full, 0) && ok;
        int un = 0;
        for (int r = rin0; r <= rin1; ++r) {
          const int y = r % H;
          for (int g = 0; g < 6; ++g, ++un) {
            const int stg = un % DU_NST;
            pt.mark(7);
            ok = mbar_wait(&full[stg], (un / DU_NST) & 1) && ok;
            pt.mark(0);
            const uint64_t dA = dadd(dS0, stg * DU_STG_B);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const int rout = r + a - 1, yout = y + a - 1;
              if (yout < 0 || yout >= H || rout < R0 || rout >= R1) continue;
              const int k = rout - R0;
              const uint32_t tacc = tbase + (k & 3) * 96;
              // the first contribution to an output row comes from the input row above it (a == 2), or from the row itself at
              // the top edge of the image
              const bool init = g == 0 && r == (yout > 0 ? rout - 1 : rout);
              if (init && k >= 4) { ok = mbar_wait(&slot_empty[k & 3], ((k >> 2) - 1) & 1) && ok; }
              tc_fence_after();
              const uint64_t dB = dadd(dW0, (a * 6 + g) * WTZ_AG_B);
              if (dbg & 2) continue;
              umma(tacc, dA, dB, idesc, !init);
              umma_c<true>(tacc, dadd(dA, 2 * 2048), dadd(dB, 2 * 96 * 16), idesc);
            }
            if (g == 0 && r >= R0 && r < R1 && !(dbg & 2))      // ddt . W_dt lands in the unshifted piece Z_1 of the row's own accumulator
              umma_c<true>(tbase + ((r - R0) & 3) * 96 + D, dadd(dA, 8192), dWdt, idesc_dt);
            umma_commit(&empty[stg]);
            pt.mark(2);
          }
          if (y > 0 && r - 1 >= R0) umma_commit(&slot_full[(r - 1 - R0) & 3]);
          if (y == H - 1 && r < R1) umma_commit(&slot_full[(r - R0) & 3]);
        }
        ADN_CTA_STAMP_ANY(1, 3);
        if (!ok) atomicExch(status, 23);
      }
    } else {
      const int x = warp * 32 + lane;
      PhaseTimer pe(3, tid == 0);
      for (int r = R0; r < R1; ++r) {
        const int k = r - R0, par = k & 1;
        pe.mark(7);
        ok = mbar_wait(&slot_full[k & 3], (k >> 2) & 1) && ok;
        tc_fence_after();
        pe.mark(0);
        const uint32_t ta = tbase + ((uint32_t)(warp * 32) << 16) + (k & 3) * 96;
        float z[6][16];
#pragma unroll
        for (int i = 0; i < 6; ++i) tmem_ld16(ta + i * 16, z[i]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[k & 3]);
        // neighbour exchange across warp boundaries, branch-free: lane 0 publishes its Z_0 row, lane 31 its Z_2 row (vector
        // stores); after the barrier EVERY lane loads the two candidate rows (broadcast reads) and keeps them only where the
        // shuffle has no source lane.  Warp 3 / warp 0 read the permanently-zero row instead (image border).
        float* xw = sX + (par * 4 + warp) * 64;
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            *reinterpret_cast<float4*>(xw + j) = make_float4(z[0][j], z[0][j + 1], z[0][j + 2], z[0][j + 3]);
            *reinterpret_cast<float4*>(xw + 16 + j) = make_float4(z[1][j], z[1][j + 1], z[1][j + 2], z[1][j + 3]);
          }
        }
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            *reinterpret_cast<float4*>(xw + 32 + j) = make_float4(z[4][j], z[4][j + 1], z[4][j + 2], z[4][j + 3]);
            *reinterpret_cast<float4*>(xw + 48 + j) = make_float4(z[5][j], z[5][j + 1], z[5][j + 2], z[5][j + 3]);
          }
        }
        pe.mark(1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        pe.mark(2);
        const float* xn = warp < 3 ? sX + (par * 4 + warp + 1) * 64 : sX + 8 * 64;        // next warp's lane 0: Z_0[x + 1]
        const float* xp = warp > 0 ? sX + (par * 4 + warp - 1) * 64 + 32 : sX + 8 * 64;   // previous warp's lane 31: Z_2[x - 1]
        const bool hi = lane == 31, lo = lane == 0;
        float o[32];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 e0 = *reinterpret_cast<const float4*>(xn + h * 16 + j4);
            const float4 e2 = *reinterpret_cast<const float4*>(xp + h * 16 + j4);
            const float e0v[4] = {e0.x, e0.y, e0.z, e0.w}, e2v[4] = {e2.x, e2.y, e2.z, e2.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = j4 + q;
              const float s0 = __shfl_down_sync(0xffffffffu, z[h][j], 1);
              const float s2 = __shfl_up_sync(0xffffffffu, z[4 + h][j], 1);
              o[h * 16 + j] = z[2 + h][j] + (hi ? e0v[q] : s0) + (lo ? e2v[q] : s2);
            }
          }
        if (ok && !(dbg & 4)) {
          bf16* dst = du + ((long long)Strip{H, TPR}.tile(r) * 128 + x) * D;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v8[j] = o[q * 8 + j];
            *reinterpret_cast<uint4*>(dst + q * 8) = pack8(v8);
          }
        }
        pe.mark(3);
      }
      if (!ok && lane == 0) atomicExch(status, 24);
    }
  }
  tc_fence_before();
  __syncthreads();
  ADN_CTA_STAMP(1, 1);
  if (warp == 5) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// k_bconv_du_edge (W > 128 only): k_bconv_du treats every 128-token strip as its own image, so the two tokens on either
// side of an interior strip edge miss the conv taps that reach across it:
//   du[y][127 of strip e]   += sum_a dpre[y-(a-1)][0 of strip e+1][:]   . Wt[a, b=0]
//   du[y][0   of strip e+1] += sum_a dpre[y-(a-1)][127 of strip e][:]   . Wt[a, b=2]
// One warp per token (lane = output column d), fp32 weights K[c][t] * W_in[c][d]; runs after k_bconv_du on the same stream.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bconv_du_edge(const bf16* __restrict__ dpre, const float* __restrict__ Kc, const float* __restrict__ Win, bf16* __restrict__ du,
                int H, int TPR, int n_edges /* B * H * (TPR - 1) */) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= 2 * n_edges) return;
  const int side = wid & 1, e = wid >> 1;                  // side 0: last token of the left strip, 1: first token of the right strip
  const int xe = e % (TPR - 1), by = e / (TPR - 1), y = by % H, b = by / H;
  const int tile_l = (b * H + y) * TPR + xe, tile_r = tile_l + 1;
  const int tile_dst = side ? tile_r : tile_l, x_dst = side ? 0 : 127;
  const int x_src = side ? 127 : 0, bt = side ? 2 : 0;     // source token inside the neighbouring tile, horizontal tap index
  float acc = 0.f;
  for (int a = 0; a < 3; ++a) {
    const int ys = y - (a - 1);
    if (ys < 0 || ys >= H) continue;
    const long long tile_src = (long long)(b * H + ys) * TPR + xe + (side ? 0 : 1);
    const bf16* row = dpre + tile_src * NA * 1024 + x_src * 8;
    for (int c = 0; c < CC; ++c) {
      const float v = __bfloat162float(row[(c >> 3) * 1024 + (c & 7)]);
      acc = fmaf(v * __ldg(Kc + c * 9 + 3 * a + bt), __ldg(Win + c * D + lane), acc);
    }
  }
  bf16* dst = du + ((long long)tile_dst * 128 + x_dst) * D + lane;
  *dst = __float2bfloat16_rn(__bfloat162float(*dst) + acc);
}

// ------------------------------------------------------------------------------------------------
// k_bconv_wg: M_t[c][d] = sum_p dpre[p][c] * u[p + d_t][d] for the 128-channel block mb of this CTA, nine 32-column TMEM
// accumulators kept for the CTA's whole row range; the epilogue contracts them into per-CTA slabs of dK and dW_in partial
// sums (added up by k_finalize_fast).   mb 0: channels 0..127 (z, x).   mb 1: channels 128..191 (B, C) + the 16 dt rows (their
// centre tap is dW_in[192 + j]) + 48 idle rows.
//   warps 0-3 epilogue (warp 0 lane 0 is also the dpre-tile producer), warp 4 u-row producer, warp 5 lane 0 MMA issue.
// ------------------------------------------------------------------------------------------------
constexpr int WG_NST = 3;
constexpr int WG_STG_B = 16 * 2048;
constexpr int WG_SMEM = WG_NST * WG_STG_B + NUW * USLOT3_B;
static_assert(WG_SMEM <= 227 * 1024, "k_bconv_wg shared memory");

__global__ void __launch_bounds__(192, 1)
k_bconv_wg(const bf16* __restrict__ dpre, const bf16* __restrict__ ddt, const bf16* __restrict__ u /* TL copy of u */,
           const float* __restrict__ Win, const float* __restrict__ Kc, float* __restrict__ dK, float* __restrict__ dWin,
           int H, int rows_total, int rows_per_cta, int ctas_per_block, int* __restrict__ status, int TPR,
           const void* __restrict__ zeros16) {
  ADN_CTA_STAMP(2, 0);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[NUW], empty[NUW], a_full[WG_NST], a_empty[WG_NST], done;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t s_tapmask;
  uint8_t* sA = smem;
  uint8_t* sU = smem + WG_NST * WG_STG_B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.x / ctas_per_block, part = blockIdx.x % ctas_per_block;
  const int R0 = part * rows_per_cta, R1 = min(rows_total, R0 + rows_per_cta);
  PhaseTimer ptc(6, tid == 64);
  for (int i = tid; i < WG_SMEM / 16; i += 192) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int i = 0; i < NUW; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < WG_NST; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(&done, 2);   // MMA completion (tcgen05.commit) + the issuing thread's release of s_tapmask
    s_tapmask = 0;
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  ADN_CTA_STAMP(2, 2);
  ptc.mark(5);
  const int gfirst = max(R0 - 1, 0), glast = min(R1, rows_total - 1);
  bool ok = true;
  if (R0 < R1) {
    if (warp == 4) {
      if (elect_one()) {
        if (!urow3_bulk_producer<NUW>(sU, u, zeros16, Strip{H, TPR}, gfirst, glast, full, empty)) atomicExch(status, 27);
      }
    } else if (warp == 5) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_rt(128, 3 * D, true, true);
        const uint32_t abase = smem_u32(sA), ubase = smem_u32(sU);
        const uint64_t dA0 = make_desc(abase, 128, 2048), dU0 = make_desc(ubase, 128, 2048);
        int next_wait = gfirst;
        uint32_t mask = 0;
        PhaseTimer pt(6, true);
        for (int R = R0; R < R1; ++R) {
          const int it = R - R0, stg = it % WG_NST, y = R % H;
          const int need = min(R + 1, glast);
          pt.mark(7);
          while (next_wait <= need) {
            const int i = next_wait - gfirst;
            ok = mbar_wait(&full[i % NUW], (i / NUW) & 1) && ok;
            ++next_wait;
          }
          pt.mark(0);
          ok = mbar_wait(&a_full[stg], (it / WG_NST) & 1) && ok;
          tc_fence_after();
          pt.mark(1);
          const uint32_t ab = abase + stg * WG_STG_B;
          // the three horizontal taps of image row y + a - 1 are one B operand with N = 96 (three shifted copies of the u row)
          if (mask == 0x1FFu && y > 0 && y < H - 1) {   // steady state: every tap accumulates, fully unrolled
            const uint64_t dA = dadd(dA0, stg * WG_STG_B);
            uint64_t dB[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) dB[a] = dadd(dU0, (uint32_t)(((R + a - 1) - gfirst) % NUW) * USLOT3_B);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_c<true>(tbase + a * 3 * D, dadd(dA, k * 256), dadd(dB[a], k * 256), idesc);
          } else {
            for (int a = 0; a < 3; ++a) {
              if (y + a - 1 < 0 || y + a - 1 >= H) continue;
              const uint32_t ub = ubase + (uint32_t)(((R + a - 1) - gfirst) % NUW) * USLOT3_B;
              const bool init = ((mask >> (3 * a)) & 1) != 0;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma(tbase + a * 3 * D, make_desc(ab + k * 256, 128, 2048), make_desc(ub + k * 256, 128, 2048), idesc, init || k > 0);
              mask |= 7u << (3 * a);
            }
          }
          umma_commit(&a_empty[stg]);
          if (R - 1 >= gfirst) umma_commit(&empty[((R - 1) - gfirst) % NUW]);
          pt.mark(2);
        }
        s_tapmask = mask;
        umma_commit(&done);
        mbar_arrive(&done);
        ADN_CTA_STAMP_ANY(2, 3);
        if (!ok) atomicExch(status, 25);
      }
    } else {
      if (warp == 0 && elect_one()) {   // dpre tile producer
        for (int R = R0; R < R1; ++R) {
          const int it = R - R0, stg = it % WG_NST;
          if (it >= WG_NST) ok = mbar_wait(&a_empty[stg], ((it / WG_NST) - 1) & 1) && ok;
          uint8_t* sb = sA + stg * WG_STG_B;
          const long long tile = Strip{H, TPR}.tile(R);
          if (mb == 0) {
            mbar_expect_tx(&a_full[stg], 16 * 2048);
            bulk_g2s(sb, dpre + tile * NA * 1024, 16 * 2048, &a_full[stg]);
          } else {
            mbar_expect_tx(&a_full[stg], 8 * 2048 + 4096);
            bulk_g2s(sb, dpre + (tile * NA + 16) * 1024, 8 * 2048, &a_full[stg]);
            bulk_g2s(sb + 8 * 2048, ddt + tile * 2 * 1024, 4096, &a_full[stg]);
          }
        }
      }
      __syncwarp();
      PhaseTimer pt(6, tid == 32);
      ok = mbar_wait(&done, 0) && ok;
      tc_fence_after();
      pt.mark(3);
      const uint32_t mask = *reinterpret_cast<volatile uint32_t*>(&s_tapmask);
      const int m = warp * 32 + lane;
      const int c = mb * 128 + m;      // conv channel (mb 1: m < 64), or dt row 192 + (m - 64)
      // every CTA writes its own slab of partial sums (no atomics: 74 CTAs per address would serialise in L2);
      // k_finalize_fast adds the slabs.  Taps this CTA never accumulated (image-border rows only) count as zero.
      float* dKs = dK + (long long)part * (CC * 9);
      float* dWs = dWin + (long long)part * (DIP * D);
      if (mb == 0 || warp < 2) {      // warp-uniform: tcgen05.ld is .sync.aligned
        float wrow[D], dw[D];
#pragma unroll
        for (int d = 0; d < D; d += 4) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(Win + c * D + d));
          wrow[d] = t4.x; wrow[d + 1] = t4.y; wrow[d + 2] = t4.z; wrow[d + 3] = t4.w;
          dw[d] = dw[d + 1] = dw[d + 2] = dw[d + 3] = 0.f;
        }
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          float dk = 0.f;
          if (ok && ((mask >> t) & 1)) {     // warp-uniform
            float v0[16], v1[16];
            tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + t * D, v0);
            tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + t * D + 16, v1);
            tmem_wait_ld();
            const float kt = __ldg(Kc + c * 9 + t);
            float dk1 = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
              dk = fmaf(wrow[d], v0[d], dk);
              dk1 = fmaf(wrow[16 + d], v1[d], dk1);
              dw[d] = fmaf(kt, v0[d], dw[d]);
              dw[16 + d] = fmaf(kt, v1[d], dw[16 + d]);
            }
            dk += dk1;
          }
          dKs[c * 9 + t] = dk;
        }
#pragma unroll
        for (int d = 0; d < D; d += 4) *reinterpret_cast<float4*>(dWs + c * D + d) = make_float4(dw[d], dw[d + 1], dw[d + 2], dw[d + 3]);
      } else if (warp == 2) {        // mb 1, rows 64..79: the dt rows of W_in (centre tap)
        float v0[16], v1[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) { v0[d] = 0.f; v1[d] = 0.f; }
        if (ok && ((mask >> 4) & 1)) {
          tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + 4 * D, v0);
          tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + 4 * D + 16, v1);
          tmem_wait_ld();
        }
        if (lane < NH) {
          float* dst = dWs + (CC + lane) * D;
#pragma unroll
          for (int d = 0; d < 16; d += 4) {
            *reinterpret_cast<float4*>(dst + d) = make_float4(v0[d], v0[d + 1], v0[d + 2], v0[d + 3]);
            *reinterpret_cast<float4*>(dst + 16 + d) = make_float4(v1[d], v1[d + 1], v1[d + 2], v1[d + 3]);
          }
        }
      }
      pt.mark(4);
      if (!ok && lane == 0) atomicExch(status, 26);
    }
  }
  tc_fence_before();
  __syncthreads();
  ptc.mark(6);
  ADN_CTA_STAMP(2, 1);
  if (warp == 5) tmem_dealloc(tbase, 512);
}

}  // namespace rowconv
