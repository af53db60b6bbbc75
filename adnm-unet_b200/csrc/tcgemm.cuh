// General bf16 x bf16 -> fp32 GEMM on the sm_100a tensor cores (tcgen05.mma, accumulator in TMEM), used by the "wide"
// ADN-SSD path (d_model 64..1024, d_state up to 128, any token-grid size): every contraction of the mixer that the
// 32-wide row / tile kernels fuse by hand is, at these widths, one launch of this kernel.
//
//   C[b][m][n] (=|+=) alpha * sum_seg sum_k opA_seg[b][m][k] * opB_seg[b][n][k]
//
// * Operands are row-major global tensors in either orientation, no copies or transposes on the host side:
//     K-major  : stored [rows = M or N][K], row pitch ld       (x @ W^T with W = [N][K]: in_proj, out_proj, readout ...)
//     MN-major : stored [K][cols = M or N], row pitch ld       (reductions over tokens: state, weight gradients; x @ W)
//   Both are fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle, one elected thread, completion on an mbarrier) through
//   tensor maps encoded per launch; boxes that reach outside the matrix are zero-filled by the hardware, so ragged M / N / K
//   need no padding and sub-matrix views (row pitch > width) never read their neighbours.  K-major tiles are [rows][128 B],
//   MN-major tiles [64 K rows][128 B] per 64-wide MN group; only the UMMA descriptor differs.  (Round-2 history: the first
//   version staged un-swizzled tiles with 16-byte cp.async copies; knock-out timing showed those copies - 8 L1 tags per warp
//   instruction - cost 200 of 395 us on the in_proj shape and capped a square 8192^3 GEMM at 207 TFLOP/s.)
// * Up to two (A, B, K) segments accumulate into the same tile: [LN(y) | zc] @ W_out^T without a concatenated buffer, and
//   fp32 states as bf16 hi + lo pairs.
// * blockIdx.z = batch * splitk + split: per-sample GEMMs (state, readout) and split-K for the token reductions
//   (fp32 atomics into a zeroed accumulator).
// * Persistent CTAs (one per SM) over 128 x BN tiles (BN <= 256, multiple of 16), BK = 64, 4-stage TMA ring; a producer
//   lane, an MMA-issuing lane and four epilogue warps run concurrently, the accumulator double-buffered in TMEM (2 x BN of
//   the 512 columns) so that a tile's epilogue overlaps the next tile's main loop.
// * Epilogue: TMEM -> registers (alpha, parity mask, optional SiLU'(aux) factor) -> 128-byte-swizzled staging slab in
//   shared memory -> TMA store (bf16 / fp32) or TMA reduce-add (split-K), 32 rows x 128 bytes per instruction, clipped to
//   the matrix by the hardware.  (Thread-per-row global stores cost 107 of 157 us on the in_proj shape.)
// * Implicit-GEMM 3x3 convolution (`Conv`, dense convs of WTLayer / OutProj, models/model_untils.py:375-386,820-831): one
//   operand is a channels-last image [B][H][W][C] read through a RANK-4 tensor map; a tile of 128 (or 64) consecutive tokens
//   is a box of whole image rows, and tap (dy, dx) is the same box moved by (dx, dy) - rows / columns / samples that fall
//   outside the image are zero-filled by the TMA unit, which IS the convolution's zero padding.  No im2col buffer exists.
//     mode 1: A = image (tokens are M), K = 9 taps x cpad channels        forward and data gradient (flipped taps)
//     mode 2: B = image (tokens are K), N = 9 taps x cpad channels        weight gradient (split-K over tokens)
#pragma once
#include "adn_common.cuh"
#include "sm100_utils.cuh"
#include "tma_utils.cuh"

namespace adn {
namespace tcg {
using namespace adn::sm100;
using namespace adn::tma;

constexpr int BM = 128, BK = 64, STAGES = 4, MAX_BN = 256;
constexpr int EPI_BUF_B = 4096, EPI_B = 4 * 2 * EPI_BUF_B;      // per TMEM lane quarter (a PAIR of epilogue warps): two 32-row x 128-byte staging slabs
constexpr int A_TILE_B = BM * BK * 2;                    // 16 KB in either orientation
constexpr int THREADS = 320;      // 8 epilogue warps (two per TMEM lane quarter, splitting each slab's columns) + MMA warp + TMA producer warp
constexpr int W_MMA = 8, W_TMA = 9;

enum { C_BF16 = 0, C_F32 = 1, C_ATOMIC_F32 = 2 };

struct Seg {
  const bf16* A; long long lda, a_bs;
  const bf16* B; long long ldb, b_bs;
  int K;
};

struct Conv {
  int mode;          // 0: plain GEMM; 1: operand A is the image (K-major); 2: operand B is the image (MN-major)
  int W, H;          // image size; L = W * H tokens per sample
  int cpad;          // image channels rounded up to 64: tap t owns k (mode 1) / n (mode 2) range [t * cpad, (t + 1) * cpad)
  int bn;            // tile width override (0: pick_bn).  The weight gradient is bound by the number of k-tile hand-offs per CTA
                     // (the token reduction is long, the tiles are few): 256-wide tiles even when the last one is mostly padding
};
static const Conv NOCONV = Conv{0, 0, 0, 0, 0};

struct Args {
  Seg seg[2];
  int nseg;
  int M, N, BN;
  int a_mn, b_mn;          // 1: operand stored [K][M] / [K][N] (MN-major), 0: [M][K] / [N][K] (K-major)
  void* C; long long ldc, c_bs;
  int c_mode;
  const float* alpha;      // optional device scalar
  int parity_mask;         // 1: keep only (m & 1) == (n & 1)   (the even/odd SSD split, models/ADNssd.py:397-404)
  int splitk, k_per_split; // split-K over segment 0 (nseg must be 1 when splitk > 1); k_per_split is a multiple of BK
  int* status;             // set to 1 on a pipeline time-out
  const bf16* aux; long long ld_aux, aux_bs;   // optional epilogue operand, indexed like C: C = acc * SiLU'(aux)   (conv backward:
                                              // the gradient w.r.t. the conv output becomes the gradient w.r.t. its input)
  int a_bz[2], b_bz[2];    // 1: the operand of that segment has a batch dimension (else every batch reads the same matrix)
  const float* bias;       // optional fp32 vector [N] added after alpha (1x1 conv / Linear bias)
  Conv conv;               // implicit-GEMM 3x3 convolution (mode 0: off)
  int dbg;                 // knock-outs for profiling (adn_set_option("gemm_dbg")): 1 no epilogue stores, 2 no operand loads, 4 no MMAs,
                           // 8 no proxy fence, 16 no epilogue math / staging, 32 no TMEM loads.  Measured (profiles/gemm_knockout.py,
                           // gpurun_out/r3g): the bare barrier skeleton costs ~0.45 us per k-tile commit whatever the ring depth (4 ... 8
                           // stages: no change) - 8192^3 is bound by it (785 of 892 us with loads, MMAs and stores all knocked out) - and on
                           // the skinny GEMMs (K <= 128) the epilogue arithmetic + staging is another 40-50 % of the launch
};

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t ssrc, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(ssrc), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t ssrc, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(ssrc), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// one box of a rank-4 tensor map (channels-last image: c, x, y, sample); out-of-range coordinates are zero-filled
__device__ __forceinline__ void tma_load_4d(uint32_t sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(sdst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// Shared-memory matrix descriptors for the 128-byte swizzled tiles TMA writes (cute/atom/mma_traits_sm100.hpp,
// make_umma_desc: canonical layouts in 16-byte units)
//   K-major  Swizzle<3,4,3> o ((8,n),2):((8,SBO),1)        rows 128 B apart, 8-row groups SBO = 1024 B apart, LBO unused (1)
//   MN-major Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO))   64 MN elements per 128-B row, K rows 128 B apart, 8-row K
//            groups SBO = 1024 B apart, 64-wide MN groups LBO apart (= one 64 x 64 box = 8192 B)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;      // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t tile, int ks) { return desc_sw128(tile + ks * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t tile, int ks) { return desc_sw128(tile + ks * 2048, 8192, 1024); }

// Persistent CTA: tiles (batch x split, m tile, n tile; n fastest so that consecutive CTAs share the A tile in L2) are taken
// round-robin.  Three warp roles run concurrently and hand work over through mbarriers only:
//   warp  9    producer: TMA ring (STAGES deep) that runs continuously across tile boundaries
//   warp  8    MMA issuer: one elected lane; accumulators double-buffered in TMEM (2 x BN columns)
//   warps 0-7  epilogue: drain accumulator buffer i & 1 while the MMAs of tile i + 1 run.  Warps w and w + 4 share TMEM lane
//              quarter w & 3 and split the column groups of every 128-byte slab (the skinny GEMMs of the mixer / FeedForward /
//              Mlp / conv stages have ONE or two k-tiles per tile and were bound by four epilogue warps: knock-out timing in
//              DESIGN.md, and GELU in the epilogue made inference slower); the pair shares the staging slabs and meets at a
//              64-thread named barrier before the slab leaves through TMA
static __global__ void __launch_bounds__(THREADS)
k_tcgemm(const Args a, int tiles_m, int tiles_n, int tiles_total, const __grid_constant__ CUtensorMap mA0,
         const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mA1, const __grid_constant__ CUtensorMap mB1,
         const __grid_constant__ CUtensorMap mC) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = a.BN;
  const int stage_b = A_TILE_B + BN * BK * 2;
  uint32_t tcols = 32;
  while ((int)tcols < BN) tcols <<= 1;
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc(&tmem_slot, 2 * tcols);
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&mC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;      // 128-byte swizzle atoms: 1024-byte aligned tiles
  bool ok = true;
  const int nk1 = a.nseg > 1 ? (a.seg[1].K + BK - 1) / BK : 0;
  const int per_z = tiles_m * tiles_n;

  // k-tile range of a tile's split
#define TCG_DECODE(tile)                                                                   \
  const int z = (tile) / per_z, mn = (tile) - z * per_z;                                  \
  const int mt = mn / tiles_n, nt_ = mn - mt * tiles_n;                                   \
  const int m0 = mt * BM, n0 = nt_ * BN;                                                  \
  const int batch = z / a.splitk, split = z - batch * a.splitk;                           \
  int kb0 = 0, ke0 = a.seg[0].K;                                                          \
  if (a.splitk > 1) { kb0 = split * a.k_per_split; ke0 = min(a.seg[0].K, kb0 + a.k_per_split); } \
  const int nk0 = (ke0 - kb0 + BK - 1) / BK;                                              \
  const int nk = nk0 + nk1;

  if (warp == W_TMA) {
    // ---------------- TMA producer (one lane)
    if (lane == 0) {
      tma_prefetch_desc(&mA0); tma_prefetch_desc(&mB0);
      if (a.nseg > 1) { tma_prefetch_desc(&mA1); tma_prefetch_desc(&mB1); }
      int it = 0;                                  // k-tile sequence number of this CTA, across tiles
      for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        TCG_DECODE(tile)
        (void)split;
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) ok &= mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
          const bool second = kt >= nk0;
          const int k0 = second ? (kt - nk0) * BK : kb0 + kt * BK;
          const CUtensorMap* pa = second ? &mA1 : &mA0;
          const CUtensorMap* pb = second ? &mB1 : &mB0;
          const uint32_t sa = s0 + s * stage_b, sb = sa + A_TILE_B;
          mbar_expect_tx(&full[s], (uint32_t)stage_b);
          if (!(a.dbg & 2)) {
            // split-K: a box may reach past this split's range into the next one's; harmless for K-major / MN-major alike
            // because splits are whole multiples of BK (k_per_split % BK == 0)
            const int za = batch * (second ? a.a_bz[1] : a.a_bz[0]), zb = batch * (second ? a.b_bz[1] : a.b_bz[0]);
            if (a.conv.mode == 1) {
              // A tile = 128 consecutive tokens of the image starting at token m0, moved by this k-tile's tap
              const int cpt = a.conv.cpad >> 6, tap = kt / cpt, c0 = (kt - tap * cpt) << 6;
              const int L = a.conv.W * a.conv.H, b0 = m0 / L, r = m0 - b0 * L, y0 = r / a.conv.W, x0 = r - y0 * a.conv.W;
              tma_load_4d(sa, pa, c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0, &full[s]);
            } else if (!a.a_mn) tma_load_3d(sa, pa, k0, m0, za, &full[s]);
            else { tma_load_3d(sa, pa, m0, k0, za, &full[s]); tma_load_3d(sa + 8192, pa, m0 + 64, k0, za, &full[s]); }
            if (a.conv.mode == 2) {
              // B tile = 64 consecutive tokens (the k range) x BN columns; every 64-column group belongs to one tap
              const int L = a.conv.W * a.conv.H, b0 = k0 / L, r = k0 - b0 * L, y0 = r / a.conv.W, x0 = r - y0 * a.conv.W;
              for (int q = 0; q < BN; q += 64) {
                const int n = n0 + q, tap = n / a.conv.cpad;
                if (tap < 9) tma_load_4d(sb + q * 128, pb, n - tap * a.conv.cpad, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0, &full[s]);
                else tma_load_4d(sb + q * 128, pb, a.conv.cpad, 0, 0, 0, &full[s]);      // past the last tap: an all-zero box
              }
            } else if (!a.b_mn) tma_load_3d(sb, pb, k0, n0, zb, &full[s]);
            else
              for (int q = 0; q < BN; q += 64) tma_load_3d(sb + q * 128, pb, n0 + q, k0, zb, &full[s]);
          } else {
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)stage_b) : "memory");
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_rt(BM, BN, a.a_mn != 0, a.b_mn != 0);
      int it = 0, i = 0;
      for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++i) {
        TCG_DECODE(tile)
        (void)m0; (void)n0; (void)batch; (void)split;
        const int ab = i & 1;
        if (i >= 2) ok &= mbar_wait(&acc_empty[ab], ((i >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tbase + ab * tcols;
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const int s = it % STAGES;
          ok &= mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = s0 + s * stage_b, sb = sa + A_TILE_B;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t da = a.a_mn ? desc_mn_sw128(sa, ks) : desc_k_sw128(sa, ks);
            const uint64_t db = a.b_mn ? desc_mn_sw128(sb, ks) : desc_k_sw128(sb, ks);
            if (!(a.dbg & 4)) umma(tacc, da, db, idesc, (kt | ks) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[ab]);
      }
    }
  } else {
    // ---------------- epilogue: TMEM lane quarter `qd`, thread = output row; slabs of 128 bytes per row leave through TMA.
    // Column groups of 16 accumulator columns: a slab holds 4 (bf16) or 2 (fp32) of them, half of which belong to this warp.
    const float alpha = a.alpha ? *a.alpha : 1.f;
    const int qd = warp & 3, half = warp >> 2;
    const uint32_t sbuf = s0 + STAGES * stage_b + qd * 2 * EPI_BUF_B;      // the pair's two staging slabs (1024-byte aligned)
    const int cw = a.c_mode == C_BF16 ? 64 : 32;                             // columns per 128-byte slab
    const int gpw = cw >> 5;                                                 // column groups per warp and slab: 2 (bf16) or 1 (fp32)
#define TCG_PAIR_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(qd + 1) : "memory")
    int i = 0, slab = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++i) {
      TCG_DECODE(tile)
      (void)split; (void)nk;
      const int ab = i & 1;
      ok &= mbar_wait(&acc_full[ab], (i >> 1) & 1);
      tc_fence_after();
      const int row = qd * 32 + lane, m = m0 + row;
      const uint32_t tacc = tbase + ab * tcols;
      for (int c0 = 0; c0 < BN && n0 + c0 < a.N; c0 += cw, ++slab) {
        const uint32_t sb = sbuf + (slab & 1) * EPI_BUF_B;
        if (slab >= 2) {                       // the TMA store that last read this slab has finished reading shared memory
          if (half == 0 && lane == 0) bulk_wait_read<1>();
          TCG_PAIR_SYNC();
        }
        // all TMEM loads of the warp's share of the slab are issued before the one wait
        float vv[2][16];
        if (!(a.dbg & 32)) {
#pragma unroll
          for (int gi = 0; gi < 2; ++gi)
            if (gi < gpw) tmem_ld16(tmem_addr(tacc, qd * 32, c0 + (half * gpw + gi) * 16), vv[gi]);
          tmem_wait_ld();
        }
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {          // 16 accumulator columns per step
          if (gi >= gpw || (a.dbg & 16)) break;
          const int q = half * gpw + gi;           // column group inside the slab
          float (&v)[16] = vv[gi];
          const int n = n0 + c0 + q * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] *= alpha;
            if (a.parity_mask && ((m ^ (n + j)) & 1)) v[j] = 0.f;
          }
          if (a.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < a.N) v[j] += __ldg(a.bias + n + j);
          }
          if (a.aux != nullptr && m < a.M) {
            const bf16* ax = a.aux + (long long)batch * a.aux_bs + (long long)m * a.ld_aux + n;
            if (n + 16 <= a.N && ((uintptr_t)ax & 15) == 0) {      // two 16-byte loads per row instead of 16 scalar ones
              float x[16];
              unpack8(reinterpret_cast<const uint4*>(ax)[0], *reinterpret_cast<float(*)[8]>(x));
              unpack8(reinterpret_cast<const uint4*>(ax)[1], *reinterpret_cast<float(*)[8]>(x + 8));
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= silu_gradf_(x[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n + j < a.N) v[j] *= silu_gradf_(__bfloat162float(ax[j]));
            }
          }
          // row `lane` of the slab, 16-byte chunk index XOR (row & 7): the 128-byte swizzle the tensor map expects
          if (a.c_mode == C_BF16) {
            const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
            const float hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
            const uint4 p0 = pack8(lo), p1 = pack8(hi);
            const int ch = q * 2;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + lane * 128 + (((ch) ^ (lane & 7)) << 4)), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + lane * 128 + (((ch + 1) ^ (lane & 7)) << 4)), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + lane * 128 + (((q * 4 + t) ^ (lane & 7)) << 4)), "f"(v[4 * t]), "f"(v[4 * t + 1]), "f"(v[4 * t + 2]), "f"(v[4 * t + 3]) : "memory");
          }
        }
        if (!(a.dbg & 8)) fence_async_smem();
        TCG_PAIR_SYNC();                       // both warps of the quarter have written (and fenced) their columns
        if (half == 0 && lane == 0 && !(a.dbg & 1)) {
          if (a.c_mode == C_ATOMIC_F32) tma_reduce_add_3d(&mC, sb, n0 + c0, m0 + qd * 32, batch);
          else tma_store_3d(&mC, sb, n0 + c0, m0 + qd * 32, batch);
          bulk_commit();
        }
      }
      // this accumulator buffer may be overwritten by the MMAs of tile i + 2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive1(&acc_empty[ab]);
    }
    if (half == 0 && lane == 0) bulk_wait_read<0>();      // shared memory must outlive the last TMA stores' reads
    __syncwarp();
#undef TCG_PAIR_SYNC
  }
#undef TCG_DECODE
  if (!ok && a.status) *a.status = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tbase, 2 * tcols);
}

// ---------------------------------------------------------------- host side
struct Op {          // one operand: pointer, row pitch, batch stride, orientation
  const bf16* p; long long ld, bs; int mn;
};
static inline Op kmaj(const bf16* p, long long ld, long long bs = 0) { return Op{p, ld, bs, 0}; }
static inline Op mnmaj(const bf16* p, long long ld, long long bs = 0) { return Op{p, ld, bs, 1}; }

struct Out {
  void* p; long long ld, bs; int mode;
};

static inline int pick_bn(int N, int b_mn) {
  // MN-major B: whole 64-wide swizzle groups (TMA zero-fills past N); K-major B: any multiple of 16 rows
  const int g = b_mn ? 64 : 16;
  const int r = (N + g - 1) / g * g;
  if (r <= MAX_BN) return r;
  // wider than one tile: 256 unless its zero-padded last tile wastes more than 1/8 of the work, then 128
  const int pad256 = (N + 255) / 256 * 256 - N;
  return pad256 * 8 <= N ? 256 : 128;
}

// Rank-3 bf16 tensor map of one operand: K-major [rows][K] -> dims (K, rows, batches), box (64, box_rows, 1);
// MN-major [K][cols] -> dims (cols, K, batches), box (64, 64, 1).  128-byte swizzle, zero fill outside the matrix.
static int make_operand_map(CUtensorMap* map, const char* name, Op o, int extent_mn, int K, int batches, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  ADN_REQUIRE(enc != nullptr, ADN_ERR_CUDA, "tcgemm %s: cuTensorMapEncodeTiled is not available from this driver", name);
  const bool has_batch = batches > 1 && o.bs != 0;
  cuuint64_t dims[3] = {(cuuint64_t)(o.mn ? extent_mn : K), (cuuint64_t)(o.mn ? K : extent_mn), (cuuint64_t)(has_batch ? batches : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)o.ld * 2, has_batch ? (cuuint64_t)o.bs * 2 : (cuuint64_t)o.ld * 2 * dims[1]};
  cuuint32_t box[3] = {64, (cuuint32_t)(o.mn ? 64 : box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)o.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADN_REQUIRE(r == CUDA_SUCCESS, ADN_ERR_CUDA, "tcgemm %s: cuTensorMapEncodeTiled failed (%d) dims %llu x %llu x %llu pitch %llu", name,
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)strides[0]);
  return ADN_OK;
}

// Rank-4 bf16 tensor map of a channels-last image [B][H][W][C]: dims (C, W, H, B), box = (64 channels, bw, bh, bb) with
// bw * bh * bb = tokens_per_box consecutive tokens (whole rows / whole samples), 128-byte swizzle, zero fill outside.
struct Image {
  const bf16* p; int B, H, W, C;
};
static int make_image_map(CUtensorMap* map, const char* name, Image im, int tokens_per_box) {
  EncodeTiledFn enc = encode_tiled_fn();
  ADN_REQUIRE(enc != nullptr, ADN_ERR_CUDA, "tcgemm %s: cuTensorMapEncodeTiled is not available from this driver", name);
  const int tpb = tokens_per_box;
  ADN_REQUIRE(im.C % 8 == 0 && ((uintptr_t)im.p % 16) == 0, ADN_ERR_SHAPE, "tcgemm %s: image channels must be a multiple of 8 and the base 16-byte aligned", name);
  ADN_REQUIRE(im.W >= tpb ? im.W % tpb == 0 : tpb % im.W == 0, ADN_ERR_SHAPE, "tcgemm %s: image width %d does not tile %d-token boxes", name, im.W, tpb);
  const int bw = im.W < tpb ? im.W : tpb, rows = tpb / bw;
  ADN_REQUIRE(im.H >= rows ? im.H % rows == 0 : rows % im.H == 0, ADN_ERR_SHAPE, "tcgemm %s: image height %d does not tile %d-row boxes", name, im.H, rows);
  const int bh = im.H < rows ? im.H : rows, bb = rows / bh;
  cuuint64_t dims[4] = {(cuuint64_t)im.C, (cuuint64_t)im.W, (cuuint64_t)im.H, (cuuint64_t)im.B};
  cuuint64_t strides[3] = {(cuuint64_t)im.C * 2, (cuuint64_t)im.W * im.C * 2, (cuuint64_t)im.H * im.W * im.C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)im.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADN_REQUIRE(r == CUDA_SUCCESS, ADN_ERR_CUDA, "tcgemm %s: cuTensorMapEncodeTiled (image) failed (%d) dims %d x %d x %d x %d box %d x %d x %d", name,
              (int)r, im.C, im.W, im.H, im.B, bw, bh, bb);
  return ADN_OK;
}
// can make_image_map() tile this image for both the 128-token (forward / data gradient) and 64-token (weight gradient) boxes?
static inline bool image_tiles(int H, int W) {
  for (int tpb = 64; tpb <= 128; tpb *= 2) {
    if (!(W >= tpb ? W % tpb == 0 : tpb % W == 0)) return false;
    const int rows = tpb / (W < tpb ? W : tpb);
    if (!(H >= rows ? H % rows == 0 : rows % H == 0)) return false;
  }
  return true;
}

struct Aux {
  const bf16* p; long long ld, bs;
};
static const Aux NOAUX = Aux{nullptr, 0, 0};

// C = alpha * (A0 . B0 [+ A1 . B1]);  batches > 1: per-sample GEMMs;  splitk > 1: atomics into a ZEROED fp32 C.
static int gemm(cudaStream_t st, const char* name, int M, int N, int K0, Op A0, Op B0, int K1, Op A1, Op B1, Out C,
                int batches, int splitk, const float* alpha, int parity_mask, int* status, Aux aux = NOAUX, const float* bias = nullptr,
                Conv conv = NOCONV, Image image = Image{nullptr, 0, 0, 0, 0}) {
  ADN_REQUIRE(M > 0 && N > 0 && K0 > 0 && batches > 0, ADN_ERR_SHAPE, "tcgemm %s: empty problem", name);
  ADN_REQUIRE(K1 == 0 || (A1.mn == A0.mn && B1.mn == B0.mn), ADN_ERR_SHAPE, "tcgemm %s: segments must share orientation", name);
  ADN_REQUIRE(splitk == 1 || (K1 == 0 && C.mode == C_ATOMIC_F32), ADN_ERR_SHAPE, "tcgemm %s: split-K needs one segment and an atomic fp32 output", name);
  // TMA granularity: 16-byte aligned bases, row pitches and batch strides
  ADN_REQUIRE(A0.ld % 8 == 0 && B0.ld % 8 == 0 && (K1 == 0 || (A1.ld % 8 == 0 && B1.ld % 8 == 0)), ADN_ERR_SHAPE, "tcgemm %s: row pitches must be multiples of 8", name);
  ADN_REQUIRE(A0.bs % 8 == 0 && B0.bs % 8 == 0 && A1.bs % 8 == 0 && B1.bs % 8 == 0, ADN_ERR_SHAPE, "tcgemm %s: batch strides must be multiples of 8", name);
  ADN_REQUIRE(((uintptr_t)A0.p | (uintptr_t)B0.p | (uintptr_t)A1.p | (uintptr_t)B1.p) % 16 == 0, ADN_ERR_SHAPE, "tcgemm %s: operands must be 16-byte aligned", name);
  Args a;
  a.seg[0] = Seg{A0.p, A0.ld, A0.bs, B0.p, B0.ld, B0.bs, K0};
  a.seg[1] = Seg{A1.p, A1.ld, A1.bs, B1.p, B1.ld, B1.bs, K1};
  a.nseg = K1 > 0 ? 2 : 1;
  a.M = M; a.N = N; a.BN = conv.bn > 0 ? conv.bn : pick_bn(N, B0.mn);
  a.a_mn = A0.mn; a.b_mn = B0.mn;
  a.C = C.p; a.ldc = C.ld; a.c_bs = C.bs; a.c_mode = C.mode;
  a.alpha = alpha; a.parity_mask = parity_mask;
  a.splitk = splitk < 1 ? 1 : splitk;
  a.k_per_split = (cdiv(cdiv(K0, a.splitk), BK)) * BK;
  a.splitk = cdiv(K0, a.k_per_split);
  a.status = status;
  a.aux = aux.p; a.ld_aux = aux.ld; a.aux_bs = aux.bs;
  a.bias = bias;
  a.a_bz[0] = batches > 1 && A0.bs != 0; a.b_bz[0] = batches > 1 && B0.bs != 0;
  a.a_bz[1] = batches > 1 && A1.bs != 0; a.b_bz[1] = batches > 1 && B1.bs != 0;
  a.dbg = env().gemm_dbg;
  a.conv = conv;
  ADN_REQUIRE(conv.mode == 0 || (K1 == 0 && batches == 1 && conv.cpad % 64 == 0 && conv.cpad >= image.C), ADN_ERR_SHAPE,
              "tcgemm %s: a convolution is one un-batched segment", name);
  ADN_REQUIRE(conv.mode != 1 || (A0.mn == 0 && a.splitk == 1 && K0 == 9 * conv.cpad && (long long)image.B * image.H * image.W == M), ADN_ERR_SHAPE,
              "tcgemm %s: inconsistent convolution (forward) extents", name);
  ADN_REQUIRE(conv.mode != 2 || (B0.mn == 1 && N == 9 * conv.cpad && (long long)image.B * image.H * image.W == K0), ADN_ERR_SHAPE,
              "tcgemm %s: inconsistent convolution (weight gradient) extents", name);
  CUtensorMap mA0, mB0, mA1, mB1;
  int rc = conv.mode == 1 ? make_image_map(&mA0, name, image, BM) : make_operand_map(&mA0, name, A0, M, K0, batches, BM);
  if (rc) return rc;
  rc = conv.mode == 2 ? make_image_map(&mB0, name, image, BK) : make_operand_map(&mB0, name, B0, N, K0, batches, a.BN);
  if (rc) return rc;
  if (K1 > 0) {
    rc = make_operand_map(&mA1, name, A1, M, K1, batches, BM);
    if (rc) return rc;
    rc = make_operand_map(&mB1, name, B1, N, K1, batches, a.BN);
    if (rc) return rc;
  } else {
    mA1 = mA0; mB1 = mB0;
  }
  // output map: dims (N, M, batches), box (128 bytes of columns, 32 rows, 1), 128-byte swizzle; stores are clipped to it
  CUtensorMap mC;
  {
    EncodeTiledFn enc = encode_tiled_fn();
    const bool bf = C.mode == C_BF16;
    const int es = bf ? 2 : 4;
    ADN_REQUIRE(((uintptr_t)C.p % 16) == 0 && (C.ld * es) % 16 == 0 && (C.bs * es) % 16 == 0, ADN_ERR_SHAPE,
                "tcgemm %s: output base / row pitch / batch stride must be 16-byte aligned", name);
    const bool has_batch = batches > 1 && C.bs != 0;
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)(has_batch ? batches : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)C.ld * es, has_batch ? (cuuint64_t)C.bs * es : (cuuint64_t)C.ld * es * M};
    cuuint32_t box[3] = {(cuuint32_t)(128 / es), 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mC, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, C.p, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADN_REQUIRE(r == CUDA_SUCCESS, ADN_ERR_CUDA, "tcgemm %s: cuTensorMapEncodeTiled (output) failed (%d)", name, (int)r);
    ADN_REQUIRE(batches == 1 || has_batch, ADN_ERR_SHAPE, "tcgemm %s: a batched GEMM needs a batch stride on its output", name);
  }
  const size_t smem = (size_t)STAGES * (A_TILE_B + a.BN * BK * 2) + EPI_B + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    ADN_CHECK_CUDA(cudaFuncSetAttribute(k_tcgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * (A_TILE_B + MAX_BN * BK * 2) + EPI_B + 1024));
    attr_done = true;
  }
  const int tiles_m = cdiv(M, BM), tiles_n = cdiv(N, a.BN);
  const long long tiles = (long long)tiles_m * tiles_n * batches * a.splitk;
  ADN_REQUIRE(tiles < (1LL << 31), ADN_ERR_SHAPE, "tcgemm %s: too many tiles", name);
  const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());      // persistent: one CTA per SM
  { ADN_KERNEL(name, st); k_tcgemm<<<grid, THREADS, smem, st>>>(a, tiles_m, tiles_n, (int)tiles, mA0, mB0, mA1, mB1, mC); }
  return ADN_OK;
}

// number of K splits that fills the machine for a token reduction with `tiles` output tiles
static inline int pick_splitk(int tiles, int K) {
  int want = cdiv(2 * sm_count(), tiles < 1 ? 1 : tiles);
  int maxs = cdiv(K, 4 * BK);            // at least 4 k-tiles per split
  if (maxs < 1) maxs = 1;
  return want < 1 ? 1 : (want > maxs ? maxs : want);
}

static const Op NOOP = Op{nullptr, 8, 0, 0};

}  // namespace tcg
}  // namespace adn
