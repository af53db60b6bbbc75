// sm_100a tensor-core (tcgen05 / TMEM) path of the ADN-SSD mixer, bf16 I/O, for the shapes ADNM-UNet's
// full-resolution refiner uses (d_model 32, headdim 4, d_state 16) and neighbours.  Same stage split and the same
// saved-tensor layout as the generic path (adnssd_generic.cuh); each stage is replaced by a fused kernel:
//   k_prep          fp32 master weights -> bf16 hi/lo operand images, conv kernel assembly
//   k_inproj        raw = u . W_in^T           tcgen05 (bf16 hi + lo weights, fp32 accumulate in TMEM)
//   k_conv_fwd8     depthwise 3x3 + SiLU       128-bit loads, register window, coalesced channels-last
//   k_conv_bwd4     conv backward + dK         same mapping, block-level reduction of the kernel gradient
#include "adnssd_sm100.cuh"

#include <stdlib.h>

#include "adnssd_generic.cuh"
#include "sm100_utils.cuh"

namespace adn {
using namespace sm100;
typedef bf16 TWf;   // workspace intermediates of the fast path are bf16 (tensor-core operands)

// ---- optional in-kernel phase timers (diagnostics, adn_phase_*): thread 0 of every CTA adds the SM-clock cycles spent
// between consecutive PHASE marks into g_phase[kernel][phase]; disabled (one predicated branch per mark) by default.
__device__ unsigned long long g_phase[8][8];
__device__ int g_phase_on = 0;
// per-CTA start / end times (globaltimer, ns) of the row kernels: [kernel 0 k_fconv, 1 k_bconv_du, 2 k_bconv_wg][CTA][2]
__device__ unsigned long long g_cta_t[3][160][4];
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef ADN_PHASE_TIMING
#define ADN_CTA_STAMP(kid, which) do { if (g_phase_on != 0 && threadIdx.x == 0 && blockIdx.x < 160) g_cta_t[kid][blockIdx.x][which] = gtime_ns(); } while (0)
#define ADN_CTA_STAMP_ANY(kid, which) do { if (g_phase_on != 0 && blockIdx.x < 160) g_cta_t[kid][blockIdx.x][which] = gtime_ns(); } while (0)
#else
#define ADN_CTA_STAMP(kid, which) do { } while (0)
#define ADN_CTA_STAMP_ANY(kid, which) do { } while (0)
#endif
#ifdef ADN_PHASE_TIMING     // build with -DADN_PHASE_TIMING (python -m adnm_unet_b200.build --phase-timing); costs registers
struct PhaseTimer {
  long long t;
  int kid;
  bool on;
  __device__ __forceinline__ PhaseTimer(int k) : t(0), kid(k), on(g_phase_on != 0 && threadIdx.x == 0) { if (on) t = clock64(); }
  __device__ __forceinline__ PhaseTimer(int k, bool mine) : t(0), kid(k), on(g_phase_on != 0 && mine) { if (on) t = clock64(); }
  __device__ __forceinline__ void mark(int ph) {
    if (on) { long long n = clock64(); atomicAdd(&g_phase[kid][ph], (unsigned long long)(n - t)); t = n; }
  }
};
#else
struct PhaseTimer {
  __device__ __forceinline__ PhaseTimer(int) {}
  __device__ __forceinline__ PhaseTimer(int, bool) {}
  __device__ __forceinline__ void mark(int) {}
};
#endif

// ------------------------------------------------------------------------------------------------
// UMMA self-test: one CTA, one 128 x N x K problem, operands staged in the T8 layout.
//   mode 0: A[128][K], B[N][K] row-major (both K-major):   C = A . B^T
//   mode 1: A[K][128], B[K][N] row-major (both MN-major):  C[m][n] = sum_k A[k][m] * B[k][n]
// Validates descriptors, instruction descriptor, commit/mbarrier, TMEM read-back (tests/test_umma_gpu.py).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_umma_selftest(int mode, int N, int K, const bf16* __restrict__ A, const bf16* __restrict__ Bm, float* __restrict__ C,
                int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  bf16* sA = reinterpret_cast<bf16*>(smem);
  bf16* sB = sA + 128 * K;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (mode == 0) {
    for (int i = tid; i < 128 * K; i += 128) { int r = i / K, c = i % K; sA[t8_off(r, c, 128)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, c = i % K; sB[t8_off(r, c, N)] = Bm[i]; }
  } else {
    for (int i = tid; i < K * 128; i += 128) { int t = i / 128, c = i % 128; sA[t8_off(t, c, K)] = A[i]; }
    for (int i = tid; i < K * N; i += 128) { int t = i / N, c = i % N; sB[t8_off(t, c, K)] = Bm[i]; }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_rt(128, N, mode == 1, mode == 1);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int k = 0; k < K; k += 16) {
      uint64_t da = mode == 0 ? desc_kmajor(a0, 128, 0, k) : desc_mnmajor(a0, K, 0, k);
      uint64_t db = mode == 0 ? desc_kmajor(b0, N, 0, k) : desc_mnmajor(b0, K, 0, k);
      umma(tbase, da, db, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) C[(long long)tid * N + c + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// Shifted / padded operand self-test (tests/test_umma_gpu.py::test_umma_shift): the fused conv kernels address a
// zero-padded row buffer [chunk][pitch][8] through descriptors whose start address is moved by whole rows (16 bytes) and
// whose chunk stride is pitch*16 rather than a multiple of 128 bytes.
//   mode 0: A[pitch][K], B[pitch][K] row-major, K-major operands:  C = A[shiftA : shiftA+128] . B[shiftB : shiftB+N]^T
//   mode 1: A[pitch][128], B[pitch][N] row-major, MN-major operands: C[m][n] = sum_{k<K} A[shiftA+k][m] * B[shiftB+k][n]
__global__ void __launch_bounds__(128)
k_umma_shift_selftest(int mode, int N, int K, int pitch, int shiftA, int shiftB, const bf16* __restrict__ A,
                      const bf16* __restrict__ Bm, float* __restrict__ C, int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int colsA = mode == 0 ? K : 128, colsB = mode == 0 ? K : N;
  bf16* sA = reinterpret_cast<bf16*>(smem);
  bf16* sB = sA + pitch * colsA;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < pitch * colsA; i += 128) { int r = i / colsA, c = i % colsA; sA[t8_off(r, c, pitch)] = A[i]; }
  for (int i = tid; i < pitch * colsB; i += 128) { int r = i / colsB, c = i % colsB; sB[t8_off(r, c, pitch)] = Bm[i]; }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_rt(128, N, mode == 1, mode == 1);
    const uint32_t a0 = smem_u32(sA) + shiftA * 16, b0 = smem_u32(sB) + shiftB * 16;
    for (int k = 0; k < K; k += 16) {
      uint64_t da, db;
      if (mode == 0) {
        da = make_desc(a0 + (k >> 3) * pitch * 16, pitch * 16, 128);
        db = make_desc(b0 + (k >> 3) * pitch * 16, pitch * 16, 128);
      } else {
        da = make_desc(a0 + k * 16, 128, pitch * 16);
        db = make_desc(b0 + k * 16, 128, pitch * 16);
      }
      umma(tbase, da, db, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) C[(long long)tid * N + c + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// tcgen05.mma issue-rate probe (diagnostics): thread 0 issues `iters` back-to-back 128 x N x 16 MMAs on operands staged with
// row pitch `pitch` and start shift `shift` rows; cycles[0] = SM clocks from first issue to completion.
__global__ void __launch_bounds__(128)
k_umma_bench(int mode, int N, int pitch, int shift, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_rt(128, N, mode == 1, mode == 1);
    const uint32_t a0 = smem_u32(smem) + shift * 16, b0 = smem_u32(smem) + 24 * 1024 + shift * 16;
    const uint64_t da = mode == 0 ? make_desc(a0, pitch * 16, 128) : make_desc(a0, 128, pitch * 16);
    const uint64_t db = mode == 0 ? make_desc(b0, pitch * 16, 128) : make_desc(b0, 128, pitch * 16);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma(tbase, da, db, idesc, true);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// ------------------------------------------------------------------------------------------------
// cp.async helpers (16-byte, L2-only caching; src_bytes = 0 zero-fills)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Load a [128 tokens][CH*8 channels] bf16 tile (row pitch ld elements, first channel c0) into T8 shared memory.
// Lane mapping: 8 consecutive tokens x consecutive 16-byte chunks per warp instruction: conflict-free shared-memory
// writes, 8 x (CH*16)-byte contiguous global segments.  Rows >= rows_valid are zero-filled.
__device__ __forceinline__ void load_tile_t8(bf16* sdst, const bf16* __restrict__ gsrc, long long ld, int CH,
                                             int rows_valid, int tid, int nthreads) {
  // lane = (token-in-block tok8, chunk-in-group cq): one warp instruction moves 8 tokens x 4 consecutive chunks.
  // Warps stride over the 16 token blocks, an inner loop strides over the chunk groups: adds only, no div / mod.
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int tok8 = lane & 7, cq = lane >> 3;
  for (int tb = warp; tb < 16; tb += nwarps) {
    const int tok = tb * 8 + tok8;
    const bool ok = tok < rows_valid;
    const bf16* src = gsrc + (long long)(ok ? tok : 0) * ld + cq * 8;
    bf16* dst = sdst + (cq * 128 + tok) * 8;
    const int nb = ok ? 16 : 0;
    for (int c = cq; c < CH; c += 4) {
      cp_async16(dst, src, nb);
      src += 32;
      dst += 4 * 128 * 8;
    }
  }
}

// thread-0 helper: one bulk copy of `nchunks` consecutive 8-channel chunks of token tile `tile` of a TL tensor
__device__ __forceinline__ void tl_bulk(bf16* sdst, const bf16* __restrict__ g, long long tile, int NCH, int chunk0,
                                        int nchunks, uint64_t* bar) {
  bulk_g2s(sdst, g + ((tile * NCH + chunk0) * 128) * 8, (uint32_t)nchunks * 128 * 16, bar);
}

// ------------------------------------------------------------------------------------------------
// k_inproj: raw[T][ldr] = u[T][D] . W_in^T   (models/ADNssd.py:309).  Persistent CTAs, 128 tokens per tile.
// A = u tile (K-major, cp.async double buffered), B = W_in rows [n0, n0+nn) as hi and lo bf16 images (K-major):
// D(tmem) = A.B_hi^T + A.B_lo^T with fp32 accumulation, so raw carries no weight-rounding error.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
k_inproj(const bf16* __restrict__ u, const bf16* __restrict__ Whi, const bf16* __restrict__ Wlo, bf16* __restrict__ raw,
         int ldr, int dip, long long T, int num_tiles, int* __restrict__ status) {
  constexpr int CH = D / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  bf16* sWhi = reinterpret_cast<bf16*>(smem);   // [CH][256][8]
  bf16* sWlo = sWhi + 256 * D;
  bf16* sA0 = sWlo + 256 * D;                   // 2 x [CH][128][8]
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  uint32_t parity = 0;
  for (int n0 = 0; n0 < dip; n0 += 256) {
    const int nn = min(256, dip - n0);
    __syncthreads();
    for (int i = tid; i < nn * CH; i += 128) {
      int j = i / CH, dc = i % CH;
      const uint4 h = *reinterpret_cast<const uint4*>(Whi + (long long)(n0 + j) * D + dc * 8);
      const uint4 l = *reinterpret_cast<const uint4*>(Wlo + (long long)(n0 + j) * D + dc * 8);
      *reinterpret_cast<uint4*>(sWhi + (dc * nn + j) * 8) = h;
      *reinterpret_cast<uint4*>(sWlo + (dc * nn + j) * 8) = l;
    }
    int tile = blockIdx.x, stage = 0;
    if (tile < num_tiles) {
      long long t0 = (long long)tile * 128;
      load_tile_t8(sA0, u + t0 * D, D, CH, (int)min((long long)128, T - t0), tid, 128);
    }
    cp_async_commit();
    for (; tile < num_tiles; tile += gridDim.x, stage ^= 1) {
      const int nxt = tile + gridDim.x;
      if (nxt < num_tiles) {
        long long t1 = (long long)nxt * 128;
        load_tile_t8(sA0 + (stage ^ 1) * 128 * D, u + t1 * D, D, CH, (int)min((long long)128, T - t1), tid, 128);
      }
      cp_async_commit();
      cp_async_wait<1>();
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA0 + stage * 128 * D), bh = smem_u32(sWhi), bl = smem_u32(sWlo);
        const uint32_t idesc = make_idesc_rt(128, nn, false, false);
#pragma unroll
        for (int k = 0; k < D; k += 16) umma(tbase, desc_kmajor(a0, 128, 0, k), desc_kmajor(bh, nn, 0, k), idesc, k > 0);
#pragma unroll
        for (int k = 0; k < D; k += 16) umma(tbase, desc_kmajor(a0, 128, 0, k), desc_kmajor(bl, nn, 0, k), idesc, true);
        umma_commit(&bar);
      }
      const bool ok = mbar_wait(&bar, parity);
      parity ^= 1;
      tc_fence_after();
      if (!ok) { if (tid == 0) atomicExch(status, 1); break; }
      // epilogue: raw is stored in the tiled layout, so thread t (token row t) writes 16 bytes at row t of each
      // 8-channel chunk: consecutive threads -> consecutive 16-byte slots, fully coalesced straight from registers
      {
        const int NR = ldr >> 3;
        bf16* dst = raw + (((long long)tile * NR + (n0 >> 3)) * 128 + tid) * 8;
        for (int c = 0; c < nn; c += 32) {
          float v0[16], v1[16];
          tmem_ld16(tmem_addr(tbase, warp * 32, c), v0);
          if (c + 16 < nn) tmem_ld16(tmem_addr(tbase, warp * 32, c + 16), v1);
          tmem_wait_ld();
          float a[8], b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { a[j] = v0[j]; b[j] = v0[8 + j]; }
          *reinterpret_cast<uint4*>(dst + (long long)(c >> 3) * 1024) = pack8(a);
          *reinterpret_cast<uint4*>(dst + (long long)((c >> 3) + 1) * 1024) = pack8(b);
          if (c + 16 < nn) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = v1[j]; b[j] = v1[8 + j]; }
            *reinterpret_cast<uint4*>(dst + (long long)((c >> 3) + 2) * 1024) = pack8(a);
            *reinterpret_cast<uint4*>(dst + (long long)((c >> 3) + 3) * 1024) = pack8(b);
          }
        }
      }
      tc_fence_before();   // TMEM reads of this tile are ordered before the next tile's MMA (after the next __syncthreads)
    }
    cp_async_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}


// bf16x8 (one 16-byte vector) -> four float2
__device__ __forceinline__ void unpack8_f2(const uint4& u, float2 (&v)[4]) {
  v[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
  v[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  v[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
  v[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
}
__device__ __forceinline__ uint4 pack8_f2(const float2 (&v)[4]) {
  return make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y), pack_bf16(v[3].x, v[3].y));
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// softplus with hardware exp2 / log2 (bf16 path only; the fp32 check mode keeps log1pf): abs error ~1e-7
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }

}  // namespace adn
#include "adnssd_rowconv.cuh"
#include "adnssd_bwdws.cuh"
namespace adn {
using namespace sm100;

// ------------------------------------------------------------------------------------------------
// k_state: S'[b][j][c] += [j%2==c%2] * sum_l Bc[l,j] * w[l,hd(c)] * xc[l,c]      (models/ADNssd.py:267-280, both parities)
// One CTA = a slice of the tiles of one sample.  Per 128-token tile: x, B and the dt columns arrive by cp.async in
// the T8 layout; each thread (= token) turns its dt values into w = softplus(dt+bias)*exp(A_log) and scales its x row
// in place; then  D(tmem)[c][j] += sum_tok wx[tok][c] * Bc[tok][j]  is ONE tcgen05 reduction over the tokens
// (A = wx and B = Bc both MN-major, M padded to 128 channels with zero chunks), accumulated in TMEM across tiles.
// headdim 4: the 8 channels of chunk cg belong to heads 2cg (even channels) and 2cg+1 (odd channels).
// ------------------------------------------------------------------------------------------------
template <int DI, int GN>
__global__ void __launch_bounds__(128)
k_state(const bf16* __restrict__ act, const bf16* __restrict__ raw, int ldr, int CC, const float* __restrict__ dt_bias,
        const float* __restrict__ A_log, float* __restrict__ S, int L, int tiles_per_batch, int ctas_per_batch,
        int* __restrict__ status) {
  constexpr int XC = DI / 8, BC = GN / 8, DC = DI / 32, NH = DI / 4;
  constexpr int STAGE = (16 + BC + DC) * 128 * 8;   // bf16 elements per stage
  constexpr uint32_t TCOLS = GN <= 32 ? 32 : (GN <= 64 ? 64 : 128);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_bias[NH], s_eA[NH];
  __shared__ uint64_t bar[2], ld_bar[2];
  __shared__ uint32_t tmem_slot;
  bf16* sbase = reinterpret_cast<bf16*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / ctas_per_batch, part = blockIdx.x % ctas_per_batch;
  const int NA = CC >> 3, NR = ldr >> 3;
  for (int i = tid; i < NH; i += 128) { s_bias[i] = dt_bias[i]; s_eA[i] = __expf(A_log[i]); }
  // zero the padding chunks [XC, 16) of the wx operand in both stages
  for (int st = 0; st < 2; ++st)
    for (int i = tid; i < (16 - XC) * 128; i += 128)
      *reinterpret_cast<uint4*>(sbase + st * STAGE + (XC * 128 + i) * 8) = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&ld_bar[0], 1); mbar_init(&ld_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  fence_async_smem();     // the zero padding written above must be visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = make_idesc_rt(128, GN, true, true);
  int it = 0;
  bool ok = true;
  // act / raw are stored tiled (TL): x, B and the dt columns of a 128-token tile are three contiguous blocks
  auto issue_loads = [&](int i, int stage) {
    if (tid == 0) {
      bf16* sX = sbase + stage * STAGE;
      bf16* sB = sX + 16 * 128 * 8;
      bf16* sDt = sB + BC * 128 * 8;
      const long long tile = ((long long)b * L >> 7) + i;
      mbar_expect_tx(&ld_bar[stage], (uint32_t)(XC + BC + DC) * 128 * 16);
      tl_bulk(sX, act, tile, NA, XC, XC, &ld_bar[stage]);
      tl_bulk(sB, act, tile, NA, 2 * XC, BC, &ld_bar[stage]);
      tl_bulk(sDt, raw, tile, NR, NA, DC, &ld_bar[stage]);
    }
  };
  if (part < tiles_per_batch) issue_loads(part, 0);
  for (int i = part; i < tiles_per_batch; i += ctas_per_batch, ++it) {
    const int stage = it & 1;
    bf16* sX = sbase + stage * STAGE;
    bf16* sB = sX + 16 * 128 * 8;
    bf16* sDt = sB + BC * 128 * 8;
    // prefetch the next tile into the other stage once the MMAs that read it (iteration it-1) are done
    if (i + ctas_per_batch < tiles_per_batch) {
      if (it >= 1) ok = ok && mbar_wait(&bar[stage ^ 1], ((it - 1) >> 1) & 1);
      issue_loads(i + ctas_per_batch, stage ^ 1);
    }
    ok = ok && mbar_wait(&ld_bar[stage], (it >> 1) & 1);
    float w[NH];
#pragma unroll
    for (int dc = 0; dc < DC; ++dc) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(sDt + (dc * 128 + tid) * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) w[dc * 8 + j] = softplus_fast(v[j] + s_bias[dc * 8 + j]) * s_eA[dc * 8 + j];
    }
#pragma unroll
    for (int cg = 0; cg < XC; ++cg) {
      uint4* slot = reinterpret_cast<uint4*>(sX + (cg * 128 + tid) * 8);
      float v[8];
      unpack8(*slot, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= w[2 * cg + (j & 1)];
      *slot = pack8(v);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sX), b0 = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < 128; k += 16)
        umma(tbase, desc_mnmajor(a0, 128, 0, k), desc_mnmajor(b0, 128, 0, k), idesc, it > 0 || k > 0);
      umma_commit(&bar[stage]);
    }
  }
  if (it > 0) {
    const int last = (it - 1) & 1;
    ok = ok && mbar_wait(&bar[last], ((it - 1) >> 1) & 1);
    tc_fence_after();
    if (ok) {
      for (int c0 = 0; c0 < GN; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_addr(tbase, warp * 32, c0), v);
        tmem_wait_ld();
        if (tid < DI) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((((c0 + j) ^ tid) & 1) == 0) atomicAdd(S + ((long long)b * GN + c0 + j) * DI + tid, v[j]);
        }
      }
    } else if (tid == 0) {
      atomicExch(status, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

// ------------------------------------------------------------------------------------------------
// k_readout: per 128-token tile  y = Cc.S' + D*xc  ->  LayerNorm  ->  out = alpha1 * [LN(y) | zc] . W_out^T
// (models/ADNssd.py:281-283, :456-461).  Two tcgen05 GEMMs per tile; between them each thread owns one token row
// (TMEM lane = token), so the LayerNorm statistics are thread-local.  S' is used as bf16 hi + lo (fp32-accurate).
// ------------------------------------------------------------------------------------------------
template <int DI, int GN>
__global__ void __launch_bounds__(128)
k_readout(const bf16* __restrict__ act, int CC, const float* __restrict__ S, const float* __restrict__ Dp,
          const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ alpha1p,
          const bf16* __restrict__ Wout, bf16* __restrict__ out, int L, int tiles_per_batch, int num_tiles,
          int* __restrict__ status) {
  constexpr int D = DI / 2, XC = DI / 8, CCH = GN / 8, CATC = 2 * DI / 8;
  constexpr uint32_t TCOLS = (DI + D) <= 128 ? 128 : 256;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float sG[DI], sBt[DI], sDh[DI];
  __shared__ uint64_t bar, ld_bar;
  __shared__ uint32_t tmem_slot;
  bf16* sC = reinterpret_cast<bf16*>(smem);   // [CCH][128][8]
  bf16* sCat = sC + CCH * 128 * 8;            // [CATC][128][8]: chunks [0,XC) = LN(y), [XC,2XC) = zc
  bf16* sX = sCat + CATC * 128 * 8;           // [XC][128][8]
  bf16* sShi = sX + XC * 128 * 8;             // [CCH][DI][8]   S'^T as a K-major B operand: row = c, K = j
  bf16* sSlo = sShi + CCH * DI * 8;
  bf16* sW = sSlo + CCH * DI * 8;             // [CATC][D][8]   W_out: row = d, K = j'
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < DI; i += 128) { sG[i] = gamma[i]; sBt[i] = beta[i]; sDh[i] = Dp[head_of_channel(i, 4)]; }
  for (int i = tid; i < D * CATC; i += 128) {
    int d = i / CATC, jc = i % CATC;
    *reinterpret_cast<uint4*>(sW + (jc * D + d) * 8) = *reinterpret_cast<const uint4*>(Wout + (long long)d * 2 * DI + jc * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&ld_bar, 1); fence_mbar_init(); }
  const int NA = CC >> 3;
  uint32_t lph = 0;
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const float a1 = *alpha1p;
  uint32_t ph = 0;
  int cur_b = -1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_batch, i = tile % tiles_per_batch;
    const int rows = min(128, L - i * 128);
    const long long tok0 = (long long)b * L + (long long)i * 128;
    if (b != cur_b) {
      for (int idx = tid; idx < CCH * DI; idx += 128) {
        int jc = idx / DI, c = idx % DI;
        float v[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q] = S[((long long)b * GN + jc * 8 + q) * DI + c];
          lo[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
        }
        *reinterpret_cast<uint4*>(sShi + (jc * DI + c) * 8) = pack8(v);
        *reinterpret_cast<uint4*>(sSlo + (jc * DI + c) * 8) = pack8(lo);
      }
      cur_b = b;
    }
    if (tid == 0) {   // act is TL: [z | x] (chunks 0..2XC) land in sCat[XC..2XC) + sX, which are adjacent; C separately
      mbar_expect_tx(&ld_bar, (uint32_t)(2 * XC + CCH) * 128 * 16);
      tl_bulk(sCat + XC * 128 * 8, act, tile, NA, 0, 2 * XC, &ld_bar);
      tl_bulk(sC, act, tile, NA, 2 * XC + CCH, CCH, &ld_bar);
      if (tile + (int)gridDim.x < num_tiles) {   // pull the next tile's inputs into L2 while this one is processed
        const long long nt = tile + gridDim.x;
        bulk_prefetch_l2(act + (nt * NA) * 1024, (uint32_t)(2 * XC) * 2048);
        bulk_prefetch_l2(act + (nt * NA + 2 * XC + CCH) * 1024, (uint32_t)CCH * 2048);
      }
    }
    bool ok = mbar_wait(&ld_bar, lph);
    lph ^= 1;
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sC), bh = smem_u32(sShi), bl = smem_u32(sSlo);
      const uint32_t idesc = make_idesc_rt(128, DI, false, false);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase, desc_kmajor(a0, 128, 0, k), desc_kmajor(bh, DI, 0, k), idesc, k > 0);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase, desc_kmajor(a0, 128, 0, k), desc_kmajor(bl, DI, 0, k), idesc, true);
      umma_commit(&bar);
    }
    ok = mbar_wait(&bar, ph) && ok;
    ph ^= 1;
    tc_fence_after();
    if (!ok) { if (tid == 0) atomicExch(status, 3); break; }
    float y[DI];
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float v[16], x0[8], x1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), v);
      unpack8(*reinterpret_cast<const uint4*>(sX + ((cb / 8) * 128 + tid) * 8), x0);
      unpack8(*reinterpret_cast<const uint4*>(sX + ((cb / 8 + 1) * 128 + tid) * 8), x1);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[cb + j] = v[j] + sDh[cb + j] * x0[j];
        y[cb + 8 + j] = v[8 + j] + sDh[cb + 8 + j] * x1[j];
      }
    }
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < DI; ++c) mu += y[c];
    mu *= (1.f / DI);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < DI; ++c) { float dlt = y[c] - mu; var = fmaf(dlt, dlt, var); }
    const float rstd = rsqrtf(var * (1.f / DI) + 1e-5f);
#pragma unroll
    for (int cg = 0; cg < XC; ++cg) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (y[cg * 8 + j] - mu) * rstd * sG[cg * 8 + j] + sBt[cg * 8 + j];
      *reinterpret_cast<uint4*>(sCat + (cg * 128 + tid) * 8) = pack8(v);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sCat), b0 = smem_u32(sW);
      const uint32_t idesc = make_idesc_rt(128, D, false, false);
#pragma unroll
      for (int k = 0; k < 2 * DI; k += 16) umma(tbase + DI, desc_kmajor(a0, 128, 0, k), desc_kmajor(b0, D, 0, k), idesc, k > 0);
      umma_commit(&bar);
    }
    ok = mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    if (!ok) { if (tid == 0) atomicExch(status, 4); break; }
#pragma unroll
    for (int cb = 0; cb < D; cb += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, DI + cb), v);
      tmem_wait_ld();
      if (tid < rows) {
        float a[8], bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = a1 * v[j]; bq[j] = a1 * v[8 + j]; }
        bf16* dst = out + (tok0 + tid) * D + cb;
        *reinterpret_cast<uint4*>(dst) = pack8(a);
        *reinterpret_cast<uint4*>(dst + 8) = pack8(bq);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
  // Fail loudly without a host synchronisation: a kernel of this forward pass that timed out on a barrier (async-pipe
  // fault) left a non-zero status word; the last kernel of the pass then poisons the output with NaNs.
  if (blockIdx.x == 0 && tid < 8 && *reinterpret_cast<volatile int*>(status) != 0) out[tid] = __float2bfloat16_rn(__int_as_float(0x7fc00000));
}

// Build the bf16 hi / lo images of a per-sample fp32 matrix M[b][j][c] (GN x DI, the state S' or its gradient)
// as K-major B operands in both orientations:
//   "a": rows = c (DI), K = j (GN): element (c, j) at ((j/8)*DI + c)*8 + j%8     (used by  X[tok][c] = sum_j A[tok][j] M[j][c])
//   "b": rows = j (GN), K = c (DI): element (j, c) at ((c/8)*GN + j)*8 + c%8     (used by  X[tok][j] = sum_c A[tok][c] M[j][c])
template <int DI, int GN>
__device__ __forceinline__ void stage_state_a(const float* __restrict__ M, bf16* hi, bf16* lo, int tid) {
  for (int idx = tid; idx < (GN / 8) * DI; idx += 128) {
    int jc = idx / DI, c = idx % DI;
    float v[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q] = M[(jc * 8 + q) * DI + c];
      l[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
    }
    *reinterpret_cast<uint4*>(hi + (jc * DI + c) * 8) = pack8(v);
    *reinterpret_cast<uint4*>(lo + (jc * DI + c) * 8) = pack8(l);
  }
}
template <int DI, int GN>
__device__ __forceinline__ void stage_state_b(const float* __restrict__ M, bf16* hi, bf16* lo, int tid) {
  for (int idx = tid; idx < (DI / 8) * GN; idx += 128) {
    int cc = idx % (DI / 8), j = idx / (DI / 8);
    float v[8], l[8];
    const float4 p0 = *reinterpret_cast<const float4*>(M + j * DI + cc * 8);
    const float4 p1 = *reinterpret_cast<const float4*>(M + j * DI + cc * 8 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
#pragma unroll
    for (int q = 0; q < 8; ++q) l[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
    *reinterpret_cast<uint4*>(hi + (cc * GN + j) * 8) = pack8(v);
    *reinterpret_cast<uint4*>(lo + (cc * GN + j) * 8) = pack8(l);
  }
}

// ------------------------------------------------------------------------------------------------
// k_bwd1 (phase B1 of oracle/adnssd_oracle.py::mixer_backward), per 128-token tile:
//   g   = dout . W_out                      tcgen05, N = 2*DI           (d out / d [LN(y) | zc], without alpha1)
//   Y   = Cc . S'                           tcgen05 (recompute), then y = Y + D*x, LayerNorm statistics per thread
//   dy  = LN backward of alpha1*g_y,  dzc = alpha1*g_z                  -> dact[:, x] and dact[:, z]
//   dCc = dy . S'^T                         tcgen05                     -> dact[:, C]
//   Rt[j][d] += sum_tok [yhat | zc][tok][j] * dout[tok][d]   tcgen05 reduction over tokens (-> dW_out, dgamma, dbeta, dalpha1)
//   dS'[c][j] += sum_tok dy[tok][c] * Cc[tok][j]             tcgen05 reduction over tokens (flushed per sample)
// ------------------------------------------------------------------------------------------------
template <int DI, int GN>
__global__ void __launch_bounds__(128)
k_bwd1(const bf16* __restrict__ dout, const bf16* __restrict__ act, int CC, const float* __restrict__ S,
       const float* __restrict__ Dp, const float* __restrict__ gamma, const float* __restrict__ alpha1p,
       const bf16* __restrict__ Wout, bf16* __restrict__ dact, float* __restrict__ Rt, float* __restrict__ sdout,
       float* __restrict__ dS, int L, int tiles_per_batch, int num_tiles, int tiles_per_cta, int* __restrict__ status,
       const bf16* __restrict__ sgrad) {
  // sgrad != nullptr (row-kernel path): the z and C column blocks are written as dpre = dact * SiLU'(pre)
  constexpr int D = DI / 2, XC = DI / 8, CCH = GN / 8, DC = D / 8;
  constexpr int COL_Y = 2 * DI, COL_RT = COL_Y + DI, COL_DS = COL_RT + D;
  constexpr uint32_t TCOLS = (COL_DS + GN) <= 256 ? 256 : 512;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float sG[DI], sDh[DI];
  __shared__ uint64_t bar, ld_bar;
  __shared__ uint32_t tmem_slot;
  bf16* sDout = reinterpret_cast<bf16*>(smem);      // [DC][128][8]
  bf16* sC = sDout + DC * 128 * 8;                  // [CCH][128][8]
  bf16* sCat = sC + CCH * 128 * 8;                  // [16][128][8]   [yhat | zc]
  bf16* sXY = sCat + 16 * 128 * 8;                  // [16][128][8]   x, then dy in place; chunks >= XC stay zero
  bf16* sWT = sXY + 16 * 128 * 8;                   // [DC][2DI][8]   W_out^T: row j', K = d
  bf16* sSa_hi = sWT + DC * 2 * DI * 8;             // [CCH][DI][8]
  bf16* sSa_lo = sSa_hi + CCH * DI * 8;
  bf16* sSb_hi = sSa_lo + CCH * DI * 8;             // [XC][GN][8]
  bf16* sSb_lo = sSb_hi + XC * GN * 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < DI; i += 128) { sG[i] = gamma[i]; sDh[i] = Dp[head_of_channel(i, 4)]; }
  for (int i = tid; i < (16 - XC) * 128; i += 128)
    *reinterpret_cast<uint4*>(sXY + (XC * 128 + i) * 8) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < DC * 2 * DI; i += 128) {
    int dc = i / (2 * DI), j = i % (2 * DI);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __bfloat162float(Wout[(long long)(dc * 8 + q) * 2 * DI + j]);
    *reinterpret_cast<uint4*>(sWT + (dc * 2 * DI + j) * 8) = pack8(v);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&ld_bar, 1); fence_mbar_init(); }
  const int NA = CC >> 3;
  uint32_t lph = 0;
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  PhaseTimer pt(0);
  const float a1 = *alpha1p;
  uint32_t ph = 0;
  int cur_b = -1;
  bool rt_fresh = true, ds_fresh = true, ok = true;
  float sd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) sd[i] = 0.f;
  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(num_tiles, tile_begin + tiles_per_cta);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int b = tile / tiles_per_batch, i = tile % tiles_per_batch;
    const int rows = min(128, L - i * 128);
    const long long tok0 = (long long)b * L + (long long)i * 128;
    if (b != cur_b) {
      stage_state_a<DI, GN>(S + (long long)b * GN * DI, sSa_hi, sSa_lo, tid);
      stage_state_b<DI, GN>(S + (long long)b * GN * DI, sSb_hi, sSb_lo, tid);
      cur_b = b;
    }
    if (tid == 0) {   // act is TL: [z | x] -> sCat[XC..16) + sXY[0..XC) (adjacent in shared memory), C -> sC
      mbar_expect_tx(&ld_bar, (uint32_t)(2 * XC + CCH) * 128 * 16);
      tl_bulk(sCat + XC * 128 * 8, act, tile, NA, 0, 2 * XC, &ld_bar);
      tl_bulk(sC, act, tile, NA, 2 * XC + CCH, CCH, &ld_bar);
      if (tile + 1 < tile_end) {   // pull the next tile's inputs into L2 while this one is processed
        const long long nt = tile + 1;
        bulk_prefetch_l2(act + (nt * NA) * 1024, (uint32_t)(2 * XC) * 2048);
        bulk_prefetch_l2(act + (nt * NA + 2 * XC + CCH) * 1024, (uint32_t)CCH * 2048);
        bulk_prefetch_l2(dout + nt * 128 * D, 128 * D * 2);
        if (sgrad) {
          bulk_prefetch_l2(sgrad + (nt * NA) * 1024, (uint32_t)XC * 2048);
          bulk_prefetch_l2(sgrad + (nt * NA + 2 * XC + CCH) * 1024, (uint32_t)CCH * 2048);
        }
      }
    }
    load_tile_t8(sDout, dout + tok0 * D, D, DC, rows, tid, 128);   // dout is an external row-major tensor
    cp_async_commit();
    pt.mark(0);
    cp_async_wait<0>();
    ok = mbar_wait(&ld_bar, lph) && ok;
    lph ^= 1;
    fence_async_smem();
    __syncthreads();
    pt.mark(1);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aD = smem_u32(sDout), bW = smem_u32(sWT), aC = smem_u32(sC), bh = smem_u32(sSa_hi), bl = smem_u32(sSa_lo);
      const uint32_t id_g = make_idesc_rt(128, 2 * DI, false, false), id_y = make_idesc_rt(128, DI, false, false);
#pragma unroll
      for (int k = 0; k < D; k += 16) umma(tbase, desc_kmajor(aD, 128, 0, k), desc_kmajor(bW, 2 * DI, 0, k), id_g, k > 0);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase + COL_Y, desc_kmajor(aC, 128, 0, k), desc_kmajor(bh, DI, 0, k), id_y, k > 0);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase + COL_Y, desc_kmajor(aC, 128, 0, k), desc_kmajor(bl, DI, 0, k), id_y, true);
      umma_commit(&bar);
    }
    // SiLU' of this row's z and C columns, fetched into registers while the MMAs run (tiles are full: L % 128 == 0)
    uint4 sgz[XC], sgc[CCH];
    if (sgrad) {
      const bf16* srow = sgrad + (((long long)tile * NA) * 128 + tid) * 8;
#pragma unroll
      for (int q = 0; q < XC; ++q) sgz[q] = __ldg(reinterpret_cast<const uint4*>(srow + (long long)q * 1024));
#pragma unroll
      for (int q = 0; q < CCH; ++q) sgc[q] = __ldg(reinterpret_cast<const uint4*>(srow + (long long)(2 * XC + CCH + q) * 1024));
    }
    ok = mbar_wait(&bar, ph) && ok;
    ph ^= 1;
    tc_fence_after();
    pt.mark(2);
    if (!ok) { if (tid == 0) atomicExch(status, 5); break; }
    // ---- y, LayerNorm statistics, yhat
    float y[DI];
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float v[16], x0[8], x1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, COL_Y + cb), v);
      unpack8(*reinterpret_cast<const uint4*>(sXY + ((cb / 8) * 128 + tid) * 8), x0);
      unpack8(*reinterpret_cast<const uint4*>(sXY + ((cb / 8 + 1) * 128 + tid) * 8), x1);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[cb + j] = v[j] + sDh[cb + j] * x0[j];
        y[cb + 8 + j] = v[8 + j] + sDh[cb + 8 + j] * x1[j];
      }
    }
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < DI; ++c) mu += y[c];
    mu *= (1.f / DI);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < DI; ++c) { y[c] -= mu; var = fmaf(y[c], y[c], var); }
    const float rstd = rsqrtf(var * (1.f / DI) + 1e-5f);
#pragma unroll
    for (int cg = 0; cg < XC; ++cg) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { y[cg * 8 + j] *= rstd; v[j] = y[cg * 8 + j]; }
      *reinterpret_cast<uint4*>(sCat + (cg * 128 + tid) * 8) = pack8(v);
    }
    pt.mark(3);
    // ---- LN backward: dyh = alpha1 * g_y * gamma ; dy = rstd * (dyh - mean(dyh) - yhat * mean(dyh*yhat))
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float dyh = a1 * v[j] * sG[cb + j];
        m1 += dyh;
        m2 = fmaf(dyh, y[cb + j], m2);
      }
    }
    m1 *= (1.f / DI);
    m2 *= (1.f / DI);
    bf16* drow = dact + (((long long)tile * NA) * 128 + tid) * 8;   // dact is TL: chunk k of this row at drow + k*1024
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float v[16], o0[8], o1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o0[j] = rstd * (a1 * v[j] * sG[cb + j] - m1 - y[cb + j] * m2);
        o1[j] = rstd * (a1 * v[8 + j] * sG[cb + 8 + j] - m1 - y[cb + 8 + j] * m2);
      }
      const uint4 p0 = pack8(o0), p1 = pack8(o1);
      *reinterpret_cast<uint4*>(sXY + ((cb / 8) * 128 + tid) * 8) = p0;
      *reinterpret_cast<uint4*>(sXY + ((cb / 8 + 1) * 128 + tid) * 8) = p1;
      if (tid < rows) {
        *reinterpret_cast<uint4*>(drow + (long long)(XC + cb / 8) * 1024) = p0;
        *reinterpret_cast<uint4*>(drow + (long long)(XC + cb / 8 + 1) * 1024) = p1;
      }
    }
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float v[16], o0[8], o1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, DI + cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) { o0[j] = a1 * v[j]; o1[j] = a1 * v[8 + j]; }
      if (tid < rows) {
        if (sgrad) {
          float s0[8], s1[8];
          unpack8(sgz[cb / 8], s0);
          unpack8(sgz[cb / 8 + 1], s1);
#pragma unroll
          for (int j = 0; j < 8; ++j) { o0[j] *= s0[j]; o1[j] *= s1[j]; }
        }
        *reinterpret_cast<uint4*>(drow + (long long)(cb / 8) * 1024) = pack8(o0);
        *reinterpret_cast<uint4*>(drow + (long long)(cb / 8 + 1) * 1024) = pack8(o1);
      }
    }
#pragma unroll
    for (int dc = 0; dc < DC; ++dc) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(sDout + (dc * 128 + tid) * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) sd[dc * 8 + j] += v[j];
    }
    pt.mark(4);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aXY = smem_u32(sXY), bh = smem_u32(sSb_hi), bl = smem_u32(sSb_lo), aCat = smem_u32(sCat),
                     bD = smem_u32(sDout), bC = smem_u32(sC);
      const uint32_t id_c = make_idesc_rt(128, GN, false, false), id_r = make_idesc_rt(128, D, true, true),
                     id_s = make_idesc_rt(128, GN, true, true);
#pragma unroll
      for (int k = 0; k < DI; k += 16) umma(tbase, desc_kmajor(aXY, 128, 0, k), desc_kmajor(bh, GN, 0, k), id_c, k > 0);
#pragma unroll
      for (int k = 0; k < DI; k += 16) umma(tbase, desc_kmajor(aXY, 128, 0, k), desc_kmajor(bl, GN, 0, k), id_c, true);
#pragma unroll
      for (int k = 0; k < 128; k += 16)
        umma(tbase + COL_RT, desc_mnmajor(aCat, 128, 0, k), desc_mnmajor(bD, 128, 0, k), id_r, !rt_fresh || k > 0);
#pragma unroll
      for (int k = 0; k < 128; k += 16)
        umma(tbase + COL_DS, desc_mnmajor(aXY, 128, 0, k), desc_mnmajor(bC, 128, 0, k), id_s, !ds_fresh || k > 0);
      umma_commit(&bar);
    }
    rt_fresh = false;
    ds_fresh = false;
    ok = mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    pt.mark(5);
    if (!ok) { if (tid == 0) atomicExch(status, 6); break; }
#pragma unroll
    for (int cb = 0; cb < GN; cb += 16) {
      float v[16], o0[8], o1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) { o0[j] = v[j]; o1[j] = v[8 + j]; }
      if (tid < rows) {
        if (sgrad) {
          float s0[8], s1[8];
          unpack8(sgc[cb / 8], s0);
          unpack8(sgc[cb / 8 + 1], s1);
#pragma unroll
          for (int j = 0; j < 8; ++j) { o0[j] *= s0[j]; o1[j] *= s1[j]; }
        }
        *reinterpret_cast<uint4*>(drow + (long long)(2 * XC + CCH + cb / 8) * 1024) = pack8(o0);
        *reinterpret_cast<uint4*>(drow + (long long)(2 * XC + CCH + cb / 8 + 1) * 1024) = pack8(o1);
      }
    }
    const bool last_of_batch = (tile + 1 == tile_end) || ((tile + 1) / tiles_per_batch != b);
    if (last_of_batch) {
      for (int cb = 0; cb < GN; cb += 16) {
        float v[16];
        tmem_ld16(tmem_addr(tbase, warp * 32, COL_DS + cb), v);
        tmem_wait_ld();
        if (tid < DI) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((((cb + j) ^ tid) & 1) == 0) atomicAdd(dS + ((long long)b * GN + cb + j) * DI + tid, v[j]);
        }
      }
      ds_fresh = true;
    }
    tc_fence_before();
    __syncthreads();
    pt.mark(6);
  }
  if (ok && !rt_fresh) {
    for (int cb = 0; cb < D; cb += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, COL_RT + cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(Rt + tid * D + cb + j, v[j]);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float v = warp_sum(sd[i]);
      if ((tid & 31) == 0) atomicAdd(sdout + i, v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
}


// ------------------------------------------------------------------------------------------------
// k_bwd2 (phase B2), per 128-token tile, after dS' is complete:
//   G   = Bc . dS'                          tcgen05
//   dxc = D*dy + w*G (in place over dy), wx = w*x, dw[h] = sum_{c in h} x*G, ddt = dw*exp(A_log)*sigmoid(dt+bias)
//   dBc = wx . dS'^T                        tcgen05
//   per-head parameter gradients dD, dA_log, ddt_bias accumulate per thread and are reduced once per CTA.
// ------------------------------------------------------------------------------------------------
template <int DI, int GN>
__global__ void __launch_bounds__(128)
k_bwd2(const bf16* __restrict__ act, const bf16* __restrict__ raw, int ldr, int CC, const float* __restrict__ dS,
       const float* __restrict__ dt_bias, const float* __restrict__ A_log, const float* __restrict__ Dp,
       bf16* __restrict__ dact, bf16* __restrict__ draw, float* __restrict__ dD, float* __restrict__ dAlog,
       float* __restrict__ ddtb, int L, int tiles_per_batch, int num_tiles, int tiles_per_cta, int* __restrict__ status,
       const bf16* __restrict__ sgrad, int dt_nch, int dt_c0) {
  // dt columns: chunks [dt_c0, dt_c0 + DC) of `raw` and of `draw`, both TL tensors with dt_nch chunks per tile.
  // sgrad != nullptr (row-kernel path): the x and B column blocks are written as dpre = dact * SiLU'(pre)
  constexpr int XC = DI / 8, BC = GN / 8, DC = DI / 32, NH = DI / 4;
  constexpr uint32_t TCOLS = (DI + GN) <= 128 ? 128 : 256;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_bias[NH], s_eA[NH], s_D[NH];
  __shared__ uint64_t bar, ld_bar[2];
  __shared__ uint32_t tmem_slot;
  // two input stages (the next tile is fetched while this one is computed), each:
  //   sX [XC][128][8] x, then w*x in place | sB [BC][128][8] (x and B are adjacent chunks of act: one bulk copy)
  //   sDy [XC][128][8] | sDt [DC][128][8]
  constexpr int STAGE = (2 * XC + BC + DC) * 128 * 8;
  bf16* sStage = reinterpret_cast<bf16*>(smem);
  bf16* sTa_hi = sStage + 2 * STAGE;             // [BC][DI][8]
  bf16* sTa_lo = sTa_hi + BC * DI * 8;
  bf16* sTb_hi = sTa_lo + BC * DI * 8;           // [XC][GN][8]
  bf16* sTb_lo = sTb_hi + XC * GN * 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NH; i += 128) { s_bias[i] = dt_bias[i]; s_eA[i] = __expf(A_log[i]); s_D[i] = Dp[i]; }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&ld_bar[0], 1); mbar_init(&ld_bar[1], 1); fence_mbar_init(); }
  const int NA = CC >> 3;
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  PhaseTimer pt(1);
  uint32_t ph = 0;
  int cur_b = -1;
  bool ok = true;
  float aD[NH], aA[NH], aB[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) { aD[h] = 0.f; aA[h] = 0.f; aB[h] = 0.f; }
  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(num_tiles, tile_begin + tiles_per_cta);
  auto issue_loads = [&](int tile, int stage) {   // thread 0 only; TL tensors: [x | B] of act, dy of dact, dt columns of raw
    bf16* sX = sStage + stage * STAGE;
    bf16* sDy = sX + (XC + BC) * 128 * 8;
    bf16* sDt = sDy + XC * 128 * 8;
    mbar_expect_tx(&ld_bar[stage], (uint32_t)(2 * XC + BC + DC) * 128 * 16);
    tl_bulk(sX, act, tile, NA, XC, XC + BC, &ld_bar[stage]);
    tl_bulk(sDy, dact, tile, NA, XC, XC, &ld_bar[stage]);
    tl_bulk(sDt, raw, tile, dt_nch, dt_c0, DC, &ld_bar[stage]);
  };
  if (tid == 0 && tile_begin < tile_end) issue_loads(tile_begin, 0);
  for (int tile = tile_begin, it = 0; tile < tile_end; ++tile, ++it) {
    const int stage = it & 1;
    bf16* sX = sStage + stage * STAGE;
    bf16* sB = sX + XC * 128 * 8;
    bf16* sDy = sB + BC * 128 * 8;
    bf16* sDt = sDy + XC * 128 * 8;
    // prefetch: the other stage was last read by iteration it-1, whose MMAs completed and whose trailing barrier passed
    if (tid == 0 && tile + 1 < tile_end) {
      issue_loads(tile + 1, stage ^ 1);
      if (sgrad) bulk_prefetch_l2(sgrad + ((long long)(tile + 1) * NA + XC) * 1024, (uint32_t)(XC + BC) * 2048);
    }
    const int b = tile / tiles_per_batch, i = tile % tiles_per_batch;
    const int rows = min(128, L - i * 128);
    const long long tok0 = (long long)b * L + (long long)i * 128;
    if (b != cur_b) {
      stage_state_a<DI, GN>(dS + (long long)b * GN * DI, sTa_hi, sTa_lo, tid);
      stage_state_b<DI, GN>(dS + (long long)b * GN * DI, sTb_hi, sTb_lo, tid);
      cur_b = b;
    }
    pt.mark(0);
    ok = mbar_wait(&ld_bar[stage], (it >> 1) & 1) && ok;
    fence_async_smem();
    __syncthreads();
    pt.mark(1);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aB_ = smem_u32(sB), bh = smem_u32(sTa_hi), bl = smem_u32(sTa_lo);
      const uint32_t idesc = make_idesc_rt(128, DI, false, false);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase, desc_kmajor(aB_, 128, 0, k), desc_kmajor(bh, DI, 0, k), idesc, k > 0);
#pragma unroll
      for (int k = 0; k < GN; k += 16) umma(tbase, desc_kmajor(aB_, 128, 0, k), desc_kmajor(bl, DI, 0, k), idesc, true);
      umma_commit(&bar);
    }
    // SiLU' of this row's x and B columns, fetched into registers while the MMA runs (tiles are full: L % 128 == 0)
    uint4 sgx[XC], sgb[BC];
    if (sgrad) {
      const bf16* srow = sgrad + (((long long)tile * NA) * 128 + tid) * 8;
#pragma unroll
      for (int q = 0; q < XC; ++q) sgx[q] = __ldg(reinterpret_cast<const uint4*>(srow + (long long)(XC + q) * 1024));
#pragma unroll
      for (int q = 0; q < BC; ++q) sgb[q] = __ldg(reinterpret_cast<const uint4*>(srow + (long long)(2 * XC + q) * 1024));
    }
    // decay weights while the MMA runs
    float w[NH], sg[NH];
#pragma unroll
    for (int dc = 0; dc < DC; ++dc) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(sDt + (dc * 128 + tid) * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = v[j] + s_bias[dc * 8 + j];
        w[dc * 8 + j] = softplus_fast(a) * s_eA[dc * 8 + j];
        sg[dc * 8 + j] = a > 20.f ? 1.f : __fdividef(1.f, 1.f + __expf(-a));
      }
    }
    pt.mark(2);
    ok = mbar_wait(&bar, ph) && ok;
    ph ^= 1;
    tc_fence_after();
    pt.mark(3);
    if (!ok) { if (tid == 0) atomicExch(status, 7); break; }
    float dw[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) dw[h] = 0.f;
    bf16* drow = dact + (((long long)tile * NA) * 128 + tid) * 8;   // TL: chunk k of this row at drow + k*1024
    const bool valid = tid < rows;
#pragma unroll
    for (int cb = 0; cb < DI; cb += 16) {
      float G[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), G);
      tmem_wait_ld();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int cg = cb / 8 + half;
        float x[8], dy[8], o[8], wx[8];
        unpack8(*reinterpret_cast<const uint4*>(sX + (cg * 128 + tid) * 8), x);
        unpack8(*reinterpret_cast<const uint4*>(sDy + (cg * 128 + tid) * 8), dy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int h = 2 * cg + (j & 1);
          const float g = G[half * 8 + j];
          o[j] = s_D[h] * dy[j] + w[h] * g;
          wx[j] = w[h] * x[j];
          dw[h] = fmaf(x[j], g, dw[h]);
          if (valid) aD[h] = fmaf(dy[j], x[j], aD[h]);
        }
        *reinterpret_cast<uint4*>(sX + (cg * 128 + tid) * 8) = pack8(wx);
        if (valid) {
          if (sgrad) {
            float sv[8];
            unpack8(sgx[cg], sv);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= sv[j];
          }
          *reinterpret_cast<uint4*>(drow + (long long)(XC + cg) * 1024) = pack8(o);
        }
      }
    }
    {
      bf16* trow = draw + tl_off(tok0 + tid, dt_c0 * 8, dt_nch);   // draw is in the tiled layout: chunk stride 128*8
#pragma unroll
      for (int dc = 0; dc < DC; ++dc) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int h = dc * 8 + j;
          o[j] = dw[h] * s_eA[h] * sg[h];
          if (valid) { aA[h] = fmaf(dw[h], w[h], aA[h]); aB[h] += o[j]; }
        }
        if (valid) *reinterpret_cast<uint4*>(trow + (long long)dc * 128 * 8) = pack8(o);
      }
    }
    pt.mark(4);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aX = smem_u32(sX), bh = smem_u32(sTb_hi), bl = smem_u32(sTb_lo);
      const uint32_t idesc = make_idesc_rt(128, GN, false, false);
#pragma unroll
      for (int k = 0; k < DI; k += 16) umma(tbase + DI, desc_kmajor(aX, 128, 0, k), desc_kmajor(bh, GN, 0, k), idesc, k > 0);
#pragma unroll
      for (int k = 0; k < DI; k += 16) umma(tbase + DI, desc_kmajor(aX, 128, 0, k), desc_kmajor(bl, GN, 0, k), idesc, true);
      umma_commit(&bar);
    }
    ok = mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    pt.mark(5);
    if (!ok) { if (tid == 0) atomicExch(status, 8); break; }
#pragma unroll
    for (int cb = 0; cb < GN; cb += 16) {
      float v[16], o0[8], o1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, DI + cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) { o0[j] = v[j]; o1[j] = v[8 + j]; }
      if (valid) {
        if (sgrad) {
          float s0[8], s1[8];
          unpack8(sgb[cb / 8], s0);
          unpack8(sgb[cb / 8 + 1], s1);
#pragma unroll
          for (int j = 0; j < 8; ++j) { o0[j] *= s0[j]; o1[j] *= s1[j]; }
        }
        *reinterpret_cast<uint4*>(drow + (long long)(2 * XC + cb / 8) * 1024) = pack8(o0);
        *reinterpret_cast<uint4*>(drow + (long long)(2 * XC + cb / 8 + 1) * 1024) = pack8(o1);
      }
    }
    tc_fence_before();
    __syncthreads();
    pt.mark(6);
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float vD = warp_sum(aD[h]), vA = warp_sum(aA[h]), vB = warp_sum(aB[h]);
    if ((tid & 31) == 0) { atomicAdd(dD + h, vD); atomicAdd(dAlog + h, vA); atomicAdd(ddtb + h, vB); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

// ------------------------------------------------------------------------------------------------
// k_bwd4 (in_proj backward), per 128-token tile:
//   du = draw . W_in                        tcgen05 (hi + lo weights), K = dip
//   dW_in[j][d] += sum_tok draw[tok][j] * u[tok][d]      tcgen05 reduction over tokens, M blocks of 128 rows j
// ------------------------------------------------------------------------------------------------
template <int D, int MB, int NS>
__global__ void __launch_bounds__(128)
k_bwd4(const bf16* __restrict__ draw, int ldr, int dip, const bf16* __restrict__ u, const bf16* __restrict__ WThi,
       const bf16* __restrict__ WTlo, bf16* __restrict__ du, float* __restrict__ dWin_part, long long T, int num_tiles,
       int tiles_per_cta, int* __restrict__ status) {
  constexpr int DC = D / 8;
  constexpr int STAGE = (16 * MB + DC) * 128 * 8;   // bf16 elements per stage: draw tile (padded to 16*MB chunks) + u tile
  constexpr uint32_t TCOLS = (D + MB * D) <= 128 ? 128 : 256;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t ld_bar[2];
  __shared__ uint32_t tmem_slot;
  const int JC = dip / 8;                          // real 8-channel chunks of draw
  bf16* sStage = reinterpret_cast<bf16*>(smem);    // NS x { [16*MB][128][8] draw (chunks >= JC stay zero), [DC][128][8] u }
  bf16* sWhi = sStage + NS * STAGE;                // [16*MB][D][8]   W_in^T: row d, K = j
  bf16* sWlo = sWhi + 16 * MB * D * 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int st = 0; st < NS; ++st)
    for (int i = tid; i < (16 * MB - JC) * 128; i += 128)
      *reinterpret_cast<uint4*>(sStage + st * STAGE + (JC * 128 + i) * 8) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 16 * MB * D; i += 128) {   // the W_in^T images were laid out by k_prep: straight 16-byte copies
    *reinterpret_cast<uint4*>(sWhi + i * 8) = __ldg(reinterpret_cast<const uint4*>(WThi + i * 8));
    *reinterpret_cast<uint4*>(sWlo + i * 8) = __ldg(reinterpret_cast<const uint4*>(WTlo + i * 8));
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&ld_bar[0], 1); mbar_init(&ld_bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  PhaseTimer pt(2);
  uint32_t ph = 0;
  bool fresh = true, ok = true;
  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(num_tiles, tile_begin + tiles_per_cta);
  auto issue_loads = [&](int tile, int stage) {
    const long long tok0 = (long long)tile * 128;
    const int rows = (int)min((long long)128, T - tok0);
    bf16* sDraw = sStage + stage * STAGE;
    // draw is stored tiled (TL): the whole 128-token x dip tile is one contiguous block -> one bulk copy
    if (tid == 0) {
      mbar_expect_tx(&ld_bar[stage], (uint32_t)JC * 128 * 16);
      bulk_g2s(sDraw, draw + (long long)tile * JC * 128 * 8, (uint32_t)JC * 128 * 16, &ld_bar[stage]);
    }
    load_tile_t8(sDraw + 16 * MB * 128 * 8, u + tok0 * D, D, DC, rows, tid, 128);
  };
  if (NS == 2) {
    if (tile_begin < tile_end) issue_loads(tile_begin, 0);
    cp_async_commit();
  }
  for (int tile = tile_begin, it = 0; tile < tile_end; ++tile, ++it) {
    const int stage = NS == 2 ? (it & 1) : 0;
    // NS == 2: prefetch the next tile into the other stage: its previous reader (the MMAs of tile-1) completed before
    // the epilogue of the previous iteration, and every thread passed that iteration's trailing __syncthreads
    if (NS == 2) {
      if (tile + 1 < tile_end) issue_loads(tile + 1, stage ^ 1);
    } else {
      issue_loads(tile, 0);
    }
    cp_async_commit();
    pt.mark(0);
    if (NS == 2) cp_async_wait<1>(); else cp_async_wait<0>();
    ok = ok && mbar_wait(&ld_bar[stage], NS == 2 ? ((it >> 1) & 1) : (it & 1));
    fence_async_smem();
    __syncthreads();
    pt.mark(1);
    const long long tok0 = (long long)tile * 128;
    const int rows = (int)min((long long)128, T - tok0);
    bf16* sDraw = sStage + stage * STAGE;
    bf16* sU = sDraw + 16 * MB * 128 * 8;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aR = smem_u32(sDraw), bh = smem_u32(sWhi), bl = smem_u32(sWlo), bU = smem_u32(sU);
      const uint32_t id_u = make_idesc_rt(128, D, false, false), id_w = make_idesc_rt(128, D, true, true);
      // descriptors advance by a constant per K step: 2 chunks of the K-major tiles, 16 token rows of the MN-major ones
      uint64_t dA = desc_kmajor(aR, 128, 0, 0), dBh = desc_kmajor(bh, D, 0, 0), dBl = desc_kmajor(bl, D, 0, 0);
      for (int k = 0; k < dip; k += 16) {
        umma(tbase, dA, dBh, id_u, k > 0);
        umma(tbase, dA, dBl, id_u, true);
        dA = desc_advance(dA, 2 * 128 * 16);
        dBh = desc_advance(dBh, 2 * D * 16);
        dBl = desc_advance(dBl, 2 * D * 16);
      }
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        uint64_t dAm = desc_mnmajor(aR, 128, mb * 128, 0), dU = desc_mnmajor(bU, 128, 0, 0);
#pragma unroll
        for (int k = 0; k < 128; k += 16) {
          umma(tbase + D + mb * D, dAm, dU, id_w, !fresh || k > 0);
          dAm = desc_advance(dAm, 16 * 16);
          dU = desc_advance(dU, 16 * 16);
        }
      }
      umma_commit(&bar);
    }
    fresh = false;
    pt.mark(2);
    ok = mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    pt.mark(3);
    if (!ok) { if (tid == 0) atomicExch(status, 9); break; }
#pragma unroll
    for (int cb = 0; cb < D; cb += 16) {
      float v[16], o0[8], o1[8];
      tmem_ld16(tmem_addr(tbase, warp * 32, cb), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) { o0[j] = v[j]; o1[j] = v[8 + j]; }
      if (tid < rows) {
        bf16* dst = du + (tok0 + tid) * D + cb;
        *reinterpret_cast<uint4*>(dst) = pack8(o0);
        *reinterpret_cast<uint4*>(dst + 8) = pack8(o1);
      }
    }
    tc_fence_before();
    __syncthreads();
    pt.mark(4);
  }
  cp_async_wait<0>();
  pt.mark(5);
  {
    // one private slab of dip*D partial sums per CTA (summed by k_finalize): no atomics, deterministic
    float* slab = dWin_part + (long long)blockIdx.x * dip * D;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int j = mb * 128 + tid;
#pragma unroll
      for (int cb = 0; cb < D; cb += 16) {
        float v[16];
        if (ok && !fresh) {
          tmem_ld16(tmem_addr(tbase, warp * 32, D + mb * D + cb), v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.f;
        }
        if (j < dip) {
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            *reinterpret_cast<float4*>(slab + (long long)j * D + cb + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
}


// One launch for every parameter gradient of the tcgen05 path:
//   blocks [0, nb_in)            dW_in = sum of the per-CTA slabs written by k_bwd4            (32 elements x 8 slab lanes)
//   blocks [nb_in, nb_in + 2Di)  dW_out, dgamma, dbeta, dalpha1 partial from Rt / sum(dout)    (one warp per column)
//   remaining blocks             dD, dA_log, ddt_bias and the ten conv weight gradients from dK (shared finalize_body)
__global__ void __launch_bounds__(256)
k_finalize_fast(GradAcc a, AdnWeights w, AdnWeightGrads g, float* __restrict__ Rt, float* __restrict__ sdout,
                int D, int Di, int GN, int nh, int dip, int nb_in, int nb_rest, int rt_parts, const int* __restrict__ status,
                bf16* __restrict__ du) {
  __shared__ float red[8][33];
  const int tid = threadIdx.x;
  // Fail loudly without a host synchronisation: if a kernel of this backward pass flagged a pipeline fault (barrier
  // time-out), the gradients that leave this last kernel are poisoned with NaNs (and so is du[0..7]).
  const bool bad = *status != 0;
  float poison_late = bad ? __int_as_float(0x7fc00000) : 0.f;
  if (bad && blockIdx.x == 0 && tid < 8) du[tid] = __float2bfloat16_rn(poison_late);
  // ---- phase 1 (row-kernel / warp-specialised path): the backward kernels left one slab of partial sums per CTA instead
  // of contended atomics.  Every block adds up 32-element chunks (8 slab lanes x 32 elements, coalesced), then the grid
  // meets at a counter (this small grid is always co-resident) and phase 2 reads the reduced arrays.
  //   set 0: dW_in   set 1: dK   set 2: [Rt | sum(dout)] (reduced in place into slab 0)   set 3: [dD | dA_log | ddt_bias]
  if (a.dWin_parts > 0 || a.dK_parts > 0 || rt_parts > 0 || a.head_parts > 0) {
    const int n0 = a.dWin_parts > 0 ? dip * D : 0, n1 = a.dK_parts > 0 ? a.dK_stride : 0;
    const int rs = 2 * Di * D + D, n2 = rt_parts > 0 ? rs : 0, n3 = a.head_parts > 0 ? 3 * nh : 0;
    const int c0 = (n0 + 31) / 32, c1 = (n1 + 31) / 32, c2 = (n2 + 31) / 32, c3 = (n3 + 31) / 32;
    const int el = tid & 31, pl = tid >> 5;
    for (int chunk = blockIdx.x; chunk < c0 + c1 + c2 + c3; chunk += gridDim.x) {
      const float* src;
      int parts, stride, n, e;
      int set;
      if (chunk < c0) { set = 0; src = a.dWin_part; parts = a.dWin_parts; stride = n0; n = n0; e = chunk * 32 + el; }
      else if (chunk < c0 + c1) { set = 1; src = a.dK_part; parts = a.dK_parts; stride = n1; n = n1; e = (chunk - c0) * 32 + el; }
      else if (chunk < c0 + c1 + c2) { set = 2; src = Rt; parts = rt_parts; stride = rs; n = n2; e = (chunk - c0 - c1) * 32 + el; }
      else { set = 3; src = a.head_part; parts = a.head_parts; stride = n3; n = n3; e = (chunk - c0 - c1 - c2) * 32 + el; }
      float v = 0.f;
      if (e < n) {      // up to 19 slabs per thread (<= 148 CTAs / 8 lanes): every load is issued before the first add
        float t[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) {
          const int p = pl + 8 * k;
          t[k] = p < parts ? __ldcg(src + (long long)p * stride + e) : 0.f;   // L2 only: nothing of the slabs may linger in L1
        }
#pragma unroll
        for (int k = 0; k < 20; k += 4) v += (t[k] + t[k + 1]) + (t[k + 2] + t[k + 3]);
      }
      __syncthreads();
      red[pl][el] = v;
      __syncthreads();
      if (pl == 0 && e < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][el];
        if (set == 0) a.dWin[e] = t;
        else if (set == 1) a.dK[e] = t;
        else if (set == 2) Rt[e] = t;
        else if (e < nh) a.dD[e] = t;
        else if (e < 2 * nh) a.dAlog[e - nh] = t;
        else a.ddtb[e - 2 * nh] = t;
      }
    }
    __threadfence();
    __syncthreads();
    // Grid hand-off through a counter.  The launch is an ordinary one, so co-residency of the grid is checked on the host
    // (launch_finalize_fast: occupancy x SMs >= grid) and the wait is BOUNDED anyway: if a block cannot be scheduled next
    // to its peers (MPS / MIG / a co-running kernel holding the SMs) the spin times out after ~2 s and every gradient this
    // block writes is poisoned with NaNs - loud downstream, never a hung GPU.
    __shared__ int s_timeout;
    if (tid == 0) {
      atomicAdd(a.sync_counter, 1);
      int ok = 0;
      for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        if (atomicAdd(a.sync_counter, 0) >= (int)gridDim.x) { ok = 1; break; }
        __nanosleep(spin < 64 ? 20 : 500);
      }
      s_timeout = !ok;
      __threadfence();
    }
    __syncthreads();
    if (s_timeout) poison_late = __int_as_float(0x7fc00000);
    a.dWin_parts = 0;
    a.dK_parts = 0;
    a.head_parts = 0;
    if (s_timeout && tid < 8) du[tid] = __float2bfloat16_rn(poison_late);
  }
  const float poison = poison_late;
  // ---- phase 2
  //   blocks [0, nb_in)            dW_in from the accumulator / the per-CTA slabs of k_bwd4        (32 elements x 8 slab lanes)
  //   blocks [nb_in, nb_in + 2Di)  dW_out, dgamma, dbeta, dalpha1 partial from Rt / sum(dout)    (one warp per column)
  //   remaining blocks             dD, dA_log, ddt_bias and the ten conv weight gradients from dK (shared finalize_body)
  int blk = blockIdx.x;
  if (blk < nb_in) {
    const int n = dip * D, e = blk * 32 + (tid & 31), pl = tid >> 5;
    // phase 2 reads what OTHER SMs wrote in phase 1: __ldcg (L2) loads, never the non-coherent L1
    float v = (e < n && pl == 0 && a.dWin_parts == 0) ? __ldcg(a.dWin + e) : 0.f;
    if (e < n)
      for (int p = pl; p < a.dWin_parts; p += 8) v += a.dWin_part[(long long)p * n + e];
    red[pl][tid & 31] = v;
    __syncthreads();
    if (pl == 0 && e < n && g.in_proj_w) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][tid & 31];
      g.in_proj_w[e] = t + poison;
    }
    return;
  }
  blk -= nb_in;
  if (blk < 2 * Di) {
    if (tid >= 32) return;
    const int c = blk;
    const float a1 = *w.alpha1;
    float dg = 0.f, db = 0.f, da = 0.f;
    for (int d = tid; d < D; d += 32) {
      const float r = __ldcg(Rt + c * D + d), wv = w.out_proj_w[d * 2 * Di + c], sd = __ldcg(sdout + d);
      const float rawv = c < Di ? w.norm_w[c] * r + w.norm_b[c] * sd : r;
      if (g.out_proj_w) g.out_proj_w[d * 2 * Di + c] = a1 * rawv + poison;
      da = fmaf(wv, rawv, da);
      dg = fmaf(wv, r, dg);
      db = fmaf(wv, sd, db);
    }
    dg = warp_sum(dg); db = warp_sum(db); da = warp_sum(da);
    if (tid == 0) {
      if (c < Di) {
        if (g.norm_w) g.norm_w[c] = a1 * dg + poison;
        if (g.norm_b) g.norm_b[c] = a1 * db + poison;
      }
      if (g.alpha1) atomicAdd(g.alpha1, da + poison);     // g.alpha1 is zeroed by the host before this launch
    }
    return;
  }
  blk -= 2 * Di;
  if (blk >= nb_rest) return;      // extra blocks only take part in phase 1
  finalize_body(a, w, g, D, Di, GN, nh, dip, (long long)blk * 256 + tid, (long long)nb_rest * 256, false);
}

// k_finalize_fast hands off between its two phases through a counter: the whole grid must be co-resident.  Verified on the
// host (occupancy query cached per process); the in-kernel wait is bounded as well.
static int check_finalize_grid(int fgrid) {
  static int occ = -1;
  if (occ < 0) ADN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_finalize_fast, 256, 0));
  ADN_REQUIRE(fgrid <= occ * sm_count(), ADN_ERR_SHAPE, "k_finalize_fast: grid %d exceeds the co-resident capacity %d x %d",
              fgrid, occ, sm_count());
  return ADN_OK;
}

// One launch for all per-call weight preparation: conv kernel assembly, in_proj hi/lo split, out_proj -> bf16.
__global__ void k_prep(ConvWeightPtrs cw, float* __restrict__ Kc, int Di, int CC, const float* __restrict__ win,
                       bf16* __restrict__ whi, bf16* __restrict__ wlo, int n_in, const float* __restrict__ wout,
                       bf16* __restrict__ wout_bf, int n_out, bf16* __restrict__ wt_hi, bf16* __restrict__ wt_lo, int n_wt,
                       int D, int dip, bf16* __restrict__ wtf, bf16* __restrict__ wtb, float* __restrict__ zero_f,
                       int n_zero_f, int* __restrict__ zero_status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // row-kernel path: this first kernel of the forward pass also clears the state accumulator and the fault word
  // (two memset nodes less per step)
  if (zero_f != nullptr)
    for (int j = i; j < n_zero_f; j += gridDim.x * blockDim.x) zero_f[j] = 0.f;
  if (zero_status != nullptr && i < 64) zero_status[i] = 0;
  if (wtf != nullptr) rowconv::prep_rowconv(cw, win, wtf, wtb, i);   // conv-as-GEMM weight images (row kernels)
  if (wt_hi != nullptr && i < n_wt) {   // element ((jc*D + d)*8 + q) of the W_in^T image = W_in[jc*8+q][d]
    const int q = i & 7, d = (i >> 3) % D, j = ((i >> 3) / D) * 8 + q;
    const float v = j < dip ? win[j * D + d] : 0.f;
    const bf16 h = __float2bfloat16_rn(v);
    wt_hi[i] = h;
    wt_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  if (i < CC) assemble_conv_channel(cw, Kc + i * 9, Di, i);
  if (i < n_in) {
    const float v = win[i];
    const bf16 h = __float2bfloat16_rn(v);
    whi[i] = h;
    wlo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  if (i < n_out) wout_bf[i] = __float2bfloat16_rn(wout[i]);
}

// ------------------------------------------------------------------------------------------------
// Shared-memory halo-tiled depthwise 3x3 (channels-last bf16).  One CTA = one sample, one 32-channel slab, one
// CT_Y x CT_X spatial tile.  The (CT_Y+2) x (CT_X+2) x 32-channel halo tile is fetched with 16-byte cp.async
// (all requests in flight at once: the memory-level parallelism the register-window kernels lacked), zero-filled
// outside the image.  Thread = (column, 8-channel chunk); it walks down its column with an fp32 register window fed
// by conflict-free LDS.128 and does the 9 taps x 8 channels as 36 FFMA2.
// ------------------------------------------------------------------------------------------------
constexpr int CT_X = 32, CT_Y = 16, CT_XH = CT_X + 2, CT_YH = CT_Y + 2;
constexpr int CB_Y = 32;   // tile height of the conv backward (fewer dK reductions, less halo)

// g: TL tensor with NCH chunks; tok_base = first token of the sample; c0 = first channel of the 32-channel slab
template <int TY>
__device__ __forceinline__ void conv_tile_load(uint4* sdst, const bf16* __restrict__ g, int NCH, int tok_base, int c0,
                                               int H, int W, int y0, int x0, int halo, int tid, int nthreads) {
  // tile rows y0-halo .. y0+TY-1+halo, cols x0-halo .. ; 4 chunks of 8 channels per token
  const int TH = TY + 2 * halo, TW = CT_X + 2 * halo;
  for (int i = tid; i < TH * TW * 4; i += nthreads) {
    const int ch = i & 3, c = (i >> 2) % TW, r = (i >> 2) / TW;
    const int y = y0 - halo + r, x = x0 - halo + c;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    const bf16* src = g + tl_off32(tok_base + (ok ? y * W + x : 0), c0 + ch * 8, NCH);
    cp_async16(sdst + i, src, ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(128)
k_conv_fwd_tile(const bf16* __restrict__ raw, int ldr, const float* __restrict__ Kc, bf16* __restrict__ pre,
                bf16* __restrict__ act, int H, int W, int CC, int slabs) {
  __shared__ uint4 tile[CT_YH * CT_XH * 4];
  const int tid = threadIdx.x;
  const int b = blockIdx.z / slabs, slab = blockIdx.z % slabs;
  const int x0 = blockIdx.x * CT_X, y0 = blockIdx.y * CT_Y, c0 = slab * 32;
  conv_tile_load<CT_Y>(tile, raw, ldr >> 3, b * H * W, c0, H, W, y0, x0, 1, tid, 128);
  cp_async_commit();
  const int xl = tid >> 2, ch = tid & 3, cc = c0 + ch * 8;
  float2 k2[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int p = 0; p < 4; ++p) k2[t][p] = make_float2(__ldg(Kc + (cc + 2 * p) * 9 + t), __ldg(Kc + (cc + 2 * p + 1) * 9 + t));
  cp_async_wait<0>();
  __syncthreads();
  const int x = x0 + xl;
  if (x >= W) return;
  float2 win[3][3][4];
  auto load_row = [&](int r, float2 (&w3)[3][4]) {
#pragma unroll
    for (int s = 0; s < 3; ++s) unpack8_f2(tile[(r * CT_XH + xl + s) * 4 + ch], w3[s]);
  };
  load_row(0, win[0]);
  load_row(1, win[1]);
  const int ny = min(CT_Y, H - y0);
  for (int rb = 0; rb < ny; rb += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int r = rb + ph;
      if (r < ny) {
        load_row(r + 2, win[(ph + 2) % 3]);
        float2 a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx)
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = __ffma2_rn(k2[rr * 3 + sx][q], win[(ph + rr) % 3][sx][q], a[q]);
        const long long off = tl_off32((b * H + y0 + r) * W + x, cc, CC >> 3);
        // On this path the "pre" slot of the saved tensors holds SiLU'(pre) = s + SiLU(pre)*(1 - s), s = sigmoid(pre):
        // the backward only ever needs that factor, which removes every exp / divide from the conv backward.
        float2 gq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 sg = make_float2(sigmoid_fast(a[q].x), sigmoid_fast(a[q].y));
          a[q] = __ffma2_rn(a[q], sg, make_float2(0.f, 0.f));
          const float2 om = __ffma2_rn(sg, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
          gq[q] = __ffma2_rn(a[q], om, sg);
        }
        if (pre) *reinterpret_cast<uint4*>(pre + off) = pack8_f2(gq);
        *reinterpret_cast<uint4*>(act + off) = pack8_f2(a);
      }
    }
  }
}

// conv backward, tiled.  dpre = dact * SiLU'(pre) is formed ONCE per element in the shared-memory halo tile (pre is
// streamed straight from global by the thread that fetched the matching dact chunk), then ONE pass per output row does
//   draw[y,x]  = sum_ab K[a][b] dpre[y-a+1][x-b+1]      and      dK[a][b] += raw[y,x] * dpre[y-a+1][x-b+1]
// sharing the unpacked register window.  256 threads: thread = (column, 4-channel half chunk) so that window (18) +
// taps (18) + kernel-gradient accumulators (18 float2) fit in registers.  dK is reduced by shuffles + shared-memory
// atomics, then one global atomic per (channel, tap) per CTA.
__device__ __forceinline__ void unpack4_f2(const uint2& u, float2 (&v)[2]) {
  v[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
  v[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

__global__ void __launch_bounds__(256, 2)
k_conv_bwd_tile(const bf16* __restrict__ dact, const bf16* __restrict__ sgrad, const bf16* __restrict__ raw, int ldr,
                const float* __restrict__ Kc, bf16* __restrict__ draw, float* __restrict__ dK, int H, int W, int CC,
                int slabs) {
  extern __shared__ __align__(16) uint8_t csm[];
  uint4* tD = reinterpret_cast<uint4*>(csm);   // dact -> dpre, (CB_Y+2) x (CT_X+2) halo tile
  __shared__ float red[32 * 9];
  const int tid = threadIdx.x;
  PhaseTimer pt(3);
  const int b = blockIdx.z / slabs, slab = blockIdx.z % slabs;
  const int x0 = blockIdx.x * CT_X, y0 = blockIdx.y * CB_Y, c0 = slab * 32;
  const int boff = b * H * W;
  conv_tile_load<CB_Y>(tD, dact, CC >> 3, boff, c0, H, W, y0, x0, 1, tid, 256);
  cp_async_commit();
  for (int i = tid; i < 32 * 9; i += 256) red[i] = 0.f;
  const int xl = tid >> 3, hc = tid & 7, cc = c0 + hc * 4;   // hc: 4-channel half chunk inside the 32-channel slab
  float2 k2[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int p = 0; p < 2; ++p) k2[t][p] = make_float2(__ldg(Kc + (cc + 2 * p) * 9 + t), __ldg(Kc + (cc + 2 * p + 1) * 9 + t));
  // each thread multiplies exactly the chunks it fetched itself: only its own cp.async group has to be complete
  pt.mark(0);
  cp_async_wait<0>();
  pt.mark(1);
  {
    // sgrad is streamed from global in batches of PB independent loads per thread (latency overlap), then multiplied in
    constexpr int PB = 6, NEL = (CB_Y + 2) * CT_XH * 4;
    for (int base = tid; base < NEL; base += 256 * PB) {
      uint4 sv[PB];
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const int i = base + u * 256;
        const int chn = i & 3, c = (i >> 2) % CT_XH, r = (i >> 2) / CT_XH;
        const int y = y0 - 1 + r, x = x0 - 1 + c;
        const bool okp = i < NEL && y >= 0 && y < H && x >= 0 && x < W;
        sv[u] = okp ? __ldg(reinterpret_cast<const uint4*>(sgrad + tl_off32(boff + y * W + x, c0 + chn * 8, CC >> 3)))
                    : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const int i = base + u * 256;
        if (i < NEL) {
          float2 g[4], sg[4];
          unpack8_f2(tD[i], g);
          unpack8_f2(sv[u], sg);
#pragma unroll
          for (int q = 0; q < 4; ++q) g[q] = __ffma2_rn(g[q], sg[q], make_float2(0.f, 0.f));
          tD[i] = pack8_f2(g);       // out-of-image elements: zero-filled tile x zero = zero
        }
      }
    }
  }
  pt.mark(2);
  __syncthreads();
  pt.mark(3);
  const int x = x0 + xl;
  const int ny = min(CB_Y, H - y0);
  const bool xin = x < W;   // columns beyond the image compute on zero-filled tile data and never store
  const uint2* tD2 = reinterpret_cast<const uint2*>(tD);
  float2 dk[9][2], win[3][3][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) { dk[t][0] = make_float2(0.f, 0.f); dk[t][1] = make_float2(0.f, 0.f); }
  auto load_row = [&](int r, float2 (&w3)[3][2]) {
#pragma unroll
    for (int s = 0; s < 3; ++s) unpack4_f2(tD2[(r * CT_XH + xl + s) * 8 + hc], w3[s]);
  };
  load_row(0, win[0]);
  load_row(1, win[1]);
  const int NR = ldr >> 3;
  const int rtok = boff + y0 * W + (xin ? x : 0);   // raw is TL: rows of one column are W tokens apart
  const uint2 zero2 = make_uint2(0u, 0u);
  // the centre raw value of row r is fetched three rows ahead (ring of 3 registers matching the 3-way unroll)
  uint2 rn[3];
#pragma unroll
  for (int q = 0; q < 3; ++q)
    rn[q] = (xin && q < ny) ? __ldg(reinterpret_cast<const uint2*>(raw + tl_off32(rtok + q * W, cc, NR))) : zero2;
  for (int rb = 0; rb < ny; rb += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int r = rb + ph;
      if (r < ny) {
        load_row(r + 2, win[(ph + 2) % 3]);
        float2 rc[2], o[2];
        unpack4_f2(rn[ph], rc);
        rn[ph] = (xin && r + 3 < ny) ? __ldg(reinterpret_cast<const uint2*>(raw + tl_off32(rtok + (r + 3) * W, cc, NR))) : zero2;
        o[0] = make_float2(0.f, 0.f);
        o[1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float2 dv = win[(ph + 2 - a) % 3][2 - bb][q];
              o[q] = __ffma2_rn(k2[a * 3 + bb][q], dv, o[q]);
              dk[a * 3 + bb][q] = __ffma2_rn(rc[q], dv, dk[a * 3 + bb][q]);
            }
        if (xin) {
          uint2 ov;
          ov.x = pack_bf16(o[0].x, o[0].y);
          ov.y = pack_bf16(o[1].x, o[1].y);
          *reinterpret_cast<uint2*>(draw + tl_off32(boff + (y0 + r) * W + x, cc, NR)) = ov;
        }
      }
    }
  }
  pt.mark(4);
  // reduce over the columns: lanes of a warp = 4 columns x 8 half chunks -> xor-shuffle over the two column bits
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float vx = dk[t][q].x, vy = dk[t][q].y;
#pragma unroll
      for (int o = 8; o < 32; o <<= 1) { vx += __shfl_xor_sync(0xffffffffu, vx, o); vy += __shfl_xor_sync(0xffffffffu, vy, o); }
      if ((tid & 31) < 8) {
        atomicAdd(&red[(hc * 4 + 2 * q) * 9 + t], vx);
        atomicAdd(&red[(hc * 4 + 2 * q + 1) * 9 + t], vy);
      }
    }
  __syncthreads();
  for (int i = tid; i < 32 * 9; i += 256)
    if (red[i] != 0.f) atomicAdd(dK + c0 * 9 + i, red[i]);
  pt.mark(5);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Per-call weight preparation (k_prep).  During training it lives at the end of the `saved` buffer, so the backward pass
// reuses what the forward prepared (the weights cannot change in between); for inference it lives in the workspace.
struct PrepBufs {
  float* Kc;             // [CC][9] assembled conv kernels
  bf16 *Whi, *Wlo, *Wout;
  bf16 *WT_hi, *WT_lo;   // W_in^T as a K-major B operand image: [ceil(dip/128)*16 chunks][D][8], zero padded
  bf16 *WtF, *WtB;       // conv-as-GEMM weight images of the row kernels (adnssd_rowconv.cuh)
  int wt_chunks;
  size_t bytes;
  PrepBufs(const MixerDims& d, void* p) {
    Carver c(p);
    Kc = c.take<float>((size_t)d.CC * 9);
    Whi = c.take<bf16>((size_t)d.dip * d.D);
    Wlo = c.take<bf16>((size_t)d.dip * d.D);
    Wout = c.take<bf16>((size_t)d.D * 2 * d.Di);
    wt_chunks = 16 * cdiv(d.dip, 128);
    WT_hi = c.take<bf16>((size_t)wt_chunks * d.D * 8);
    WT_lo = c.take<bf16>((size_t)wt_chunks * d.D * 8);
    WtF = c.take<bf16>((size_t)rowconv::WTF_B / 2);
    WtB = c.take<bf16>((size_t)rowconv::WTB_B / 2);
    bytes = c.off;
  }
};

struct FastWs {           // placed after the generic workspace of the same pass
  float* dWin_part;      // [148][dip*D] per-CTA partial sums of dW_in (k_bwd4 / k_bconv_wg)
  float* dK_part;        // [148][CC*9] per-CTA partial sums of dK (k_bconv_wg)
  float *Rt, *sdout;     // k_bwd1 accumulators: Rt[2Di][D], sdout[D] (contiguous, zeroed together); slabs on the ws path
  float* head_part;      // [148][3*nh] per-CTA partial sums of dD, dA_log, ddt_bias (k_bwd2_ws)
  int* status;
  size_t bytes;
  FastWs(const MixerDims& d, void* p) {
    Carver c(p);
    dWin_part = c.take<float>((size_t)sm_count() * d.dip * d.D);
    dK_part = c.take<float>((size_t)sm_count() * d.CC * 9);
    Rt = c.take<float>((size_t)sm_count() * (2 * d.Di * d.D + d.D));      // one slab per CTA on the warp-specialised path
    sdout = Rt ? Rt + (size_t)2 * d.Di * d.D : nullptr;
    head_part = c.take<float>((size_t)sm_count() * 3 * d.nh);
    status = c.take<int>(64);
    bytes = c.off;
  }
};

// shapes served by the conv-as-GEMM row kernels: the full-resolution refiner mixers at 128-token-wide grids
// rows per CTA of the row kernels: one CTA per SM by default; ADN_ROWS_PER_CTA overrides (diagnostics)
static int rows_per_cta(int rows_total, int ctas) {
  // the override may only coarsen the split: the per-CTA slab buffers are sized for `ctas` CTAs
  if (env().rows_per_cta > 0 && cdiv(rows_total, env().rows_per_cta) <= ctas) return env().rows_per_cta;
  return cdiv(rows_total, ctas);
}

static bool rowconv_supported(const MixerDims& d) {
  if (!env().rowconv) return false;             // diagnostics: ADN_ROWCONV=0 keeps these shapes on the tile kernels
  const bool wide = env().row_wide;             // diagnostics: ADN_ROW_WIDE=0 restricts the row kernels to W == 128
  return d.D == 32 && d.Di == 64 && d.P == 4 && d.GN == 32 && d.dip == 208 && d.ldr == 208 &&
         (d.W == 128 || (wide && d.W % 128 == 0 && d.W <= 1024));
}

bool sm100_rowconv(const MixerDims& d) { return rowconv_supported(d); }

bool sm100_supported(const MixerDims& d) {
  // instantiated tile shapes: d_model 32 (d_inner 64), headdim 4, ngroups*d_state in {32, 64, 128} (d_state 16, 32, 64)
  return d.D == 32 && d.Di == 64 && d.P == 4 && (d.GN == 32 || d.GN == 64 || d.GN == 128) && d.dip % 16 == 0 && d.dip <= 512 && d.ldr == d.dip && d.CC % 32 == 0 && d.L % 128 == 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  ADN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return ADN_OK;
}

template <int DI, int GN>
static int launch_state(const MixerDims& d, const bf16* act, const bf16* raw, const AdnWeights& w, float* S, int* status,
                        cudaStream_t st) {
  constexpr size_t smem = 2 * (size_t)(16 + GN / 8 + DI / 32) * 128 * 8 * sizeof(bf16);
  int rc = set_smem(k_state<DI, GN>, smem);
  if (rc) return rc;
  const int tpb = cdiv(d.L, 128);
  const int cpb = max(1, min(tpb, cdiv(sm_count() * 2, d.B)));
  { ADN_KERNEL("k_state", st); k_state<DI, GN><<<d.B * cpb, 128, smem, st>>>(act, raw, d.ldr, d.CC, w.dt_bias, w.A_log, S, d.L, tpb, cpb, status); }
  return ADN_OK;
}

template <int DI, int GN>
static int launch_readout(const MixerDims& d, const bf16* act, const float* S, const AdnWeights& w, const bf16* Wout,
                          bf16* out, int* status, cudaStream_t st) {
  constexpr size_t smem = ((size_t)(GN / 8 + 2 * DI / 8 + DI / 8) * 128 * 8 + 2 * (GN / 8) * DI * 8 + (2 * DI / 8) * (DI / 2) * 8) * sizeof(bf16);
  int rc = set_smem(k_readout<DI, GN>, smem);
  if (rc) return rc;
  const int tpb = cdiv(d.L, 128), nt = tpb * d.B;
  const int per_sm = smem > 110 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : 3);
  { ADN_KERNEL("k_readout", st); k_readout<DI, GN><<<min(nt, sm_count() * per_sm), 128, smem, st>>>(act, d.CC, S, w.D, w.norm_w, w.norm_b, w.alpha1, Wout, out, d.L, tpb, nt, status); }
  return ADN_OK;
}

size_t sm100_saved_extra_bytes(const MixerDims& d) { return sm100_supported(d) ? PrepBufs(d, nullptr).bytes : 0; }

void sm100_workspace_bytes(const MixerDims& d, size_t* f, size_t* b) {
  size_t extra = FastWs(d, nullptr).bytes + PrepBufs(d, nullptr).bytes;
  *f = FwdWs<bf16, TWf>(d, nullptr).bytes + extra;
  *b = BwdWs<bf16, TWf>(d, nullptr).bytes + extra;
}

template <int D>
static int launch_inproj(const MixerDims& d, const bf16* u, const PrepBufs& P, const FastWs& F, bf16* raw, cudaStream_t st) {
  const int num_tiles = cdiv(d.T, 128);
  const size_t smem = (size_t)(2 * 256 * D + 2 * 128 * D) * sizeof(bf16);
  ADN_CHECK_CUDA(cudaFuncSetAttribute(k_inproj<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(num_tiles, sm_count() * 2);
  { ADN_KERNEL("k_inproj", st); k_inproj<D><<<grid, 128, smem, st>>>(u, P.Whi, P.Wlo, raw, d.ldr, d.dip, d.T, num_tiles, F.status); }
  return ADN_OK;
}

static inline void split_tiles(int num_tiles, int per_sm, int* grid, int* per_cta) {
  const int target = max(1, min(num_tiles, sm_count() * per_sm));
  *per_cta = cdiv(num_tiles, target);
  *grid = cdiv(num_tiles, *per_cta);
}

template <int DI, int GN>
static int launch_bwd1(const MixerDims& d, const bf16* dout, const bf16* act, const float* S, const AdnWeights& w,
                       const PrepBufs& P, const FastWs& F, bf16* dact, float* dS, cudaStream_t st, const bf16* sgrad) {
  constexpr int D = DI / 2;
  constexpr size_t smem = ((size_t)(D / 8 + GN / 8 + 32) * 128 * 8 + (D / 8) * 2 * DI * 8 + 2 * (GN / 8) * DI * 8 + 2 * (DI / 8) * GN * 8) * sizeof(bf16);
  int rc = set_smem(k_bwd1<DI, GN>, smem);
  if (rc) return rc;
  const int tpb = cdiv(d.L, 128), nt = tpb * d.B;
  int grid, per;
  split_tiles(nt, smem > 110 * 1024 ? 1 : 2, &grid, &per);
  { ADN_KERNEL("k_bwd1", st); k_bwd1<DI, GN><<<grid, 128, smem, st>>>(dout, act, d.CC, S, w.D, w.norm_w, w.alpha1, P.Wout, dact, F.Rt, F.sdout, dS, d.L, tpb, nt, per, F.status, sgrad); }
  return ADN_OK;
}

template <int DI, int GN>
static int launch_bwd2(const MixerDims& d, const bf16* act, const bf16* raw, const float* dS, const AdnWeights& w,
                       const FastWs& F, bf16* dact, bf16* draw, const GradAcc& acc, cudaStream_t st, const bf16* sgrad,
                       int dt_nch, int dt_c0) {
  constexpr size_t smem = (2 * (size_t)(GN / 8 + 2 * (DI / 8) + DI / 32) * 128 * 8 + 2 * (GN / 8) * DI * 8 + 2 * (DI / 8) * GN * 8) * sizeof(bf16);
  static_assert(smem <= 227 * 1024, "k_bwd2 stages do not fit shared memory");
  int rc = set_smem(k_bwd2<DI, GN>, smem);
  if (rc) return rc;
  const int tpb = cdiv(d.L, 128), nt = tpb * d.B;
  int grid, per;
  split_tiles(nt, smem > 110 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : 3), &grid, &per);
  { ADN_KERNEL("k_bwd2", st); k_bwd2<DI, GN><<<grid, 128, smem, st>>>(act, raw, d.ldr, d.CC, dS, w.dt_bias, w.A_log, w.D, dact, draw, acc.dD, acc.dAlog, acc.ddtb, d.L, tpb, nt, per, F.status, sgrad, dt_nch, dt_c0); }
  return ADN_OK;
}

template <int D, int MB, int NS>
static int launch_bwd4(const MixerDims& d, const bf16* draw, const bf16* u, const PrepBufs& P, const FastWs& F, bf16* du, GradAcc* acc,
                       cudaStream_t st) {
  constexpr size_t smem = (NS * (size_t)(16 * MB + D / 8) * 128 * 8 + 2 * 16 * MB * D * 8) * sizeof(bf16);
  static_assert(smem <= 227 * 1024, "k_bwd4 stages do not fit shared memory");
  int rc = set_smem(k_bwd4<D, MB, NS>, smem);
  if (rc) return rc;
  const int nt = cdiv(d.T, 128);
  int grid, per;
  split_tiles(nt, 1, &grid, &per);      // one CTA per SM (<= 148 partial slabs)
  { ADN_KERNEL("k_bwd4", st); k_bwd4<D, MB, NS><<<grid, 128, smem, st>>>(draw, d.ldr, d.dip, u, P.WT_hi, P.WT_lo, du, F.dWin_part, d.T, nt, per, F.status); }
  acc->dWin_part = F.dWin_part;
  acc->dWin_parts = grid;
  return ADN_OK;
}

int sm100_forward(const MixerDims& d, const AdnWeights& w, const bf16* u, bf16* out, void* saved, void* ws,
                  cudaStream_t st) {
  typedef bf16 T;
  FwdWs<T, TWf> W(d, ws);
  FastWs F(d, (char*)ws + W.bytes);
  SavedBufs<T> S = saved ? SavedBufs<T>(d, saved) : W.tmp;
  const bool training = saved != nullptr;
  PrepBufs P(d, training ? (char*)saved + S.bytes : (char*)ws + W.bytes + F.bytes);
  const long long Tt = d.T;
  {
    const int n_in = d.dip * d.D, n_out = d.D * 2 * d.Di, n_wt = P.wt_chunks * d.D * 8;
    const bool rows = rowconv_supported(d);
    if (!rows) ADN_CHECK_CUDA(cudaMemsetAsync(F.status, 0, 256, st));
    const int n = max(max(max(max(n_in, n_out), d.CC), n_wt), rows ? 9 * rowconv::DIP * rowconv::D : 0);
    { ADN_KERNEL("k_prep", st); k_prep<<<cdiv(n, 256), 256, 0, st>>>(conv_ptrs(w), P.Kc, d.Di, d.CC, w.in_proj_w, P.Whi, P.Wlo, n_in, w.out_proj_w, P.Wout, n_out, P.WT_hi, P.WT_lo, n_wt, d.D, d.dip, rows ? P.WtF : nullptr, P.WtB, rows ? S.S : nullptr, d.B * d.GN * d.Di, rows ? F.status : nullptr); }
  }
  if (rowconv_supported(d)) {
    // (1)-(4a) fused: in_proj + conv + SiLU + decay weights + state, one image row per step (adnssd_rowconv.cuh).
    // The `raw` slot of the saved tensors only holds the dt columns, as a TL tensor with 2 chunks per tile, and the `wdec`
    // slot (T x nh floats = T x D bf16 for headdim 4) holds a TL copy of u for the bulk copies of the backward pass.
    static_assert(sizeof(float) * rowconv::NH == sizeof(bf16) * rowconv::D, "u_tl does not fit the wdec slot");
    int rc = set_smem(rowconv::k_fconv, rowconv::FC_SMEM);
    if (rc) return rc;
    const int rows_total = d.B * d.H * (d.W / 128), per = rows_per_cta(rows_total, sm_count()), grid = cdiv(rows_total, per);   // strip rows
    { ADN_KERNEL("k_fconv", st); rowconv::k_fconv<<<grid, rowconv::FC_THREADS, rowconv::FC_SMEM, st>>>(u, P.WtF, w.dt_bias, w.A_log, S.act, training ? S.pre : nullptr, S.raw, S.S, d.H, rows_total, per, F.status, training ? reinterpret_cast<bf16*>(S.wdec) : nullptr, d.W / 128); }
    // (A warp-specialised readout, one CTA per SM with two tiles in flight, measured SLOWER than this monolithic tile
    // kernel at three CTAs per SM - 41.7 vs 33.2 us at the benchmark shape - and was removed in round 2.)
    rc = launch_readout<64, 32>(d, S.act, S.S, w, P.Wout, out, F.status, st);
    if (rc) return rc;
    ADN_CHECK_LAUNCH();
    return ADN_OK;
  }
  // (1) in_proj on tcgen05
  int rc = d.D == 16 ? launch_inproj<16>(d, u, P, F, S.raw, st) : d.D == 32 ? launch_inproj<32>(d, u, P, F, S.raw, st)
                                                                         : launch_inproj<64>(d, u, P, F, S.raw, st);
  if (rc) return rc;
  // (2) depthwise 3x3 + SiLU
  {
    const int slabs = d.CC / 32;
    dim3 grid(cdiv(d.W, CT_X), cdiv(d.H, CT_Y), d.B * slabs);
    { ADN_KERNEL("k_conv_fwd_tile", st); k_conv_fwd_tile<<<grid, 128, 0, st>>>(S.raw, d.ldr, P.Kc, training ? S.pre : nullptr, S.act, d.H, d.W, d.CC, slabs); }
  }
  (void)Tt;
  // (4a) state on tcgen05 (reduction over tokens), (4b)+(5) readout + LayerNorm + out_proj on tcgen05
  ADN_CHECK_CUDA(cudaMemsetAsync(S.S, 0, (size_t)d.B * d.GN * d.Di * sizeof(float), st));
  rc = d.GN == 32 ? launch_state<64, 32>(d, S.act, S.raw, w, S.S, F.status, st)
       : d.GN == 64 ? launch_state<64, 64>(d, S.act, S.raw, w, S.S, F.status, st)
                    : launch_state<64, 128>(d, S.act, S.raw, w, S.S, F.status, st);
  if (rc) return rc;
  rc = d.GN == 32 ? launch_readout<64, 32>(d, S.act, S.S, w, P.Wout, out, F.status, st)
       : d.GN == 64 ? launch_readout<64, 64>(d, S.act, S.S, w, P.Wout, out, F.status, st)
                    : launch_readout<64, 128>(d, S.act, S.S, w, P.Wout, out, F.status, st);
  if (rc) return rc;
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

int sm100_backward(const MixerDims& d, const AdnWeights& w, const bf16* u, const void* saved, const bf16* dout,
                   bf16* du, const AdnWeightGrads& g, void* ws, cudaStream_t st) {
  typedef bf16 T;
  BwdWs<T, TWf> W(d, ws);
  FastWs F(d, (char*)ws + W.bytes);
  SavedBufs<T> S(d, const_cast<void*>(saved));
  PrepBufs P(d, (char*)const_cast<void*>(saved) + S.bytes);   // prepared by the forward pass of this step
  ADN_CHECK_CUDA(cudaMemsetAsync(W.zero_begin, 0, W.zero_bytes, st));     // accumulators, dS', hand-off counter, W.status
  // ---- phase B1: dout -> dy, dzc, dCc ; reductions Rt, dS'
  if (rowconv_supported(d)) {
    // B1 / B2 write dpre = dact * SiLU'(pre) directly; ddt goes to a compact TL tensor (2 chunks per tile) in W.draw
    int rc, rt_parts = 0;
    const bool ws_path = env().bwd_ws;            // diagnostics: ADN_BWD_WS=0 keeps the monolithic tile kernels
    // warp-specialised path: ONE memset per backward pass (the fault word lives in the zeroed region, Rt / sum(dout) are
    // slabs that are fully overwritten, and k_bwd1_ws clears the alpha1 gradient that k_finalize_fast accumulates into)
    int* status = ws_path ? W.status : F.status;
    if (!ws_path) {
      ADN_CHECK_CUDA(cudaMemsetAsync(F.Rt, 0, ((size_t)2 * d.Di * d.D + d.D) * sizeof(float), st));
      ADN_CHECK_CUDA(cudaMemsetAsync(F.status, 0, 256, st));
    }
    if (!ws_path) {
      rc = launch_bwd1<64, 32>(d, dout, S.act, S.S, w, P, F, W.dact, W.dS, st, S.pre);
      if (rc) return rc;
      rc = launch_bwd2<64, 32>(d, S.act, S.raw, W.dS, w, F, W.dact, W.draw, W.acc, st, S.pre, 2, 0);
      if (rc) return rc;
    } else {
      const int tpb = d.L / 128, nt = tpb * d.B, per = cdiv(nt, sm_count()), grid = cdiv(nt, per);
      rc = set_smem(bwdws::k_bwd1_ws, bwdws::B1_SMEM);
      if (rc) return rc;
      { ADN_KERNEL("k_bwd1_ws", st); bwdws::k_bwd1_ws<<<grid, bwdws::WS_THREADS, bwdws::B1_SMEM, st>>>(dout, S.act, S.pre, S.S, w.D, w.norm_w, w.alpha1, P.Wout, W.dact, F.Rt, F.sdout, W.dS, tpb, nt, per, status, g.alpha1); }
      rc = set_smem(bwdws::k_bwd2_ws, bwdws::B2_SMEM);
      if (rc) return rc;
      { ADN_KERNEL("k_bwd2_ws", st); bwdws::k_bwd2_ws<<<grid, 320, bwdws::B2_SMEM, st>>>(S.act, S.pre, S.raw, W.dS, w.dt_bias, w.A_log, w.D, W.dact, W.draw, F.head_part, tpb, nt, per, status); }
      rt_parts = grid;
      W.acc.head_part = F.head_part;
      W.acc.head_parts = grid;
    }
    const int TPR = d.W / 128, rows_total = d.B * d.H * TPR;      // rows of the 128-wide strip images
    {
      rc = set_smem(rowconv::k_bconv_du, rowconv::DU_SMEM);
      if (rc) return rc;
      const int per = rows_per_cta(rows_total, sm_count()), grid = cdiv(rows_total, per);
      { ADN_KERNEL("k_bconv_du", st); rowconv::k_bconv_du<<<grid, 192, rowconv::DU_SMEM, st>>>(W.dact, W.draw, P.WtB, du, d.H, rows_total, per, status, env().du_dbg, TPR); }
      if (TPR > 1) {
        const int n_edges = d.B * d.H * (TPR - 1);
        { ADN_KERNEL("k_bconv_du_edge", st); rowconv::k_bconv_du_edge<<<cdiv(2LL * n_edges * 32, 256), 256, 0, st>>>(W.dact, P.Kc, w.in_proj_w, du, d.H, TPR, n_edges); }
      }
    }
    {
      rc = set_smem(rowconv::k_bconv_wg, rowconv::WG_SMEM);
      if (rc) return rc;
      const int cpb = max(1, min(sm_count() / 2, rows_total)), per = cdiv(rows_total, cpb), parts = cdiv(rows_total, per);
      { ADN_KERNEL("k_bconv_wg", st); rowconv::k_bconv_wg<<<2 * parts, 192, rowconv::WG_SMEM, st>>>(W.dact, W.draw, reinterpret_cast<const bf16*>(S.wdec), w.in_proj_w, P.Kc, F.dK_part, F.dWin_part, d.H, rows_total, per, parts, status, TPR, W.acc.sync_counter + 16); }
      W.acc.dWin_part = F.dWin_part;
      W.acc.dWin_parts = parts;
      W.acc.dK_part = F.dK_part;
      W.acc.dK_parts = parts;
      W.acc.dK_stride = d.CC * 9;
    }
    if (g.alpha1 && !ws_path) ADN_CHECK_CUDA(cudaMemsetAsync(g.alpha1, 0, sizeof(float), st));
    const int nb_in = cdiv(d.dip * d.D, 32), nb_rest = 8;
    // every block of this grid (344 x 256 threads at the benchmark shape, < 3 blocks per SM) is co-resident, which the
    // counter hand-off between the two phases relies on; a 148 x 6 grid for phase 1 measured no faster
    const int fgrid = nb_in + 2 * d.Di + nb_rest;
    rc = check_finalize_grid(fgrid);
    if (rc) return rc;
    { ADN_KERNEL("k_finalize_fast", st); k_finalize_fast<<<fgrid, 256, 0, st>>>(W.acc, w, g, F.Rt, F.sdout, d.D, d.Di, d.GN, d.nh, d.dip, nb_in, nb_rest, rt_parts, status, du); }
    ADN_CHECK_LAUNCH();
    return ADN_OK;
  }
  ADN_CHECK_CUDA(cudaMemsetAsync(F.Rt, 0, ((size_t)2 * d.Di * d.D + d.D) * sizeof(float), st));
  ADN_CHECK_CUDA(cudaMemsetAsync(F.status, 0, 256, st));
  int rc = d.GN == 32 ? launch_bwd1<64, 32>(d, dout, S.act, S.S, w, P, F, W.dact, W.dS, st, nullptr)
           : d.GN == 64 ? launch_bwd1<64, 64>(d, dout, S.act, S.S, w, P, F, W.dact, W.dS, st, nullptr)
                        : launch_bwd1<64, 128>(d, dout, S.act, S.S, w, P, F, W.dact, W.dS, st, nullptr);
  if (rc) return rc;
  // ---- phase B2: dS' -> dxc, dBc, ddt
  rc = d.GN == 32 ? launch_bwd2<64, 32>(d, S.act, S.raw, W.dS, w, F, W.dact, W.draw, W.acc, st, nullptr, d.ldr >> 3, d.CC >> 3)
       : d.GN == 64 ? launch_bwd2<64, 64>(d, S.act, S.raw, W.dS, w, F, W.dact, W.draw, W.acc, st, nullptr, d.ldr >> 3, d.CC >> 3)
                    : launch_bwd2<64, 128>(d, S.act, S.raw, W.dS, w, F, W.dact, W.draw, W.acc, st, nullptr, d.ldr >> 3, d.CC >> 3);
  if (rc) return rc;
  // ---- conv backward (dpre formed in shared memory; transposed conv + kernel gradient)
  {
    const int slabs = d.CC / 32;
    const size_t smem = (size_t)(CB_Y + 2) * CT_XH * 4 * 16;
    rc = set_smem(k_conv_bwd_tile, smem);
    if (rc) return rc;
    dim3 grid(cdiv(d.W, CT_X), cdiv(d.H, CB_Y), d.B * slabs);
    { ADN_KERNEL("k_conv_bwd_tile", st); k_conv_bwd_tile<<<grid, 256, smem, st>>>(W.dact, S.pre, S.raw, d.ldr, P.Kc, W.draw, W.acc.dK, d.H, d.W, d.CC, slabs); }
  }
  // ---- in_proj backward
  rc = d.dip <= 128 ? launch_bwd4<32, 1, 2>(d, W.draw, u, P, F, du, &W.acc, st)
       : d.dip <= 256 ? launch_bwd4<32, 2, 2>(d, W.draw, u, P, F, du, &W.acc, st)
                      : launch_bwd4<32, 4, 1>(d, W.draw, u, P, F, du, &W.acc, st);   // wide in_proj: one stage fits
  if (rc) return rc;
  {
    if (g.alpha1) ADN_CHECK_CUDA(cudaMemsetAsync(g.alpha1, 0, sizeof(float), st));
    const int nb_in = cdiv(d.dip * d.D, 32), nb_rest = 8;
    rc = check_finalize_grid(nb_in + 2 * d.Di + nb_rest);
    if (rc) return rc;
    { ADN_KERNEL("k_finalize_fast", st); k_finalize_fast<<<nb_in + 2 * d.Di + nb_rest, 256, 0, st>>>(W.acc, w, g, F.Rt, F.sdout, d.D, d.Di, d.GN, d.nh, d.dip, nb_in, nb_rest, 0, F.status, du); }
  }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

}  // namespace adn

extern "C" int adn_phase_enable(int on) {
  using namespace adn;
  unsigned long long zero[64] = {0};
  ADN_CHECK_CUDA(cudaMemcpyToSymbol(g_phase, zero, sizeof(zero)));
  ADN_CHECK_CUDA(cudaMemcpyToSymbol(g_phase_on, &on, sizeof(int)));
  return ADN_OK;
}
extern "C" int adn_cta_times_read(unsigned long long* out960) {
  using namespace adn;
  ADN_REQUIRE(out960 != nullptr, ADN_ERR_NULL, "adn_cta_times_read: NULL");
  ADN_CHECK_CUDA(cudaDeviceSynchronize());
  ADN_CHECK_CUDA(cudaMemcpyFromSymbol(out960, g_cta_t, 3 * 160 * 4 * sizeof(unsigned long long)));
  return ADN_OK;
}
extern "C" int adn_phase_read(unsigned long long* out64) {
  using namespace adn;
  ADN_REQUIRE(out64 != nullptr, ADN_ERR_NULL, "adn_phase_read: NULL");
  ADN_CHECK_CUDA(cudaDeviceSynchronize());
  ADN_CHECK_CUDA(cudaMemcpyFromSymbol(out64, g_phase, 64 * sizeof(unsigned long long)));
  return ADN_OK;
}

extern "C" int adn_selftest_umma(int mode, int N, int K, const void* A, const void* B, float* C, int* status, void* stream) {
  using namespace adn;
  ADN_REQUIRE(A && B && C && status, ADN_ERR_NULL, "adn_selftest_umma: NULL argument");
  ADN_REQUIRE((mode == 0 || mode == 1) && N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256,
              ADN_ERR_SHAPE, "adn_selftest_umma: mode in {0,1}, N,K multiples of 16 in [16,256]");
  size_t smem = (size_t)(128 + N) * K * sizeof(bf16);
  ADN_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  ADN_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  { ADN_KERNEL("k_umma_selftest", st); k_umma_selftest<<<1, 128, smem, st>>>(mode, N, K, (const bf16*)A, (const bf16*)B, C, status); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

extern "C" int adn_selftest_umma_shift(int mode, int N, int K, int pitch, int shiftA, int shiftB, const void* A, const void* B,
                                       float* C, int* status, void* stream) {
  using namespace adn;
  ADN_REQUIRE(A && B && C && status, ADN_ERR_NULL, "adn_selftest_umma_shift: NULL argument");
  ADN_REQUIRE((mode == 0 || mode == 1) && N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256 &&
                  shiftA >= 0 && shiftB >= 0 && pitch <= 512,
              ADN_ERR_SHAPE, "adn_selftest_umma_shift: bad shape");
  ADN_REQUIRE(mode == 0 ? (shiftA + 128 <= pitch && shiftB + N <= pitch) : (shiftA + K <= pitch && shiftB + K <= pitch),
              ADN_ERR_SHAPE, "adn_selftest_umma_shift: shift + extent exceeds pitch");
  const int colsA = mode == 0 ? K : 128, colsB = mode == 0 ? K : N;
  size_t smem = (size_t)pitch * (colsA + colsB) * sizeof(bf16);
  ADN_REQUIRE(smem <= 200 * 1024, ADN_ERR_SHAPE, "adn_selftest_umma_shift: tile too large");
  ADN_CHECK_CUDA(cudaFuncSetAttribute(k_umma_shift_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  ADN_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  { ADN_KERNEL("k_umma_shift_selftest", st); k_umma_shift_selftest<<<1, 128, smem, st>>>(mode, N, K, pitch, shiftA, shiftB, (const bf16*)A, (const bf16*)B, C, status); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

extern "C" int adn_bench_umma(int mode, int N, int pitch, int shift, int iters, int ctas, long long* cycles, void* stream) {
  using namespace adn;
  ADN_REQUIRE(cycles && (mode == 0 || mode == 1) && N % 16 == 0 && N >= 16 && N <= 256 && pitch >= 128 && pitch <= 160 && shift >= 0 &&
                  shift <= 8 && ctas >= 1, ADN_ERR_SHAPE, "adn_bench_umma: bad arguments");
  const size_t smem = 64 * 1024;
  ADN_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  { ADN_KERNEL("k_umma_bench", st); k_umma_bench<<<ctas, 128, smem, st>>>(mode, N, pitch, shift, iters, cycles); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}
