// General bf16 x bf16 -> fp32 GEMM on the sm_100a tensor cores (tcgen05.mma, accumulator in TMEM), used by the "wide"
// ADN-SSD path (d_model 64..1024, d_state up to 128, any token-grid size): every contraction of the mixer that the
// 32-wide row / tile kernels fuse by hand is, at these widths, one launch of this kernel.
//
//   C[b][m][n] (=|+=) alpha * sum_seg sum_k opA_seg[b][m][k] * opB_seg[b][n][k]
//
// * Operands are row-major global tensors in either orientation, no copies or transposes on the host side:
//     K-major  : stored [rows = M or N][K], row pitch ld       (x @ W^T with W = [N][K]: in_proj, out_proj, readout ...)
//     MN-major : stored [K][cols = M or N], row pitch ld       (reductions over tokens: state, weight gradients; x @ W)
//   Both land in shared memory in the same un-swizzled "T8" core-matrix layout (sm100_utils.cuh) with 16-byte cp.async
//   copies (zero-filled outside the matrix: ragged M / N / K need no padding), and only the UMMA descriptor differs.
// * Up to two (A, B, K) segments accumulate into the same tile: [LN(y) | zc] @ W_out^T without a concatenated buffer, and
//   fp32 states as bf16 hi + lo pairs.
// * blockIdx.z = batch * splitk + split: per-sample GEMMs (state, readout) and split-K for the token reductions
//   (fp32 atomics into a zeroed accumulator).
// * CTA = 128 x BN tile (BN <= 128, multiple of 16), BK = 64, 3-stage cp.async ring; warps 0-3 produce, then drain the
//   accumulator (TMEM lane quarter = warp), warp 4 issues the MMAs.  96 KB of shared memory: two CTAs per SM, so one CTA's
//   epilogue overlaps the other's main loop.
#pragma once
#include "adn_common.cuh"
#include "sm100_utils.cuh"

namespace adn {
namespace tcg {
using namespace adn::sm100;

constexpr int BM = 128, BK = 64, STAGES = 3, LAG = 2, MAX_BN = 128;
constexpr int A_TILE_B = BM * BK * 2;                    // 16 KB in either orientation
constexpr int THREADS = 160;

enum { C_BF16 = 0, C_F32 = 1, C_ATOMIC_F32 = 2 };

struct Seg {
  const bf16* A; long long lda, a_bs;
  const bf16* B; long long ldb, b_bs;
  int K;
};

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
};                                            // the gradient w.r.t. the conv output becomes the gradient w.r.t. its input)

__device__ __forceinline__ void cp16(uint32_t sdst, const void* gsrc, int nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// rows x chunks (8 bf16 each) of a row-major global matrix -> T8 tile [chunk][R][8].  128 producer threads; one warp
// instruction covers 8 rows x 4 chunks: conflict-free 128-byte shared-memory runs, 64-byte global segments per row.
__device__ __forceinline__ void load_t8(uint32_t sdst, const bf16* __restrict__ g, long long ld, int R, int nrows, int nchunks,
                                        int rows_valid, int chunks_valid, int warp, int lane) {
  const int r8 = lane & 7, cq = lane >> 3;
  for (int rb = warp; rb * 8 < nrows; rb += 4) {
    const int r = rb * 8 + r8;
    const bool rok = r < rows_valid;
    const bf16* src = g + (long long)(rok ? r : 0) * ld;
    for (int c = cq; c < nchunks; c += 4) {
      const bool ok = rok && c < chunks_valid;
      cp16(sdst + (uint32_t)(c * R + r) * 16, ok ? (const void*)(src + c * 8) : (const void*)g, ok ? 16 : 0);
    }
  }
}

__global__ void __launch_bounds__(THREADS)
k_tcgemm(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = a.BN;
  const int stage_b = A_TILE_B + BN * BK * 2;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / a.splitk, split = blockIdx.z % a.splitk;
  // k-tile list: segment 0 restricted to this split's range, then segment 1
  int kb0 = 0, ke0 = a.seg[0].K;
  if (a.splitk > 1) { kb0 = split * a.k_per_split; ke0 = min(a.seg[0].K, kb0 + a.k_per_split); }
  const int nk0 = ke0 > kb0 ? (ke0 - kb0 + BK - 1) / BK : 0;
  const int nk1 = a.nseg > 1 ? (a.seg[1].K + BK - 1) / BK : 0;
  const int nk = nk0 + nk1;
  if (nk == 0) return;
  uint32_t tcols = 32;
  while ((int)tcols < BN) tcols <<= 1;
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t s0 = smem_u32(smem);
  bool ok = true;

  if (warp < 4) {
    // ---------------- producer
    for (int it = 0; it < nk + LAG; ++it) {
      if (it < nk) {
        const int s = it % STAGES;
        if (it >= STAGES) ok &= mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
        const bool second = it >= nk0;      // scalar selects (no dynamic indexing of the kernel parameter struct)
        const long long lda = second ? a.seg[1].lda : a.seg[0].lda, ldb = second ? a.seg[1].ldb : a.seg[0].ldb;
        const int k0 = second ? (it - nk0) * BK : kb0 + it * BK;
        const int kvalid = (second ? a.seg[1].K : ke0) - k0;                       // > 0
        const bf16* A = second ? a.seg[1].A + (long long)batch * a.seg[1].a_bs : a.seg[0].A + (long long)batch * a.seg[0].a_bs;
        const bf16* B = second ? a.seg[1].B + (long long)batch * a.seg[1].b_bs : a.seg[0].B + (long long)batch * a.seg[0].b_bs;
        const uint32_t sa = s0 + s * stage_b, sb = sa + A_TILE_B;
        if (!a.a_mn) load_t8(sa, A + (long long)m0 * lda + k0, lda, BM, BM, BK / 8, a.M - m0, (kvalid + 7) >> 3, warp, lane);
        else         load_t8(sa, A + (long long)k0 * lda + m0, lda, BK, BK, BM / 8, kvalid, (a.M - m0 + 7) >> 3, warp, lane);
        if (!a.b_mn) load_t8(sb, B + (long long)n0 * ldb + k0, ldb, BN, BN, BK / 8, a.N - n0, (kvalid + 7) >> 3, warp, lane);
        else         load_t8(sb, B + (long long)k0 * ldb + n0, ldb, BK, BK, BN / 8, kvalid, (a.N - n0 + 7) >> 3, warp, lane);
      }
      cp_commit();
      if (it >= LAG) {
        cp_wait<LAG>();
        fence_async_smem();
        mbar_arrive1(&full[(it - LAG) % STAGES]);
      }
    }
    // ---------------- epilogue: TMEM lane quarter `warp`, thread = output row
    ok &= mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const float alpha = a.alpha ? *a.alpha : 1.f;
    const bool row_ok = m < a.M;
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
      tmem_wait_ld();
      const int n = n0 + c;
      if (!row_ok || n >= a.N) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] *= alpha;
        if (a.parity_mask && ((m ^ (n + j)) & 1)) v[j] = 0.f;
      }
      const int nv = min(16, a.N - n);
      if (a.aux != nullptr) {
        const bf16* q = a.aux + (long long)batch * a.aux_bs + (long long)m * a.ld_aux + n;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nv) v[j] *= silu_gradf_(__bfloat162float(q[j]));
      }
      if (a.c_mode == C_BF16) {
        bf16* p = (bf16*)a.C + (long long)batch * a.c_bs + (long long)m * a.ldc + n;
        if (nv == 16 && ((uintptr_t)p & 15) == 0) {
          const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
          const float hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
          reinterpret_cast<uint4*>(p)[0] = pack8(lo);
          reinterpret_cast<uint4*>(p)[1] = pack8(hi);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < nv) p[j] = __float2bfloat16_rn(v[j]);
        }
      } else {
        float* p = (float*)a.C + (long long)batch * a.c_bs + (long long)m * a.ldc + n;
        if (a.c_mode == C_F32) {
          if (nv == 16 && ((uintptr_t)p & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < nv) p[j] = v[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nv && !(a.parity_mask && ((m ^ (n + j)) & 1))) atomicAdd(p + j, v[j]);
        }
      }
    }
  } else {
    // ---------------- MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_rt(BM, BN, a.a_mn != 0, a.b_mn != 0);
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        ok &= mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = s0 + s * stage_b, sb = sa + A_TILE_B;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t da = a.a_mn ? desc_mnmajor(sa, BK, 0, ks * 16) : desc_kmajor(sa, BM, 0, ks * 16);
          const uint64_t db = a.b_mn ? desc_mnmajor(sb, BK, 0, ks * 16) : desc_kmajor(sb, BN, 0, ks * 16);
          umma(tbase, da, db, idesc, (it | ks) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(&acc_full);
    }
  }
  if (!ok && a.status) *a.status = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, tcols);
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

static inline int pick_bn(int N) {
  int r = (N + 15) / 16 * 16;
  return r < MAX_BN ? r : MAX_BN;
}

// C = alpha * (A0 . B0 [+ A1 . B1]);  batches > 1: per-sample GEMMs;  splitk > 1: atomics into a ZEROED fp32 C.
struct Aux {
  const bf16* p; long long ld, bs;
};
static const Aux NOAUX = Aux{nullptr, 0, 0};

static int gemm(cudaStream_t st, const char* name, int M, int N, int K0, Op A0, Op B0, int K1, Op A1, Op B1, Out C,
                int batches, int splitk, const float* alpha, int parity_mask, int* status, Aux aux = NOAUX) {
  ADN_REQUIRE(M > 0 && N > 0 && K0 > 0 && batches > 0, ADN_ERR_SHAPE, "tcgemm %s: empty problem", name);
  ADN_REQUIRE(K1 == 0 || (A1.mn == A0.mn && B1.mn == B0.mn), ADN_ERR_SHAPE, "tcgemm %s: segments must share orientation", name);
  ADN_REQUIRE(splitk == 1 || (K1 == 0 && C.mode == C_ATOMIC_F32), ADN_ERR_SHAPE, "tcgemm %s: split-K needs one segment and an atomic fp32 output", name);
  // 16-byte cp.async granularity: pitches and the contiguous extent in units of 8 bf16
  ADN_REQUIRE(A0.ld % 8 == 0 && B0.ld % 8 == 0 && (K1 == 0 || (A1.ld % 8 == 0 && B1.ld % 8 == 0)), ADN_ERR_SHAPE, "tcgemm %s: row pitches must be multiples of 8", name);
  ADN_REQUIRE(((uintptr_t)A0.p | (uintptr_t)B0.p | (uintptr_t)A1.p | (uintptr_t)B1.p) % 16 == 0, ADN_ERR_SHAPE, "tcgemm %s: operands must be 16-byte aligned", name);
  ADN_REQUIRE((A0.mn || K0 % 8 == 0) && (B0.mn || K0 % 8 == 0) && (K1 % 8 == 0 || (A1.mn && B1.mn)), ADN_ERR_SHAPE, "tcgemm %s: K-major operands need K %% 8 == 0", name);
  Args a;
  a.seg[0] = Seg{A0.p, A0.ld, A0.bs, B0.p, B0.ld, B0.bs, K0};
  a.seg[1] = Seg{A1.p, A1.ld, A1.bs, B1.p, B1.ld, B1.bs, K1};
  a.nseg = K1 > 0 ? 2 : 1;
  a.M = M; a.N = N; a.BN = pick_bn(N);
  a.a_mn = A0.mn; a.b_mn = B0.mn;
  a.C = C.p; a.ldc = C.ld; a.c_bs = C.bs; a.c_mode = C.mode;
  a.alpha = alpha; a.parity_mask = parity_mask;
  a.splitk = splitk < 1 ? 1 : splitk;
  a.k_per_split = (cdiv(cdiv(K0, a.splitk), BK)) * BK;
  a.splitk = cdiv(K0, a.k_per_split);
  a.status = status;
  a.aux = aux.p; a.ld_aux = aux.ld; a.aux_bs = aux.bs;
  const size_t smem = (size_t)STAGES * (A_TILE_B + a.BN * BK * 2);
  static bool attr_done = false;
  if (!attr_done) {
    ADN_CHECK_CUDA(cudaFuncSetAttribute(k_tcgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * (A_TILE_B + MAX_BN * BK * 2)));
    attr_done = true;
  }
  dim3 grid(cdiv(M, BM), cdiv(N, a.BN), batches * a.splitk);
  ADN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, ADN_ERR_SHAPE, "tcgemm %s: grid too large", name);
  { ADN_KERNEL(name, st); k_tcgemm<<<grid, THREADS, smem, st>>>(a); }
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
