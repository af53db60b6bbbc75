// sm_100a tensor-core (tcgen05 / TMEM) path of the ADN-SSD mixer, bf16 I/O.
#include "adnssd_sm100.cuh"

#include "sm100_utils.cuh"

namespace adn {
using namespace sm100;

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
    // tiles are [K tokens][channels]: rows = tokens
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

bool sm100_supported(const MixerDims&) { return false; }
void sm100_workspace_bytes(const MixerDims&, size_t* f, size_t* b) { *f = 0; *b = 0; }
int sm100_forward(const MixerDims&, const AdnWeights&, const bf16*, bf16*, void*, void*, cudaStream_t) {
  set_error("sm100 path not built"); return ADN_ERR_ARCH; }
int sm100_backward(const MixerDims&, const AdnWeights&, const bf16*, const void*, const bf16*, bf16*,
                   const AdnWeightGrads&, void*, cudaStream_t) { set_error("sm100 path not built"); return ADN_ERR_ARCH; }
}  // namespace adn

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
