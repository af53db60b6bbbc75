// Threshold counts on device: float2int + _cal_frame of the reference (datasets/Shanghai_metrics.py:45-47,105-114).
// Integer work, HBM-bound: one pass over obs and sim (8 bytes per element), per-thread counters, one atomic per
// (block, threshold, class).
#include "adn_common.cuh"

namespace adn {

constexpr int MAX_THR = 8;
struct ThrList { int n; int t[MAX_THR]; };

__device__ __forceinline__ int quantise(float v, float scale) {
  // np.clip(arr, 0, 1) * value_scale -> astype(uint16): fp32 multiply, then truncation toward zero
  float c = fminf(fmaxf(v, 0.0f), 1.0f);
  return (int)(unsigned short)(int)__fmul_rn(c, scale);
}

__global__ void __launch_bounds__(256)
k_threshold_counts(const float* __restrict__ obs, const float* __restrict__ sim, long long n, ThrList thr, float scale,
                   unsigned long long* __restrict__ table) {
  unsigned int cnt[MAX_THR][3];  // TP, FN, FP ; TN = total - sum
#pragma unroll
  for (int i = 0; i < MAX_THR; ++i) cnt[i][0] = cnt[i][1] = cnt[i][2] = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n >> 2;
  unsigned int seen = 0;
  for (; i < n4; i += stride) {
    float4 o = reinterpret_cast<const float4*>(obs)[i], s = reinterpret_cast<const float4*>(sim)[i];
    const float ov[4] = {o.x, o.y, o.z, o.w}, sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int qo = quantise(ov[j], scale), qs = quantise(sv[j], scale);
#pragma unroll
      for (int t = 0; t < MAX_THR; ++t)
        if (t < thr.n) {
          bool a = qo >= thr.t[t], b = qs >= thr.t[t];
          cnt[t][0] += (a && b);
          cnt[t][1] += (a && !b);
          cnt[t][2] += (!a && b);
        }
    }
    seen += 4;
  }
  // tail (n % 4 elements), handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    long long j = (n4 << 2) + threadIdx.x;
    int qo = quantise(obs[j], scale), qs = quantise(sim[j], scale);
#pragma unroll
    for (int t = 0; t < MAX_THR; ++t)
      if (t < thr.n) {
        bool a = qo >= thr.t[t], b = qs >= thr.t[t];
        cnt[t][0] += (a && b);
        cnt[t][1] += (a && !b);
        cnt[t][2] += (!a && b);
      }
    seen += 1;
  }
  __shared__ unsigned int red[8][MAX_THR * 4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < MAX_THR; ++t) {
    if (t >= thr.n) break;
    unsigned int v[4] = {cnt[t][0], cnt[t][1], cnt[t][2], seen - cnt[t][0] - cnt[t][1] - cnt[t][2]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int x = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) red[wid][t * 4 + k] = x;
    }
  }
  __syncthreads();
  if (threadIdx.x < thr.n * 4) {
    unsigned long long s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (s) atomicAdd(table + threadIdx.x, s);
  }
}

// One evaluation batch of SimplifiedEvaluator.evaluate (datasets/Shanghai_metrics.py:49-103): frames f = b * seq_len + t of
// `elems` pixels each.  Per frame: the threshold counts (accumulated into the running table) and the squared error of the
// clipped, value_scale-d frames (:116-121), whose per-frame mean is accumulated per lead time t (what `done` turns into the
// RMSE, :276).  grid (chunks per frame, frames); integer counts bit-exact, squared error summed in fp64.
__global__ void __launch_bounds__(256)
k_eval_frames(const float* __restrict__ tru, const float* __restrict__ pred, long long elems, int seq_len, ThrList thr, float scale,
              unsigned long long* __restrict__ table, double* __restrict__ mse_t) {
  const long long f = blockIdx.y;
  const float* o = tru + f * elems;
  const float* s = pred + f * elems;
  unsigned int cnt[MAX_THR][3];
#pragma unroll
  for (int i = 0; i < MAX_THR; ++i) cnt[i][0] = cnt[i][1] = cnt[i][2] = 0;
  unsigned int seen = 0;
  double sq = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
    const float ov = o[i], sv = s[i];
    const int qo = quantise(ov, scale), qs = quantise(sv, scale);
#pragma unroll
    for (int t = 0; t < MAX_THR; ++t)
      if (t < thr.n) {
        const bool a = qo >= thr.t[t], b = qs >= thr.t[t];
        cnt[t][0] += (a && b);
        cnt[t][1] += (a && !b);
        cnt[t][2] += (!a && b);
      }
    const float d = __fmul_rn(fminf(fmaxf(sv, 0.f), 1.f), scale) - __fmul_rn(fminf(fmaxf(ov, 0.f), 1.f), scale);
    sq += (double)d * (double)d;
    seen += 1;
  }
  __shared__ unsigned int red[8][MAX_THR * 4];
  __shared__ double redsq[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < MAX_THR; ++t) {
    if (t >= thr.n) break;
    unsigned int v[4] = {cnt[t][0], cnt[t][1], cnt[t][2], seen - cnt[t][0] - cnt[t][1] - cnt[t][2]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int x = v[k];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      if (lane == 0) red[wid][t * 4 + k] = x;
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, m);
  if (lane == 0) redsq[wid] = sq;
  __syncthreads();
  if (threadIdx.x < thr.n * 4) {
    unsigned long long a = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    if (a) atomicAdd(table + threadIdx.x, a);
  }
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += redsq[w];
    atomicAdd(mse_t + (int)(f % seq_len), a / (double)elems);
  }
}

}  // namespace adn

extern "C" int adn_eval_batch(const float* true_batch, const float* pred_batch, int64_t batch, int32_t seq_len, int64_t frame_elems,
                              const int32_t* thresholds, int32_t n_thresholds, float value_scale, int64_t* table, double* mse_t,
                              void* stream) {
  using namespace adn;
  ADN_REQUIRE(true_batch && pred_batch && thresholds && table && mse_t, ADN_ERR_NULL, "adn_eval_batch: NULL argument");
  ADN_REQUIRE(batch > 0 && seq_len > 0 && frame_elems > 0 && batch * seq_len <= 65535 && n_thresholds > 0 && n_thresholds <= MAX_THR,
              ADN_ERR_SHAPE, "adn_eval_batch: batch * seq_len in 1..65535, frame_elems > 0 and 1..%d thresholds required", MAX_THR);
  ThrList thr;
  thr.n = n_thresholds;
  for (int i = 0; i < MAX_THR; ++i) thr.t[i] = i < n_thresholds ? thresholds[i] : 0;
  cudaStream_t st = (cudaStream_t)stream;
  int chunks = cdiv(frame_elems, 256 * 16);
  const int cap = cdiv(8LL * sm_count(), batch * seq_len);
  chunks = chunks > cap ? (cap < 1 ? 1 : cap) : chunks;
  dim3 grid(chunks, (unsigned)(batch * seq_len));
  { ADN_KERNEL("k_eval_frames", st); k_eval_frames<<<grid, 256, 0, st>>>(true_batch, pred_batch, frame_elems, seq_len, thr, value_scale, (unsigned long long*)table, mse_t); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}

extern "C" int adn_threshold_counts(const float* obs, const float* sim, int64_t n, const int32_t* thresholds,
                                    int32_t n_thresholds, float value_scale, int64_t* table, void* stream) {
  using namespace adn;
  ADN_REQUIRE(table && thresholds && ((obs && sim) || n == 0), ADN_ERR_NULL, "adn_threshold_counts: NULL argument");
  ADN_REQUIRE(n >= 0 && n_thresholds > 0 && n_thresholds <= MAX_THR, ADN_ERR_SHAPE,
              "adn_threshold_counts: n >= 0 and 1..%d thresholds required", MAX_THR);
  ADN_REQUIRE(((uintptr_t)obs % 16 == 0) && ((uintptr_t)sim % 16 == 0), ADN_ERR_SHAPE,
              "adn_threshold_counts: obs / sim must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ThrList thr;
  thr.n = n_thresholds;
  for (int i = 0; i < MAX_THR; ++i) thr.t[i] = i < n_thresholds ? thresholds[i] : 0;
  ADN_CHECK_CUDA(cudaMemsetAsync(table, 0, sizeof(int64_t) * 4 * n_thresholds, st));
  if (n == 0) return ADN_OK;
  // each thread may count at most 2^32 events: 148*8 blocks x 256 threads x 4 => fine up to 2^50 elements
  long long blocks = (n / 4 + 255) / 256;
  int grid = (int)(blocks < 1 ? 1 : (blocks > sm_count() * 8 ? sm_count() * 8 : blocks));
  { ADN_KERNEL("k_threshold_counts", st); k_threshold_counts<<<grid, 256, 0, st>>>(obs, sim, n, thr, value_scale, (unsigned long long*)table); }
  ADN_CHECK_LAUNCH();
  return ADN_OK;
}
