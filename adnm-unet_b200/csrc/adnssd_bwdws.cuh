// Warp-specialised backward phases B1 / B2 of the ADN-SSD mixer for d_model 32 / d_state 16 (DI 64, GN 32), any token
// count that is a multiple of 128.  Same math as k_bwd1 / k_bwd2 (adnssd_sm100.cu, oracle/adnssd_oracle.py::mixer_backward)
// but with the three roles of a tile decoupled so that nothing waits on a serial load -> MMA -> epilogue chain:
//   warps 0-15 epilogue: thread = (token row = TMEM lane, column quarter); the four quarters of a row exchange their
//              partial LayerNorm sums through shared memory
//   warp  16   producer: bulk copies of the TL operands + cp.async of the row-major dout tile, two tiles in flight
//   warp  17   one elected lane issues every tcgen05.mma
// Two tiles are in flight (two shared-memory stages, two TMEM accumulator sets); the per-sample state images (bf16 hi + lo)
// are double buffered by sample parity and staged by the producer warp.
#pragma once

namespace bwdws {
using namespace adn;
using namespace adn::sm100;
using rowconv::elect_one;
using rowconv::mbar_arrive;
using rowconv::umma_c;
using rowconv::dadd;

constexpr int D = 32, DI = 64, GN = 32, NA = 24, NH = 16;
constexpr int XC = 8, CCH = 4, DC = 4;

// ------------------------------------------------------------------------------------------------
// State images: S[j][c] (GN x DI fp32) -> bf16 hi / lo as K-major B operands in both orientations (32 lanes):
//   "a": rows = c, K = j : element (c, j) at ((j/8)*DI + c)*8 + j%8       Y[tok][c]  = sum_j C[tok][j]  S[j][c]
//   "b": rows = j, K = c : element (j, c) at ((c/8)*GN + j)*8 + c%8       dC[tok][j] = sum_c dy[tok][c] S[j][c]
// ------------------------------------------------------------------------------------------------
constexpr int SIMG_B = 4 * (GN / 8) * DI * 16;    // a_hi, a_lo, b_hi, b_lo: 4 x 4096 bytes
__device__ __forceinline__ void stage_state_warp(const float* __restrict__ M, uint8_t* img, int lane) {
  bf16* a_hi = reinterpret_cast<bf16*>(img);
  bf16* a_lo = a_hi + (GN / 8) * DI * 8;
  bf16* b_hi = a_lo + (GN / 8) * DI * 8;
  bf16* b_lo = b_hi + (DI / 8) * GN * 8;
  for (int idx = lane; idx < (GN / 8) * DI; idx += 32) {
    const int jc = idx / DI, c = idx % DI;
    float v[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q] = M[(jc * 8 + q) * DI + c];
      l[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
    }
    *reinterpret_cast<uint4*>(a_hi + (jc * DI + c) * 8) = pack8(v);
    *reinterpret_cast<uint4*>(a_lo + (jc * DI + c) * 8) = pack8(l);
  }
  for (int idx = lane; idx < (DI / 8) * GN; idx += 32) {
    const int cc = idx % (DI / 8), j = idx / (DI / 8);
    float v[8], l[8];
    const float4 p0 = *reinterpret_cast<const float4*>(M + j * DI + cc * 8);
    const float4 p1 = *reinterpret_cast<const float4*>(M + j * DI + cc * 8 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
#pragma unroll
    for (int q = 0; q < 8; ++q) l[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
    *reinterpret_cast<uint4*>(b_hi + (cc * GN + j) * 8) = pack8(v);
    *reinterpret_cast<uint4*>(b_lo + (cc * GN + j) * 8) = pack8(l);
  }
}

// ------------------------------------------------------------------------------------------------
// k_bwd1_ws (phase B1), per 128-token tile:
//   MMA1: g = dout . W_out (N = 128),  Y = Cc . S' (hi + lo)
//   EPI1: y = Y + D*x, LayerNorm statistics, yhat -> sCat;  dy = LN backward of alpha1*g_y -> sXY (over x) and global;
//         dpre_z = alpha1 * g_z * SiLU'(pre_z) -> global;  sum(dout)
//   MMA2: dCc = dy . S'^T (hi + lo);  Rt += [yhat | zc]^T . dout;  dS' += dy^T . Cc
//   EPI2: dpre_C = dCc * SiLU'(pre_C) -> global
// ------------------------------------------------------------------------------------------------
constexpr int B1_STG_B = (DC + CCH + 16 + 16) * 2048;      // dout | C | [yhat | zc] | [x -> dy | zero padding]
constexpr int B1_WT_B = DC * 2 * DI * 16;                  // W_out^T image: [4 chunks of d][128 rows j'][8]
constexpr int B1_XCH_B = 2 * 4 * 128 * 16;                 // two exchange buffers [column quarter][row] of float4
constexpr int WS_EPI_WARPS = 16, WS_THREADS = (WS_EPI_WARPS + 2) * 32;   // 4 column quarters x 4 lane quarters + producer + MMA
constexpr int B1_SMEM = 2 * B1_STG_B + B1_WT_B + 2 * SIMG_B + B1_XCH_B;
constexpr int B1_TSTG = 224, B1_COL_Y = 128, B1_COL_DC = 192, B1_COL_RT = 448, B1_COL_DS = 480;

__global__ void __launch_bounds__(WS_THREADS, 1)
k_bwd1_ws(const bf16* __restrict__ dout, const bf16* __restrict__ act, const bf16* __restrict__ sgrad, const float* __restrict__ S,
          const float* __restrict__ Dp, const float* __restrict__ gamma, const float* __restrict__ alpha1p,
          const bf16* __restrict__ Wout, bf16* __restrict__ dact, float* __restrict__ Rt, float* __restrict__ sdout,
          float* __restrict__ dS, int tiles_per_batch, int num_tiles, int tiles_per_cta, int* __restrict__ status,
          float* __restrict__ dalpha1) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[2], mma1_done[2], epi1_done[2], mma2_done[2], acc_free[2], gy_free[2], s_done, s_free;
  if (blockIdx.x == 0 && threadIdx.x == 0 && dalpha1 != nullptr) *dalpha1 = 0.f;   // k_finalize_fast accumulates into it
  __shared__ uint32_t tmem_slot;
  __shared__ float sG[DI], sDh[DI];
  uint8_t* sStg = smem;
  uint8_t* sWT = smem + 2 * B1_STG_B;
  uint8_t* sImg = sWT + B1_WT_B;
  float2* sXch = reinterpret_cast<float2*>(sImg + 2 * SIMG_B);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T0 = blockIdx.x * tiles_per_cta, T1 = min(num_tiles, T0 + tiles_per_cta);
  for (int i = tid; i < DI; i += WS_THREADS) { sG[i] = gamma[i]; sDh[i] = Dp[head_of_channel(i, 4)]; }
  // zero padding chunks [XC, 16) of the x -> dy operand of both stages (M = 128 rows of the dS' reduction)
  for (int st = 0; st < 2; ++st)
    for (int i = tid; i < 8 * 128; i += WS_THREADS)
      reinterpret_cast<uint4*>(sStg + st * B1_STG_B + (DC + CCH + 16 + XC) * 2048)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < DC * 2 * DI; i += WS_THREADS) {     // W_out^T image: row j', K = d
    const int dc = i / (2 * DI), j = i % (2 * DI);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __bfloat162float(Wout[(long long)(dc * 8 + q) * 2 * DI + j]);
    *reinterpret_cast<uint4*>(sWT + (dc * 2 * DI + j) * 16) = pack8(v);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 33);            // 32 producer lanes (cp.async + state image) + 1 expect_tx arrival
      mbar_init(&mma1_done[i], 1);
      mbar_init(&epi1_done[i], WS_EPI_WARPS);
      mbar_init(&mma2_done[i], 1);
      mbar_init(&acc_free[i], WS_EPI_WARPS);      // dCc columns consumed (second epilogue)
      mbar_init(&gy_free[i], WS_EPI_WARPS);       // g / Y columns consumed (first epilogue): MMA1 of tile t+2 does not wait for EPI2
    }
    mbar_init(&s_done, 1);
    mbar_init(&s_free, WS_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == WS_EPI_WARPS + 1) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  bool ok = true;
  if (T0 < T1) {
    if (warp == WS_EPI_WARPS) {
      // ---------------- producer
      int cur_b = -1;
      for (int t = T0; t < T1; ++t) {
        const int it = t - T0, s = it & 1, b = t / tiles_per_batch;
        uint8_t* sb = sStg + s * B1_STG_B;
        if (it >= 2) ok = mbar_wait(&mma2_done[s], ((it >> 1) - 1) & 1) && ok;     // every reader of this stage has finished
        if (b != cur_b) {      // new sample: its state images (buffer b & 1; the sample before last is long finished)
          stage_state_warp(S + (long long)b * GN * DI, sImg + (b & 1) * SIMG_B, lane);
          cur_b = b;
        }
        if (lane == 0) {
          mbar_expect_tx(&full[s], (uint32_t)(2 * XC + CCH) * 2048);
          // act TL: z (chunks 0..7) -> sCat[8..16), x (8..15) -> sXY[0..8): adjacent in the stage; C (20..23) -> sC
          bulk_g2s(sb + (DC + CCH + XC) * 2048, act + ((long long)t * NA) * 1024, 2 * XC * 2048, &full[s]);
          bulk_g2s(sb + DC * 2048, act + ((long long)t * NA + 2 * XC + CCH) * 1024, CCH * 2048, &full[s]);
          if (t + 1 < T1) {     // only two shared-memory stages: the next tile's loads cannot start before this stage's
            const long long nt = t + 1;   // reader is done, so pull its operands into L2 now (their latency is then ~1/3)
            bulk_prefetch_l2(act + (nt * NA) * 1024, 2 * XC * 2048);
            bulk_prefetch_l2(act + (nt * NA + 2 * XC + CCH) * 1024, CCH * 2048);
            bulk_prefetch_l2(dout + nt * 128 * D, 128 * D * 2);
            bulk_prefetch_l2(sgrad + (nt * NA) * 1024, XC * 2048);
            bulk_prefetch_l2(sgrad + (nt * NA + 2 * XC + CCH) * 1024, CCH * 2048);
          }
        }
        // dout tile (row-major external tensor) -> T8: 512 16-byte pieces, 16 per lane
        {
          const bf16* src = dout + (long long)t * 128 * D;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int piece = i * 32 + lane, tok = piece >> 2, ch = piece & 3;
            cp_async16(sb + (ch * 128 + tok) * 16, src + piece * 8, 16);
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_async_smem();
        mbar_arrive(&full[s]);
      }
    } else if (warp == WS_EPI_WARPS + 1) {
      // ---------------- MMA issue
      if (elect_one()) {
        const uint32_t sbase = smem_u32(sStg), wbase = smem_u32(sWT), ibase = smem_u32(sImg);
        const uint32_t id_g = make_idesc_rt(128, 2 * DI, false, false), id_y = make_idesc_rt(128, DI, false, false),
                       id_c = make_idesc_rt(128, GN, false, false), id_r = make_idesc_rt(128, D, true, true),
                       id_s = make_idesc_rt(128, GN, true, true);
        int fl = 0;
        bool rt_fresh = true;
        auto mma2 = [&](int t) {
          const int it = t - T0, s = it & 1, b = t / tiles_per_batch, i_in_b = t % tiles_per_batch;
          ok = mbar_wait(&epi1_done[s], (it >> 1) & 1) && ok;
          if (it >= 2) ok = mbar_wait(&acc_free[s], ((it >> 1) - 1) & 1) && ok;     // dCc of tile t-2 has been read
          tc_fence_after();
          const uint32_t sb = sbase + s * B1_STG_B, tb = tbase + s * B1_TSTG;
          const uint32_t aDout = sb, aC = sb + DC * 2048, aCat = sb + (DC + CCH) * 2048, aXY = sb + (DC + CCH + 16) * 2048;
          const uint32_t img = ibase + (b & 1) * SIMG_B;
          const uint64_t dXYk = make_desc(aXY, 2048, 128), dBh = make_desc(img + 8192, 512, 128), dBl = make_desc(img + 12288, 512, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k == 0) umma_c<false>(tb + B1_COL_DC, dXYk, dBh, id_c);
            else umma_c<true>(tb + B1_COL_DC, dadd(dXYk, k * 2 * 2048), dadd(dBh, k * 2 * 512), id_c);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_c<true>(tb + B1_COL_DC, dadd(dXYk, k * 2 * 2048), dadd(dBl, k * 2 * 512), id_c);
          const uint64_t dCat = make_desc(aCat, 128, 2048), dDo = make_desc(aDout, 128, 2048);
#pragma unroll
          for (int k = 0; k < 8; ++k) umma(tbase + B1_COL_RT, dadd(dCat, k * 256), dadd(dDo, k * 256), id_r, !(rt_fresh && k == 0));
          rt_fresh = false;
          const bool first_of_sample = (t == T0) || (i_in_b == 0);
          if (first_of_sample && fl > 0) { ok = mbar_wait(&s_free, (fl - 1) & 1) && ok; tc_fence_after(); }
          const uint64_t dXYm = make_desc(aXY, 128, 2048), dCm = make_desc(aC, 128, 2048);
#pragma unroll
          for (int k = 0; k < 8; ++k) umma(tbase + B1_COL_DS, dadd(dXYm, k * 256), dadd(dCm, k * 256), id_s, !(first_of_sample && k == 0));
          umma_commit(&mma2_done[s]);
          if ((t == T1 - 1) || (i_in_b == tiles_per_batch - 1)) { umma_commit(&s_done); ++fl; }
        };
        for (int t = T0; t < T1; ++t) {
          const int it = t - T0, s = it & 1, b = t / tiles_per_batch;
          ok = mbar_wait(&full[s], (it >> 1) & 1) && ok;
          if (it >= 2) ok = mbar_wait(&gy_free[s], ((it >> 1) - 1) & 1) && ok;
          tc_fence_after();
          const uint32_t sb = sbase + s * B1_STG_B, tb = tbase + s * B1_TSTG;
          const uint32_t img = ibase + (b & 1) * SIMG_B;
          const uint64_t dDo = make_desc(sb, 2048, 128), dW = make_desc(wbase, 2 * DI * 16, 128);
          umma_c<false>(tb, dDo, dW, id_g);
          umma_c<true>(tb, dadd(dDo, 2 * 2048), dadd(dW, 2 * 2 * DI * 16), id_g);
          const uint64_t dC = make_desc(sb + DC * 2048, 2048, 128), dAh = make_desc(img, DI * 16, 128), dAl = make_desc(img + 4096, DI * 16, 128);
          umma_c<false>(tb + B1_COL_Y, dC, dAh, id_y);
          umma_c<true>(tb + B1_COL_Y, dadd(dC, 2 * 2048), dadd(dAh, 2 * DI * 16), id_y);
          umma_c<true>(tb + B1_COL_Y, dC, dAl, id_y);
          umma_c<true>(tb + B1_COL_Y, dadd(dC, 2 * 2048), dadd(dAl, 2 * DI * 16), id_y);
          umma_commit(&mma1_done[s]);
          if (t > T0) mma2(t - 1);
        }
        mma2(T1 - 1);
        if (!ok) atomicExch(status, 30);
      }
    } else {
      // ---------------- epilogue: thread = (token row, column quarter cq): channels [16 cq, 16 cq + 16)
      const int q = warp & 3, cq = warp >> 2, row = q * 32 + lane;
      const float a1 = *alpha1p;
      float sd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) sd[i] = 0.f;
      int fl = 0;
      PhaseTimer pt(0, tid == 0);
      uint4 sgc_prev, sgc_cur;     // SiLU' of this thread's C columns: fetched one tile ahead of their use in epi2
      sgc_prev = make_uint4(0u, 0u, 0u, 0u);
      // (sample, tile-in-sample) of the tile handled by epi2, tracked without per-tile integer divisions
      int e2_b = T0 / tiles_per_batch, e2_i = T0 - e2_b * tiles_per_batch;
      auto epi2 = [&](int t) {
        const int it = t - T0, s = it & 1, b = e2_b, i_in_b = e2_i;
        if (++e2_i == tiles_per_batch) { e2_i = 0; ++e2_b; }
        ok = mbar_wait(&mma2_done[s], (it >> 1) & 1) && ok;
        tc_fence_after();
        pt.mark(5);
        const uint32_t ta = tbase + ((uint32_t)(q * 32) << 16) + s * B1_TSTG;
        float v[8];
        tmem_ld8(ta + B1_COL_DC + cq * 8, v);
        float s0[8];
        unpack8(sgc_prev, s0);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free[s]);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= s0[j];
        *reinterpret_cast<uint4*>(dact + (((long long)t * NA + 2 * XC + CCH + cq) * 128 + row) * 8) = pack8(v);
        if ((t == T1 - 1) || (i_in_b == tiles_per_batch - 1)) {     // flush dS' of sample b
          ok = mbar_wait(&s_done, fl & 1) && ok;
          ++fl;
          tc_fence_after();
          if (q < 2 && ok) {
            float w[8];
            tmem_ld8(tbase + ((uint32_t)(q * 32) << 16) + B1_COL_DS + cq * 8, w);
            tmem_wait_ld();
            const int c = q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int jj = cq * 8 + j;
              if (((jj ^ c) & 1) == 0) atomicAdd(dS + ((long long)b * GN + jj) * DI + c, w[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free);
        }
      };
      for (int t = T0; t < T1; ++t) {
        const int it = t - T0, s = it & 1;
        uint8_t* sb = sStg + s * B1_STG_B;
        const uint8_t* sDoutRow = sb + row * 16;
        uint8_t* sCatRow = sb + (DC + CCH) * 2048 + row * 16;
        uint8_t* sXYRow = sb + (DC + CCH + 16) * 2048 + row * 16;
        // SiLU' of this thread's z and C columns, fetched while MMA1 runs
        uint4 sgz[2];
        {
          const bf16* srow = sgrad + (((long long)t * NA + 2 * cq) * 128 + row) * 8;
          sgz[0] = __ldg(reinterpret_cast<const uint4*>(srow));
          sgz[1] = __ldg(reinterpret_cast<const uint4*>(srow + 1024));
          sgc_cur = __ldg(reinterpret_cast<const uint4*>(sgrad + (((long long)t * NA + 2 * XC + CCH + cq) * 128 + row) * 8));
        }
        pt.mark(6);
        ok = mbar_wait(&full[s], (it >> 1) & 1) && ok;      // x and dout of this stage are read from shared memory below
        ok = mbar_wait(&mma1_done[s], (it >> 1) & 1) && ok;
        tc_fence_after();
        pt.mark(0);
        const uint32_t ta = tbase + ((uint32_t)(q * 32) << 16) + s * B1_TSTG;
        // ---- all three accumulator slices of this thread (Y, g_y, g_z: 48 columns) are requested up front so that the
        // TMEM latency is exposed once per tile instead of three times
        float y[16], gy[16], gz[16];
        float p1 = 0.f, p2 = 0.f;
        {
          float x0[8], x1[8];
          tmem_ld16(ta + B1_COL_Y + cq * 16, y);
          tmem_ld16(ta + cq * 16, gy);
          tmem_ld16(ta + DI + cq * 16, gz);
          unpack8(*reinterpret_cast<const uint4*>(sXYRow + (2 * cq) * 2048), x0);
          unpack8(*reinterpret_cast<const uint4*>(sXYRow + (2 * cq + 1) * 2048), x1);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            y[j] = fmaf(sDh[cq * 16 + j], x0[j], y[j]);
            y[8 + j] = fmaf(sDh[cq * 16 + 8 + j], x1[j], y[8 + j]);
          }
        }
        // ---- ONE exchange per tile: the LayerNorm statistics (sum y, sum y^2) and the two sums of its backward
        // (sum gy, sum gy*y with gy = alpha1 * g_y * gamma) all come straight from the accumulators, and
        // mean(gy * yhat) = rstd * (mean(gy * y) - mu * mean(gy)).  Exchange buffers alternate per tile (one barrier per tile).
        float p3 = 0.f, p4 = 0.f;
        {
          float q1 = 0.f, q2 = 0.f, q3 = 0.f, q4 = 0.f;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            gy[c] = a1 * gy[c] * sG[cq * 16 + c];
            gy[c + 1] = a1 * gy[c + 1] * sG[cq * 16 + c + 1];
            p1 += y[c]; p2 = fmaf(y[c], y[c], p2); p3 += gy[c]; p4 = fmaf(gy[c], y[c], p4);
            q1 += y[c + 1]; q2 = fmaf(y[c + 1], y[c + 1], q2); q3 += gy[c + 1]; q4 = fmaf(gy[c + 1], y[c + 1], q4);
          }
          p1 += q1; p2 += q2; p3 += q3; p4 += q4;
        }
        float4* xch = reinterpret_cast<float4*>(sXch) + (it & 1) * 4 * 128;
        xch[cq * 128 + row] = make_float4(p1, p2, p3, p4);
        pt.mark(1);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        pt.mark(2);
#pragma unroll
        for (int o = 1; o < 4; ++o) {
          const float4 v = xch[((cq + o) & 3) * 128 + row];
          p1 += v.x; p2 += v.y; p3 += v.z; p4 += v.w;
        }
        const float mu = p1 * (1.f / DI);
        const float rstd = rsqrtf(fmaxf(p2 * (1.f / DI) - mu * mu, 0.f) + 1e-5f);
        const float m1 = p3 * (1.f / DI);
        const float m2 = rstd * (p4 * (1.f / DI) - mu * m1);
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { y[cg * 8 + j] = (y[cg * 8 + j] - mu) * rstd; v[j] = y[cg * 8 + j]; }
          *reinterpret_cast<uint4*>(sCatRow + (2 * cq + cg) * 2048) = pack8(v);
        }
        pt.mark(3);
        bf16* drow = dact + (((long long)t * NA) * 128 + row) * 8;
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (gy[cg * 8 + j] - m1 - y[cg * 8 + j] * m2);
          const uint4 pk = pack8(o);
          *reinterpret_cast<uint4*>(sXYRow + (2 * cq + cg) * 2048) = pk;
          *reinterpret_cast<uint4*>(drow + (long long)(XC + 2 * cq + cg) * 1024) = pk;
        }
        // ---- dpre_z = alpha1 * g_z * SiLU'(pre_z)
        {
          float s0[8], s1[8], o0[8], o1[8];
          unpack8(sgz[0], s0);
          unpack8(sgz[1], s1);
#pragma unroll
          for (int j = 0; j < 8; ++j) { o0[j] = a1 * gz[j] * s0[j]; o1[j] = a1 * gz[8 + j] * s1[j]; }
          *reinterpret_cast<uint4*>(drow + (long long)(2 * cq) * 1024) = pack8(o0);
          *reinterpret_cast<uint4*>(drow + (long long)(2 * cq + 1) * 1024) = pack8(o1);
        }
        // ---- sum(dout) over tokens: this quarter's 8 columns
        {
          float v[8];
          unpack8(*reinterpret_cast<const uint4*>(sDoutRow + cq * 2048), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) sd[j] += v[j];
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&epi1_done[s]); mbar_arrive(&gy_free[s]); }
        pt.mark(4);
        if (t > T0) epi2(t - 1);
        sgc_prev = sgc_cur;
      }
      epi2(T1 - 1);
      pt.mark(6);
      // ---- flush Rt (TMEM lanes = rows j' of [yhat | zc], 32 columns d) and sum(dout) into this CTA's slab (no atomics:
      // one address would receive an update from every CTA); k_finalize_fast adds the slabs
      {
        float* slab = Rt + (long long)blockIdx.x * (2 * DI * D + D);
        float v[8];
        tmem_ld8(tbase + ((uint32_t)(q * 32) << 16) + B1_COL_RT + cq * 8, v);
        tmem_wait_ld();
        *reinterpret_cast<float4*>(slab + row * D + cq * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(slab + row * D + cq * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        // sum(dout): reduce the four lane quarters of this column quarter through shared memory (exchange buffer is free now)
        float* red = reinterpret_cast<float*>(sXch);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float w = warp_sum(sd[i]);
          if (lane == 0) red[(cq * 4 + q) * 8 + i] = w;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (q == 0 && lane < 8)
          slab[2 * DI * D + cq * 8 + lane] = red[(cq * 4 + 0) * 8 + lane] + red[(cq * 4 + 1) * 8 + lane] + red[(cq * 4 + 2) * 8 + lane] + red[(cq * 4 + 3) * 8 + lane];
      }
      if (!ok && lane == 0) atomicExch(status, 31);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_EPI_WARPS + 1) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// k_bwd2_ws (phase B2), per 128-token tile, after dS' is complete:
//   MMA1: G = Bc . dS' (hi + lo)
//   EPI1: w = softplus(dt + bias) * exp(A_log);  dpre_x = (D*dy + w*G) * SiLU'(pre_x) -> global;  wx = w*x -> sX (over x);
//         dw[h] = sum_{c in h} x*G;  ddt = dw * exp(A_log) * sigmoid(dt + bias) -> global;  dD, dA_log, ddt_bias partial sums
//   MMA2: dBc = wx . dS'^T (hi + lo)
//   EPI2: dpre_B = dBc * SiLU'(pre_B) -> global
// A column half owns channels [32h, 32h+32) = heads [8h, 8h+8) (headdim 4), so the halves never exchange anything.
// (8 epilogue warps: with 16 warps / column quarters this lighter kernel measured 5 % slower; k_bwd1_ws gained 14 %.)
// ------------------------------------------------------------------------------------------------
constexpr int B2_STG_B = (XC + CCH + XC + 2) * 2048;       // x | B | dy | dt
constexpr int B2_SMEM = 2 * B2_STG_B + 2 * SIMG_B;
constexpr int B2_TSTG = 96, B2_COL_DB = 64;

__global__ void __launch_bounds__(320, 1)
k_bwd2_ws(const bf16* __restrict__ act, const bf16* __restrict__ sgrad, const bf16* __restrict__ dtraw, const float* __restrict__ dS,
          const float* __restrict__ dt_bias, const float* __restrict__ A_log, const float* __restrict__ Dp,
          bf16* __restrict__ dact, bf16* __restrict__ ddt, float* __restrict__ head_part /* [CTA][dD | dA_log | ddt_bias] */,
          int tiles_per_batch, int num_tiles, int tiles_per_cta, int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[2], mma1_done[2], epi1_done[2], mma2_done[2], acc_free[2], g_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[NH], s_eA[NH], s_D[NH];
  __shared__ float s_red[8][24];
  uint8_t* sStg = smem;
  uint8_t* sImg = smem + 2 * B2_STG_B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T0 = blockIdx.x * tiles_per_cta, T1 = min(num_tiles, T0 + tiles_per_cta);
  if (tid < NH) { s_bias[tid] = dt_bias[tid]; s_eA[tid] = __expf(A_log[tid]); s_D[tid] = Dp[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 33);            // 32 producer lanes (state image) + 1 expect_tx arrival
      mbar_init(&mma1_done[i], 1);
      mbar_init(&epi1_done[i], 8);
      mbar_init(&mma2_done[i], 1);
      mbar_init(&acc_free[i], 8);        // dBc columns consumed (second epilogue)
      mbar_init(&g_free[i], 8);          // G columns consumed (first epilogue): MMA1 of tile t+2 does not wait for EPI2
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  bool ok = true;
  if (T0 < T1) {
    if (warp == 8) {
      int cur_b = -1;
      for (int t = T0; t < T1; ++t) {
        const int it = t - T0, s = it & 1, b = t / tiles_per_batch;
        uint8_t* sb = sStg + s * B2_STG_B;
        if (it >= 2) ok = mbar_wait(&mma2_done[s], ((it >> 1) - 1) & 1) && ok;
        if (b != cur_b) {
          stage_state_warp(dS + (long long)b * GN * DI, sImg + (b & 1) * SIMG_B, lane);
          cur_b = b;
        }
        if (lane == 0) {
          mbar_expect_tx(&full[s], (uint32_t)(XC + CCH + XC + 2) * 2048);
          bulk_g2s(sb, act + ((long long)t * NA + XC) * 1024, (XC + CCH) * 2048, &full[s]);                 // x | B
          bulk_g2s(sb + (XC + CCH) * 2048, dact + ((long long)t * NA + XC) * 1024, XC * 2048, &full[s]);   // dy (written by B1)
          bulk_g2s(sb + (XC + CCH + XC) * 2048, dtraw + (long long)t * 2 * 1024, 2 * 2048, &full[s]);      // dt columns
          if (t + 1 < T1) {     // L2 prefetch of the next tile (see k_bwd1_ws)
            const long long nt = t + 1;
            bulk_prefetch_l2(act + (nt * NA + XC) * 1024, (XC + CCH) * 2048);
            bulk_prefetch_l2(dact + (nt * NA + XC) * 1024, XC * 2048);
            bulk_prefetch_l2(dtraw + nt * 2 * 1024, 2 * 2048);
            bulk_prefetch_l2(sgrad + (nt * NA + XC) * 1024, (XC + CCH) * 2048);
          }
        }
        fence_async_smem();
        mbar_arrive(&full[s]);
      }
    } else if (warp == 9) {
      if (elect_one()) {
        const uint32_t sbase = smem_u32(sStg), ibase = smem_u32(sImg);
        const uint32_t id_g = make_idesc_rt(128, DI, false, false), id_b = make_idesc_rt(128, GN, false, false);
        auto mma2 = [&](int t) {
          const int it = t - T0, s = it & 1, b = t / tiles_per_batch;
          ok = mbar_wait(&epi1_done[s], (it >> 1) & 1) && ok;
          if (it >= 2) ok = mbar_wait(&acc_free[s], ((it >> 1) - 1) & 1) && ok;     // dBc of tile t-2 has been read
          tc_fence_after();
          const uint32_t tb = tbase + s * B2_TSTG, img = ibase + (b & 1) * SIMG_B;
          const uint64_t dX = make_desc(sbase + s * B2_STG_B, 2048, 128), dBh = make_desc(img + 8192, 512, 128), dBl = make_desc(img + 12288, 512, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k == 0) umma_c<false>(tb + B2_COL_DB, dX, dBh, id_b);
            else umma_c<true>(tb + B2_COL_DB, dadd(dX, k * 2 * 2048), dadd(dBh, k * 2 * 512), id_b);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_c<true>(tb + B2_COL_DB, dadd(dX, k * 2 * 2048), dadd(dBl, k * 2 * 512), id_b);
          umma_commit(&mma2_done[s]);
        };
        for (int t = T0; t < T1; ++t) {
          const int it = t - T0, s = it & 1, b = t / tiles_per_batch;
          ok = mbar_wait(&full[s], (it >> 1) & 1) && ok;
          if (it >= 2) ok = mbar_wait(&g_free[s], ((it >> 1) - 1) & 1) && ok;
          tc_fence_after();
          const uint32_t tb = tbase + s * B2_TSTG, img = ibase + (b & 1) * SIMG_B;
          const uint64_t dB = make_desc(sbase + s * B2_STG_B + XC * 2048, 2048, 128), dAh = make_desc(img, DI * 16, 128), dAl = make_desc(img + 4096, DI * 16, 128);
          umma_c<false>(tb, dB, dAh, id_g);
          umma_c<true>(tb, dadd(dB, 2 * 2048), dadd(dAh, 2 * DI * 16), id_g);
          umma_c<true>(tb, dB, dAl, id_g);
          umma_c<true>(tb, dadd(dB, 2 * 2048), dadd(dAl, 2 * DI * 16), id_g);
          umma_commit(&mma1_done[s]);
          if (t > T0) mma2(t - 1);
        }
        mma2(T1 - 1);
        if (!ok) atomicExch(status, 32);
      }
    } else {
      const int q = warp & 3, h = warp >> 2, row = q * 32 + lane;
      float aD[8], aA[8], aB[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { aD[i] = 0.f; aA[i] = 0.f; aB[i] = 0.f; }
      uint4 sgb_prev[2], sgb_cur[2];     // SiLU' of the B columns: fetched one tile ahead of their use in epi2
      auto epi2 = [&](int t) {
        const int it = t - T0, s = it & 1;
        ok = mbar_wait(&mma2_done[s], (it >> 1) & 1) && ok;
        tc_fence_after();
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(q * 32) << 16) + s * B2_TSTG + B2_COL_DB + h * 16, v);
        float s0[8], s1[8];
        unpack8(sgb_prev[0], s0);
        unpack8(sgb_prev[1], s1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free[s]);
        float o0[8], o1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o0[j] = v[j] * s0[j]; o1[j] = v[8 + j] * s1[j]; }
        bf16* drow = dact + (((long long)t * NA + 2 * XC + 2 * h) * 128 + row) * 8;
        *reinterpret_cast<uint4*>(drow) = pack8(o0);
        *reinterpret_cast<uint4*>(drow + 1024) = pack8(o1);
      };
      for (int t = T0; t < T1; ++t) {
        const int it = t - T0, s = it & 1;
        uint8_t* sb = sStg + s * B2_STG_B;
        uint8_t* sXRow = sb + row * 16;
        const uint8_t* sDyRow = sb + (XC + CCH) * 2048 + row * 16;
        const uint8_t* sDtRow = sb + (XC + CCH + XC) * 2048 + row * 16;
        uint4 sgx[4];
        {
          const bf16* srow = sgrad + (((long long)t * NA + XC + 4 * h) * 128 + row) * 8;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) sgx[c4] = __ldg(reinterpret_cast<const uint4*>(srow + c4 * 1024));
          const bf16* brow = sgrad + (((long long)t * NA + 2 * XC + 2 * h) * 128 + row) * 8;
          sgb_cur[0] = __ldg(reinterpret_cast<const uint4*>(brow));
          sgb_cur[1] = __ldg(reinterpret_cast<const uint4*>(brow + 1024));
        }
        ok = mbar_wait(&full[s], (it >> 1) & 1) && ok;     // the dt columns are read before MMA1 completes
        float w[8], sg[8];
        {
          float v[8];
          unpack8(*reinterpret_cast<const uint4*>(sDtRow + h * 2048), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = v[j] + s_bias[8 * h + j];
            w[j] = softplus_fast(a) * s_eA[8 * h + j];
            sg[j] = a > 20.f ? 1.f : __fdividef(1.f, 1.f + __expf(-a));
          }
        }
        ok = mbar_wait(&mma1_done[s], (it >> 1) & 1) && ok;
        tc_fence_after();
        const uint32_t ta = tbase + ((uint32_t)(q * 32) << 16) + s * B2_TSTG;
        float dw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dw[i] = 0.f;
        bf16* drow = dact + (((long long)t * NA) * 128 + row) * 8;
#pragma unroll
        for (int cb = 0; cb < 32; cb += 16) {
          float G[16];
          tmem_ld16(ta + h * 32 + cb, G);
          tmem_wait_ld();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cl = cb / 8 + half;          // chunk inside this half (0..3): heads 2*cl, 2*cl + 1 of the half
            float x[8], dy[8], o[8], wx[8], sv[8];
            unpack8(*reinterpret_cast<const uint4*>(sXRow + (4 * h + cl) * 2048), x);
            unpack8(*reinterpret_cast<const uint4*>(sDyRow + (4 * h + cl) * 2048), dy);
            unpack8(sgx[cl], sv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int hh = 2 * cl + (j & 1);
              const float g = G[half * 8 + j];
              o[j] = (s_D[8 * h + hh] * dy[j] + w[hh] * g) * sv[j];
              wx[j] = w[hh] * x[j];
              dw[hh] = fmaf(x[j], g, dw[hh]);
              aD[hh] = fmaf(dy[j], x[j], aD[hh]);
            }
            *reinterpret_cast<uint4*>(sXRow + (4 * h + cl) * 2048) = pack8(wx);
            *reinterpret_cast<uint4*>(drow + (long long)(XC + 4 * h + cl) * 1024) = pack8(o);
          }
        }
        {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = dw[j] * s_eA[8 * h + j] * sg[j];
            aA[j] = fmaf(dw[j], w[j], aA[j]);
            aB[j] += o[j];
          }
          *reinterpret_cast<uint4*>(ddt + (((long long)t * 2 + h) * 128 + row) * 8) = pack8(o);
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&epi1_done[s]); mbar_arrive(&g_free[s]); }
        if (t > T0) epi2(t - 1);
        sgb_prev[0] = sgb_cur[0];
        sgb_prev[1] = sgb_cur[1];
      }
      epi2(T1 - 1);
      // per-head parameter gradients: warp sums -> shared memory -> one slab of 3*NH floats per CTA (no atomics: every CTA
      // would hit the same 48 addresses); k_finalize_fast adds the slabs
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float vD = warp_sum(aD[j]), vA = warp_sum(aA[j]), vB = warp_sum(aB[j]);
        if (lane == 0) { s_red[warp][j] = vD; s_red[warp][8 + j] = vA; s_red[warp][16 + j] = vB; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (q == 0 && lane < 24) {
        const float v = s_red[4 * h + 0][lane] + s_red[4 * h + 1][lane] + s_red[4 * h + 2][lane] + s_red[4 * h + 3][lane];
        head_part[(long long)blockIdx.x * 3 * NH + (lane >> 3) * NH + 8 * h + (lane & 7)] = v;
      }
      if (!ok && lane == 0) atomicExch(status, 33);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 256);
}

}  // namespace bwdws
