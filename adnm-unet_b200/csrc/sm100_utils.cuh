// Blackwell (sm_100a) building blocks used by the fused ADN-SSD kernels: tcgen05.mma with shared-memory operand
// descriptors, TMEM allocation / read-back, mbarrier completion tracking, proxy fences.
//
// Shared-memory operand layout used everywhere ("T8", no swizzle): a tile of R rows x C columns of bf16 is stored
// as  [C/8][R][8]  i.e. element (r, c) at ((c >> 3) * R + r) * 8 + (c & 7).  Each 8x8 block (8 consecutive rows of
// one 8-column chunk) is one contiguous 128-byte UMMA "core matrix".  The same bytes serve
//   * as a K-major operand   (rows = M or N index, columns = K):   LBO = R*16 (next 8 K-columns), SBO = 128
//   * as an MN-major operand (columns = M or N index, rows = K):   SBO = R*16 (next 8 MN-columns), LBO = 128
// so one activation tile [tokens x channels] feeds both the per-token GEMMs (tokens are M) and the reductions
// over tokens (tokens are K).  Descriptor bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor) of the CUTLASS tree vendored with the image; canonical layouts: cute/atom/mma_traits_sm100.hpp.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace adn {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element offset inside a T8 tile of R rows
__device__ __forceinline__ int t8_off(int r, int c, int R) { return (((c >> 3) * R + r) << 3) + (c & 7); }

// ---- shared memory matrix descriptor (SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// K-major operand: tile base (element (0,0)), R rows in the tile, first row r0 (multiple of 8), first K column k0 (multiple of 8)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int R, int r0, int k0) {
  return make_desc(tile_saddr + (uint32_t)(((k0 >> 3) * R + r0) * 16), (uint32_t)R * 16, 128);
}
// MN-major operand: MN index = tile column c0.., K index = tile row (token) t0..
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int R, int c0, int t0) {
  return make_desc(tile_saddr + (uint32_t)(((c0 >> 3) * R + t0) * 16), 128, (uint32_t)R * 16);
}

// Advance a descriptor's start address by `bytes` (the 14-bit start-address field is in 16-byte units): consecutive
// K steps of one operand differ by a constant, so the issuing thread adds instead of rebuilding the descriptor.
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// ---- instruction descriptor, kind::f16, bf16 x bf16 -> fp32
template <int M, int N, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  static_assert(M == 64 || M == 128, "UMMA M");
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N (M=128 needs N % 16 == 0)");
  return (1u << 4)                       // D format: F32
         | (1u << 7) | (1u << 10)        // A, B format: BF16
         | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t make_idesc_rt(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a fault in the async pipe must not hang the GPU box.  Returns false on time-out (~seconds).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ---- bulk asynchronous copy global -> shared (one instruction moves `bytes`, a multiple of 16; completion is signalled
// on an mbarrier as a byte count).  Issued by ONE thread after mbar_expect_tx() for the same barrier.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of a contiguous global range (multiple of 16 bytes); fire-and-forget, issued by ONE thread
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---- tiled global layout "TL" of the internal activation tensors: tokens in tiles of 128, inside a tile the T8 layout
// [chunk][128 rows][8]: element (tok, c) of a tensor with NCH = channels/8 chunks.  A tile's chunk range is contiguous,
// so it moves global -> shared with ONE bulk copy, and "thread t owns token row t" kernels store fully coalesced.
__host__ __device__ __forceinline__ long long tl_off(long long tok, int c, int NCH) {
  return (((tok >> 7) * NCH + (c >> 3)) * 128 + (tok & 127)) * 8 + (c & 7);
}
// 32-bit token index variant for hot loops (tokens < 2^31 is validated at the ABI): one 64-bit multiply-add
__device__ __forceinline__ long long tl_off32(int tok, int c, int NCH) {
  return (long long)((tok >> 7) * NCH + (c >> 3)) * 1024 + (((tok & 127) << 3) + (c & 7));
}

// ---- fences
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM
// called by one full warp; writes the base address to *slot (shared memory); ncols: power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, x16 columns: thread t of the warp gets row (lane base + t), columns [col, col+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// TMEM address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) {
  return base + ((uint32_t)lane << 16) + (uint32_t)col;
}

// ---- packing helpers
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16(uint32_t u, float& a, float& b) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  float2 f = __bfloat1622float2(h);
  a = f.x; b = f.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  unpack_bf16(u.x, v[0], v[1]); unpack_bf16(u.y, v[2], v[3]); unpack_bf16(u.z, v[4], v[5]); unpack_bf16(u.w, v[6], v[7]);
}

}  // namespace sm100
}  // namespace adn
