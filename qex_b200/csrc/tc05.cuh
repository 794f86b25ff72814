// Blackwell (sm_100a) 5th-generation tensor-core primitives used by the FP32 network path: tcgen05.mma
// (kind::tf32, operands in shared memory, accumulators in tensor memory), TMEM allocation / loads, mbarriers
// and the shared-memory matrix descriptors of the 128-byte-swizzled canonical layouts.
//
// Storage convention used everywhere in this library ("plane"): a [R rows x Ccols] float matrix, R % 8 == 0,
// Ccols % 32 == 0, is stored as Ccols/32 column atoms; atom a holds columns [32a, 32a+32) of all rows as
// R rows x 128 bytes, rows consecutive, and inside each 8-row x 128-byte block (1024 bytes, 1024-aligned) the
// 16-byte chunk index is XOR-ed with (row & 7)  (the hardware's SWIZZLE_128B pattern).  The SAME bytes are
// a valid K-major operand (rows = M or N index, columns = K) and a valid MN-major operand (columns = M or N
// index, rows = K) -- which is what lets one set of activation planes feed the forward product, the
// back-propagation product and the weight-gradient product.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qexxc {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) of a plane with R rows
__device__ __host__ __forceinline__ uint32_t plane_off(int r, int c, int R) {
    const int a = c >> 5, cc = c & 31;
    return (uint32_t)a * (uint32_t)R * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
           (uint32_t)((((cc >> 2) ^ (r & 7)) << 4) + ((cc & 3) << 2));
}

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// true in exactly one lane of a converged warp (the pattern ptxas recognises for single-thread issue of UTC* instructions
// from warp-uniform code: no per-lane "waterfall" loop around the uniform-register operands)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "@P mov.s32 %0, 1;\n\t"
        "}"
        : "+r"(pred));
    return pred != 0;
}
// warp index as a provably warp-uniform value
__device__ __forceinline__ int warp_uniform_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- proxies / fences -------------------------------------------------------------------------------------------
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory ----------------------------------------------------------------------------------------------
// one full warp; writes the base address (lane 0, column c0) to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) {
    return base + ((uint32_t)lane << 16) + (uint32_t)col;
}
// warp w reads lanes 32*(w%4) .. +31 (one row per thread), 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// same, 8 consecutive 32-bit columns, raw words (several loads can be in flight before one tmem_ld_wait)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// zero 16 consecutive 32-bit columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).  Offsets in bytes (multiples of 16).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// instruction descriptor of kind::tf32 with FP32 accumulation
__device__ __host__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major operand: rows = M/N index, K along the columns of the plane; one instruction consumes 8 columns (32 bytes)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t plane_saddr, int R, int kstep) {
    const uint32_t a = plane_saddr + (uint32_t)(kstep >> 2) * (uint32_t)R * 128u + (uint32_t)(kstep & 3) * 32u;
    return smem_desc(a, 16, 1024);
}
// MN-major operand: columns of the plane = M/N index (atoms of 32 at stride R*128), rows = K; one instruction
// consumes 8 rows (one 1024-byte block row)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t plane_saddr, int R, int kstep) {
    return smem_desc(plane_saddr + (uint32_t)kstep * 1024u, (uint32_t)R * 128u, 1024);
}

// split an FP32 value into a TF32-exact head and the remainder (3xTF32 products recover ~21 mantissa bits)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

}  // namespace tc05
}  // namespace qexxc

// ---- 16-bit (bf16) operands -------------------------------------------------------------------------------------
// MN-major TF32 operands exist only in a different swizzle (SWIZZLE_128B_BASE32B), so a TF32 plane cannot serve
// both views.  16-bit planes can: a [R rows x 64] bf16 matrix stored as R rows x 128 bytes with the 16-byte chunk
// index XOR-ed with (row & 7) is a K-major operand (rows = M/N, columns = K, 16 columns = 32 bytes per
// instruction) AND an MN-major operand (columns = M/N, rows = K, 16 rows = two 1024-byte blocks per instruction).
namespace qexxc {
namespace tc05 {

__device__ __host__ __forceinline__ uint32_t plane16_off(int r, int c) {  // byte offset of element (r, c), c < 64
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
}
__device__ __host__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major: one instruction consumes 16 columns (32 bytes); kstep = 0..3 for a 64-column plane
__device__ __forceinline__ uint64_t desc16_k(uint32_t plane_saddr, int kstep) {
    return smem_desc(plane_saddr + (uint32_t)kstep * 32u, 16, 1024);
}
// MN-major: one instruction consumes 16 rows (2048 bytes); the next 64 M/N elements are the next plane (stride R*128)
__device__ __forceinline__ uint64_t desc16_mn(uint32_t plane_saddr, int R, int kstep) {
    return smem_desc(plane_saddr + (uint32_t)kstep * 2048u, (uint32_t)R * 128u, 1024);
}
// x = b1 + b2 + b3 (+ O(2^-24 x)), each part exactly representable in bf16 (truncation splits are exact)
__device__ __forceinline__ void split_bf16x3(float x, uint32_t& b1, uint32_t& b2, uint32_t& b3) {
    b1 = __float_as_uint(x) & 0xFFFF0000u;
    const float r1 = x - __uint_as_float(b1);
    b2 = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = r1 - __uint_as_float(b2);
    b3 = __float_as_uint(r2) & 0xFFFF0000u;
}
// pack the high halves of two such words: lo 16 bits <- a, hi 16 bits <- b
__device__ __forceinline__ uint32_t pack_hi16(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }

}  // namespace tc05
}  // namespace qexxc

// ---- 8-bit integer operands (kind::i8, INT32 accumulation): the exact Ozaki-split contractions (contract_i8.cu) --------------------
namespace qexxc {
namespace tc05 {
__device__ __host__ constexpr uint32_t idesc_i8(int M, int N, int a_mn_major = 0) {  // signed 8-bit A and B, B K-major, S32 accumulators
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, the two 64-bit shared-memory descriptors given as (low word, common high word)
__device__ __forceinline__ void mma_i8_lohi(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
}  // namespace tc05
}  // namespace qexxc
