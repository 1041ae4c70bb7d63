// Blackwell (sm_100a) tensor-core primitives used by the fp32-accurate contraction kernels:
// tcgen05.mma kind::tf32 with accumulators in TMEM, operands in shared memory in the UMMA
// K-major / 128-byte-swizzle canonical layout, mbarrier completion, tcgen05.ld epilogues.
//
// Precision: every fp32 operand x is split as x = hi + lo with hi = rna_tf32(x),
// lo = rna_tf32(x - hi); a product a*b is evaluated as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
// ("3xTF32") with fp32 accumulation in TMEM.  The dropped a_lo*b_lo term and the roundings are
// <= 2^-22 relative per product -- inside the 1e-5 parity budget with two orders of margin.
#pragma once

#include "common.cuh"

namespace ihg {
namespace tc {

constexpr int kTileM = 128;          // rows per UMMA tile (M = 128, cta_group::1)
constexpr int kChunkK = 32;          // fp32 elements per 128-byte swizzle row (one K chunk)
constexpr int kChunkBytesPerRow = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- operand split ---------------------------------------------------------------------
// Round-to-nearest (ties away from zero, == cvt.rna.tf32.f32) to the 10 explicit mantissa bits of
// tf32, done with integer add + mask on the full-rate ALU pipe: the cvt instruction runs on the
// 16-lane conversion pipe and was measured to dominate the operand producers (an in-kernel
// clock64 trace showed ~900 cycles per 128x32 stage, ~60 % of it in cvt).
__device__ __forceinline__ uint32_t rna_tf32(float x) {
    return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - __uint_as_float(hi));
}

// Byte offset of element (row, 16-byte chunk c in [0,8)) inside one [rows x 128 B] operand
// tile in the K-major SWIZZLE_128B canonical layout (tile base 1024-byte aligned): rows are
// grouped by 8 (1024 B per group, SBO), the 16-byte chunk index is XORed with row % 8.
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}

// Store 4 consecutive K elements (one 16-byte chunk) of a row, split into hi / lo tiles.
__device__ __forceinline__ void store_split_chunk(uint32_t hi_tile, uint32_t lo_tile, int row,
                                                  int chunk16, const float4& v) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32(v.x, h0, l0);
    split_tf32(v.y, h1, l1);
    split_tf32(v.z, h2, l2);
    split_tf32(v.w, h3, l3);
    const uint32_t off = sw128_offset(row, chunk16);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hi_tile + off), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lo_tile + off), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
//  [46,48), layout_type=2 [61,64).)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major): 16 B
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// Advance a K-major SW128 descriptor by `k_elems` fp32 elements inside the 128-byte row.
__device__ __forceinline__ uint64_t advance_desc_k(uint64_t desc, int k_elems) {
    return desc + (uint64_t)((k_elems * 4) >> 4);
}

// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M = 128.
// (cute::UMMA::InstrDescriptor: c_format [4,6)=1 F32, a_format [7,10)=2 TF32, b_format
//  [10,13)=2, a_major [15]=0, b_major [16]=0, n_dim [17,23)=N>>3, m_dim [24,29)=M>>4.)
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// MN-major tf32 operands must use the SWIZZLE_128B_BASE32B canonical layout (the only MN-major
// layout the tensor core accepts for 32-bit types): rows of 128 B (32 consecutive MN elements)
// per K index, K atoms of 4 rows (512 B), the 32-byte chunk index XORed with row % 4.
__device__ __forceinline__ uint32_t sw128b32_offset(int row, int chunk16) {
    return (uint32_t)(row * 128 + ((((chunk16 >> 1) ^ (row & 3)) << 5) | ((chunk16 & 1) << 4)));
}
__device__ __forceinline__ void store_split_chunk_mn(uint32_t hi_tile, uint32_t lo_tile, int row,
                                                     int chunk16, const float4& v) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32(v.x, h0, l0);
    split_tf32(v.y, h1, l1);
    split_tf32(v.z, h2, l2);
    split_tf32(v.w, h3, l3);
    const uint32_t off = sw128b32_offset(row, chunk16);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hi_tile + off), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lo_tile + off), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
}
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // stride between 32-feature MN blocks
    d |= (uint64_t)(512 >> 4) << 32;                     // stride between 4-edge K atoms
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int n) {
    return make_idesc_tf32(n) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}


// ---- tcgen05 wrappers -------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo for one K step of 8
__device__ __forceinline__ void mma_3xtf32(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo,
                                           uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
    mma_tf32(tmem_d, a_lo, b_hi, idesc, accumulate);   // small terms first
    mma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
    mma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
}
// ---- MMA issue discipline -----------------------------------------------------------------
// The issuing warp must run its loop in warp-uniform control flow and guard only the asm with
// elect_one(): under `if (lane == 0)` ptxas cannot prove the operands uniform and wraps EVERY
// tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall, which measured 57-186 cycles
// per MMA against a 32-cycle tensor floor (profiles/microbench_mma2.cu: 32.0 cycles with this
// shape).  Values loaded from shared memory (the TMEM base) are made provably uniform with
// warp_uniform() so they live in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }

// D[tmem] (+)= A[tmem] . B[smem descriptor]: A operand read from tensor memory (lane = row,
// one 32-bit column per K element), one K step of 8
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// commit that arrives on the barrier at the same shared-memory offset in every CTA of `cta_mask`
__device__ __forceinline__ void mma_commit_multicast(uint32_t mbar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar), "h"(cta_mask) : "memory");
}
// ---- thread-block clusters ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> shared bulk copy delivered to the same offset in every CTA of `cta_mask`, completing
// `bytes` on each destination CTA's mbarrier at the same offset
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar,
                                                   uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(mbar), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
    return n <= 32 ? 32u : (n <= 64 ? 64u : (n <= 128 ? 128u : (n <= 256 ? 256u : 512u)));
}

// 32 lanes (this warp's TMEM quadrant) x 16 consecutive fp32 columns -> 16 registers/thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Per-warp staging tile for the epilogues: 32 rows x 128 B, 16-byte chunks XOR-swizzled by
// row & 7.  Written thread-per-row (TMEM order), read with 8 lanes per row (coalesced global
// access: one full 128-byte line per row instead of 32 partial lines per request).
constexpr int kEpiStageBytes = 32 * 128;
__device__ __forceinline__ uint32_t epi_off(int row, int chunk) {
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void sts4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}


// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}

}  // namespace tc
}  // namespace ihg
