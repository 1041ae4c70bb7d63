// Dev microbenchmark 2: how the issue loop is written decides the tcgen05.mma rate.
// Models the chunk loop of the interaction kernels (per chunk: 4 K steps x 3 MMAs, A stage
// and W stage rings) in three code shapes:
//   0  "lane0": the loop runs under if (lane == 0)         (divergent: per-MMA ELECT/R2UR loops)
//   1  "elect": whole warp runs the loop, elect.sync guards the asm, operands from a
//               redux.sync'ed (uniform) TMEM base
//   2  "elect+ptx": same, but the 12 MMAs of a chunk are one asm block that derives the
//               operands with PTX adds from 4 inputs (fewer register->uniform moves)
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace ihg::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// one chunk = 4 K steps x (a_lo*b_hi, a_hi*b_lo, a_hi*b_hi); a_lo = a_hi + 32 columns,
// b_lo descriptor = b_hi + lo_off (16-byte units); first MMA overwrites when acc0 == 0
__device__ __forceinline__ void mma_chunk_ts(uint32_t d, uint32_t a_hi, uint64_t b_hi, uint32_t lo_off, uint32_t idesc, uint32_t acc0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 al, ah;\n\t.reg .b64 bh, bl, off;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "cvt.u64.u32 off, %3;\n\t"
        "mov.b32 ah, %1;\n\tadd.u32 al, %1, 32;\n\tmov.b64 bh, %2;\n\tadd.u64 bl, %2, off;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %4, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %4, 1;\n\t"
        "add.u32 ah, ah, 8;\n\tadd.u32 al, al, 8;\n\tadd.u64 bh, bh, 2;\n\tadd.u64 bl, bl, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %4, 1;\n\t"
        "add.u32 ah, ah, 8;\n\tadd.u32 al, al, 8;\n\tadd.u64 bh, bh, 2;\n\tadd.u64 bl, bl, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %4, 1;\n\t"
        "add.u32 ah, ah, 8;\n\tadd.u32 al, al, 8;\n\tadd.u64 bh, bh, 2;\n\tadd.u64 bl, bl, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %4, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %4, 1;\n\t"
        "}\n" ::"r"(d), "r"(a_hi), "l"(b_hi), "r"(lo_off), "r"(idesc), "r"(acc0)
        : "memory");
}

template <int kShape>
__global__ void __launch_bounds__(128, 1) k(int n, int chunks, int a_stages, int w_stages, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(raw)[i] = 0x3f800000u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = slot;
    const uint32_t idesc = make_idesc_tf32(n);
    const uint32_t b_tile = (uint32_t)n * 128u, w_stage = 2 * b_tile;
    if (kShape == 0) {
        if (threadIdx.x == 0) {
            long long t0 = clock64();
            for (int it = 0; it < chunks; ++it) {
                const int sa = it % a_stages, sw = it % w_stages;
                const uint32_t a_hi = tm + 256u + 64u * sa, a_lo = a_hi + 32u;
                const uint64_t dbh = make_kmajor_sw128_desc(base + sw * w_stage), dbl = make_kmajor_sw128_desc(base + sw * w_stage + b_tile);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    mma_ts(tm, a_lo + 8u * ks, advance_desc_k(dbh, 8 * ks), idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    mma_ts(tm, a_hi + 8u * ks, advance_desc_k(dbl, 8 * ks), idesc, 1u);
                    mma_ts(tm, a_hi + 8u * ks, advance_desc_k(dbh, 8 * ks), idesc, 1u);
                }
            }
            mma_commit(smem_u32(&bar));
            long long t1 = clock64();
            mbar_wait(smem_u32(&bar), 0);
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (threadIdx.x < 32) {
        const uint32_t tmu = __reduce_or_sync(0xffffffffu, tm);          // uniform register copy
        long long t0 = clock64();
        int sa = 0, sw = 0;
        for (int it = 0; it < chunks; ++it) {
            const uint32_t a_hi = tmu + 256u + 64u * sa, a_lo = a_hi + 32u;
            const uint64_t dbh = make_kmajor_sw128_desc(base + sw * w_stage);
            if (kShape == 1) {
                const uint64_t dbl = make_kmajor_sw128_desc(base + sw * w_stage + b_tile);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        mma_ts(tmu, a_lo + 8u * ks, advance_desc_k(dbh, 8 * ks), idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        mma_ts(tmu, a_hi + 8u * ks, advance_desc_k(dbl, 8 * ks), idesc, 1u);
                        mma_ts(tmu, a_hi + 8u * ks, advance_desc_k(dbh, 8 * ks), idesc, 1u);
                    }
                }
            } else {
                if (elect_one()) mma_chunk_ts(tmu, a_hi, dbh, b_tile >> 4, idesc, it > 0 ? 1u : 0u);
            }
            __syncwarp();
            if (++sa == a_stages) sa = 0;
            if (++sw == w_stages) sw = 0;
        }
        if (elect_one()) mma_commit(smem_u32(&bar));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    const int smem = 201 * 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int chunks = 1024;
    const char* names[3] = {"lane0    ", "elect    ", "elect+ptx"};
    for (int shape = 0; shape < 3; ++shape)
        for (int n : {64, 128}) {
            const int a_stages = n == 64 ? 4 : 4, w_stages = n == 64 ? 6 : 3;
            for (int rep = 0; rep < 2; ++rep) {
                if (shape == 0) k<0><<<148, 128, smem>>>(n, chunks, a_stages, w_stages, out);
                if (shape == 1) k<1><<<148, 128, smem>>>(n, chunks, a_stages, w_stages, out);
                if (shape == 2) k<2><<<148, 128, smem>>>(n, chunks, a_stages, w_stages, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            }
            printf("%s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (tensor floor %d)\n", names[shape], n,
                   (double)out[0] / (12 * chunks), (double)out[1] / (12 * chunks), n / 2);
        }
    return 0;
}
