// Dev microbenchmark: issue rate of tcgen05.mma kind::tf32 (M = 128) chains on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ihgnn_b200/csrc -I include profiles/microbench_mma.cu -o build/microbench_mma
// Prints cycles per MMA for: A from shared memory (SS) / A from tensor memory (TS), N in
// {64,128,256}, 1..3 accumulators used round-robin (dependent vs independent accumulation).
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

// variant: the whole warp runs the loop (uniform control flow, operands in uniform registers),
// one elected lane issues
__global__ void __launch_bounds__(128, 1) k_uniform(int ts, int n, int n_acc, int iters, int a_rot, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(raw)[i] = 0x3f800000u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_tf32(n);
        const uint64_t da = make_kmajor_sw128_desc(base);
        const uint64_t db = make_kmajor_sw128_desc(base + 16384);
        long long t0 = clock64();
        int acc_i = 0;
        for (int i = 0; i < iters; ++i) {
            const int ks = i & 3;
            const uint32_t d = tm + (uint32_t)(acc_i * n);
            if (++acc_i == n_acc) acc_i = 0;
            if (elect_one()) {
                if (ts) mma_ts(d, tm + 448u + 8u * (a_rot ? ks : 0), advance_desc_k(db, 8 * ks), idesc, i >= n_acc);
                else mma_tf32(d, advance_desc_k(da, 8 * (a_rot ? ks : 0)), advance_desc_k(db, 8 * ks), idesc, i >= n_acc);
            }
            __syncwarp();
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

__global__ void __launch_bounds__(128, 1) k(int ts, int n, int n_acc, int iters, int a_rot, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(raw)[i] = 0x3f800000u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(n);
        const uint64_t da = make_kmajor_sw128_desc(base);
        const uint64_t db = make_kmajor_sw128_desc(base + 16384);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int ks = i & 3;
            const uint32_t d = tm + (uint32_t)((i % n_acc) * n);
            if (ts) mma_ts(d, tm + 448u + 8u * (a_rot ? ks : 0), advance_desc_k(db, 8 * ks), idesc, i >= n_acc);
            else mma_tf32(d, advance_desc_k(da, 8 * (a_rot ? ks : 0)), advance_desc_k(db, 8 * ks), idesc, i >= n_acc);
        }
        mma_commit(smem_u32(&bar));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int iters = 4096;
    cudaFuncSetAttribute(k_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    for (int variant = 0; variant < 2; ++variant)
    for (int grid : {1, 148})
        for (int ts = 0; ts < 2; ++ts)
            for (int n : {64, 128, 256})
                for (int n_acc = 1; n_acc <= 3; ++n_acc) {
                    if (n_acc * n > 384) continue;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (variant) k_uniform<<<grid, 128, 66 * 1024>>>(ts, n, n_acc, iters, 1, out);
                        else k<<<grid, 128, 66 * 1024>>>(ts, n, n_acc, iters, 1, out);
                        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    }
                    printf("%s grid=%3d %s N=%3d accumulators=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", variant ? "elect  " : "lane0  ", grid,
                           ts ? "TS" : "SS", n, n_acc, (double)out[0] / iters, (double)out[1] / iters, n / 2);
                }
    return 0;
}
