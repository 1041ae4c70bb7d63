// Dev microbenchmark: how fast can W warps of one CTA per SM gather random 128-byte row slices
// into shared memory?  Variants: cp.async 16 B (with / without the zero-fill operand),
// ids preloaded or loaded in the loop, 4..16 warps; 128-byte cp.async.bulk per row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ihgnn_b200/csrc -I include profiles/microbench_gather.cu -o build/microbench_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace ihg::tc;

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool ok) {
    const uint32_t n = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void bulk128(uint32_t dst, const void* src, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(dst), "l"(src), "r"(mbar) : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}

// granule = 384 rows (128 edges x 3) x 128 B; mode 0: cp16, 1: cp16 zfill, 2: cp16 + ids loaded in loop (two halves),
// 3: bulk 128 B per row (one lane per row)
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ table, int64_t ld, const int* __restrict__ ids,
                                            int granules, int mode, int warps, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar[2];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); mbar_init_fence(); }
    __syncthreads();
    if (tid >= warps * 32) return;
    const int nt = warps * 32;
    const int per = 3072 / nt;                    // 16-byte copies per thread per granule
    const int chk = tid & 7, row0 = tid >> 3, rstep = nt >> 3;
    const int* my = ids + (int64_t)blockIdx.x * granules * 384;
    long long t0 = clock64();
    if (mode <= 2) {
        for (int g = 0; g < granules; ++g) {
            const uint32_t buf = base + (g & 1) * 49152;
            const int* gi = my + g * 384;
            if (mode == 2) {
                for (int h = 0; h < 2; ++h) {
                    int id[12];
#pragma unroll
                    for (int j = 0; j < 12; ++j) id[j] = (j + 12 * h) < per ? __ldg(gi + row0 + rstep * (j + 12 * h)) : 0;
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        if (j + 12 * h < per) {
                            const int r = row0 + rstep * (j + 12 * h);
                            cp16z(buf + r * 128 + ((chk ^ (r & 7)) << 4), table + (int64_t)id[j] * ld + 4 * chk, id[j] >= 0);
                        }
                }
            } else {
                for (int j = 0; j < per; ++j) {
                    const int r = row0 + rstep * j;
                    const int id = __ldg(gi + r);
                    if (mode == 0) cp16(buf + r * 128 + ((chk ^ (r & 7)) << 4), table + (int64_t)id * ld + 4 * chk);
                    else cp16z(buf + r * 128 + ((chk ^ (r & 7)) << 4), table + (int64_t)id * ld + 4 * chk, id >= 0);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");     // keep one granule in flight
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        // bulk: lane per row, rows padded to 144 B
        uint32_t ph[2] = {0, 0};
        for (int g = 0; g < granules; ++g) {
            const int b = g & 1;
            const uint32_t buf = base + b * 55296;
            const int* gi = my + g * 384;
            if (g >= 2) { if (tid == 0) mbar_wait(smem_u32(&bar[b]), ph[b]); ph[b] ^= 1u; }
            asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
            if (tid == 0) expect_tx(smem_u32(&bar[b]), 384 * 128);
            asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
            for (int r = tid; r < 384; r += nt) bulk128(buf + r * 144, table + (int64_t)__ldg(gi + r) * ld, smem_u32(&bar[b]));
        }
        if (tid == 0) {
            for (int b = 0; b < 2; ++b) if (granules > b) mbar_wait(smem_u32(&bar[b]), ph[b]);
        }
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    const int64_t rows = 261000, ld = 64;      // 67 MB table (amazon-full node table)
    float* table; int* ids; long long* out;
    cudaMalloc(&table, rows * ld * 4);
    cudaMemset(table, 0, rows * ld * 4);
    const int granules = 64;
    std::vector<int> h((size_t)148 * granules * 384);
    srand(1);
    for (auto& v : h) v = (int)(((int64_t)rand() * 7919 + rand()) % rows);
    cudaMalloc(&ids, h.size() * 4);
    cudaMemcpy(ids, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    const char* names[4] = {"cp.async 16B", "cp.async 16B zfill", "cp.async zfill, ids in 2 halves", "cp.async.bulk 128B/row"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {2, 4, 8, 16}) {
            if (mode == 2 && warps != 4) continue;
            for (int rep = 0; rep < 2; ++rep) {
                k<<<148, 512, 120 * 1024>>>(table, ld, ids, granules, mode, warps, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            }
            printf("%-34s warps=%2d: %6.0f cycles per 48 KB granule (%.1f B/clk/SM)\n", names[mode], warps,
                   (double)out[0] / granules, 49152.0 * granules / out[0]);
        }
    return 0;
}
