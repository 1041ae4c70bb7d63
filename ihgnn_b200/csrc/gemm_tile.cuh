// Shared-memory tiled fp32 FFMA micro-kernels used by the SIMT (exact-fp32) contraction paths:
// the node-level Linear layers and the per-hyperedge interaction contraction.
//
// Block = 256 threads viewed as 16 x 16: thread (tx, ty) owns rows ty*4 .. ty*4+3 of a 64-row
// tile and the strided columns tx + 16*j (j < DPT).  K is streamed through shared memory in
// chunks of kKC = 32.  All accumulation is plain fp32 FFMA in a fixed order (deterministic).
#pragma once

#include "common.cuh"

namespace ihg {

constexpr int kTileRows = 64;   // rows (edges / nodes) per block tile
constexpr int kKC = 32;         // K chunk
constexpr int kAtPitch = 68;    // pitch of the transposed A chunk At[k][row] (16B-aligned rows)
constexpr int kGemmThreads = 256;

// acc[r][j] += sum_k At[k][ty*4+r] * Bs[k][tx+16j]
template <int DPT>
__device__ __forceinline__ void tile_fma_at(const float* __restrict__ At, const float* __restrict__ Bs,
                                            int b_pitch, int klen, int tx, int ty,
                                            float (&acc)[4][DPT]) {
#pragma unroll 4
    for (int k = 0; k < klen; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(At + k * kAtPitch + ty * 4);
        float b[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) b[j] = Bs[k * b_pitch + tx + 16 * j];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
            acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
            acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
            acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
        }
    }
}

// Outer-product accumulation for weight gradients:
//   acc[rg][r][j] += sum_{e<rows} As[e][ty*4 + r + 64*rg] * Bs[e][tx + 16j]
// As: [rows][a_pitch] (a_pitch % 4 == 0), Bs: [rows][b_pitch].
template <int RG, int DPT>
__device__ __forceinline__ void tile_outer(const float* __restrict__ As, int a_pitch,
                                           const float* __restrict__ Bs, int b_pitch, int rows,
                                           int tx, int ty, float (&acc)[RG][4][DPT]) {
#pragma unroll 2
    for (int e = 0; e < rows; ++e) {
        float b[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) b[j] = Bs[e * b_pitch + tx + 16 * j];
#pragma unroll
        for (int rg = 0; rg < RG; ++rg) {
            const float4 a = *reinterpret_cast<const float4*>(As + e * a_pitch + ty * 4 + 64 * rg);
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                acc[rg][0][j] = fmaf(a.x, b[j], acc[rg][0][j]);
                acc[rg][1][j] = fmaf(a.y, b[j], acc[rg][1][j]);
                acc[rg][2][j] = fmaf(a.z, b[j], acc[rg][2][j]);
                acc[rg][3][j] = fmaf(a.w, b[j], acc[rg][3][j]);
            }
        }
    }
}

// dst[i] = sum_{g<G} src[g*stride + i]  in ascending g (deterministic second pass)
static __global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ src, int64_t stride, int G, int64_t n,
                    float* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += src[(int64_t)g * stride + i];
        dst[i] = s;
    }
}

}  // namespace ihg
