// Weight gradient of the node-level typed Linear on the 5th-generation tensor cores (tcgen05.mma
// kind::tf32, 3xTF32 split precision, fp32 accumulators in TMEM).  The forward / input-gradient
// product lives in tc_linear_ts.cu (A operand in tensor memory, weights resident in shared memory).
//
// Roofline: HBM (4*(n_in+n_out) bytes per row vs 6*n_in*n_out tf32 flops per row).
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

// =========================================================================================
// Weight gradient of the typed node Linear on tensor cores:
//   dw[t][n][k] = sum_{r in type t} dy[r][n] * x[r][k]
// A = dy tile, B = x tile, both [32 rows x 32 features] sub-tiles in the MN-major
// SWIZZLE_128B_BASE32B layout (the row index is the MMA K dimension); M is padded to 128 with
// zero sub-tiles.  One accumulator [128 x n_in] per node type stays in TMEM across all tiles of
// the CTA; partials go to the workspace and are summed in CTA order (deterministic).
// 8 producer warps + 1 MMA warp per CTA, two CTAs per SM.  (With 4 producer warps the kernel was
// issue-latency bound: ncu showed the few resident warps mostly "selected" / waiting on fixed-latency
// ALU dependencies of the tf32 split, tensor pipe 23 %, DRAM 25 %.)
// =========================================================================================
constexpr int kNwTe = 32;
constexpr int kNwProdWarps = 8;                   // one tile row x one 16-byte chunk column per producer thread
constexpr int kNwThreads = (kNwProdWarps + 1) * 32;
constexpr int kNwMaxBlk = 4;

__global__ void __launch_bounds__(kNwThreads)
node_wgrad_tc_kernel(const float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ x,
                     int64_t x_ld, int64_t b0, int64_t b1, int64_t n_rows, int n_types, int n_out,
                     int n_in, float* __restrict__ ws_dw) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KA = n_out / kChunkK, KB = n_in / kChunkK;
    constexpr uint32_t sub_bytes = kNwTe * kChunkBytesPerRow;             // 4 KB
    const uint32_t a_bytes = 8 * sub_bytes;                               // 4 hi + 4 lo sub-tiles (M = 128)
    const int NB = KB + 1;                                                // + the constant "ones" block (bias gradient)
    const uint32_t stage_bytes = a_bytes + 2 * (uint32_t)NB * sub_bytes;
    const bool is_mma_warp = warp == kNwProdWarps;
    const int acc_cols = n_in + 16;                                       // column n_in accumulates sum_r dy[r][n]
    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(n_types * acc_cols));

    // tiles never straddle a node-type boundary
    const int64_t lo[3] = {0, b0, b1}, hi[3] = {b0, b1, n_rows};
    int64_t tiles_t[3], tile_base[4];
    tile_base[0] = 0;
    for (int t = 0; t < 3; ++t) {
        tiles_t[t] = (hi[t] - lo[t] + kNwTe - 1) / kNwTe;
        tile_base[t + 1] = tile_base[t] + tiles_t[t];
    }
    const int64_t n_tiles = tile_base[3];

    // once: zero the whole operand area (M-padding sub-tiles of A stay zero), then set column 0 of
    // the extra B block to 1.0 so that accumulator column n_in collects the bias gradient
    for (uint32_t off = tid * 16; off < 2 * stage_bytes; off += kNwThreads * 16) sts4(smem_base + off, f4_zero());
    __syncthreads();
    if (tid < 2 * kNwTe) {
        const int s = tid / kNwTe, r = tid % kNwTe;
        sts4(smem_base + (uint32_t)s * stage_bytes + a_bytes + (uint32_t)KB * sub_bytes + sw128b32_offset(r, 0),
             make_float4(1.f, 0.f, 0.f, 0.f));
    }
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_full[s]), kNwProdWarps);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_done), 1);
        mbar_init_fence();
    }
    if (is_mma_warp) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;

    if (!is_mma_warp) {
        const int c = tid & 7, r0 = tid >> 3;            // row r0 of the tile; chunk c of every block
        uint32_t it = 0;
        uint32_t started = 0;                            // bit t set once type t's accumulator is live
        // Software pipeline: the loads of tile i+1 are issued BEFORE tile i is split and stored, so
        // every thread keeps its 128-bit loads in flight across the wait / store phase.
        float4 av[kNwMaxBlk], bv[kNwMaxBlk];
        auto load_tile = [&](int64_t tile, float4 (&a)[kNwMaxBlk], float4 (&b)[kNwMaxBlk]) {
            int t = 0;
            while (t < 2 && tile >= tile_base[t + 1]) ++t;
            const int64_t r = lo[t] + (tile - tile_base[t]) * kNwTe + r0;
            const bool ok = tile < n_tiles && r < hi[t];
#pragma unroll
            for (int blk = 0; blk < kNwMaxBlk; ++blk) {
                a[blk] = (ok && blk < KA) ? ldg4(dy + r * dy_ld + blk * kChunkK + 4 * c) : f4_zero();
                b[blk] = (ok && blk < KB) ? ldg4(x + r * x_ld + blk * kChunkK + 4 * c) : f4_zero();
            }
        };
        if ((int64_t)blockIdx.x < n_tiles) load_tile(blockIdx.x, av, bv);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            int t = 0;
            while (t < 2 && tile >= tile_base[t + 1]) ++t;
            started |= 1u << (n_types > 1 ? t : 0);
            float4 an[kNwMaxBlk], bn[kNwMaxBlk];
            load_tile(tile + gridDim.x, an, bn);
            const int s = it & 1;
            mbar_wait(smem_u32(&bar_empty[s]), ((it >> 1) & 1u) ^ 1u);
            const uint32_t ah = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t bh = ah + a_bytes;
#pragma unroll
            for (int blk = 0; blk < kNwMaxBlk; ++blk) {
                if (blk < KA)
                    store_split_chunk_mn(ah + (uint32_t)blk * sub_bytes, ah + (uint32_t)(4 + blk) * sub_bytes, r0, c, av[blk]);
                if (blk < KB)
                    store_split_chunk_mn(bh + (uint32_t)blk * sub_bytes, bh + (uint32_t)(NB + blk) * sub_bytes, r0, c, bv[blk]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
#pragma unroll
            for (int blk = 0; blk < kNwMaxBlk; ++blk) {
                av[blk] = an[blk];
                bv[blk] = bn[blk];
            }
        }
        // final epilogue: warp == TMEM lane quadrant; rows n < n_out of every type's accumulator
        const int n = warp * 32 + lane;
        float* out = ws_dw + (int64_t)blockIdx.x * n_types * n_out * acc_cols;
        const bool any = blockIdx.x < n_tiles;
        if (any) {
            mbar_wait(smem_u32(&bar_done), 0);
            fence_after_sync();
        }
        for (int t = 0; t < n_types && warp < 4; ++t)    // warps 0-3 own the four TMEM lane quadrants
            for (int c0 = 0; c0 < acc_cols; c0 += 16) {
                float acc[16];
                if (any && ((started >> t) & 1u)) {      // block-uniform: untouched accumulators are undefined
                    tmem_ld16(tmem_base + (uint32_t)(t * acc_cols + c0) + ((uint32_t)(warp * 32) << 16), acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
                }
                if (n < n_out) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stg4(out + ((int64_t)t * n_out + n) * acc_cols + c0 + 4 * j,
                             make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
                }
            }
    } else {
        // warp-uniform loop, one elected lane issues (tc_common.cuh: MMA issue discipline)
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32_mn(acc_cols);
        uint32_t it = 0;
        uint32_t started = 0;                            // bit t set once type t's accumulator is live
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            int t = 0;
            while (t < 2 && tile >= tile_base[t + 1]) ++t;
            const int tt = n_types > 1 ? t : 0;
            const int s = it & 1;
            mbar_wait(smem_u32(&bar_full[s]), (it >> 1) & 1u);
            fence_after_sync();
            const uint32_t ah = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t bh = ah + a_bytes;
            const uint32_t tmem_d = tmu + (uint32_t)(tt * acc_cols);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kNwTe / 8; ++ks) {
                    const uint32_t koff = (uint32_t)ks * 1024u;
                    mma_3xtf32(tmem_d, make_mnmajor_sw128_desc(ah + koff, sub_bytes),
                               make_mnmajor_sw128_desc(ah + 4 * sub_bytes + koff, sub_bytes),
                               make_mnmajor_sw128_desc(bh + koff, sub_bytes),
                               make_mnmajor_sw128_desc(bh + (uint32_t)NB * sub_bytes + koff, sub_bytes), idesc,
                               ((started >> tt) & 1u) ? 1u : (ks > 0 ? 1u : 0u));
                }
                mma_commit(smem_u32(&bar_empty[s]));
            }
            __syncwarp();
            started |= 1u << tt;
        }
        if (elect_one()) mma_commit(smem_u32(&bar_done));
        __syncwarp();
    }
    fence_before_sync();
    __syncthreads();
    if (is_mma_warp) tmem_dealloc(tmem_base, tmem_cols);
}

bool node_wgrad_tc_eligible(int n_types, int n_out, int n_in) {
    return n_out % 32 == 0 && n_in % 32 == 0 && n_out <= 128 && n_in <= 128 && n_types * (n_in + 16) <= 512;
}
constexpr int kNwCtas = 2 * kNumSMs;

int64_t node_wgrad_tc_workspace_bytes(int n_types, int n_out, int n_in) {
    return (int64_t)kNwCtas * n_types * n_out * (n_in + 16) * 4 + 1024;
}

// dw[t][n][k] = sum_g ws[g][t][n][k],  db[t][n] = sum_g ws[g][t][n][n_in]   (ascending g)
// 256 threads = 32 outputs x 8 slices: slice s sums the contiguous range of partials
// [s*G/8, (s+1)*G/8) in ascending g, the slice sums are combined in ascending s (fixed order,
// deterministic; 8x the loads in flight of a one-thread-per-output loop).
__global__ void __launch_bounds__(256)
wgrad_partials_sum_kernel(const float* __restrict__ ws, int G, int n_types, int n_out, int n_in,
                          float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float part[8][32];
    const int acc_cols = n_in + 16;
    const int64_t rows = (int64_t)n_types * n_out;
    const int64_t total = rows * (n_in + 1);
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    const int64_t r = i / (n_in + 1);
    const int k = (int)(i % (n_in + 1));
    float s = 0.f;
    if (i < total) {
        const int g0 = (int)((int64_t)slice * G / 8), g1 = (int)((int64_t)(slice + 1) * G / 8);
        for (int g = g0; g < g1; ++g) s += ws[((int64_t)g * rows + r) * acc_cols + k];
    }
    part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && i < total) {
        float t = part[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += part[q][lane];
        if (k < n_in) dw[r * n_in + k] = t;
        else if (db) db[r] = t;
    }
}

int launch_node_wgrad_tc(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, int64_t n_rows,
                         int64_t b0, int64_t b1, int n_types, int n_out, int n_in, float* dw, float* db,
                         void* workspace, cudaStream_t st) {
    float* ws_dw = static_cast<float*>(workspace);
    const int KB = n_in / kChunkK;
    const int smem = 2 * (8 + 2 * (KB + 1)) * kNwTe * kChunkBytesPerRow + 1024;
    static int attr_smem = 0;
    if (attr_smem < smem) {
        IHG_CUDA(cudaFuncSetAttribute(node_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    int64_t n_tiles = 0;
    const int64_t lo[3] = {0, b0, b1}, hi[3] = {b0, b1, n_rows};
    for (int t = 0; t < 3; ++t) n_tiles += (hi[t] - lo[t] + kNwTe - 1) / kNwTe;
    const int grid = (int)(n_tiles < kNwCtas ? n_tiles : kNwCtas);
    node_wgrad_tc_kernel<<<grid, kNwThreads, smem, st>>>(dy, dy_ld, x, x_ld, b0, b1, n_rows, n_types, n_out, n_in, ws_dw);
    IHG_LAUNCH_CHECK();
    const int64_t total = (int64_t)n_types * n_out * (n_in + 1);
    wgrad_partials_sum_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(ws_dw, grid, n_types, n_out, n_in, dw, db);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg
