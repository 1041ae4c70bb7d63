// Node-level typed Linear on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32
// split precision, fp32 accumulators in TMEM).  Same contract as the SIMT kernel in
// node_linear.cu -- y[r,:] = x[r,:] . W[t]^T (+bias[t]) (+addend[r,:]) -- selected when
// n_in % 32 == 0 and n_out % 16 == 0.
//
// One CTA owns one tile of 128 rows (tiles never straddle a node-type boundary): row slices of x
// are split into tf32 hi/lo and written into the UMMA K-major SWIZZLE_128B layout together with
// the weight rows; one elected thread issues the MMAs; completion is tracked with an mbarrier via
// tcgen05.commit; the epilogue reads the accumulator with tcgen05.ld.  Several CTAs are resident
// per SM, so staging, MMA and epilogue of different tiles overlap.
//
// Roofline: HBM (4*(n_in+n_out) bytes per row vs 6*n_in*n_out tf32 flops per row).
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

struct TcTypeTiles {
    int64_t b0, b1, n_rows;
    __host__ __device__ int64_t lo(int t) const { return t == 0 ? 0 : (t == 1 ? b0 : b1); }
    __host__ __device__ int64_t hi(int t) const { return t == 0 ? b0 : (t == 1 ? b1 : n_rows); }
    __host__ __device__ int64_t tiles(int t) const { return (hi(t) - lo(t) + kTileM - 1) / kTileM; }
};

// 256 threads per CTA, one 128-row tile per CTA, K processed in phases of at most 64 features
// (2 chunks): per phase every thread issues all of its 128-bit row loads in one burst with the
// coalesced (row, chunk) mapping (8 lanes = one 128-byte row slice), splits to tf32 hi/lo and
// stages A and the weight rows; one thread issues the MMAs; the epilogue transposes the
// accumulator through shared memory (reusing the A region) so that bias / addend reads and the
// stores of y are full-line coalesced.  96 KB of shared memory per CTA => two CTAs per SM
// overlap each other's load latency.
constexpr int kNlThreads = 256;
constexpr int kNlPhaseChunks = 2;

__global__ void __launch_bounds__(kNlThreads)
node_linear_tc_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ w,
                      int n_types, int n_out, int n_in, int transpose_w,
                      const float* __restrict__ bias, const float* __restrict__ addend,
                      int64_t addend_ld, TcTypeTiles tt, float* __restrict__ y, int64_t y_ld) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // per chunk: A hi, A lo (16 KB each); then per chunk: B hi, B lo (n_out x 128 B each)
    const uint32_t b_tile = (uint32_t)n_out * kChunkBytesPerRow;
    const uint32_t a_base = base, b_base = base + kNlPhaseChunks * 2 * 16384;
    __shared__ __align__(8) uint64_t mbar_storage;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t mbar = smem_u32(&mbar_storage);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int64_t bid = blockIdx.x;
    int type = 0;
    while (type < 2 && bid >= tt.tiles(type)) { bid -= tt.tiles(type); ++type; }
    const int64_t row0 = tt.lo(type) + bid * kTileM;
    const int rows = (int)min((int64_t)kTileM, tt.hi(type) - row0);
    const int wt = n_types > 1 ? type : 0;
    const float* W = w + (int64_t)wt * n_out * n_in;

    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)n_out);
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init_fence();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = make_idesc_tf32(n_out);

    const int c = tid & 7, r0 = tid >> 3;               // rows r0 + 32 j, chunk c
    uint32_t phase = 0;
    const int n_chunks = n_in / kChunkK;
    for (int kc0 = 0; kc0 < n_chunks; kc0 += kNlPhaseChunks) {
        const int nch = min(kNlPhaseChunks, n_chunks - kc0);
        // ---- one burst of loads: A rows and weight rows of this phase
        float4 av[kNlPhaseChunks][4], bv[kNlPhaseChunks][4];
#pragma unroll
        for (int ch = 0; ch < kNlPhaseChunks; ++ch) {
            const int k0 = (kc0 + ch) * kChunkK + 4 * c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + 32 * j;
                av[ch][j] = (ch < nch && r < rows) ? ldg4(x + (row0 + r) * x_ld + k0) : f4_zero();
                float4 v = f4_zero();
                if (ch < nch && r < n_out) {
                    if (!transpose_w) {
                        v = ldg4(W + (int64_t)r * n_in + k0);
                    } else {
                        const float* p = W + (int64_t)k0 * n_out + r;
                        v = make_float4(__ldg(p), __ldg(p + n_out), __ldg(p + 2 * n_out), __ldg(p + 3 * n_out));
                    }
                }
                bv[ch][j] = v;
            }
        }
#pragma unroll
        for (int ch = 0; ch < kNlPhaseChunks; ++ch)
            if (ch < nch) {
                const uint32_t a_hi = a_base + (uint32_t)ch * 32768, a_lo = a_hi + 16384;
                const uint32_t b_hi = b_base + (uint32_t)ch * 2 * b_tile, b_lo = b_hi + b_tile;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + 32 * j;
                    store_split_chunk(a_hi, a_lo, r, c, av[ch][j]);
                    if (r < n_out) store_split_chunk(b_hi, b_lo, r, c, bv[ch][j]);
                }
            }
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            // warp-uniform issue, one elected lane (tc_common.cuh: MMA issue discipline)
            fence_after_sync();
            const uint32_t tmu = warp_uniform(tmem_d);
            for (int ch = 0; ch < nch; ++ch) {
                const uint32_t a_hi = a_base + (uint32_t)ch * 32768;
                const uint32_t b_hi = b_base + (uint32_t)ch * 2 * b_tile;
                const uint64_t dah = make_kmajor_sw128_desc(a_hi), dal = make_kmajor_sw128_desc(a_hi + 16384);
                const uint64_t dbh = make_kmajor_sw128_desc(b_hi), dbl = make_kmajor_sw128_desc(b_hi + b_tile);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks)
                        mma_3xtf32(tmu, advance_desc_k(dah, 8 * ks), advance_desc_k(dal, 8 * ks),
                                   advance_desc_k(dbh, 8 * ks), advance_desc_k(dbl, 8 * ks), idesc,
                                   (kc0 + ch > 0 || ks > 0) ? 1u : 0u);
                    if (ch == nch - 1) mma_commit(mbar);
                }
                __syncwarp();
            }
        }
        // the operand tiles are reused by the next phase / the epilogue: wait for the MMAs
        mbar_wait(mbar, phase);
        phase ^= 1u;
    }
    fence_after_sync();
    // ---- epilogue: warps (q, half) read 32-column slabs of quadrant q, transpose through a
    // per-warp staging tile (the A region is free now) and write coalesced rows
    const int q4 = warp & 3, half = warp >> 2;
    const uint32_t stg = a_base + (uint32_t)warp * kEpiStageBytes;
    const int rs = lane >> 3;
    for (int c0 = 32 * half; c0 < n_out; c0 += 64) {
        const int ncol = min(32, n_out - c0);
        __syncwarp();
        for (int hc = 0; hc < ncol; hc += 16) {
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c0 + hc), v);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                sts4(stg + epi_off(lane, hc / 4 + j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        __syncwarp();
        if (4 * c < ncol) {
            float4 bb = f4_zero();
            if (bias) bb = ldg4(bias + (int64_t)wt * n_out + c0 + 4 * c);
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
                const int r = q4 * 32 + itr * 4 + rs;
                if (r < rows) {
                    float4 o = lds4(stg + epi_off(itr * 4 + rs, c));
                    f4_add(o, bb);
                    if (addend) f4_add(o, ldg4(addend + (row0 + r) * addend_ld + c0 + 4 * c));
                    stg4(y + (row0 + r) * y_ld + c0 + 4 * c, o);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

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
    static const bool disabled = getenv("IHG_DISABLE_TC") != nullptr;
    if (disabled) return false;
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

bool node_linear_tc_eligible(int n_out, int n_in, int64_t x_ld, int64_t y_ld, const float* addend,
                             int64_t addend_ld) {
    static const bool disabled = getenv("IHG_DISABLE_TC") != nullptr;
    if (disabled) return false;
    if (n_in % 32 != 0 || n_out % 16 != 0 || n_out < 16 || n_out > 128 || n_in > 128) return false;
    if (x_ld % 4 != 0 || y_ld % 4 != 0) return false;
    if (addend && addend_ld % 4 != 0) return false;
    return true;
}

int launch_node_linear_tc(const float* x, int64_t x_ld, const float* w, int n_types, int n_out,
                          int n_in, int transpose_w, const float* bias, const float* addend,
                          int64_t addend_ld, int64_t n_rows, int64_t bound0, int64_t bound1,
                          float* y, int64_t y_ld, cudaStream_t st) {
    TcTypeTiles tt{bound0, bound1, n_rows};
    const int64_t blocks = tt.tiles(0) + tt.tiles(1) + tt.tiles(2);
    const int smem = kNlPhaseChunks * (2 * 16384 + 2 * n_out * kChunkBytesPerRow) + 1024;
    static int attr_smem = 0;
    if (attr_smem < smem) {
        IHG_CUDA(cudaFuncSetAttribute(node_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    node_linear_tc_kernel<<<(unsigned)blocks, kNlThreads, smem, st>>>(x, x_ld, w, n_types, n_out, n_in, transpose_w,
                                                               bias, addend, addend_ld, tt, y, y_ld);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg
