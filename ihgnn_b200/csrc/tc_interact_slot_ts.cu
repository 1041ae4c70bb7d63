// Backward of the order 2/3 FeatureInteractor, per-slot input gradients, A operand in tensor memory.
//
//   dz_b[e][k] = sum_n def[e][n] * W_b[n][k]          b = u*q, q*i, i*u [, u*q*i]
//   du = dz_uq*q + dz_iu*i + dz_uqi*q*i   (dq, di alike)  ->  slot_grad[e][slot][k]
//   (autograd of /root/reference/Models/CommonLayers.py:68-85 w.r.t. the gathered rows)
//
// Work unit = (tile of 128 hyperedges, slice j of 32 output columns).  The nb blocks of a slice are
// ONE MMA of N = 32*nb (the weight tile stacks the blocks' rows), so a unit is KC chunks of
// 4 K-steps x 3 MMAs (3xTF32) into a 32*nb-column accumulator, double-buffered in TMEM.
// The split def tile (A operand, K = dim) is written to TMEM once per tile with tcgen05.st and
// reused by all slices; only weight chunks are read from shared memory.
//
// Persistent, warp-specialised, one CTA per SM:
//   warps 0-3   A producers (warp = TMEM lane quadrant): cp.async their 32 def rows chunk by
//               chunk into a private double-buffered staging tile, split hi/lo, tcgen05.st;
//   warps 4-11  epilogue, two warps per quadrant (16 columns each): tcgen05.ld the dz slices,
//               product rule against the u/q/i slices staged by the gather warps (thread = row),
//               results overwrite the staged inputs, then coalesced slot_grad stores;
//   warp 12     MMA issuer (warp-uniform loop + elect.sync);
//   warp 13     weight loader (cp.async.bulk of pre-split, pre-swizzled chunks);
//   warps 14-17 gather: cp.async of the u/q/i 128-byte slices of a unit into a granule ring.
// TMEM map: [0,256) two accumulator buffers (stride 128); [256,512) def tile(s): per 32-wide
// K chunk 32 hi + 32 lo columns.
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

#ifdef IHG_TRACE
__device__ long long g_sl_trace[8][8192];
#define SL_PROBE(cond, rowi, idx)                                                     \
    do {                                                                              \
        if (blockIdx.x == 0 && (cond) && (idx) < 8192) g_sl_trace[rowi][idx] = clock64(); \
    } while (0)
#else
#define SL_PROBE(cond, rowi, idx) \
    do {                          \
    } while (0)
#endif

namespace {

constexpr int kSlAWarps = 4;
constexpr int kSlEpiWarp0 = 4;
constexpr int kSlEpiWarps = 8;
constexpr int kSlMmaWarp = 12;
constexpr int kSlLoadWarp = 13;
constexpr int kSlGatherWarp0 = 14;
constexpr int kSlGatherWarps = 8;
constexpr int kSlGatherThreads = kSlGatherWarps * 32;
constexpr int kSlThreads = (kSlGatherWarp0 + kSlGatherWarps) * 32;
constexpr int kSlGranuleBytes = 3 * kTileM * kChunkBytesPerRow;           // 48 KB: u, q, i slices
constexpr int kSlCopies = 3 * kTileM * 8 / kSlGatherThreads;              // cp.async per gather thread per granule
constexpr int kSlDefStageBytes = 32 * kChunkBytesPerRow;                  // 4 KB: one warp's 32 rows x 128 B
constexpr int kSlAccStride = 128;                                         // TMEM columns per accumulator buffer
constexpr int kSlMaxW = 4, kSlMaxG = 3;

__device__ __forceinline__ void sl_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void sl_cp16_zfill(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void sl_cp_arrive(uint32_t mbar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void sl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sl_cp_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void sl_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void sl_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// issue only; pair with sl_tmem_ld_wait()
__device__ __forceinline__ void sl_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void sl_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// granule slice layout: [table][row][128 B], 16-byte chunk index XORed with row % 8
__device__ __forceinline__ uint32_t sl_stage_off(int table, int row, int chunk) {
    return (uint32_t)((table * kTileM + row) * kChunkBytesPerRow + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void sl_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = rna_tf32(x);
    lo = __float_as_uint(x - __uint_as_float(hi));     // the tensor core reads its top 19 bits
}

// Weight tiles of the slot kernel: tile (j, nc) = rows r = b*32 + kk (block b, output column
// 32 j + kk), K = contraction index n in chunk nc:  w_hi[(nc*32 + n) * w_ld + b*dim + 32 j + kk];
// hi tile then lo tile, K-major SWIZZLE_128B.
__global__ void __launch_bounds__(256)
slot_prep_weights_kernel(const float* __restrict__ w_hi, int64_t w_ld, int nb, int dim, uint8_t* __restrict__ wprep) {
    const int KC = dim / kChunkK;
    const int rows = nb * 32;
    const int tile_bytes = rows * kChunkBytesPerRow;
    const int64_t total = (int64_t)KC * KC * rows * 8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = idx & 7;
        const int row = (int)((idx >> 3) % rows);
        const int tile = (int)(idx / (8 * rows));
        const int nc = tile % KC, j = tile / KC;
        const int b = row >> 5, kk = row & 31;
        uint32_t h[4], l[4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
            split_tf32(__ldg(w_hi + (int64_t)(nc * kChunkK + 4 * c + x) * w_ld + (int64_t)b * dim + 32 * j + kk), h[x], l[x]);
        uint8_t* t = wprep + (int64_t)tile * 2 * tile_bytes;
        const uint32_t off = sw128_offset(row, c);
        *reinterpret_cast<uint4*>(t + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(t + tile_bytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(kSlThreads, 1)
interact_bwd_slot_ts_kernel(const float* __restrict__ xp, int64_t xp_ld, const float* __restrict__ def,
                            int64_t def_ld, const uint8_t* __restrict__ wprep, int nb,
                            const int32_t* __restrict__ i3, int64_t E, float* __restrict__ slot_grad, int dim,
                            int w_stages, int n_gran, int n_abuf) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[8], bar_aempty[8];
    __shared__ __align__(8) uint64_t bar_wfull[kSlMaxW], bar_wempty[kSlMaxW];
    __shared__ __align__(8) uint64_t bar_gfull[kSlMaxG], bar_gempty[kSlMaxG];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = dim / kChunkK;                         // K chunks == column slices per tile
    const uint32_t w_tile_bytes = (uint32_t)nb * 32u * kChunkBytesPerRow;
    const uint32_t w_stage_bytes = 2u * w_tile_bytes;
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    // every CTA runs the same number of tiles (a cluster shares its weight stages in lock step);
    // tiles past the end are all-invalid rows: zero-filled loads, no stores
    const int64_t my_tiles = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t cl_rank = cluster_ctarank(), cl_size = cluster_nctarank();
    const uint16_t cl_mask = (uint16_t)((1u << cl_size) - 1u);
    // shared memory map: [W ring][granule ring][def staging: 4 warps x 2 x 4 KB]
    const uint32_t gran_base = smem_base + (uint32_t)w_stages * w_stage_bytes;
    const uint32_t def_base = gran_base + (uint32_t)n_gran * kSlGranuleBytes;

    if (tid == 0) {
        for (int s = 0; s < n_abuf * KC; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), kSlAWarps);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < w_stages; ++s) {
            mbar_init(smem_u32(&bar_wfull[s]), 1);
            mbar_init(smem_u32(&bar_wempty[s]), cl_size);                // one commit from every CTA of the cluster
        }
        for (int s = 0; s < n_gran; ++s) {
            mbar_init(smem_u32(&bar_gfull[s]), kSlGatherThreads);
            mbar_init(smem_u32(&bar_gempty[s]), kSlEpiWarps);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), kSlEpiWarps);
        }
        mbar_init_fence();
    }
    if (warp == kSlMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (cl_size > 1) cluster_sync();            // peers' barriers are initialised before anyone signals them
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_a0 = tmem_base + 2u * kSlAccStride;

    if (warp < kSlAWarps) {
        // ======================= A producers: def tile -> TMEM =======================
        const int quad = warp;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t stg = def_base + (uint32_t)warp * 2u * kSlDefStageBytes;
        const int c = lane & 7, rr = lane >> 3;           // copy jj: row rr + 4 jj of this warp's 32, chunk c
        const int64_t total = my_tiles * KC;              // def chunks of this CTA
        auto issue = [&](int64_t ai) {
            if (ai < total) {
                const int64_t k = ai / KC;
                const int nc = (int)(ai - k * KC);
                const int64_t e0 = (blockIdx.x + k * gridDim.x) * kTileM + quad * 32;
                const uint32_t buf = stg + (uint32_t)(ai & 1) * kSlDefStageBytes;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int r = rr + 4 * jj;
                    const bool ok = e0 + r < E;
                    sl_cp16_zfill(buf + epi_off(r, c), def + (ok ? e0 + r : 0) * def_ld + nc * kChunkK + 4 * c, ok);
                }
            }
            sl_cp_commit();
        };
        issue(0);
        int64_t ai = 0;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int ab = (int)(k % n_abuf);
            const uint32_t aph = (uint32_t)(k / n_abuf) & 1u;
            for (int nc = 0; nc < KC; ++nc, ++ai) {
                issue(ai + 1);
                SL_PROBE(tid == 0, 6, ai);
                sl_cp_wait1();                            // chunk ai landed (ai + 1 may be in flight)
                __syncwarp();
                SL_PROBE(tid == 0, 6, 2000 + ai);
                const uint32_t buf = stg + (uint32_t)(ai & 1) * kSlDefStageBytes;
                const int slot = ab * KC + nc;
                mbar_wait(smem_u32(&bar_aempty[slot]), aph ^ 1u);
                SL_PROBE(tid == 0, 6, 4000 + ai);
                fence_after_sync();
                const uint32_t ta = tmem_a0 + (uint32_t)(ab * 2 * dim + nc * 64) + lane_addr;
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const float4 a = lds4(buf + epi_off(lane, 2 * pass));
                    const float4 b = lds4(buf + epi_off(lane, 2 * pass + 1));
                    const float z[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int x = 0; x < 8; ++x) sl_split(z[x], hi[x], lo[x]);
                    sl_tmem_st8(ta + 8u * pass, hi);
                    sl_tmem_st8(ta + 32u + 8u * pass, lo);
                }
                sl_tmem_st_wait();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_afull[slot]));
                SL_PROBE(tid == 0, 7, ai);
            }
        }
        cp_async_wait_all();
    } else if (warp >= kSlGatherWarp0) {
        // ======================= gather: u/q/i slices of each unit =======================
        const int gt = tid - kSlGatherWarp0 * 32;
        constexpr int kRowStep = kSlGatherThreads / 8;
        const int chk = gt & 7, row0 = gt >> 3;          // copy jj: chunk chk of row row0 + kRowStep (jj / 3), table jj % 3
        const uint32_t off0 = sl_stage_off(0, row0, chk);
        int gb = 0;
        uint32_t gph = 0;
        // the node ids of a tile serve all of its KC units: loaded once per tile, the next tile's ids are
        // requested before this tile's copies are issued (in-kernel trace: with the ids re-read inside every
        // unit the gather warps needed ~3 900 cycles to issue one granule and bounded the unit period)
        auto load_ids = [&](int64_t k, int (&ids)[kSlCopies]) {
            const int64_t k0 = 3 * ((blockIdx.x + k * gridDim.x) * kTileM + row0);
#pragma unroll
            for (int jj = 0; jj < kSlCopies; ++jj) {
                const int64_t kk = k0 + 3 * kRowStep * (jj / 3) + (jj % 3);
                ids[jj] = (k < my_tiles && kk < 3 * E) ? __ldg(i3 + kk) : -1;
            }
        };
        int ids[kSlCopies], nxt[kSlCopies];
        load_ids(0, ids);
        for (int64_t k = 0; k < my_tiles; ++k) {
            load_ids(k + 1, nxt);
            for (int j = 0; j < KC; ++j) {
                SL_PROBE(gt == 0, 0, k * KC + j);
                mbar_wait(smem_u32(&bar_gempty[gb]), gph ^ 1u);
                SL_PROBE(gt == 0, 0, 2000 + k * KC + j);
                const uint32_t gbuf = gran_base + (uint32_t)gb * kSlGranuleBytes + off0;
                const float* col = xp + j * kChunkK + 4 * chk;
#pragma unroll
                for (int jj = 0; jj < kSlCopies; ++jj) {
                    const bool ok = ids[jj] >= 0;
                    sl_cp16_zfill(gbuf + (uint32_t)(((jj % 3) * kTileM + kRowStep * (jj / 3)) * kChunkBytesPerRow),
                                  col + (int64_t)(ok ? ids[jj] : 0) * xp_ld, ok);
                }
                sl_cp_arrive(smem_u32(&bar_gfull[gb]));
                SL_PROBE(gt == 0, 0, 4000 + k * KC + j);
                if (++gb == n_gran) gb = 0, gph ^= 1u;
            }
#pragma unroll
            for (int jj = 0; jj < kSlCopies; ++jj) ids[jj] = nxt[jj];
        }
        cp_async_wait_all();
    } else if (warp == kSlLoadWarp) {
        // ======================= weight loader =======================
        // each CTA of the cluster fetches 1 / cl_size of every chunk and multicasts it to all
        if (lane == 0) {
            const uint32_t piece = w_stage_bytes / cl_size;
            int sw = 0;
            uint32_t pw = 0;
            for (int64_t k = 0; k < my_tiles; ++k)
                for (int j = 0; j < KC; ++j)
                    for (int nc = 0; nc < KC; ++nc) {
                        mbar_wait(smem_u32(&bar_wempty[sw]), pw ^ 1u);
                        const uint32_t full = smem_u32(&bar_wfull[sw]);
                        sl_expect_tx(full, w_stage_bytes);
                        const uint32_t dst = smem_base + (uint32_t)sw * w_stage_bytes + cl_rank * piece;
                        const uint8_t* src = wprep + (int64_t)(j * KC + nc) * w_stage_bytes + cl_rank * piece;
                        if (cl_size > 1) bulk_g2s_multicast(dst, src, piece, full, cl_mask);
                        else sl_bulk_g2s(dst, src, piece, full);
                        if (++sw == w_stages) sw = 0, pw ^= 1u;
                    }
        }
    } else if (warp == kSlMmaWarp) {
        // ======================= MMA issuer =======================
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32(nb * 32);
        int sw = 0;
        uint32_t pw = 0, uc = 0;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int ab = (int)(k % n_abuf);
            const uint32_t aph = (uint32_t)(k / n_abuf) & 1u;
            for (int j = 0; j < KC; ++j, ++uc) {
                const uint32_t buf = uc & 1u;
                SL_PROBE(lane == 0, 1, uc);
                mbar_wait(smem_u32(&bar_tempty[buf]), ((uc >> 1) & 1u) ^ 1u);
                SL_PROBE(lane == 0, 1, 2000 + uc);
                const uint32_t tmem_d = tmu + buf * kSlAccStride;
                for (int nc = 0; nc < KC; ++nc) {
                    SL_PROBE(lane == 0, 2, uc * KC + nc);
                    mbar_wait(smem_u32(&bar_wfull[sw]), pw);
                    SL_PROBE(lane == 0, 3, uc * KC + nc);
                    if (j == 0) mbar_wait(smem_u32(&bar_afull[ab * KC + nc]), aph);
                    SL_PROBE(lane == 0, 2, 4000 + uc * KC + nc);
                    fence_after_sync();
                    const uint32_t a_hi = tmu + 2u * kSlAccStride + (uint32_t)(ab * 2 * dim + nc * 64), a_lo = a_hi + 32u;
                    const uint32_t w_hi = smem_base + (uint32_t)sw * w_stage_bytes;
                    const uint64_t dbh = make_kmajor_sw128_desc(w_hi);
                    const uint64_t dbl = make_kmajor_sw128_desc(w_hi + w_tile_bytes);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < kChunkK / 8; ++ks) {
                            const uint64_t bh = advance_desc_k(dbh, 8 * ks), bl = advance_desc_k(dbl, 8 * ks);
                            mma_tf32_ts(tmem_d, a_lo + 8u * ks, bh, idesc, (nc > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(tmem_d, a_hi + 8u * ks, bl, idesc, 1u);
                            mma_tf32_ts(tmem_d, a_hi + 8u * ks, bh, idesc, 1u);
                        }
                        if (cl_size > 1) mma_commit_multicast(smem_u32(&bar_wempty[sw]), cl_mask);
                        else mma_commit(smem_u32(&bar_wempty[sw]));
                        if (j == KC - 1) mma_commit(smem_u32(&bar_aempty[ab * KC + nc]));   // last reader of this def chunk
                        if (nc == KC - 1) mma_commit(smem_u32(&bar_tfull[buf]));
                    }
                    __syncwarp();
                    if (++sw == w_stages) sw = 0, pw ^= 1u;
                }
            }
        }
    } else {
        // ======================= epilogue: product rule =======================
        const int ew = warp - kSlEpiWarp0;
        const int quad = warp & 3, half = ew >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int cc = lane & 3, rs = lane >> 2;           // store phase: 4 lanes per 64-byte row piece
        int gb = 0;
        uint32_t gph = 0, uc = 0;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int64_t e0 = (blockIdx.x + k * gridDim.x) * kTileM + quad * 32;
            for (int j = 0; j < KC; ++j, ++uc) {
                const uint32_t buf = uc & 1u;
                const uint32_t gbuf = gran_base + (uint32_t)gb * kSlGranuleBytes;
                SL_PROBE(ew == 0 && lane == 0, 4, uc);
                mbar_wait(smem_u32(&bar_gfull[gb]), gph);
                SL_PROBE(ew == 0 && lane == 0, 4, 2000 + uc);
                mbar_wait(smem_u32(&bar_tfull[buf]), (uc >> 1) & 1u);
                SL_PROBE(ew == 0 && lane == 0, 4, 4000 + uc);
                fence_after_sync();
                const uint32_t taddr = tmem_base + buf * kSlAccStride + lane_addr;
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    const int pass = 2 * half + pp;           // columns 8 pass .. 8 pass + 7 of the slice
                    uint32_t z0[8], z1[8], z2[8], z3[8];
                    sl_tmem_ld8(taddr + (uint32_t)(8 * pass), z0);
                    sl_tmem_ld8(taddr + (uint32_t)(32 + 8 * pass), z1);
                    sl_tmem_ld8(taddr + (uint32_t)(64 + 8 * pass), z2);
                    if (nb == 4) sl_tmem_ld8(taddr + (uint32_t)(96 + 8 * pass), z3);
                    float u[8], q[8], v[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = lds4(gbuf + sl_stage_off(0, row, 2 * pass + h));
                        const float4 b = lds4(gbuf + sl_stage_off(1, row, 2 * pass + h));
                        const float4 c = lds4(gbuf + sl_stage_off(2, row, 2 * pass + h));
                        u[4 * h] = a.x, u[4 * h + 1] = a.y, u[4 * h + 2] = a.z, u[4 * h + 3] = a.w;
                        q[4 * h] = b.x, q[4 * h + 1] = b.y, q[4 * h + 2] = b.z, q[4 * h + 3] = b.w;
                        v[4 * h] = c.x, v[4 * h + 1] = c.y, v[4 * h + 2] = c.z, v[4 * h + 3] = c.w;
                    }
                    sl_tmem_ld_wait();
                    float du[8], dq[8], di[8];
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        const float a = __uint_as_float(z0[x]), b = __uint_as_float(z1[x]), c = __uint_as_float(z2[x]);
                        du[x] = fmaf(c, v[x], a * q[x]);          // uq: dz*q ; iu: dz*i
                        dq[x] = fmaf(b, v[x], a * u[x]);          // uq: dz*u ; qi: dz*i
                        di[x] = fmaf(c, u[x], b * q[x]);          // qi: dz*q ; iu: dz*u
                        if (nb == 4) {
                            const float d3 = __uint_as_float(z3[x]);
                            du[x] = fmaf(d3, q[x] * v[x], du[x]);
                            dq[x] = fmaf(d3, u[x] * v[x], dq[x]);
                            di[x] = fmaf(d3, u[x] * q[x], di[x]);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        sts4(gbuf + sl_stage_off(0, row, 2 * pass + h), make_float4(du[4 * h], du[4 * h + 1], du[4 * h + 2], du[4 * h + 3]));
                        sts4(gbuf + sl_stage_off(1, row, 2 * pass + h), make_float4(dq[4 * h], dq[4 * h + 1], dq[4 * h + 2], dq[4 * h + 3]));
                        sts4(gbuf + sl_stage_off(2, row, 2 * pass + h), make_float4(di[4 * h], di[4 * h + 1], di[4 * h + 2], di[4 * h + 3]));
                    }
                }
                // accumulator slice fully read: hand the buffer back to the MMA warp
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
                SL_PROBE(ew == 0 && lane == 0, 5, uc);
                // coalesced stores of this warp's 32 rows x 16 columns x 3 slots
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                    const int r = itr * 8 + rs;
                    if (e0 + r < E) {
                        float* out = slot_grad + (e0 + r) * 3 * (int64_t)dim + j * kChunkK + 16 * half + 4 * cc;
                        const int rowt = quad * 32 + r;
                        stg4(out, lds4(gbuf + sl_stage_off(0, rowt, 4 * half + cc)));
                        stg4(out + dim, lds4(gbuf + sl_stage_off(1, rowt, 4 * half + cc)));
                        stg4(out + 2 * dim, lds4(gbuf + sl_stage_off(2, rowt, 4 * half + cc)));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_gempty[gb]));
                SL_PROBE(ew == 0 && lane == 0, 5, 2000 + uc);
                if (++gb == n_gran) gb = 0, gph ^= 1u;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (cl_size > 1) cluster_sync();            // no peer may still multicast into / signal this CTA
    if (warp == kSlMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// slot_grad[e][slot][:] for all hyperedges; wprep = workspace for the pre-split weight tiles
// (nb * (dim/32) * 2 * dim * 128 bytes, 1024-byte aligned)
int launch_interact_bwd_slot_ts(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                                const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                                float* slot_grad, int dim, uint8_t* wprep, cudaStream_t st) {
    const int KC = dim / kChunkK;
    const int64_t total = (int64_t)KC * KC * nb * 32 * 8;
    slot_prep_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_hi, w_ld, nb, dim, wprep);
    IHG_LAUNCH_CHECK();
    const int w_stage_bytes = 2 * nb * 32 * kChunkBytesPerRow;
    const int n_gran = 2;
    int w_stages = (226 * 1024 - n_gran * kSlGranuleBytes - kSlAWarps * 2 * kSlDefStageBytes - 1024) / w_stage_bytes;
    if (w_stages > kSlMaxW) w_stages = kSlMaxW;
    int n_abuf = 256 / (2 * dim);
    if (n_abuf > 2) n_abuf = 2;
    if (n_abuf < 1) n_abuf = 1;
    const int smem = w_stages * w_stage_bytes + n_gran * kSlGranuleBytes + kSlAWarps * 2 * kSlDefStageBytes + 1024;
    IHG_CUDA(cudaFuncSetAttribute(interact_bwd_slot_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    static int cluster = -1, max_ctas = 0;
    if (cluster < 0) {
        IHG_CUDA(cudaFuncSetAttribute(interact_bwd_slot_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        cluster = pick_cluster(interact_bwd_slot_ts_kernel, kSlThreads, 226 * 1024, 4, &max_ctas);
    }
    int64_t ctas = n_tiles < max_ctas ? n_tiles : max_ctas;
    ctas = (ctas + cluster - 1) / cluster * cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kSlThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    IHG_CUDA(cudaLaunchKernelEx(&cfg, interact_bwd_slot_ts_kernel, xp, xp_ld, def, def_ld, (const uint8_t*)wprep, nb, i3, E,
                                slot_grad, dim, w_stages, n_gran, n_abuf));
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg

#ifdef IHG_TRACE
extern "C" int ihg_debug_read_trace_slot(long long* dst, int n) {
    if (n > 8 * 8192) n = 8 * 8192;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(dst, ihg::g_sl_trace, (size_t)n * sizeof(long long));
}
#endif
