// Order 2/3 FeatureInteractor.forward, A operand in tensor memory ("TS" form of tcgen05.mma).
//
//   ef[e,:] = aggregation.weight . cat(u, q, i, u*q, q*i, i*u [, u*q*i]) + bias
//   (/root/reference/Models/CommonLayers.py:68-85; u,q,i = projected rows of the hyperedge's nodes)
//
// The split (3xTF32) A operand never touches shared memory: the producers build it in registers
// and write it to TMEM with tcgen05.st (thread = one hyperedge row = one TMEM lane: no swizzle,
// no proxy fence), the MMA reads A from TMEM and only the weight chunk from shared memory.
//
// Persistent, warp-specialised, one CTA per SM, 128 hyperedges per tile.  A "granule" is
// (tile, 32-column slice kc); a "chunk" is (granule, operand block b) = 12 MMAs (4 K steps x 3).
//   warps 22-25 gather: cp.async the 128-byte slices of the u/q/i rows of a granule into a ring
//               of staging buffers (8 lanes per row: coalesced, no registers), completion on an
//               mbarrier (cp.async.mbarrier.arrive.noinc);
//   warps 0-15  producers, four groups of four warps; group g owns chunks c = g (mod 4), so four
//               chunks are in flight and the per-chunk barrier traffic is paid once per 32
//               columns: thread = row 32*(warp%4)+lane reads its slices, forms the block's
//               products, splits hi/lo and tcgen05.st's them into the TMEM A ring;
//   warp 20     MMA issuer (warp-uniform loop + elect.sync, see tc_common.cuh);
//   warp 21     weight loader: cp.async.bulk of the pre-split, pre-swizzled weight chunks;
//   warps 16-19 epilogue: tcgen05.ld the accumulator, add bias, staged coalesced store of ef.
// TMEM map (512 columns): [0, 2*dim) two accumulator buffers; then A stages of 64 columns
// (32 hi + 32 lo).
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

#ifdef IHG_TRACE
// dev-only clock64 probes of block 0 (python -m ihgnn_b200.build --trace; profiles/trace_fwd_kernel.py)
__device__ long long g_ts_trace[8][8192];
__device__ long long g_ts_cta[160][2];             // per-CTA start / end (globaltimer ns)
__device__ __forceinline__ long long ts_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TS_PROBE(cond, rowi, idx)                                                     \
    do {                                                                              \
        if (blockIdx.x == 0 && (cond) && (idx) < 8192) g_ts_trace[rowi][idx] = clock64(); \
    } while (0)
#else
#define TS_PROBE(cond, rowi, idx) \
    do {                          \
    } while (0)
#endif

namespace {

constexpr int kTsProducerWarps = 12;
constexpr int kTsGroups = kTsProducerWarps / 4;
constexpr int kTsEpiWarp0 = kTsProducerWarps;
constexpr int kTsMmaWarp = kTsProducerWarps + 4;
constexpr int kTsLoadWarp = kTsProducerWarps + 5;
constexpr int kTsGatherWarp0 = kTsProducerWarps + 6;
constexpr int kTsGatherWarps = 8;
constexpr int kTsGatherThreads = kTsGatherWarps * 32;
constexpr int kTsThreads = (kTsProducerWarps + 6 + kTsGatherWarps) * 32;
constexpr int kTsGranuleBytes = 3 * kTileM * kChunkBytesPerRow;           // 48 KB: u, q, i slices
constexpr int kTsCopies = 3 * kTileM * 8 / kTsGatherThreads;              // 16-byte cp.async per gather thread per granule
constexpr int kTsMaxAStages = 6;
constexpr int kTsMaxWStages = 6;
constexpr int kTsMaxGranules = 3;

__device__ __forceinline__ void mbar_expect_tx_(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// 16-byte async copy, zero-filled when !valid (src-size 0)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// arrive on `mbar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t mbar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}
// this warp's TMEM lane quadrant x 8 consecutive 32-bit columns <- 8 registers per thread
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// staging slice: [table][row][128 B], the 16-byte chunk index XORed with row % 8 (conflict-free
// both for the 8-lanes-per-row writes and the thread-per-row reads)
__device__ __forceinline__ uint32_t stage_off(int table, int row, int chunk) {
    return (uint32_t)((table * kTileM + row) * kChunkBytesPerRow + ((chunk ^ (row & 7)) << 4));
}
// hi = tf32(x) rounded to nearest; lo = x - hi is passed as is: the tensor core reads its top
// 19 bits, |lo - tf32(lo)| <= 2^-10 |lo| <= 2^-21 |x|
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = rna_tf32(x);
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// One chunk of operand block B for this thread's row: 32 columns in four passes of 8.
//   B: 0,1,2 = u, q, i;  3,4,5 = u*q, q*i, i*u;  6 = u*q*i
template <int B>
__device__ __forceinline__ void produce_chunk(uint32_t gbuf, int row, uint32_t ta) {
    constexpr bool need_u = B == 0 || B == 3 || B == 5 || B == 6;
    constexpr bool need_q = B == 1 || B == 3 || B == 4 || B == 6;
    constexpr bool need_v = B == 2 || B == 4 || B == 5 || B == 6;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        float z[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 a = f4_zero(), b = f4_zero(), c = f4_zero();
            if (need_u) a = lds4(gbuf + stage_off(0, row, 2 * pass + h));
            if (need_q) b = lds4(gbuf + stage_off(1, row, 2 * pass + h));
            if (need_v) c = lds4(gbuf + stage_off(2, row, 2 * pass + h));
            float4 r;
            if (B == 0) r = a;
            else if (B == 1) r = b;
            else if (B == 2) r = c;
            else if (B == 3) r = f4_mul(a, b);
            else if (B == 4) r = f4_mul(b, c);
            else if (B == 5) r = f4_mul(c, a);
            else r = f4_mul(f4_mul(a, b), c);
            z[4 * h] = r.x, z[4 * h + 1] = r.y, z[4 * h + 2] = r.z, z[4 * h + 3] = r.w;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) split_fast(z[k], hi[k], lo[k]);
        tmem_st8(ta + 8u * pass, hi);
        tmem_st8(ta + 32u + 8u * pass, lo);
    }
}

__global__ void __launch_bounds__(kTsThreads, 1)
feature_interact_fwd_ts_kernel(const float* __restrict__ xp, int64_t xp_ld, const float* __restrict__ bias,
                               const uint8_t* __restrict__ wprep, int nb, const int32_t* __restrict__ i3,
                               int64_t E, float* __restrict__ ef, int64_t ef_ld, int dim, int a_stages,
                               int w_stages, int n_gran) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kTsMaxAStages], bar_aempty[kTsMaxAStages];
    __shared__ __align__(8) uint64_t bar_wfull[kTsMaxWStages], bar_wempty[kTsMaxWStages];
    __shared__ __align__(8) uint64_t bar_gfull[kTsMaxGranules], bar_gempty[kTsMaxGranules];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = dim / kChunkK;
    const int chunks_per_tile = KC * nb;
    const uint32_t w_stage_bytes = 2u * (uint32_t)dim * kChunkBytesPerRow;     // hi + lo weight chunk
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    // every CTA runs the same number of tiles (the CTAs of a cluster share weight stages in lock
    // step); tiles past the end are all-invalid rows: zero-filled gathers, no stores
    const int64_t my_tiles = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t cl_rank = cluster_ctarank(), cl_size = cluster_nctarank();
    const uint16_t cl_mask = (uint16_t)((1u << cl_size) - 1u);
    // shared memory map: [W ring][granule ring][4 epilogue staging tiles]
    const uint32_t gran_base = smem_base + (uint32_t)w_stages * w_stage_bytes;
    const uint32_t epi_base = gran_base + (uint32_t)n_gran * kTsGranuleBytes;

    if (tid == 0) {
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), 4);                       // the four warps of one group
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < w_stages; ++s) {
            mbar_init(smem_u32(&bar_wfull[s]), 1);
            mbar_init(smem_u32(&bar_wempty[s]), cl_size);                // one commit from every CTA of the cluster
        }
        for (int s = 0; s < n_gran; ++s) {
            mbar_init(smem_u32(&bar_gfull[s]), kTsGatherThreads);
            mbar_init(smem_u32(&bar_gempty[s]), 4 * nb);                 // one arrival per warp per chunk
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), 4);
        }
        mbar_init_fence();
    }
    if (warp == kTsMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (cl_size > 1) cluster_sync();            // peers' barriers are initialised before anyone signals them
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_a0 = tmem_base + 2u * (uint32_t)dim;
#ifdef IHG_TRACE
    if (tid == 0 && blockIdx.x < 160) g_ts_cta[blockIdx.x][0] = ts_gtime();
#endif

    if (warp < kTsProducerWarps) {
        // ======================= producers =======================
        const int group = warp >> 2, quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int64_t total_chunks = my_tiles * chunks_per_tile;
        // chunk c = granule * nb + b; this group takes c = group, group + 4, ...
        int b = group, gb = 0, s = group;            // block, granule ring slot, A ring slot
        uint32_t gph = 0, ph = 0;
        // (kTsGroups <= nb and kTsGroups <= a_stages: no wrap in the initial position)
        for (int64_t c = group; c < total_chunks; c += kTsGroups) {
            TS_PROBE(tid == 0, 1, c >> 2);
            mbar_wait(smem_u32(&bar_gfull[gb]), gph);
            mbar_wait(smem_u32(&bar_aempty[s]), ph ^ 1u);
            TS_PROBE(tid == 0, 2, c >> 2);
            fence_after_sync();
            const uint32_t gbuf = gran_base + (uint32_t)gb * kTsGranuleBytes;
            const uint32_t ta = tmem_a0 + (uint32_t)s * 64u + lane_addr;
            switch (b) {
                case 0: produce_chunk<0>(gbuf, row, ta); break;
                case 1: produce_chunk<1>(gbuf, row, ta); break;
                case 2: produce_chunk<2>(gbuf, row, ta); break;
                case 3: produce_chunk<3>(gbuf, row, ta); break;
                case 4: produce_chunk<4>(gbuf, row, ta); break;
                case 5: produce_chunk<5>(gbuf, row, ta); break;
                default: produce_chunk<6>(gbuf, row, ta); break;
            }
            TS_PROBE(tid == 0, 3, c >> 2);
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&bar_afull[s]));
                mbar_arrive(smem_u32(&bar_gempty[gb]));
            }
            TS_PROBE(tid == 0, 6, c >> 2);
            b += kTsGroups;
            while (b >= nb) {
                b -= nb;
                if (++gb == n_gran) gb = 0, gph ^= 1u;
            }
            s += kTsGroups;
            if (s >= a_stages) s -= a_stages, ph ^= 1u;
        }
    } else if (warp >= kTsGatherWarp0) {
        // ======================= gather =======================
        const int gt = tid - kTsGatherWarp0 * 32;
        constexpr int kRowStep = kTsGatherThreads / 8;      // rows covered by one pass of the gather threads
        const int chk = gt & 7, row0 = gt >> 3;       // copy j: 16-byte chunk chk of row row0 + kRowStep (j / 3), table j % 3
        const uint32_t off0 = stage_off(0, row0, chk);    // (row0 + kRowStep m) % 8 == row0 % 8: same swizzle for all j
        int gb = 0;
        uint32_t gph = 0;
        // the node ids of a tile serve all of its KC column slices: loaded once per tile, the next tile's ids
        // are requested before this tile's copies are issued (the id latency -- two dependent global loads per
        // slice before -- dominated the issue time of a granule)
        auto load_ids = [&](int64_t t, int (&ids)[kTsCopies]) {
            const int64_t k0 = 3 * ((blockIdx.x + t * gridDim.x) * kTileM + row0);     // offset of i3[e0 + row0][0]
#pragma unroll
            for (int jj = 0; jj < kTsCopies; ++jj) {
                const int64_t k = k0 + 3 * kRowStep * (jj / 3) + (jj % 3);
                ids[jj] = (t < my_tiles && k < 3 * E) ? __ldg(i3 + k) : -1;
            }
        };
        int ids[kTsCopies], nxt[kTsCopies];
        load_ids(0, ids);
        for (int64_t t = 0; t < my_tiles; ++t) {
            load_ids(t + 1, nxt);
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(smem_u32(&bar_gempty[gb]), gph ^ 1u);
                const uint32_t gbuf = gran_base + (uint32_t)gb * kTsGranuleBytes + off0;
                const float* col = xp + kc * kChunkK + 4 * chk;
#pragma unroll
                for (int jj = 0; jj < kTsCopies; ++jj) {
                    const bool ok = ids[jj] >= 0;
                    cp_async16_zfill(gbuf + (uint32_t)(((jj % 3) * kTileM + kRowStep * (jj / 3)) * kChunkBytesPerRow),
                                     col + (int64_t)(ok ? ids[jj] : 0) * xp_ld, ok);
                }
                cp_async_arrive(smem_u32(&bar_gfull[gb]));
                if (++gb == n_gran) gb = 0, gph ^= 1u;
            }
#pragma unroll
            for (int jj = 0; jj < kTsCopies; ++jj) ids[jj] = nxt[jj];
        }
        cp_async_wait_all();
    } else if (warp == kTsLoadWarp) {
        // ======================= weight loader =======================
        // Each CTA of the cluster fetches 1 / cl_size of every weight chunk and multicasts it to
        // all of them: the chunk crosses the L2 -> SM fabric once per cluster, not once per CTA
        // (weight streaming was 2/3 of this kernel's L2 traffic and the fabric, ~33 B/clk/SM with
        // every SM pulling, was the bound).
        if (lane == 0) {
            const uint32_t piece = w_stage_bytes / cl_size;
            uint32_t it = 0;
            for (int64_t t = 0; t < my_tiles; ++t)
                for (int kc = 0; kc < KC; ++kc)
                    for (int b = 0; b < nb; ++b, ++it) {
                        const int s = it % w_stages;
                        const uint32_t ph = (it / w_stages) & 1u;
                        mbar_wait(smem_u32(&bar_wempty[s]), ph ^ 1u);      // every CTA of the cluster released the stage
                        const uint32_t full = smem_u32(&bar_wfull[s]);
                        mbar_expect_tx_(full, w_stage_bytes);
                        const uint32_t dst = smem_base + (uint32_t)s * w_stage_bytes + cl_rank * piece;
                        const uint8_t* src = wprep + (int64_t)(b * KC + kc) * w_stage_bytes + cl_rank * piece;
                        if (cl_size > 1) bulk_g2s_multicast(dst, src, piece, full, cl_mask);
                        else bulk_g2s_(dst, src, piece, full);
                    }
        }
    } else if (warp == kTsMmaWarp) {
        // ======================= MMA issuer =======================
        // warp-uniform loop, one elected lane issues (see tc_common.cuh: MMA issue discipline)
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32(dim);
        const uint32_t b_tile_bytes = (uint32_t)dim * kChunkBytesPerRow;
        int sa = 0, sw = 0;
        uint32_t pa = 0, pw = 0;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = (uint32_t)t & 1u;
            mbar_wait(smem_u32(&bar_tempty[buf]), (((uint32_t)t >> 1) & 1u) ^ 1u);
            const uint32_t tmem_d = tmu + buf * (uint32_t)dim;
            for (int ci = 0; ci < chunks_per_tile; ++ci) {
                TS_PROBE(lane == 0, 4, t * chunks_per_tile + ci);
                mbar_wait(smem_u32(&bar_wfull[sw]), pw);
                mbar_wait(smem_u32(&bar_afull[sa]), pa);
                TS_PROBE(lane == 0, 5, t * chunks_per_tile + ci);
                fence_after_sync();
                const uint32_t a_hi = tmu + 2u * (uint32_t)dim + (uint32_t)sa * 64u, a_lo = a_hi + 32u;
                const uint32_t w_hi = smem_base + (uint32_t)sw * w_stage_bytes;
                const uint64_t dbh = make_kmajor_sw128_desc(w_hi);
                const uint64_t dbl = make_kmajor_sw128_desc(w_hi + b_tile_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks) {
                        const uint64_t bh = advance_desc_k(dbh, 8 * ks), bl = advance_desc_k(dbl, 8 * ks);
                        mma_tf32_ts(tmem_d, a_lo + 8u * ks, bh, idesc, (ci > 0 || ks > 0) ? 1u : 0u);   // small terms first
                        mma_tf32_ts(tmem_d, a_hi + 8u * ks, bl, idesc, 1u);
                        mma_tf32_ts(tmem_d, a_hi + 8u * ks, bh, idesc, 1u);
                    }
                    mma_commit(smem_u32(&bar_aempty[sa]));
                    if (cl_size > 1) mma_commit_multicast(smem_u32(&bar_wempty[sw]), cl_mask);
                    else mma_commit(smem_u32(&bar_wempty[sw]));
                    if (ci == chunks_per_tile - 1) mma_commit(smem_u32(&bar_tfull[buf]));
                }
                __syncwarp();
                if (++sa == a_stages) sa = 0, pa ^= 1u;
                if (++sw == w_stages) sw = 0, pw ^= 1u;
            }
        }
    } else {
        // ======================= epilogue =======================
        const int q4 = warp - kTsEpiWarp0;              // TMEM lane quadrant == warp % 4
        const uint32_t stg = epi_base + (uint32_t)q4 * kEpiStageBytes;
        const int c = lane & 7, rs = lane >> 3;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = (uint32_t)t & 1u;
            const int64_t e0 = (blockIdx.x + t * gridDim.x) * kTileM + q4 * 32;
            TS_PROBE(q4 == 0 && lane == 0, 1, 6000 + t);
            mbar_wait(smem_u32(&bar_tfull[buf]), ((uint32_t)t >> 1) & 1u);
            TS_PROBE(q4 == 0 && lane == 0, 2, 6000 + t);
            fence_after_sync();
            const uint32_t taddr = tmem_base + buf * (uint32_t)dim + ((uint32_t)(q4 * 32) << 16);
            for (int c0 = 0; c0 < dim; c0 += 32) {
                float acc[32];
                tmem_ld32(taddr + (uint32_t)c0, acc);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts4(stg + epi_off(lane, j), make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
                __syncwarp();
                const float4 bv = bias ? ldg4(bias + c0 + 4 * c) : f4_zero();
#pragma unroll
                for (int itr = 0; itr < 8; ++itr) {
                    const int r = itr * 4 + rs;
                    if (e0 + r < E) {
                        float4 o = lds4(stg + epi_off(r, c));
                        f4_add(o, bv);
                        stg4(ef + (e0 + r) * ef_ld + c0 + 4 * c, o);
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
            TS_PROBE(q4 == 0 && lane == 0, 3, 6000 + t);
        }
    }
    fence_before_sync();
    __syncthreads();
#ifdef IHG_TRACE
    if (tid == 0 && blockIdx.x < 160) g_ts_cta[blockIdx.x][1] = ts_gtime();
#endif
    if (cl_size > 1) cluster_sync();            // no peer may still multicast into / signal this CTA
    if (warp == kTsMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// full form: w = aggregation.weight (3 + nb blocks), bias [dim] or null
int launch_interact_fwd_full_ts(const float* xp, int64_t xp_ld, const float* w_agg, int64_t w_ld,
                                const float* bias, int nb, const int32_t* i3, int64_t E, float* ef,
                                int64_t ef_ld, int dim, void* workspace, cudaStream_t st) {
    uint8_t* wprep = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    if (int rc = launch_interact_prep(w_agg, w_ld, 3 + nb, dim, 0, wprep, st)) return rc;
    int a_stages = (512 - 2 * dim) / 64;
    if (a_stages > kTsMaxAStages) a_stages = kTsMaxAStages;
    const int w_stage_bytes = 2 * dim * kChunkBytesPerRow;
    const int n_gran = dim <= 64 ? 3 : 2;            // d = 128 needs the room for a third weight stage
    int w_stages = (226 * 1024 - n_gran * kTsGranuleBytes - 4 * kEpiStageBytes - 1024) / w_stage_bytes;
    if (w_stages > kTsMaxWStages) w_stages = kTsMaxWStages;
    const int smem = w_stages * w_stage_bytes + n_gran * kTsGranuleBytes + 4 * kEpiStageBytes + 1024;
    IHG_CUDA(cudaFuncSetAttribute(feature_interact_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    // clusters of CTAs share the weight stream; the grid is what the device can keep resident
    static int cluster = -1, max_ctas = 0;
    if (cluster < 0) {
        IHG_CUDA(cudaFuncSetAttribute(feature_interact_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        cluster = pick_cluster(feature_interact_fwd_ts_kernel, kTsThreads, 226 * 1024, 4, &max_ctas);
    }
    int64_t ctas = n_tiles < max_ctas ? n_tiles : max_ctas;
    ctas = (ctas + cluster - 1) / cluster * cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kTsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    IHG_CUDA(cudaLaunchKernelEx(&cfg, feature_interact_fwd_ts_kernel, xp, xp_ld, bias, (const uint8_t*)wprep, 3 + nb, i3, E,
                                ef, ef_ld, dim, a_stages, w_stages, n_gran));
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg

#ifdef IHG_TRACE
extern "C" int ihg_debug_read_trace(long long* dst, int n) {
    if (n > 8 * 8192) n = 8 * 8192;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(dst, ihg::g_ts_trace, (size_t)n * sizeof(long long));
}
extern "C" int ihg_debug_read_cta_times(long long* dst) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(dst, ihg::g_ts_cta, sizeof(long long) * 320);
}
#endif
