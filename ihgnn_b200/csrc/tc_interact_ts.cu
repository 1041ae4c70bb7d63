// Order 2/3 FeatureInteractor.forward, A operand in tensor memory ("TS" form of tcgen05.mma).
//
//   ef[e,:] = aggregation.weight . cat(u, q, i, u*q, q*i, i*u [, u*q*i]) + bias
//   (/root/reference/Models/CommonLayers.py:68-85; u,q,i = projected rows of the hyperedge's nodes)
//
// Why: with both operands in shared memory the 3xTF32 contraction is shared-memory-bandwidth
// bound -- per 32-wide K chunk the tensor core reads A_hi twice, A_lo once (48 KB) and the
// producers write 32 KB of split operand, ~100 KB of traffic for 384 cycles of MMA at d = 64
// (measured: 13 K cycles per 128-hyperedge tile against a 5.4 K-cycle tensor floor).  Here the
// split A operand never touches shared memory: the producers build it in registers and write it
// to TMEM with tcgen05.st (thread = one hyperedge row = one TMEM lane, no swizzle, no proxy
// fence), and the MMA reads A from TMEM and only the weight chunk from shared memory.
//
// Persistent, warp-specialised, one CTA per SM, 128 hyperedges per tile:
//   warps 0-15  producers.  A "granule" is (tile, 32-column slice): the 128-byte slices of the
//               u/q/i rows are gathered with cp.async (8 lanes per row: coalesced, no
//               registers) into a double-buffered staging area one granule ahead; then thread
//               (row = 32*(warp%4)+lane, column group = warp/4) reads its 8 columns of u, q, i
//               once and produces the 6-7 operand blocks hi/lo into the TMEM A ring;
//   warp 20     MMA issuer: 4 K-steps x 3 tcgen05.mma (A in TMEM, B descriptor) per chunk;
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

constexpr int kTsProducerWarps = 16;
constexpr int kTsProducerThreads = kTsProducerWarps * 32;
constexpr int kTsEpiWarp0 = kTsProducerWarps;
constexpr int kTsMmaWarp = kTsProducerWarps + 4;
constexpr int kTsLoadWarp = kTsProducerWarps + 5;
constexpr int kTsThreads = (kTsProducerWarps + 6) * 32;
constexpr int kTsGranuleBytes = 3 * kTileM * kChunkBytesPerRow;           // 48 KB: u, q, i slices
constexpr int kTsCopies = 3 * kTileM * 8 / kTsProducerThreads;            // 16-byte cp.async per thread per granule
constexpr int kTsMaxAStages = 6;
constexpr int kTsMaxWStages = 6;

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
__device__ __forceinline__ void producer_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(kTsProducerThreads) : "memory");
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

__global__ void __launch_bounds__(kTsThreads, 1)
feature_interact_fwd_ts_kernel(const float* __restrict__ xp, int64_t xp_ld, const float* __restrict__ bias,
                               const uint8_t* __restrict__ wprep, int nb, const int32_t* __restrict__ i3,
                               int64_t E, float* __restrict__ ef, int64_t ef_ld, int dim, int a_stages,
                               int w_stages) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kTsMaxAStages], bar_aempty[kTsMaxAStages];
    __shared__ __align__(8) uint64_t bar_wfull[kTsMaxWStages], bar_wempty[kTsMaxWStages];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = dim / kChunkK;
    const int chunks_per_tile = KC * nb;
    const uint32_t w_stage_bytes = 2u * (uint32_t)dim * kChunkBytesPerRow;     // hi + lo weight chunk
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // shared memory map: [W ring][2 granules][4 epilogue staging tiles]
    const uint32_t gran_base = smem_base + (uint32_t)w_stages * w_stage_bytes;
    const uint32_t epi_base = gran_base + 2u * kTsGranuleBytes;

    if (tid == 0) {
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), kTsProducerWarps);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < w_stages; ++s) {
            mbar_init(smem_u32(&bar_wfull[s]), 1);
            mbar_init(smem_u32(&bar_wempty[s]), 1);
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
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_a0 = tmem_base + 2u * (uint32_t)dim;

    if (warp < kTsProducerWarps) {
        // ======================= producers =======================
        const int64_t G = my_tiles * KC;                    // granules of this CTA
        const int quad = warp & 3, cg = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        // cp.async mapping: copy j of this thread = (row, table, 16-byte chunk)
        int crow[kTsCopies], ctab[kTsCopies], cchk[kTsCopies];
#pragma unroll
        for (int j = 0; j < kTsCopies; ++j) {
            const int idx = tid + kTsProducerThreads * j;
            crow[j] = idx / 24;
            ctab[j] = (idx % 24) >> 3;
            cchk[j] = idx & 7;
        }
        int ids_cur[kTsCopies], ids_nxt[kTsCopies];
        auto load_ids = [&](int64_t tile_local, int (&ids)[kTsCopies]) {
#pragma unroll
            for (int j = 0; j < kTsCopies; ++j) {
                const int64_t e = (blockIdx.x + tile_local * gridDim.x) * kTileM + crow[j];
                ids[j] = (tile_local < my_tiles && e < E) ? __ldg(i3 + 3 * e + ctab[j]) : -1;
            }
        };
        auto issue = [&](int64_t g) {
            const int kc = (int)(g % KC);
            const uint32_t buf = gran_base + (uint32_t)(g & 1) * kTsGranuleBytes;
#pragma unroll
            for (int j = 0; j < kTsCopies; ++j) {
                const bool ok = ids_cur[j] >= 0;
                const float* src = xp + (int64_t)(ok ? ids_cur[j] : 0) * xp_ld + kc * kChunkK + 4 * cchk[j];
                cp_async16_zfill(buf + stage_off(ctab[j], crow[j], cchk[j]), src, ok);
            }
        };
        load_ids(0, ids_cur);
        if (G > 0) issue(0);
        load_ids(1, ids_nxt);
        uint32_t it = 0;                                    // A-ring position
        int pending = -1;                                   // A stage written but not yet published
        for (int64_t g = 0; g < G; ++g) {
            TS_PROBE(tid == 0, 0, g);
            cp_async_wait_all();
            TS_PROBE(tid == 0, 6, g);
            producer_barrier();                             // granule g landed; everyone is done reading g-1
            TS_PROBE(tid == 0, 7, g);
            if (g + 1 < G) {
                const bool new_tile = (g + 1) % KC == 0;
                if (new_tile) {
#pragma unroll
                    for (int j = 0; j < kTsCopies; ++j) ids_cur[j] = ids_nxt[j];
                }
                issue(g + 1);
                if (new_tile) load_ids((g + 1) / KC + 1, ids_nxt);
            }
            TS_PROBE(tid == 0, 4, 7000 + g);
            const uint32_t buf = gran_base + (uint32_t)(g & 1) * kTsGranuleBytes;
            float u[8], q[8], v[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 a = lds4(buf + stage_off(0, row, 2 * cg + h));
                const float4 b = lds4(buf + stage_off(1, row, 2 * cg + h));
                const float4 c = lds4(buf + stage_off(2, row, 2 * cg + h));
                u[4 * h] = a.x, u[4 * h + 1] = a.y, u[4 * h + 2] = a.z, u[4 * h + 3] = a.w;
                q[4 * h] = b.x, q[4 * h + 1] = b.y, q[4 * h + 2] = b.z, q[4 * h + 3] = b.w;
                v[4 * h] = c.x, v[4 * h + 1] = c.y, v[4 * h + 2] = c.z, v[4 * h + 3] = c.w;
            }
            TS_PROBE(tid == 0, 5, 7000 + g);
            for (int b = 0; b < nb; ++b, ++it) {
                const int s = it % a_stages;
                const uint32_t ph = (it / a_stages) & 1u;
                float z[8];
                switch (b) {
                    case 0:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = u[k];
                        break;
                    case 1:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = q[k];
                        break;
                    case 2:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = v[k];
                        break;
                    case 3:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = u[k] * q[k];
                        break;
                    case 4:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = q[k] * v[k];
                        break;
                    case 5:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = v[k] * u[k];
                        break;
                    default:
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = u[k] * q[k] * v[k];
                        break;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) split_tf32(z[k], hi[k], lo[k]);
                TS_PROBE(tid == 0, 6, 1000 + it);
                // publish the previous chunk only now: its tcgen05.st latency overlapped the split above
                if (pending >= 0) {
                    tmem_st_wait();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_afull[pending]));
                }
                TS_PROBE(tid == 0, 1, it);
                mbar_wait(smem_u32(&bar_aempty[s]), ph ^ 1u);
                TS_PROBE(tid == 0, 2, it);
                fence_after_sync();
                const uint32_t ta = tmem_a0 + (uint32_t)s * 64u + (uint32_t)(8 * cg) + lane_addr;
                tmem_st8(ta, hi);
                tmem_st8(ta + 32u, lo);
                pending = s;
                TS_PROBE(tid == 0, 3, it);
            }
        }
        if (pending >= 0) {
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_afull[pending]));
        }
    } else if (warp == kTsLoadWarp) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t t = 0; t < my_tiles; ++t)
                for (int kc = 0; kc < KC; ++kc)
                    for (int b = 0; b < nb; ++b, ++it) {
                        const int s = it % w_stages;
                        const uint32_t ph = (it / w_stages) & 1u;
                        mbar_wait(smem_u32(&bar_wempty[s]), ph ^ 1u);
                        const uint32_t full = smem_u32(&bar_wfull[s]);
                        mbar_expect_tx_(full, w_stage_bytes);
                        bulk_g2s_(smem_base + (uint32_t)s * w_stage_bytes,
                                  wprep + (int64_t)(b * KC + kc) * w_stage_bytes, w_stage_bytes, full);
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
            fence_after_sync();
            const uint32_t tmem_d = tmu + buf * (uint32_t)dim;
            for (int ci = 0; ci < chunks_per_tile; ++ci) {
                TS_PROBE(lane == 0, 4, t * chunks_per_tile + ci);
                mbar_wait(smem_u32(&bar_wfull[sw]), pw);
                TS_PROBE(lane == 0, 0, 4096 + t * chunks_per_tile + ci);
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
                    mma_commit(smem_u32(&bar_wempty[sw]));
                }
                __syncwarp();
                if (++sa == a_stages) sa = 0, pa ^= 1u;
                if (++sw == w_stages) sw = 0, pw ^= 1u;
            }
            if (elect_one()) mma_commit(smem_u32(&bar_tfull[buf]));
            __syncwarp();
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
    int w_stages = (226 * 1024 - 2 * kTsGranuleBytes - 4 * kEpiStageBytes - 1024) / w_stage_bytes;
    if (w_stages > kTsMaxWStages) w_stages = kTsMaxWStages;
    const int smem = w_stages * w_stage_bytes + 2 * kTsGranuleBytes + 4 * kEpiStageBytes + 1024;
    IHG_CUDA(cudaFuncSetAttribute(feature_interact_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    feature_interact_fwd_ts_kernel<<<grid, kTsThreads, smem, st>>>(xp, xp_ld, bias, wprep, 3 + nb, i3, E, ef, ef_ld,
                                                                   dim, a_stages, w_stages);
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
#endif
