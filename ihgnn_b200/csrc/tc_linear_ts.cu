// Typed node Linear, A operand in tensor memory, weights resident in shared memory.
//
//   y[r,:] = x[r,:] . W[t(r)]^T (+ bias[t(r)]) (+ addend[r,:])        t(r) = node type of row r
//   (feature_transform Models/GnnLayers.py:224; the first-order blocks of `aggregation`
//    Models/CommonLayers.py:66,85 hoisted to node level; and their input-gradient transposes)
//
// Persistent CTAs, one per SM.  Each CTA serves ONE node type (CTAs are dealt to the types in
// proportion to their tile counts), so it splits that type's weight matrix to tf32 hi/lo once,
// keeps it in shared memory in the UMMA K-major SWIZZLE_128B layout for its whole life, and then
// streams 128-row tiles:
//   warps 17-20 gather: cp.async 128-byte row slices (32 columns) into a ring of 16 KB granules;
//   warps 0-7   producers, two groups of four warps alternating granules: thread = row, split the
//               32 values to hi/lo and tcgen05.st them into the TMEM A ring;
//   warp 16     MMA issuer: per granule 4 K-steps x 3 tcgen05.mma (A from TMEM, B = resident W);
//   warps 8-15  epilogue, two warps per TMEM lane quadrant alternating 32-column slabs (an
//               in-kernel trace showed the epilogue, ~1.9 K cycles per slab, pacing the kernel):
//               tcgen05.ld, bias (preloaded), addend, staged
//               coalesced store.
// TMEM: [0, 2*n_out) two accumulator buffers, then A stages of 64 columns (32 hi + 32 lo).
// Roofline: HBM, 4*(n_in + n_out) bytes per row (+ 4*n_out with an addend).
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

#ifdef IHG_TRACE
__device__ long long g_lt_trace[8][4096];
#define LT_PROBE(cond, rowi, idx)                                                     \
    do {                                                                              \
        if (blockIdx.x == 0 && (cond) && (idx) < 4096) g_lt_trace[rowi][idx] = clock64(); \
    } while (0)
#else
#define LT_PROBE(cond, rowi, idx) \
    do {                          \
    } while (0)
#endif

namespace {

constexpr int kLtProducerWarps = 8;
constexpr int kLtGroups = kLtProducerWarps / 4;
constexpr int kLtEpiWarp0 = 8;
constexpr int kLtEpiWarps = 8;             // two per TMEM lane quadrant, alternating 32-column slabs
constexpr int kLtMmaWarp = 16;
constexpr int kLtGatherWarp0 = 17;
constexpr int kLtGatherWarps = 4;
constexpr int kLtGatherThreads = kLtGatherWarps * 32;
constexpr int kLtThreads = (kLtGatherWarp0 + kLtGatherWarps) * 32;
constexpr int kLtGranuleBytes = kTileM * kChunkBytesPerRow;      // 16 KB: 128 rows x 32 columns
constexpr int kLtMaxGranules = 10;        // ring depth is set per launch from the shared memory left by W
constexpr int kLtMaxAStages = 6;

struct LtTypes {
    int64_t lo[3], hi[3];      // row range of each node type
    int cta0[4];               // CTAs [cta0[t], cta0[t+1]) serve type t
};

__device__ __forceinline__ void lt_cp16_zfill(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void lt_cp_arrive(uint32_t mbar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void lt_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void lt_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kLtThreads, 1)
node_linear_ts_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ w, int n_types,
                      int n_out, int n_in, int transpose_w, const float* __restrict__ bias,
                      const float* __restrict__ addend, int64_t addend_ld, LtTypes tt, float* __restrict__ y,
                      int64_t y_ld, int a_stages, int n_gran) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kLtMaxAStages], bar_aempty[kLtMaxAStages];
    __shared__ __align__(8) uint64_t bar_gfull[kLtMaxGranules], bar_gempty[kLtMaxGranules];
    __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = n_in / kChunkK;
    const uint32_t w_tile_bytes = (uint32_t)n_out * kChunkBytesPerRow;       // one [n_out x 32] hi or lo tile
    // shared memory map: [W: KC x (hi, lo)][granule ring][4 epilogue staging tiles]
    const uint32_t gran_base = smem_base + (uint32_t)KC * 2u * w_tile_bytes;
    const uint32_t epi_base = gran_base + (uint32_t)n_gran * kLtGranuleBytes;

    // (selects, not indexed loads: a dynamically indexed by-value struct would live in local memory)
    const int bx = (int)blockIdx.x;
    const int type = bx >= tt.cta0[2] ? 2 : (bx >= tt.cta0[1] ? 1 : 0);
    const int c_lo = type == 2 ? tt.cta0[2] : (type == 1 ? tt.cta0[1] : tt.cta0[0]);
    const int c_hi = type == 2 ? tt.cta0[3] : (type == 1 ? tt.cta0[2] : tt.cta0[1]);
    const int cta_in_type = bx - c_lo, ctas_of_type = c_hi - c_lo;
    const int64_t row_lo = type == 2 ? tt.lo[2] : (type == 1 ? tt.lo[1] : tt.lo[0]);
    const int64_t row_hi = type == 2 ? tt.hi[2] : (type == 1 ? tt.hi[1] : tt.hi[0]);
    const int64_t n_tiles = (row_hi - row_lo + kTileM - 1) / kTileM;
    const int64_t my_tiles = n_tiles > cta_in_type ? (n_tiles - cta_in_type + ctas_of_type - 1) / ctas_of_type : 0;
    const int wt = n_types > 1 ? type : 0;

    LT_PROBE(tid == 0, 7, 0);
    if (tid == 0) {
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), 4);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < n_gran; ++s) {
            mbar_init(smem_u32(&bar_gfull[s]), kLtGatherThreads);
            mbar_init(smem_u32(&bar_gempty[s]), 4);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), kLtEpiWarps);
        }
        mbar_init_fence();
    }
    if (warp == kLtMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), 512);
    // resident weights: element (n, k) of this type's matrix -> chunk k / 32, row n, split hi / lo
    {
        const float* W = w + (int64_t)wt * n_out * n_in;
        const int total = KC * n_out * 8;                      // 16-byte chunks
        for (int idx = tid; idx < total; idx += kLtThreads) {
            const int c = idx & 7;
            const int n = (idx >> 3) % n_out;
            const int kc = idx / (8 * n_out);
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kc * kChunkK + 4 * c + j;
                const float v = transpose_w ? __ldg(W + (int64_t)k * n_out + n) : __ldg(W + (int64_t)n * n_in + k);
                split_tf32(v, h[j], l[j]);
            }
            const uint32_t tile = smem_base + (uint32_t)kc * 2u * w_tile_bytes;
            const uint32_t off = sw128_offset(n, c);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile + w_tile_bytes + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
        fence_async_smem();                                   // generic-proxy writes -> tensor-core reads
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_a0 = tmem_base + 2u * (uint32_t)n_out;
    LT_PROBE(tid == 0, 7, 1);

    if (warp < kLtProducerWarps) {
        // ======================= producers =======================
        const int group = warp >> 2, quad = warp & 3;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int64_t total = my_tiles * KC;                   // granules of this CTA
        int s = group, gb = group;                              // A ring slot, granule ring slot (kLtGroups <= both ring sizes)
        uint32_t ph = 0, gph = 0;
        for (int64_t c = group; c < total; c += kLtGroups) {
            LT_PROBE(tid == 0, 0, c >> 1);
            mbar_wait(smem_u32(&bar_gfull[gb]), gph);
            LT_PROBE(tid == 0, 1, c >> 1);
            mbar_wait(smem_u32(&bar_aempty[s]), ph ^ 1u);
            LT_PROBE(tid == 0, 1, 2048 + (c >> 1));
            fence_after_sync();
            const uint32_t gbuf = gran_base + (uint32_t)gb * kLtGranuleBytes + (uint32_t)(quad * 32) * kChunkBytesPerRow;
            const uint32_t ta = tmem_a0 + (uint32_t)s * 64u + lane_addr;
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                const float4 a = lds4(gbuf + epi_off(lane, 2 * pass));
                const float4 b = lds4(gbuf + epi_off(lane, 2 * pass + 1));
                const float z[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    hi[k] = rna_tf32(z[k]);
                    lo[k] = __float_as_uint(z[k] - __uint_as_float(hi[k]));     // the tensor core reads its top 19 bits
                }
                lt_tmem_st8(ta + 8u * pass, hi);
                lt_tmem_st8(ta + 32u + 8u * pass, lo);
            }
            lt_tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&bar_afull[s]));
                mbar_arrive(smem_u32(&bar_gempty[gb]));
            }
            LT_PROBE(tid == 0, 2, c >> 1);
            s += kLtGroups;
            if (s >= a_stages) s -= a_stages, ph ^= 1u;
            gb += kLtGroups;
            if (gb >= n_gran) gb -= n_gran, gph ^= 1u;
        }
    } else if (warp >= kLtGatherWarp0) {
        // ======================= gather =======================
        const int gt = tid - kLtGatherWarp0 * 32;
        const int chk = gt & 7, row0 = gt >> 3;                 // copy j: chunk chk of row row0 + 16 j
        int gb = 0;
        uint32_t gph = 0;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int64_t r0 = row_lo + (cta_in_type + t * ctas_of_type) * kTileM;
            for (int kc = 0; kc < KC; ++kc) {
                LT_PROBE(gt == 0, 3, t * KC + kc);
                mbar_wait(smem_u32(&bar_gempty[gb]), gph ^ 1u);
                LT_PROBE(gt == 0, 3, 2048 + t * KC + kc);
                const uint32_t gbuf = gran_base + (uint32_t)gb * kLtGranuleBytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = row0 + 16 * j;
                    const bool ok = r0 + r < row_hi;
                    // granule layout: four 32-row groups, each [32 rows x 128 B] with the epi_off swizzle
                    lt_cp16_zfill(gbuf + (uint32_t)(r >> 5) * 4096u + epi_off(r & 31, chk),
                                  x + (ok ? r0 + r : row_lo) * x_ld + kc * kChunkK + 4 * chk, ok);
                }
                lt_cp_arrive(smem_u32(&bar_gfull[gb]));
                LT_PROBE(gt == 0, 4, t * KC + kc);
                if (++gb == n_gran) gb = 0, gph ^= 1u;
            }
        }
        cp_async_wait_all();
    } else if (warp == kLtMmaWarp) {
        // ======================= MMA issuer =======================
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32(n_out);
        int sa = 0;
        uint32_t pa = 0;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = (uint32_t)t & 1u;
            LT_PROBE(lane == 0, 5, t);
            mbar_wait(smem_u32(&bar_tempty[buf]), (((uint32_t)t >> 1) & 1u) ^ 1u);
            LT_PROBE(lane == 0, 5, 2048 + t);
            const uint32_t tmem_d = tmu + buf * (uint32_t)n_out;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(smem_u32(&bar_afull[sa]), pa);
                fence_after_sync();
                const uint32_t a_hi = tmu + 2u * (uint32_t)n_out + (uint32_t)sa * 64u, a_lo = a_hi + 32u;
                const uint32_t w_hi = smem_base + (uint32_t)kc * 2u * w_tile_bytes;
                const uint64_t dbh = make_kmajor_sw128_desc(w_hi);
                const uint64_t dbl = make_kmajor_sw128_desc(w_hi + w_tile_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks) {
                        const uint64_t bh = advance_desc_k(dbh, 8 * ks), bl = advance_desc_k(dbl, 8 * ks);
                        mma_tf32_ts(tmem_d, a_lo + 8u * ks, bh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ts(tmem_d, a_hi + 8u * ks, bl, idesc, 1u);
                        mma_tf32_ts(tmem_d, a_hi + 8u * ks, bh, idesc, 1u);
                    }
                    mma_commit(smem_u32(&bar_aempty[sa]));
                    if (kc == KC - 1) mma_commit(smem_u32(&bar_tfull[buf]));
                }
                __syncwarp();
                if (++sa == a_stages) sa = 0, pa ^= 1u;
            }
        }
    } else {
        // ======================= epilogue =======================
        const int ew = warp - kLtEpiWarp0;
        const int q4 = ew & 3, half = ew >> 2;
        const uint32_t stg = epi_base + (uint32_t)ew * kEpiStageBytes;
        const int c = lane & 7, rs = lane >> 3;
        const float* bias_t = bias ? bias + (int64_t)wt * n_out : nullptr;
        // this warp's slabs: columns 32 (2 i + half); at most two of them for n_out <= 128
        float4 bv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c0 = 32 * (2 * i + half);
            bv[i] = (bias_t && c0 + 4 * c < n_out) ? ldg4(bias_t + c0 + 4 * c) : f4_zero();
        }
        for (int64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = (uint32_t)t & 1u;
            const int64_t r0 = row_lo + (cta_in_type + t * ctas_of_type) * kTileM + q4 * 32;
            LT_PROBE(ew == 0 && lane == 0, 6, t);
            mbar_wait(smem_u32(&bar_tfull[buf]), ((uint32_t)t >> 1) & 1u);
            LT_PROBE(ew == 0 && lane == 0, 6, 1024 + t);
            fence_after_sync();
            const uint32_t taddr = tmem_base + buf * (uint32_t)n_out + ((uint32_t)(q4 * 32) << 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c0 = 32 * (2 * i + half);
                if (c0 >= n_out) break;
                const int width = min(32, n_out - c0);          // n_out % 16 == 0: the last slab may be 16 wide
                const bool col_ok = 4 * c < width;
                float acc[32];
                if (width == 32) {
                    tmem_ld32(taddr + (uint32_t)c0, acc);
                } else {
                    float lo16[16];
                    tmem_ld16(taddr + (uint32_t)c0, lo16);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = lo16[j], acc[16 + j] = 0.f;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts4(stg + epi_off(lane, j), make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
                __syncwarp();
                if (col_ok) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int r = itr * 4 + rs;
                        if (r0 + r < row_hi) {
                            float4 o = lds4(stg + epi_off(r, c));
                            f4_add(o, bv[i]);
                            if (addend) f4_add(o, ldg4(addend + (r0 + r) * addend_ld + c0 + 4 * c));
                            stg4(y + (r0 + r) * y_ld + c0 + 4 * c, o);
                        }
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
            LT_PROBE(ew == 0 && lane == 0, 6, 2048 + t);
        }
    }
    fence_before_sync();
    __syncthreads();
    LT_PROBE(tid == 0, 7, 2);
    if (warp == kLtMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool node_linear_ts_eligible(int n_out, int n_in, int64_t x_ld, int64_t y_ld, const float* addend, int64_t addend_ld) {
    if (n_in % 32 != 0 || n_out % 16 != 0 || n_out < 16 || n_out > 128 || n_in > 128) return false;
    if (x_ld % 4 != 0 || y_ld % 4 != 0 || (addend && addend_ld % 4 != 0)) return false;
    const int64_t smem = (int64_t)8 * n_in * n_out + 3 * kLtGranuleBytes + kLtEpiWarps * kEpiStageBytes + 1024;
    return smem <= 226 * 1024;
}

int launch_node_linear_ts(const float* x, int64_t x_ld, const float* w, int n_types, int n_out, int n_in,
                          int transpose_w, const float* bias, const float* addend, int64_t addend_ld,
                          int64_t n_rows, int64_t bound0, int64_t bound1, float* y, int64_t y_ld, cudaStream_t st) {
    LtTypes tt;
    tt.lo[0] = 0, tt.hi[0] = bound0 < n_rows ? bound0 : n_rows;
    tt.lo[1] = tt.hi[0], tt.hi[1] = bound1 < n_rows ? bound1 : n_rows;
    tt.lo[2] = tt.hi[1], tt.hi[2] = n_rows;
    int64_t tiles[3], total = 0;
    for (int t = 0; t < 3; ++t) tiles[t] = (tt.hi[t] - tt.lo[t] + kTileM - 1) / kTileM, total += tiles[t];
    if (total == 0) return IHG_OK;
    // deal the CTAs to the types in proportion to their tile counts (>= 1 for a non-empty type)
    const int64_t grid = total < kNumSMs ? total : kNumSMs;
    int ctas[3], used = 0;
    for (int t = 0; t < 3; ++t) {
        ctas[t] = tiles[t] == 0 ? 0 : (int)((tiles[t] * grid) / total);
        if (tiles[t] > 0 && ctas[t] == 0) ctas[t] = 1;
        used += ctas[t];
    }
    while (used > grid) {                                        // rounding pushed us over: shrink the largest share
        int big = 0;
        for (int t = 1; t < 3; ++t) if (ctas[t] > ctas[big]) big = t;
        --ctas[big], --used;
    }
    while (used < grid) {                                        // leftovers go where a CTA has the most tiles
        int best = -1;
        for (int t = 0; t < 3; ++t)
            if (tiles[t] > ctas[t] && (best < 0 || tiles[t] * ctas[best] > tiles[best] * ctas[t])) best = t;
        if (best < 0) break;
        ++ctas[best], ++used;
    }
    tt.cta0[0] = 0;
    for (int t = 0; t < 3; ++t) tt.cta0[t + 1] = tt.cta0[t] + ctas[t];
    int a_stages = (512 - 2 * n_out) / 64;
    if (a_stages > kLtMaxAStages) a_stages = kLtMaxAStages;
    int n_gran = (226 * 1024 - 8 * n_in * n_out - kLtEpiWarps * kEpiStageBytes - 1024) / kLtGranuleBytes;
    if (n_gran > kLtMaxGranules) n_gran = kLtMaxGranules;     // bytes in flight per SM = ring depth x 16 KB
    const int smem = 8 * n_in * n_out + n_gran * kLtGranuleBytes + kLtEpiWarps * kEpiStageBytes + 1024;
    static int attr_smem = 0;
    if (attr_smem < smem) {
        IHG_CUDA(cudaFuncSetAttribute(node_linear_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    node_linear_ts_kernel<<<(unsigned)used, kLtThreads, smem, st>>>(x, x_ld, w, n_types, n_out, n_in, transpose_w, bias,
                                                                    addend, addend_ld, tt, y, y_ld, a_stages, n_gran);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg

#ifdef IHG_TRACE
extern "C" int ihg_debug_read_trace_linear(long long* dst, int n) {
    if (n > 8 * 4096) n = 8 * 4096;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(dst, ihg::g_lt_trace, (size_t)n * sizeof(long long));
}
#endif
