// Order 2/3 FeatureInteractor contraction on the 5th-generation tensor cores.
//
//   ef[e,:] = p[u]+p[q]+p[i] + W_hi . cat(u*q, q*i, i*u [, u*q*i])
//   (/root/reference/Models/CommonLayers.py:68-85; u,q,i = projected rows of the hyperedge's nodes)
//
// Warp-specialised persistent kernel, one CTA per SM, 128 hyperedges per tile:
//   warps 0-7   producers: gather 128-byte slices of the u/q/i rows, form the Hadamard products
//               in registers, split them into tf32 hi/lo and write the A operand straight into
//               the UMMA K-major SWIZZLE_128B layout in shared memory (the [E, K*d] concatenation
//               of the reference never exists anywhere);
//   warp  12    MMA issuer: tcgen05.mma kind::tf32, 3xTF32, fp32 accumulators in TMEM
//               (two accumulator buffers so the epilogue of tile t overlaps the MMAs of t+1);
//   warp  13    weight loader: cp.async.bulk of the pre-split, pre-swizzled weight chunks
//               (written once per call by interact_prep_weights_kernel) onto the stage mbarrier;
//   warps 8-11  epilogue: tcgen05.ld the accumulator row of each hyperedge, add the hoisted
//               first-order part p[u]+p[q]+p[i], store ef.
// Stages are handed over with mbarriers (full: 8 producer warps + weight bytes; empty:
// tcgen05.commit).
//
// Roofline: HBM-bound by design once the contraction is on tensor cores: per hyperedge
// 12 + 28*d bytes (i3, three xp rows, three p rows, one ef row) vs 3 * 2*nb*d^2 tf32 flops.
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

constexpr int kProducerWarps = 8;
constexpr int kEpilogueWarp0 = 8;     // warps 8..11  (warp % 4 == TMEM lane quadrant)
constexpr int kMmaWarp = 12;
constexpr int kLoadWarp = 13;
constexpr int kInteractThreads = 14 * 32;
constexpr int kATileBytes = kTileM * kChunkBytesPerRow;   // 16 KB: 128 rows x 128 B

__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// wprep[(b*KC + kc)] = { hi tile [dim rows x 128 B, SW128], lo tile } of
//   W_b[n][kc*32 .. kc*32+32) = w_hi[n*w_ld + b*dim + kc*32 + k]           (transposed == 0)
//   W_b^T[k][nc*32 .. +32)    = w_hi[(nc*32 + n)*w_ld + b*dim + k]         (transposed == 1)
__global__ void __launch_bounds__(256)
interact_prep_weights_kernel(const float* __restrict__ w_hi, int64_t w_ld, int nb, int dim,
                             int transposed, uint8_t* __restrict__ wprep) {
    const int KC = dim / kChunkK;
    const int tile_bytes = dim * kChunkBytesPerRow;
    const int64_t total = (int64_t)nb * KC * dim * 8;      // one 16-byte chunk per work item
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = idx & 7;
        const int row = (idx >> 3) % dim;
        const int kc = (idx / (8 * dim)) % KC;
        const int b = idx / ((int64_t)8 * dim * KC);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = kc * kChunkK + 4 * c + j;
            v[j] = transposed ? __ldg(w_hi + (int64_t)kk * w_ld + (int64_t)b * dim + row)
                              : __ldg(w_hi + (int64_t)row * w_ld + (int64_t)b * dim + kk);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_tf32(v[j], h[j], l[j]);
        uint8_t* tile = wprep + (int64_t)(b * KC + kc) * 2 * tile_bytes;
        const uint32_t off = sw128_offset(row, c);
        *reinterpret_cast<uint4*>(tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(tile + tile_bytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

struct InteractSmem {
    int stages;
    uint32_t stage_bytes;
    uint32_t b_tile_bytes;
};
static inline InteractSmem interact_smem(int dim) {
    InteractSmem s;
    s.b_tile_bytes = (uint32_t)dim * kChunkBytesPerRow;
    s.stage_bytes = 2 * kATileBytes + 2 * s.b_tile_bytes;
    s.stages = (int)((200 * 1024) / s.stage_bytes);   // + 16 KB epilogue staging stays < 227 KB
    if (s.stages > 6) s.stages = 6;
    return s;
}

// kFull = false: hoisted form -- only the nb product blocks are contracted, the epilogue adds the
//                 gathered first-order rows p[u]+p[q]+p[i] (used by the multi-GPU layer, where p
//                 travels with the halo exchange).
// kFull = true:  the whole FeatureInteractor.forward -- the raw u, q, i slices (already in the
//                 producers' registers) are three more A blocks contracted with
//                 aggregation.weight[:, :3d], so there is no p table, no p gather and the epilogue
//                 only adds the bias and stores.  nb counts ALL blocks (6 or 7) in this mode.
// Forward-kernel roles: 16 producer warps (an in-kernel clock64 trace showed the operand
// producers, not the gathers or the MMAs, bound the tile time; with 2 warps per scheduler their
// dependent ALU chains and proxy fences left the issue slots mostly idle), 4 epilogue warps,
// 1 MMA warp, 1 weight-loader warp.
constexpr int kFwdProducerWarps = 16;
constexpr int kFwdEpiWarp0 = kFwdProducerWarps;          // multiple of 4: warp % 4 == TMEM quadrant
constexpr int kFwdMmaWarp = kFwdProducerWarps + 4;
constexpr int kFwdLoadWarp = kFwdProducerWarps + 5;
constexpr int kFwdThreads = (kFwdProducerWarps + 6) * 32;
constexpr int kFwdItems = kTileM * 8 / (kFwdProducerWarps * 32);   // (row, chunk) items per producer thread
constexpr int kFwdRowStep = kFwdProducerWarps * 4;                 // rows covered by one pass of the producers

template <bool kFull>
__global__ void __launch_bounds__(kFwdThreads, 1)
edge_interact_fwd_tc_kernel(const float* __restrict__ xp, int64_t xp_ld, const float* __restrict__ p,
                            int64_t p_ld, const uint8_t* __restrict__ wprep, int nb,
                            const int32_t* __restrict__ i3, int64_t E, float* __restrict__ ef,
                            int64_t ef_ld, int dim, int stages, uint32_t stage_bytes) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[6], bar_empty[6], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = dim / kChunkK;
    const int chunks_per_tile = KC * nb;
    const uint32_t b_tile_bytes = (uint32_t)dim * kChunkBytesPerRow;
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(2 * dim));
    const uint32_t epi_base = smem_base + (uint32_t)stages * stage_bytes;   // 4 x 4 KB staging tiles

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), kFwdProducerWarps + 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), 4);
        }
        mbar_init_fence();
    }
    if (warp == kFwdMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < kFwdProducerWarps) {
        // ======================= producers =======================
        // lane mapping (row, chunk): 8 consecutive lanes read one row's 128-byte slice, so a warp
        // request touches 4 full lines; thread owns chunk c of rows r0 + kFwdRowStep * j.
        const int c = tid & 7, r0 = tid >> 3;
        uint32_t it = 0;                    // global chunk counter (stage ring position)
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const float *pu[kFwdItems], *pq[kFwdItems], *pi[kFwdItems];
            bool ok[kFwdItems];
#pragma unroll
            for (int j = 0; j < kFwdItems; ++j) {
                const int64_t e = tile * kTileM + r0 + kFwdRowStep * j;
                ok[j] = e < E;
                int nu = 0, nq = 0, ni = 0;
                if (ok[j]) {
                    nu = __ldg(i3 + 3 * e);
                    nq = __ldg(i3 + 3 * e + 1);
                    ni = __ldg(i3 + 3 * e + 2);
                }
                pu[j] = xp + (int64_t)nu * xp_ld + 4 * c;
                pq[j] = xp + (int64_t)nq * xp_ld + 4 * c;
                pi[j] = xp + (int64_t)ni * xp_ld + 4 * c;
            }
            for (int kc = 0; kc < KC; ++kc) {
                float4 u[kFwdItems], q[kFwdItems], v[kFwdItems];
#pragma unroll
                for (int j = 0; j < kFwdItems; ++j) {
                    u[j] = ok[j] ? ldg4(pu[j] + kc * kChunkK) : f4_zero();
                    q[j] = ok[j] ? ldg4(pq[j] + kc * kChunkK) : f4_zero();
                    v[j] = ok[j] ? ldg4(pi[j] + kc * kChunkK) : f4_zero();
                }
                for (int b = 0; b < nb; ++b, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t a_lo = a_hi + kATileBytes;
                    const int pb = kFull ? b - 3 : b;             // product block; < 0: raw row block b
#pragma unroll
                    for (int j = 0; j < kFwdItems; ++j) {
                        float4 z;
                        if (kFull && b == 0) z = u[j];
                        else if (kFull && b == 1) z = q[j];
                        else if (kFull && b == 2) z = v[j];
                        else if (pb == 0) z = f4_mul(u[j], q[j]);
                        else if (pb == 1) z = f4_mul(q[j], v[j]);
                        else if (pb == 2) z = f4_mul(v[j], u[j]);
                        else z = f4_mul(f4_mul(u[j], q[j]), v[j]);
                        store_split_chunk(a_hi, a_lo, r0 + kFwdRowStep * j, c, z);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
                }
            }
        }
    } else if (warp == kFwdLoadWarp) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < KC; ++kc)
                    for (int b = 0; b < nb; ++b, ++it) {
                        const int s = it % stages;
                        const uint32_t ph = (it / stages) & 1u;
                        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                        const uint32_t full = smem_u32(&bar_full[s]);
                        const uint32_t dst = smem_base + (uint32_t)s * stage_bytes + 2 * kATileBytes;
                        mbar_expect_tx(full, 2 * b_tile_bytes);
                        bulk_g2s(dst, wprep + (int64_t)(b * KC + kc) * 2 * b_tile_bytes, 2 * b_tile_bytes, full);
                    }
            }
        }
    } else if (warp == kFwdMmaWarp) {
        // ======================= MMA issuer =======================
        // warp-uniform loop, one elected lane issues (tc_common.cuh: MMA issue discipline)
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32(dim);
        uint32_t it = 0, t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            mbar_wait(smem_u32(&bar_tempty[buf]), ((t >> 1) & 1u) ^ 1u);
            const uint32_t tmem_d = tmu + buf * (uint32_t)dim;
            for (int ci = 0; ci < chunks_per_tile; ++ci, ++it) {
                const int s = it % stages;
                const uint32_t ph = (it / stages) & 1u;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                fence_after_sync();
                const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes;
                const uint64_t dah = make_kmajor_sw128_desc(a_hi);
                const uint64_t dal = make_kmajor_sw128_desc(a_hi + kATileBytes);
                const uint64_t dbh = make_kmajor_sw128_desc(a_hi + 2 * kATileBytes);
                const uint64_t dbl = make_kmajor_sw128_desc(a_hi + 2 * kATileBytes + b_tile_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks)
                        mma_3xtf32(tmem_d, advance_desc_k(dah, 8 * ks), advance_desc_k(dal, 8 * ks),
                                   advance_desc_k(dbh, 8 * ks), advance_desc_k(dbl, 8 * ks), idesc,
                                   (ci > 0 || ks > 0) ? 1u : 0u);
                    mma_commit(smem_u32(&bar_empty[s]));          // stage may be refilled
                    if (ci == chunks_per_tile - 1) mma_commit(smem_u32(&bar_tfull[buf]));   // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ======================= epilogue =======================
        const int q4 = warp - kFwdEpiWarp0;             // TMEM lane quadrant == warp % 4
        const uint32_t stg = epi_base + (uint32_t)q4 * kEpiStageBytes;
        const int c = lane & 7, rs = lane >> 3;         // (row-in-group, chunk) mapping for global access
        uint32_t t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            const int64_t e0 = tile * kTileM + q4 * 32;
            // node ids of the 8 rows this lane serves in the coalesced phase (rows rs, rs+4, ...)
            int nu[8], nq[8], ni[8];
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
                const int64_t e = e0 + itr * 4 + rs;
                const bool ok = e < E;
                nu[itr] = ok ? (kFull ? 0 : __ldg(i3 + 3 * e)) : -1;
                nq[itr] = (ok && !kFull) ? __ldg(i3 + 3 * e + 1) : 0;
                ni[itr] = (ok && !kFull) ? __ldg(i3 + 3 * e + 2) : 0;
            }
            mbar_wait(smem_u32(&bar_tfull[buf]), (t >> 1) & 1u);
            fence_after_sync();
            const uint32_t taddr = tmem_base + buf * (uint32_t)dim + ((uint32_t)(q4 * 32) << 16);
            for (int c0 = 0; c0 < dim; c0 += 32) {
                float acc[32];
                tmem_ld32(taddr + (uint32_t)c0, acc);
                __syncwarp();                             // previous slab fully read back
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts4(stg + epi_off(lane, j), make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
                __syncwarp();
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float4 base[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int itr = hb * 4 + k;
                        if (kFull) {
                            base[k] = p ? ldg4(p + c0 + 4 * c) : f4_zero();      // p = aggregation bias [dim]
                        } else if (nu[itr] >= 0) {
                            base[k] = ldg4(p + (int64_t)nu[itr] * p_ld + c0 + 4 * c);
                            f4_add(base[k], ldg4(p + (int64_t)nq[itr] * p_ld + c0 + 4 * c));
                            f4_add(base[k], ldg4(p + (int64_t)ni[itr] * p_ld + c0 + 4 * c));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int itr = hb * 4 + k;
                        if (nu[itr] >= 0) {
                            const int r = itr * 4 + rs;
                            float4 o = lds4(stg + epi_off(r, c));
                            f4_add(o, base[k]);
                            stg4(ef + (e0 + r) * ef_ld + c0 + 4 * c, o);
                        }
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kFwdMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}


// =========================================================================================
// backward (a): per-slot input gradients on tensor cores
//   dz_b[e][k] = sum_n def[e][n] * W_b[n][k]        (A = def tile, B = W_b^T chunks)
//   du = dz_uq*q + dz_iu*i + dz_uqi*q*i, ... written as slot_grad[e][slot][k]
// Work unit = (tile of 128 hyperedges, column half h): NU = dim (dim <= 64) or dim/2 output
// columns per unit, so that nb accumulators of NU columns fit TMEM twice (double buffering).
// =========================================================================================
constexpr int kSlotAStages = 2;
// Roles of the slot-gradient kernel: the def-tile producers are light, the product-rule epilogue
// (three gathered rows in, three gradient rows out per hyperedge) is the heavy part, so it gets
// 8 warps: two per TMEM lane quadrant, alternating 16-column slabs.
constexpr int kSlotProducerWarps = 4;
constexpr int kSlotEpiWarp0 = 4;          // warps 4..11; quadrant = warp % 4
constexpr int kSlotEpiWarps = 8;
constexpr int kSlotStageBytes = 32 * 64;  // per-warp staging tile: 32 rows x 16 fp32
__device__ __forceinline__ uint32_t epi_off16(int row, int chunk) {      // chunk in [0,4)
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__global__ void __launch_bounds__(kInteractThreads, 1)
edge_interact_bwd_slot_tc_kernel(const float* __restrict__ xp, int64_t xp_ld,
                                 const float* __restrict__ def, int64_t def_ld,
                                 const uint8_t* __restrict__ wprep_t, int nb,
                                 const int32_t* __restrict__ i3, int64_t E,
                                 float* __restrict__ slot_grad, int dim, int nu, int b_stages) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kSlotAStages], bar_aempty[kSlotAStages];
    __shared__ __align__(8) uint64_t bar_bfull[8], bar_bempty[8], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int32_t slot_ids[kSlotEpiWarps][3][32];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = dim / kChunkK;                 // chunks along the contraction (n)
    const int UH = dim / nu;                      // units per tile
    const uint32_t b_tile_bytes = (uint32_t)nu * kChunkBytesPerRow;
    const uint32_t a_stage_bytes = 2 * kATileBytes;
    const uint32_t b_base = smem_base + kSlotAStages * a_stage_bytes;
    const uint32_t b_stage_bytes = 2 * b_tile_bytes;
    const uint32_t epi_base = b_base + (uint32_t)b_stages * b_stage_bytes;   // 24 x 2 KB staging tiles
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const int64_t n_units = n_tiles * UH;
    const uint32_t acc_cols = (uint32_t)(nb * nu);
    const uint32_t tmem_cols = tmem_cols_pow2(2 * acc_cols);

    if (tid == 0) {
        for (int s = 0; s < kSlotAStages; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), kSlotProducerWarps);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < b_stages; ++s) {
            mbar_init(smem_u32(&bar_bfull[s]), 1);
            mbar_init(smem_u32(&bar_bempty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), kSlotEpiWarps);
        }
        mbar_init_fence();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < kSlotProducerWarps) {
        // ---- A producer: def tile rows, split to tf32 hi/lo; (row, chunk) lane mapping
        const int c = tid & 7, r0 = tid >> 3;                 // rows r0 + 16 j
        uint32_t it = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int64_t e0 = (u / UH) * kTileM;
            for (int nc = 0; nc < KC; ++nc, ++it) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int64_t e = e0 + r0 + 16 * j;
                    v[j] = e < E ? ldg4(def + e * def_ld + nc * kChunkK + 4 * c) : f4_zero();
                }
                const int s = it % kSlotAStages;
                mbar_wait(smem_u32(&bar_aempty[s]), ((it / kSlotAStages) & 1u) ^ 1u);
                const uint32_t a_hi = smem_base + (uint32_t)s * a_stage_bytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) store_split_chunk(a_hi, a_hi + kATileBytes, r0 + 16 * j, c, v[j]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_afull[s]));
            }
        }
    } else if (warp == kLoadWarp) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int h = (int)(u % UH);
                for (int nc = 0; nc < KC; ++nc)
                    for (int b = 0; b < nb; ++b, ++it) {
                        const int s = it % b_stages;
                        mbar_wait(smem_u32(&bar_bempty[s]), ((it / b_stages) & 1u) ^ 1u);
                        const uint32_t full = smem_u32(&bar_bfull[s]);
                        mbar_expect_tx(full, b_stage_bytes);
                        bulk_g2s(b_base + (uint32_t)s * b_stage_bytes,
                                 wprep_t + ((int64_t)(b * KC + nc) * UH + h) * b_stage_bytes, b_stage_bytes, full);
                    }
            }
        }
    } else if (warp == kMmaWarp) {
        // warp-uniform loop, one elected lane issues (tc_common.cuh: MMA issue discipline)
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32(nu);
        uint32_t ita = 0, itb = 0, t = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            mbar_wait(smem_u32(&bar_tempty[buf]), ((t >> 1) & 1u) ^ 1u);
            for (int nc = 0; nc < KC; ++nc, ++ita) {
                const int sa = ita % kSlotAStages;
                mbar_wait(smem_u32(&bar_afull[sa]), (ita / kSlotAStages) & 1u);
                const uint32_t a_hi = smem_base + (uint32_t)sa * a_stage_bytes;
                const uint64_t dah = make_kmajor_sw128_desc(a_hi);
                const uint64_t dal = make_kmajor_sw128_desc(a_hi + kATileBytes);
                for (int b = 0; b < nb; ++b, ++itb) {
                    const int sb = itb % b_stages;
                    mbar_wait(smem_u32(&bar_bfull[sb]), (itb / b_stages) & 1u);
                    fence_after_sync();
                    const uint32_t bh = b_base + (uint32_t)sb * b_stage_bytes;
                    const uint64_t dbh = make_kmajor_sw128_desc(bh);
                    const uint64_t dbl = make_kmajor_sw128_desc(bh + b_tile_bytes);
                    const uint32_t tmem_d = tmu + buf * acc_cols + (uint32_t)(b * nu);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < kChunkK / 8; ++ks)
                            mma_3xtf32(tmem_d, advance_desc_k(dah, 8 * ks), advance_desc_k(dal, 8 * ks),
                                       advance_desc_k(dbh, 8 * ks), advance_desc_k(dbl, 8 * ks), idesc,
                                       (nc > 0 || ks > 0) ? 1u : 0u);
                        mma_commit(smem_u32(&bar_bempty[sb]));
                        if (b == nb - 1) {
                            mma_commit(smem_u32(&bar_aempty[sa]));
                            if (nc == KC - 1) mma_commit(smem_u32(&bar_tfull[buf]));
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= kSlotEpiWarp0 && warp < kSlotEpiWarp0 + kSlotEpiWarps) {
        // ---- epilogue: product rule.  Global traffic (u,q,i gathers, slot_grad stores) uses a
        // coalesced (row, chunk) mapping; the math runs thread-per-row (TMEM order); three per-warp
        // staging tiles transpose between the two.  Slabs of 16 columns alternate between the two
        // warps of a quadrant.
        const int ew = warp - kSlotEpiWarp0;
        const int q4 = warp & 3, half = ew >> 2;
        const uint32_t su = epi_base + (uint32_t)(ew * 3) * kSlotStageBytes;
        const uint32_t sq = su + kSlotStageBytes, si = sq + kSlotStageBytes;
        const int c = lane & 3, rs = lane >> 2;              // 8 rows x 4 chunks per request
        uint32_t t = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            const int h = (int)(u % UH);
            const int64_t e0 = (u / UH) * kTileM + q4 * 32;
            __syncwarp();
            {
                const int64_t e = e0 + lane;
                const bool ok = e < E;
                slot_ids[ew][0][lane] = ok ? __ldg(i3 + 3 * e) : -1;
                slot_ids[ew][1][lane] = ok ? __ldg(i3 + 3 * e + 1) : 0;
                slot_ids[ew][2][lane] = ok ? __ldg(i3 + 3 * e + 2) : 0;
            }
            __syncwarp();
            const uint32_t taddr = tmem_base + buf * acc_cols + ((uint32_t)(q4 * 32) << 16);
            bool waited = false;
            for (int c0 = 16 * half; c0 < nu; c0 += 32) {
                const int col = h * nu + c0;
                __syncwarp();
                // load phase: asynchronous 16-byte copies straight into the staging tiles -- all 12
                // gathers of this lane are in flight at once and cost no registers
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                    const int r = itr * 8 + rs;
                    const int n0 = slot_ids[ew][0][r];
                    if (n0 >= 0) {
                        cp_async16(su + epi_off16(r, c), xp + (int64_t)n0 * xp_ld + col + 4 * c);
                        cp_async16(sq + epi_off16(r, c), xp + (int64_t)slot_ids[ew][1][r] * xp_ld + col + 4 * c);
                        cp_async16(si + epi_off16(r, c), xp + (int64_t)slot_ids[ew][2][r] * xp_ld + col + 4 * c);
                    } else {
                        sts4(su + epi_off16(r, c), f4_zero());
                        sts4(sq + epi_off16(r, c), f4_zero());
                        sts4(si + epi_off16(r, c), f4_zero());
                    }
                }
                cp_async_wait_all();
                __syncwarp();
                if (!waited) {
                    mbar_wait(smem_u32(&bar_tfull[buf]), (t >> 1) & 1u);
                    fence_after_sync();
                    waited = true;
                }
                // compute phase: thread = row `lane`, results overwrite the inputs
                {
                    float uu[16], qq[16], ii[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a = lds4(su + epi_off16(lane, j));
                        const float4 b = lds4(sq + epi_off16(lane, j));
                        const float4 d = lds4(si + epi_off16(lane, j));
                        uu[4 * j] = a.x; uu[4 * j + 1] = a.y; uu[4 * j + 2] = a.z; uu[4 * j + 3] = a.w;
                        qq[4 * j] = b.x; qq[4 * j + 1] = b.y; qq[4 * j + 2] = b.z; qq[4 * j + 3] = b.w;
                        ii[4 * j] = d.x; ii[4 * j + 1] = d.y; ii[4 * j + 2] = d.z; ii[4 * j + 3] = d.w;
                    }
                    float du[16], dq[16], di[16], dz[16];
                    tmem_ld16(taddr + (uint32_t)c0, dz);                              // b = 0: u*q
#pragma unroll
                    for (int j = 0; j < 16; ++j) { du[j] = dz[j] * qq[j]; dq[j] = dz[j] * uu[j]; }
                    tmem_ld16(taddr + (uint32_t)(nu + c0), dz);                       // b = 1: q*i
#pragma unroll
                    for (int j = 0; j < 16; ++j) { dq[j] = fmaf(dz[j], ii[j], dq[j]); di[j] = dz[j] * qq[j]; }
                    tmem_ld16(taddr + (uint32_t)(2 * nu + c0), dz);                   // b = 2: i*u
#pragma unroll
                    for (int j = 0; j < 16; ++j) { di[j] = fmaf(dz[j], uu[j], di[j]); du[j] = fmaf(dz[j], ii[j], du[j]); }
                    if (nb == 4) {
                        tmem_ld16(taddr + (uint32_t)(3 * nu + c0), dz);               // b = 3: u*q*i
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            du[j] = fmaf(dz[j], qq[j] * ii[j], du[j]);
                            dq[j] = fmaf(dz[j], uu[j] * ii[j], dq[j]);
                            di[j] = fmaf(dz[j], uu[j] * qq[j], di[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sts4(su + epi_off16(lane, j), make_float4(du[4 * j], du[4 * j + 1], du[4 * j + 2], du[4 * j + 3]));
                        sts4(sq + epi_off16(lane, j), make_float4(dq[4 * j], dq[4 * j + 1], dq[4 * j + 2], dq[4 * j + 3]));
                        sts4(si + epi_off16(lane, j), make_float4(di[4 * j], di[4 * j + 1], di[4 * j + 2], di[4 * j + 3]));
                    }
                }
                __syncwarp();
                // store phase: coalesced slot_grad rows
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                    const int r = itr * 8 + rs;
                    if (slot_ids[ew][0][r] >= 0) {
                        float* out = slot_grad + (e0 + r) * 3 * (int64_t)dim + col + 4 * c;
                        stg4(out, lds4(su + epi_off16(r, c)));
                        stg4(out + dim, lds4(sq + epi_off16(r, c)));
                        stg4(out + 2 * dim, lds4(si + epi_off16(r, c)));
                    }
                }
            }
            if (!waited) {                       // a warp without a slab still has to hand the buffer back
                mbar_wait(smem_u32(&bar_tfull[buf]), (t >> 1) & 1u);
                fence_after_sync();
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}

// Weight chunks for the slot kernel: tile (b, nc, h) = rows k in [h*nu, (h+1)*nu) of
//   W_b^T[k][nc*32 .. +32) = w_hi[(nc*32 + n)*w_ld + b*dim + k], hi tile then lo tile.
__global__ void __launch_bounds__(256)
interact_prep_weights_t_kernel(const float* __restrict__ w_hi, int64_t w_ld, int nb, int dim, int nu,
                               uint8_t* __restrict__ wprep) {
    const int KC = dim / kChunkK, UH = dim / nu;
    const int tile_bytes = nu * kChunkBytesPerRow;
    const int64_t total = (int64_t)nb * KC * dim * 8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = idx & 7;
        const int k = (idx >> 3) % dim;                 // output column == B-tile row
        const int nc = (idx / (8 * dim)) % KC;
        const int b = idx / ((int64_t)8 * dim * KC);
        const int h = k / nu, r = k % nu;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = nc * kChunkK + 4 * c + j;
            split_tf32(__ldg(w_hi + (int64_t)n * w_ld + (int64_t)b * dim + k), hi[j], lo[j]);
        }
        uint8_t* tile = wprep + ((int64_t)(b * KC + nc) * UH + h) * 2 * tile_bytes;
        const uint32_t off = sw128_offset(r, c);
        *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(tile + tile_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// =========================================================================================
// backward (b): weight gradient on tensor cores, MN-major operands
//   dw_b^T[k][n] = sum_e z_b[e][k] * def[e][n]
// Both operands are stored as [edge rows x 32 features] sub-tiles (128 B per edge row) in the
// MN-major SWIZZLE_128B_BASE32B layout; the edge index is the MMA K dimension.  A = 128 product features (a "group": 4 sub-tiles), B = def (dim/32 sub-tiles).
// The accumulators (G groups x dim columns) stay in TMEM across all tiles of the CTA; one
// partial [G*128, dim] per CTA goes to the workspace and a second kernel sums them in order.
// =========================================================================================
// Small CTAs (4 producer warps + 1 MMA warp, 32 hyperedges per tile, two or more CTAs resident
// per SM) so that the gather latency of one CTA overlaps the staging / MMAs of the others.  All
// u/q/i and def slices of a tile are fetched in ONE burst of independent 128-bit loads per thread
// (coalesced (row, chunk) mapping), then every operand sub-tile is produced from registers.
// blockIdx.y selects a range of feature groups when all groups would not fit TMEM twice.
constexpr int kWgTe = 32;                       // hyperedges per tile
constexpr int kWgProducerWarps = 8;               // (row, chunk) per thread: 32 rows x 8 chunks = 256 threads
constexpr int kWgRowsPerThread = kWgTe * 8 / (kWgProducerWarps * 32);
constexpr int kWgRowStep = kWgProducerWarps * 4;  // rows covered by one pass of the producers
constexpr int kWgThreads = (kWgProducerWarps + 1) * 32;
constexpr int kWgMaxKC = 4;                     // dim <= 128

// TKC = dim / 32 as a compile-time constant (1, 2, 4) lets the register-resident row slices be
// indexed statically (the kernel is bound by the producers' instruction issue); TKC = 0 is the
// generic path (dim = 96).
template <int TKC>
__global__ void __launch_bounds__(kWgThreads)
edge_interact_bwd_wgrad_tc_kernel(const float* __restrict__ xp, int64_t xp_ld,
                                  const float* __restrict__ def, int64_t def_ld, int nb,
                                  const int32_t* __restrict__ i3, int64_t E, int dim, int G, int gpc,
                                  int nbs, float* __restrict__ ws_dw) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[2], bar_aempty[2], bar_bfull[2], bar_bempty[2], bar_done;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = TKC > 0 ? TKC : dim / kChunkK;          // 32-feature blocks of def / of one product
    constexpr int NBLK = TKC > 0 ? TKC : kWgMaxKC;         // register-resident blocks per row
    const int g0 = blockIdx.y * gpc;                       // first feature group of this CTA
    const int ng = min(gpc, G - g0);                       // groups handled here
    constexpr uint32_t sub_bytes = kWgTe * kChunkBytesPerRow;      // one [32 x 128 B] sub-tile
    constexpr uint32_t a_stage_bytes = 8 * sub_bytes;      // 4 sub-tiles hi + 4 lo
    const uint32_t b_stage_bytes = 2 * KC * sub_bytes;     // KC sub-tiles hi + KC lo
    const uint32_t b_base = smem_base + 2 * a_stage_bytes;
    const int64_t n_tiles = (E + kWgTe - 1) / kWgTe;
    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(gpc * dim));
    const bool is_mma_warp = warp == kWgProducerWarps;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), kWgProducerWarps);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
            mbar_init(smem_u32(&bar_bfull[s]), kWgProducerWarps);
            mbar_init(smem_u32(&bar_bempty[s]), 1);
        }
        mbar_init(smem_u32(&bar_done), 1);
        mbar_init_fence();
    }
    if (is_mma_warp) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (!is_mma_warp) {
        // (row, chunk) lane mapping: 8 consecutive lanes cover one row's 128-byte slice; the 256
        // producer threads cover the 32 rows of a tile (8 warps: with 4 the producers' instruction
        // issue -- products, tf32 splits, swizzled stores -- had too few warps to hide ALU latency)
        const int c = tid & 7, r0 = tid >> 3;
        uint32_t ita = 0, itb = 0;
        // node ids of the next tile are fetched one tile ahead: the row gathers of a tile then start
        // without waiting for their own index loads
        int nxt_u[kWgRowsPerThread], nxt_q[kWgRowsPerThread], nxt_i[kWgRowsPerThread];
        auto load_ids = [&](int64_t tile) {
#pragma unroll
            for (int j = 0; j < kWgRowsPerThread; ++j) {
                const int64_t e = tile * kWgTe + r0 + kWgRowStep * j;
                const bool ok = tile < n_tiles && e < E;
                nxt_u[j] = ok ? __ldg(i3 + 3 * e) : 0;
                nxt_q[j] = ok ? __ldg(i3 + 3 * e + 1) : 0;
                nxt_i[j] = ok ? __ldg(i3 + 3 * e + 2) : 0;
            }
        };
        load_ids(blockIdx.x);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++itb) {
            // ---- one burst of loads: def and u/q/i slices of this thread's row(s), all 32-column blocks
            float4 dv[kWgRowsPerThread][NBLK], uv[kWgRowsPerThread][NBLK], qv[kWgRowsPerThread][NBLK], iv[kWgRowsPerThread][NBLK];
#pragma unroll
            for (int j = 0; j < kWgRowsPerThread; ++j) {
                const int64_t e = tile * kWgTe + r0 + kWgRowStep * j;
                const bool ok = e < E;
                const int nu_ = nxt_u[j], nq = nxt_q[j], ni = nxt_i[j];
#pragma unroll
                for (int blk = 0; blk < NBLK; ++blk) {
                    const bool okb = ok && blk < KC;
                    dv[j][blk] = okb ? ldg4(def + e * def_ld + blk * kChunkK + 4 * c) : f4_zero();
                    uv[j][blk] = okb ? ldg4(xp + (int64_t)nu_ * xp_ld + blk * kChunkK + 4 * c) : f4_zero();
                    qv[j][blk] = okb ? ldg4(xp + (int64_t)nq * xp_ld + blk * kChunkK + 4 * c) : f4_zero();
                    iv[j][blk] = okb ? ldg4(xp + (int64_t)ni * xp_ld + blk * kChunkK + 4 * c) : f4_zero();
                }
            }
            load_ids(tile + gridDim.x);
            // ---- B stage: def tile
            {
                const int sb = itb % nbs;
                mbar_wait(smem_u32(&bar_bempty[sb]), ((itb / nbs) & 1u) ^ 1u);
                const uint32_t bh = b_base + (uint32_t)sb * b_stage_bytes;
#pragma unroll
                for (int blk = 0; blk < NBLK; ++blk)
                    if (blk < KC) {
#pragma unroll
                        for (int j = 0; j < kWgRowsPerThread; ++j)
                            store_split_chunk_mn(bh + (uint32_t)blk * sub_bytes, bh + (uint32_t)(KC + blk) * sub_bytes,
                                                 r0 + kWgRowStep * j, c, dv[j][blk]);
                    }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_bfull[sb]));
            }
            // ---- A stages: one group of 128 product features each, straight from registers
            for (int g = g0; g < g0 + ng; ++g, ++ita) {
                const int sa = ita & 1;
                mbar_wait(smem_u32(&bar_aempty[sa]), ((ita >> 1) & 1u) ^ 1u);
                const uint32_t ah = smem_base + (uint32_t)sa * a_stage_bytes;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    // sub-tile st = 4 g + j4 of the product features: block b = st / KC, columns blk = st % KC
                    int b, blk;
                    if (TKC == 4) { b = g; blk = j4; }
                    else if (TKC == 2) { b = 2 * g + (j4 >> 1); blk = j4 & 1; }
                    else if (TKC == 1) { b = 4 * g + j4; blk = 0; }
                    else { const int st = 4 * g + j4; b = st / KC; blk = st % KC; }
#pragma unroll
                    for (int j = 0; j < kWgRowsPerThread; ++j) {
                        float4 u = f4_zero(), q = f4_zero(), v = f4_zero();
                        if (TKC > 0) {
                            // blk is a compile-time constant after unrolling: plain register reads
                            u = uv[j][TKC == 4 ? j4 : (TKC == 2 ? (j4 & 1) : 0)];
                            q = qv[j][TKC == 4 ? j4 : (TKC == 2 ? (j4 & 1) : 0)];
                            v = iv[j][TKC == 4 ? j4 : (TKC == 2 ? (j4 & 1) : 0)];
                        } else {
#pragma unroll
                            for (int k = 0; k < NBLK; ++k)
                                if (k == blk) { u = uv[j][k]; q = qv[j][k]; v = iv[j][k]; }
                        }
                        float4 z = f4_zero();
                        if (b == 0) z = f4_mul(u, q);
                        else if (b == 1) z = f4_mul(q, v);
                        else if (b == 2) z = f4_mul(v, u);
                        else if (b == 3 && nb == 4) z = f4_mul(f4_mul(u, q), v);
                        store_split_chunk_mn(ah + (uint32_t)j4 * sub_bytes, ah + (uint32_t)(4 + j4) * sub_bytes,
                                             r0 + kWgRowStep * j, c, z);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_afull[sa]));
            }
        }
        // ---- final epilogue (same warps: warp == TMEM lane quadrant): dump the partial dw^T
        const int r = warp * 32 + lane;                             // product feature within the group
        float* out = ws_dw + (int64_t)blockIdx.x * G * 128 * dim;
        if (my_tiles > 0) {
            mbar_wait(smem_u32(&bar_done), 0);
            fence_after_sync();
        }
        for (int g = 0; g < ng && warp < 4; ++g)            // warps 0-3 own the four TMEM lane quadrants
            for (int c0 = 0; c0 < dim; c0 += 16) {
                float acc[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + (uint32_t)(g * dim + c0) + ((uint32_t)(warp * 32) << 16), acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    stg4(out + ((int64_t)(g0 + g) * 128 + r) * dim + c0 + 4 * j,
                         make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
            }
    } else {
        // warp-uniform loop, one elected lane issues (tc_common.cuh: MMA issue discipline)
        const uint32_t tmu = warp_uniform(tmem_base);
        const uint32_t idesc = make_idesc_tf32_mn(dim);
        uint32_t ita = 0, itb = 0;
        bool first = true;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++itb) {
            const int sb = itb % nbs;
            mbar_wait(smem_u32(&bar_bfull[sb]), (itb / nbs) & 1u);
            const uint32_t bh = b_base + (uint32_t)sb * b_stage_bytes;
            for (int g = 0; g < ng; ++g, ++ita) {
                const int sa = ita & 1;
                mbar_wait(smem_u32(&bar_afull[sa]), (ita >> 1) & 1u);
                fence_after_sync();
                const uint32_t ah = smem_base + (uint32_t)sa * a_stage_bytes;
                const uint32_t tmem_d = tmu + (uint32_t)(g * dim);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kWgTe / 8; ++ks) {
                        const uint32_t koff = (uint32_t)ks * 1024u;   // 8 edge rows x 128 B
                        const uint64_t dah = make_mnmajor_sw128_desc(ah + koff, sub_bytes);
                        const uint64_t dal = make_mnmajor_sw128_desc(ah + 4 * sub_bytes + koff, sub_bytes);
                        const uint64_t dbh = make_mnmajor_sw128_desc(bh + koff, sub_bytes);
                        const uint64_t dbl = make_mnmajor_sw128_desc(bh + (uint32_t)KC * sub_bytes + koff, sub_bytes);
                        mma_3xtf32(tmem_d, dah, dal, dbh, dbl, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                    mma_commit(smem_u32(&bar_aempty[sa]));
                    if (g == ng - 1) mma_commit(smem_u32(&bar_bempty[sb]));
                }
                __syncwarp();
            }
            first = false;
        }
        if (elect_one()) mma_commit(smem_u32(&bar_done));
        __syncwarp();
    }
    fence_before_sync();
    __syncthreads();
    if (is_mma_warp) tmem_dealloc(tmem_base, tmem_cols);
}

// dw_hi[n][b*dim + k] = sum_cta ws[cta][b*dim + k][n]   (ascending cta: deterministic)
// 256 threads = 32 outputs x 8 slices of the CTA range, combined in ascending slice order.
__global__ void __launch_bounds__(256)
interact_wgrad_tc_reduce_kernel(const float* __restrict__ ws, int n_cta, int G, int nb, int dim,
                                float* __restrict__ dw_hi) {
    __shared__ float part[8][32];
    const int64_t total = (int64_t)nb * dim * dim;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int64_t idx = (int64_t)blockIdx.x * 32 + lane;
    const int n = (int)(idx % dim);              // fastest: coalesced reads of ws rows
    const int64_t f = idx / dim;                 // product feature b*dim + k
    float s = 0.f;
    if (idx < total) {
        const int c0 = (int)((int64_t)slice * n_cta / 8), c1 = (int)((int64_t)(slice + 1) * n_cta / 8);
        for (int c = c0; c < c1; ++c) s += ws[((int64_t)c * G * 128 + f) * dim + n];
    }
    part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && idx < total) {
        float t = part[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += part[q][lane];
        dw_hi[(int64_t)n * nb * dim + f] = t;
    }
}

bool interact_tc_eligible(int dim) {
    static const bool disabled = getenv("IHG_DISABLE_TC") != nullptr;
    return !disabled && dim % 32 == 0 && dim >= 32 && dim <= 128;
}

int64_t interact_fwd_tc_workspace_bytes(int dim, int nb) {
    return (int64_t)(3 + nb) * (dim / kChunkK) * 2 * dim * kChunkBytesPerRow + 1024;   // sized for the full form
}

int launch_interact_prep(const float* w_hi, int64_t w_ld, int nb, int dim, int transposed,
                         void* wprep, cudaStream_t st) {
    const int64_t total = (int64_t)nb * (dim / kChunkK) * dim * 8;
    interact_prep_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        w_hi, w_ld, nb, dim, transposed, static_cast<uint8_t*>(wprep));
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

template <bool kFull>
static int launch_interact_fwd_tc_impl(const float* xp, int64_t xp_ld, const float* p, int64_t p_ld,
                                       const float* w, int64_t w_ld, int nblk, const int32_t* i3, int64_t E,
                                       float* ef, int64_t ef_ld, int dim, void* workspace, cudaStream_t st) {
    // 1024-byte aligned weight staging area inside the caller's workspace
    uint8_t* wprep = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    if (int rc = launch_interact_prep(w, w_ld, nblk, dim, 0, wprep, st)) return rc;
    const InteractSmem cfg = interact_smem(dim);
    const int smem = cfg.stages * (int)cfg.stage_bytes + 4 * kEpiStageBytes + 1024;
    IHG_CUDA(cudaFuncSetAttribute(edge_interact_fwd_tc_kernel<kFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    edge_interact_fwd_tc_kernel<kFull><<<grid, kFwdThreads, smem, st>>>(xp, xp_ld, p, p_ld, wprep, nblk, i3, E, ef,
                                                                             ef_ld, dim, cfg.stages, cfg.stage_bytes);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

// hoisted form: w = aggregation.weight[:, 3d:] (nb product blocks), p = first-order rows [N, dim]
int launch_interact_fwd_tc(const float* xp, int64_t xp_ld, const float* p, int64_t p_ld,
                           const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                           float* ef, int64_t ef_ld, int dim, void* workspace, cudaStream_t st) {
    return launch_interact_fwd_tc_impl<false>(xp, xp_ld, p, p_ld, w_hi, w_ld, nb, i3, E, ef, ef_ld, dim, workspace, st);
}
// full form: w = aggregation.weight (3 + nb blocks), bias [dim] or null
int launch_interact_fwd_full_tc(const float* xp, int64_t xp_ld, const float* w_agg, int64_t w_ld,
                                const float* bias, int nb, const int32_t* i3, int64_t E, float* ef,
                                int64_t ef_ld, int dim, void* workspace, cudaStream_t st) {
    return launch_interact_fwd_tc_impl<true>(xp, xp_ld, bias, 0, w_agg, w_ld, 3 + nb, i3, E, ef, ef_ld, dim, workspace, st);
}

static int slot_nu(int dim) { return dim > 64 ? dim / 2 : dim; }
static int wgrad_groups(int dim, int nb) { return (nb * dim + 127) / 128; }
// groups per CTA: keep the persistent accumulators within 256 TMEM columns so that two CTAs fit an SM
static int wgrad_groups_per_cta(int dim, int nb) {
    const int G = wgrad_groups(dim, nb);
    int gpc = 256 / dim;
    if (gpc < 1) gpc = 1;
    return gpc < G ? gpc : G;
}
constexpr int kWgCtasX = 2 * kNumSMs;     // CTAs along x times the group split = resident CTAs

int64_t interact_bwd_tc_workspace_bytes(int dim, int nb) {
    const int64_t wprep = (int64_t)nb * (dim / kChunkK) * 2 * dim * kChunkBytesPerRow + 1024;
    const int64_t partial = (int64_t)kWgCtasX * wgrad_groups(dim, nb) * 128 * dim * 4;
    return wprep + partial + 1024;
}

int launch_interact_bwd_tc(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                           const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                           float* slot_grad, float* dw_hi, int dim, void* workspace, cudaStream_t st) {
    uint8_t* wprep = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const int64_t wprep_bytes = (int64_t)nb * (dim / kChunkK) * 2 * dim * kChunkBytesPerRow;
    float* partial = reinterpret_cast<float*>(wprep + ((wprep_bytes + 1023) & ~(int64_t)1023));
    // ---- (a) slot gradients
    const int nu = slot_nu(dim);
    static const bool slot_ss = getenv("IHG_SLOT_SS") != nullptr;      // A/B switch: both operands in shared memory
    if (!slot_ss) {
        if (int rc = launch_interact_bwd_slot_ts(xp, xp_ld, def, def_ld, w_hi, w_ld, nb, i3, E, slot_grad, dim, wprep, st))
            return rc;
    } else {
        const int64_t total = (int64_t)nb * (dim / kChunkK) * dim * 8;
        interact_prep_weights_t_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_hi, w_ld, nb, dim, nu, wprep);
        IHG_LAUNCH_CHECK();
        const uint32_t b_stage = 2u * (uint32_t)nu * kChunkBytesPerRow;
        const int epi_bytes = kSlotEpiWarps * 3 * kSlotStageBytes;                 // 48 KB
        int b_stages = (int)((220 * 1024 - epi_bytes - kSlotAStages * 2 * kATileBytes) / b_stage);
        if (b_stages > 8) b_stages = 8;
        const int smem = kSlotAStages * 2 * kATileBytes + b_stages * (int)b_stage + epi_bytes + 1024;
        static int attr_smem = 0;
        if (attr_smem < smem) {
            IHG_CUDA(cudaFuncSetAttribute(edge_interact_bwd_slot_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr_smem = smem;
        }
        const int64_t n_units = ((E + kTileM - 1) / kTileM) * (dim / nu);
        const unsigned grid = (unsigned)(n_units < kNumSMs ? n_units : kNumSMs);
        edge_interact_bwd_slot_tc_kernel<<<grid, kInteractThreads, smem, st>>>(xp, xp_ld, def, def_ld, wprep, nb, i3, E,
                                                                               slot_grad, dim, nu, b_stages);
        IHG_LAUNCH_CHECK();
    }
    // ---- (b) weight gradient
    {
        const int G = wgrad_groups(dim, nb);
        const int KC = dim / kChunkK;
        const int gpc = wgrad_groups_per_cta(dim, nb);            // groups per CTA (TMEM: gpc*dim <= 256 cols)
        const int gy = (G + gpc - 1) / gpc;
        const int nbs = dim > 64 ? 1 : 2;                          // def-tile stages (two CTAs must fit an SM)
        const int smem = 2 * 8 * kWgTe * kChunkBytesPerRow + nbs * 2 * KC * kWgTe * kChunkBytesPerRow + 1024;
        const int64_t n_tiles = (E + kWgTe - 1) / kWgTe;
        const int gx = (int)(n_tiles < kWgCtasX / gy ? n_tiles : kWgCtasX / gy);
        dim3 grid(gx, gy);
#define IHG_WG_LAUNCH(T)                                                                                            \
        do {                                                                                                        \
            IHG_CUDA(cudaFuncSetAttribute(edge_interact_bwd_wgrad_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            edge_interact_bwd_wgrad_tc_kernel<T><<<grid, kWgThreads, smem, st>>>(xp, xp_ld, def, def_ld, nb, i3, E, dim, G, gpc, nbs, partial); \
        } while (0)
        if (KC == 4) IHG_WG_LAUNCH(4);
        else if (KC == 2) IHG_WG_LAUNCH(2);
        else if (KC == 1) IHG_WG_LAUNCH(1);
        else IHG_WG_LAUNCH(0);
#undef IHG_WG_LAUNCH
        IHG_LAUNCH_CHECK();
        const int64_t total = (int64_t)nb * dim * dim;
        interact_wgrad_tc_reduce_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(partial, gx, G, nb, dim, dw_hi);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}

}  // namespace ihg
