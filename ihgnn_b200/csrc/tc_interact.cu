// Order 2/3 FeatureInteractor on the 5th-generation tensor cores: the pieces shared by the
// forward (tc_interact_ts.cu) and backward (tc_interact_slot_ts.cu) kernels plus the weight
// gradient of the product blocks.
//
//   ef[e,:] = W_a . cat(u, q, i, u*q, q*i, i*u [, u*q*i]) + b
//   (/root/reference/Models/CommonLayers.py:68-85; u,q,i = projected rows of the hyperedge's nodes)
//
//   interact_prep_weights_kernel       weight blocks split into tf32 hi/lo, written once per call in
//                                      the UMMA K-major SWIZZLE_128B layout (cp.async.bulk-able chunks)
//   edge_interact_bwd_wgrad_tc_kernel  dW_hi = sum_e z[e]^T def[e] (MN-major operands, accumulators
//                                      resident in TMEM, fixed-order cross-CTA sum)
//   launch_interact_bwd_tc             backward = slot gradients (tc_interact_slot_ts.cu) + dW_hi
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

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

bool interact_tc_eligible(int dim) { return dim % 32 == 0 && dim >= 32 && dim <= 128; }

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
    if (int rc = launch_interact_bwd_slot_ts(xp, xp_ld, def, def_ld, w_hi, w_ld, nb, i3, E, slot_grad, dim, wprep, st))
        return rc;
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
