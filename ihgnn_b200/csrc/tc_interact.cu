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
    s.stages = (int)((200 * 1024) / s.stage_bytes);
    if (s.stages > 6) s.stages = 6;
    return s;
}

__global__ void __launch_bounds__(kInteractThreads, 1)
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

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), kProducerWarps + 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_tfull[s]), 1);
            mbar_init(smem_u32(&bar_tempty[s]), 4);
        }
        mbar_init_fence();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < kProducerWarps) {
        // ======================= producers =======================
        const int row = tid & 127;          // hyperedge row of the tile
        const int half = tid >> 7;          // which 64-byte half of the 128-byte slice
        uint32_t it = 0;                    // global chunk counter (stage ring position)
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t e = tile * kTileM + row;
            const bool ok = e < E;
            int nu = 0, nq = 0, ni = 0;
            if (ok) {
                nu = __ldg(i3 + 3 * e);
                nq = __ldg(i3 + 3 * e + 1);
                ni = __ldg(i3 + 3 * e + 2);
            }
            const float* pu = xp + (int64_t)nu * xp_ld + 16 * half;
            const float* pq = xp + (int64_t)nq * xp_ld + 16 * half;
            const float* pi = xp + (int64_t)ni * xp_ld + 16 * half;
            for (int kc = 0; kc < KC; ++kc) {
                float4 u[4], q[4], v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    u[c] = ok ? ldg4(pu + kc * kChunkK + 4 * c) : f4_zero();
                    q[c] = ok ? ldg4(pq + kc * kChunkK + 4 * c) : f4_zero();
                    v[c] = ok ? ldg4(pi + kc * kChunkK + 4 * c) : f4_zero();
                }
                for (int b = 0; b < nb; ++b, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t a_lo = a_hi + kATileBytes;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float4 z;
                        if (b == 0) z = f4_mul(u[c], q[c]);
                        else if (b == 1) z = f4_mul(q[c], v[c]);
                        else if (b == 2) z = f4_mul(v[c], u[c]);
                        else z = f4_mul(f4_mul(u[c], q[c]), v[c]);
                        store_split_chunk(a_hi, a_lo, row, 4 * half + c, z);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
                }
            }
        }
    } else if (warp == kLoadWarp) {
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
    } else if (warp == kMmaWarp) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(dim);
            uint32_t it = 0, t = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1u;
                mbar_wait(smem_u32(&bar_tempty[buf]), ((t >> 1) & 1u) ^ 1u);
                fence_after_sync();
                const uint32_t tmem_d = tmem_base + buf * (uint32_t)dim;
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
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks)
                        mma_3xtf32(tmem_d, advance_desc_k(dah, 8 * ks), advance_desc_k(dal, 8 * ks),
                                   advance_desc_k(dbh, 8 * ks), advance_desc_k(dbl, 8 * ks), idesc,
                                   (ci > 0 || ks > 0) ? 1u : 0u);
                    mma_commit(smem_u32(&bar_empty[s]));          // stage may be refilled
                }
                mma_commit(smem_u32(&bar_tfull[buf]));            // accumulator complete
            }
        }
    } else {
        // ======================= epilogue =======================
        const int q4 = warp - kEpilogueWarp0;           // TMEM lane quadrant == warp % 4
        const int row = q4 * 32 + lane;
        uint32_t t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1u;
            const int64_t e = tile * kTileM + row;
            const bool ok = e < E;
            int nu = 0, nq = 0, ni = 0;
            if (ok) {
                nu = __ldg(i3 + 3 * e);
                nq = __ldg(i3 + 3 * e + 1);
                ni = __ldg(i3 + 3 * e + 2);
            }
            const float* pu = p + (int64_t)nu * p_ld;
            const float* pq = p + (int64_t)nq * p_ld;
            const float* pi = p + (int64_t)ni * p_ld;
            mbar_wait(smem_u32(&bar_tfull[buf]), (t >> 1) & 1u);
            fence_after_sync();
            const uint32_t taddr = tmem_base + buf * (uint32_t)dim + ((uint32_t)(q4 * 32) << 16);
            for (int c0 = 0; c0 < dim; c0 += 16) {
                float4 base[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ok) {
                        base[j] = ldg4(pu + c0 + 4 * j);
                        f4_add(base[j], ldg4(pq + c0 + 4 * j));
                        f4_add(base[j], ldg4(pi + c0 + 4 * j));
                    } else {
                        base[j] = f4_zero();
                    }
                }
                float acc[16];
                tmem_ld16(taddr + (uint32_t)c0, acc);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 o = make_float4(base[j].x + acc[4 * j], base[j].y + acc[4 * j + 1],
                                               base[j].z + acc[4 * j + 2], base[j].w + acc[4 * j + 3]);
                        stg4(ef + e * ef_ld + c0 + 4 * j, o);
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
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}

bool interact_tc_eligible(int dim) {
    static const bool disabled = getenv("IHG_DISABLE_TC") != nullptr;
    return !disabled && dim % 32 == 0 && dim >= 32 && dim <= 128;
}

int64_t interact_fwd_tc_workspace_bytes(int dim, int nb) {
    return (int64_t)nb * (dim / kChunkK) * 2 * dim * kChunkBytesPerRow + 1024;
}

int launch_interact_prep(const float* w_hi, int64_t w_ld, int nb, int dim, int transposed,
                         void* wprep, cudaStream_t st) {
    const int64_t total = (int64_t)nb * (dim / kChunkK) * dim * 8;
    interact_prep_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        w_hi, w_ld, nb, dim, transposed, static_cast<uint8_t*>(wprep));
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int launch_interact_fwd_tc(const float* xp, int64_t xp_ld, const float* p, int64_t p_ld,
                           const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                           float* ef, int64_t ef_ld, int dim, void* workspace, cudaStream_t st) {
    // 1024-byte aligned weight staging area inside the caller's workspace
    uint8_t* wprep = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    if (int rc = launch_interact_prep(w_hi, w_ld, nb, dim, 0, wprep, st)) return rc;
    const InteractSmem cfg = interact_smem(dim);
    const int smem = cfg.stages * (int)cfg.stage_bytes + 1024;
    static int attr_smem = 0;
    if (attr_smem < smem) {
        IHG_CUDA(cudaFuncSetAttribute(edge_interact_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    const int64_t n_tiles = (E + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    edge_interact_fwd_tc_kernel<<<grid, kInteractThreads, smem, st>>>(xp, xp_ld, p, p_ld, wprep, nb, i3, E, ef,
                                                                      ef_ld, dim, cfg.stages, cfg.stage_bytes);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg
