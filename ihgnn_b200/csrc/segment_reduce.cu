// K-B: deterministic segmented reduction over CSR rows (the edge->node scatter-mean and all
// of its transposes).  Replaces torch_sparse.matmul(incidence, Ef) * Dv^-1
// (/root/reference/Models/GnnLayers.py:233-234), the index_put_(accumulate=True) backward of
// the three row gathers (CommonLayers.py:70-72) and EmbeddingBag(mean) (EmbeddingLayers.py:79).
//
// Work decomposition: the plan (ihg_segment_plan_build) lists row chunks of <= chunk_len
// incidences as 16-byte records {begin, end, row, partial slot}.  A group of LPR lanes (LPR x
// float4 spans the feature dimension; 32/LPR groups per warp) owns one chunk at a time (and can
// walk a strided sequence of chunks with the next records prefetched), keeping kSegUnroll
// independent 128-bit row gathers in flight.  Rows longer than chunk_len (Zipf head) are split;
// their partial sums are combined by a second kernel (one warp per split row, fixed interleave
// + shuffle tree).
// Summation order is a pure function of the plan => bitwise run-to-run determinism, no float
// atomics.
//
// Roofline: HBM.  Algorithmic bytes per incidence: 4 (col) + 4*dim (source row); per row:
// 16 (plan) + 4 (scale) + 4*dim (output row).
#include <stdlib.h>

#include "common.cuh"

namespace ihg {

// Routed output (multi-GPU, ihg_*_routed): consecutive row ranges of the result go to different destinations --
// the own rows to a local matrix, each peer's halo rows straight into that peer's receive buffer over NVLink
// (posted stores that overlap the kernel's own gathers), so the partial sums need no separate transfer pass.
constexpr int kMaxRoute = 16;
struct OutRoute {
    int32_t n;                        // ranges in use
    int32_t start[kMaxRoute + 1];     // range g = rows [start[g], start[g+1])
    float* base[kMaxRoute];           // where row start[g] goes (rows of a range are out_ld floats apart)
};
template <bool ROUTED>
__device__ __forceinline__ float* out_row(const OutRoute& r, float* out, int64_t out_ld, int row) {
    if (!ROUTED) return out + (int64_t)row * out_ld;
    int g = 0;
#pragma unroll
    for (int k = 1; k < kMaxRoute; ++k) g += (k < r.n && row >= r.start[k]) ? 1 : 0;
    return r.base[g] + (int64_t)(row - r.start[g]) * out_ld;
}

constexpr int kSegWarpsPerBlock = 8;
constexpr int kSegUnroll = 4;        // independent 128-bit row gathers in flight per lane group
constexpr int kSegPerGroup = 1;      // chunks a lane group walks through (strided)

template <int LPR, int VPL, int UNR, int SEGS, int MINB, bool ROUTED>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32, MINB)
segment_reduce_kernel(const float* __restrict__ src, int64_t src_ld, int32_t src_row_mul,
                      int64_t bound0, int64_t bound1, const int32_t* __restrict__ row_slot,
                      const float* __restrict__ init, int64_t init_ld,
                      const float* __restrict__ src_scale,
                      const float* __restrict__ row_scale, const int32_t* __restrict__ col,
                      int64_t n_seg, int64_t n_groups, const int4* __restrict__ seg,
                      float* __restrict__ partial, float* __restrict__ out, int64_t out_ld, int dim,
                      int accumulate, const __grid_constant__ OutRoute route) {
    constexpr int G = 32 / LPR;
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR;                       // lane within the group == float4 column
    const int gbase = lane - gl;                     // first lane of the group (shuffle source base)
    const int64_t group = ((int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5)) * G + lane / LPR;
    const int nvec = dim >> 2;

    // chunk sequence of this group: group, group + n_groups, ...
    int4 cur = make_int4(0, 0, 0, -1), nxt = make_int4(0, 0, 0, -1);
    int64_t s = group;
    if (s < n_seg) cur = __ldg(seg + s);
    if (s + n_groups < n_seg) nxt = __ldg(seg + s + n_groups);
    int colv = (s < n_seg && cur.x + gl < cur.y) ? __ldg(col + cur.x + gl) : 0;

    for (int i = 0; i < SEGS; ++i, s += n_groups) {
        const bool live = s < n_seg;                 // not warp-uniform: keep shuffles unconditional
        // prefetch: plan record two chunks ahead, first column batch of the next chunk
        int4 nxt2 = make_int4(0, 0, 0, -1);
        if (s + 2 * n_groups < n_seg && i + 2 < SEGS) nxt2 = __ldg(seg + s + 2 * n_groups);
        const bool nlive = (s + n_groups < n_seg) && (i + 1 < SEGS);
        int ncolv = (nlive && nxt.x + gl < nxt.y) ? __ldg(col + nxt.x + gl) : 0;

        const int begin = cur.x, end = live ? cur.y : cur.x, row = cur.z, part = cur.w;
        const int slot = row_slot ? (live ? __ldg(row_slot + row) : 0) : (row >= bound0) + (row >= bound1);
        float4 acc[VPL];
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = gl + w * LPR;
            // an optional initial row (e.g. the rank's own partial sum) opens the ordered sum
            acc[w] = (init && !accumulate && live && part < 0 && cv < nvec) ? ldg4(init + (int64_t)row * init_ld + 4 * cv) : f4_zero();
        }
        // longest chunk among the groups of this warp drives the (warp-uniform) trip count
        int len = end - begin;
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(kFull, len, o));
        for (int j0 = 0; j0 < len; j0 += LPR) {
            if (j0 > 0) colv = (begin + j0 + gl < end) ? __ldg(col + begin + j0 + gl) : 0;
            const int cnt = min(LPR, end - begin - j0);          // may be <= 0 for a shorter group
            for (int k = 0; k < LPR; k += UNR) {
                // warp-uniform early exit: every group is past its count
                if (__all_sync(kFull, k >= cnt)) break;
                float4 v[UNR][VPL];
                float sc[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = k + u;
                    const int e = __shfl_sync(kFull, colv, gbase + (idx % LPR));
                    const bool ok = idx < cnt;
                    const int64_t sr = (int64_t)e * src_row_mul + slot;
                    sc[u] = (ok && src_scale) ? __ldg(src_scale + e) : 1.0f;
#pragma unroll
                    for (int w = 0; w < VPL; ++w) {
                        const int cv = gl + w * LPR;
                        v[u][w] = (ok && cv < nvec) ? ldg4(src + sr * src_ld + 4 * cv) : f4_zero();
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u)
#pragma unroll
                    for (int w = 0; w < VPL; ++w) {
                        if (src_scale) f4_fma(acc[w], sc[u], v[u][w]);
                        else f4_add(acc[w], v[u][w]);
                    }
            }
        }
        if (live) {
            if (part < 0) {
                const float rs = row_scale ? __ldg(row_scale + row) : 1.0f;
                if (!accumulate) {
                    float* dst = out_row<ROUTED>(route, out, out_ld, row);
#pragma unroll
                    for (int w = 0; w < VPL; ++w) {
                        const int cv = gl + w * LPR;
                        if (cv < nvec) stg4(dst + 4 * cv, f4_scale(rs, acc[w]));
                    }
                } else if (end > begin) {
                    // accumulate mode: out[row] = init[row] + rs * sum; rows without incidences stay untouched
                    // (init may alias out: plain loads, not the read-only path)
#pragma unroll
                    for (int w = 0; w < VPL; ++w) {
                        const int cv = gl + w * LPR;
                        if (cv >= nvec) continue;
                        float4 o = *reinterpret_cast<const float4*>(init + (int64_t)row * init_ld + 4 * cv);
                        f4_fma(o, rs, acc[w]);
                        stg4(out + (int64_t)row * out_ld + 4 * cv, o);
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = gl + w * LPR;
                    if (cv < nvec) stg4(partial + (int64_t)part * dim + 4 * cv, acc[w]);
                }
            }
        }
        cur = nxt;
        nxt = nxt2;
        colv = ncolv;
    }
}


// One block per split row: its 8 x 32/LPR lane groups take interleaved partial rows (the
// compiler keeps ~4 loads in flight per lane, so parallelism comes from the groups), a fixed
// shuffle tree combines the groups of a warp and warp 0 adds the 8 warp sums in order:
//   out[row] = row_scale * sum of partial[p0 .. p1)        (fixed order => deterministic)
template <int LPR, int VPL, bool ROUTED>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32)
segment_fixup_kernel(const float* __restrict__ partial, const int32_t* __restrict__ split_row,
                     const int32_t* __restrict__ split_ptr, int64_t n_split,
                     const float* __restrict__ init, int64_t init_ld,
                     const float* __restrict__ row_scale, float* __restrict__ out, int64_t out_ld,
                     int dim, int accumulate, const __grid_constant__ OutRoute route) {
    constexpr int G = 32 / LPR;
    __shared__ float4 warp_sum[kSegWarpsPerBlock][LPR * VPL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPR, g = lane / LPR;
    const int64_t i = blockIdx.x;
    const int row = split_row[i];
    const int p0 = split_ptr[i], p1 = split_ptr[i + 1];
    const int nvec = dim >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) acc[w] = f4_zero();
    for (int p = p0 + warp * G + g; p < p1; p += kSegWarpsPerBlock * G)
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = gl + w * LPR;
            if (cv < nvec) f4_add(acc[w], ldg4(partial + (int64_t)p * dim + 4 * cv));
        }
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1)
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            acc[w].x += __shfl_xor_sync(0xffffffffu, acc[w].x, o);
            acc[w].y += __shfl_xor_sync(0xffffffffu, acc[w].y, o);
            acc[w].z += __shfl_xor_sync(0xffffffffu, acc[w].z, o);
            acc[w].w += __shfl_xor_sync(0xffffffffu, acc[w].w, o);
        }
    if (g == 0) {
#pragma unroll
        for (int w = 0; w < VPL; ++w) warp_sum[warp][gl + w * LPR] = acc[w];
    }
    __syncthreads();
    if (warp == 0 && g == 0) {
        const float rs = row_scale ? row_scale[row] : 1.0f;
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = gl + w * LPR;
            if (cv >= nvec) continue;
            if (!accumulate) {
                float4 t = init ? ldg4(init + (int64_t)row * init_ld + 4 * cv) : f4_zero();
#pragma unroll
                for (int k = 0; k < kSegWarpsPerBlock; ++k) f4_add(t, warp_sum[k][cv]);
                stg4(out_row<ROUTED>(route, out, out_ld, row) + 4 * cv, f4_scale(rs, t));
            } else {
                float4 t = f4_zero();
#pragma unroll
                for (int k = 0; k < kSegWarpsPerBlock; ++k) f4_add(t, warp_sum[k][cv]);
                float4 o = *reinterpret_cast<const float4*>(init + (int64_t)row * init_ld + 4 * cv);
                f4_fma(o, rs, t);
                stg4(out + (int64_t)row * out_ld + 4 * cv, o);
            }
        }
    }
}

template <int LPR, int VPL>
static int launch_segment_reduce(const ihg_csr* g, const float* src, int64_t src_ld, int32_t mul,
                                 int64_t b0, int64_t b1, const int32_t* row_slot, const float* init,
                                 int64_t init_ld, const float* src_scale,
                                 const float* row_scale, float* partial, float* out, int64_t out_ld,
                                 int dim, int accumulate, int l2_source, const OutRoute* route, cudaStream_t st) {
    constexpr int G = 32 / LPR;
    static const OutRoute kNoRoute = {};
    // (unroll, chunks per group) = (4, 1) measured best on both bench workloads (profiles/
    // microbench_segment.py): 5.3 TB/s at d=128 (82% of the measured copy bandwidth); at d=64 the
    // 256-byte random rows cap every variant near 3.1 TB/s (DRAM page locality, not the kernel).
    constexpr int SEGS = kSegPerGroup;
    const int64_t groups_per_block = (int64_t)kSegWarpsPerBlock * G;
    int64_t blocks = ceil_div(ceil_div(g->n_seg, SEGS), groups_per_block);
    if (blocks < 1) blocks = 1;
    const int64_t n_groups = blocks * groups_per_block;
    // unroll 4, no forced occupancy: forcing >= 4 blocks/SM wins in the isolated microbenchmark
    // but loses in the real step (1.14 vs 1.07 ms at amazon-full), so the in-situ winner stays.
    // Also tried and rejected (r01): staging a whole column batch (16 rows) per lane group with
    // cp.async into a per-thread shared-memory ring -- 4x the bytes in flight per group but 64 KB of
    // shared memory per block (3 blocks/SM) and a drain per batch: 1.29 vs 1.07 ms.  Numbering the
    // hyperedges by user (sequential reads for a third of the incidences) moved it by < 1 %.
    // Launch-bound variants re-measured in situ (profiles/r01_bench_segreduce_variants.txt): unlike the
    // L2-served two-hop gathers, these DRAM-served 256/512-byte random row reads get SLOWER with more
    // resident warps or a shorter unroll (unroll 2 with >= 5/6/8 blocks: 2.3 / 2.7 / 3.2 TB/s vs 3.6-4.0).
    // A source that is L2-resident (the L2-sized hyperedge ranges of the phased reduction) wants the
    // opposite: unroll 2 and >= 5 resident blocks, like the two-hop gathers.
    if (route)
        segment_reduce_kernel<LPR, VPL, kSegUnroll, SEGS, 0, true><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
            src, src_ld, mul, b0, b1, row_slot, init, init_ld, src_scale, row_scale, g->col, g->n_seg, n_groups,
            reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim, 0, *route);
    else if (l2_source)
        segment_reduce_kernel<LPR, VPL, 2, SEGS, 5, false><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
            src, src_ld, mul, b0, b1, row_slot, init, init_ld, src_scale, row_scale, g->col, g->n_seg, n_groups,
            reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim, accumulate, kNoRoute);
    else
        segment_reduce_kernel<LPR, VPL, kSegUnroll, SEGS, 0, false><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
            src, src_ld, mul, b0, b1, row_slot, init, init_ld, src_scale, row_scale, g->col, g->n_seg, n_groups,
            reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim, accumulate, kNoRoute);
    IHG_LAUNCH_CHECK();
    if (g->n_split > 0) {
        if (route)
            segment_fixup_kernel<LPR, VPL, true><<<(unsigned)g->n_split, kSegWarpsPerBlock * 32, 0, st>>>(
                partial, g->split_row, g->split_ptr, g->n_split, init, init_ld, row_scale, out, out_ld, dim, 0, *route);
        else
            segment_fixup_kernel<LPR, VPL, false><<<(unsigned)g->n_split, kSegWarpsPerBlock * 32, 0, st>>>(
                partial, g->split_row, g->split_ptr, g->n_split, init, init_ld, row_scale, out, out_ld, dim, accumulate, kNoRoute);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}


// =========================================================================================
// Two-hop reduce: the order-1 node -> hyperedge -> node round trip without the hyperedge
// intermediate.
//   out[r] = row_scale[r] * alpha * sum_{e contains r} sum_{n in e} node_scale[n] * src[n]
// which is  _ScatterMean(edge_gather_sum(src))  of an order-1 IHGNN layer
// (/root/reference/Models/CommonLayers.py:58-66 + GnnLayers.py:233-234 after hoisting the
// aggregation Linear to node level), HGCN's H De^-1 H^T product (GnnLayers.py:148-151) and both of
// their backward passes (H H^T is symmetric).  A row's own contribution is deg(r) * ns[r] * src[r];
// the two OTHER nodes of every incident hyperedge come from the precomputed neighbour list
// nbr[j] = (i3[col[j], slot+1 mod 3], i3[col[j], slot+2 mod 3]).  Instead of writing [E,dim] and
// reading it back three times, the kernel gathers 2 node rows per incidence -- a win whenever the
// node table ([N,dim]) is L2-resident while the hyperedge table ([E,dim]) is not.
// Same chunk plan / fixed summation order / fix-up kernel as segment_reduce.
// =========================================================================================
__global__ void two_hop_index_kernel(const int4* __restrict__ seg, int64_t n_seg,
                                     const int32_t* __restrict__ col, const int32_t* __restrict__ i3,
                                     int64_t bound0, int64_t bound1,
                                     const int32_t* __restrict__ row_slot, int2* __restrict__ nbr) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    const int4 c = __ldg(seg + s);
    const int row = c.z;
    const int slot = row_slot ? __ldg(row_slot + row) : (row >= bound0) + (row >= bound1);
    const int a = slot == 2 ? 0 : slot + 1, b = slot == 0 ? 2 : slot - 1;      // (slot+1)%3, (slot+2)%3
    for (int j = c.x + lane; j < c.y; j += 32) {
        const int64_t e = __ldg(col + j);
        nbr[j] = make_int2(__ldg(i3 + 3 * e + a), __ldg(i3 + 3 * e + b));
    }
}

template <int LPR, int VPL, int UNR, int MINB, bool ROUTED>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32, MINB)
two_hop_reduce_kernel(const float* __restrict__ src, int64_t src_ld,
                      const float* __restrict__ node_scale, float alpha, float own_per_inc, float own_const,
                      const float* __restrict__ row_scale, const int2* __restrict__ nbr,
                      const int32_t* __restrict__ rowptr,
                      int64_t n_seg, const int4* __restrict__ seg, float* __restrict__ partial,
                      float* __restrict__ out, int64_t out_ld, int dim, const __grid_constant__ OutRoute route) {
    constexpr int G = 32 / LPR;
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR;
    const int gbase = lane - gl;
    const int64_t group = ((int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5)) * G + lane / LPR;
    const int nvec = dim >> 2;
    const bool live = group < n_seg;                 // not warp-uniform: keep shuffles unconditional
    const int4 cur = live ? __ldg(seg + group) : make_int4(0, 0, 0, -1);
    const int begin = cur.x, end = cur.y, row = cur.z, part = cur.w;
    int2 nv = (begin + gl < end) ? __ldg(nbr + begin + gl) : make_int2(0, 0);

    // the row's own term, issued early so that its latency hides behind the gathers
    float4 own[VPL];
    float own_w = 0.0f;
    // own weight: own_per_inc per incidence of the chunk; own_const once per row (by the row's first chunk,
    // including rows without incidences -- a self connection does not need a hyperedge)
    const bool first_chunk = live && (part < 0 || begin == __ldg(rowptr + row));
    if (live) own_w = ((float)(end - begin) * own_per_inc + (first_chunk ? own_const : 0.0f)) *
                      (node_scale ? __ldg(node_scale + row) : 1.0f);
#pragma unroll
    for (int w = 0; w < VPL; ++w) {
        const int cv = gl + w * LPR;
        own[w] = (own_w != 0.0f && cv < nvec) ? ldg4(src + (int64_t)row * src_ld + 4 * cv) : f4_zero();
    }

    float4 acc[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) acc[w] = f4_zero();
    int len = end - begin;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(kFull, len, o));
    for (int j0 = 0; j0 < len; j0 += LPR) {
        // prefetch the next neighbour batch while this one is being gathered
        const int jn = begin + j0 + LPR + gl;
        const int2 nn = (jn < end) ? __ldg(nbr + jn) : make_int2(0, 0);
        const int cnt = min(LPR, end - begin - j0);              // may be <= 0 for a shorter group
        for (int k = 0; k < LPR; k += UNR) {
            if (__all_sync(kFull, k >= cnt)) break;
            float4 v[UNR][2][VPL];
            float sc[UNR][2];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = k + u;
                const int na = __shfl_sync(kFull, nv.x, gbase + (idx % LPR));
                const int nb = __shfl_sync(kFull, nv.y, gbase + (idx % LPR));
                const bool ok = idx < cnt;
                sc[u][0] = (ok && node_scale) ? __ldg(node_scale + na) : 1.0f;
                sc[u][1] = (ok && node_scale) ? __ldg(node_scale + nb) : 1.0f;
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = gl + w * LPR;
                    const bool okc = ok && cv < nvec;
                    v[u][0][w] = okc ? ldg4(src + (int64_t)na * src_ld + 4 * cv) : f4_zero();
                    v[u][1][w] = okc ? ldg4(src + (int64_t)nb * src_ld + 4 * cv) : f4_zero();
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    if (node_scale) {
                        f4_fma(acc[w], sc[u][0], v[u][0][w]);
                        f4_fma(acc[w], sc[u][1], v[u][1][w]);
                    } else {
                        f4_add(acc[w], v[u][0][w]);
                        f4_add(acc[w], v[u][1][w]);
                    }
                }
        }
        nv = nn;
    }
    if (live) {
        const float rs = (part < 0 && row_scale) ? alpha * __ldg(row_scale + row) : alpha;
        float* drow = part < 0 ? out_row<ROUTED>(route, out, out_ld, row) : partial + (int64_t)part * dim;
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = gl + w * LPR;
            if (cv >= nvec) continue;
            f4_fma(acc[w], own_w, own[w]);
            stg4(drow + 4 * cv, f4_scale(rs, acc[w]));
        }
    }
}

template <int LPR, int VPL>
static int launch_two_hop(const ihg_csr* g, const int32_t* nbr, const float* src, int64_t src_ld,
                          const float* node_scale, float alpha, float own_per_inc, float own_const,
                          const float* row_scale,
                          float* partial, float* out, int64_t out_ld, int dim, const OutRoute* route, cudaStream_t st) {
    constexpr int G = 32 / LPR;
    static const OutRoute kNoRoute = {};
    const int64_t groups_per_block = (int64_t)kSegWarpsPerBlock * G;
    int64_t blocks = ceil_div(g->n_seg, groups_per_block);
    if (blocks < 1) blocks = 1;
    // (unroll 2, >= 5 resident blocks = 48 registers) measured best in situ: occupancy beats per-thread
    // ILP for these L2-served gathers (per call 0.274 vs 0.327 ms for unroll 4 / 80 registers at the
    // amazon-full shape, 1.59 vs 1.99 ms at cikm; profiles/r01_bench_twohop_variants.txt)
    if (route)
        two_hop_reduce_kernel<LPR, VPL, 2, 5, true><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
            src, src_ld, node_scale, alpha, own_per_inc, own_const, row_scale, reinterpret_cast<const int2*>(nbr),
            g->rowptr, g->n_seg, reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim, *route);
    else
        two_hop_reduce_kernel<LPR, VPL, 2, 5, false><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
            src, src_ld, node_scale, alpha, own_per_inc, own_const, row_scale, reinterpret_cast<const int2*>(nbr),
            g->rowptr, g->n_seg, reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim, kNoRoute);
    IHG_LAUNCH_CHECK();
    if (g->n_split > 0) {
        if (route)
            segment_fixup_kernel<LPR, VPL, true><<<(unsigned)g->n_split, kSegWarpsPerBlock * 32, 0, st>>>(
                partial, g->split_row, g->split_ptr, g->n_split, nullptr, 0, row_scale, out, out_ld, dim, 0, *route);
        else
            segment_fixup_kernel<LPR, VPL, false><<<(unsigned)g->n_split, kSegWarpsPerBlock * 32, 0, st>>>(
                partial, g->split_row, g->split_ptr, g->n_split, nullptr, 0, row_scale, out, out_ld, dim, 0, kNoRoute);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}

}  // namespace ihg

using namespace ihg;

static int make_route(const char* what, const ihg_csr* g, const int64_t* start, void* const* base, int32_t n,
                      OutRoute* r) {
    IHG_REQUIRE(start && base && n >= 1 && n <= kMaxRoute, "%s: needs 1..%d output ranges", what, kMaxRoute);
    IHG_REQUIRE(start[0] == 0 && start[n] == g->n_rows, "%s: the output ranges must cover rows [0, n_rows)", what);
    *r = OutRoute{};
    r->n = n;
    for (int k = 0; k < n; ++k) {
        IHG_REQUIRE(start[k] <= start[k + 1], "%s: output ranges must ascend", what);
        IHG_REQUIRE(base[k] || start[k] == start[k + 1], "%s: null destination of a non-empty range", what);
        r->start[k] = (int32_t)start[k];
        r->base[k] = static_cast<float*>(base[k]);
    }
    for (int k = n; k <= kMaxRoute; ++k) r->start[k] = (int32_t)start[n];
    return IHG_OK;
}

static int segment_reduce_impl(const ihg_csr* g, const float* src, int64_t src_ld,
                               int32_t src_row_mul, int64_t bound0, int64_t bound1,
                               const int32_t* row_slot, const float* init, int64_t init_ld,
                               const float* src_scale, const float* row_scale, float* partial,
                               float* out, int64_t out_ld, int32_t dim, int32_t flags,
                               const OutRoute* route, void* stream) {
    IHG_REQUIRE(g && src && (out || route), "segment_reduce: null pointer");
    const int accumulate = (flags & IHG_SEG_ACCUMULATE) != 0, l2_source = (flags & IHG_SEG_L2_SOURCE) != 0;
    IHG_REQUIRE(!accumulate || init, "segment_reduce: accumulate mode needs init (usually == out)");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 256, "segment_reduce: dim=%d must be a multiple of 4, <= 256", dim);
    IHG_REQUIRE(src_ld % 4 == 0 && out_ld % 4 == 0 && src_ld >= dim && out_ld >= dim,
                "segment_reduce: leading dimensions must be multiples of 4 and >= dim");
    // a plan may list a subset of the rows (accumulate mode: the non-empty rows; multi-GPU: the halo rows or
    // the own rows of a local table): rows without work items are left untouched
    IHG_REQUIRE(g->n_rows > 0 && g->n_seg >= 0 && g->seg, "segment_reduce: incomplete csr plan");
    if (g->n_seg == 0) return IHG_OK;
    IHG_REQUIRE(g->nnz == 0 || g->col, "segment_reduce: null col");
    IHG_REQUIRE(g->n_split == 0 || (partial && g->split_row && g->split_ptr),
                "segment_reduce: split rows need the partial buffer");
    IHG_REQUIRE(src_row_mul >= 1, "segment_reduce: src_row_mul must be >= 1");
    IHG_REQUIRE(!init || (init_ld % 4 == 0 && init_ld >= dim), "segment_reduce: bad init leading dimension");
    cudaStream_t st = as_stream(stream);
    const int nvec = dim / 4;
#define IHG_SEG_CASE(L, V) \
    return launch_segment_reduce<L, V>(g, src, src_ld, src_row_mul, bound0, bound1, row_slot, init, init_ld, src_scale, row_scale, partial, out, out_ld, dim, accumulate, l2_source, route, st)
    if (nvec <= 1) IHG_SEG_CASE(1, 1);
    if (nvec <= 2) IHG_SEG_CASE(2, 1);
    if (nvec <= 4) IHG_SEG_CASE(4, 1);
    if (nvec <= 8) IHG_SEG_CASE(8, 1);
    if (nvec <= 16) IHG_SEG_CASE(16, 1);
    if (nvec <= 32) IHG_SEG_CASE(32, 1);
    IHG_SEG_CASE(32, 2);
#undef IHG_SEG_CASE
}

extern "C" int ihg_segment_reduce(const ihg_csr* g, const float* src, int64_t src_ld,
                                  int32_t src_row_mul, int64_t bound0, int64_t bound1,
                                  const int32_t* row_slot, const float* init, int64_t init_ld,
                                  const float* src_scale, const float* row_scale, float* partial,
                                  float* out, int64_t out_ld, int32_t dim, int32_t flags,
                                  void* stream) {
    return segment_reduce_impl(g, src, src_ld, src_row_mul, bound0, bound1, row_slot, init, init_ld, src_scale,
                               row_scale, partial, out, out_ld, dim, flags, nullptr, stream);
}

extern "C" int ihg_segment_reduce_routed(const ihg_csr* g, const float* src, int64_t src_ld,
                                         int32_t src_row_mul, int64_t bound0, int64_t bound1,
                                         const int32_t* row_slot, float* partial,
                                         const int64_t* route_start, void* const* route_base,
                                         int32_t n_route, int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(g, "segment_reduce_routed: null csr");
    OutRoute route;
    if (int rc = make_route("segment_reduce_routed", g, route_start, route_base, n_route, &route)) return rc;
    return segment_reduce_impl(g, src, src_ld, src_row_mul, bound0, bound1, row_slot, nullptr, 0, nullptr, nullptr,
                               partial, nullptr, out_ld, dim, 0, &route, stream);
}

extern "C" int ihg_two_hop_index_build(const ihg_csr* g, const int32_t* i3, int64_t bound0,
                                       int64_t bound1, const int32_t* row_slot, int32_t* nbr,
                                       void* stream) {
    IHG_REQUIRE(g && i3 && nbr, "two_hop_index_build: null pointer");
    IHG_REQUIRE(g->n_rows > 0 && g->n_seg >= g->n_rows && g->seg, "two_hop_index_build: incomplete csr plan");
    if (g->nnz == 0) return IHG_OK;
    IHG_REQUIRE(g->col, "two_hop_index_build: null col");
    const int warps = 8;
    two_hop_index_kernel<<<(unsigned)ceil_div(g->n_seg, warps), warps * 32, 0, as_stream(stream)>>>(
        reinterpret_cast<const int4*>(g->seg), g->n_seg, g->col, i3, bound0, bound1, row_slot,
        reinterpret_cast<int2*>(nbr));
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

static int two_hop_reduce_impl(const ihg_csr* g, const int32_t* nbr, const float* src,
                               int64_t src_ld, const float* node_scale, float alpha,
                               float own_per_incidence, float own_const,
                               const float* row_scale, float* partial, float* out,
                               int64_t out_ld, int32_t dim, const OutRoute* route, void* stream) {
    IHG_REQUIRE(g && src && (out || route), "two_hop_reduce: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 256, "two_hop_reduce: dim=%d must be a multiple of 4, <= 256", dim);
    IHG_REQUIRE(src_ld % 4 == 0 && out_ld % 4 == 0 && src_ld >= dim && out_ld >= dim,
                "two_hop_reduce: leading dimensions must be multiples of 4 and >= dim");
    IHG_REQUIRE(g->n_rows > 0 && g->n_seg >= 0 && g->seg && g->rowptr, "two_hop_reduce: incomplete csr plan");
    if (g->n_seg == 0) return IHG_OK;                 // a row-range view may be empty
    IHG_REQUIRE(g->nnz == 0 || nbr, "two_hop_reduce: null neighbour list");
    IHG_REQUIRE(g->n_split == 0 || (partial && g->split_row && g->split_ptr),
                "two_hop_reduce: split rows need the partial buffer");
    cudaStream_t st = as_stream(stream);
    const int nvec = dim / 4;
#define IHG_TH_CASE(L, V) \
    return launch_two_hop<L, V>(g, nbr, src, src_ld, node_scale, alpha, own_per_incidence, own_const, row_scale, partial, out, out_ld, dim, route, st)
    if (nvec <= 1) IHG_TH_CASE(1, 1);
    if (nvec <= 2) IHG_TH_CASE(2, 1);
    if (nvec <= 4) IHG_TH_CASE(4, 1);
    if (nvec <= 8) IHG_TH_CASE(8, 1);
    if (nvec <= 16) IHG_TH_CASE(16, 1);
    if (nvec <= 32) IHG_TH_CASE(32, 1);
    IHG_TH_CASE(32, 2);
#undef IHG_TH_CASE
}

extern "C" int ihg_two_hop_reduce(const ihg_csr* g, const int32_t* nbr, const float* src,
                                  int64_t src_ld, const float* node_scale, float alpha,
                                  float own_per_incidence, float own_const,
                                  const float* row_scale, float* partial, float* out,
                                  int64_t out_ld, int32_t dim, void* stream) {
    return two_hop_reduce_impl(g, nbr, src, src_ld, node_scale, alpha, own_per_incidence, own_const, row_scale,
                               partial, out, out_ld, dim, nullptr, stream);
}

extern "C" int ihg_two_hop_reduce_routed(const ihg_csr* g, const int32_t* nbr, const float* src,
                                         int64_t src_ld, const float* node_scale, float alpha,
                                         float own_per_incidence, float own_const, float* partial,
                                         const int64_t* route_start, void* const* route_base,
                                         int32_t n_route, int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(g, "two_hop_reduce_routed: null csr");
    OutRoute route;
    if (int rc = make_route("two_hop_reduce_routed", g, route_start, route_base, n_route, &route)) return rc;
    return two_hop_reduce_impl(g, nbr, src, src_ld, node_scale, alpha, own_per_incidence, own_const, nullptr,
                               partial, nullptr, out_ld, dim, &route, stream);
}
