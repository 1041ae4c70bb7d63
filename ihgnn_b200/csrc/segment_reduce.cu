// K-B: deterministic segmented reduction over CSR rows (the edge->node scatter-mean and all
// of its transposes).  Replaces torch_sparse.matmul(incidence, Ef) * Dv^-1
// (/root/reference/Models/GnnLayers.py:233-234), the index_put_(accumulate=True) backward of
// the three row gathers (CommonLayers.py:70-72) and EmbeddingBag(mean) (EmbeddingLayers.py:79).
//
// Work decomposition: the plan (ihg_segment_plan_build) lists row chunks of <= chunk_len
// incidences as 16-byte records {begin, end, row, partial slot}.  A group of LPR lanes (LPR x
// float4 spans the feature dimension; 32/LPR groups per warp) advances a quad of consecutive
// chunks in lockstep -- 4 accumulation streams x 4 independent 128-bit row gathers in flight --
// and walks a strided sequence of quads; the plan records and first column indices of the next
// quad are prefetched while the current rows are in flight.  Rows longer than chunk_len (Zipf
// head) are split; their partial sums are combined by a second kernel (one warp per split row,
// fixed interleave + shuffle tree).
// Summation order is a pure function of the plan => bitwise run-to-run determinism, no float
// atomics.
//
// Roofline: HBM.  Algorithmic bytes per incidence: 4 (col) + 4*dim (source row); per row:
// 16 (plan) + 4 (scale) + 4*dim (output row).
#include "common.cuh"

namespace ihg {

constexpr int kSegWarpsPerBlock = 8;
constexpr int kSegUnroll = 4;        // gathers in flight per chunk stream
constexpr int kQuadsPerGroup = 2;    // quads a lane group walks through (strided)
constexpr int kFixUnroll = 16;       // partial rows in flight per lane group in the fix-up

// A lane group advances Q consecutive chunks ("quad") in lockstep: Q independent accumulation
// streams x kSegUnroll gathers each are in flight, so short rows (degree < unroll) still fill
// the memory pipeline.
template <int LPR, int VPL, int Q>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32, 2)
segment_reduce_kernel(const float* __restrict__ src, int64_t src_ld, int32_t src_row_mul,
                      int64_t bound0, int64_t bound1, const float* __restrict__ src_scale,
                      const float* __restrict__ row_scale, const int32_t* __restrict__ col,
                      int64_t n_seg, int64_t n_groups, const int4* __restrict__ seg,
                      float* __restrict__ partial, float* __restrict__ out, int64_t out_ld, int dim) {
    constexpr int G = 32 / LPR;
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR;                       // lane within the group == float4 column
    const int gbase = lane - gl;                     // first lane of the group (shuffle source base)
    const int64_t group = ((int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5)) * G + lane / LPR;
    const int nvec = dim >> 2;
    const int4 kDead = make_int4(0, 0, 0, -1);

    int4 cur[Q], nxt[Q];
    int colv[Q];
    int64_t qd = group;                              // quad index; chunks qd*Q .. qd*Q+Q-1
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int64_t s = qd * Q + j;
        cur[j] = s < n_seg ? __ldg(seg + s) : kDead;
        colv[j] = (cur[j].x + gl < cur[j].y) ? __ldg(col + cur[j].x + gl) : 0;
    }

    for (int i = 0; i < kQuadsPerGroup; ++i, qd += n_groups) {
        // prefetch the plan records of the next quad
        const bool more = i + 1 < kQuadsPerGroup;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const int64_t s = (qd + n_groups) * Q + j;
            nxt[j] = (more && s < n_seg) ? __ldg(seg + s) : kDead;
        }
        float4 acc[Q][VPL];
        int len = 0;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            len = max(len, cur[j].y - cur[j].x);
#pragma unroll
            for (int w = 0; w < VPL; ++w) acc[j][w] = f4_zero();
        }
        // longest chunk among all streams of this warp drives the (warp-uniform) trip count
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(kFull, len, o));
        for (int j0 = 0; j0 < len; j0 += LPR) {
            if (j0 > 0) {
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    colv[j] = (cur[j].x + j0 + gl < cur[j].y) ? __ldg(col + cur[j].x + j0 + gl) : 0;
            }
            int cnt[Q], cmax = 0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                cnt[j] = min(LPR, cur[j].y - cur[j].x - j0);     // <= 0 for a finished stream
                cmax = max(cmax, cnt[j]);
            }
            for (int k = 0; k < LPR; k += kSegUnroll) {
                if (__all_sync(kFull, k >= cmax)) break;         // warp-uniform early exit
                float4 v[Q][kSegUnroll][VPL];
                float sc[Q][kSegUnroll];
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const int slot = (cur[j].z >= bound0) + (cur[j].z >= bound1);
#pragma unroll
                    for (int u = 0; u < kSegUnroll; ++u) {
                        const int idx = k + u;
                        const int e = __shfl_sync(kFull, colv[j], gbase + (idx % LPR));
                        const bool ok = idx < cnt[j];
                        const int64_t sr = (int64_t)e * src_row_mul + slot;
                        sc[j][u] = (ok && src_scale) ? __ldg(src_scale + e) : 1.0f;
#pragma unroll
                        for (int w = 0; w < VPL; ++w) {
                            const int cv = gl + w * LPR;
                            v[j][u][w] = (ok && cv < nvec) ? ldg4(src + sr * src_ld + 4 * cv) : f4_zero();
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < Q; ++j)
#pragma unroll
                    for (int u = 0; u < kSegUnroll; ++u)
#pragma unroll
                        for (int w = 0; w < VPL; ++w) {
                            if (src_scale) f4_fma(acc[j][w], sc[j][u], v[j][u][w]);
                            else f4_add(acc[j][w], v[j][u][w]);
                        }
            }
        }
        // first column batch of the next quad (its plan records have arrived by now)
        int ncolv[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) ncolv[j] = (nxt[j].x + gl < nxt[j].y) ? __ldg(col + nxt[j].x + gl) : 0;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const int row = cur[j].z, part = cur[j].w;
            const bool live = qd * Q + j < n_seg;
            if (!live) continue;
            if (part < 0) {
                const float rs = row_scale ? __ldg(row_scale + row) : 1.0f;
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = gl + w * LPR;
                    if (cv < nvec) stg4(out + (int64_t)row * out_ld + 4 * cv, f4_scale(rs, acc[j][w]));
                }
            } else {
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = gl + w * LPR;
                    if (cv < nvec) stg4(partial + (int64_t)part * dim + 4 * cv, acc[j][w]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            cur[j] = nxt[j];
            colv[j] = ncolv[j];
        }
    }
}

// One warp per split row: its 32/LPR lane groups take alternating partial rows (kFixUnroll loads
// in flight each), a fixed shuffle tree combines the groups:
//   out[row] = row_scale * sum of partial[p0 .. p1)        (fixed order => deterministic)
template <int LPR, int VPL>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32)
segment_fixup_kernel(const float* __restrict__ partial, const int32_t* __restrict__ split_row,
                     const int32_t* __restrict__ split_ptr, int64_t n_split,
                     const float* __restrict__ row_scale, float* __restrict__ out, int64_t out_ld,
                     int dim) {
    constexpr int G = 32 / LPR;
    constexpr int U = VPL > 1 ? kFixUnroll / 2 : kFixUnroll;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR, g = lane / LPR;
    const int64_t i = (int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5);
    if (i >= n_split) return;                        // warp-uniform
    const int row = split_row[i];
    const int p0 = split_ptr[i], p1 = split_ptr[i + 1];
    const int nvec = dim >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) acc[w] = f4_zero();
    for (int p = p0 + g; p < p1; p += G * U) {
        float4 v[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
                const int cv = gl + w * LPR;
                const int pp = p + u * G;
                v[u][w] = (pp < p1 && cv < nvec) ? ldg4(partial + (int64_t)pp * dim + 4 * cv) : f4_zero();
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int w = 0; w < VPL; ++w) f4_add(acc[w], v[u][w]);
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
        const float rs = row_scale ? row_scale[row] : 1.0f;
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = gl + w * LPR;
            if (cv < nvec) stg4(out + (int64_t)row * out_ld + 4 * cv, f4_scale(rs, acc[w]));
        }
    }
}

template <int LPR, int VPL>
static int launch_segment_reduce(const ihg_csr* g, const float* src, int64_t src_ld, int32_t mul,
                                 int64_t b0, int64_t b1, const float* src_scale,
                                 const float* row_scale, float* partial, float* out, int64_t out_ld,
                                 int dim, cudaStream_t st) {
    constexpr int G = 32 / LPR;
    constexpr int Q = VPL > 1 ? 2 : 4;
    const int64_t groups_per_block = (int64_t)kSegWarpsPerBlock * G;
    const int64_t n_quads = ceil_div(g->n_seg, Q);
    int64_t blocks = ceil_div(ceil_div(n_quads, kQuadsPerGroup), groups_per_block);
    if (blocks < 1) blocks = 1;
    const int64_t n_groups = blocks * groups_per_block;        // stride of the quad sequences
    segment_reduce_kernel<LPR, VPL, Q><<<(unsigned)blocks, kSegWarpsPerBlock * 32, 0, st>>>(
        src, src_ld, mul, b0, b1, src_scale, row_scale, g->col, g->n_seg, n_groups,
        reinterpret_cast<const int4*>(g->seg), partial, out, out_ld, dim);
    IHG_LAUNCH_CHECK();
    if (g->n_split > 0) {
        const unsigned fb = (unsigned)ceil_div(g->n_split, kSegWarpsPerBlock);
        segment_fixup_kernel<LPR, VPL><<<fb, kSegWarpsPerBlock * 32, 0, st>>>(
            partial, g->split_row, g->split_ptr, g->n_split, row_scale, out, out_ld, dim);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}

}  // namespace ihg

using namespace ihg;

extern "C" int ihg_segment_reduce(const ihg_csr* g, const float* src, int64_t src_ld,
                                  int32_t src_row_mul, int64_t bound0, int64_t bound1,
                                  const float* src_scale, const float* row_scale, float* partial,
                                  float* out, int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(g && src && out, "segment_reduce: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 256, "segment_reduce: dim=%d must be a multiple of 4, <= 256", dim);
    IHG_REQUIRE(src_ld % 4 == 0 && out_ld % 4 == 0 && src_ld >= dim && out_ld >= dim,
                "segment_reduce: leading dimensions must be multiples of 4 and >= dim");
    IHG_REQUIRE(g->n_rows > 0 && g->n_seg >= g->n_rows && g->seg, "segment_reduce: incomplete csr plan");
    IHG_REQUIRE(g->nnz == 0 || g->col, "segment_reduce: null col");
    IHG_REQUIRE(g->n_split == 0 || (partial && g->split_row && g->split_ptr),
                "segment_reduce: split rows need the partial buffer");
    IHG_REQUIRE(src_row_mul >= 1, "segment_reduce: src_row_mul must be >= 1");
    cudaStream_t st = as_stream(stream);
    const int nvec = dim / 4;
#define IHG_SEG_CASE(L, V) \
    return launch_segment_reduce<L, V>(g, src, src_ld, src_row_mul, bound0, bound1, src_scale, row_scale, partial, out, out_ld, dim, st)
    if (nvec <= 1) IHG_SEG_CASE(1, 1);
    if (nvec <= 2) IHG_SEG_CASE(2, 1);
    if (nvec <= 4) IHG_SEG_CASE(4, 1);
    if (nvec <= 8) IHG_SEG_CASE(8, 1);
    if (nvec <= 16) IHG_SEG_CASE(16, 1);
    if (nvec <= 32) IHG_SEG_CASE(32, 1);
    IHG_SEG_CASE(32, 2);
#undef IHG_SEG_CASE
}
