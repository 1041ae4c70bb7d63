// K-E: batched inference ranking -- fused HEM scoring of a candidate list + per-query top-k.
//
// Replaces, per (user, query) of the evaluation loop:
//   TestSearchLogDataLoader.__iter__   /root/reference/Dataset.py:324-329  (index vectors of length I)
//   RawGnn.forward, eval branch        /root/reference/Models/RawGnn.py:124-142 (gathers the SAME
//                                       user / query row I times, scores every item)
//   HemPredictionLayer.forward         /root/reference/Models/PredictionLayers.py:21-44
//   torch.sort(outputs, descending)[:10] + .cpu()   /root/reference/Helpers/Metrics.py:60-61
// by ONE launch for a whole batch of queries: one CTA per query builds m = lambda*q + (1-lambda)*u
// once (registers), its 8 warps stream the candidate item rows (128-bit gathers, 4 rows in flight
// per warp), and the block keeps a running top-k in shared memory (k rounds of block arg-max per
// chunk of 1024 candidates; ties go to the earlier candidate).  Scores use the same per-lane
// accumulation order + shuffle tree as hem_score_fwd_kernel, so they are bit-identical to the
// training-time scorer.
//
// Roofline: HBM / L2 gather bandwidth.  Algorithmic bytes per query:
//   8*D (u, q rows) + C * (8 [candidate id] + 4*D [item row] + 4 [bias]) + k * 12.
#include <math_constants.h>

#include "common.cuh"

namespace ihg {

constexpr int kRankWarps = 8;
constexpr int kRankChunk = 1024;      // candidates scored between two top-k merges
constexpr int kRankMaxK = 32;
constexpr int kRankUnroll = 4;

__device__ __forceinline__ float rank_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (value desc, position asc) ordering
__device__ __forceinline__ bool rank_better(float v, int p, float bv, int bp) {
    return v > bv || (v == bv && p < bp);
}

template <int VPL>
__global__ void __launch_bounds__(kRankWarps * 32)
rank_topk_kernel(const float* __restrict__ feat, int64_t feat_ld, const int64_t* __restrict__ users,
                 const int64_t* __restrict__ queries, int64_t query_row0,
                 const int64_t* __restrict__ cand, int64_t n_cand, int64_t item_row0,
                 int64_t item_count, const float* __restrict__ items_bias, float lam, int nvec, int k,
                 int cosine, int64_t* __restrict__ top_items, float* __restrict__ top_scores) {
    // positions [0, kRankMaxK): the running top-k; [kRankMaxK, kRankMaxK + kRankChunk): this chunk
    __shared__ float s_val[kRankMaxK + kRankChunk];
    __shared__ int64_t s_id[kRankMaxK + kRankChunk];
    __shared__ float s_wv[kRankWarps];
    __shared__ int s_wp[kRankWarps];
    __shared__ float s_newv[kRankMaxK];
    __shared__ int64_t s_newid[kRankMaxK];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    const float oml = 1.0f - lam;

    // m = lambda * q + (1 - lambda) * u  (PredictionLayers.py:28-31; users == null: m = q)
    float4 m[VPL];
    {
        const float* qrow = feat + (__ldg(queries + b) + query_row0) * feat_ld;
        const float* urow = users ? feat + __ldg(users + b) * feat_ld : nullptr;
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int c = lane + 32 * w;
            m[w] = f4_zero();
            if (c < nvec) {
                const float4 q = ldg4(qrow + 4 * c);
                m[w] = q;
                if (urow) {
                    const float4 u = ldg4(urow + 4 * c);
                    m[w] = make_float4(lam * q.x + oml * u.x, lam * q.y + oml * u.y,
                                       lam * q.z + oml * u.z, lam * q.w + oml * u.w);
                }
            }
        }
    }
    float norm_m = 1.0f;                             // max(|m|, eps) for the cosine scorer
    if (cosine) {
        float nm2 = 0.f;
#pragma unroll
        for (int w = 0; w < VPL; ++w) nm2 += m[w].x * m[w].x + m[w].y * m[w].y + m[w].z * m[w].z + m[w].w * m[w].w;
        norm_m = fmaxf(sqrtf(rank_warp_sum(nm2)), 1e-8f);
    }
    for (int i = tid; i < kRankMaxK; i += blockDim.x) {
        s_val[i] = -CUDART_INF_F;
        s_id[i] = -1;
    }
    __syncthreads();

    for (int64_t c0 = 0; c0 < n_cand; c0 += kRankChunk) {
        const int n = (int)min((int64_t)kRankChunk, n_cand - c0);
        // ---- score this chunk: warp per candidate, kRankUnroll rows in flight -------------------
        for (int j0 = warp * kRankUnroll; j0 < n; j0 += kRankWarps * kRankUnroll) {
            int64_t id[kRankUnroll];
            float4 v[kRankUnroll][VPL];
#pragma unroll
            for (int u = 0; u < kRankUnroll; ++u) {
                const int j = j0 + u;
                id[u] = -1;
                if (j < n) id[u] = cand ? __ldg(cand + b * n_cand + c0 + j) : c0 + j;
                const bool ok = id[u] >= 0 && id[u] < item_count;      // out-of-range candidates never rank
                if (!ok) id[u] = -1;
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int c = lane + 32 * w;
                    v[u][w] = (ok && c < nvec) ? ldg4(feat + (id[u] + item_row0) * feat_ld + 4 * c) : f4_zero();
                }
            }
#pragma unroll
            for (int u = 0; u < kRankUnroll; ++u) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < VPL; ++w)
                    s += v[u][w].x * m[w].x + v[u][w].y * m[w].y + v[u][w].z * m[w].z + v[u][w].w * m[w].w;
                s = rank_warp_sum(s);
                if (cosine) {                        // PredictionLayers.py:39: cosine_similarity(item, m)
                    float ni2 = 0.f;
#pragma unroll
                    for (int w = 0; w < VPL; ++w)
                        ni2 += v[u][w].x * v[u][w].x + v[u][w].y * v[u][w].y + v[u][w].z * v[u][w].z + v[u][w].w * v[u][w].w;
                    s = s / (fmaxf(sqrtf(rank_warp_sum(ni2)), 1e-8f) * norm_m);      // same expression as hem_score_fwd_kernel
                }
                const int j = j0 + u;
                if (lane == 0 && j < n) {
                    s_val[kRankMaxK + j] = id[u] >= 0 ? s + __ldg(items_bias + id[u]) : -CUDART_INF_F;
                    s_id[kRankMaxK + j] = id[u];
                }
            }
        }
        __syncthreads();
        // ---- merge: k rounds of block arg-max over [running top-k | chunk] --------------------
        const int total = kRankMaxK + n;
        for (int r = 0; r < k; ++r) {
            float bv = -CUDART_INF_F;
            int bp = 0x7fffffff;
            for (int p = tid; p < total; p += blockDim.x) {
                const float x = s_val[p];
                if (s_id[p] >= 0 && rank_better(x, p, bv, bp)) { bv = x; bp = p; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (op != 0x7fffffff && (bp == 0x7fffffff || rank_better(ov, op, bv, bp))) { bv = ov; bp = op; }
            }
            if (lane == 0) { s_wv[warp] = bv; s_wp[warp] = bp; }
            __syncthreads();
            if (tid == 0) {
                float fv = s_wv[0];
                int fp = s_wp[0];
                for (int w = 1; w < kRankWarps; ++w) {
                    const float ov = s_wv[w];
                    const int op = s_wp[w];
                    if (op != 0x7fffffff && (fp == 0x7fffffff || rank_better(ov, op, fv, fp))) { fv = ov; fp = op; }
                }
                if (fp != 0x7fffffff) {
                    s_newv[r] = fv;
                    s_newid[r] = s_id[fp];
                    s_id[fp] = -1;                   // taken
                } else {
                    s_newv[r] = -CUDART_INF_F;
                    s_newid[r] = -1;
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < kRankMaxK; i += blockDim.x) {
            s_val[i] = i < k ? s_newv[i] : -CUDART_INF_F;
            s_id[i] = i < k ? s_newid[i] : -1;
        }
        __syncthreads();
    }
    for (int i = tid; i < k; i += blockDim.x) {
        top_items[b * k + i] = s_id[i];
        top_scores[b * k + i] = s_val[i];
    }
}

}  // namespace ihg

using namespace ihg;

extern "C" int ihg_rank_topk(const float* feat, int64_t feat_ld, const int64_t* users,
                             const int64_t* queries, int64_t n_queries, int64_t query_row0,
                             const int64_t* cand, int64_t n_cand, int64_t item_row0,
                             int64_t item_count, const float* items_bias, float lambda_muq,
                             int32_t dim, int32_t k, int32_t cosine, int64_t* top_items,
                             float* top_scores, void* stream) {
    IHG_REQUIRE(feat && queries && items_bias && top_items && top_scores, "rank_topk: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 1024 && feat_ld % 4 == 0 && feat_ld >= dim,
                "rank_topk: dim=%d must be a multiple of 4, <= 1024, leading dimension a multiple of 4", dim);
    IHG_REQUIRE(k >= 1 && k <= kRankMaxK, "rank_topk: k=%d must be in [1, %d]", k, kRankMaxK);
    IHG_REQUIRE(n_cand >= 0 && item_count >= 0, "rank_topk: negative count");
    if (n_queries <= 0) return IHG_OK;
    cudaStream_t st = as_stream(stream);
    const int nvec = dim / 4;
    const unsigned grid = (unsigned)n_queries;
#define IHG_RANK_CASE(V)                                                                              \
    rank_topk_kernel<V><<<grid, kRankWarps * 32, 0, st>>>(feat, feat_ld, users, queries, query_row0, \
        cand, n_cand, item_row0, item_count, items_bias, lambda_muq, nvec, k, cosine, top_items, top_scores)
    if (nvec <= 32) IHG_RANK_CASE(1);
    else if (nvec <= 64) IHG_RANK_CASE(2);
    else if (nvec <= 128) IHG_RANK_CASE(4);
    else IHG_RANK_CASE(8);
#undef IHG_RANK_CASE
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}
