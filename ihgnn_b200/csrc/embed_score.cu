// K-C / K-D: embedding row movement and the fused HEM scorer.
//
//   ihg_copy_rows / ihg_gather_rows / ihg_scatter_add_rows
//        EmbeddingLayer.embed_user / embed_item (/root/reference/Models/EmbeddingLayers.py:70-74)
//        and RawGnn's batch row selects (Models/RawGnn.py:128-133) with their backward.
//        (EmbeddingBag(mean), EmbeddingLayers.py:79, is ihg_segment_reduce over the bag CSR.)
//   ihg_hem_score_fwd / bwd
//        HemPredictionLayer.forward (Models/PredictionLayers.py:21-44): replaces ~8 elementwise
//        launches (2 mul, add, mul, sum, index, add) by one warp-per-row kernel.
//
// All HBM/latency-bound: 128-bit accesses, one warp (or lane group) per row, reductions by
// fixed shuffle trees and ordered loops -- no float atomics.
#include "common.cuh"

namespace ihg {

__global__ void __launch_bounds__(256)
copy_rows_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst,
                 int64_t dst_ld, int64_t n_rows, int nvec) {
    const int64_t total = n_rows * nvec;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / nvec;
        const int c = (int)(idx % nvec);
        stg4(dst + r * dst_ld + 4 * c, ldg4_stream(src + r * src_ld + 4 * c));
    }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, int64_t table_ld,
                   const int64_t* __restrict__ idx, int64_t idx_offset, int64_t count,
                   float* __restrict__ out, int64_t out_ld, int nvec) {
    const int64_t total = count * nvec;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / nvec;
        const int c = (int)(t % nvec);
        const int64_t row = __ldg(idx + b) + idx_offset;
        stg4(out + b * out_ld + 4 * c, ldg4(table + row * table_ld + 4 * c));
    }
}

// Segmented row copy with one (source base, destination base) pair per segment: the halo
// exchange over NVLink peer memory.  Flat row r belongs to segment s with seg_off[s] <= r <
// seg_off[s+1]; it is read from seg_src[s] at row (src_rows ? src_rows[r] : r - seg_off[s]) and
// written to seg_dst[s] at row r - seg_off[s].  Either side may be a peer-mapped pointer.
struct HaloSegs {
    const float* src[16];
    float* dst[16];
    int64_t off[17];
    int n;
};
__global__ void __launch_bounds__(256)
halo_copy_kernel(HaloSegs segs, const int64_t* __restrict__ src_rows, int64_t src_ld, int64_t dst_ld, int nvec) {
    const int64_t total = segs.off[segs.n] * nvec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int kU = 4;                           // independent 16-byte copies in flight per thread
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += kU * stride) {
        float4 v[kU];
        float* dst[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t t = t0 + u * stride;
            dst[u] = nullptr;
            if (t < total) {
                const int64_t r = t / nvec;
                const int c = (int)(t - r * nvec);
                int sg = 0;
                while (sg + 1 < segs.n && r >= segs.off[sg + 1]) ++sg;
                const int64_t k = r - segs.off[sg];
                const int64_t srow = src_rows ? __ldg(src_rows + r) : k;
                // peer memory is not read-only cached: plain loads
                v[u] = *reinterpret_cast<const float4*>(segs.src[sg] + srow * src_ld + 4 * c);
                dst[u] = segs.dst[sg] + k * dst_ld + 4 * c;
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (dst[u]) *reinterpret_cast<float4*>(dst[u]) = v[u];
    }
}

// One block per source row b.  The first occurrence of an index ("leader") sums every later
// duplicate in ascending b and adds the total to the destination row once => deterministic.
// Both scans over idx are block-parallel: windows of 128 positions are tested at once, the matches
// of a window are compacted IN ORDER (ballot + prefix) into shared memory, then added in that order.
constexpr int kScatterSmemIdx = 2048;          // batches up to this size keep the whole index list in shared memory

template <bool SMEM_IDX>
__global__ void __launch_bounds__(128)
scatter_add_rows_kernel(const float* __restrict__ g, int64_t g_ld, const int64_t* __restrict__ idx,
                        int64_t idx_offset, int64_t count, float* __restrict__ out,
                        int64_t out_ld, int nvec) {
    __shared__ int dup_before;
    __shared__ int warp_cnt[4];
    __shared__ int match[128];
    __shared__ int64_t s_idx[SMEM_IDX ? kScatterSmemIdx : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x;
    if (SMEM_IDX) {
        // one coalesced sweep: every later scan reads shared memory instead of chaining global loads
        for (int64_t j = tid; j < count; j += blockDim.x) s_idx[j] = __ldg(idx + j);
    }
    if (tid == 0) dup_before = 0;
    __syncthreads();
    auto index_at = [&](int64_t j) -> int64_t { return SMEM_IDX ? s_idx[j] : __ldg(idx + j); };
    const int64_t my = index_at(b);
    int found = 0;
    for (int64_t j = tid; j < b; j += blockDim.x) found |= (index_at(j) == my);
    if (found) dup_before = 1;   // benign race: every writer stores 1
    __syncthreads();
    if (dup_before) return;
    float4 acc[2];                                   // nvec <= 256 (dim <= 1024): two float4 per thread
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int c = tid + 128 * w;
        acc[w] = c < nvec ? ldg4(g + b * g_ld + 4 * c) : f4_zero();
    }
    for (int64_t j0 = b + 1; j0 < count; j0 += 128) {
        const int64_t j = j0 + tid;
        const bool hit = j < count && index_at(j) == my;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (w < warp) base += warp_cnt[w];
            total += warp_cnt[w];
        }
        if (hit) match[base + __popc(bal & ((1u << lane) - 1u))] = (int)(j - j0);
        __syncthreads();
        for (int m = 0; m < total; ++m) {
            const int64_t jj = j0 + match[m];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int c = tid + 128 * w;
                if (c < nvec) f4_add(acc[w], ldg4(g + jj * g_ld + 4 * c));
            }
        }
        __syncthreads();                             // match[] / warp_cnt[] are rewritten by the next window
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int c = tid + 128 * w;
        if (c >= nvec) continue;
        float* o = out + (my + idx_offset) * out_ld + 4 * c;
        float4 base = *reinterpret_cast<const float4*>(o);
        f4_add(base, acc[w]);
        stg4(o, base);
    }
}

constexpr float kCosEps = 1e-8f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
hem_score_fwd_kernel(const float* __restrict__ uf, int64_t u_ld, const float* __restrict__ qf,
                     int64_t q_ld, const float* __restrict__ itf, int64_t i_ld,
                     const float* __restrict__ items_bias, const int64_t* __restrict__ item_idx,
                     float lam, int64_t count, int nvec, float* __restrict__ score, int cosine,
                     float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float oml = 1.0f - lam;
    for (int64_t b = warp; b < count; b += nwarps) {
        float s = 0.f, ni2 = 0.f, nm2 = 0.f;
        for (int c = lane; c < nvec; c += 32) {
            const float4 q = ldg4(qf + b * q_ld + 4 * c);
            const float4 it = ldg4(itf + b * i_ld + 4 * c);
            float4 m = q;
            if (uf) {
                const float4 u = ldg4(uf + b * u_ld + 4 * c);
                m = make_float4(lam * q.x + oml * u.x, lam * q.y + oml * u.y,
                                lam * q.z + oml * u.z, lam * q.w + oml * u.w);
            }
            s += it.x * m.x + it.y * m.y + it.z * m.z + it.w * m.w;
            if (cosine) {
                ni2 += it.x * it.x + it.y * it.y + it.z * it.z + it.w * it.w;
                nm2 += m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w;
            }
        }
        s = warp_sum(s);
        if (cosine) { ni2 = warp_sum(ni2); nm2 = warp_sum(nm2); }
        if (lane == 0) {
            const int64_t bi = item_idx ? __ldg(item_idx + b) : b;
            if (cosine) {
                // torch.cosine_similarity (PredictionLayers.py:39): x.y / (max(|x|, eps) max(|y|, eps)), eps = 1e-8
                const float a = fmaxf(sqrtf(ni2), kCosEps), bm = fmaxf(sqrtf(nm2), kCosEps);
                norms[3 * b] = s; norms[3 * b + 1] = a; norms[3 * b + 2] = bm;
                score[b] = s / (a * bm) + __ldg(items_bias + bi);
            } else {
                score[b] = s + __ldg(items_bias + bi);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
hem_score_bwd_kernel(const float* __restrict__ dscore, const float* __restrict__ uf, int64_t u_ld,
                     const float* __restrict__ qf, int64_t q_ld, const float* __restrict__ itf,
                     int64_t i_ld, float lam, int64_t count, int nvec, float* __restrict__ d_user,
                     float* __restrict__ d_query, float* __restrict__ d_item,
                     const float* __restrict__ norms) {
    const int64_t total = count * nvec;
    const float oml = 1.0f - lam;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / nvec;
        const int c = (int)(t % nvec);
        const float gsc = __ldg(dscore + b);
        const float4 q = ldg4(qf + b * q_ld + 4 * c);
        const float4 it = ldg4(itf + b * i_ld + 4 * c);
        float4 m = q;
        if (uf) {
            const float4 u = ldg4(uf + b * u_ld + 4 * c);
            m = make_float4(lam * q.x + oml * u.x, lam * q.y + oml * u.y,
                            lam * q.z + oml * u.z, lam * q.w + oml * u.w);
        }
        const int64_t o = b * (int64_t)nvec * 4 + 4 * c;   // gradients are dense [count, dim]
        float4 gi = f4_scale(gsc, m), gm = f4_scale(gsc, it);            // d score / d item, d score / d m  (dot product)
        if (norms) {
            // cosine: s = dot / (a bm);  ds/d item = m/(a bm) - dot item/(a^3 bm);  ds/d m symmetric
            const float dot = __ldg(norms + 3 * b), a = __ldg(norms + 3 * b + 1), bm = __ldg(norms + 3 * b + 2);
            const float inv = gsc / (a * bm), ci = dot / (a * a), cm = dot / (bm * bm);
            gi = make_float4(inv * (m.x - ci * it.x), inv * (m.y - ci * it.y), inv * (m.z - ci * it.z), inv * (m.w - ci * it.w));
            gm = make_float4(inv * (it.x - cm * m.x), inv * (it.y - cm * m.y), inv * (it.z - cm * m.z), inv * (it.w - cm * m.w));
        }
        if (d_item) stg4(d_item + o, gi);
        if (uf) {
            if (d_query) stg4(d_query + o, f4_scale(lam, gm));
            if (d_user) stg4(d_user + o, f4_scale(oml, gm));
        } else {
            if (d_query) stg4(d_query + o, gm);
        }
    }
}

// d_bias[i] = sum of dscore[b] over b with item_idx[b] == i.  Order-independent (hence
// bit-reproducible) without an O(B^2) scan: the addends are converted to 64-bit fixed point
// relative to max|dscore| (2^-38 of the largest addend, 14 bits finer than fp32) and summed with
// integer atomics, which are associative.
//   ws[0] = max |dscore| bits, ws[1 ..] = per-item accumulators
__global__ void __launch_bounds__(256)
hem_bias_max_kernel(const float* __restrict__ dscore, int64_t count, unsigned long long* __restrict__ ws) {
    unsigned m = 0;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < count; b += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(__ldg(dscore + b))));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(ws, (unsigned long long)m);
}
__device__ __forceinline__ int hem_bias_shift(const unsigned long long* ws) {
    const int e = (int)((unsigned)ws[0] >> 23) - 127;          // floor(log2(max |g|)) (max is finite, >= 0)
    return 38 - e;
}
__global__ void __launch_bounds__(256)
hem_bias_accum_kernel(const float* __restrict__ dscore, const int64_t* __restrict__ item_idx, int64_t count,
                      unsigned long long* __restrict__ ws) {
    const int sh = hem_bias_shift(ws);
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < count; b += (int64_t)gridDim.x * blockDim.x) {
        const long long q = __double2ll_rn(scalbn((double)__ldg(dscore + b), sh));
        atomicAdd(ws + 1 + __ldg(item_idx + b), (unsigned long long)q);
    }
}
__global__ void __launch_bounds__(256)
hem_bias_finish_kernel(const unsigned long long* __restrict__ ws, int64_t item_count, float* __restrict__ d_bias) {
    const int sh = hem_bias_shift(ws);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < item_count; i += (int64_t)gridDim.x * blockDim.x)
        d_bias[i] = (float)scalbn((double)(long long)ws[1 + i], -sh);
}

static unsigned blocks_for(int64_t n, int threads) {
    int64_t b = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = (int64_t)kNumSMs * 32;
    return (unsigned)(b < cap ? b : cap);
}

}  // namespace ihg

using namespace ihg;

extern "C" {

int ihg_copy_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows,
                  int32_t dim, void* stream) {
    IHG_REQUIRE(src && dst, "copy_rows: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0,
                "copy_rows: dim and leading dimensions must be multiples of 4");
    if (n_rows <= 0) return IHG_OK;
    copy_rows_kernel<<<blocks_for(n_rows * (dim / 4), 256), 256, 0, as_stream(stream)>>>(
        src, src_ld, dst, dst_ld, n_rows, dim / 4);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int ihg_gather_rows(const float* table, int64_t table_ld, const int64_t* idx, int64_t idx_offset,
                    int64_t count, float* out, int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(table && idx && out, "gather_rows: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && table_ld % 4 == 0 && out_ld % 4 == 0,
                "gather_rows: dim and leading dimensions must be multiples of 4");
    if (count <= 0) return IHG_OK;
    gather_rows_kernel<<<blocks_for(count * (dim / 4), 256), 256, 0, as_stream(stream)>>>(
        table, table_ld, idx, idx_offset, count, out, out_ld, dim / 4);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int ihg_halo_copy(const void* const* seg_src, void* const* seg_dst, const int64_t* seg_off, int32_t n_seg,
                  const int64_t* src_rows, int64_t src_ld, int64_t dst_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(seg_src && seg_dst && seg_off, "halo_copy: null pointer");
    IHG_REQUIRE(n_seg >= 1 && n_seg <= 16, "halo_copy: n_seg=%d outside [1, 16]", n_seg);
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0,
                "halo_copy: dim and leading dimensions must be multiples of 4");
    HaloSegs segs;
    segs.n = n_seg;
    for (int i = 0; i < n_seg; ++i) {
        segs.src[i] = static_cast<const float*>(seg_src[i]);
        segs.dst[i] = static_cast<float*>(seg_dst[i]);
        segs.off[i] = seg_off[i];
        IHG_REQUIRE(seg_off[i + 1] >= seg_off[i], "halo_copy: seg_off must be non-decreasing");
    }
    segs.off[n_seg] = seg_off[n_seg];
    const int64_t rows = seg_off[n_seg] - seg_off[0];
    IHG_REQUIRE(seg_off[0] == 0, "halo_copy: seg_off[0] must be 0");
    if (rows <= 0) return IHG_OK;
    halo_copy_kernel<<<blocks_for(rows * (dim / 4), 256), 256, 0, as_stream(stream)>>>(segs, src_rows, src_ld,
                                                                                       dst_ld, dim / 4);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int ihg_scatter_add_rows(const float* g, int64_t g_ld, const int64_t* idx, int64_t idx_offset,
                         int64_t count, float* out, int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(g && idx && out, "scatter_add_rows: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && g_ld % 4 == 0 && out_ld % 4 == 0,
                "scatter_add_rows: dim and leading dimensions must be multiples of 4");
    IHG_REQUIRE(count <= 65536, "scatter_add_rows: count=%lld exceeds the 65536-row batch limit", (long long)count);
    IHG_REQUIRE(dim <= 1024, "scatter_add_rows: dim=%d exceeds 1024", dim);
    if (count <= 0) return IHG_OK;
    if (count <= kScatterSmemIdx)
        scatter_add_rows_kernel<true><<<(unsigned)count, 128, 0, as_stream(stream)>>>(g, g_ld, idx, idx_offset,
                                                                                      count, out, out_ld, dim / 4);
    else
        scatter_add_rows_kernel<false><<<(unsigned)count, 128, 0, as_stream(stream)>>>(g, g_ld, idx, idx_offset,
                                                                                       count, out, out_ld, dim / 4);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int ihg_hem_score_fwd(const float* user_f, int64_t user_ld, const float* query_f, int64_t query_ld,
                      const float* item_f, int64_t item_ld, const float* items_bias,
                      const int64_t* item_idx, float lambda_muq, int64_t count, int32_t dim,
                      float* score, int32_t cosine, float* norms, void* stream) {
    IHG_REQUIRE(query_f && item_f && items_bias && score, "hem_score_fwd: null pointer");
    IHG_REQUIRE(!cosine || norms, "hem_score_fwd: the cosine scorer needs the norms buffer [count, 3]");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && query_ld % 4 == 0 && item_ld % 4 == 0 && (!user_f || user_ld % 4 == 0),
                "hem_score_fwd: dim and leading dimensions must be multiples of 4");
    if (count <= 0) return IHG_OK;
    hem_score_fwd_kernel<<<blocks_for(count, 8), 256, 0, as_stream(stream)>>>(
        user_f, user_ld, query_f, query_ld, item_f, item_ld, items_bias, item_idx, lambda_muq, count,
        dim / 4, score, cosine, norms);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int64_t ihg_hem_score_bwd_workspace_bytes(int64_t item_count) {
    return item_count > 0 ? (item_count + 1) * 8 : 0;
}

int ihg_hem_score_bwd(const float* dscore, const float* user_f, int64_t user_ld,
                      const float* query_f, int64_t query_ld, const float* item_f, int64_t item_ld,
                      const int64_t* item_idx, float lambda_muq, int64_t count, int32_t dim,
                      float* d_user, float* d_query, float* d_item, float* d_bias,
                      int64_t item_count, void* workspace, int64_t workspace_bytes, const float* norms,
                      void* stream) {
    IHG_REQUIRE(dscore && query_f && item_f, "hem_score_bwd: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && query_ld % 4 == 0 && item_ld % 4 == 0 && (!user_f || user_ld % 4 == 0),
                "hem_score_bwd: dim and leading dimensions must be multiples of 4");
    cudaStream_t st = as_stream(stream);
    if (d_bias) {
        IHG_REQUIRE(item_count > 0, "hem_score_bwd: item_count must be given with d_bias");
        if (item_idx) {
            IHG_REQUIRE(count <= 65536, "hem_score_bwd: count=%lld exceeds the 65536-row batch limit", (long long)count);
            IHG_REQUIRE(workspace && workspace_bytes >= ihg_hem_score_bwd_workspace_bytes(item_count),
                        "hem_score_bwd: workspace too small");
            unsigned long long* ws = static_cast<unsigned long long*>(workspace);
            IHG_CUDA(cudaMemsetAsync(ws, 0, (size_t)(item_count + 1) * 8, st));
            if (count > 0) {
                hem_bias_max_kernel<<<blocks_for(count, 256), 256, 0, st>>>(dscore, count, ws);
                IHG_LAUNCH_CHECK();
                hem_bias_accum_kernel<<<blocks_for(count, 256), 256, 0, st>>>(dscore, item_idx, count, ws);
                IHG_LAUNCH_CHECK();
            }
            hem_bias_finish_kernel<<<blocks_for(item_count, 256), 256, 0, st>>>(ws, item_count, d_bias);
            IHG_LAUNCH_CHECK();
        } else {
            IHG_REQUIRE(count == item_count, "hem_score_bwd: all-items form needs count == item_count");
            IHG_CUDA(cudaMemcpyAsync(d_bias, dscore, (size_t)count * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (count <= 0) return IHG_OK;
    hem_score_bwd_kernel<<<blocks_for(count * (dim / 4), 256), 256, 0, st>>>(
        dscore, user_f, user_ld, query_f, query_ld, item_f, item_ld, lambda_muq, count, dim / 4,
        d_user, d_query, d_item, norms);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // extern "C"
