// K-B: deterministic segmented reduction over CSR rows (the edge->node scatter-mean and all
// of its transposes).  Replaces torch_sparse.matmul(incidence, Ef) * Dv^-1
// (/root/reference/Models/GnnLayers.py:233-234), the index_put_(accumulate=True) backward of
// the three row gathers (CommonLayers.py:70-72) and EmbeddingBag(mean) (EmbeddingLayers.py:79).
//
// Work decomposition: one warp per row chunk (<= chunk_len incidences, from the plan built by
// ihg_segment_plan_build), LPR lanes x float4 across the feature dimension and 32/LPR rows in
// flight per step, `kUnroll` steps of independent 128-bit gathers issued before the adds.
// Rows longer than chunk_len (Zipf head nodes) are split; their partial sums are combined by
// a second tiny kernel in ascending chunk order.  Summation order is a pure function of the
// plan => bitwise run-to-run determinism, no float atomics.
//
// Roofline: HBM.  Algorithmic bytes per incidence: 4 (col) + 4*dim (source row), per row:
// 8 (plan) + 4 (scale) + 4*dim (output row).
#include "common.cuh"

namespace ihg {

constexpr int kSegWarpsPerBlock = 8;
constexpr int kSegUnroll = 4;

template <int LPR, int VPL>
__device__ __forceinline__ void seg_accumulate(const float* __restrict__ src, int64_t src_ld,
                                               int32_t src_row_mul, int slot,
                                               const float* __restrict__ src_scale,
                                               const int32_t* __restrict__ col, int begin, int end,
                                               int nvec, int lane, float4 (&acc)[VPL]) {
    constexpr int G = 32 / LPR;
    const int g = lane / LPR, c = lane % LPR;
    for (int j0 = begin; j0 < end; j0 += 32) {
        const int cnt = min(32, end - j0);
        const int my_col = (lane < cnt) ? __ldg(col + j0 + lane) : 0;
        for (int k = 0; k < cnt; k += G * kSegUnroll) {
            float4 v[kSegUnroll][VPL];
            float sc[kSegUnroll];
#pragma unroll
            for (int u = 0; u < kSegUnroll; ++u) {
                const int idx = k + u * G + g;
                const int e = __shfl_sync(0xffffffffu, my_col, idx & 31);
                const bool ok = idx < cnt;
                const int64_t s = (int64_t)e * src_row_mul + slot;
                sc[u] = (ok && src_scale) ? __ldg(src_scale + e) : 1.0f;
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = c + w * LPR;
                    v[u][w] = (ok && cv < nvec) ? ldg4(src + s * src_ld + 4 * cv) : f4_zero();
                }
            }
#pragma unroll
            for (int u = 0; u < kSegUnroll; ++u)
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    if (src_scale) f4_fma(acc[w], sc[u], v[u][w]);
                    else f4_add(acc[w], v[u][w]);
                }
        }
    }
    // combine the 32/LPR row groups (fixed tree)
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1)
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            acc[w].x += __shfl_xor_sync(0xffffffffu, acc[w].x, o);
            acc[w].y += __shfl_xor_sync(0xffffffffu, acc[w].y, o);
            acc[w].z += __shfl_xor_sync(0xffffffffu, acc[w].z, o);
            acc[w].w += __shfl_xor_sync(0xffffffffu, acc[w].w, o);
        }
}

template <int LPR, int VPL>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32)
segment_reduce_kernel(const float* __restrict__ src, int64_t src_ld, int32_t src_row_mul,
                      int64_t bound0, int64_t bound1, const float* __restrict__ src_scale,
                      const float* __restrict__ row_scale, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, int32_t chunk_len, int64_t n_seg,
                      const int32_t* __restrict__ seg_row, const int32_t* __restrict__ seg_begin,
                      const int32_t* __restrict__ seg_part, float* __restrict__ partial,
                      float* __restrict__ out, int64_t out_ld, int dim) {
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    const int row = __ldg(seg_row + seg);
    const int begin = __ldg(seg_begin + seg);
    const int part = __ldg(seg_part + seg);
    const int row_end = __ldg(rowptr + row + 1);
    const int end = min(begin + chunk_len, row_end);
    const int slot = (row >= bound0) + (row >= bound1);
    const int nvec = dim >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) acc[w] = f4_zero();
    seg_accumulate<LPR, VPL>(src, src_ld, src_row_mul, slot, src_scale, col, begin, end, nvec, lane, acc);
    if (lane < LPR) {
        if (part < 0) {
            const float rs = row_scale ? __ldg(row_scale + row) : 1.0f;
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
                const int cv = lane + w * LPR;
                if (cv < nvec) stg4(out + (int64_t)row * out_ld + 4 * cv, f4_scale(rs, acc[w]));
            }
        } else {
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
                const int cv = lane + w * LPR;
                if (cv < nvec) stg4(partial + (int64_t)part * dim + 4 * cv, acc[w]);
            }
        }
    }
}

// one warp per split row: out[row] = row_scale * (partial[p0] + partial[p0+1] + ...) in order
template <int LPR, int VPL>
__global__ void __launch_bounds__(kSegWarpsPerBlock * 32)
segment_fixup_kernel(const float* __restrict__ partial, const int32_t* __restrict__ split_row,
                     const int32_t* __restrict__ split_ptr, int64_t n_split,
                     const float* __restrict__ row_scale, float* __restrict__ out, int64_t out_ld,
                     int dim) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kSegWarpsPerBlock + (threadIdx.x >> 5);
    if (i >= n_split || lane >= LPR) return;
    const int row = split_row[i];
    const int p0 = split_ptr[i], p1 = split_ptr[i + 1];
    const int nvec = dim >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) acc[w] = f4_zero();
    for (int p = p0; p < p1; ++p)
#pragma unroll
        for (int w = 0; w < VPL; ++w) {
            const int cv = lane + w * LPR;
            if (cv < nvec) f4_add(acc[w], ldg4(partial + (int64_t)p * dim + 4 * cv));
        }
    const float rs = row_scale ? row_scale[row] : 1.0f;
#pragma unroll
    for (int w = 0; w < VPL; ++w) {
        const int cv = lane + w * LPR;
        if (cv < nvec) stg4(out + (int64_t)row * out_ld + 4 * cv, f4_scale(rs, acc[w]));
    }
}

template <int LPR, int VPL>
static int launch_segment_reduce(const ihg_csr* g, const float* src, int64_t src_ld, int32_t mul,
                                 int64_t b0, int64_t b1, const float* src_scale,
                                 const float* row_scale, float* partial, float* out, int64_t out_ld,
                                 int dim, cudaStream_t st) {
    const unsigned blocks = (unsigned)ceil_div(g->n_seg, kSegWarpsPerBlock);
    segment_reduce_kernel<LPR, VPL><<<blocks, kSegWarpsPerBlock * 32, 0, st>>>(
        src, src_ld, mul, b0, b1, src_scale, row_scale, g->rowptr, g->col, g->chunk_len, g->n_seg,
        g->seg_row, g->seg_begin, g->seg_part, partial, out, out_ld, dim);
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
    IHG_REQUIRE(g->n_rows > 0 && g->n_seg >= g->n_rows && g->rowptr && g->seg_row && g->seg_begin && g->seg_part,
                "segment_reduce: incomplete csr plan");
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
