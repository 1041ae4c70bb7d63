// Node-level typed Linear layers (exact fp32 FFMA path).
//
// Replaces the `feature_transform` Linear (/root/reference/Models/GnnLayers.py:224) and -- after
// hoisting -- the first-order blocks of the FeatureInteractor aggregation Linear
// (Models/CommonLayers.py:66,85): users only occupy slot 0, queries slot 1, items slot 2
// (Helpers/Graph.py:110-117), so  W_a[:, s*d:(s+1)*d] . X'[node of slot s]  is a per-node-type
// [N,d]x[d,d] product instead of three per-edge ones.  N << K*E, so these are minor kernels;
// accumulation is plain fp32 (no TF32) to stay inside the 1e-5 parity budget.
//
// Roofline: HBM for realistic d (arithmetic intensity d/4 flop/B at fp32): bytes per row =
// 4*(n_in + n_out) (+ addend), flops per row = 2*n_in*n_out.
#include "gemm_tile.cuh"
#include "tc_linear.h"

namespace ihg {

struct TypeTiles {
    int64_t b0, b1, n_rows;
    __host__ __device__ int64_t tiles(int t) const {
        const int64_t lo = t == 0 ? 0 : (t == 1 ? b0 : b1);
        const int64_t hi = t == 0 ? b0 : (t == 1 ? b1 : n_rows);
        return (hi - lo + kTileRows - 1) / kTileRows;
    }
    __host__ __device__ int64_t lo(int t) const { return t == 0 ? 0 : (t == 1 ? b0 : b1); }
    __host__ __device__ int64_t hi(int t) const { return t == 0 ? b0 : (t == 1 ? b1 : n_rows); }
};

template <int DPT>
__global__ void __launch_bounds__(kGemmThreads)
node_linear_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ w,
                   int n_types, int n_out, int n_in, int transpose_w,
                   const float* __restrict__ bias, const float* __restrict__ addend,
                   int64_t addend_ld, TypeTiles tt, float* __restrict__ y, int64_t y_ld) {
    constexpr int BN = 16 * DPT;
    constexpr int BP = BN + 1;
    __shared__ __align__(16) float At[kKC * kAtPitch];
    __shared__ float Bs[kKC * BP];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // locate this block's tile: tiles never straddle a node-type boundary
    int64_t bid = blockIdx.x;
    int type = 0;
    while (type < 2 && bid >= tt.tiles(type)) { bid -= tt.tiles(type); ++type; }
    const int64_t row0 = tt.lo(type) + bid * kTileRows;
    const int rows = (int)min((int64_t)kTileRows, tt.hi(type) - row0);
    const int wt = n_types > 1 ? type : 0;
    const float* W = w + (int64_t)wt * n_out * n_in;

    float acc[4][DPT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < DPT; ++j) acc[r][j] = 0.f;

    for (int kc0 = 0; kc0 < n_in; kc0 += kKC) {
        const int klen = min(kKC, n_in - kc0);
        {   // A chunk, transposed: At[k][r] = x[row0+r][kc0+k]   (global reads coalesced along k)
            const int k = tid & 31;
            for (int r = tid >> 5; r < kTileRows; r += 8)
                At[k * kAtPitch + r] = (k < klen && r < rows) ? __ldg(x + (row0 + r) * x_ld + kc0 + k) : 0.f;
        }
        if (!transpose_w) {   // Bs[k][n] = W[n][kc0+k]
            const int k = tid & 31;
            for (int n = tid >> 5; n < BN; n += 8)
                Bs[k * BP + n] = (k < klen && n < n_out) ? __ldg(W + (int64_t)n * n_in + kc0 + k) : 0.f;
        } else {              // Bs[k][n] = W[kc0+k][n]
            for (int idx = tid; idx < kKC * BN; idx += kGemmThreads) {
                const int k = idx / BN, n = idx % BN;
                Bs[k * BP + n] = (k < klen && n < n_out) ? __ldg(W + (int64_t)(kc0 + k) * n_out + n) : 0.f;
            }
        }
        __syncthreads();
        tile_fma_at<DPT>(At, Bs, BP, klen, tx, ty, acc);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int rr = ty * 4 + r;
        if (rr >= rows) continue;
        const int64_t row = row0 + rr;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int c = tx + 16 * j;
            if (c >= n_out) continue;
            float v = acc[r][j];
            if (bias) v += __ldg(bias + (int64_t)wt * n_out + c);
            if (addend) v += __ldg(addend + row * addend_ld + c);
            y[row * y_ld + c] = v;
        }
    }
}

// dw[t][n][k] partial over this block's row tiles; db[t][n] likewise.
template <int RG, int DPT>
__global__ void __launch_bounds__(kGemmThreads)
node_wgrad_kernel(const float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ x,
                  int64_t x_ld, TypeTiles tt, int n_types, int n_out, int n_in,
                  float* __restrict__ ws_dw, float* __restrict__ ws_db) {
    constexpr int BN = 16 * DPT;
    constexpr int BP = BN + 1;
    constexpr int AP = RG * 64 + 4;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                   // [64][AP]
    float* Bs = smem + kTileRows * AP;  // [64][BP]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int g = blockIdx.x, G = gridDim.x, t = blockIdx.y;
    const int64_t s0 = n_types > 1 ? tt.lo(t) : 0;
    const int64_t s1 = n_types > 1 ? tt.hi(t) : tt.n_rows;
    const int64_t tiles = (s1 - s0 + kTileRows - 1) / kTileRows;

    float acc[RG][4][DPT];
    float dbacc[RG][4];
#pragma unroll
    for (int rg = 0; rg < RG; ++rg)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            dbacc[rg][r] = 0.f;
#pragma unroll
            for (int j = 0; j < DPT; ++j) acc[rg][r][j] = 0.f;
        }
    const int nvo = n_out >> 2;
    for (int64_t tile = g; tile < tiles; tile += G) {
        const int64_t row0 = s0 + tile * kTileRows;
        const int rows = (int)min((int64_t)kTileRows, s1 - row0);
        for (int idx = tid; idx < kTileRows * (AP / 4); idx += kGemmThreads) {
            const int r = idx / (AP / 4), c = idx % (AP / 4);
            float4 v = f4_zero();
            if (r < rows && c < nvo) v = ldg4(dy + (row0 + r) * dy_ld + 4 * c);
            *reinterpret_cast<float4*>(As + r * AP + 4 * c) = v;
        }
        for (int idx = tid; idx < kTileRows * BN; idx += kGemmThreads) {
            const int r = idx / BN, k = idx % BN;
            Bs[r * BP + k] = (r < rows && k < n_in) ? __ldg(x + (row0 + r) * x_ld + k) : 0.f;
        }
        __syncthreads();
        tile_outer<RG, DPT>(As, AP, Bs, BP, kTileRows, tx, ty, acc);
        if (tx == 0) {
            for (int e = 0; e < kTileRows; ++e)
#pragma unroll
                for (int rg = 0; rg < RG; ++rg) {
                    const float4 a = *reinterpret_cast<const float4*>(As + e * AP + ty * 4 + 64 * rg);
                    dbacc[rg][0] += a.x; dbacc[rg][1] += a.y; dbacc[rg][2] += a.z; dbacc[rg][3] += a.w;
                }
        }
        __syncthreads();
    }
    const int T = gridDim.y;
    float* out_w = ws_dw + ((int64_t)g * T + t) * n_out * n_in;
    float* out_b = ws_db + ((int64_t)g * T + t) * n_out;
#pragma unroll
    for (int rg = 0; rg < RG; ++rg)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = ty * 4 + r + 64 * rg;
            if (n >= n_out) continue;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int k = tx + 16 * j;
                if (k < n_in) out_w[(int64_t)n * n_in + k] = acc[rg][r][j];
            }
            if (tx == 0) out_b[n] = dbacc[rg][r];
        }
}

static int wgrad_grid_g(int n_types) { return n_types > 1 ? 98 : 296; }  // 2 blocks per SM in total

}  // namespace ihg

using namespace ihg;

extern "C" {

int ihg_node_linear(const float* x, int64_t x_ld, const float* w, int32_t n_types, int32_t n_out,
                    int32_t n_in, int32_t transpose_w, const float* bias, const float* addend,
                    int64_t addend_ld, int64_t n_rows, int64_t bound0, int64_t bound1, float* y,
                    int64_t y_ld, void* stream) {
    IHG_REQUIRE(x && w && y, "node_linear: null pointer");
    IHG_REQUIRE(n_types == 1 || n_types == 3, "node_linear: n_types must be 1 or 3");
    IHG_REQUIRE(n_out > 0 && n_in > 0 && n_out <= 128 && n_in <= 128 && n_out % 4 == 0 && n_in % 4 == 0,
                "node_linear: n_out=%d n_in=%d must be multiples of 4, <= 128", n_out, n_in);
    IHG_REQUIRE(n_rows > 0 && x_ld >= n_in && y_ld >= n_out, "node_linear: bad shapes");
    if (n_types == 1) bound0 = bound1 = n_rows;
    IHG_REQUIRE(0 <= bound0 && bound0 <= bound1 && bound1 <= n_rows, "node_linear: bad type bounds");
    cudaStream_t st = as_stream(stream);
    if (node_linear_ts_eligible(n_out, n_in, x_ld, y_ld, addend, addend_ld))
        return launch_node_linear_ts(x, x_ld, w, n_types, n_out, n_in, transpose_w, bias, addend, addend_ld, n_rows,
                                     bound0, bound1, y, y_ld, st);
    TypeTiles tt{bound0, bound1, n_rows};
    const int64_t blocks = tt.tiles(0) + tt.tiles(1) + tt.tiles(2);
#define IHG_NL_CASE(D)                                                                              \
    node_linear_kernel<D><<<(unsigned)blocks, kGemmThreads, 0, st>>>(x, x_ld, w, n_types, n_out, n_in, \
        transpose_w, bias, addend, addend_ld, tt, y, y_ld)
    if (n_out <= 16) IHG_NL_CASE(1);
    else if (n_out <= 32) IHG_NL_CASE(2);
    else if (n_out <= 64) IHG_NL_CASE(4);
    else IHG_NL_CASE(8);
#undef IHG_NL_CASE
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int64_t ihg_node_linear_wgrad_workspace_bytes(int32_t n_types, int32_t n_out, int32_t n_in) {
    const int64_t G = wgrad_grid_g(n_types);
    const int64_t simt = ws_slice(G * n_types * (int64_t)n_out * n_in, 4) + ws_slice(G * n_types * (int64_t)n_out, 4) + 1024;
    const int64_t tcb = node_wgrad_tc_eligible(n_types, n_out, n_in) ? node_wgrad_tc_workspace_bytes(n_types, n_out, n_in) : 0;
    return simt > tcb ? simt : tcb;
}

int ihg_node_linear_wgrad(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld,
                          int64_t n_rows, int64_t bound0, int64_t bound1, int32_t n_types,
                          int32_t n_out, int32_t n_in, float* dw, float* db, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(dy && x && dw && workspace, "node_linear_wgrad: null pointer");
    IHG_REQUIRE(n_types == 1 || n_types == 3, "node_linear_wgrad: n_types must be 1 or 3");
    IHG_REQUIRE(n_out > 0 && n_in > 0 && n_out <= 128 && n_in <= 128 && n_out % 4 == 0 && n_in % 4 == 0,
                "node_linear_wgrad: n_out=%d n_in=%d must be multiples of 4, <= 128", n_out, n_in);
    IHG_REQUIRE(dy_ld % 4 == 0, "node_linear_wgrad: dy_ld must be a multiple of 4");
    IHG_REQUIRE(workspace_bytes >= ihg_node_linear_wgrad_workspace_bytes(n_types, n_out, n_in),
                "node_linear_wgrad: workspace too small");
    if (n_types == 1) bound0 = bound1 = n_rows;
    IHG_REQUIRE(n_rows > 0 && 0 <= bound0 && bound0 <= bound1 && bound1 <= n_rows, "node_linear_wgrad: bad bounds");
    cudaStream_t st = as_stream(stream);
    if (node_wgrad_tc_eligible(n_types, n_out, n_in) && x_ld % 4 == 0)
        return launch_node_wgrad_tc(dy, dy_ld, x, x_ld, n_rows, bound0, bound1, n_types, n_out, n_in, dw, db,
                                    workspace, st);
    TypeTiles tt{bound0, bound1, n_rows};
    const int G = wgrad_grid_g(n_types);
    Workspace ws(workspace, workspace_bytes);
    float* ws_dw = ws.take<float>((int64_t)G * n_types * n_out * n_in);
    float* ws_db = ws.take<float>((int64_t)G * n_types * n_out);
    dim3 grid(G, n_types);
#define IHG_WG_CASE(RG, D)                                                                        \
    do {                                                                                          \
        const size_t smem = (size_t)kTileRows * ((RG * 64 + 4) + (16 * D + 1)) * sizeof(float);   \
        IHG_CUDA(cudaFuncSetAttribute(node_wgrad_kernel<RG, D>,                                   \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        node_wgrad_kernel<RG, D><<<grid, kGemmThreads, smem, st>>>(dy, dy_ld, x, x_ld, tt, n_types, \
                                                                   n_out, n_in, ws_dw, ws_db);   \
    } while (0)
    const int dsel = n_in <= 16 ? 1 : (n_in <= 32 ? 2 : (n_in <= 64 ? 4 : 8));
    if (n_out <= 64) {
        if (dsel == 1) IHG_WG_CASE(1, 1); else if (dsel == 2) IHG_WG_CASE(1, 2);
        else if (dsel == 4) IHG_WG_CASE(1, 4); else IHG_WG_CASE(1, 8);
    } else {
        if (dsel == 1) IHG_WG_CASE(2, 1); else if (dsel == 2) IHG_WG_CASE(2, 2);
        else if (dsel == 4) IHG_WG_CASE(2, 4); else IHG_WG_CASE(2, 8);
    }
#undef IHG_WG_CASE
    IHG_LAUNCH_CHECK();
    const int64_t nw = (int64_t)n_types * n_out * n_in;
    sum_partials_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, st>>>(ws_dw, nw, G, nw, dw);
    IHG_LAUNCH_CHECK();
    if (db) {
        const int64_t nb = (int64_t)n_types * n_out;
        sum_partials_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(ws_db, nb, G, nb, db);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}

}  // extern "C"
