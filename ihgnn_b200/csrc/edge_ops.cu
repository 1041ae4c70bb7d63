// Node -> hyperedge kernels (K-A family), exact-fp32 SIMT path.
//
//   ihg_edge_gather_sum      out[e] = alpha * sum_s scale[n_s] * src[n_s] (+bias)
//        order-1 FeatureInteractor after hoisting (/root/reference/Models/CommonLayers.py:58-66),
//        HGCN's H^T product (Models/GnnLayers.py:148-149) and the backward of the edge->node
//        SpMM (dEf = H^T (Dv^-1 dOut), autograd of GnnLayers.py:233-234).  HBM-bound:
//        12 (i3) + 3*4*dim (rows) + 4*dim (out) bytes per hyperedge.
//   ihg_edge_interact_fwd    order 2/3 FeatureInteractor (CommonLayers.py:68-85): the Hadamard
//        products u*q, q*i, i*u (, u*q*i) are formed on the fly and contracted with
//        aggregation.weight[:, 3d:] -- the [E, K*d] concatenation of the reference
//        (CommonLayers.py:82-84; 17.9 GB at the CIKM shape) never exists.
//   ihg_edge_interact_bwd    its backward: per-slot input gradients [E,3,d] and the weight
//        gradient (recomputing the products from re-gathered rows; two-pass deterministic sum).
//
// The contraction here runs on fp32 FFMA so that results stay within ~1e-7 of the reference;
// flops per hyperedge = 2*nb*d^2 (nb = 3 or 4 product blocks) forward, 2x that backward.
#include "gemm_tile.cuh"
#include "tc_linear.h"

namespace ihg {

// =========================================================================================
// gather-sum
// =========================================================================================
template <int LPR, int VPL, int kGsUnroll, int MINB>
__global__ void __launch_bounds__(256, MINB)
edge_gather_sum_kernel(const float* __restrict__ src, int64_t src_ld,
                       const float* __restrict__ node_scale, float alpha,
                       const float* __restrict__ bias, const int32_t* __restrict__ i3,
                       int64_t E, float* __restrict__ out, int64_t out_ld, int dim) {
    constexpr int G = 32 / LPR;                       // edges per warp step
    const int lane = threadIdx.x & 31;
    const int g = lane / LPR, c = lane % LPR;
    const int nvec = dim >> 2;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float4 bv[VPL];
#pragma unroll
    for (int w = 0; w < VPL; ++w) {
        const int cv = c + w * LPR;
        bv[w] = (bias && cv < nvec) ? ldg4(bias + 4 * cv) : f4_zero();
    }
    for (int64_t e0 = warp * (G * kGsUnroll); e0 < E; e0 += nwarps * (G * kGsUnroll)) {
        float4 v[kGsUnroll][3][VPL];
        float sc[kGsUnroll][3];
#pragma unroll
        for (int u = 0; u < kGsUnroll; ++u) {
            const int64_t e = e0 + u * G + g;
            const bool ok = e < E;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int n = ok ? __ldg(i3 + 3 * e + s) : 0;
                sc[u][s] = (ok && node_scale) ? __ldg(node_scale + n) : 1.0f;
#pragma unroll
                for (int w = 0; w < VPL; ++w) {
                    const int cv = c + w * LPR;
                    v[u][s][w] = (ok && cv < nvec) ? ldg4(src + (int64_t)n * src_ld + 4 * cv) : f4_zero();
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kGsUnroll; ++u) {
            const int64_t e = e0 + u * G + g;
            if (e >= E) continue;
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
                const int cv = c + w * LPR;
                if (cv >= nvec) continue;
                float4 a = f4_scale(sc[u][0], v[u][0][w]);
                f4_fma(a, sc[u][1], v[u][1][w]);
                f4_fma(a, sc[u][2], v[u][2][w]);
                float4 r = bv[w];
                f4_fma(r, alpha, a);
                stg4(out + e * out_ld + 4 * cv, r);
            }
        }
    }
}

// =========================================================================================
// interaction contraction, forward
// =========================================================================================
// product block b of hyperedge (u,q,i) at feature k:  0: u*q  1: q*i  2: i*u  3: u*q*i
__device__ __forceinline__ float interact_value(int b, float u, float q, float i) {
    return b == 0 ? u * q : (b == 1 ? q * i : (b == 2 ? i * u : (u * q) * i));
}

template <int DPT>
__global__ void __launch_bounds__(kGemmThreads)
edge_interact_fwd_kernel(const float* __restrict__ xp, int64_t xp_ld, const float* __restrict__ p,
                         int64_t p_ld, const float* __restrict__ w_hi, int64_t w_ld, int nb,
                         const int32_t* __restrict__ i3, int64_t E, float* __restrict__ ef,
                         int64_t ef_ld, int dim) {
    constexpr int BN = 16 * DPT;
    constexpr int BP = BN + 1;
    __shared__ __align__(16) float Zt[kKC * kAtPitch];   // Zt[k][edge]
    __shared__ float Ws[kKC * BP];                       // Ws[k][n] = w_hi[n][b*dim + kc0 + k]
    __shared__ int32_t sI3[kTileRows * 3];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t e0 = (int64_t)blockIdx.x * kTileRows;
    const int rows = (int)min((int64_t)kTileRows, E - e0);
    for (int idx = tid; idx < kTileRows * 3; idx += kGemmThreads)
        sI3[idx] = idx < rows * 3 ? __ldg(i3 + 3 * e0 + idx) : 0;
    __syncthreads();

    float acc[4][DPT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < DPT; ++j) acc[r][j] = 0.f;

    for (int b = 0; b < nb; ++b) {
        for (int kc0 = 0; kc0 < dim; kc0 += kKC) {
            const int klen = min(kKC, dim - kc0);
            const int k = tid & 31;
            // Z chunk: lanes run along k => each gathered row is read in 128-byte pieces
            for (int r = tid >> 5; r < kTileRows; r += 8) {
                float z = 0.f;
                if (k < klen && r < rows) {
                    const float u = __ldg(xp + (int64_t)sI3[3 * r + 0] * xp_ld + kc0 + k);
                    const float q = __ldg(xp + (int64_t)sI3[3 * r + 1] * xp_ld + kc0 + k);
                    const float i = __ldg(xp + (int64_t)sI3[3 * r + 2] * xp_ld + kc0 + k);
                    z = interact_value(b, u, q, i);
                }
                Zt[k * kAtPitch + r] = z;
            }
            for (int n = tid >> 5; n < BN; n += 8)
                Ws[k * BP + n] = (k < klen && n < dim)
                    ? __ldg(w_hi + (int64_t)n * w_ld + (int64_t)b * dim + kc0 + k) : 0.f;
            __syncthreads();
            tile_fma_at<DPT>(Zt, Ws, BP, klen, tx, ty, acc);
            __syncthreads();
        }
    }
    // epilogue: add the hoisted first-order part p[u]+p[q]+p[i] (bias folded into p)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int rr = ty * 4 + r;
        if (rr >= rows) continue;
        const int64_t nu = sI3[3 * rr + 0], nq = sI3[3 * rr + 1], ni = sI3[3 * rr + 2];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int c = tx + 16 * j;
            if (c >= dim) continue;
            const float base = (__ldg(p + nu * p_ld + c) + __ldg(p + nq * p_ld + c)) + __ldg(p + ni * p_ld + c);
            ef[(e0 + rr) * ef_ld + c] = base + acc[r][j];
        }
    }
}

// =========================================================================================
// interaction contraction, backward (a): per-slot input gradients
//   dz_b[e][k] = sum_n def[e][n] * w_hi[n][b*dim + k];   fold the product rule into du,dq,di
// =========================================================================================
template <int DPT>
__global__ void __launch_bounds__(kGemmThreads)
edge_interact_bwd_slot_kernel(const float* __restrict__ xp, int64_t xp_ld,
                              const float* __restrict__ def, int64_t def_ld,
                              const float* __restrict__ w_hi, int64_t w_ld, int nb,
                              const int32_t* __restrict__ i3, int64_t E,
                              float* __restrict__ slot_grad, int dim) {
    constexpr int BN = 16 * DPT;
    constexpr int BP = BN + 1;
    __shared__ __align__(16) float Gt[kKC * kAtPitch];   // Gt[n][edge] = def[edge][nc0+n]
    __shared__ float Ws[kKC * BP];                       // Ws[n][k] = w_hi[nc0+n][b*dim + c0 + k]
    __shared__ int32_t sI3[kTileRows * 3];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t e0 = (int64_t)blockIdx.x * kTileRows;
    const int rows = (int)min((int64_t)kTileRows, E - e0);
    for (int idx = tid; idx < kTileRows * 3; idx += kGemmThreads)
        sI3[idx] = idx < rows * 3 ? __ldg(i3 + 3 * e0 + idx) : 0;
    __syncthreads();

    // column groups of BN features keep the register footprint bounded for dim = 128
    for (int c0 = 0; c0 < dim; c0 += BN) {
        float uv[4][DPT], qv[4][DPT], iv[4][DPT];
        float du[4][DPT], dq[4][DPT], di[4][DPT];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int c = c0 + tx + 16 * j;
                const bool ok = rr < rows && c < dim;
                uv[r][j] = ok ? __ldg(xp + (int64_t)sI3[3 * rr + 0] * xp_ld + c) : 0.f;
                qv[r][j] = ok ? __ldg(xp + (int64_t)sI3[3 * rr + 1] * xp_ld + c) : 0.f;
                iv[r][j] = ok ? __ldg(xp + (int64_t)sI3[3 * rr + 2] * xp_ld + c) : 0.f;
                du[r][j] = dq[r][j] = di[r][j] = 0.f;
            }
        }
        for (int b = 0; b < nb; ++b) {
            float acc[4][DPT];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < DPT; ++j) acc[r][j] = 0.f;
            for (int nc0 = 0; nc0 < dim; nc0 += kKC) {
                const int nlen = min(kKC, dim - nc0);
                const int n = tid & 31;
                for (int r = tid >> 5; r < kTileRows; r += 8)
                    Gt[n * kAtPitch + r] = (n < nlen && r < rows)
                        ? __ldg(def + (e0 + r) * def_ld + nc0 + n) : 0.f;
                for (int idx = tid; idx < kKC * BN; idx += kGemmThreads) {
                    const int nn = idx / BN, k = idx % BN;
                    Ws[nn * BP + k] = (nn < nlen && c0 + k < dim)
                        ? __ldg(w_hi + (int64_t)(nc0 + nn) * w_ld + (int64_t)b * dim + c0 + k) : 0.f;
                }
                __syncthreads();
                tile_fma_at<DPT>(Gt, Ws, BP, nlen, tx, ty, acc);
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < DPT; ++j) {
                    const float dz = acc[r][j];
                    const float u = uv[r][j], q = qv[r][j], i = iv[r][j];
                    if (b == 0) { du[r][j] = fmaf(dz, q, du[r][j]); dq[r][j] = fmaf(dz, u, dq[r][j]); }
                    else if (b == 1) { dq[r][j] = fmaf(dz, i, dq[r][j]); di[r][j] = fmaf(dz, q, di[r][j]); }
                    else if (b == 2) { di[r][j] = fmaf(dz, u, di[r][j]); du[r][j] = fmaf(dz, i, du[r][j]); }
                    else {
                        du[r][j] = fmaf(dz, q * i, du[r][j]);
                        dq[r][j] = fmaf(dz, u * i, dq[r][j]);
                        di[r][j] = fmaf(dz, u * q, di[r][j]);
                    }
                }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
            if (rr >= rows) continue;
            float* o = slot_grad + (e0 + rr) * 3 * (int64_t)dim;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int c = c0 + tx + 16 * j;
                if (c >= dim) continue;
                o[c] = du[r][j];
                o[dim + c] = dq[r][j];
                o[2 * dim + c] = di[r][j];
            }
        }
    }
}

// =========================================================================================
// interaction contraction, backward (b): weight gradient
//   dw_hi[n][b*dim + k] = sum_e def[e][n] * z_b[e][k]     block (g, b) owns product block b
// =========================================================================================
template <int RG, int DPT>
__global__ void __launch_bounds__(kGemmThreads)
edge_interact_bwd_wgrad_kernel(const float* __restrict__ xp, int64_t xp_ld,
                               const float* __restrict__ def, int64_t def_ld,
                               const int32_t* __restrict__ i3, int64_t E, int dim,
                               float* __restrict__ ws_dw) {
    constexpr int BN = 16 * DPT;
    constexpr int BP = BN + 1;
    constexpr int AP = RG * 64 + 4;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                    // [64][AP]  def tile
    float* Zs = smem + kTileRows * AP;   // [64][BP]  product tile
    __shared__ int32_t sI3[kTileRows * 3];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int g = blockIdx.x, G = gridDim.x, b = blockIdx.y, nb = gridDim.y;
    const int64_t tiles = (E + kTileRows - 1) / kTileRows;
    const int nvo = dim >> 2;

    float acc[RG][4][DPT];
#pragma unroll
    for (int rg = 0; rg < RG; ++rg)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < DPT; ++j) acc[rg][r][j] = 0.f;

    for (int64_t tile = g; tile < tiles; tile += G) {
        const int64_t e0 = tile * kTileRows;
        const int rows = (int)min((int64_t)kTileRows, E - e0);
        for (int idx = tid; idx < kTileRows * 3; idx += kGemmThreads)
            sI3[idx] = idx < rows * 3 ? __ldg(i3 + 3 * e0 + idx) : 0;
        for (int idx = tid; idx < kTileRows * (AP / 4); idx += kGemmThreads) {
            const int r = idx / (AP / 4), c = idx % (AP / 4);
            float4 v = f4_zero();
            if (r < rows && c < nvo) v = ldg4(def + (e0 + r) * def_ld + 4 * c);
            *reinterpret_cast<float4*>(As + r * AP + 4 * c) = v;
        }
        __syncthreads();
        for (int idx = tid; idx < kTileRows * BN; idx += kGemmThreads) {
            const int r = idx / BN, k = idx % BN;
            float z = 0.f;
            if (r < rows && k < dim) {
                const float u = __ldg(xp + (int64_t)sI3[3 * r + 0] * xp_ld + k);
                const float q = __ldg(xp + (int64_t)sI3[3 * r + 1] * xp_ld + k);
                const float i = __ldg(xp + (int64_t)sI3[3 * r + 2] * xp_ld + k);
                z = interact_value(b, u, q, i);
            }
            Zs[r * BP + k] = z;
        }
        __syncthreads();
        tile_outer<RG, DPT>(As, AP, Zs, BP, kTileRows, tx, ty, acc);
        __syncthreads();
    }
    float* out = ws_dw + ((int64_t)g * nb + b) * dim * dim;   // [n][k] of block b
#pragma unroll
    for (int rg = 0; rg < RG; ++rg)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = ty * 4 + r + 64 * rg;
            if (n >= dim) continue;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int k = tx + 16 * j;
                if (k < dim) out[(int64_t)n * dim + k] = acc[rg][r][j];
            }
        }
}

// dw_hi[n][b*dim + k] = sum_g ws[g][b][n][k]
__global__ void __launch_bounds__(256)
interact_wgrad_reduce_kernel(const float* __restrict__ ws, int G, int nb, int dim,
                             float* __restrict__ dw_hi) {
    const int64_t total = (int64_t)nb * dim * dim;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = idx % dim;
        const int n = (idx / dim) % dim;
        const int b = idx / ((int64_t)dim * dim);
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += ws[(((int64_t)g * nb + b) * dim + n) * dim + k];
        dw_hi[(int64_t)n * nb * dim + (int64_t)b * dim + k] = s;
    }
}

constexpr int kInteractWgradG = 74;   // x nb (3 or 4) blocks ~= 2 per SM

}  // namespace ihg

using namespace ihg;

extern "C" {

int ihg_edge_gather_sum(const float* src, int64_t src_ld, const float* node_scale, float alpha,
                        const float* bias, const int32_t* i3, int64_t E, float* out,
                        int64_t out_ld, int32_t dim, void* stream) {
    IHG_REQUIRE(src && i3 && out, "edge_gather_sum: null pointer");
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 256, "edge_gather_sum: dim=%d must be a multiple of 4, <= 256", dim);
    IHG_REQUIRE(src_ld % 4 == 0 && out_ld % 4 == 0 && src_ld >= dim && out_ld >= dim,
                "edge_gather_sum: leading dimensions must be multiples of 4 and >= dim");
    if (E == 0) return IHG_OK;
    cudaStream_t st = as_stream(stream);
    const int nvec = dim / 4;
    // (unroll 2, >= 4 resident blocks) measured best: occupancy beats per-thread ILP for this
    // gather (113 us vs 154 us per call at the amazon-full shape, 903 vs 1064 us at cikm)
#define IHG_GS_CASE(L, V)                                                                         \
    do {                                                                                          \
        const int64_t per_block = 8 * (32 / L) * 2;                                               \
        int64_t blocks = ceil_div(E, per_block);                                                  \
        if (blocks > kNumSMs * 64) blocks = kNumSMs * 64;                                         \
        edge_gather_sum_kernel<L, V, 2, 4><<<(unsigned)blocks, 256, 0, st>>>(src, src_ld,         \
            node_scale, alpha, bias, i3, E, out, out_ld, dim);                                    \
    } while (0)
    if (nvec <= 1) IHG_GS_CASE(1, 1);
    else if (nvec <= 2) IHG_GS_CASE(2, 1);
    else if (nvec <= 4) IHG_GS_CASE(4, 1);
    else if (nvec <= 8) IHG_GS_CASE(8, 1);
    else if (nvec <= 16) IHG_GS_CASE(16, 1);
    else if (nvec <= 32) IHG_GS_CASE(32, 1);
    else IHG_GS_CASE(32, 2);
#undef IHG_GS_CASE
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

static int check_interact(const char* what, int32_t order, int32_t dim) {
    IHG_REQUIRE(order == 2 || order == 3, "%s: order=%d must be 2 or 3", what, order);
    IHG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 128, "%s: dim=%d must be a multiple of 4, <= 128", what, dim);
    return IHG_OK;
}

int64_t ihg_edge_interact_fwd_workspace_bytes(int32_t dim, int32_t order) {
    const int nb = order == 3 ? 4 : 3;
    return interact_tc_eligible(dim) ? interact_fwd_tc_workspace_bytes(dim, nb) : 0;
}

int ihg_edge_interact_fwd(const float* xp, int64_t xp_ld, const float* p, int64_t p_ld,
                          const float* w_hi, int64_t w_ld, int32_t order, const int32_t* i3,
                          int64_t E, float* ef, int64_t ef_ld, int32_t dim, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(xp && p && w_hi && i3 && ef, "edge_interact_fwd: null pointer");
    if (int rc = check_interact("edge_interact_fwd", order, dim)) return rc;
    if (E == 0) return IHG_OK;
    const int nb = order == 3 ? 4 : 3;
    cudaStream_t st = as_stream(stream);
    (void)workspace; (void)workspace_bytes;      // the hoisted form is the exact-fp32 FFMA kernel for every dim
    const unsigned blocks = (unsigned)ceil_div(E, kTileRows);
#define IHG_IF_CASE(D) edge_interact_fwd_kernel<D><<<blocks, kGemmThreads, 0, st>>>(xp, xp_ld, p, p_ld, w_hi, w_ld, nb, i3, E, ef, ef_ld, dim)
    if (dim <= 16) IHG_IF_CASE(1);
    else if (dim <= 32) IHG_IF_CASE(2);
    else if (dim <= 64) IHG_IF_CASE(4);
    else IHG_IF_CASE(8);
#undef IHG_IF_CASE
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

int ihg_feature_interact_supported(int32_t dim) {
    return interact_tc_eligible(dim) ? 1 : 0;
}

int ihg_feature_interact_fwd(const float* xp, int64_t xp_ld, const float* w_agg, int64_t w_ld,
                             const float* bias, int32_t order, const int32_t* i3, int64_t E,
                             float* ef, int64_t ef_ld, int32_t dim, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(xp && w_agg && i3 && ef, "feature_interact_fwd: null pointer");
    if (int rc = check_interact("feature_interact_fwd", order, dim)) return rc;
    IHG_REQUIRE(interact_tc_eligible(dim) && xp_ld % 4 == 0 && ef_ld % 4 == 0,
                "feature_interact_fwd: dim=%d is not supported by the tensor-core path", dim);
    IHG_REQUIRE(workspace && workspace_bytes >= ihg_edge_interact_fwd_workspace_bytes(dim, order),
                "feature_interact_fwd: workspace too small");
    if (E == 0) return IHG_OK;
    return launch_interact_fwd_full_ts(xp, xp_ld, w_agg, w_ld, bias, order == 3 ? 4 : 3, i3, E, ef, ef_ld, dim,
                                       workspace, as_stream(stream));
}

int64_t ihg_edge_interact_bwd_workspace_bytes(int32_t dim, int32_t order) {
    const int nb = order == 3 ? 4 : 3;
    const int64_t simt = ws_slice((int64_t)kInteractWgradG * nb * dim * dim, 4) + 1024;
    const int64_t tcb = interact_tc_eligible(dim) ? interact_bwd_tc_workspace_bytes(dim, nb) : 0;
    return simt > tcb ? simt : tcb;
}

int ihg_edge_interact_bwd(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                          const float* w_hi, int64_t w_ld, int32_t order, const int32_t* i3,
                          int64_t E, float* slot_grad, float* dw_hi, int32_t dim, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(xp && def && w_hi && i3 && slot_grad && dw_hi && workspace, "edge_interact_bwd: null pointer");
    if (int rc = check_interact("edge_interact_bwd", order, dim)) return rc;
    IHG_REQUIRE(def_ld % 4 == 0, "edge_interact_bwd: def_ld must be a multiple of 4");
    IHG_REQUIRE(workspace_bytes >= ihg_edge_interact_bwd_workspace_bytes(dim, order),
                "edge_interact_bwd: workspace too small");
    const int nb = order == 3 ? 4 : 3;
    cudaStream_t st = as_stream(stream);
    if (E == 0) {
        IHG_CUDA(cudaMemsetAsync(dw_hi, 0, (size_t)nb * dim * dim * 4, st));
        return IHG_OK;
    }
    if (interact_tc_eligible(dim) && xp_ld % 4 == 0)
        return launch_interact_bwd_tc(xp, xp_ld, def, def_ld, w_hi, w_ld, nb, i3, E, slot_grad, dw_hi, dim, workspace, st);
    const unsigned blocks = (unsigned)ceil_div(E, kTileRows);
    // (a) slot gradients; column groups of at most 64 features per pass
#define IHG_IB_CASE(D) edge_interact_bwd_slot_kernel<D><<<blocks, kGemmThreads, 0, st>>>(xp, xp_ld, def, def_ld, w_hi, w_ld, nb, i3, E, slot_grad, dim)
    if (dim <= 16) IHG_IB_CASE(1);
    else if (dim <= 32) IHG_IB_CASE(2);
    else IHG_IB_CASE(4);
#undef IHG_IB_CASE
    IHG_LAUNCH_CHECK();
    // (b) weight gradient
    float* ws_dw = static_cast<float*>(workspace);
    dim3 grid(kInteractWgradG, nb);
#define IHG_IW_CASE(RG, D)                                                                          \
    do {                                                                                            \
        const size_t smem = (size_t)kTileRows * ((RG * 64 + 4) + (16 * D + 1)) * sizeof(float);     \
        IHG_CUDA(cudaFuncSetAttribute(edge_interact_bwd_wgrad_kernel<RG, D>,                        \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        edge_interact_bwd_wgrad_kernel<RG, D><<<grid, kGemmThreads, smem, st>>>(xp, xp_ld, def, def_ld, \
                                                                                i3, E, dim, ws_dw); \
    } while (0)
    if (dim <= 16) IHG_IW_CASE(1, 1);
    else if (dim <= 32) IHG_IW_CASE(1, 2);
    else if (dim <= 64) IHG_IW_CASE(1, 4);
    else IHG_IW_CASE(2, 8);
#undef IHG_IW_CASE
    IHG_LAUNCH_CHECK();
    const int64_t total = (int64_t)nb * dim * dim;
    interact_wgrad_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(ws_dw, kInteractWgradG, nb, dim, dw_hi);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // extern "C"
