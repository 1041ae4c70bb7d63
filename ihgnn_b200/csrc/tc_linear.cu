// Node-level typed Linear on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32
// split precision, fp32 accumulators in TMEM).  Same contract as the SIMT kernel in
// node_linear.cu -- y[r,:] = x[r,:] . W[t]^T (+bias[t]) (+addend[r,:]) -- selected when
// n_in % 32 == 0 and n_out % 16 == 0.
//
// One CTA of 128 threads owns one tile of 128 rows (tiles never straddle a node-type
// boundary).  Thread t stages row t: a 128-byte slice of x per K chunk is split into tf32
// hi/lo and written into the UMMA K-major SWIZZLE_128B layout; the same threads stage the
// weight rows.  One elected thread issues the MMAs; completion is tracked with an mbarrier via
// tcgen05.commit; the epilogue reads the accumulator row of each thread with tcgen05.ld.
// Several CTAs are resident per SM, so staging, MMA and epilogue of different tiles overlap.
//
// Roofline: HBM (4*(n_in+n_out) bytes per row vs 6*n_in*n_out tf32 flops per row).
#include "tc_common.cuh"
#include "tc_linear.h"

namespace ihg {

using namespace tc;

struct TcTypeTiles {
    int64_t b0, b1, n_rows;
    __host__ __device__ int64_t lo(int t) const { return t == 0 ? 0 : (t == 1 ? b0 : b1); }
    __host__ __device__ int64_t hi(int t) const { return t == 0 ? b0 : (t == 1 ? b1 : n_rows); }
    __host__ __device__ int64_t tiles(int t) const { return (hi(t) - lo(t) + kTileM - 1) / kTileM; }
};

__global__ void __launch_bounds__(128)
node_linear_tc_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ w,
                      int n_types, int n_out, int n_in, int transpose_w,
                      const float* __restrict__ bias, const float* __restrict__ addend,
                      int64_t addend_ld, TcTypeTiles tt, float* __restrict__ y, int64_t y_ld) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned operand tiles: A hi, A lo (128 rows), B hi, B lo (n_out rows)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768, b_lo = base + 49152;
    __shared__ __align__(8) uint64_t mbar_storage;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t mbar = smem_u32(&mbar_storage);
    const int tid = threadIdx.x, warp = tid >> 5;

    int64_t bid = blockIdx.x;
    int type = 0;
    while (type < 2 && bid >= tt.tiles(type)) { bid -= tt.tiles(type); ++type; }
    const int64_t row0 = tt.lo(type) + bid * kTileM;
    const int rows = (int)min((int64_t)kTileM, tt.hi(type) - row0);
    const int wt = n_types > 1 ? type : 0;
    const float* W = w + (int64_t)wt * n_out * n_in;

    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)n_out);
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init_fence();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = make_idesc_tf32(n_out);

    uint32_t phase = 0;
    const int n_chunks = n_in / kChunkK;
    for (int kc = 0; kc < n_chunks; ++kc) {
        const int k0 = kc * kChunkK;
        // ---- stage A: row `tid` of the tile, 32 floats
        {
            const bool ok = tid < rows;
            const float* src = x + (row0 + tid) * x_ld + k0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = ok ? ldg4(src + 4 * c) : f4_zero();
                store_split_chunk(a_hi, a_lo, tid, c, v);
            }
        }
        // ---- stage B: weight row n = tid (B[n][k] = W[n][k] or W[k][n])
        if (tid < n_out) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v;
                if (!transpose_w) {
                    v = ldg4(W + (int64_t)tid * n_in + k0 + 4 * c);
                } else {
                    const float* p = W + (int64_t)(k0 + 4 * c) * n_out + tid;
                    v = make_float4(__ldg(p), __ldg(p + n_out), __ldg(p + 2 * n_out), __ldg(p + 3 * n_out));
                }
                store_split_chunk(b_hi, b_lo, tid, c, v);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint64_t dah = make_kmajor_sw128_desc(a_hi), dal = make_kmajor_sw128_desc(a_lo);
            const uint64_t dbh = make_kmajor_sw128_desc(b_hi), dbl = make_kmajor_sw128_desc(b_lo);
#pragma unroll
            for (int ks = 0; ks < kChunkK / 8; ++ks) {
                mma_3xtf32(tmem_d, advance_desc_k(dah, 8 * ks), advance_desc_k(dal, 8 * ks),
                           advance_desc_k(dbh, 8 * ks), advance_desc_k(dbl, 8 * ks), idesc,
                           (kc > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(mbar);
        }
        // the operand tiles are reused by the next chunk: wait until the MMAs have read them
        mbar_wait(mbar, phase);
        phase ^= 1u;
    }
    fence_after_sync();
    // ---- epilogue: thread t owns accumulator row t (TMEM lane t)
    const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
    const bool ok = tid < rows;
    const int64_t row = row0 + tid;
    for (int c0 = 0; c0 < n_out; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)c0, v);     // warp-collective: all lanes participate
        if (ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                if (bias) f4_add(o, ldg4(bias + (int64_t)wt * n_out + c0 + 4 * q));
                if (addend) f4_add(o, ldg4(addend + row * addend_ld + c0 + 4 * q));
                stg4(y + row * y_ld + c0 + 4 * q, o);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

bool node_linear_tc_eligible(int n_out, int n_in, int64_t x_ld, int64_t y_ld, const float* addend,
                             int64_t addend_ld) {
    static const bool disabled = getenv("IHG_DISABLE_TC") != nullptr;
    if (disabled) return false;
    if (n_in % 32 != 0 || n_out % 16 != 0 || n_out < 16 || n_out > 128 || n_in > 128) return false;
    if (x_ld % 4 != 0 || y_ld % 4 != 0) return false;
    if (addend && addend_ld % 4 != 0) return false;
    return true;
}

int launch_node_linear_tc(const float* x, int64_t x_ld, const float* w, int n_types, int n_out,
                          int n_in, int transpose_w, const float* bias, const float* addend,
                          int64_t addend_ld, int64_t n_rows, int64_t bound0, int64_t bound1,
                          float* y, int64_t y_ld, cudaStream_t st) {
    TcTypeTiles tt{bound0, bound1, n_rows};
    const int64_t blocks = tt.tiles(0) + tt.tiles(1) + tt.tiles(2);
    const int smem = 4 * 16384 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        IHG_CUDA(cudaFuncSetAttribute(node_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    node_linear_tc_kernel<<<(unsigned)blocks, 128, smem, st>>>(x, x_ld, w, n_types, n_out, n_in, transpose_w,
                                                               bias, addend, addend_ld, tt, y, y_ld);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // namespace ihg
