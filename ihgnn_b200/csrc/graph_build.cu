// Device-resident index construction for the 3-uniform hypergraph.
//
// Replaces PpsHyperGraph.from_interactions (/root/reference/Helpers/Graph.py:94-134: a Python
// loop with one device fancy-index `+=` per edge, then a COO sort inside coalesce()) by
//   histogram (integer atomics)  ->  exclusive scan  ->  stable LSD radix sort per slot.
// Users only ever occupy slot 0, queries slot 1, items slot 2 (Graph.py:110-117), so the CSR
// column array is simply [edges sorted by user | edges sorted by query | edges sorted by item],
// each sort stable in the edge id -- exactly the (node, edge) lexicographic order coalesce()
// produces.  Everything is integer work: results are bit-exact and run-to-run deterministic.
//
// All kernels here are HBM/latency-bound one-time setup work (12E bytes in, ~28E+8N out).
#include <stdarg.h>

#include <atomic>
#include <string>

#include "common.cuh"

namespace ihg {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// =========================================================================================
// exclusive scan of int32 (multi-level, in-place capable)
// =========================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const int32_t* in, int32_t* out, int64_t n,  // in == out allowed
                 int32_t* __restrict__ tile_sums) {
    __shared__ int32_t warp_tot[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int32_t local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : 0;
        local += v[k];
    }
    // inclusive scan of `local` across the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int32_t warp_base = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w)
        if (w < warp) warp_base += warp_tot[w];
    int32_t run = warp_base + incl - local;  // exclusive prefix of this thread
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    if (threadIdx.x == kScanThreads - 1 && tile_sums) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ tile_offsets) {
    const int32_t off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        if (i < n) data[i] += off;
    }
}

static int64_t scan_scratch_elems(int64_t n) {
    int64_t total = 0;
    while (n > kScanTile) {
        n = ceil_div(n, kScanTile);
        total += align_up(n, 64);
    }
    return total + 64;
}

// out[i] = sum_{j<i} in[j]; in == out allowed.  scratch: scan_scratch_elems(n) int32.
static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch,
                          cudaStream_t st) {
    if (n <= 0) return IHG_OK;
    int64_t tiles = ceil_div(n, kScanTile);
    if (tiles == 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
        IHG_LAUNCH_CHECK();
        return IHG_OK;
    }
    int32_t* sums = scratch;
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums);
    IHG_LAUNCH_CHECK();
    int rc = exclusive_scan(sums, sums, tiles, scratch + align_up(tiles, 64), st);
    if (rc) return rc;
    scan_add_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(out, n, sums);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

// =========================================================================================
// stable LSD radix sort of (key, position) pairs, 8 bits per pass
// =========================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;  // 4096

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift,
                  int32_t* __restrict__ block_hist, int nblocks) {
    __shared__ int32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortRounds; ++r) {
        int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255], 1);
    }
    __syncthreads();
    block_hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                     int64_t n, int shift, const int32_t* __restrict__ block_off, int nblocks,
                     int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
    __shared__ int32_t warp_hist[kSortThreads / 32][256];
    __shared__ int32_t running[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    running[tid] = block_off[(int64_t)tid * nblocks + blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t i = base + r * kSortThreads + tid;
        if (base + (int64_t)r * kSortThreads >= n) break;  // block-uniform
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) warp_hist[w][tid] = 0;
        __syncthreads();
        const bool valid = i < n;
        int32_t key = 0, val = 0;
        int digit = 0x1000;  // sentinel groups the out-of-range lanes together
        if (valid) {
            key = keys[i];
            val = vals ? vals[i] : (int32_t)i;
            digit = (key >> shift) & 255;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & lt_mask);
        if (valid && rank == 0) warp_hist[warp][digit] = __popc(peers);
        __syncthreads();
        {   // thread `tid` owns digit `tid`: turn per-warp counts into per-warp start offsets
            int32_t run = running[tid];
#pragma unroll
            for (int w = 0; w < kSortThreads / 32; ++w) {
                int32_t c = warp_hist[w][tid];
                warp_hist[w][tid] = run;
                run += c;
            }
            running[tid] = run;
        }
        __syncthreads();
        if (valid) {
            const int32_t dst = warp_hist[warp][digit] + rank;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
    }
}

static int key_bits(int64_t num_keys) {
    int bits = 1;
    while (((int64_t)1 << bits) < num_keys) ++bits;
    return bits;
}
static int sort_passes(int64_t num_keys) { return (int)ceil_div(key_bits(num_keys), 8); }

static int64_t sort_ws_bytes(int64_t n) {
    int64_t nb = ceil_div(n > 0 ? n : 1, kSortTile);
    return 4 * ws_slice(n, 4) + ws_slice(256 * nb + 1, 4) + ws_slice(scan_scratch_elems(256 * nb + 1), 4);
}

// perm_out[j] = position of the j-th smallest key (stable).  keys are not modified.
static int radix_sort_perm(const int32_t* keys, int64_t n, int64_t num_keys, int32_t* perm_out,
                           Workspace& ws, cudaStream_t st) {
    if (n <= 0) return IHG_OK;
    const int nb = (int)ceil_div(n, kSortTile);
    int32_t* kbuf[2] = {ws.take<int32_t>(n), ws.take<int32_t>(n)};
    int32_t* vbuf[2] = {ws.take<int32_t>(n), ws.take<int32_t>(n)};
    int32_t* hist = ws.take<int32_t>(256 * (int64_t)nb + 1);
    int32_t* scratch = ws.take<int32_t>(scan_scratch_elems(256 * (int64_t)nb + 1));
    if (!ws.ok()) {
        set_error("radix sort: workspace too small (%lld needed, %lld given)", (long long)ws.used,
                  (long long)ws.size);
        return IHG_ERR_WORKSPACE;
    }
    const int passes = sort_passes(num_keys);
    const int32_t* kin = keys;
    const int32_t* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        int32_t* kout = kbuf[p & 1];
        int32_t* vout = (p == passes - 1) ? perm_out : vbuf[p & 1];
        radix_hist_kernel<<<nb, kSortThreads, 0, st>>>(kin, n, shift, hist, nb);
        IHG_LAUNCH_CHECK();
        int rc = exclusive_scan(hist, hist, 256 * (int64_t)nb, scratch, st);
        if (rc) return rc;
        radix_scatter_kernel<<<nb, kSortThreads, 0, st>>>(kin, vin, n, shift, hist, nb, kout, vout);
        IHG_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
    }
    return IHG_OK;
}

// =========================================================================================
// hypergraph build
// =========================================================================================
__global__ void __launch_bounds__(256)
make_i3_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ query,
               const int64_t* __restrict__ item, int64_t E, int64_t U, int64_t Q, int64_t I,
               int32_t* __restrict__ i3, int32_t* __restrict__ key_u, int32_t* __restrict__ key_q,
               int32_t* __restrict__ key_i, int32_t* __restrict__ counts,
               int32_t* __restrict__ error_flag) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t u = user[e], q = query[e], i = item[e];
        if (u < 0 || u >= U || q < 0 || q >= Q || i < 0 || i >= I) {
            atomicOr(error_flag, 1);
            u = 0; q = 0; i = 0;
        }
        const int32_t gu = (int32_t)u, gq = (int32_t)(q + U), gi = (int32_t)(i + U + Q);
        i3[3 * e + 0] = gu;
        i3[3 * e + 1] = gq;
        i3[3 * e + 2] = gi;
        key_u[e] = (int32_t)u;
        key_q[e] = (int32_t)q;
        key_i[e] = (int32_t)i;
        atomicAdd(&counts[gu], 1);  // integer atomics: order-independent result
        atomicAdd(&counts[gq], 1);
        atomicAdd(&counts[gi], 1);
    }
}

__global__ void __launch_bounds__(256)
degree_kernel(const int32_t* __restrict__ counts, int64_t N, float* __restrict__ deg,
              float* __restrict__ dv_inv, float* __restrict__ dv_inv_sqrt) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < N;
         v += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = counts[v];
        const float d = c > 0 ? (float)c : 1e-8f;             // Graph.py:120
        if (deg) deg[v] = d;
        if (dv_inv) dv_inv[v] = __fdiv_rn(1.0f, d);            // VertexDegrees.pow(-1)
        if (dv_inv_sqrt) dv_inv_sqrt[v] = __fdiv_rn(1.0f, __fsqrt_rn(d));  // .pow(-0.5)
    }
}

__global__ void __launch_bounds__(256)
count_keys_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t num_keys,
                  int32_t* __restrict__ counts, int32_t* __restrict__ error_flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t k = keys[i];
        if (k < 0 || k >= num_keys) {
            atomicOr(error_flag, 1);
        } else {
            atomicAdd(&counts[k], 1);
        }
    }
}

__global__ void __launch_bounds__(256)
permute_values_kernel(const int32_t* __restrict__ values, const int32_t* __restrict__ perm,
                      int64_t n, int32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = values[perm[i]];
}

static unsigned grid_for(int64_t n, int threads = 256, int max_blocks = kNumSMs * 16) {
    int64_t b = ceil_div(n > 0 ? n : 1, threads);
    return (unsigned)(b < max_blocks ? b : max_blocks);
}

// =========================================================================================
// segment plan
// =========================================================================================
__global__ void __launch_bounds__(256)
plan_count_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk_len,
                  int32_t* __restrict__ nseg, int32_t* __restrict__ nsplit,
                  int32_t* __restrict__ npart) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        int32_t a = 0, b = 0, c = 0;
        if (r < n_rows) {
            const int32_t deg = rowptr[r + 1] - rowptr[r];
            a = deg > chunk_len ? (deg + chunk_len - 1) / chunk_len : 1;
            b = a > 1;
            c = b ? a : 0;
        }
        nseg[r] = a;  // element n_rows is the 0 that turns the scan's last slot into the total
        nsplit[r] = b;
        npart[r] = c;
    }
}

__global__ void __launch_bounds__(256)
plan_fill_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk_len,
                 const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ split_idx,
                 const int32_t* __restrict__ part_ptr, int4* __restrict__ seg,
                 int32_t* __restrict__ split_row, int32_t* __restrict__ split_ptr,
                 int64_t* __restrict__ counts) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s0 = seg_ptr[r];
        const int32_t nch = seg_ptr[r + 1] - s0;
        const int32_t p0 = part_ptr[r];
        const int32_t begin = rowptr[r], row_end = rowptr[r + 1];
        const bool split = nch > 1;
        for (int32_t c = 0; c < nch; ++c) {
            const int32_t b = begin + c * chunk_len;
            seg[s0 + c] = make_int4(b, min(b + chunk_len, row_end), (int32_t)r, split ? p0 + c : -1);
        }
        if (split) {
            split_row[split_idx[r]] = (int32_t)r;
            split_ptr[split_idx[r]] = p0;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[0] = seg_ptr[n_rows];
        counts[1] = split_idx[n_rows];
        counts[2] = part_ptr[n_rows];
        split_ptr[split_idx[n_rows]] = part_ptr[n_rows];
    }
}

}  // namespace ihg

using namespace ihg;

extern "C" {

int ihg_abi_version(void) { return IHG_ABI_VERSION; }
const char* ihg_last_error(void) { return g_last_error.c_str(); }
int64_t ihg_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int64_t ihg_graph_workspace_bytes(int64_t E, int64_t N) {
    if (E < 0) E = 0;
    return 3 * ws_slice(E, 4) + ws_slice(N + 1, 4) + ws_slice(scan_scratch_elems(N + 1), 4) +
           sort_ws_bytes(E) + 4096;
}

int ihg_graph_build(const int64_t* user, const int64_t* query, const int64_t* item, int64_t E,
                    int64_t U, int64_t Q, int64_t I, int32_t* i3, int32_t* rowptr, int32_t* col,
                    float* vertex_degrees, float* dv_inv, float* dv_inv_sqrt, int32_t* error_flag,
                    void* workspace, int64_t workspace_bytes, void* stream) {
    const int64_t N = U + Q + I;
    IHG_REQUIRE(E >= 0 && U > 0 && Q > 0 && I > 0, "graph_build: bad counts E=%lld U=%lld Q=%lld I=%lld",
                (long long)E, (long long)U, (long long)Q, (long long)I);
    IHG_REQUIRE(3 * E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX,
                "graph_build: 3E=%lld or N=%lld exceeds int32 indexing", (long long)(3 * E), (long long)N);
    IHG_REQUIRE(rowptr && error_flag && (E == 0 || (user && query && item && i3 && col)),
                "graph_build: null pointer");
    IHG_REQUIRE(workspace_bytes >= ihg_graph_workspace_bytes(E, N), "graph_build: workspace too small");
    cudaStream_t st = as_stream(stream);
    Workspace ws(workspace, workspace_bytes);
    int32_t* key_u = ws.take<int32_t>(E);
    int32_t* key_q = ws.take<int32_t>(E);
    int32_t* key_i = ws.take<int32_t>(E);
    int32_t* counts = ws.take<int32_t>(N + 1);
    int32_t* scratch = ws.take<int32_t>(scan_scratch_elems(N + 1));
    IHG_CUDA(cudaMemsetAsync(counts, 0, (size_t)(N + 1) * 4, st));
    IHG_CUDA(cudaMemsetAsync(error_flag, 0, 4, st));
    if (E > 0) {
        make_i3_kernel<<<grid_for(E), 256, 0, st>>>(user, query, item, E, U, Q, I, i3, key_u, key_q,
                                                    key_i, counts, error_flag);
        IHG_LAUNCH_CHECK();
    }
    degree_kernel<<<grid_for(N), 256, 0, st>>>(counts, N, vertex_degrees, dv_inv, dv_inv_sqrt);
    IHG_LAUNCH_CHECK();
    int rc = exclusive_scan(counts, rowptr, N + 1, scratch, st);
    if (rc) return rc;
    const int32_t* keys[3] = {key_u, key_q, key_i};
    const int64_t nk[3] = {U, Q, I};
    for (int s = 0; s < 3 && E > 0; ++s) {
        Workspace sub(ws.base + ws.used, ws.size - ws.used);  // the sort scratch is reused per slot
        rc = radix_sort_perm(keys[s], E, nk[s], col + (int64_t)s * E, sub, st);
        if (rc) return rc;
    }
    return IHG_OK;
}

int64_t ihg_csr_from_keys_workspace_bytes(int64_t n, int64_t num_keys) {
    if (n < 0) n = 0;
    return ws_slice(num_keys + 1, 4) + ws_slice(scan_scratch_elems(num_keys + 1), 4) + sort_ws_bytes(n) + 4096;
}

int ihg_csr_from_keys(const int32_t* keys, const int32_t* values, int64_t n, int64_t num_keys,
                      int32_t* rowptr, int32_t* perm, int32_t* out_values, int32_t* error_flag,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(n >= 0 && num_keys > 0 && n < (int64_t)INT32_MAX && num_keys < (int64_t)INT32_MAX,
                "csr_from_keys: bad sizes n=%lld num_keys=%lld", (long long)n, (long long)num_keys);
    IHG_REQUIRE(rowptr && error_flag && (n == 0 || (keys && perm)), "csr_from_keys: null pointer");
    IHG_REQUIRE(workspace_bytes >= ihg_csr_from_keys_workspace_bytes(n, num_keys),
                "csr_from_keys: workspace too small");
    cudaStream_t st = as_stream(stream);
    Workspace ws(workspace, workspace_bytes);
    int32_t* counts = ws.take<int32_t>(num_keys + 1);
    int32_t* scratch = ws.take<int32_t>(scan_scratch_elems(num_keys + 1));
    IHG_CUDA(cudaMemsetAsync(counts, 0, (size_t)(num_keys + 1) * 4, st));
    IHG_CUDA(cudaMemsetAsync(error_flag, 0, 4, st));
    if (n > 0) {
        count_keys_kernel<<<grid_for(n), 256, 0, st>>>(keys, n, num_keys, counts, error_flag);
        IHG_LAUNCH_CHECK();
    }
    int rc = exclusive_scan(counts, rowptr, num_keys + 1, scratch, st);
    if (rc) return rc;
    if (n > 0) {
        Workspace sub(ws.base + ws.used, ws.size - ws.used);
        rc = radix_sort_perm(keys, n, num_keys, perm, sub, st);
        if (rc) return rc;
        if (values && out_values) {
            permute_values_kernel<<<grid_for(n), 256, 0, st>>>(values, perm, n, out_values);
            IHG_LAUNCH_CHECK();
        }
    }
    return IHG_OK;
}

int64_t ihg_segment_plan_workspace_bytes(int64_t n_rows) {
    return 3 * ws_slice(n_rows + 1, 4) + ws_slice(scan_scratch_elems(n_rows + 1), 4) + 4096;
}

int ihg_segment_plan_build(const int32_t* rowptr, int64_t n_rows, int32_t chunk_len,
                           int32_t* seg, int32_t* split_row, int32_t* split_ptr, int64_t* counts,
                           void* workspace, int64_t workspace_bytes, void* stream) {
    IHG_REQUIRE(n_rows > 0 && chunk_len >= 32 && n_rows < (int64_t)INT32_MAX,
                "segment_plan: bad arguments n_rows=%lld chunk_len=%d", (long long)n_rows, chunk_len);
    IHG_REQUIRE(rowptr && seg && split_row && split_ptr && counts, "segment_plan: null pointer");
    IHG_REQUIRE((reinterpret_cast<uintptr_t>(seg) & 15) == 0, "segment_plan: seg must be 16-byte aligned");
    IHG_REQUIRE(workspace_bytes >= ihg_segment_plan_workspace_bytes(n_rows),
                "segment_plan: workspace too small");
    cudaStream_t st = as_stream(stream);
    Workspace ws(workspace, workspace_bytes);
    int32_t* nseg = ws.take<int32_t>(n_rows + 1);
    int32_t* nsplit = ws.take<int32_t>(n_rows + 1);
    int32_t* npart = ws.take<int32_t>(n_rows + 1);
    int32_t* scratch = ws.take<int32_t>(scan_scratch_elems(n_rows + 1));
    plan_count_kernel<<<grid_for(n_rows + 1), 256, 0, st>>>(rowptr, n_rows, chunk_len, nseg, nsplit, npart);
    IHG_LAUNCH_CHECK();
    int rc;
    if ((rc = exclusive_scan(nseg, nseg, n_rows + 1, scratch, st))) return rc;
    if ((rc = exclusive_scan(nsplit, nsplit, n_rows + 1, scratch, st))) return rc;
    if ((rc = exclusive_scan(npart, npart, n_rows + 1, scratch, st))) return rc;
    plan_fill_kernel<<<grid_for(n_rows), 256, 0, st>>>(rowptr, n_rows, chunk_len, nseg, nsplit, npart,
                                                       reinterpret_cast<int4*>(seg), split_row,
                                                       split_ptr, counts);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}

}  // extern "C"
