/*
 * ihgnn_b200.h -- C ABI of libihgnn_b200.so: the B200 (sm_100a) implementation of IHGNN's
 * interactive hypergraph-convolution hot path.
 *
 * Conventions (last row):
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - the caller owns every buffer, the library allocates nothing and keeps no state
 *     (workspaces are caller-provided, their sizes come from the *_workspace_bytes calls);
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*); there is no
 *     hidden synchronisation, so the calls can be captured in a CUDA graph;
 *   - fp32 values, row-major, leading dimensions (`*_ld`) in elements; int32 indices inside
 *     the library, int64 where the reference hands int64 tensors in (graph input, batch
 *     indices);
 *   - return value 0 = ok, non-zero = error (IHG_ERR_*), text via ihg_last_error();
 *   - results are deterministic: no floating-point atomics anywhere.
 *
 * The reference is pure PyTorch, so it has no FFI of its own; each entry point below names
 * the reference Python code (file:line under /root/reference) whose computation it replaces.
 * The reference-side binding (a ctypes stub + autograd.Function) is shown in INTEGRATION.md
 * and implemented in ihgnn_b200/_lib.py and ihgnn_b200/functional.py.
 */
#ifndef IHGNN_B200_H
#define IHGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IHG_ABI_VERSION 4   /* 4: ihg_segment_reduce_routed / ihg_two_hop_reduce_routed; 3: ihg_adam_step; 2: ihg_segment_reduce gained `flags`, two-hop / ranking / sampler */

#define IHG_OK 0
#define IHG_ERR_INVALID_ARGUMENT 1   /* bad shape / null pointer / unsupported dimension   */
#define IHG_ERR_CUDA 2               /* a CUDA runtime call or launch failed              */
#define IHG_ERR_WORKSPACE 3          /* caller workspace too small                        */

int ihg_abi_version(void);
/* Thread-local text of the last error raised on the calling thread ("" if none). */
const char* ihg_last_error(void);
/* Number of CUDA kernels this library has launched in the process so far (all threads). */
int64_t ihg_launch_count(void);

/* Host-side struct describing a CSR matrix with unit values plus its deterministic
 * load-balancing plan (rows longer than `chunk_len` are split into chunks whose partial
 * sums are combined in fixed order).  All pointers are device pointers. */
typedef struct ihg_csr {
    int64_t n_rows;
    int64_t nnz;
    const int32_t* rowptr;     /* [n_rows+1]                                             */
    const int32_t* col;        /* [nnz]                                                  */
    int32_t chunk_len;
    int64_t n_seg;             /* work items: one per row chunk (>= n_rows)              */
    int64_t n_split;           /* rows that were split into more than one chunk          */
    int64_t n_part;            /* partial-sum rows needed by split rows                  */
    const int32_t* seg;        /* [n_seg][4] = {begin, end, row, partial slot or -1}     */
    const int32_t* split_row;  /* [n_split]                                              */
    const int32_t* split_ptr;  /* [n_split+1] range of partial slots of each split row   */
} ihg_csr;
/* `seg` may list the work items of a subset of the rows: the reductions then touch only those rows of `out`. */

/* ------------------------------------------------------------------------------------
 * a1  PpsHyperGraph.from_interactions            Helpers/Graph.py:94-134
 *     (+ the coalesce() of GnnLayers.py:190).  Replaces the per-edge Python loop and the
 *     COO sort by a device histogram + scan + stable LSD radix sort.
 * in : user/query/item  int64 [E], 0-based per-type ids, in file (= hyperedge) order
 * out: i3      int32 [E,3]   global node ids (u, q+U, i+U+Q)            Graph.py:110-117,129
 *      rowptr  int32 [N+1], col int32 [3E]: CSR of Adjacency.coalesce() (edge ids ascending
 *              inside a node)                                            Graph.py:123-128
 *      vertex_degrees fp32 [N] with 0 stored as 1e-8                     Graph.py:112,120
 *      dv_inv = vertex_degrees^-1 (GnnLayers.py:187), dv_inv_sqrt = ^-1/2 (GnnLayers.py:133)
 *      error_flag int32 [1]: set non-zero if any id is out of range
 * ------------------------------------------------------------------------------------ */
int64_t ihg_graph_workspace_bytes(int64_t edge_count, int64_t node_count);
int ihg_graph_build(const int64_t* user, const int64_t* query, const int64_t* item,
                    int64_t edge_count, int64_t user_count, int64_t query_count,
                    int64_t item_count, int32_t* i3, int32_t* rowptr, int32_t* col,
                    float* vertex_degrees, float* dv_inv, float* dv_inv_sqrt,
                    int32_t* error_flag, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Stable CSR of a key array: rowptr[k] = #keys < k, perm = positions sorted by key with
 * ascending position inside a key.  Used for the vocabulary->query transpose that the
 * EmbeddingBag backward needs (EmbeddingLayers.py:79; torch _embedding_bag_dense_backward).
 * `values` (nullable, int32 [n]) is permuted along: out_values[j] = values[perm[j]]. */
int64_t ihg_csr_from_keys_workspace_bytes(int64_t n, int64_t num_keys);
int ihg_csr_from_keys(const int32_t* keys, const int32_t* values, int64_t n, int64_t num_keys,
                      int32_t* rowptr, int32_t* perm, int32_t* out_values,
                      int32_t* error_flag, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* Load-balancing plan for ihg_segment_reduce.  Capacities the caller must provide:
 *   seg        : (n_rows + nnz/chunk_len + 1) entries of 4 x int32 (16-byte aligned)
 *   split_row  : nnz/chunk_len + 1,  split_ptr: nnz/chunk_len + 2
 * counts (device int64 [3]) receives n_seg, n_split, n_part. */
int64_t ihg_segment_plan_workspace_bytes(int64_t n_rows);
int ihg_segment_plan_build(const int32_t* rowptr, int64_t n_rows, int32_t chunk_len,
                           int32_t* seg, int32_t* split_row, int32_t* split_ptr, int64_t* counts,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * a7  torch_sparse.matmul(incidence, Ef) * Dv^-1     Models/GnnLayers.py:233-234
 *     (and its transpose products in backward; EmbeddingBag(mean) EmbeddingLayers.py:79).
 * Deterministic segmented reduction over CSR rows:
 *   out[r, 0:dim] = row_scale[r] * sum_{j in row r, ascending}
 *                       src_scale[s(j)] * src[s(j), 0:dim],
 *   s(j) = col[j] * src_row_mul + slot(r),  slot(r) = row_slot[r] if row_slot is given, else
 *   (r >= bound0) + (r >= bound1)
 * (src_row_mul = 3 with the node-type bounds reads the per-slot gradient rows [E,3,dim];
 * src_row_mul = 1 with bounds = INT64_MAX is the plain SpMM).  `init` (nullable, [n_rows, dim],
 * row stride init_ld) is added first: out = row_scale * (init[r] + sum ...) -- the multi-GPU
 * reduce adds the rank's own partial before the received ones.  row_scale / src_scale may be
 * null (= 1).  `partial` is scratch of csr->n_part * dim floats.  dim % 4 == 0, dim <= 256.
 * flags & IHG_SEG_ACCUMULATE: out[r] = init[r] + row_scale[r] * sum instead, rows without
 * incidences are left untouched (the plan may then list the non-empty rows only), and init may
 * alias out -- used to run one reduction as a few passes over L2-sized hyperedge ranges (each pass
 * a CSR restricted to the range), so that the three reads of a hyperedge row (one per slot) hit L2
 * after the first.  flags & IHG_SEG_L2_SOURCE: the source rows are expected in L2 -- launch the
 * high-occupancy variant (DRAM-served random rows want the opposite).
 * ------------------------------------------------------------------------------------ */
#define IHG_SEG_ACCUMULATE 1
#define IHG_SEG_L2_SOURCE 2
int ihg_segment_reduce(const ihg_csr* csr_host, const float* src, int64_t src_ld,
                       int32_t src_row_mul, int64_t bound0, int64_t bound1,
                       const int32_t* row_slot, const float* init, int64_t init_ld,
                       const float* src_scale, const float* row_scale, float* partial,
                       float* out, int64_t out_ld, int32_t dim, int32_t flags, void* stream);

/* ------------------------------------------------------------------------------------
 * a5-a8  node -> hyperedge -> node round trip of an order-1 layer in one pass (no [E,dim]
 *     intermediate):   Models/CommonLayers.py:58-66 (order-1 FeatureInteractor, aggregation
 *     Linear hoisted to node level) followed by thsp.matmul(incidence, ef) * Dv^-1
 *     (GnnLayers.py:233-234); HGCN's H De^-1 H^T (GnnLayers.py:148-151); and their backward
 *     products (H H^T is symmetric).
 *   out[r,:] = row_scale[r] * alpha * sum_{e contains r} sum_{n in e} node_scale[n] * src[n,:]
 * ihg_two_hop_index_build fills nbr int32 [nnz,2]: for incidence j of row r (slot(r) as in
 * ihg_segment_reduce) the two OTHER nodes of hyperedge col[j], slots (slot+1)%3 and (slot+2)%3.
 * The row's own term is (own_per_incidence * deg(r) + own_const) * node_scale[r] * src[r]:
 * (1, 0) is the hypergraph round trip above; (0, 0) / (0, 1) is the pairwise adjacency of
 * Pps2DGraph (Helpers/Graph.py:19-81, u-q, q-i, i-u both ways per interaction, duplicates summed
 * as coalesce() does) without / with self connections, i.e. GCNLayer's D^-1/2 A D^-1/2 product
 * (Models/GnnLayers.py:33-43).  Deterministic (same chunk plan and fix-up as ihg_segment_reduce);
 * node_scale / row_scale nullable; `partial` as there.
 * ------------------------------------------------------------------------------------ */
int ihg_two_hop_index_build(const ihg_csr* csr_host, const int32_t* i3, int64_t bound0,
                            int64_t bound1, const int32_t* row_slot, int32_t* nbr, void* stream);
int ihg_two_hop_reduce(const ihg_csr* csr_host, const int32_t* nbr, const float* src,
                       int64_t src_ld, const float* node_scale, float alpha,
                       float own_per_incidence, float own_const, const float* row_scale,
                       float* partial, float* out, int64_t out_ld, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------
 * (e) multi-GPU reduce-scatter of node partial sums (SURVEY.md section 8e "Collectives per layer":
 *     local partials over owned + halo rows, then the cross-GPU reduce), fused into the producer:
 * the same reductions as ihg_segment_reduce / ihg_two_hop_reduce (no init / row scale / flags), but
 * consecutive row ranges of the result go to different destinations: range k = rows
 * [route_start[k], route_start[k+1]) is written to route_base[k] + (row - route_start[k]) * out_ld.
 * The sharded layers route the own rows to a local matrix and every peer's halo rows straight into
 * that peer's receive buffer over NVLink (peer-mapped pointers): the partial sums travel as posted
 * stores while the kernel is still gathering, and the owner finishes with a local ordered sum
 * (ihg_segment_reduce over its receive buffer).  route_start (n_route + 1 entries, route_start[0] == 0,
 * route_start[n_route] == csr->n_rows, ascending) and route_base (n_route entries) are HOST arrays;
 * 1 <= n_route <= 16.  Deterministic like the un-routed calls.
 * ------------------------------------------------------------------------------------ */
int ihg_segment_reduce_routed(const ihg_csr* csr_host, const float* src, int64_t src_ld,
                              int32_t src_row_mul, int64_t bound0, int64_t bound1,
                              const int32_t* row_slot, float* partial,
                              const int64_t* route_start, void* const* route_base, int32_t n_route,
                              int64_t out_ld, int32_t dim, void* stream);
int ihg_two_hop_reduce_routed(const ihg_csr* csr_host, const int32_t* nbr, const float* src,
                              int64_t src_ld, const float* node_scale, float alpha,
                              float own_per_incidence, float own_const, float* partial,
                              const int64_t* route_start, void* const* route_base, int32_t n_route,
                              int64_t out_ld, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------
 * a6/a8  node -> hyperedge gather-sum.
 *   out[e,:] = alpha * sum_{s<3} node_scale[i3[e,s]] * src[i3[e,s], :]  (+ bias)
 * Order-1 FeatureInteractor after hoisting the aggregation Linear to node level
 * (CommonLayers.py:58-66), HGCN's incidence_t SpMM (GnnLayers.py:148-149), and the
 * backward of the edge->node SpMM (dEf = H^T (Dv^-1 dOut)).  node_scale / bias nullable.
 * ------------------------------------------------------------------------------------ */
int ihg_edge_gather_sum(const float* src, int64_t src_ld, const float* node_scale,
                        float alpha, const float* bias, const int32_t* i3, int64_t edge_count,
                        float* out, int64_t out_ld, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------
 * a6  FeatureInteractor.forward, order 2/3          Models/CommonLayers.py:68-85
 *   ef[e,:] = p[u]+p[q]+p[i]                      (first-order blocks, hoisted: p = X' Wa_s^T
 *                                                  per node type, bias folded into p)
 *           + W_hi . cat(u*q, q*i, i*u [, u*q*i])  u,q,i = xp rows of the edge's nodes
 * w_hi = aggregation.weight[:, 3*dim:] ([dim, nb*dim], row stride w_ld), nb = order+... 3 or 4.
 * Backward: given def = dL/def [E,dim] writes slot_grad[e,s,:] = dL/d(xp row of slot s)
 * through the products only, and dw_hi [dim, nb*dim] (dense, deterministic two-pass sum).
 * dim % 4 == 0, dim <= 128.  dim % 32 == 0 runs on the tensor cores (tcgen05, 3xTF32 split
 * precision, fp32 accumulate); other dims on fp32 FFMA.
 * ------------------------------------------------------------------------------------ */
int64_t ihg_edge_interact_fwd_workspace_bytes(int32_t dim, int32_t order);
int ihg_edge_interact_fwd(const float* xp, int64_t xp_ld, const float* p, int64_t p_ld,
                          const float* w_hi, int64_t w_ld, int32_t order, const int32_t* i3,
                          int64_t edge_count, float* ef, int64_t ef_ld, int32_t dim,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* The whole FeatureInteractor.forward of order 2/3 in one call (CommonLayers.py:68-85):
 *   ef[e,:] = aggregation.weight . cat(u, q, i, u*q, q*i, i*u [, u*q*i]) + bias
 * w_agg = aggregation.weight [dim, K*dim] (K = 6 or 7, row stride w_ld), bias [dim] or null.
 * Tensor-core path only (dim % 32 == 0, dim <= 128): the raw rows are three more operand blocks
 * next to the products, so no first-order table is gathered.  Same workspace size as
 * ihg_edge_interact_fwd.  Returns IHG_ERR_INVALID_ARGUMENT for other dims (use the hoisted
 * ihg_edge_interact_fwd there). */
int ihg_feature_interact_fwd(const float* xp, int64_t xp_ld, const float* w_agg, int64_t w_ld,
                             const float* bias, int32_t order, const int32_t* i3,
                             int64_t edge_count, float* ef, int64_t ef_ld, int32_t dim,
                             void* workspace, int64_t workspace_bytes, void* stream);
/* 1 if ihg_feature_interact_fwd supports this dim on this build, else 0. */
int ihg_feature_interact_supported(int32_t dim);
int64_t ihg_edge_interact_bwd_workspace_bytes(int32_t dim, int32_t order);
int ihg_edge_interact_bwd(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                          const float* w_hi, int64_t w_ld, int32_t order, const int32_t* i3,
                          int64_t edge_count, float* slot_grad, float* dw_hi, int32_t dim,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * a5  feature_transform / hoisted aggregation Linear  GnnLayers.py:224, CommonLayers.py:66
 * Node-level typed Linear: rows [0,bound0) use weight 0, [bound0,bound1) weight 1, the rest
 * weight 2 (n_types = 1: one weight for all rows).
 *   transpose_w = 0:  y[r,:] = x[r,:] . W[t]^T (+ bias[t]) (+ addend[r,:]),  W[t]: [n_out, n_in]
 *   transpose_w = 1:  y[r,:] = x[r,:] . W[t]   (+ ...),                      W[t]: [n_in, n_out]
 * wgrad: dw[t] = sum_{r in type t} dy[r]^T x[r]  ([n_out, n_in]), db[t] = sum dy[r]
 * (deterministic two-pass).  n_in, n_out multiples of 4, <= 128.
 * ------------------------------------------------------------------------------------ */
int ihg_node_linear(const float* x, int64_t x_ld, const float* w, int32_t n_types,
                    int32_t n_out, int32_t n_in, int32_t transpose_w, const float* bias,
                    const float* addend, int64_t addend_ld, int64_t n_rows, int64_t bound0,
                    int64_t bound1, float* y, int64_t y_ld, void* stream);
int64_t ihg_node_linear_wgrad_workspace_bytes(int32_t n_types, int32_t n_out, int32_t n_in);
int ihg_node_linear_wgrad(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld,
                          int64_t n_rows, int64_t bound0, int64_t bound1, int32_t n_types,
                          int32_t n_out, int32_t n_in, float* dw, float* db, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * a3  EmbeddingLayer lookups                       Models/EmbeddingLayers.py:70-81
 * copy_rows   : dst[r,:] = src[r,:] for a row range (the identity-index lookup of all
 *               users / items: weight[1:], 128-bit vectorised)
 * gather_rows : out[b,:] = table[idx[b] + idx_offset, :]   (indexed form; RawGnn.py:128-133)
 * scatter_add_rows: out[idx[b]+idx_offset,:] += g[b,:], duplicates summed in ascending b
 *               (backward of gather_rows; `out` must hold the base values, e.g. zeros).
 * ------------------------------------------------------------------------------------ */
int ihg_copy_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows,
                  int32_t dim, void* stream);
int ihg_gather_rows(const float* table, int64_t table_ld, const int64_t* idx,
                    int64_t idx_offset, int64_t count, float* out, int64_t out_ld, int32_t dim,
                    void* stream);
int ihg_scatter_add_rows(const float* g, int64_t g_ld, const int64_t* idx, int64_t idx_offset,
                         int64_t count, float* out, int64_t out_ld, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------
 * f2  optimizer.step()                              Main.py:192, Helpers/TrainTestHelper.py:142-143
 *   torch.optim.Adam(params, lr, weight_decay) over a list of fp32 parameter tensors in one launch:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps),
 *   g += weight_decay * p first when weight_decay != 0 (L2 form, not AdamW).  Bit-identical with
 *   torch's fused CUDA Adam.  t = step[0] + 1; the fp32 step counters live on the device and are
 *   incremented by the call, `lr` is a device scalar: nothing host-side changes between replays of
 *   a captured graph.  `tensors_host` is a HOST array; all pointers inside are device pointers.
 * ------------------------------------------------------------------------------------ */
typedef struct ihg_adam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    float* step;               /* fp32 [1] */
    int64_t numel;
} ihg_adam_tensor;
int ihg_adam_step(const ihg_adam_tensor* tensors_host, int32_t n_tensors, const float* lr,
                  float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU halo exchange over NVLink peer memory (no reference counterpart: the reference is
 * single-device).  Segmented row copy, one (source base, destination base) pair per
 * peer; the pointer and offset tables are HOST arrays (n_seg <= 16), either side may be a
 * peer-mapped device pointer.  Flat row r of segment s (seg_off[s] <= r < seg_off[s+1]) is read
 * from seg_src[s] at row (src_rows ? src_rows[r] : r - seg_off[s]) and written to seg_dst[s] at
 * row r - seg_off[s].  src_rows: device int64 or null.
 * ------------------------------------------------------------------------------------ */
int ihg_halo_copy(const void* const* seg_src, void* const* seg_dst, const int64_t* seg_off,
                  int32_t n_seg, const int64_t* src_rows, int64_t src_ld, int64_t dst_ld,
                  int32_t dim, void* stream);

/* Owner side of the halo reduce-scatter, fused with the pull over NVLink:
 *   out[v] = row_scale[v] * (own[v] + sum_{j in [rowptr[v], rowptr[v+1])} peer_base[e.peer][e.row])   e = entries[j]
 * `entries` int32 [nnz][2] = {peer slot, row inside that peer's chunk}, ascending source rank inside a
 * row (fixed summation order); `peer_base_host` is a HOST array of n_peers <= 16 peer-mapped device
 * pointers (row stride peer_ld).  own / row_scale may be null (0 / 1). */
int ihg_halo_reduce(const float* own, int64_t own_ld, const int32_t* rowptr, const int32_t* entries,
                    const void* const* peer_base_host, int32_t n_peers, int64_t peer_ld,
                    const float* row_scale, float* out, int64_t out_ld, int64_t n_rows, int32_t dim,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * a10 HemPredictionLayer.forward                   Models/PredictionLayers.py:21-44
 *   m = lambda*q + (1-lambda)*u   (u null: m = q);  score[b] = sum_D item[b]*m[b] + bias[b']
 *   b' = item_idx[b] (item_idx null: b' = b, "all items").
 *   cosine != 0 (Gs.Prediction.use_cosine_similarity, PredictionLayers.py:38-40):
 *   score[b] = item.m / (max(|item|, 1e-8) max(|m|, 1e-8)) + bias[b'];  fwd then fills norms [count,3] =
 *   (item.m, |item|, |m|) which bwd takes back (norms null in bwd = dot-product branch).
 * bwd: d_item = g*m, d_query = g*lambda*item, d_user = g*(1-lambda)*item (null = skip),
 *      d_bias[i] = sum_{b: idx[b]==i} g[b], summed in 64-bit fixed point (order-independent,
 *      bit-reproducible; resolution 2^-38 of max|g|).  workspace: 8-byte aligned device scratch of
 *      ihg_hem_score_bwd_workspace_bytes(item_count) bytes (only read when d_bias && item_idx).
 * ------------------------------------------------------------------------------------ */
int ihg_hem_score_fwd(const float* user_f, int64_t user_ld, const float* query_f,
                      int64_t query_ld, const float* item_f, int64_t item_ld,
                      const float* items_bias, const int64_t* item_idx, float lambda_muq,
                      int64_t count, int32_t dim, float* score, int32_t cosine, float* norms,
                      void* stream);
int ihg_hem_score_bwd(const float* dscore, const float* user_f, int64_t user_ld,
                      const float* query_f, int64_t query_ld, const float* item_f,
                      int64_t item_ld, const int64_t* item_idx, float lambda_muq, int64_t count,
                      int32_t dim, float* d_user, float* d_query, float* d_item, float* d_bias,
                      int64_t item_count, void* workspace, int64_t workspace_bytes,
                      const float* norms, void* stream);
int64_t ihg_hem_score_bwd_workspace_bytes(int64_t item_count);

/* ------------------------------------------------------------------------------------
 * Inference ranking (BASELINE.json configs[4]): per (user, query) score a
 * candidate item list with the HEM scorer and keep the k best, for a whole batch of queries in
 * one launch.  Replaces the per-query loop  TestSearchLogDataLoader.__iter__ (Dataset.py:324-329)
 * -> RawGnn.forward eval branch (Models/RawGnn.py:124-142) -> HemPredictionLayer.forward
 * (Models/PredictionLayers.py:21-44) -> torch.sort(descending)[:10] (Helpers/Metrics.py:60-61).
 *   feat       [N, dim] saved output features (RawGnn.save_features_for_test, RawGnn.py:147-155)
 *   users      int64 [n_queries] user ids (row = id) or null (m = q);  queries int64 [n_queries],
 *              row = id + query_row0
 *   cand       int64 [n_queries, n_cand] item ids, or null = all items 0..n_cand-1
 *              (the reference's "predict on all items"); ids outside [0, item_count) never rank
 *   score      = sum_D feat[item_row0 + id] * (lambda*q + (1-lambda)*u) + items_bias[id]
 *                (cosine != 0: the cosine scorer of ihg_hem_score_fwd instead of the dot product)
 *   top_items  int64 [n_queries, k], top_scores fp32 [n_queries, k]: descending score, ties to the
 *              earlier candidate; unfilled places (fewer than k valid candidates) = (-1, -inf).
 * dim % 4 == 0, dim <= 1024, 1 <= k <= 32.
 * ------------------------------------------------------------------------------------ */
int ihg_rank_topk(const float* feat, int64_t feat_ld, const int64_t* users, const int64_t* queries,
                  int64_t n_queries, int64_t query_row0, const int64_t* cand, int64_t n_cand,
                  int64_t item_row0, int64_t item_count, const float* items_bias, float lambda_muq,
                  int32_t dim, int32_t k, int32_t cosine, int64_t* top_items, float* top_scores,
                  void* stream);

/* ------------------------------------------------------------------------------------
 * Training-batch sampler on the device.  Replaces GraphDataset.__getitem__
 * (Dataset.py:107-119, default nonrand_neg_sample_size == 0: the positive plus
 * random.sample(range(item_count), K) = K distinct uniform items) and GraphDataset.collate_fn
 * (Dataset.py:260-293) by one launch writing the 8-tuple collate_fn returns, in its order and
 * dtypes (all int64): positives (users, queries, items, flags == 1) for the interactions
 * pick[0..batch), then per positive its K negatives (users, queries repeated; items drawn; flags 0).
 * Draws are a pure function of (seed, step, slot, draw): reproducible; the reference is unseeded, so
 * parity is distributional.  pos_* : device int64 [E] positive interactions; neg_per_positive <= 64.
 * nonrandom_per_positive > 0 (Dataset.py:110-119): neg_items / neg_ptr = CSR, by (user, query) pair,
 * of the items that pair was shown without interacting (log order, duplicates kept), pos_pair [E] =
 * pair id of every positive.  A pair with fewer logged negatives than requested gets
 * [random fill | all of them]; otherwise [nonrandom distinct list positions | the remaining random].
 * ------------------------------------------------------------------------------------ */
int ihg_sample_batch(const int64_t* pos_user, const int64_t* pos_query, const int64_t* pos_item,
                     const int64_t* pick, int64_t batch, int32_t neg_per_positive, int64_t item_count,
                     uint64_t seed, uint64_t step, int64_t* p_users, int64_t* p_queries,
                     int64_t* p_items, int64_t* p_flags, int64_t* n_users, int64_t* n_queries,
                     int64_t* n_items, int64_t* n_flags, const int64_t* pos_pair, const int64_t* neg_ptr,
                     const int64_t* neg_items, int32_t nonrandom_per_positive, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IHGNN_B200_H */
