// Host-side entry points of the tensor-core (tcgen05) kernels, called from the C-ABI wrappers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ihg {

// typed node Linear: persistent kernel, A operand in tensor memory, weights resident in shared memory
// (tc_linear_ts.cu); dimensions it does not take go to the fp32 FFMA kernel of node_linear.cu
bool node_linear_ts_eligible(int n_out, int n_in, int64_t x_ld, int64_t y_ld, const float* addend, int64_t addend_ld);
int launch_node_linear_ts(const float* x, int64_t x_ld, const float* w, int n_types, int n_out, int n_in,
                          int transpose_w, const float* bias, const float* addend, int64_t addend_ld,
                          int64_t n_rows, int64_t bound0, int64_t bound1, float* y, int64_t y_ld, cudaStream_t st);

bool node_wgrad_tc_eligible(int n_types, int n_out, int n_in);
int64_t node_wgrad_tc_workspace_bytes(int n_types, int n_out, int n_in);
int launch_node_wgrad_tc(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, int64_t n_rows,
                         int64_t b0, int64_t b1, int n_types, int n_out, int n_in, float* dw, float* db,
                         void* workspace, cudaStream_t st);
bool interact_tc_eligible(int dim);
int64_t interact_fwd_tc_workspace_bytes(int dim, int nb);
int launch_interact_prep(const float* w_hi, int64_t w_ld, int nb, int dim, int transposed,
                         void* wprep, cudaStream_t st);

// slot gradients with the def tile in tensor memory (tc_interact_slot_ts.cu)
int launch_interact_bwd_slot_ts(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                                const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                                float* slot_grad, int dim, uint8_t* wprep, cudaStream_t st);
// the whole order-2/3 FeatureInteractor.forward, A operand in tensor memory (tc_interact_ts.cu)
int launch_interact_fwd_full_ts(const float* xp, int64_t xp_ld, const float* w_agg, int64_t w_ld,
                                const float* bias, int nb, const int32_t* i3, int64_t E, float* ef,
                                int64_t ef_ld, int dim, void* workspace, cudaStream_t st);
int64_t interact_bwd_tc_workspace_bytes(int dim, int nb);
int launch_interact_bwd_tc(const float* xp, int64_t xp_ld, const float* def, int64_t def_ld,
                           const float* w_hi, int64_t w_ld, int nb, const int32_t* i3, int64_t E,
                           float* slot_grad, float* dw_hi, int dim, void* workspace, cudaStream_t st);

// Largest cluster size <= `want` (a power of two) for which the device keeps (almost) one CTA
// per SM resident; *max_ctas = resident CTAs at that size.  A persistent kernel must not need a
// second wave, so a size that strands more than 8 SMs is rejected.
template <class Kernel>
inline int pick_cluster(Kernel kernel, int threads, int smem_bytes, int want, int* max_ctas) {
    for (; want > 1; want >>= 1) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(kNumSMs / want * want);
        q.blockDim = dim3(threads);
        q.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = want, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
        q.attrs = qa, q.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kernel, &q) == cudaSuccess && n_clusters * want >= kNumSMs - 8) {
            *max_ctas = n_clusters * want > kNumSMs ? kNumSMs / want * want : n_clusters * want;
            return want;
        }
        (void)cudaGetLastError();
    }
    *max_ctas = kNumSMs;
    return 1;
}

}  // namespace ihg
