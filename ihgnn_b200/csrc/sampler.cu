// K-F: device-side training-batch sampler.
//
// Replaces, per training step,
//   GraphDataset.__getitem__   /root/reference/Dataset.py:107-119  (default nonrand_neg_sample_size == 0:
//       the positive (u, q, i, flag) plus random.sample(range(item_count), rand_neg_sample_size), i.e.
//       K DISTINCT items drawn uniformly from all items -- the positive item is not excluded; with
//       nonrand_neg_sample_size > 0 part of the group comes from the items the (user, query) pair was
//       shown without interacting, :110-119)
//   GraphDataset.collate_fn    /root/reference/Dataset.py:260-293  (Python lists -> 8 device tensors)
// by one launch that writes the same 8-tuple straight into device memory: no Python per-positive
// loop, no host->device copies.  The reference is unseeded, so parity is distributional: the same
// tuple layout / dtypes, negatives uniform over [0, item_count) and distinct inside a positive's
// group.  Draws are a pure function of (seed, step, positive slot, draw, attempt) -- a counter-based
// generator (splitmix64 finaliser), reproducible and independent of the launch shape.
#include "common.cuh"

namespace ihg {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr int kMaxNeg = 64;

__global__ void __launch_bounds__(128)
sample_batch_kernel(const int64_t* __restrict__ pos_user, const int64_t* __restrict__ pos_query,
                    const int64_t* __restrict__ pos_item, const int64_t* __restrict__ pick,
                    int64_t batch, int neg, int64_t item_count, uint64_t seed, uint64_t step,
                    const int64_t* __restrict__ pos_pair, const int64_t* __restrict__ neg_ptr,
                    const int64_t* __restrict__ neg_items, int nonrand,
                    int64_t* __restrict__ p_users, int64_t* __restrict__ p_queries,
                    int64_t* __restrict__ p_items, int64_t* __restrict__ p_flags,
                    int64_t* __restrict__ n_users, int64_t* __restrict__ n_queries,
                    int64_t* __restrict__ n_items, int64_t* __restrict__ n_flags) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int64_t e = __ldg(pick + b);
    const int64_t u = __ldg(pos_user + e), q = __ldg(pos_query + e);
    p_users[b] = u;
    p_queries[b] = q;
    p_items[b] = __ldg(pos_item + e);
    p_flags[b] = 1;                                        // treat_all_1 (Dataset.py:200)
    int64_t drawn[kMaxNeg];
    const uint64_t base = mix64(seed ^ mix64(step)) ^ mix64((uint64_t)b * 0x100000001B3ull);
    // how the `neg` slots of this positive are filled (Dataset.py:110-119):
    //   no logged negatives in use (nonrand == 0)       -> [neg random]
    //   fewer logged negatives than nonrand (n < nonrand) -> [neg - n random | all n logged]
    //   otherwise                                        -> [nonrand logged, distinct positions | neg - nonrand random]
    int64_t l0 = 0, n_logged = 0;
    if (nonrand > 0) {
        const int64_t pr = __ldg(pos_pair + e);
        l0 = __ldg(neg_ptr + pr);
        n_logged = __ldg(neg_ptr + pr + 1) - l0;
    }
    const bool short_list = nonrand > 0 && n_logged < nonrand;
    const int n_list = nonrand == 0 ? 0 : (short_list ? (int)n_logged : nonrand);     // slots taken from the log
    const int n_rand = neg - n_list;
    const int rand0 = short_list || nonrand == 0 ? 0 : n_list;                        // first random slot
    const int list0 = short_list ? n_rand : 0;                                        // first logged slot
    auto emit = [&](int slot, int64_t it) {
        const int64_t o = b * neg + slot;
        n_users[o] = u;
        n_queries[o] = q;
        n_items[o] = it;
        n_flags[o] = 0;
    };
    // random.sample(range(item_count), n_rand): distinct uniform items
    for (int k = 0; k < n_rand; ++k) {
        int64_t it = 0;
        for (uint64_t attempt = 0;; ++attempt) {           // rejection keeps the group distinct
            const uint64_t r = mix64(base + ((uint64_t)k << 20) + attempt);
            it = (int64_t)__umul64hi(r, (uint64_t)item_count);
            bool dup = false;
            for (int j = 0; j < k; ++j) dup |= (drawn[j] == it);
            if (!dup) break;
        }
        drawn[k] = it;
        emit(rand0 + k, it);
    }
    if (short_list) {                                       // all logged negatives, in log order
        for (int k = 0; k < n_list; ++k) emit(list0 + k, __ldg(neg_items + l0 + k));
    } else {                                                // random.sample(list, nonrand): distinct POSITIONS
        for (int k = 0; k < n_list; ++k) {
            int64_t pos = 0;
            for (uint64_t attempt = 0;; ++attempt) {
                const uint64_t r = mix64(base + 0x8000000000ull + ((uint64_t)k << 20) + attempt);
                pos = (int64_t)__umul64hi(r, (uint64_t)n_logged);
                bool dup = false;
                for (int j = 0; j < k; ++j) dup |= (drawn[n_rand + j] == pos);
                if (!dup) break;
            }
            drawn[n_rand + k] = pos;
            emit(list0 + k, __ldg(neg_items + l0 + pos));
        }
    }
}

}  // namespace ihg

using namespace ihg;

extern "C" int ihg_sample_batch(const int64_t* pos_user, const int64_t* pos_query, const int64_t* pos_item,
                                const int64_t* pick, int64_t batch, int32_t neg_per_positive,
                                int64_t item_count, uint64_t seed, uint64_t step, int64_t* p_users,
                                int64_t* p_queries, int64_t* p_items, int64_t* p_flags, int64_t* n_users,
                                int64_t* n_queries, int64_t* n_items, int64_t* n_flags,
                                const int64_t* pos_pair, const int64_t* neg_ptr, const int64_t* neg_items,
                                int32_t nonrandom_per_positive, void* stream) {
    IHG_REQUIRE(pos_user && pos_query && pos_item && pick && p_users && p_queries && p_items && p_flags,
                "sample_batch: null pointer");
    IHG_REQUIRE(neg_per_positive >= 0 && neg_per_positive <= kMaxNeg,
                "sample_batch: neg_per_positive=%d must be in [0, %d]", neg_per_positive, kMaxNeg);
    IHG_REQUIRE(neg_per_positive == 0 || (n_users && n_queries && n_items && n_flags), "sample_batch: null pointer");
    IHG_REQUIRE(nonrandom_per_positive >= 0 && nonrandom_per_positive <= neg_per_positive,
                "sample_batch: nonrandom_per_positive=%d must be in [0, neg_per_positive]", nonrandom_per_positive);
    IHG_REQUIRE(nonrandom_per_positive == 0 || (pos_pair && neg_ptr && neg_items),
                "sample_batch: logged negatives need pos_pair / neg_ptr / neg_items");
    IHG_REQUIRE(item_count >= neg_per_positive && item_count > 0,
                "sample_batch: cannot draw %d distinct items out of %lld", neg_per_positive, (long long)item_count);
    if (batch <= 0) return IHG_OK;
    sample_batch_kernel<<<(unsigned)ceil_div(batch, 128), 128, 0, as_stream(stream)>>>(
        pos_user, pos_query, pos_item, pick, batch, neg_per_positive, item_count, seed, step, pos_pair, neg_ptr,
        neg_items, nonrandom_per_positive, p_users,
        p_queries, p_items, p_flags, n_users, n_queries, n_items, n_flags);
    IHG_LAUNCH_CHECK();
    return IHG_OK;
}
