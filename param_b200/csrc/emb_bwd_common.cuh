// emb_bwd_common.cuh — what the EmbeddingBag backward variants share (emb_bwd.cu: ATOMIC, SORTED;
// emb_bwd_exact.cu: EXACT with the optimizer fused in): the launch parameters and the sort plan
// (sort_plan.cuh, built by radix_sort.cu) the sort-based variants reduce over.
#pragma once
#include "sort_plan.cuh"

namespace pb200 {

__device__ __forceinline__ void split_bag_bwd(const BwdParams &p, long long gb, int &t,
                                              long long &b) {
    if (p.num_tables == 1) {
        t = 0;
        b = gb;
    } else if (p.n_bags < (1ll << 31)) {
        unsigned q = (unsigned)gb / (unsigned)p.batch;
        t = (int)q;
        b = (long long)((unsigned)gb - q * (unsigned)p.batch);
    } else {
        t = (int)(gb / p.batch);
        b = gb - (long long)t * p.batch;
    }
}

// What a reducer launch gets: the globally sorted pairs of the whole request.
struct SortedView {
    long long n;
    const unsigned *keys, *vals, *goff_of;
    const float *w_of;
    unsigned char *extra;
    long long n_seg;
    const long long *count;
};

static inline SortedView sorted_view(void *plan, const PlanLayout &L, long long n) {
    unsigned char *b = (unsigned char *)plan;
    SortedView v{};
    v.n = n;
    v.keys = (const unsigned *)(b + L.keys);
    v.vals = (const unsigned *)(b + L.vals);
    v.goff_of = (const unsigned *)(b + L.goff_of);
    v.w_of = (const float *)(b + L.w_of);
    v.extra = b + L.extra;
    v.n_seg = L.n_seg;
    v.count = (const long long *)(b + L.count);
    return v;
}

}  // namespace pb200
