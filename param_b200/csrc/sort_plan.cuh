// sort_plan.cuh — the index-only half of the sort-based EmbeddingBag backward ("sort plan"): geometry and
// buffer layout shared by radix_sort.cu (which builds the plan) and the reducers (emb_bwd.cu: SORTED,
// emb_bwd_exact.cu: EXACT), which consume it.
//
// A plan is ONE device buffer that holds, for a TBE request (indices, offsets) over T tables:
//   count     int64, offsets[T * B] - offsets[0]: how many entries of keys / vals are valid
//   keys[n]   arena row of every lookup, ascending (tables are consecutive row ranges of the arena and every
//             table's lookups are sorted by row, so the whole array is sorted)
//   vals[n]   plain sum: offset of the lookup's gradient row inside grad_out, in float4 units
//             weighted / mean: the lookup's position in the request; goff_of[pos] / w_of[pos] then hold the
//             gradient offset and the weight (per-sample weight x 1/bag length)
// followed by what the sort needs while it runs (the intermediate pair buffer, per-tile digit histograms) and
// what the EXACT reducer needs (boundary-run partial sums, work list).
//
// Everything the host decides — digit plan, tile size, grid sizes, byte offsets — depends only on values
// the caller passes by value (n_indices, num_tables, batch, dim, max_table_rows): nothing is read back from
// the device, so building and consuming a plan never synchronises and is CUDA-graph capturable.  It also
// needs the indices only, not the gradient: the host layer builds it on a side stream while the forward
// lookup runs (pb200_tbe_plan_build), and the backward proper is the segmented reduce alone.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace pb200 {

// sorted entries per lane group of the reducers.  Measured at 64 tables (profiles/README.md,
// r01e_variant_*_seg256.log): 128 beats 64 by 4.5 % under Zipf and costs 1.6 % under uniform indices; 256
// beats 128 by another 3.8 % under Zipf (half as many boundary runs and partial sums) and costs 0.9 %
// under uniform indices.
constexpr int kSeg = 256;

static inline int seg_len_from_env() {
    static const int v = [] {
        const char *e = getenv("PB200_SEG");   // sorted entries per lane group
        const int x = e ? atoi(e) : kSeg;
        return x >= 8 ? x : kSeg;
    }();
    return v;
}

static inline int bits_for(unsigned long long n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    return b;
}

constexpr int kSortMaxPasses = 4;
constexpr int kSortMaxDigit = 10;      // digits are 8 or 10 bits wide (compile-time variants of the kernels)
constexpr int kSortMaxTileBags = 2048;

static inline long long sort_tile_target_from_env() {
    static const long long v = [] {
        const char *e = getenv("PB200_SORT_TILE");    // lookups per tile the tile size aims at
        const long long x = e ? atoll(e) : 12288;
        return x >= 256 ? x : 12288;
    }();
    return v;
}

struct SortGeom {
    int passes;
    int bits[kSortMaxPasses];
    int shift[kSortMaxPasses];
    int tile_bags;         // bags of one table per tile
    int tiles_per_table;
};

// max_table_rows <= 0: unknown — the keys are sorted on all 32 bits
static inline SortGeom sort_geometry(long long n_indices, int num_tables, long long batch,
                                     long long max_table_rows) {
    SortGeom g{};
    const int key_bits = max_table_rows > 0 ? bits_for((unsigned long long)max_table_rows) : 32;
    // fewest passes of 8- or 10-bit digits that cover the key: 1-8 bits -> 8; 9-10 -> 10; 11-16 -> 8+8;
    // 17-20 -> 10+10 (a 1 M-row table); 21-24 -> 8+8+8 (10 M rows); 25-30 -> 10+10+10; 31-32 -> 4 x 8
    int passes = 1, width = 8;
    if (key_bits <= 8) passes = 1, width = 8;
    else if (key_bits <= 10) passes = 1, width = 10;
    else if (key_bits <= 16) passes = 2, width = 8;
    else if (key_bits <= 20) passes = 2, width = 10;
    else if (key_bits <= 24) passes = 3, width = 8;
    else if (key_bits <= 30) passes = 3, width = 10;
    else passes = 4, width = 8;
    g.passes = passes;
    for (int p = 0; p < passes; ++p) {
        g.bits[p] = width;
        g.shift[p] = p * width;
    }
    const long long bags = (long long)num_tables * batch;
    const long long avg = bags > 0 ? (n_indices + bags - 1) / bags : 1;
    // a tile = as many bags as give ~target lookups (3 sub-tiles of 4096), a multiple of 16
    long long tb = sort_tile_target_from_env() / (avg > 0 ? avg : 1) / 16 * 16;
    if (tb < 16) tb = 16;
    if (tb > kSortMaxTileBags) tb = kSortMaxTileBags;
    const int p2 = (int)tb;
    g.tile_bags = p2;
    g.tiles_per_table = (int)((batch + p2 - 1) / p2);
    if (g.tiles_per_table < 1) g.tiles_per_table = 1;
    return g;
}

struct PlanLayout {
    size_t keys, vals, goff_of, w_of, tmp, hist, bin_total, count, extra;   // byte offsets into the plan buffer
    size_t extra_bytes, total;
    long long n_seg;
};

// extra_per_seg: reducer-private bytes per segment of seg_len sorted entries (EXACT: partial sums + work list)
static inline PlanLayout plan_layout(long long n_indices, int num_tables, long long batch,
                                     size_t extra_per_seg, int seg_len) {
    PlanLayout L{};
    const SortGeom g = sort_geometry(n_indices, num_tables, batch, 0);
    const size_t arr = ((size_t)n_indices * 4 + 255) & ~(size_t)255;
    const size_t bins = (size_t)1 << kSortMaxDigit;
    L.keys = 0;
    L.vals = arr;
    L.goff_of = 2 * arr;
    L.w_of = 3 * arr;
    L.tmp = 4 * arr;                      // uint2 pairs between two passes
    L.hist = 6 * arr;
    const size_t hist_bytes = ((size_t)num_tables * g.tiles_per_table * bins * 4 + 255) & ~(size_t)255;
    L.bin_total = L.hist + hist_bytes;
    const size_t bt_bytes = ((size_t)num_tables * bins * 4 + 255) & ~(size_t)255;
    L.count = L.bin_total + bt_bytes;     // int64: offsets[T * B] - offsets[0], the number of sorted entries
    L.extra = L.count + 256;
    L.n_seg = (n_indices + seg_len - 1) / seg_len;
    L.extra_bytes = ((size_t)L.n_seg * extra_per_seg + 16 + 255) & ~(size_t)255;
    L.total = L.extra + (extra_per_seg ? L.extra_bytes : 0);
    return L;
}

struct BwdParams {
    float *dst;
    const long long *table_row_offsets;
    const void *indices;
    const void *offsets;
    const float *psw;
    const float *grad_out;
    long long n_indices;
    long long batch;
    long long n_bags;
    long long go_stride_t;
    long long go_stride_b;
    float scale;
    int num_tables;
    int dim;
    int mean;
    // sort-based variants: the true number of lookups (offsets[T * B] - offsets[0]), written by the plan
    // build.  n_indices is the host's value and may be a capacity (a device-side redistribution hands over
    // an indices buffer whose valid prefix only the device knows): the reducers use min(n_indices, *n_dev).
    const long long *n_dev;
    // SORTED reduce of a table group: only sorted positions in [offsets[table_lo * batch], offsets[table_hi *
    // batch]) - offsets[0] are reduced (read on the device); table_hi <= table_lo: the whole request
    int table_lo, table_hi;
    int idx_is_i32;
    // SORTED reduce: L2 policies — gradient rows (re-read ~bag-size times while their table is reduced) evict_last,
    // the read-modify-write of the arena rows (touched once) evict_first.  0: no hints.
    int l2_hints;
};

// Builds the plan in `plan` (layout above) on `st`.  Returns a PB200 code; never synchronises.
int build_sort_plan(const BwdParams &p, int idx_type, long long max_table_rows, void *plan,
                    const PlanLayout &L, cudaStream_t st);

__device__ __forceinline__ void add2b(float &a0, float &a1, float b0, float b1) {
    asm("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; "
        "mov.b64 {%0,%1}, ra; }"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}

}  // namespace pb200
