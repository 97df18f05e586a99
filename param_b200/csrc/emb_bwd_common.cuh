// emb_bwd_common.cuh — what the sort-based EmbeddingBag backward variants share (emb_bwd.cu: SORTED,
// emb_bwd_exact.cu: EXACT with the optimizer fused in): the launch parameters, the (row, gradient
// offset) pair builder, the scratch plan, and the chunked build -> radix sort -> reduce pipeline.
#pragma once
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pb200 {

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
};

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

// step 1: (key, val) pairs for bags [gb_lo, gb_hi) whose lookups are [i_lo, i_hi).
//   key = arena row relative to the chunk's first row;
//   val = offset of the bag's gradient row inside grad_out, in float4 units (plain sum), or the
//         lookup position relative to i_lo (weighted / mean: weight and gradient offset come from
//         side arrays).  The segmented reduce then needs no division to find a gradient row.
template <typename index_t, bool SIDE>
__global__ void __launch_bounds__(256) build_pairs_kernel(const BwdParams p, long long gb_lo,
                                                          long long gb_hi, long long i_lo,
                                                          long long chunk_row0, unsigned *keys,
                                                          unsigned *vals, unsigned *goff_of,
                                                          float *w_of) {
    // one lane group of 8 per bag keeps the index reads coalesced for typical bag sizes
    constexpr int G = 8;
    const int lane_g = threadIdx.x & (G - 1);
    const long long gb = gb_lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (gb >= gb_hi) return;
    const index_t *off = (const index_t *)p.offsets;
    const index_t *idx = (const index_t *)p.indices;
    const long long begin = ld_index<index_t>(off + gb);
    const long long end = ld_index<index_t>(off + gb + 1);
    int t;
    long long b;
    split_bag_bwd(p, gb, t, b);
    const long long base_row = p.table_row_offsets[t] - chunk_row0;
    const unsigned goff4 = (unsigned)(((long long)t * p.go_stride_t + b * p.go_stride_b) >> 2);
    const float inv = (p.mean && end > begin) ? 1.f / (float)(end - begin) : 1.f;
    for (long long i = begin + lane_g; i < end; i += G) {
        const long long o = i - i_lo;
        keys[o] = (unsigned)(base_row + ld_index<index_t>(idx + i));
        if (SIDE) {
            vals[o] = (unsigned)o;
            goff_of[o] = goff4;
            w_of[o] = (p.psw ? p.psw[i] : 1.f) * inv;
        } else {
            vals[o] = goff4;
        }
    }
}

// sorted entries per lane group.  Measured at 64 tables (profiles/README.md, r01e_variant_*_seg256.log): 128 beats 64
// by 4.5 % under Zipf and costs 1.6 % under uniform indices; 256 beats 128 by another 3.8 % under Zipf (SORTED 6.18 ->
// 5.94 ms, EXACT 6.78 -> 6.34 ms: half as many boundary runs and partial sums) and costs 0.9 % under uniform indices.
constexpr int kSeg = 256;

__device__ __forceinline__ void add2b(float &a0, float &a1, float b0, float b1) {
    asm("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; "
        "mov.b64 {%0,%1}, ra; }"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}

static inline int bits_for(unsigned long long n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    return b;
}

static inline int seg_len_from_env() {
    static const int v = [] {
        const char *e = getenv("PB200_SEG");   // sorted entries per lane group
        const int x = e ? atoi(e) : kSeg;
        return x >= 8 ? x : kSeg;
    }();
    return v;
}

// Chunking: tables are processed in chunks whose lookups fit the scratch buffers.
struct SortedPlan {
    long long max_pairs;   // capacity of keys/vals arrays
    size_t cub_bytes;
    size_t extra_bytes;    // per-set extra the reducer asked for (EXACT: boundary-run partial sums)
    size_t total_bytes;    // one set
};

// extra_per_seg: bytes of reducer-private scratch per segment of seg_len sorted entries
static inline SortedPlan plan_sorted(long long n_indices, int num_tables, bool side,
                                     long long pair_cap = 64ll << 20, size_t extra_per_seg = 0,
                                     int seg_len = kSeg) {
    SortedPlan pl{};
    // aim for <= pair_cap pairs per chunk (64 M = 0.5 GB of key/val double buffers), at least one table
    long long per_table = num_tables > 0 ? (n_indices + num_tables - 1) / num_tables : n_indices;
    long long cap = pair_cap;
    if (cap < 2 * per_table) cap = 2 * per_table;  // slack for ragged tables; verified at run time
    if (cap > n_indices) cap = n_indices;
    if (cap < 1) cap = 1;
    pl.max_pairs = cap;
    size_t cub_bytes = 0;
    cub::DoubleBuffer<unsigned> dk(nullptr, nullptr), dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, (int)(cap > 0x7fffffffll ? 0x7fffffff : cap), 0, 32);
    pl.cub_bytes = (cub_bytes + 255) & ~(size_t)255;
    size_t arr = ((size_t)cap * 4 + 255) & ~(size_t)255;
    const size_t n_seg = (size_t)((cap + seg_len - 1) / seg_len);
    pl.extra_bytes = (n_seg * extra_per_seg + 255) & ~(size_t)255;
    pl.total_bytes = pl.cub_bytes + 4 * arr + (side ? 2 * arr : 0) + pl.extra_bytes;
    return pl;
}

static inline long long sorted_scratch_need(const SortedPlan &pl, int num_tables) {
    // two chunk sets + (T+1) int64 lookup bounds that are read back once per call
    return 2 * (long long)pl.total_bytes + 256 + (long long)(num_tables + 1) * 8;
}

// gathers offsets[t*B] for t = 0..T into a small device array (then copied to the host) so the
// host can chunk tables by lookup count without reading the whole offsets array
template <typename index_t>
__global__ void table_bounds_kernel(const index_t *offsets, long long batch, int num_tables,
                                    long long *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= num_tables) out[t] = (long long)offsets[(long long)t * batch];
}

// Scratch for one chunk in flight; two sets let chunk i+1 be built and sorted (side stream) while
// the reduce of chunk i runs (caller's stream).
struct SortSet {
    void *cub_tmp;
    unsigned *k0, *k1, *v0, *v1, *goff_of;
    float *w_of;
    unsigned char *extra;
};

struct SortedChunk {
    int t0, t1;
    long long i_lo, n, row0, row1, gb_lo, gb_hi;
    const unsigned *ks, *vs;
};

// The chunk pipeline.  reduce_fn(chunk, set, stream) launches the reducer of one sorted chunk on
// `stream` and returns a PB200 code.  full_key: the reducer needs equal rows to be contiguous
// (the PB200_SORT_BITS partial-key experiment knob is ignored).
template <typename index_t, typename ReduceFn>
static int bwd_sorted_pipeline(const BwdParams &p, const SortedPlan &pl, void *scratch,
                               long long scratch_bytes, cudaStream_t st, bool full_key,
                               ReduceFn &&reduce_fn) {
    const bool side = (p.psw != nullptr) || p.mean;
    const int T = p.num_tables;
    static const int sort_bits = [] {
        const char *e = getenv("PB200_SORT_BITS");   // 0 = full key (default); 16 = two radix passes
        return e ? atoi(e) : 0;
    }();
    static const int overlap = [] {
        const char *e = getenv("PB200_BWD_OVERLAP");  // 1 (default): sort of chunk i+1 under reduce of chunk i
        return e ? atoi(e) : 1;
    }();
    if (!scratch || scratch_bytes < sorted_scratch_need(pl, T)) return PB200_EINVAL;
    unsigned char *base = (unsigned char *)scratch;
    const size_t arr = ((size_t)pl.max_pairs * 4 + 255) & ~(size_t)255;
    SortSet sets[2];
    for (int k = 0; k < 2; ++k) {
        unsigned char *b0 = base + (size_t)k * pl.total_bytes;
        sets[k].cub_tmp = b0;
        sets[k].k0 = (unsigned *)(b0 + pl.cub_bytes);
        sets[k].k1 = (unsigned *)(b0 + pl.cub_bytes + arr);
        sets[k].v0 = (unsigned *)(b0 + pl.cub_bytes + 2 * arr);
        sets[k].v1 = (unsigned *)(b0 + pl.cub_bytes + 3 * arr);
        sets[k].goff_of = (unsigned *)(b0 + pl.cub_bytes + 4 * arr);
        sets[k].w_of = (float *)(b0 + pl.cub_bytes + 5 * arr);
        sets[k].extra = b0 + pl.total_bytes - pl.extra_bytes;
    }
    long long *d_bounds = (long long *)(base + 2 * pl.total_bytes);

    // table boundaries in lookup space and row space -> host (T+1 values each; tiny, one sync)
    table_bounds_kernel<index_t><<<(T + 1 + 127) / 128, 128, 0, st>>>((const index_t *)p.offsets,
                                                                      p.batch, T, d_bounds);
    count_launch();
    PB200_LAUNCH_CHECK();
    static thread_local long long *h_bounds = nullptr;
    static thread_local long long *h_rows = nullptr;
    static thread_local int h_cap = 0;
    if (h_cap < T + 1) {
        if (h_bounds) cudaFreeHost(h_bounds);
        if (h_rows) cudaFreeHost(h_rows);
        h_bounds = h_rows = nullptr;
        h_cap = 0;
        PB200_CUDA_TRY(cudaMallocHost(&h_bounds, (size_t)(T + 1) * 8));
        PB200_CUDA_TRY(cudaMallocHost(&h_rows, (size_t)(T + 1) * 8));
        h_cap = T + 1;
    }
    PB200_CUDA_TRY(cudaMemcpyAsync(h_bounds, d_bounds, (size_t)(T + 1) * 8, cudaMemcpyDeviceToHost, st));
    PB200_CUDA_TRY(cudaMemcpyAsync(h_rows, p.table_row_offsets, (size_t)(T + 1) * 8,
                                   cudaMemcpyDeviceToHost, st));
    PB200_CUDA_TRY(cudaStreamSynchronize(st));

    // gradient row offsets travel as 32-bit float4 indices
    {
        const long long last = (long long)(T - 1) * p.go_stride_t + (p.batch - 1) * p.go_stride_b + p.dim;
        if ((last >> 2) >= 0xffffffffll) return PB200_EUNSUPPORTED;
    }

    // ---- chunk plan (host) ----
    static thread_local SortedChunk *chunks = nullptr;
    static thread_local int chunks_cap = 0;
    if (chunks_cap < T) {
        free(chunks);
        chunks_cap = 0;
        chunks = (SortedChunk *)malloc(sizeof(SortedChunk) * (size_t)T);
        if (!chunks) return PB200_EINVAL;
        chunks_cap = T;
    }
    static const long long min_pairs = [] {
        const char *e = getenv("PB200_CHUNK_MIN_PAIRS");
        return e ? atoll(e) : (8ll << 20);
    }();
    int n_chunks = 0;
    for (int t0 = 0; t0 < T;) {
        int t1 = t0 + 1;
        // grow the chunk while its lookups fit the scratch AND its rows fit 24 key bits: the radix
        // sort then needs 3 passes instead of 4 (measured: 1/4 of the sort time at 48 tables/chunk).
        // Tables with many rows but few lookups (10 M rows, 1.3 M lookups each) would end up one per
        // chunk, each paying the launch latency of ~10 small kernels: below min_pairs lookups the
        // 24-bit rule gives way (a fourth radix pass is cheaper than 7x the launches).
        while (t1 < T && h_bounds[t1 + 1] - h_bounds[t0] <= pl.max_pairs &&
               h_rows[t1 + 1] - h_rows[t0] < 0xffffffffll &&
               (h_rows[t1 + 1] - h_rows[t0] <= (1ll << 24) || h_bounds[t1] - h_bounds[t0] < min_pairs))
            ++t1;
        SortedChunk c{};
        c.t0 = t0;
        c.t1 = t1;
        c.i_lo = h_bounds[t0];
        c.n = h_bounds[t1] - h_bounds[t0];
        c.row0 = h_rows[t0];
        c.row1 = h_rows[t1];
        c.gb_lo = (long long)t0 * p.batch;
        c.gb_hi = (long long)t1 * p.batch;
        if (c.n > pl.max_pairs) return PB200_EUNSUPPORTED;  // one table larger than the scratch plan
        if (c.row1 - c.row0 >= 0xffffffffll) return PB200_EUNSUPPORTED;   // 0xffffffff = "no row" sentinel
        if (c.gb_hi - c.gb_lo > 0xffffffffll || c.n > 0x7fffffffll) return PB200_EUNSUPPORTED;
        if ((c.gb_hi - c.gb_lo) * 8 / 256 > 0x7fffffffll) return PB200_EUNSUPPORTED;
        if (c.n > 0) chunks[n_chunks++] = c;
        t0 = t1;
    }
    if (n_chunks == 0) return PB200_OK;

    // ---- side stream + events (created once per host thread) ----
    static thread_local cudaStream_t s2 = nullptr;
    static thread_local cudaEvent_t ev_start = nullptr, ev_sorted[2] = {nullptr, nullptr},
                                    ev_seg[2] = {nullptr, nullptr};
    const bool use_overlap = overlap && n_chunks > 1;
    if (use_overlap && !s2) {
        PB200_CUDA_TRY(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        PB200_CUDA_TRY(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&ev_sorted[k], cudaEventDisableTiming));
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&ev_seg[k], cudaEventDisableTiming));
        }
    }

    // build (row, gradient offset) pairs of chunk c and sort them by row, on stream s
    auto prepare = [&](int ci, cudaStream_t s) -> int {
        SortedChunk &c = chunks[ci];
        const SortSet &ss = sets[ci & 1];
        const long long threads = (c.gb_hi - c.gb_lo) * 8;
        const long long grid = (threads + 255) / 256;
        if (side)
            build_pairs_kernel<index_t, true><<<(unsigned)grid, 256, 0, s>>>(
                p, c.gb_lo, c.gb_hi, c.i_lo, c.row0, ss.k0, ss.v0, ss.goff_of, ss.w_of);
        else
            build_pairs_kernel<index_t, false><<<(unsigned)grid, 256, 0, s>>>(
                p, c.gb_lo, c.gb_hi, c.i_lo, c.row0, ss.k0, ss.v0, nullptr, nullptr);
        count_launch();
        PB200_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned> dk(ss.k0, ss.k1), dv(ss.v0, ss.v1);
        size_t tmp = pl.cub_bytes;
        // Grouping, not ordering, is what the segmented reduce needs: PB200_SORT_BITS > 0 sorts on
        // the low bits of the row id only (stable), one radix pass less at the price of more reds.
        int key_bits = bits_for((unsigned long long)(c.row1 - c.row0));
        if (!full_key && sort_bits > 0 && key_bits > sort_bits) key_bits = sort_bits;
        PB200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ss.cub_tmp, tmp, dk, dv, (int)c.n, 0, key_bits, s));
        count_launch(4);  // onesweep: histogram + scan + digit passes (library kernels)
        c.ks = dk.Current();
        c.vs = dv.Current();
        return PB200_OK;
    };

    if (!use_overlap) {
        for (int ci = 0; ci < n_chunks; ++ci) {
            int rc = prepare(ci, st);
            if (rc != PB200_OK) return rc;
            rc = reduce_fn(chunks[ci], sets[ci & 1], st);
            if (rc != PB200_OK) return rc;
        }
        return PB200_OK;
    }

    // software pipeline over chunks: prepare(i+1) on the side stream while reduce(i) runs on `st`
    PB200_CUDA_TRY(cudaEventRecord(ev_start, st));          // inputs are ready at this point of `st`
    PB200_CUDA_TRY(cudaStreamWaitEvent(s2, ev_start, 0));
    int rc = prepare(0, st);
    if (rc != PB200_OK) return rc;
    for (int ci = 0; ci < n_chunks; ++ci) {
        if (ci + 1 < n_chunks) {
            // set (ci+1)&1 was last read by reduce(ci-1): wait for it before overwriting
            if (ci >= 1) PB200_CUDA_TRY(cudaStreamWaitEvent(s2, ev_seg[(ci + 1) & 1], 0));
            rc = prepare(ci + 1, s2);
            if (rc != PB200_OK) return rc;
            PB200_CUDA_TRY(cudaEventRecord(ev_sorted[(ci + 1) & 1], s2));
        }
        if (ci >= 1) PB200_CUDA_TRY(cudaStreamWaitEvent(st, ev_sorted[ci & 1], 0));
        rc = reduce_fn(chunks[ci], sets[ci & 1], st);
        if (rc != PB200_OK) return rc;
        PB200_CUDA_TRY(cudaEventRecord(ev_seg[ci & 1], st));
    }
    return PB200_OK;
}

}  // namespace pb200
