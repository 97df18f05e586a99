// radix_sort.cu — builds the sort plan of the EmbeddingBag backward (sort_plan.cuh): a hand-written LSD
// radix sort for sm_100a that sorts every table's lookups by row, fused with the construction of the
// (row, gradient offset) pairs.  No library kernel, no host round trip.
//
// Replaces, on the backward the reference drives at train/comms/pt/pytorch_dist_backend.py:849-857 /
// dlrm.py:1296 / split_table_batched_embeddings_ops.py:318-324, the cub::DeviceRadixSort::SortPairs call
// (plus the separate pair-build kernel and the D2H of the table bounds) of the first version.
//
// What makes this sort different from a general one:
//  * the table of a lookup is known from its POSITION (the request is table-major), so the sort is
//    segmented per table and only the table-relative row is a key: 20 bits for a 1 M-row table — two
//    passes of 10 bits instead of the three to four 8-bit passes a generic 32-bit sort needs for
//    chunk-relative arena rows;
//  * the first pass reads the int64/int32 indices directly and derives the value (the gradient row of the
//    lookup's bag) from the offsets staged in shared memory: unsorted pairs are never written;
//  * the last pass adds the table's first arena row and writes keys / values as two arrays, the form the
//    segmented reducers read: the result is one globally sorted array over all tables, so the reducers
//    run as ONE launch, with no chunk plan on the host.
//
// One pass = three kernels over tiles (a tile = the lookups of `tile_bags` consecutive bags of one table,
// grid = tiles x tables, all sizes host-known):
//   radix_hist_kernel     digit histogram of the tile -> hist[table][tile][bin]
//   radix_scan_kernel     per (table, bin): exclusive prefix over the tiles, bin totals
//   radix_scatter_kernel  re-reads the tile in sub-tiles of 4096 lookups; per-warp ranking (the lanes of a
//                         warp that hold the same digit are found with one vote.ballot per digit bit and
//                         elect a leader that bumps the warp's private counter: no atomics,
//                         deterministic), cross-warp and cross-bin scans in shared memory, the sub-tile
//                         is permuted into digit order in shared memory and written out as runs of
//                         consecutive addresses per bin.
// Digits are 8 or 10 bits wide (compile-time: the ballot loop is fully unrolled); all positions inside a
// tile are 32-bit offsets from the tile's first lookup.
// The sort is stable, so the order of equal rows — hence the summation order of the reducers — is fixed:
// lookups of a row are summed in request order (bag-major), run to run identical.
//
// Traffic per lookup for a 2-pass plan with int64 indices: 8 (hist 1: index) + 8 + 8 (scatter 1: index in,
// pair out) + 8 (hist 2: pair) + 8 + 8 (scatter 2) = 48 B, against the 1057.6 B per lookup of the
// backward's algorithmic figure.
#include "sort_plan.cuh"

namespace pb200 {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                                   // lookups per thread per sub-tile
constexpr int kSortSub = kSortThreads * kSortItems;             // 4096 lookups per sub-tile

struct SortArgs {
    const void *indices;
    const void *offsets;
    const long long *table_row_offsets;
    const float *psw;
    long long batch;
    long long go_stride_t, go_stride_b;
    int mean;
    int t_base;              // first table of this launch (blockIdx.y is relative to it)
    int tile_bags, tiles_per_table;
    int shift;
    const uint2 *src;        // pairs of the previous pass (nullptr: first pass, read the request)
    uint2 *dst_pairs;        // pairs for the next pass (nullptr: last pass)
    unsigned *dst_keys;      // last pass: arena rows, sorted
    unsigned *dst_vals;
    unsigned *goff_of;       // weighted / mean: per-position side arrays, written by the first pass
    float *w_of;
    unsigned *hist;          // [T][tiles_per_table][1 << BITS]
    unsigned *bin_total;     // [T][1 << BITS]
    long long *count;        // out: offsets[T * B] - offsets[0]
    int num_tables;
    int hist_plain;          // histogram kernel: one shared-memory atomic per lookup, no warp aggregation
};

// Positions p0, p1 are absolute (they address `indices` / `psw`, which the caller may have shifted so that
// offsets need not start at 0); the plan's arrays are addressed relative to origin = offsets[0].
template <typename index_t>
__device__ __forceinline__ void tile_range(const SortArgs &a, int t, int tile, long long &bag0,
                                           int &nb, long long &p0, long long &p1, long long &origin) {
    const index_t *off = (const index_t *)a.offsets;
    origin = (long long)off[0];
    const long long tb0 = (long long)t * a.batch;
    const long long b0 = (long long)tile * a.tile_bags;
    long long b1 = b0 + a.tile_bags;
    if (b1 > a.batch) b1 = a.batch;
    bag0 = tb0 + b0;
    nb = (int)(b1 - b0);
    p0 = (long long)off[tb0 + b0];
    p1 = (long long)off[tb0 + b1];
}

// Lanes of the warp that hold the same digit as this lane, among the lanes of `valid_mask`.
// One vote.ballot per digit bit, fully unrolled: predicate from the bit, vote, one select, one and —
// four issue slots per bit spread over the four schedulers of an SM.  (The hardware match.any does this in
// one instruction but runs on the ADU pipe at ~2 cycles per lane: the first version of these kernels was
// ADU-bound, 95 % busy in the histogram — profiles/r02b_ncu_full_sort_v1.md.)
template <int B>
__device__ __forceinline__ void peer_bit(unsigned &peers, unsigned d) {
    // peers &= bit ? ballot(bit) : ~ballot(bit)   (and + setp fuse into one LOP3 with a predicate result)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, m;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "@!p not.b32 m, m;\n\t"
        "and.b32 %0, %0, m;\n\t"
        "}"
        : "+r"(peers)
        : "r"(d), "n"(1u << B));
}
template <int B, int BITS>
__device__ __forceinline__ void peer_bits(unsigned &peers, unsigned d) {
    if constexpr (B < BITS) {
        peer_bit<B>(peers, d);
        peer_bits<B + 1, BITS>(peers, d);
    }
}
template <int BITS>
__device__ __forceinline__ unsigned digit_peers(unsigned d, unsigned valid_mask) {
    unsigned peers = valid_mask;
    peer_bits<0, BITS>(peers, d);
    return peers;
}

// ---- K1: digit histogram of one tile ---------------------------------------------------------------
template <typename index_t, bool FIRST, int BITS>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const SortArgs a) {
    constexpr int BINS = 1 << BITS;
    __shared__ unsigned s_hist[BINS];
    const int t = a.t_base + blockIdx.y;
    const int tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    long long bag0, p0, p1, origin;
    int nb;
    tile_range<index_t>(a, t, tile, bag0, nb, p0, p1, origin);
    for (int b = threadIdx.x; b < BINS; b += kSortThreads) s_hist[b] = 0;
    if (FIRST && t == 0 && tile == 0 && threadIdx.x == 0) {
        const index_t *off = (const index_t *)a.offsets;
        *a.count = (long long)off[(long long)a.num_tables * a.batch] - origin;
    }
    __syncthreads();
    const unsigned n_tile = (unsigned)(p1 - p0);
    const index_t *idx_t = (const index_t *)a.indices + p0;
    const uint2 *src_t = a.src + (p0 - origin);
    constexpr int U = 4;
    for (unsigned base = 0; base < n_tile; base += kSortThreads * U) {   // CTA-uniform trip count
        unsigned d[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned rel = base + u * kSortThreads + threadIdx.x;
            ok[u] = rel < n_tile;
            unsigned key = 0;
            if (ok[u]) key = FIRST ? (unsigned)ld_index<index_t>(idx_t + rel) : src_t[rel].x;
            d[u] = (key >> a.shift) & (BINS - 1);
        }
        if (a.hist_plain) {
            // first pass: the digits of a warp's lookups are (nearly) distinct — the low bits of the distinct rows
            // of one or two bags — so aggregation finds nothing to merge; later passes: colliding atomics are
            // resolved by the shared-memory unit faster than ten ballots per lookup can aggregate them
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) atomicAdd(&s_hist[d[u]], 1u);
            continue;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned peers = digit_peers<BITS>(d[u], __ballot_sync(0xffffffffu, ok[u]));
            if (ok[u] && lane == __ffs(peers) - 1) atomicAdd(&s_hist[d[u]], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    unsigned *h = a.hist + ((size_t)t * a.tiles_per_table + tile) * BINS;
    for (int b = threadIdx.x; b < BINS; b += kSortThreads) h[b] = s_hist[b];
}

// ---- K2: per (table, bin) exclusive prefix over the tiles --------------------------------------------
// grid (ceil(bins / 32), tables); a warp = one slice of the tiles, a lane = one bin
__global__ void __launch_bounds__(kSortThreads) radix_scan_kernel(const SortArgs a, int bins) {
    __shared__ unsigned s_sum[kSortWarps][32];
    const int t = a.t_base + blockIdx.y;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + lane;
    const int nt = a.tiles_per_table;
    const int per = (nt + kSortWarps - 1) / kSortWarps;
    const int lo = slice * per;
    const int hi = min(nt, lo + per);
    unsigned *h = a.hist + (size_t)t * nt * bins + b;
    unsigned sum = 0;
    if (b < bins)
        for (int i = lo; i < hi; ++i) sum += h[(size_t)i * bins];
    s_sum[slice][lane] = sum;
    __syncthreads();
    unsigned run = 0, total = 0;
#pragma unroll
    for (int s = 0; s < kSortWarps; ++s) {
        const unsigned v = s_sum[s][lane];
        if (s < slice) run += v;
        total += v;
    }
    if (b < bins) {
        for (int i = lo; i < hi; ++i) {
            const unsigned v = h[(size_t)i * bins];
            h[(size_t)i * bins] = run;
            run += v;
        }
        if (slice == 0) a.bin_total[(size_t)t * bins + b] = total;
    }
}

// exclusive scan of N values in shared memory, in place; all threads of the CTA call it
template <int N>
__device__ __forceinline__ void block_excl_scan(unsigned *v, unsigned *s_warp) {
    constexpr int PER = (N + kSortThreads - 1) / kSortThreads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lo = min(N, (int)threadIdx.x * PER);
    const int hi = min(N, lo + PER);
    unsigned sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i)
        if (lo + i < hi) sum += v[lo + i];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = lane < kSortWarps ? s_warp[lane] : 0u;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < kSortWarps; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += x;
        }
        if (lane < kSortWarps) s_warp[lane] = wi - w;
    }
    __syncthreads();
    unsigned run = s_warp[warp] + incl - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (lo + i < hi) {
            const unsigned x = v[lo + i];
            v[lo + i] = run;
            run += x;
        }
    }
    __syncthreads();
}

__host__ __device__ constexpr size_t union_bytes(int bins) {
    return ((size_t)kSortWarps * bins * 4 > (size_t)kSortSub * 8) ? (size_t)kSortWarps * bins * 4
                                                                   : (size_t)kSortSub * 8;
}

// ---- K3: rank and scatter one tile -------------------------------------------------------------------
// dynamic shared memory: { wh[kSortWarps][BINS] (ranking)  UNION  stage[kSortSub] uint2 (permute) }
//                        | bin_base[BINS] | sub_start[BINS + 1] | offs[tile_bags + 1] (FIRST)
template <typename index_t, bool FIRST, bool LAST, bool SIDE, int BITS>
__global__ void __launch_bounds__(kSortThreads, 2) radix_scatter_kernel(const SortArgs a) {
    constexpr int BINS = 1 << BITS;
    constexpr unsigned MASK = BINS - 1;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned s_warp[kSortWarps];
    unsigned *wh = (unsigned *)s_raw;
    uint2 *stage = (uint2 *)s_raw;               // the counters are dead once the staging positions are known
    unsigned *bin_base = (unsigned *)(s_raw + union_bytes(BINS));
    unsigned *sub_start = bin_base + BINS;
    unsigned *offs = sub_start + BINS + 1;

    const int t = a.t_base + blockIdx.y;
    const int tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    long long bag0, p0, p1, origin;
    int nb;
    tile_range<index_t>(a, t, tile, bag0, nb, p0, p1, origin);
    if (p1 <= p0) return;                                           // CTA-uniform
    const unsigned n_tile = (unsigned)(p1 - p0);
    const index_t *off = (const index_t *)a.offsets;
    const index_t *idx_t = (const index_t *)a.indices + p0;
    const float *psw_t = a.psw ? a.psw + p0 : nullptr;
    const unsigned rel_origin = (unsigned)(p0 - origin);            // plan position of the tile's first lookup
    const uint2 *src_t = a.src + rel_origin;
    const unsigned table_p0 = (unsigned)((long long)off[(long long)t * a.batch] - origin);
    const unsigned row_base = LAST ? (unsigned)a.table_row_offsets[t] : 0u;
    const int shift = a.shift;

    // where this tile's lookups of every bin go: table start + bins before + same bin in earlier tiles
    {
        const unsigned *bt = a.bin_total + (size_t)t * BINS;
        for (int b = threadIdx.x; b < BINS; b += kSortThreads) bin_base[b] = bt[b];
        __syncthreads();
        block_excl_scan<BINS>(bin_base, s_warp);
        const unsigned *h = a.hist + ((size_t)t * a.tiles_per_table + tile) * BINS;
        for (int b = threadIdx.x; b < BINS; b += kSortThreads) bin_base[b] += table_p0 + h[b];
    }
    if (FIRST) {
        // bag boundaries of the tile relative to p0; bag j of the tile covers [offs[j], offs[j + 1])
        for (int j = threadIdx.x; j <= nb; j += kSortThreads) offs[j] = (unsigned)((long long)off[bag0 + j] - p0);
    }
    __syncthreads();

    // FIRST: how the bag of a position is found.  Equal bag lengths in the tile (fixed-size bags, the
    // benchmark and DLRM case): one multiply-high.  Otherwise: a binary search for the thread's first
    // item, then a walk along the boundaries (a thread's positions ascend by 32 per item).
    bool uniform = false;
    unsigned len0 = 0, magic = 0;
    unsigned goff_base = 0, goff_step = 0;       // gradient row of bag j of this tile, in float4 units
    if (FIRST) {
        len0 = offs[1] - offs[0];
        bool same = true;
        for (int j = threadIdx.x; j < nb; j += kSortThreads) same &= (offs[j + 1] - offs[j]) == len0;
        uniform = __syncthreads_and(same) && len0 > 0 &&
                  (unsigned long long)nb * len0 * len0 < (1ull << 32);      // q = mulhi(rel, magic) is exact
        magic = uniform ? (unsigned)((1ull << 32) / len0) + 1u : 0u;
        goff_base = (unsigned)(((long long)t * a.go_stride_t + (long long)tile * a.tile_bags * a.go_stride_b) >> 2);
        goff_step = (unsigned)(a.go_stride_b >> 2);
    }

    for (unsigned sub = 0; sub < n_tile; sub += kSortSub) {
        const unsigned n_sub = min((unsigned)kSortSub, n_tile - sub);
        {
            uint4 *z = (uint4 *)wh;
            for (int i = threadIdx.x; i < kSortWarps * BINS / 4; i += kSortThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();

        // ---- load + per-warp ranking: warp w owns lookups [w * 32 * ITEMS, +32 * ITEMS) of the sub-tile
        unsigned key[kSortItems], val[kSortItems], rank[kSortItems];
        unsigned *my_wh = wh + warp * BINS;
        const unsigned wbase = sub + warp * 32 * kSortItems + lane;   // item k: position wbase + 32 k of the tile
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const unsigned rel = wbase + k * 32;
            key[k] = 0;
            val[k] = 0;
            if (rel < n_tile) {
                if (FIRST) {
                    key[k] = (unsigned)ld_index<index_t>(idx_t + rel);
                } else {
                    const uint2 pr = src_t[rel];
                    key[k] = pr.x;
                    val[k] = pr.y;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const bool valid = (wbase + k * 32) < n_tile;
            const unsigned d = (key[k] >> shift) & MASK;
            const unsigned peers = digit_peers<BITS>(d, __ballot_sync(0xffffffffu, valid));
            const int leader = (__ffs(peers) - 1) & 31;
            unsigned prev = 0;
            if (valid && lane == leader) {
                prev = my_wh[d];
                my_wh[d] = prev + (unsigned)__popc(peers);
            }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            rank[k] = prev + (unsigned)__popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();

        // ---- per bin: exclusive prefix over the warps (in place), count of the CTA -> sub_start
        for (int b = threadIdx.x; b < BINS; b += kSortThreads) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const unsigned v = wh[w * BINS + b];
                wh[w * BINS + b] = run;
                run += v;
            }
            sub_start[b] = run;
        }
        if (threadIdx.x == 0) sub_start[BINS] = n_sub;
        __syncthreads();
        block_excl_scan<BINS>(sub_start, s_warp);

        // ---- position of every lookup in the digit-ordered sub-tile (the counters die after this)
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const unsigned d = (key[k] >> shift) & MASK;
            rank[k] += sub_start[d] + my_wh[d];      // lookups past the tile end read bin 0: harmless
        }
        if (FIRST) {
            // the value of a lookup: the gradient row of its bag (plain sum), or its position
            int j = 0;
            bool have_j = false;
#pragma unroll
            for (int k = 0; k < kSortItems; ++k) {
                const unsigned rel = wbase + k * 32;
                if (rel < n_tile) {
                    if (uniform) {
                        j = len0 == 1 ? (int)rel : (int)__umulhi(rel, magic);
                    } else if (!have_j) {
                        // last bag boundary <= rel (empty bags repeat a boundary: the last one is the bag)
                        j = 0;
                        for (int step = 1 << (31 - __clz(nb)); step >= 1; step >>= 1) {
                            const int c = j + step;
                            if (c <= nb && offs[c] <= rel) j = c;
                        }
                        have_j = true;
                    } else {
                        while (offs[j + 1] <= rel) ++j;      // offs[nb] = lookups of the tile > rel: stops
                    }
                    const unsigned goff4 = goff_base + (unsigned)j * goff_step;
                    if (SIDE) {
                        val[k] = rel_origin + rel;
                        const float inv = a.mean ? 1.f / (float)(offs[j + 1] - offs[j]) : 1.f;
                        a.goff_of[rel_origin + rel] = goff4;
                        a.w_of[rel_origin + rel] = (psw_t ? psw_t[rel] : 1.f) * inv;
                    } else {
                        val[k] = goff4;
                    }
                }
            }
        }
        __syncthreads();

        // ---- permute the sub-tile into digit order in shared memory (over the counters)
#pragma unroll
        for (int k = 0; k < kSortItems; ++k)
            if ((wbase + k * 32) < n_tile) stage[rank[k]] = make_uint2(key[k], val[k]);
        __syncthreads();

        // ---- write out: consecutive threads -> consecutive addresses inside a bin's run
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const unsigned i = threadIdx.x + k * kSortThreads;
            if (i < n_sub) {
                const uint2 pr = stage[i];
                const unsigned d = (pr.x >> shift) & MASK;
                const unsigned o = bin_base[d] + (i - sub_start[d]);
                if (LAST) {
                    a.dst_keys[o] = pr.x + row_base;
                    a.dst_vals[o] = pr.y;
                } else {
                    a.dst_pairs[o] = pr;
                }
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < BINS; b += kSortThreads) bin_base[b] += sub_start[b + 1] - sub_start[b];
        // the barrier after the next sub-tile's counter reset orders this update before its readers
    }
}

static size_t scatter_smem_bytes(int bits, int tile_bags, bool first) {
    const size_t bins = (size_t)1 << bits;
    size_t bytes = union_bytes((int)bins) + (2 * bins + 1) * 4;
    if (first) bytes += ((size_t)tile_bags + 1) * 4;
    return bytes;
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
    if (bytes > 48 * 1024)
        PB200_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PB200_OK;
}

template <typename index_t, bool FIRST, bool LAST, int BITS>
static int launch_pass(const SortArgs &a, bool side, dim3 grid, cudaStream_t st) {
    radix_hist_kernel<index_t, FIRST, BITS><<<grid, kSortThreads, 0, st>>>(a);
    PB200_LAUNCH_CHECK();
    dim3 sgrid((unsigned)(((1u << BITS) + 31) / 32), grid.y);
    radix_scan_kernel<<<sgrid, kSortThreads, 0, st>>>(a, 1 << BITS);
    PB200_LAUNCH_CHECK();
    const size_t smem = scatter_smem_bytes(BITS, a.tile_bags, FIRST);
    int rc;
    if (FIRST && side) {
        auto k = radix_scatter_kernel<index_t, FIRST, LAST, true, BITS>;
        if ((rc = set_smem(k, smem)) != PB200_OK) return rc;
        k<<<grid, kSortThreads, smem, st>>>(a);
    } else {
        auto k = radix_scatter_kernel<index_t, FIRST, LAST, false, BITS>;
        if ((rc = set_smem(k, smem)) != PB200_OK) return rc;
        k<<<grid, kSortThreads, smem, st>>>(a);
    }
    PB200_LAUNCH_CHECK();
    count_launch(3);
    return PB200_OK;
}

template <typename index_t, int BITS>
static int launch_pass_fl(const SortArgs &a, bool side, bool first, bool last, dim3 grid, cudaStream_t st) {
    if (first && last) return launch_pass<index_t, true, true, BITS>(a, side, grid, st);
    if (first) return launch_pass<index_t, true, false, BITS>(a, side, grid, st);
    if (last) return launch_pass<index_t, false, true, BITS>(a, side, grid, st);
    return launch_pass<index_t, false, false, BITS>(a, side, grid, st);
}

template <typename index_t>
static int build_sort_plan_t(const BwdParams &p, long long max_table_rows, void *plan,
                             const PlanLayout &L, cudaStream_t st) {
    const bool side = (p.psw != nullptr) || p.mean;
    const SortGeom g = sort_geometry(p.n_indices, p.num_tables, p.batch, max_table_rows);
    unsigned char *base = (unsigned char *)plan;
    // tables per launch: all of them by default; PB200_SORT_GROUP = g runs the passes group by group
    static const int group_env = [] {
        const char *e = getenv("PB200_SORT_GROUP");
        return e ? atoi(e) : 0;
    }();
    // default 3: both passes count with one shared-memory atomic per lookup.  Measured at 64 tables
    // (profiles/r02n_sort_*.log): plan 1.814 -> 1.699 (first pass only) -> 1.594 ms under Zipf 1.15 — even the second
    // pass, where 59 % of the lookups share digit 0, is faster than the ballot aggregation it replaces
    static const int hist_plain_env = [] {
        const char *e = getenv("PB200_SORT_HIST_PLAIN");
        return e ? atoi(e) : 3;
    }();
    int group = group_env > 0 ? group_env : p.num_tables;
    if (group > 65535) group = 65535;

    SortArgs a{};
    a.indices = p.indices;
    a.offsets = p.offsets;
    a.table_row_offsets = p.table_row_offsets;
    a.psw = p.psw;
    a.batch = p.batch;
    a.go_stride_t = p.go_stride_t;
    a.go_stride_b = p.go_stride_b;
    a.mean = p.mean;
    a.tile_bags = g.tile_bags;
    a.tiles_per_table = g.tiles_per_table;
    a.goff_of = (unsigned *)(base + L.goff_of);
    a.w_of = (float *)(base + L.w_of);
    a.hist = (unsigned *)(base + L.hist);
    a.bin_total = (unsigned *)(base + L.bin_total);
    a.count = (long long *)(base + L.count);
    a.num_tables = p.num_tables;
    unsigned *keys = (unsigned *)(base + L.keys);
    unsigned *vals = (unsigned *)(base + L.vals);
    // ping-pong between the pair buffer and (as 8 B pairs) the final key/value area, arranged so that the
    // last pass reads the pair buffer and writes keys / vals:  P = 1: request -> out;  P = 2: request ->
    // tmp -> out;  P = 3: request -> out-as-pairs -> tmp -> out; ...
    uint2 *tmp = (uint2 *)(base + L.tmp);
    uint2 *out_as_pairs = (uint2 *)(base + L.keys);   // keys | vals are adjacent: n * 8 bytes

    for (int t0 = 0; t0 < p.num_tables; t0 += group) {
        const int tg = (t0 + group <= p.num_tables) ? group : p.num_tables - t0;
        dim3 grid((unsigned)g.tiles_per_table, (unsigned)tg);
        a.t_base = t0;
        for (int ps = 0; ps < g.passes; ++ps) {
            const bool first = ps == 0, last = ps == g.passes - 1;
            a.shift = g.shift[ps];
            a.hist_plain = (hist_plain_env & (first ? 1 : 2)) ? 1 : 0;   // bit 0: first pass, bit 1: later passes
            // pass ps writes buffer (passes - 1 - ps) & 1: 0 = out area, 1 = tmp
            uint2 *wr = ((g.passes - 1 - ps) & 1) ? tmp : out_as_pairs;
            const uint2 *rd = ((g.passes - ps) & 1) ? tmp : out_as_pairs;
            a.src = first ? nullptr : rd;
            a.dst_pairs = last ? nullptr : wr;
            a.dst_keys = last ? keys : nullptr;
            a.dst_vals = last ? vals : nullptr;
            int rc;
            if (g.bits[ps] == 8) rc = launch_pass_fl<index_t, 8>(a, side, first, last, grid, st);
            else if (g.bits[ps] == 10) rc = launch_pass_fl<index_t, 10>(a, side, first, last, grid, st);
            else rc = PB200_EUNSUPPORTED;
            if (rc != PB200_OK) return rc;
        }
    }
    return PB200_OK;
}

int build_sort_plan(const BwdParams &p, int idx_type, long long max_table_rows, void *plan,
                    const PlanLayout &L, cudaStream_t st) {
    if (p.n_indices <= 0 || p.n_bags <= 0) return PB200_OK;
    // positions, gradient-row offsets (float4 units) and arena rows travel as 32-bit values
    if (p.n_indices >= 0xffffffffll) return PB200_EUNSUPPORTED;
    {
        const long long last = (long long)(p.num_tables - 1) * p.go_stride_t + (p.batch - 1) * p.go_stride_b + p.dim;
        if ((last >> 2) >= 0xffffffffll) return PB200_EUNSUPPORTED;
    }
    if (idx_type == PB200_IDX_I64) return build_sort_plan_t<long long>(p, max_table_rows, plan, L, st);
    if (idx_type == PB200_IDX_I32) return build_sort_plan_t<int>(p, max_table_rows, plan, L, st);
    return PB200_EINVAL;
}

}  // namespace pb200
