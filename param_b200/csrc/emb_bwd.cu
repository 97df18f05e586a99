// emb_bwd.cu — EmbeddingBag backward for sm_100a: scatter-add of the pooled gradient into the
// table arena (fused SGD, scale = -lr) or into a dense gradient buffer (scale = 1).
//
// Replaces the autograd backward of nn.EmbeddingBag as driven from
//   train/comms/pt/pytorch_dist_backend.py:849-857   (LookupOut.backward(grad_output))
//   train/comms/pt/dlrm.py:1296                       (tempB.backward(C))
//   train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:318-324
// Dense-equivalent semantics: dst[row(t, idx[i]), :] += scale * w_i * grad_out[(t, bag(i)), :].
//
// ATOMIC : one lane group per bag; the bag's gradient row is read once (16 B per lane) and
//          red.global.add.v4.f32 is issued per lookup.  Summation order across bags is not fixed.
// SORTED : (1) the sort plan — every table's lookups sorted by row, a hand-written radix sort fused with
//          the pair build (radix_sort.cu); it needs the indices only, so the host layer builds it on a
//          side stream at forward time (pb200_tbe_plan_build) — and (2) ONE segmented-reduce launch: a
//          lane group per 256 sorted entries accumulates runs of equal rows in registers and issues ONE
//          red per (segment, row).  Under Zipf skew ~87 % of the lookups of a table-batch
//          are duplicates, so the number of L2 read-modify-writes drops ~8x and hot rows no longer
//          serialise on one L2 slice; rows wholly inside a segment are updated exactly once
//          (deterministic), only runs crossing a segment boundary are combined by a few reds.
// EXACT  : emb_bwd_exact.cu — same sort, but every touched row is written exactly once (no atomics),
//          which is what a non-linear fused optimizer (rowwise Adagrad) and fp16 tables need.
#include "emb_bwd_common.cuh"

namespace pb200 {

// ------------------------------------------------------------------------------------
// ATOMIC
// ------------------------------------------------------------------------------------
template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__(256) tbe_bwd_atomic_kernel(const BwdParams p) {
    constexpr int BPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long gb = warp_global * BPW + grp;
    const bool active = gb < p.n_bags;
    const index_t *off = (const index_t *)p.offsets;
    const index_t *idx = (const index_t *)p.indices;

    long long begin = 0, end = 0;
    int t = 0;
    long long b = 0;
    if (active) {
        begin = ld_index<index_t>(off + gb);
        end = ld_index<index_t>(off + gb + 1);
        split_bag_bwd(p, gb, t, b);
    }
    const int len = (int)(end - begin);
    const int maxlen = (BPW == 1) ? len : __reduce_max_sync(0xffffffffu, len);
    const long long base_row = active ? p.table_row_offsets[t] : 0;

    float4 g[C];
    float s = p.scale;
    if (p.mean && len > 0) s = p.scale / (float)len;
    const float4 *go4 = (const float4 *)(p.grad_out + (long long)t * p.go_stride_t + b * p.go_stride_b);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int col = c * G + lane_g;
        if (active && col < vec4) {
            g[c] = ld_stream_f4(go4 + col);
            g[c].x *= s; g[c].y *= s; g[c].z *= s; g[c].w *= s;
        } else {
            g[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float4 *d4 = (float4 *)p.dst;
    const unsigned row_stride4 = (unsigned)vec4;
    for (int base = 0; base < maxlen; base += G) {
        unsigned my_row = 0;
        float my_w = 1.f;
        if (base + lane_g < len) {
            my_row = (unsigned)(base_row + ld_index<index_t>(idx + begin + base + lane_g));
            if (WEIGHTED) my_w = ld_stream_f32(p.psw + begin + base + lane_g);
        }
        const int cnt = min(G, maxlen - base);
        for (int j = 0; j < cnt; ++j) {
            const unsigned row = __shfl_sync(0xffffffffu, my_row, j, G);
            float w = 1.f;
            if (WEIGHTED) w = __shfl_sync(0xffffffffu, my_w, j, G);
            if (base + j < len) {
                float4 *rp = d4 + (unsigned long long)row * row_stride4;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c * G + lane_g;
                    if (col < vec4) {
                        float4 v = g[c];
                        if (WEIGHTED) { v.x *= w; v.y *= w; v.z *= w; v.w *= w; }
                        red_add_f4(rp + col, v);
                    }
                }
            }
        }
    }
}

// generic fallback (any dim/alignment): scalar atomics
template <typename index_t>
__global__ void __launch_bounds__(256) tbe_bwd_generic_kernel(const BwdParams p) {
    const int lane = threadIdx.x & 31;
    const long long gb = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gb >= p.n_bags) return;
    const index_t *off = (const index_t *)p.offsets;
    const index_t *idx = (const index_t *)p.indices;
    const long long begin = (long long)off[gb], end = (long long)off[gb + 1];
    int t;
    long long b;
    split_bag_bwd(p, gb, t, b);
    const long long base_row = p.table_row_offsets[t];
    const int len = (int)(end - begin);
    const float s = (p.mean && len > 0) ? p.scale / (float)len : p.scale;
    const float *go = p.grad_out + (long long)t * p.go_stride_t + b * p.go_stride_b;
    for (int d = lane; d < p.dim; d += 32) {
        const float g = go[d] * s;
        for (long long i = begin; i < end; ++i) {
            const long long row = base_row + (long long)idx[i];
            atomicAdd(p.dst + row * p.dim + d, p.psw ? g * p.psw[i] : g);
        }
    }
}

// ------------------------------------------------------------------------------------
// SPARSE: the values of the uncoalesced COO gradient nn.EmbeddingBag(sparse=True) produces
// ------------------------------------------------------------------------------------
// values[i, :] = w_i * grad_out[bag(i), :] for every lookup i (ATen: _embedding_bag_sparse_backward ->
// index_select of grad by offset2bag, scaled; the reference allocates its tables with sparse=True,
// pytorch_dist_backend.py:923-934).  One warp per bag: the gradient row is read once and written once per
// lookup of the bag, 16 B per lane, so a table's gradient costs nnz * dim * 4 bytes instead of a dense
// rows * dim * 4 zero-filled buffer (5.1 GB for a 10 M x 128 table).
template <typename index_t>
__global__ void __launch_bounds__(256) embbag_bwd_sparse_values_kernel(
    const float *__restrict__ grad_out, long long go_row_stride, int dim, const index_t *__restrict__ offsets,
    long long n_bags, int has_last, long long n_indices, const float *__restrict__ psw, int mean,
    float *__restrict__ values) {
    const int lane = threadIdx.x & 31;
    const long long bag = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bag >= n_bags) return;
    const long long begin = (long long)offsets[bag];
    const long long end = (bag + 1 < n_bags || has_last) ? (long long)offsets[bag + 1] : n_indices;
    const float inv = (mean && end > begin) ? 1.f / (float)(end - begin) : 1.f;
    const float *g = grad_out + bag * go_row_stride;
    const bool vec = (dim % 4 == 0) && ((((uintptr_t)g | (uintptr_t)values) & 15) == 0);
    if (vec) {
        const int v4 = dim >> 2;
        for (int c = lane; c < v4; c += 32) {
            float4 x = ld_stream_f4((const float4 *)g + c);
            x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
            for (long long i = begin; i < end; ++i) {
                float4 y = x;
                if (psw) {
                    const float w = psw[i];
                    y.x *= w; y.y *= w; y.z *= w; y.w *= w;
                }
                st_stream_f4((float4 *)(values + i * dim) + c, y);
            }
        }
    } else {
        for (int d = lane; d < dim; d += 32) {
            const float x = g[d] * inv;
            for (long long i = begin; i < end; ++i) values[i * dim + d] = psw ? x * psw[i] : x;
        }
    }
}

// ------------------------------------------------------------------------------------
// SORTED (pair builder, scratch plan and chunk pipeline: emb_bwd_common.cuh)
// ------------------------------------------------------------------------------------
// step 3: segmented reduce over the sorted pairs.  One lane group per kSeg sorted entries.
// Batches of U gradient rows are loaded unconditionally (16 B per lane, rows of past-the-end
// entries alias gradient row 0 and are masked); a batch whose first and last key equal the running
// key — the common case under skew, where one hot row spans thousands of entries — is added without
// any per-entry bookkeeping.  Each (segment, row) ends in ONE red.global.add.v4.f32.

template <int G, int C, bool SIDE>
__device__ __forceinline__ void segment_reduce_body(const BwdParams &p, long long n,
                                                    long long chunk_row0,
                                                    const unsigned *__restrict__ keys,
                                                    const unsigned *__restrict__ vals,
                                                    const unsigned *__restrict__ goff_of,
                                                    const float *__restrict__ w_of, int seg_len) {
    constexpr int BPW = 32 / G;
    constexpr int U = (C == 1) ? 8 : (C == 2 ? 4 : 2);
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;
    n = min(n, *p.n_dev);    // n is the host's capacity, the plan knows the count
    long long lo = 0;
    if (p.table_hi > p.table_lo) {
        // a table group: its sorted positions are [off[table_lo * B], off[table_hi * B]) relative to off[0]
        const long long a = (long long)p.table_lo * p.batch, b = (long long)p.table_hi * p.batch;
        long long o0, oa, ob;
        if (p.idx_is_i32) {
            const int *off = (const int *)p.offsets;
            o0 = off[0]; oa = off[a]; ob = off[b];
        } else {
            const long long *off = (const long long *)p.offsets;
            o0 = off[0]; oa = off[a]; ob = off[b];
        }
        lo = oa - o0;
        n = min(n, ob - o0);
    }
    const long long seg = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BPW + grp;
    const long long s0 = max(seg * seg_len, lo);
    const long long s1 = min((seg + 1) * (long long)seg_len, n);
    const int my_n = (s0 < s1) ? (int)(s1 - s0) : 0;
    const int max_n = (BPW == 1) ? my_n : __reduce_max_sync(0xffffffffu, my_n);

    float4 *d4 = (float4 *)p.dst + (unsigned long long)chunk_row0 * (unsigned)vec4;
    const unsigned row_stride4 = (unsigned)vec4;
    const float4 *colp[C];
    bool col_ok[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int col = c * G + lane_g;
        col_ok[c] = col < vec4;
        colp[c] = (const float4 *)p.grad_out + (col_ok[c] ? col : 0);
    }
    float4 acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned cur_key = 0xffffffffu;   // no arena row has this id (row count < 2^32 - 1)
    const unsigned long long pol_grad = l2_policy(p.l2_hints ? 1 : 0);
    const unsigned long long pol_row = l2_policy(p.l2_hints ? 2 : 0);

    auto flush = [&]() {
        if (cur_key != 0xffffffffu) {
            float4 *rp = d4 + (unsigned long long)cur_key * row_stride4;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (col_ok[c]) {
                    float4 v = acc[c];
                    v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
                    red_add_f4_hint(rp + c * G + lane_g, v, pol_row);
                }
                acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };

    for (int base = 0; base < max_n; base += G) {
        unsigned my_key = 0, my_goff = 0;
        float my_w = 0.f;
        if (base + lane_g < my_n) {
            my_key = keys[s0 + base + lane_g];
            my_goff = vals[s0 + base + lane_g];
            if (SIDE) {
                my_w = w_of[my_goff];
                my_goff = goff_of[my_goff];
            }
        }
        const int cnt = min(G, max_n - base);    // warp-uniform
        const int valid = my_n - base;           // this group's remaining entries
        for (int j0 = 0; j0 < cnt; j0 += U) {
            float4 v[U][C];
            unsigned kk[U];
            float ww[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = (j0 + u) & (G - 1);
                kk[u] = __shfl_sync(0xffffffffu, my_key, src, G);
                const unsigned goff = __shfl_sync(0xffffffffu, my_goff, src, G);
                if (SIDE) ww[u] = __shfl_sync(0xffffffffu, my_w, src, G);
#pragma unroll
                for (int c = 0; c < C; ++c) v[u][c] = ld_row_f4_hint(colp[c] + goff, pol_grad);
            }
            const bool all_valid = (j0 + U <= valid) && (j0 + U <= G);
            // every key of the batch must match (the sort may be on the low key bits only, so
            // first == last does not imply the ones in between are equal)
            bool same = true;
#pragma unroll
            for (int u = 1; u < U; ++u) same &= (kk[u] == kk[0]);
            if (!SIDE && all_valid && same && (kk[0] == cur_key || cur_key == 0xffffffffu)) {
                cur_key = kk[0];
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        add2b(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                        add2b(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (j0 + u < valid && j0 + u < G) {
                        if (kk[u] != cur_key) {
                            flush();
                            cur_key = kk[u];
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if (SIDE) {
                                acc[c].x = fmaf(ww[u], v[u][c].x, acc[c].x);
                                acc[c].y = fmaf(ww[u], v[u][c].y, acc[c].y);
                                acc[c].z = fmaf(ww[u], v[u][c].z, acc[c].z);
                                acc[c].w = fmaf(ww[u], v[u][c].w, acc[c].w);
                            } else {
                                add2b(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                                add2b(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                            }
                        }
                    }
                }
            }
        }
    }
    flush();
}

template <int G, int C, bool SIDE>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const BwdParams p, long long n,
                                                             long long chunk_row0,
                                                             const unsigned *__restrict__ keys,
                                                             const unsigned *__restrict__ vals,
                                                             const unsigned *__restrict__ goff_of,
                                                             const float *__restrict__ w_of,
                                                             int seg_len) {
    segment_reduce_body<G, C, SIDE>(p, n, chunk_row0, keys, vals, goff_of, w_of, seg_len);
}
// same body compiled for 4 resident CTAs/SM (64 registers): selectable with PB200_SEG_OCC4=1
template <int G, int C, bool SIDE>
__global__ void __launch_bounds__(256, 4) segment_reduce_kernel_occ4(const BwdParams p, long long n,
                                                                     long long chunk_row0,
                                                                     const unsigned *__restrict__ keys,
                                                                     const unsigned *__restrict__ vals,
                                                                     const unsigned *__restrict__ goff_of,
                                                                     const float *__restrict__ w_of,
                                                                     int seg_len) {
    segment_reduce_body<G, C, SIDE>(p, n, chunk_row0, keys, vals, goff_of, w_of, seg_len);
}

}  // namespace pb200

using namespace pb200;

extern "C" int64_t pb200_tbe_bwd_scratch_bytes(int64_t n_indices, int32_t num_tables, int64_t batch,
                                               int64_t total_rows, int32_t algo) {
    (void)total_rows;
    if (algo == PB200_BWD_EXACT)   // dim is not known here: size for the widest supported row
        return pb200_tbe_bwd_fused_scratch_bytes(n_indices, num_tables, batch, 512);
    if (algo != PB200_BWD_SORTED && algo != PB200_BWD_AUTO) return 0;
    if (n_indices <= 0 || num_tables < 1 || batch < 1) return 0;
    return (int64_t)plan_layout(n_indices, num_tables, batch, 0, seg_len_from_env()).total;
}

extern "C" int pb200_tbe_plan_build(void *scratch, int64_t scratch_bytes,
                                    const int64_t *table_row_offsets, int64_t max_table_rows,
                                    int32_t num_tables, int32_t dim, const void *indices,
                                    int64_t n_indices, const void *offsets, int64_t batch,
                                    int32_t idx_type, const float *psw, int32_t pool_mode,
                                    int64_t go_stride_t, int64_t go_stride_b, void *stream) {
    if (!table_row_offsets || !offsets || (!indices && n_indices > 0)) return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (idx_type != PB200_IDX_I64 && idx_type != PB200_IDX_I32) return PB200_EINVAL;
    if (batch == 0 || n_indices == 0) return PB200_OK;
    if (go_stride_t % 4 != 0 || go_stride_b % 4 != 0) return PB200_EALIGN;
    // the sort part of the layout is the same for SORTED and EXACT (reducer extras sit behind it)
    const PlanLayout L = plan_layout(n_indices, num_tables, batch, 0, seg_len_from_env());
    if (!scratch || scratch_bytes < (int64_t)L.total) return PB200_EINVAL;
    BwdParams p{};
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.n_indices = n_indices;
    p.batch = batch;
    p.n_bags = (long long)num_tables * batch;
    p.go_stride_t = go_stride_t;
    p.go_stride_b = go_stride_b;
    p.num_tables = num_tables;
    p.dim = dim;
    p.mean = pool_mode == PB200_POOL_MEAN;
    return build_sort_plan(p, idx_type, max_table_rows, scratch, L, (cudaStream_t)stream);
}

extern "C" int pb200_sort_plan_geometry(int64_t n_indices, int32_t num_tables, int64_t batch,
                                        int64_t max_table_rows, int32_t *passes, int32_t *bits, int32_t *shifts,
                                        int32_t *tile_bags, int32_t *tiles_per_table, int64_t *keys_offset,
                                        int64_t *vals_offset) {
    if (n_indices < 0 || num_tables < 1 || batch < 1 || !passes || !bits || !shifts || !tile_bags ||
        !tiles_per_table)
        return PB200_EINVAL;
    const SortGeom g = sort_geometry(n_indices, num_tables, batch, max_table_rows);
    *passes = g.passes;
    for (int i = 0; i < kSortMaxPasses; ++i) {
        bits[i] = i < g.passes ? g.bits[i] : 0;
        shifts[i] = i < g.passes ? g.shift[i] : 0;
    }
    *tile_bags = g.tile_bags;
    *tiles_per_table = g.tiles_per_table;
    const PlanLayout L = plan_layout(n_indices, num_tables, batch, 0, seg_len_from_env());
    if (keys_offset) *keys_offset = (int64_t)L.keys;
    if (vals_offset) *vals_offset = (int64_t)L.vals;
    return PB200_OK;
}

namespace pb200 {

static int bwd_sorted(const BwdParams &p, int idx_type, long long max_table_rows, void *scratch,
                      long long scratch_bytes, bool plan_ready, cudaStream_t st, int table_lo = 0,
                      int table_hi = 0) {
    const bool side = (p.psw != nullptr) || p.mean;
    const int seg_len = seg_len_from_env();
    static const int seg_occ4 = [] {
        const char *e = getenv("PB200_SEG_OCC4");
        return e ? atoi(e) : 0;
    }();
    const PlanLayout L = plan_layout(p.n_indices, p.num_tables, p.batch, 0, seg_len);
    if (!scratch || scratch_bytes < (long long)L.total) return PB200_EINVAL;
    if (p.n_indices > 0x7fffffffll) return PB200_EUNSUPPORTED;
    if (!plan_ready) {
        const int rc = build_sort_plan(p, idx_type, max_table_rows, scratch, L, st);
        if (rc != PB200_OK) return rc;
    }
    const SortedView v = sorted_view(scratch, L, p.n_indices);
    BwdParams pr = p;
    pr.n_dev = v.count;
    pr.table_lo = table_lo;
    pr.table_hi = table_hi;
    pr.idx_is_i32 = idx_type == PB200_IDX_I32;
    // measured at 64 tables (profiles/r02n_sort_*l2hint*.log): reduce 3.321 -> 3.289 ms (Zipf), 9.504 -> 9.368 ms (uniform)
    static const int l2_hints = [] {
        const char *e = getenv("PB200_SEG_L2HINT");
        return e ? atoi(e) : 1;
    }();
    pr.l2_hints = l2_hints;
    static const int seg_group = [] {
        const char *e = getenv("PB200_SEG_GROUP");
        return e ? atoi(e) : 32;
    }();
    const int vec4 = p.dim >> 2;
    const long long n = v.n, n_seg = v.n_seg;
    // segmented reduce of the whole request: ONE launch (the keys are arena rows, globally sorted)
#define PB200_SEG_LAUNCH(G_, C_)                                                                \
    do {                                                                                        \
        const long long per_block = 8ll * (32 / G_);                                            \
        const long long g2 = (n_seg + per_block - 1) / per_block;                               \
        if (g2 > 0x7fffffffll) return PB200_EUNSUPPORTED;                                       \
        if (side)                                                                               \
            segment_reduce_kernel<G_, C_, true><<<(unsigned)g2, 256, 0, st>>>(                  \
                pr, n, 0, v.keys, v.vals, v.goff_of, v.w_of, seg_len);                           \
        else if (seg_occ4)                                                                      \
            segment_reduce_kernel_occ4<G_, C_, false><<<(unsigned)g2, 256, 0, st>>>(            \
                pr, n, 0, v.keys, v.vals, nullptr, nullptr, seg_len);                            \
        else                                                                                    \
            segment_reduce_kernel<G_, C_, false><<<(unsigned)g2, 256, 0, st>>>(                 \
                pr, n, 0, v.keys, v.vals, nullptr, nullptr, seg_len);                            \
    } while (0)
    if (vec4 <= 4) PB200_SEG_LAUNCH(4, 1);
    else if (vec4 <= 8) PB200_SEG_LAUNCH(8, 1);
    else if (vec4 <= 16) PB200_SEG_LAUNCH(16, 1);
    else if (vec4 <= 32 && seg_group == 16) PB200_SEG_LAUNCH(16, 2);   // two segments per warp (trial knob)
    else if (vec4 <= 32) PB200_SEG_LAUNCH(32, 1);
    else if (vec4 <= 64) PB200_SEG_LAUNCH(32, 2);
    else PB200_SEG_LAUNCH(32, 4);
#undef PB200_SEG_LAUNCH
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

template <typename index_t, int G, int C>
static int launch_atomic(const BwdParams &p, cudaStream_t st) {
    constexpr int BPW = 32 / G;
    const long long per_block = 8ll * BPW;
    const long long grid = (p.n_bags + per_block - 1) / per_block;
    if (grid > 0x7fffffffll) return PB200_EUNSUPPORTED;
    if (p.psw)
        tbe_bwd_atomic_kernel<index_t, G, C, true><<<(unsigned)grid, 256, 0, st>>>(p);
    else
        tbe_bwd_atomic_kernel<index_t, G, C, false><<<(unsigned)grid, 256, 0, st>>>(p);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

template <typename index_t>
static int dispatch_bwd(const BwdParams &p, int algo, int idx_type, long long max_table_rows,
                        void *scratch, long long scratch_bytes, bool plan_ready, cudaStream_t st) {
    if (p.n_bags == 0 || p.n_indices == 0) return PB200_OK;
    const bool vec_ok = (p.dim % 4 == 0) && (p.dim <= 512) && (((uintptr_t)p.dst & 15) == 0) &&
                        (((uintptr_t)p.grad_out & 15) == 0) && (p.go_stride_t % 4 == 0) &&
                        (p.go_stride_b % 4 == 0);
    if (!vec_ok) {
        const long long grid = (p.n_bags + 7) / 8;
        if (grid > 0x7fffffffll) return PB200_EUNSUPPORTED;
        tbe_bwd_generic_kernel<index_t><<<(unsigned)grid, 256, 0, st>>>(p);
        count_launch();
        PB200_LAUNCH_CHECK();
        return PB200_OK;
    }
    if (algo == PB200_BWD_AUTO) algo = scratch ? PB200_BWD_SORTED : PB200_BWD_ATOMIC;
    if (algo == PB200_BWD_SORTED)
        return bwd_sorted(p, idx_type, max_table_rows, scratch, scratch_bytes, plan_ready, st);
    const int vec4 = p.dim >> 2;
    if (vec4 <= 4) return launch_atomic<index_t, 4, 1>(p, st);
    if (vec4 <= 8) return launch_atomic<index_t, 8, 1>(p, st);
    if (vec4 <= 16) return launch_atomic<index_t, 16, 1>(p, st);
    if (vec4 <= 32) return launch_atomic<index_t, 32, 1>(p, st);
    if (vec4 <= 64) return launch_atomic<index_t, 32, 2>(p, st);
    return launch_atomic<index_t, 32, 4>(p, st);
}

}  // namespace pb200

extern "C" int pb200_tbe_bwd(float *dst, const int64_t *table_row_offsets, int32_t num_tables,
                             int32_t dim, const void *indices, int64_t n_indices,
                             const void *offsets, int64_t batch, int32_t idx_type, const float *psw,
                             int32_t pool_mode, const float *grad_out, int64_t go_stride_t,
                             int64_t go_stride_b, float scale, int32_t algo, int64_t max_table_rows,
                             void *scratch, int64_t scratch_bytes, int32_t plan_ready, void *stream) {
    if (!dst || !table_row_offsets || !offsets || !grad_out || (!indices && n_indices > 0))
        return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (algo < PB200_BWD_AUTO || algo > PB200_BWD_EXACT) return PB200_EINVAL;
    if (plan_ready && (!scratch || algo == PB200_BWD_ATOMIC)) return PB200_EINVAL;
    if (algo == PB200_BWD_EXACT)   // dst -= (-scale) * dW with every touched row written exactly once
        return pb200_tbe_bwd_fused(dst, PB200_W_F32, nullptr, table_row_offsets, num_tables, dim,
                                   indices, n_indices, offsets, batch, idx_type, psw, pool_mode,
                                   grad_out, go_stride_t, go_stride_b, PB200_OPT_SGD, -scale, 0.f, 0,
                                   0, max_table_rows, scratch, scratch_bytes, plan_ready, stream);
    BwdParams p{};
    p.dst = dst;
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.grad_out = grad_out;
    p.n_indices = n_indices;
    p.batch = batch;
    p.n_bags = (long long)num_tables * batch;
    p.go_stride_t = go_stride_t;
    p.go_stride_b = go_stride_b;
    p.scale = scale;
    p.num_tables = num_tables;
    p.dim = dim;
    p.mean = pool_mode == PB200_POOL_MEAN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ready = plan_ready != 0;
    if (idx_type == PB200_IDX_I64)
        return dispatch_bwd<long long>(p, algo, idx_type, max_table_rows, scratch, scratch_bytes, ready, st);
    if (idx_type == PB200_IDX_I32)
        return dispatch_bwd<int>(p, algo, idx_type, max_table_rows, scratch, scratch_bytes, ready, st);
    return PB200_EINVAL;
}

extern "C" int pb200_embbag_bwd_sparse(const float *grad_out, int64_t go_row_stride, int32_t dim,
                                       const void *offsets, int64_t n_bags, int32_t include_last_offset,
                                       int64_t n_indices, int32_t idx_type, const float *psw,
                                       int32_t pool_mode, float *values, void *stream) {
    if (!grad_out || !values || (!offsets && n_bags > 0)) return PB200_EINVAL;
    if (dim < 1 || n_bags < 0 || n_indices < 0 || go_row_stride < dim) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (n_bags == 0 || n_indices == 0) return PB200_OK;
    const long long grid = (n_bags + 7) / 8;
    if (grid > 0x7fffffffll) return PB200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int mean = pool_mode == PB200_POOL_MEAN;
    if (idx_type == PB200_IDX_I64)
        embbag_bwd_sparse_values_kernel<long long><<<(unsigned)grid, 256, 0, st>>>(
            grad_out, go_row_stride, dim, (const long long *)offsets, n_bags, include_last_offset ? 1 : 0,
            n_indices, psw, mean, values);
    else if (idx_type == PB200_IDX_I32)
        embbag_bwd_sparse_values_kernel<int><<<(unsigned)grid, 256, 0, st>>>(
            grad_out, go_row_stride, dim, (const int *)offsets, n_bags, include_last_offset ? 1 : 0, n_indices,
            psw, mean, values);
    else
        return PB200_EINVAL;
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

extern "C" int pb200_tbe_bwd_tables(float *dst, const int64_t *table_row_offsets, int32_t num_tables,
                                    int32_t dim, const void *indices, int64_t n_indices,
                                    const void *offsets, int64_t batch, int32_t idx_type, const float *psw,
                                    int32_t pool_mode, const float *grad_out, int64_t go_stride_t,
                                    int64_t go_stride_b, float scale, int32_t table_lo, int32_t table_hi,
                                    void *plan, int64_t plan_bytes, void *stream) {
    if (!dst || !table_row_offsets || !offsets || !grad_out || !plan || (!indices && n_indices > 0))
        return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (table_lo < 0 || table_hi > num_tables || table_hi < table_lo) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (idx_type != PB200_IDX_I64 && idx_type != PB200_IDX_I32) return PB200_EINVAL;
    if (dim % 4 != 0 || dim > 512) return PB200_EUNSUPPORTED;
    if (((uintptr_t)dst & 15) || ((uintptr_t)grad_out & 15) || go_stride_t % 4 || go_stride_b % 4) return PB200_EALIGN;
    if (table_hi == table_lo || batch == 0 || n_indices == 0) return PB200_OK;
    BwdParams p{};
    p.dst = dst;
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.grad_out = grad_out;
    p.n_indices = n_indices;
    p.batch = batch;
    p.n_bags = (long long)num_tables * batch;
    p.go_stride_t = go_stride_t;
    p.go_stride_b = go_stride_b;
    p.scale = scale;
    p.num_tables = num_tables;
    p.dim = dim;
    p.mean = pool_mode == PB200_POOL_MEAN;
    return bwd_sorted(p, idx_type, 0, plan, plan_bytes, /*plan_ready=*/true, (cudaStream_t)stream, table_lo, table_hi);
}
