// emb_fwd.cu — EmbeddingBag forward for sm_100a: single-table and batched multi-table (TBE layout).
//
// Replaces aten::_embedding_bag as called from
//   train/compute/pt/pytorch_emb.py:61,179        (nn.EmbeddingBag(features, embdim, mode="sum"))
//   train/comms/pt/dlrm.py:363-388                (apply_emb: per-table loop + torch.stack)
//   train/comms/pt/pytorch_dist_backend.py:832-848 (emb_lookup fwd over TBE requests)
//
// Work decomposition: a "lane group" of G lanes (G = 4..32, G*C float4 >= dim/4) owns one bag;
// a warp owns 32/G bags.  Row reads are 16 B per lane, i.e. one fully coalesced 512 B request
// per row at dim = 128.  Indices are read once, coalesced, and broadcast inside the group with
// a 32-bit shuffle (the arena row id, not the 64-bit index).  U independent row loads are in
// flight per lane before the first add; adds happen in index order so fp32 sums are bit-identical
// to a sequential CPU accumulation.
//
// Two variants, identical bits:
//   DIRECT : grid covers all bags, indices/offsets read straight from global memory.
//   STAGED : persistent CTAs (multiple of the SM count); each tile's offsets and index bucket are
//            staged into shared memory with cp.async.bulk (TMA unit, UBLKCP) on an mbarrier by a
//            producer warp running kStages tiles ahead, so the offsets -> indices -> rows
//            dependency chain is off the consumers' critical path.
#include <stdlib.h>

#include "emb_core.cuh"

namespace pb200 {

// ------------------------------------------------------------------------------------
// DIRECT variant
// ------------------------------------------------------------------------------------
template <typename index_t, int G, int C, bool WEIGHTED, typename WT = float, int UU = 0>
__device__ __forceinline__ void tbe_fwd_direct_body(const FwdParams &p) {
    constexpr int BPW = 32 / G;  // bags per warp
    constexpr int U = UU > 0 ? UU : UnrollFor<C>::value;
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long gb = warp_global * BPW + grp;
    const bool active = gb < p.n_bags;

    long long begin = 0, end = 0;
    int t = 0;
    long long b = 0;
    if (active) {
        bag_range<index_t>(p, gb, begin, end);
        split_bag(p, gb, t, b);
    }
    const int len = (int)(end - begin);
    const int maxlen = (BPW == 1) ? len : __reduce_max_sync(0xffffffffu, len);
    const int minlen = (BPW == 1) ? len : __reduce_min_sync(0xffffffffu, len);
    const long long base_row = (active && p.table_row_offsets) ? p.table_row_offsets[t] : 0;

    BagAccum<index_t, G, C, WEIGHTED, U, WT> acc;
    acc.zero();
    acc.template run<false>(p, (const index_t *)p.indices + begin,
                            WEIGHTED ? p.psw + begin : nullptr, base_row, len, minlen, maxlen,
                            lane_g, vec4);
    if (active) acc.store(p, t, b, len, lane_g, vec4);
}

template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__(256) tbe_fwd_direct_kernel(const FwdParams p) {
    tbe_fwd_direct_body<index_t, G, C, WEIGHTED>(p);
}
// same body compiled for 5 resident CTAs/SM (48 registers): selectable with PB200_FWD_OCC5=1
template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__(256, 5) tbe_fwd_direct_kernel_occ5(const FwdParams p) {
    tbe_fwd_direct_body<index_t, G, C, WEIGHTED>(p);
}

// Two bags per warp for rows of 17..32 float4 (dim 68..128): 16 lanes per bag, 2 float4 per lane and row, 4 rows in
// flight per lane group, compiled for 5 resident CTAs per SM (48 registers).  Same bytes in flight per lane as the
// one-bag-per-warp form, but twice as many independent offsets -> indices -> rows chains per warp and 80 instead of
// 32 bags in flight per SM: the one-bag form was latency x occupancy bound (no unit above 55 %,
// profiles/r02n_ncu_full_fwd_sort_reduce.md).  Measured at 64 tables (profiles/r02o_*, r02p_*): 3.17 -> 2.51 ms under
// Zipf 1.15, 5.74 -> 5.71 ms under uniform indices; four bags per warp (8 lanes, 4 float4) 3.10 / 6.41 ms;
// 4 or 6 resident CTAs and 8 rows in flight within 3 % of the chosen point.  Identical bits (same order of adds).
template <typename index_t, int G, int C, int MINB, int UU>
__global__ void __launch_bounds__(256, MINB) tbe_fwd_direct_kernel_var(const FwdParams p) {
    tbe_fwd_direct_body<index_t, G, C, false, float, UU>(p);
}

// fp16 tables (fbgemm weights_precision = fp16, split_table_batched_embeddings_ops.py:291): the same
// body, rows read as 8 B vectors of 4 halves, fp32 accumulation and fp32 output.
template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__(256) tbe_fwd_direct_f16_kernel(const FwdParams p) {
    tbe_fwd_direct_body<index_t, G, C, WEIGHTED, __half>(p);
}

// ------------------------------------------------------------------------------------
// PIPELINED variant: persistent grid, register software pipeline across bags
// ------------------------------------------------------------------------------------
// Each lane group walks bags gb, gb + S, gb + 2S, ... (S = lane groups in the grid) and keeps
// three bags in flight: the offsets of bag i+2 and the first index chunk of bag i+1 are requested
// BEFORE the rows of bag i are gathered, so the offsets -> indices -> rows chain of a bag overlaps
// the row gather of its predecessors instead of being paid in sequence (no shared memory, so the
// whole 228 KB stays L1 for hot rows).
template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__(256) tbe_fwd_pipelined_kernel(const FwdParams p) {
    constexpr int BPW = 32 / G;
    constexpr int U = UnrollFor<C>::value;
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * BPW;
    long long gb = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BPW + grp;
    const index_t *idx = (const index_t *)p.indices;

    // stage registers: (b0,e0,row0,w0) current bag, (b1,e1) next, (b2,e2) the one after
    long long b0 = 0, e0 = 0, b1 = 0, e1 = 0, b2 = 0, e2 = 0;
    unsigned row0 = 0, row1 = 0;
    float w0 = 0.f, w1 = 0.f;
    int t0 = 0, t1 = 0;
    long long bb0 = 0, bb1 = 0, base_row0 = 0, base_row1 = 0;

    auto fetch_range = [&](long long g, long long &b, long long &e) {
        b = 0;
        e = 0;
        if (g < p.n_bags) bag_range<index_t>(p, g, b, e);
    };
    auto fetch_first_chunk = [&](long long g, long long b, long long e, int &t, long long &bb,
                                 long long &base_row, unsigned &row, float &w) {
        t = 0;
        bb = 0;
        base_row = 0;
        row = 0;
        w = 0.f;
        if (g < p.n_bags) {
            split_bag(p, g, t, bb);
            base_row = p.table_row_offsets ? p.table_row_offsets[t] : 0;
            row = (unsigned)base_row;
            if (lane_g < (int)(e - b)) {
                row = (unsigned)(base_row + ld_index<index_t>(idx + b + lane_g));
                if (WEIGHTED) w = ld_stream_f32(p.psw + b + lane_g);
            }
        }
    };

    fetch_range(gb, b0, e0);
    fetch_range(gb + stride, b1, e1);
    fetch_first_chunk(gb, b0, e0, t0, bb0, base_row0, row0, w0);

    // warp-uniform trip count: the first group of the warp has the smallest bag id
    const long long gb_warp = gb - grp;
    for (long long it = gb_warp; it < p.n_bags; it += stride, gb += stride) {
        // requests for the bags behind the current one (consumed one / two iterations later)
        fetch_range(gb + 2 * stride, b2, e2);
        fetch_first_chunk(gb + stride, b1, e1, t1, bb1, base_row1, row1, w1);

        const bool active = gb < p.n_bags;
        const int len = (int)(e0 - b0);
        const int maxlen = (BPW == 1) ? len : __reduce_max_sync(0xffffffffu, len);
        const int minlen = (BPW == 1) ? len : __reduce_min_sync(0xffffffffu, len);
        BagAccum<index_t, G, C, WEIGHTED, U> acc;
        acc.zero();
        acc.template run<false, true>(p, idx + b0, WEIGHTED ? p.psw + b0 : nullptr, base_row0, len,
                                      minlen, maxlen, lane_g, vec4, row0, w0);
        if (active) acc.store(p, t0, bb0, len, lane_g, vec4);

        b0 = b1; e0 = e1; row0 = row1; w0 = w1; t0 = t1; bb0 = bb1; base_row0 = base_row1;
        b1 = b2; e1 = e2;
    }
}

// ------------------------------------------------------------------------------------
// STAGED variant: persistent, warp-specialised CTAs + cp.async.bulk index/offset staging
// ------------------------------------------------------------------------------------
// CTA = kStagedWarps consumer warps + 1 producer warp.  Tile = NB consecutive bags
// (NB = kStagedWarps * BPW * kBagsPerGroup).  Per stage in shared memory:
//   s_off[OFF_PAD] (bulk copy of offsets[gb0 .. gb0+NB], padded to 16 B)
//   s_idx[cap]     (bulk copy of indices[begin_al .. end_al))
// Producer lane 0 runs up to kStages tiles ahead: waits empty[s], reads the tile's two
// boundary offsets, arms full[s] with the byte count and issues both bulk copies.
// Consumers wait full[s], gather rows, and release the stage with one arrive per warp.
// A tile falls back to direct global reads (same accumulate code) when it is partial, touches
// the end of the offsets array, or its index range does not fit / cannot be 16 B-aligned.
constexpr int kStagedWarps = 8;
constexpr int kBagsPerGroup = 2;
constexpr int kStages = 4;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct StagePlan {
    long long begin_al;
    int staged;
    int pad;
};

template <typename index_t, int G, int C, bool WEIGHTED>
__global__ void __launch_bounds__((kStagedWarps + 1) * 32) tbe_fwd_staged_kernel(const FwdParams p) {
    constexpr int BPW = 32 / G;
    constexpr int U = UnrollFor<C>::value;
    constexpr int NB = kStagedWarps * BPW * kBagsPerGroup;
    constexpr int ALIGN_ELEMS = 16 / (int)sizeof(index_t);
    constexpr int OFF_PAD = ((NB + 1 + ALIGN_ELEMS - 1) / ALIGN_ELEMS) * ALIGN_ELEMS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = p.stage_cap;  // multiple of ALIGN_ELEMS
    index_t *s_off0 = (index_t *)smem_raw;                  // [kStages][OFF_PAD]
    index_t *s_idx0 = s_off0 + kStages * OFF_PAD;           // [kStages][cap]
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ StagePlan plan[kStages];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long n_tiles = (p.n_bags + NB - 1) / NB;
    const index_t *g_off = (const index_t *)p.offsets;
    const index_t *g_idx = (const index_t *)p.indices;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kStagedWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kStagedWarps) {
        // ===== producer =====
        if (lane != 0) return;
        const long long n_idx_al = p.n_indices & ~(long long)(ALIGN_ELEMS - 1);
        const long long n_off_entries = p.n_bags + (p.has_last_offset ? 1 : 0);
        long long it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = (int)(it % kStages);
            const unsigned use = (unsigned)(it / kStages);
            if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1u);
            const long long gb0 = tile * NB;
            bool staged = (gb0 + OFF_PAD <= n_off_entries) && (gb0 + NB <= p.n_bags);
            long long begin_al = 0;
            unsigned idx_bytes = 0;
            if (staged) {
                const long long begin = (long long)g_off[gb0];
                const long long end = (long long)g_off[gb0 + NB];
                begin_al = begin & ~(long long)(ALIGN_ELEMS - 1);
                const long long end_al =
                    (end + ALIGN_ELEMS - 1) & ~(long long)(ALIGN_ELEMS - 1);
                if (end < begin || begin < 0 || end_al > n_idx_al || end_al - begin_al > cap)
                    staged = false;
                else
                    idx_bytes = (unsigned)((end_al - begin_al) * (long long)sizeof(index_t));
            }
            plan[s].begin_al = begin_al;
            plan[s].staged = staged ? 1 : 0;
            if (staged) {
                const unsigned off_bytes = (unsigned)(OFF_PAD * sizeof(index_t));
                mbar_arrive_expect_tx(&full_bar[s], idx_bytes + off_bytes);
                bulk_g2s(s_off0 + s * OFF_PAD, g_off + gb0, off_bytes, &full_bar[s]);
                if (idx_bytes)
                    bulk_g2s(s_idx0 + (long long)s * cap, g_idx + begin_al, idx_bytes,
                             &full_bar[s]);
            } else {
                mbar_arrive(&full_bar[s]);
            }
        }
        return;
    }

    // ===== consumers =====
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it % kStages);
        const unsigned use = (unsigned)(it / kStages);
        mbar_wait(&full_bar[s], use & 1u);
        const bool staged = plan[s].staged != 0;
        const long long begin_al = plan[s].begin_al;
        const long long gb0 = tile * NB;
        const index_t *s_off = s_off0 + s * OFF_PAD;
        const index_t *s_idx = s_idx0 + (long long)s * cap;

#pragma unroll 1
        for (int k = 0; k < kBagsPerGroup; ++k) {
            const int local = (k * kStagedWarps + warp) * BPW + grp;
            const long long gb = gb0 + local;
            const bool active = gb < p.n_bags;
            long long begin = 0, end = 0;
            int t = 0;
            long long b = 0;
            if (active) {
                if (staged) {
                    begin = (long long)s_off[local];
                    end = (long long)s_off[local + 1];
                } else {
                    bag_range<index_t>(p, gb, begin, end);
                }
                split_bag(p, gb, t, b);
            }
            const int len = (int)(end - begin);
            const int maxlen = (BPW == 1) ? len : __reduce_max_sync(0xffffffffu, len);
            const int minlen = (BPW == 1) ? len : __reduce_min_sync(0xffffffffu, len);
            const long long base_row =
                (active && p.table_row_offsets) ? p.table_row_offsets[t] : 0;
            BagAccum<index_t, G, C, WEIGHTED, U> acc;
            acc.zero();
            if (staged)
                acc.template run<true>(p, s_idx + (begin - begin_al),
                                       WEIGHTED ? p.psw + begin : nullptr, base_row, len, minlen,
                                       maxlen, lane_g, vec4);
            else
                acc.template run<false>(p, g_idx + begin, WEIGHTED ? p.psw + begin : nullptr,
                                        base_row, len, minlen, maxlen, lane_g, vec4);
            if (active) acc.store(p, t, b, len, lane_g, vec4);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
}

// ------------------------------------------------------------------------------------
// Generic fallback: any dim / alignment, one warp per bag, scalar loads
// ------------------------------------------------------------------------------------
template <typename index_t>
__global__ void __launch_bounds__(256) tbe_fwd_generic_kernel(const FwdParams p) {
    const int lane = threadIdx.x & 31;
    const long long gb = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gb >= p.n_bags) return;
    long long begin, end;
    bag_range<index_t>(p, gb, begin, end);
    int t;
    long long b;
    split_bag(p, gb, t, b);
    const long long base_row = p.table_row_offsets ? p.table_row_offsets[t] : 0;
    const index_t *idx = (const index_t *)p.indices;
    const int len = (int)(end - begin);
    const float cnt = (float)(len > 0 ? len : 1);
    float *o = p.out + (long long)t * p.out_stride_t + b * p.out_stride_b;
    for (int d = lane; d < p.dim; d += 32) {
        float a = 0.f;
        for (long long i = begin; i < end; ++i) {
            const long long row = base_row + (long long)idx[i];
            const float v = __ldg(p.weights + row * p.dim + d);
            if (p.psw)
                a = fmaf(p.psw[i], v, a);
            else
                a += v;
        }
        o[d] = p.mean ? __fdiv_rn(a, cnt) : a;
    }
}

// ------------------------------------------------------------------------------------
// bounds check
// ------------------------------------------------------------------------------------
template <typename index_t>
__global__ void check_indices_kernel(const long long *table_row_offsets, int num_tables,
                                     const index_t *indices, long long n_indices,
                                     const index_t *offsets, long long batch,
                                     unsigned long long *bad) {
    // one thread per bag: walks its index range and compares against its table's row count
    const long long n_bags = (long long)num_tables * batch;
    unsigned long long local_bad = 0;
    for (long long gb = (long long)blockIdx.x * blockDim.x + threadIdx.x; gb < n_bags;
         gb += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(gb / batch);
        const long long rows = table_row_offsets[t + 1] - table_row_offsets[t];
        const long long begin = (long long)offsets[gb];
        const long long end = (gb + 1 < n_bags) ? (long long)offsets[gb + 1] : n_indices;
        if (begin < 0 || end > n_indices || end < begin) {
            ++local_bad;
            continue;
        }
        for (long long i = begin; i < end; ++i) {
            const long long ix = (long long)indices[i];
            if (ix < 0 || ix >= rows) ++local_bad;
        }
    }
    if (local_bad) atomicAdd(bad, local_bad);
}

// ------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------
template <typename index_t, int G, int C>
static int launch_fwd(const FwdParams &p, int algo, cudaStream_t st) {
    constexpr int BPW = 32 / G;
    const bool weighted = p.psw != nullptr;
    if (algo == PB200_FWD_STAGED) {
        constexpr int NB = kStagedWarps * BPW * kBagsPerGroup;
        constexpr int ALIGN_ELEMS = 16 / (int)sizeof(index_t);
        constexpr int OFF_PAD = ((NB + 1 + ALIGN_ELEMS - 1) / ALIGN_ELEMS) * ALIGN_ELEMS;
        constexpr int THREADS = (kStagedWarps + 1) * 32;
        const size_t smem = (size_t)kStages * (OFF_PAD + (size_t)p.stage_cap) * sizeof(index_t);
        const long long n_tiles = (p.n_bags + NB - 1) / NB;
        auto kern = weighted ? tbe_fwd_staged_kernel<index_t, G, C, true>
                             : tbe_fwd_staged_kernel<index_t, G, C, false>;
        PB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        int per_sm = 1;
        PB200_CUDA_TRY(
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)sm_count() * per_sm;  // persistent: a multiple of the SM count
        if (grid > n_tiles) grid = n_tiles;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, THREADS, smem, st>>>(p);
    } else if (algo == PB200_FWD_PIPELINED) {
        auto kern = weighted ? tbe_fwd_pipelined_kernel<index_t, G, C, true>
                             : tbe_fwd_pipelined_kernel<index_t, G, C, false>;
        int per_sm = 1;
        PB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)sm_count() * per_sm;  // persistent: a multiple of the SM count
        const long long need = (p.n_bags + 8ll * BPW - 1) / (8ll * BPW);
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, 256, 0, st>>>(p);
    } else {
        const int warps_per_block = 8;
        const long long bags_per_block = (long long)warps_per_block * BPW;
        const long long grid = (p.n_bags + bags_per_block - 1) / bags_per_block;
        if (grid > 0x7fffffffll) return PB200_EUNSUPPORTED;
        static const int occ5 = [] {
            const char *e = getenv("PB200_FWD_OCC5");
            return e ? atoi(e) : 0;
        }();
        if constexpr (G == 16 && C == 2) {
            if (!p.weights_f16 && !weighted) {
                tbe_fwd_direct_kernel_var<index_t, G, C, 5, 4><<<(unsigned)grid, 256, 0, st>>>(p);
                count_launch();
                PB200_LAUNCH_CHECK();
                return PB200_OK;
            }
        }
        if (p.weights_f16 && weighted)
            tbe_fwd_direct_f16_kernel<index_t, G, C, true><<<(unsigned)grid, 256, 0, st>>>(p);
        else if (p.weights_f16)
            tbe_fwd_direct_f16_kernel<index_t, G, C, false><<<(unsigned)grid, 256, 0, st>>>(p);
        else if (weighted)
            tbe_fwd_direct_kernel<index_t, G, C, true><<<(unsigned)grid, 256, 0, st>>>(p);
        else if (occ5)
            tbe_fwd_direct_kernel_occ5<index_t, G, C, false><<<(unsigned)grid, 256, 0, st>>>(p);
        else
            tbe_fwd_direct_kernel<index_t, G, C, false><<<(unsigned)grid, 256, 0, st>>>(p);
    }
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

template <typename index_t>
static int dispatch_fwd(FwdParams &p, int algo, long long num_rows, cudaStream_t st) {
    if (p.n_bags == 0) return PB200_OK;
    const bool vec_ok = (p.dim % 4 == 0) && (p.dim <= 512) &&
                        (((uintptr_t)p.weights & (p.weights_f16 ? 7 : 15)) == 0) &&
                        (((uintptr_t)p.out & 15) == 0) &&
                        (p.out_stride_t % 4 == 0) && (p.out_stride_b % 4 == 0) &&
                        (num_rows < (1ll << 32));
    if (p.weights_f16) {
        // fp16 tables: vector path, DIRECT variant only (no scalar fallback)
        if (!vec_ok) return PB200_EUNSUPPORTED;
        algo = PB200_FWD_DIRECT;
    }
    if (!vec_ok) {
        const long long grid = (p.n_bags + 7) / 8;
        if (grid > 0x7fffffffll) return PB200_EUNSUPPORTED;
        tbe_fwd_generic_kernel<index_t><<<(unsigned)grid, 256, 0, st>>>(p);
        count_launch();
        PB200_LAUNCH_CHECK();
        return PB200_OK;
    }
    // measured on B200 (profiles/): DIRECT beats STAGED by 15-35 %, so AUTO = DIRECT.  (A persistent variant with
    // the head of every table staged into shared memory by cp.async.bulk was built and measured in round 2:
    // 4.30 ms with one contiguous bag range per CTA, 6.26 ms with all CTAs walking the tables in lockstep,
    // against 3.09 ms for DIRECT at 64 tables under Zipf 1.15 — profiles/r02f_ncu_fwd_hot_v1.md,
    // r02h_*.log; removed.  So was an L1-priority split inside DIRECT — head rows ld ... L1::evict_last, the
    // rest L1::no_allocate: 4.32 ms, the warm rows behind the head live on L1 too — profiles/r02i_head*.log.)
    if (algo == PB200_FWD_AUTO) algo = PB200_FWD_DIRECT;
    // bulk copies need 16 B-aligned index/offset arrays
    if ((((uintptr_t)p.indices | (uintptr_t)p.offsets) & 15) != 0) algo = PB200_FWD_DIRECT;
    if (algo == PB200_FWD_STAGED) {
        // stage capacity: ~2x the average tile footprint (+ alignment slack), a multiple of 64
        // elements, bounded so that several CTAs fit an SM and most of the 228 KB stays L1
        const double avg_len = p.n_bags ? (double)p.n_indices / (double)p.n_bags : 0.0;
        const int v4 = p.dim >> 2;
        const int G = v4 <= 4 ? 4 : v4 <= 8 ? 8 : v4 <= 16 ? 16 : 32;
        const int NB = kStagedWarps * (32 / G) * kBagsPerGroup;
        long long want = (long long)(2.0 * avg_len * NB) + 64;
        want = (want + 63) & ~63ll;
        if (want < 256) want = 256;
        if (want > 6144) want = 6144;
        p.stage_cap = (int)want;
    }
    const int vec4 = p.dim >> 2;
    if (vec4 <= 4) return launch_fwd<index_t, 4, 1>(p, algo, st);
    if (vec4 <= 8) return launch_fwd<index_t, 8, 1>(p, algo, st);
    if (vec4 <= 16) return launch_fwd<index_t, 16, 1>(p, algo, st);
    if (vec4 <= 32) {
        // two bags per warp (see tbe_fwd_direct_kernel_var); PB200_FWD_GROUP=32 selects the one-bag form
        static const int group_env = [] {
            const char *e = getenv("PB200_FWD_GROUP");
            return e ? atoi(e) : 16;
        }();
        if (algo == PB200_FWD_DIRECT && !p.weights_f16 && !p.psw && vec4 > 16 && group_env == 16)
            return launch_fwd<index_t, 16, 2>(p, algo, st);
        return launch_fwd<index_t, 32, 1>(p, algo, st);
    }
    if (vec4 <= 64) return launch_fwd<index_t, 32, 2>(p, algo, st);
    return launch_fwd<index_t, 32, 4>(p, algo, st);
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_tbe_fwd(const float *weights, const int64_t *table_row_offsets,
                             int32_t num_tables, int32_t dim, const void *indices,
                             int64_t n_indices, const void *offsets, int64_t batch,
                             int32_t idx_type, const float *psw, int32_t pool_mode, float *out,
                             int64_t out_stride_t, int64_t out_stride_b, int32_t algo,
                             void *stream) {
    if (!weights || !out || !offsets || (!indices && n_indices > 0) || !table_row_offsets)
        return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (algo < PB200_FWD_AUTO || algo > PB200_FWD_PIPELINED) return PB200_EINVAL;
    FwdParams p{};
    p.weights = weights;
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.out = out;
    p.n_indices = n_indices;
    p.batch = batch;
    p.n_bags = (long long)num_tables * batch;
    p.out_stride_t = out_stride_t;
    p.out_stride_b = out_stride_b;
    p.num_tables = num_tables;
    p.dim = dim;
    p.has_last_offset = 1;
    p.mean = pool_mode == PB200_POOL_MEAN;
    cudaStream_t st = (cudaStream_t)stream;
    // Contract (param_b200.h): the arena holds fewer than 2^32 rows — arena row ids are 32-bit inside
    // the kernels.  table_row_offsets lives on the device, so the bound is the caller's to keep
    // (ops.TableArena.allocate enforces it; 180 GB of HBM / 512 B rows is 3.5e8).
    if (idx_type == PB200_IDX_I64) return dispatch_fwd<long long>(p, algo, 0, st);
    if (idx_type == PB200_IDX_I32) return dispatch_fwd<int>(p, algo, 0, st);
    return PB200_EINVAL;
}

extern "C" int pb200_tbe_fwd_f16(const void *weights_f16, const int64_t *table_row_offsets,
                                 int32_t num_tables, int32_t dim, const void *indices,
                                 int64_t n_indices, const void *offsets, int64_t batch,
                                 int32_t idx_type, const float *psw, int32_t pool_mode, float *out,
                                 int64_t out_stride_t, int64_t out_stride_b, void *stream) {
    if (!weights_f16 || !out || !offsets || (!indices && n_indices > 0) || !table_row_offsets)
        return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    FwdParams p{};
    p.weights = (const float *)weights_f16;
    p.weights_f16 = 1;
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.out = out;
    p.n_indices = n_indices;
    p.batch = batch;
    p.n_bags = (long long)num_tables * batch;
    p.out_stride_t = out_stride_t;
    p.out_stride_b = out_stride_b;
    p.num_tables = num_tables;
    p.dim = dim;
    p.has_last_offset = 1;
    p.mean = pool_mode == PB200_POOL_MEAN;
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_type == PB200_IDX_I64) return dispatch_fwd<long long>(p, PB200_FWD_DIRECT, 0, st);
    if (idx_type == PB200_IDX_I32) return dispatch_fwd<int>(p, PB200_FWD_DIRECT, 0, st);
    return PB200_EINVAL;
}

extern "C" int pb200_embbag_fwd(const float *weight, int64_t num_rows, int32_t dim,
                                const void *indices, int64_t n_indices, const void *offsets,
                                int64_t n_bags, int32_t include_last_offset, int32_t idx_type,
                                const float *psw, int32_t pool_mode, float *out,
                                int64_t out_row_stride, int32_t algo, void *stream) {
    if (!weight || !out || (!offsets && n_bags > 0) || (!indices && n_indices > 0))
        return PB200_EINVAL;
    if (num_rows < 0 || dim < 1 || n_bags < 0 || n_indices < 0 || out_row_stride < dim)
        return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (algo < PB200_FWD_AUTO || algo > PB200_FWD_PIPELINED) return PB200_EINVAL;
    FwdParams p{};
    p.weights = weight;
    p.table_row_offsets = nullptr;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = psw;
    p.out = out;
    p.n_indices = n_indices;
    p.batch = n_bags;
    p.n_bags = n_bags;
    p.out_stride_t = 0;
    p.out_stride_b = out_row_stride;
    p.num_tables = 1;
    p.dim = dim;
    p.has_last_offset = include_last_offset ? 1 : 0;
    p.mean = pool_mode == PB200_POOL_MEAN;
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_type == PB200_IDX_I64) return dispatch_fwd<long long>(p, algo, num_rows, st);
    if (idx_type == PB200_IDX_I32) return dispatch_fwd<int>(p, algo, num_rows, st);
    return PB200_EINVAL;
}

extern "C" int pb200_check_indices(const int64_t *table_row_offsets, int32_t num_tables,
                                   const void *indices, int64_t n_indices, const void *offsets,
                                   int64_t batch, int32_t idx_type, int64_t *bad_count_dev,
                                   void *stream) {
    if (!table_row_offsets || !offsets || !bad_count_dev || num_tables < 1 || batch < 0)
        return PB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    PB200_CUDA_TRY(cudaMemsetAsync(bad_count_dev, 0, sizeof(int64_t), st));
    const long long n_bags = (long long)num_tables * batch;
    if (n_bags == 0) return PB200_OK;
    long long grid = (n_bags + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    if (idx_type == PB200_IDX_I64)
        check_indices_kernel<long long><<<(unsigned)grid, 256, 0, st>>>(
            (const long long *)table_row_offsets, num_tables, (const long long *)indices, n_indices,
            (const long long *)offsets, batch, (unsigned long long *)bad_count_dev);
    else if (idx_type == PB200_IDX_I32)
        check_indices_kernel<int><<<(unsigned)grid, 256, 0, st>>>(
            (const long long *)table_row_offsets, num_tables, (const int *)indices, n_indices,
            (const int *)offsets, batch, (unsigned long long *)bad_count_dev);
    else
        return PB200_EINVAL;
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}
