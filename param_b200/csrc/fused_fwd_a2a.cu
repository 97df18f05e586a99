// fused_fwd_a2a.cu — ONE kernel: batched EmbeddingBag lookup whose epilogue writes each pooled row
// straight into the destination rank's final [lN, T_global*E] tensor over NVLink / NVSwitch.
//
// Replaces, as a single launch per rank:
//   paramDLRM_Net.apply_emb (per-table nn.EmbeddingBag + torch.stack)   train/comms/pt/dlrm.py:363-388
//   All2Allv_Req.forward (cat + all_to_allv)                            dlrm.py:86-134
//   All2Allv_Wait.forward + torch.cat(B, dim=1)                         dlrm.py:157-177, :1253
//
// Why fuse: with separate kernels the pooled [N, T_l*E] tensor is written to HBM, read back and
// pushed while the SMs that computed it sit idle; here the NVLink transfer of bag k overlaps the
// gather of bag k+1 on every warp, the intermediate never exists, and the handshake round trip is
// hidden behind the first lookups (a warp only needs the destination's ready flag when it is about
// to store, not when it starts gathering).
//
// Grid: persistent, (SM count x resident CTAs) CTAs of 8 warps; warps stride over the bags
// (table-major, batch-minor, start rotated by rank so the W ranks do not all hit destination 0
// first).  Completion: every CTA fences (fence.acq_rel.sys) after its last store and bumps a
// device counter; the last CTA releases the done flags to all peers and then waits for theirs —
// the same epoch protocol as a2a.cu, so both kernels can be mixed on one communicator.
#include "a2a_common.cuh"
#include "emb_core.cuh"

namespace pb200 {

struct FusedArgs {
    FwdParams f;                                     // lookup request (out* unused)
    unsigned char *peer_data[PB200_A2A_MAX_RANKS];
    SignalPad *peer_pad[PB200_A2A_MAX_RANKS];
    long long recv_off[PB200_A2A_MAX_RANKS];         // where source r's columns start in MY window
    long long n_base[PB200_A2A_MAX_RANKS + 1];       // batch prefix: rows [n_base[j], n_base[j+1]) -> rank j
    long long row_bytes;                             // T_global * E * 4
    unsigned long long *epoch;
    unsigned *grid_cnt;
    unsigned *error;
    long long spin_cycles;
    long long bag_rotate;                            // start offset inside the batch (rank-dependent)
    int rank;
    int world;
};

// MINB: resident CTAs per SM asked of the compiler (1: no request)
template <typename index_t, int G, int C, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) tbe_fwd_a2a_kernel(const FusedArgs a) {
    constexpr int BPW = 32 / G;
    constexpr int U = UnrollFor<C>::value;
    __shared__ unsigned long long s_epoch;
    __shared__ volatile long long s_dst_off[PB200_A2A_MAX_RANKS];   // -1: not yet known, -2: timed out
    const FwdParams &p = a.f;
    const int W = a.world, me = a.rank;
    SignalPad *my_pad = a.peer_pad[me];
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int vec4 = p.dim >> 2;

    if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long *)a.epoch + 1ull;
    if (threadIdx.x < PB200_A2A_MAX_RANKS) s_dst_off[threadIdx.x] = (threadIdx.x == me) ? a.recv_off[me] : -1;
    __syncthreads();
    const unsigned long long e = s_epoch;

    // ready: CTA 0 posts, for every source, where its columns start in my window
    if (blockIdx.x == 0 && threadIdx.x < W && (int)threadIdx.x != me) {
        SignalPad *pp = a.peer_pad[threadIdx.x];
        st_relaxed_sys(&pp->ready_payload[me], (unsigned long long)a.recv_off[threadIdx.x]);
        st_release_sys(&pp->ready_epoch[me], e);
    }

    const long long total_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long N = p.batch;
    for (long long w = warp_global; w * BPW < p.n_bags; w += total_warps) {
        const long long lin = w * BPW + grp;
        const bool active = lin < p.n_bags;
        // bag order: table-major; inside a table the batch index is rotated by the rank
        int t = 0;
        long long b = 0;
        if (active) {
            split_bag(p, lin, t, b);
            b += a.bag_rotate;
            if (b >= N) b -= N;
        }
        const long long gb = (long long)t * N + b;
        long long begin = 0, end = 0;
        if (active) bag_range<index_t>(p, gb, begin, end);
        const int len = (int)(end - begin);
        const int maxlen = (BPW == 1) ? len : __reduce_max_sync(0xffffffffu, len);
        const int minlen = (BPW == 1) ? len : __reduce_min_sync(0xffffffffu, len);
        const long long base_row = (active && p.table_row_offsets) ? p.table_row_offsets[t] : 0;
        BagAccum<index_t, G, C, false, U> acc;
        acc.zero();
        acc.template run<false>(p, (const index_t *)p.indices + begin, nullptr, base_row, len,
                                minlen, maxlen, lane_g, vec4);
        if (active) {
            // destination of batch row b
            int j = 0;
            while (j + 1 < W && b >= a.n_base[j + 1]) ++j;
            // The group leader resolves the destination offset (first store of this CTA to rank j:
            // wait for j's ready flag on the local pad, then cache the offset in shared memory) and
            // broadcasts it.  Explicit group mask + __syncwarp: lanes of a group may not be
            // converged here, and different groups of a warp can target different ranks.
            const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
            __syncwarp(gmask);
            long long dst_off = 0;
            if (lane_g == 0) {
                dst_off = s_dst_off[j];
                if (dst_off == -1) {
                    const bool ok = wait_flag_ge(&my_pad->ready_epoch[j], e, a.spin_cycles, a.error);
                    dst_off = ok ? (long long)ld_relaxed_sys(&my_pad->ready_payload[j]) : -2;
                    s_dst_off[j] = dst_off;
                }
            }
            dst_off = __shfl_sync(gmask, dst_off, lane & ~(G - 1));
            if (dst_off >= 0) {
                float4 *o = (float4 *)(a.peer_data[j] + dst_off + (b - a.n_base[j]) * a.row_bytes) +
                            (long long)t * vec4;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int col = c * G + lane_g;
                    if (col < vec4) {
                        float4 r = acc.acc[c];
                        if (p.mean) {
                            const float cnt = (float)(len > 0 ? len : 1);
                            r.x = __fdiv_rn(r.x, cnt); r.y = __fdiv_rn(r.y, cnt);
                            r.z = __fdiv_rn(r.z, cnt); r.w = __fdiv_rn(r.w, cnt);
                        }
                        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + col),
                                     "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w)
                                     : "memory");
                    }
                }
            }
        }
    }

    // done: all of this CTA's peer stores are ordered before its counter increment
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(a.grid_cnt, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int r = 0; r < W; ++r)
                if (r != me) st_release_sys(&a.peer_pad[r]->done_epoch[me], e);
            for (int r = 0; r < W; ++r)
                if (r != me) wait_flag_ge(&my_pad->done_epoch[r], e, a.spin_cycles, a.error);
            *a.grid_cnt = 0;
            *(volatile unsigned long long *)a.epoch = e;
            __threadfence();
        }
    }
}

template <typename index_t, int G, int C, int MINB = 1>
static int launch_fused(const FusedArgs &a, int max_ctas, cudaStream_t st) {
    auto kern = tbe_fwd_a2a_kernel<index_t, G, C, MINB>;
    int per_sm = 1;
    PB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) per_sm = 1;
    constexpr int BPW = 32 / G;
    long long grid = (long long)sm_count() * per_sm;   // persistent: a multiple of the SM count
    const long long need = (a.f.n_bags + 8ll * BPW - 1) / (8ll * BPW);
    if (grid > need) grid = need;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;   // comm config (single-GPU test groups)
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, 256, 0, st>>>(a);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_tbe_fwd_a2a(pb200_a2a_comm *c, const float *weights,
                                 const int64_t *table_row_offsets, int32_t num_tables_local,
                                 int32_t dim, const void *indices, int64_t n_indices,
                                 const void *offsets, int32_t idx_type, int32_t pool_mode,
                                 const int64_t *batch_split, const int64_t *tables_split,
                                 int64_t out_window_off, void *stream) {
    if (!c || !weights || !table_row_offsets || !offsets || !batch_split || !tables_split ||
        (!indices && n_indices > 0))
        return PB200_EINVAL;
    if (dim < 4 || dim % 4 != 0 || dim > 512 || out_window_off < 0) return PB200_EUNSUPPORTED;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    const int W = c->world, me = c->rank;
    if (tables_split[me] != num_tables_local || num_tables_local < 1) return PB200_EINVAL;
    FusedArgs a{};
    long long N = 0, Tg = 0, table_base[PB200_A2A_MAX_RANKS];
    for (int r = 0; r < W; ++r) {
        if (batch_split[r] < 0 || tables_split[r] < 0) return PB200_EINVAL;
        a.n_base[r] = N;
        N += batch_split[r];
        table_base[r] = Tg;
        Tg += tables_split[r];
    }
    a.n_base[W] = N;
    const long long E = dim;
    a.row_bytes = Tg * E * 4;
    if (out_window_off + batch_split[me] * a.row_bytes > c->window_bytes) return PB200_EINVAL;
    if ((out_window_off & 15) != 0 || ((uintptr_t)weights & 15) != 0) return PB200_EALIGN;
    for (int r = 0; r < W; ++r) {
        a.peer_data[r] = c->peer_data[r];
        a.peer_pad[r] = c->peer_pad[r];
        a.recv_off[r] = out_window_off + table_base[r] * E * 4;   // source r's columns in my rows
    }
    a.epoch = c->d_epoch;
    a.grid_cnt = c->d_grid_cnt;
    a.error = c->d_error;
    a.spin_cycles = c->spin_cycles;
    a.rank = me;
    a.world = W;
    a.bag_rotate = a.n_base[(me + 1) % W];   // start with the next rank's rows
    FwdParams &p = a.f;
    p.weights = weights;
    p.table_row_offsets = (const long long *)table_row_offsets;
    p.indices = indices;
    p.offsets = offsets;
    p.psw = nullptr;
    p.out = nullptr;
    p.n_indices = n_indices;
    p.batch = N;
    p.n_bags = (long long)num_tables_local * N;
    p.num_tables = num_tables_local;
    p.dim = dim;
    p.has_last_offset = 1;
    p.mean = pool_mode == PB200_POOL_MEAN;
    if (p.n_bags == 0) return PB200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int vec4 = dim >> 2;
    // two bags per warp, as in the DIRECT forward (emb_fwd.cu); PB200_FUSED_GROUP=32: one bag per warp
    static const int fused_group = [] {
        const char *e = getenv("PB200_FUSED_GROUP");
        return e ? atoi(e) : 16;
    }();
#define PB200_FUSED(IDX)                                                     \
    do {                                                                     \
        if (vec4 <= 4) return launch_fused<IDX, 4, 1>(a, c->max_ctas, st);                \
        if (vec4 <= 8) return launch_fused<IDX, 8, 1>(a, c->max_ctas, st);                \
        if (vec4 <= 16) return launch_fused<IDX, 16, 1>(a, c->max_ctas, st);              \
        if (vec4 <= 32 && fused_group == 16) return launch_fused<IDX, 16, 2, 5>(a, c->max_ctas, st); \
        if (vec4 <= 32) return launch_fused<IDX, 32, 1>(a, c->max_ctas, st);              \
        if (vec4 <= 64) return launch_fused<IDX, 32, 2>(a, c->max_ctas, st);              \
        return launch_fused<IDX, 32, 4>(a, c->max_ctas, st);                              \
    } while (0)
    if (idx_type == PB200_IDX_I64) PB200_FUSED(long long);
    if (idx_type == PB200_IDX_I32) PB200_FUSED(int);
#undef PB200_FUSED
    return PB200_EINVAL;
}
