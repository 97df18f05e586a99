// emb_fwd_hot.cu — batched EmbeddingBag forward with a per-table hot-row cache in shared memory (sm_100a).
//
// Same contract and same bits as the DIRECT variant of emb_fwd.cu (it replaces the same reference calls:
// train/compute/pt/pytorch_emb.py:61,179, train/comms/pt/dlrm.py:363-388, pytorch_dist_backend.py:832-848);
// what changes is where the hot rows come from.
//
// Under the skew the reference's generator produces (init_indices, pytorch_emb.py:138-160: Zipf over the row
// RANK = row id, so the popular rows sit at the head of every table) the first 256 rows of a 1 M-row table
// take 59 % of the lookups at alpha = 1.15.  DIRECT leaves them to L1: four resident CTAs per SM compete for
// it with the streaming cold rows, the hit rate is 50 %, and every miss is a 512 B round trip to L2 —
// 86 GB of L2 -> SM traffic per step at cfg2, which is what bounds that kernel (ncu: no unit saturated, DRAM at
// 31 %, profiles/r01c_ncu_full_fwd_256tables.md).  Here:
//   * persistent grid, ONE CTA of 32 warps per SM; all CTAs walk the tables in the same order, each taking
//     every gridDim-th group of 32 bags of the current table (all SMs stay on one table: L2 locality as DIRECT);
//   * at every table the CTA copies the head of the table (K rows, 128 KB at dim 128) into shared memory
//     with cp.async.bulk (TMA unit, SASS UBLKCP) on an mbarrier — one copy per SM instead of one L1 image
//     per resident CTA;
//   * a lookup whose row is < K is served by LDS.128, the others by the same LDG.128 path as DIRECT.  With
//     G = 32 (dim 65..256) the row of a lookup is warp-uniform, so the choice is a uniform predicate: no
//     divergence, loads stay batched 8 deep, adds stay in index order (bit-identical sums).
// A table whose popular rows are elsewhere loses nothing but the L1 capacity handed to the cache.
#include <stdlib.h>

#include "emb_core.cuh"

namespace pb200 {

constexpr int kHotThreads = 768;
constexpr int kHotWarps = kHotThreads / 32;
constexpr unsigned kHotCopyChunk = 32768;   // bytes per bulk copy

template <int C, bool WEIGHTED>
struct HotAccum {
    float4 acc[C];

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // one 16 B vector of a row: from the shared-memory image if the row is cached (warp-uniform predicate), else
    // from global memory — two predicated loads into the same registers, no select, no branch
    static __device__ __forceinline__ float4 ld_row(unsigned hot, unsigned saddr, const float4 *gptr) {
        float4 v;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %4, 0;\n\t"
            "@p ld.shared.v4.f32 {%0,%1,%2,%3}, [%5];\n\t"
            "@!p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%6];\n\t"
            "}"
            : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
            : "r"(hot), "r"(saddr), "l"(gptr));
        return v;
    }

    // N rows whose table-relative ids sit in lanes j .. j+N-1 of my_rel
    template <int N>
    __device__ __forceinline__ void batch(const float4 *const (&gcol)[C], const unsigned (&scol)[C],
                                          unsigned vec4, unsigned hot_n, unsigned my_rel, float my_w, int j) {
        float4 v[N][C];
        float wv[N];
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const unsigned rel = __shfl_sync(0xffffffffu, my_rel, j + u);
            if (WEIGHTED) wv[u] = __shfl_sync(0xffffffffu, my_w, j + u);
            const unsigned hot = rel < hot_n ? 1u : 0u;                  // warp-uniform
            const unsigned soff = (hot ? rel : 0u) * vec4 * 16u;         // cached rows only
#pragma unroll
            for (int c = 0; c < C; ++c)
                v[u][c] = ld_row(hot, scol[c] + soff, gcol[c] + (unsigned long long)rel * vec4);
        }
#pragma unroll
        for (int u = 0; u < N; ++u) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (WEIGHTED) {
                    fma2(acc[c].x, acc[c].y, wv[u], v[u][c].x, v[u][c].y);
                    fma2(acc[c].z, acc[c].w, wv[u], v[u][c].z, v[u][c].w);
                } else {
                    add2(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                    add2(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                }
            }
        }
    }
};

template <typename index_t, int C, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1) tbe_fwd_hot_kernel(const FwdParams p, int hot_cap) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    // rows in flight per lane: 4 (x 32 warps per SM = 64 KB of requests in flight, several times the
    // latency-bandwidth product of one SM's share of HBM); 8 as in DIRECT does not fit 64 registers
    constexpr int U = (C == 1) ? 4 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned vec4 = (unsigned)(p.dim >> 2);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    unsigned phase = 0;
    const index_t *off = (const index_t *)p.offsets;
    const index_t *idx = (const index_t *)p.indices;
    const unsigned s_base = smem_u32(s_raw);
    // Every CTA walks the tables in the same order and takes every gridDim-th group of 32 bags of the current
    // table: at any moment all SMs gather from the same table, so its warm rows (beyond the cached head) are
    // shared through L2 exactly as under DIRECT's linear block order.  (The first version gave each CTA one
    // contiguous range of bags: 148 CTAs in 148 different places of the arena, L2 hit rate 12 % instead of
    // 55 %, 2.6x the DRAM reads — profiles/r02f_ncu_fwd_hot_v1.md.)
    const unsigned stride = gridDim.x * kHotWarps;
    for (int t = 0; t < p.num_tables; ++t) {
        const long long base_row = p.table_row_offsets[t];
        const long long rows_t = p.table_row_offsets[t + 1] - base_row;
        const unsigned hot_n = (unsigned)min((long long)hot_cap, rows_t);
        __syncthreads();                    // every warp is done with the previous table's cache
        if (hot_n > 0) {
            if (threadIdx.x == 0) {
                const unsigned bytes = hot_n * vec4 * 16u;
                const unsigned char *src = (const unsigned char *)(p.weights + base_row * p.dim);
                mbar_arrive_expect_tx(&s_bar, bytes);
                for (unsigned o = 0; o < bytes; o += kHotCopyChunk)
                    bulk_g2s(s_raw + o, src + o, min(kHotCopyChunk, bytes - o), &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
        }
        const float4 *gtab = (const float4 *)p.weights + (unsigned long long)base_row * vec4;
        const float4 *gcol[C];
        unsigned scol[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const unsigned col = (unsigned)(c * 32 + lane);
            gcol[c] = gtab + (col < vec4 ? col : 0);
            scol[c] = s_base + (col < vec4 ? col : 0) * 16u;
        }
        const index_t *off_t = off + (long long)t * p.batch;
        const unsigned n_here = (unsigned)p.batch;
        float *out_t = p.out + t * p.out_stride_t;
        for (unsigned r = blockIdx.x * kHotWarps + warp; r < n_here; r += stride) {
            const long long begin = ld_index<index_t>(off_t + r);
            const int len = (int)(ld_index<index_t>(off_t + r + 1) - begin);
            const index_t *ip = idx + begin + lane;
            HotAccum<C, WEIGHTED> a;
            a.zero();
            for (int base = 0; base < len; base += 32) {
                unsigned my_rel = 0;
                float my_w = 0.f;
                if (base + lane < len) {
                    my_rel = (unsigned)ld_index<index_t>(ip + base);
                    if (WEIGHTED) my_w = ld_stream_f32(p.psw + begin + base + lane);
                }
                const int cnt = min(32, len - base);
                int j = 0;
                for (; j + U <= cnt; j += U) a.template batch<U>(gcol, scol, vec4, hot_n, my_rel, my_w, j);
                if (U > 2 && j + 2 <= cnt) {
                    a.template batch<2>(gcol, scol, vec4, hot_n, my_rel, my_w, j);
                    j += 2;
                }
                for (; j < cnt; ++j) a.template batch<1>(gcol, scol, vec4, hot_n, my_rel, my_w, j);
            }
            // store (same epilogue as BagAccum::store)
            const float cntf = (float)(len > 0 ? len : 1);
            float4 *o = (float4 *)(out_t + (long long)r * p.out_stride_b);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned col = (unsigned)(c * 32 + lane);
                if (col < vec4) {
                    float4 rr = a.acc[c];
                    if (p.mean) {
                        rr.x = __fdiv_rn(rr.x, cntf); rr.y = __fdiv_rn(rr.y, cntf);
                        rr.z = __fdiv_rn(rr.z, cntf); rr.w = __fdiv_rn(rr.w, cntf);
                    }
                    st_stream_f4(o + col, rr);
                }
            }
        }
    }
}

static int hot_rows_from_env() {
    static const int v = [] {
        const char *e = getenv("PB200_FWD_HOT_ROWS");   // rows of every table kept in shared memory
        return e ? atoi(e) : 256;
    }();
    return v;
}

template <typename index_t, int C>
static int launch_hot(const FwdParams &p, cudaStream_t st) {
    const size_t row_bytes = (size_t)p.dim * 4;
    long long k = hot_rows_from_env();
    const long long fit = (long long)((160 * 1024) / row_bytes);     // leave ~64 KB of the 228 KB to L1
    if (k > fit) k = fit;
    if (k < 0) k = 0;
    const size_t smem = (size_t)k * row_bytes;
    auto kern = p.psw ? tbe_fwd_hot_kernel<index_t, C, true> : tbe_fwd_hot_kernel<index_t, C, false>;
    if (smem > 48 * 1024)
        PB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = sm_count();                                     // persistent: one CTA per SM
    const long long groups = (p.batch + kHotWarps - 1) / kHotWarps;  // groups of 32 bags per table
    if (grid > groups) grid = groups;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kHotThreads, smem, st>>>(p, (int)k);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

// entry used by emb_fwd.cu's dispatcher: fp32 tables, dim % 4 == 0, 64 < dim <= 256, TBE layout
int launch_fwd_hot(const FwdParams &p, int idx_type, cudaStream_t st) {
    const int vec4 = p.dim >> 2;
    // dim > 256 would need four accumulator vectors per lane: that does not fit the 64 registers a
    // 1024-thread CTA leaves per thread — those shapes stay on DIRECT
    if (!p.table_row_offsets || vec4 <= 16 || vec4 > 64 || p.weights_f16 || !p.has_last_offset)
        return PB200_EUNSUPPORTED;
    if (idx_type == PB200_IDX_I64) {
        if (vec4 <= 32) return launch_hot<long long, 1>(p, st);
        return launch_hot<long long, 2>(p, st);
    }
    if (vec4 <= 32) return launch_hot<int, 1>(p, st);
    return launch_hot<int, 2>(p, st);
}

}  // namespace pb200
