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
//   * persistent grid, ONE CTA of 32 warps per SM, each CTA owns a contiguous range of bags (table-major), so
//     it changes table two or three times in its life;
//   * on a table change the CTA copies the head of the table (K rows, 128 KB at dim 128) into shared memory
//     with cp.async.bulk (TMA unit, SASS UBLKCP) on an mbarrier — one copy per SM instead of one L1 image
//     per resident CTA;
//   * a lookup whose row is < K is served by LDS.128, the others by the same LDG.128 path as DIRECT.  With
//     G = 32 (dim 65..256) the row of a lookup is warp-uniform, so the choice is a uniform predicate: no
//     divergence, loads stay batched 8 deep, adds stay in index order (bit-identical sums).
// A table whose popular rows are elsewhere loses nothing but the L1 capacity handed to the cache.
#include <stdlib.h>

#include "emb_core.cuh"

namespace pb200 {

constexpr int kHotThreads = 1024;
constexpr int kHotWarps = kHotThreads / 32;
constexpr unsigned kHotCopyChunk = 32768;   // bytes per bulk copy

template <int C, bool WEIGHTED>
struct HotAccum {
    float4 acc[C];

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // N rows whose table-relative ids sit in lanes j .. j+N-1 of my_rel
    template <int N>
    __device__ __forceinline__ void batch(const float4 *const (&gcol)[C], const float4 *const (&scol)[C],
                                          unsigned vec4, unsigned hot_n, unsigned my_rel, float my_w, int j) {
        float4 v[N][C];
        float wv[N];
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const unsigned rel = __shfl_sync(0xffffffffu, my_rel, j + u);
            if (WEIGHTED) wv[u] = __shfl_sync(0xffffffffu, my_w, j + u);
            const unsigned roff = rel * vec4;          // used for cached rows only (rel < hot_n)
            if (rel < hot_n) {                          // warp-uniform
#pragma unroll
                for (int c = 0; c < C; ++c) v[u][c] = scol[c][roff];
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) v[u][c] = ld_row_f4(gcol[c] + (unsigned long long)rel * vec4);
            }
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
__global__ void __launch_bounds__(kHotThreads, 1) tbe_fwd_hot_kernel(const FwdParams p, int hot_cap,
                                                                     long long bags_per_cta) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    // rows in flight per lane: 4 (x 32 warps per SM = 64 KB of requests in flight, several times the
    // latency-bandwidth product of one SM's share of HBM); 8 as in DIRECT does not fit 64 registers
    constexpr int U = (C == 1) ? 4 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned vec4 = (unsigned)(p.dim >> 2);
    const long long gb0 = (long long)blockIdx.x * bags_per_cta;
    const long long gb1 = min(p.n_bags, gb0 + bags_per_cta);
    if (gb0 >= gb1) return;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    unsigned phase = 0;
    const index_t *off = (const index_t *)p.offsets;
    const index_t *idx = (const index_t *)p.indices;
    const float4 *s_hot = (const float4 *)s_raw;

    for (long long t = gb0 / p.batch; t * p.batch < gb1; ++t) {
        const long long lo = max(gb0, t * p.batch);
        const long long hi = min(gb1, (t + 1) * p.batch);
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
        const float4 *scol[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const unsigned col = (unsigned)(c * 32 + lane);
            gcol[c] = gtab + (col < vec4 ? col : 0);
            scol[c] = s_hot + (col < vec4 ? col : 0);
        }
        const index_t *off_t = off + lo;
        const unsigned n_here = (unsigned)(hi - lo);
        float *out_t = p.out + t * p.out_stride_t + (lo - t * p.batch) * p.out_stride_b;
        for (unsigned r = warp; r < n_here; r += kHotWarps) {
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
                if (U > 4 && j + 4 <= cnt) {
                    a.template batch<4>(gcol, scol, vec4, hot_n, my_rel, my_w, j);
                    j += 4;
                }
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
                    float4 r = a.acc[c];
                    if (p.mean) {
                        r.x = __fdiv_rn(r.x, cntf); r.y = __fdiv_rn(r.y, cntf);
                        r.z = __fdiv_rn(r.z, cntf); r.w = __fdiv_rn(r.w, cntf);
                    }
                    st_stream_f4(o + col, r);
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
    const long long min_bags = 4 * kHotWarps;
    if (grid * min_bags > p.n_bags) grid = (p.n_bags + min_bags - 1) / min_bags;
    if (grid < 1) grid = 1;
    const long long per = (p.n_bags + grid - 1) / grid;
    kern<<<(unsigned)grid, kHotThreads, smem, st>>>(p, (int)k, per);
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
