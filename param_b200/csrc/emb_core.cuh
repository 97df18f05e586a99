// emb_core.cuh — the gather/accumulate core shared by the forward kernels (emb_fwd.cu) and the fused
// lookup + all-to-all kernel (fused_fwd_a2a.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace pb200 {

struct FwdParams {
    const float *weights;                // fp32 arena; the fp16 kernels reinterpret it as __half
    const long long *table_row_offsets;  // device [T+1] or nullptr (single table at row 0)
    const void *indices;
    const void *offsets;
    const float *psw;
    float *out;
    long long n_indices;
    long long batch;        // bags per table
    long long n_bags;       // T * batch
    long long out_stride_t;
    long long out_stride_b;
    int num_tables;
    int dim;
    int has_last_offset;    // offsets has n_bags + 1 entries
    int mean;
    int stage_cap;          // STAGED: index elements per stage buffer
    int weights_f16;        // table elements are __half (DIRECT variant only)
};

template <typename index_t>
__device__ __forceinline__ void bag_range(const FwdParams &p, long long gb, long long &begin,
                                          long long &end) {
    const index_t *off = (const index_t *)p.offsets;
    begin = ld_index<index_t>(off + gb);
    end = (gb + 1 < p.n_bags || p.has_last_offset) ? ld_index<index_t>(off + gb + 1) : p.n_indices;
}

__device__ __forceinline__ void split_bag(const FwdParams &p, long long gb, int &t, long long &b) {
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

// Packed fp32 adds (Blackwell FADD2 / FFMA2): two IEEE round-to-nearest fp32 lanes per instruction,
// bit-identical to two scalar adds, half the issue slots.
__device__ __forceinline__ void add2(float &a0, float &a1, float b0, float b1) {
    asm("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; "
        "mov.b64 {%0,%1}, ra; }"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float &a0, float &a1, float w, float b0, float b1) {
    asm("{ .reg .b64 ra, rb, rw; mov.b64 ra, {%0,%1}; mov.b64 rb, {%3,%4}; mov.b64 rw, {%2,%2}; "
        "fma.rn.f32x2 ra, rw, rb, ra; mov.b64 {%0,%1}, ra; }"
        : "+f"(a0), "+f"(a1)
        : "f"(w), "f"(b0), "f"(b1));
}

// A table row is read in vectors of 4 elements: 16 B of an fp32 table, 8 B of an fp16 table (converted
// to fp32 on load; accumulation is always fp32).
template <typename WT>
struct RowVec;
template <>
struct RowVec<float> {
    using type = float4;
    static __device__ __forceinline__ float4 ld(const float4 *p) { return ld_row_f4(p); }
};
template <>
struct RowVec<__half> {
    using type = uint2;
    static __device__ __forceinline__ float4 ld(const uint2 *p) {
        uint2 raw;
        asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
        const float2 lo = __half22float2(*(const __half2 *)&raw.x);
        const float2 hi = __half22float2(*(const __half2 *)&raw.y);
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
};

// Accumulate one bag (or, for G < 32, 32/G bags side by side) given per-group [begin, end).
// The hot loop is branch- and predicate-free: rows are consumed in batches of 8/4/2/1 whose size is
// warp-uniform, every lane issues its loads unconditionally (lanes beyond dim/4 read column 0 and
// drop the result at the store), and only groups shorter than the longest bag of the warp mask
// their tail.  Per row and warp that is 1 SHFL + 1 IMAD.WIDE + 1 LDG.128 + 2 FADD2.
template <typename index_t, int G, int C, bool WEIGHTED, int U, typename WT = float>
struct BagAccum {
    using V = typename RowVec<WT>::type;
    float4 acc[C];

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // N rows starting at lane j of the group; MASKED: rows at or beyond `valid` are dropped
    template <int N, bool MASKED>
    __device__ __forceinline__ void batch(const V *const (&colp)[C], unsigned row_stride4,
                                          unsigned my_row, float my_w, int j, int valid) {
        float4 v[N][C];
        float wv[N];
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const unsigned row = __shfl_sync(0xffffffffu, my_row, j + u, G);
            if (WEIGHTED) wv[u] = __shfl_sync(0xffffffffu, my_w, j + u, G);
            const unsigned long long roff = (unsigned long long)row * row_stride4;
#pragma unroll
            for (int c = 0; c < C; ++c) v[u][c] = RowVec<WT>::ld(colp[c] + roff);
        }
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const bool drop = MASKED && (j + u >= valid);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float4 x = v[u][c];
                if (MASKED && drop) x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (WEIGHTED) {
                    const float w = (MASKED && drop) ? 0.f : wv[u];
                    fma2(acc[c].x, acc[c].y, w, x.x, x.y);
                    fma2(acc[c].z, acc[c].w, w, x.z, x.w);
                } else {
                    add2(acc[c].x, acc[c].y, x.x, x.y);
                    add2(acc[c].z, acc[c].w, x.z, x.w);
                }
            }
        }
    }

    template <bool MASKED>
    __device__ __forceinline__ void span(const V *const (&colp)[C], unsigned row_stride4,
                                         unsigned my_row, float my_w, int j, int end, int valid) {
        for (; j + U <= end; j += U) batch<U, MASKED>(colp, row_stride4, my_row, my_w, j, valid);
        if (U > 4 && j + 4 <= end) {
            batch<4, MASKED>(colp, row_stride4, my_row, my_w, j, valid);
            j += 4;
        }
        if (U > 2 && j + 2 <= end) {
            batch<2, MASKED>(colp, row_stride4, my_row, my_w, j, valid);
            j += 2;
        }
        for (; j < end; ++j) batch<1, MASKED>(colp, row_stride4, my_row, my_w, j, valid);
    }

    // idx_ptr: pointer to this group's first index (global or shared); len: this group's bag
    // length; minlen / maxlen: warp-uniform min / max over the groups of the warp.
    template <bool FROM_SMEM, bool PRELOADED = false>
    __device__ __forceinline__ void run(const FwdParams &p, const index_t *idx_ptr,
                                        const float *psw_ptr, long long base_row, int len,
                                        int minlen, int maxlen, int lane_g, int vec4,
                                        unsigned pre_row = 0, float pre_w = 0.f) {
        const unsigned row_stride4 = (unsigned)vec4;
        const V *colp[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int col = c * G + lane_g;
            colp[c] = (const V *)p.weights + (col < vec4 ? col : 0);
        }
        // one coalesced read of up to G indices per group and chunk; out-of-bag lanes keep row 0 of the
        // group's table, which is a valid address and masked out below.  The chunk after the current one is
        // requested before the current chunk's rows are gathered, so a bag longer than G pays the index
        // latency once, not once per chunk.
        auto load_chunk = [&](int base, unsigned &row, float &w) {
            row = (unsigned)base_row;
            w = 0.f;
            if (base + lane_g < len) {
                long long ix;
                if (FROM_SMEM)
                    ix = (long long)idx_ptr[base + lane_g];
                else
                    ix = ld_index<index_t>(idx_ptr + base + lane_g);
                row = (unsigned)(base_row + ix);
                if (WEIGHTED) w = ld_stream_f32(psw_ptr + base + lane_g);
            }
        };
        unsigned next_row = (unsigned)base_row;
        float next_w = 0.f;
        if (PRELOADED) {
            // first chunk of indices was fetched one bag ahead (software pipeline)
            next_row = pre_row;
            next_w = pre_w;
        } else if (maxlen > 0) {
            load_chunk(0, next_row, next_w);
        }
        for (int base = 0; base < maxlen; base += G) {
            const unsigned my_row = next_row;
            const float my_w = next_w;
            if (base + G < maxlen) load_chunk(base + G, next_row, next_w);
            const int full = min(G, minlen - base);   // rows every group of the warp still has
            const int most = min(G, maxlen - base);   // rows the longest group still has
            int j = 0;
            if (full > 0) {
                span<false>(colp, row_stride4, my_row, my_w, 0, full, 0);
                j = full;
            }
            if (j < most) span<true>(colp, row_stride4, my_row, my_w, j, most, len - base);
        }
    }

    __device__ __forceinline__ void store(const FwdParams &p, int t, long long b, int len,
                                          int lane_g, int vec4) {
        const float cnt = (float)(len > 0 ? len : 1);
        float4 *o = (float4 *)(p.out + (long long)t * p.out_stride_t + b * p.out_stride_b);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int col = c * G + lane_g;
            if (col < vec4) {
                float4 r = acc[c];
                if (p.mean) {  // true division, as ATen's mean does (sum / bag_size)
                    r.x = __fdiv_rn(r.x, cnt); r.y = __fdiv_rn(r.y, cnt);
                    r.z = __fdiv_rn(r.z, cnt); r.w = __fdiv_rn(r.w, cnt);
                }
                st_stream_f4(o + col, r);
            }
        }
    }
};

template <int C>
struct UnrollFor {
    static constexpr int value = (C == 1) ? 8 : (C == 2 ? 4 : 2);
};

}  // namespace pb200
