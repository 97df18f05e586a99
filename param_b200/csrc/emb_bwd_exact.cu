// emb_bwd_exact.cu — EmbeddingBag backward with the optimizer fused in, every touched table row
// written EXACTLY ONCE (no atomics, run-to-run deterministic), fp32 or fp16 tables (sm_100a).
//
// Replaces the fused backward+optimizer of fbgemm's SplitTableBatchedEmbeddingBagsCodegen as PARAM
// builds it:
//   train/comms/pt/comms_utils.py:1995-2017        optimizer=OptimType.EXACT_ROWWISE_ADAGRAD
//   train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:279-301
//                                                   optimizer / weights_precision / lr / eps /
//                                                   stochastic_rounding=True from the op config
//   backward driven from pytorch_dist_backend.py:849-857 and ...CodegenOp.backward :318-324.
// (fbgemm_gpu itself is absent from the reference tree: the arithmetic follows its published
// "exact" optimizers — the gradient of a row is summed over ALL its lookups of the batch first,
// then the update is applied once.  SGD: w -= lr*g.  Rowwise Adagrad: m += mean_d(g_d^2);
// w -= lr / (sqrt(m) + eps) * g.  Parity for this op is unpinned)
//
// A non-linear update needs the COMPLETE gradient of a row, so the "one red per (segment, row)" of
// the SORTED variant (emb_bwd.cu) is not enough.  Over the same sort plan (sort_plan.cuh, radix_sort.cu — all
// lookups of the request sorted by arena row), one launch each:
//   E1 exact_reduce_kernel   one lane group per 256 sorted entries, runs of equal rows summed in
//                            registers.  A run that lies wholly inside the segment is final: the
//                            optimizer is applied on the spot (row read once, written once).  The
//                            first run if it continues the previous segment's last row, and the
//                            last run if it continues into the next segment, go to a partial-sum
//                            buffer instead ([segment][head|tail][dim] fp32).
//   E2 exact_boundary_kernel one lane group per segment whose tail run STARTS a multi-segment
//                            run: tail partial + head partials of the following segments while
//                            they begin with the same row, then the update; runs longer than 8
//                            segments go to a work list.
//   E3 exact_long_run_kernel persistent CTAs, one listed run at a time, the CTA's lane groups
//                            walking its segments interleaved (the hottest Zipf row spans
//                            hundreds of segments).
// Summation order is fixed by the sort (stable radix sort of a fixed pair order) and the segment
// structure, so two runs on the same inputs give identical bits.
#include <cuda_fp16.h>

#include "emb_bwd_common.cuh"

namespace pb200 {

struct OptParams {
    void *weights;               // arena base, fp32 or fp16 [rows, dim]
    float *state;                // ROWWISE_ADAGRAD: running sum of mean squared gradients, per arena row
    float lr;
    float eps;
    int optimizer;
    int stochastic;              // fp16 tables: stochastic rounding of the updated weights
    unsigned long long sr_seed;
};

// ---- row access in units of 4 elements -------------------------------------------------------
template <typename WT>
struct Row4;
template <>
struct Row4<float> {
    static __device__ __forceinline__ float4 load(const void *base, unsigned long long v) {
        return *((const float4 *)base + v);
    }
    static __device__ __forceinline__ void store(void *base, unsigned long long v, const float4 &x,
                                                 const OptParams &) {
        *((float4 *)base + v) = x;
    }
};
template <>
struct Row4<__half> {
    static __device__ __forceinline__ float4 load(const void *base, unsigned long long v) {
        const uint2 raw = *((const uint2 *)base + v);
        const float2 lo = __half22float2(*(const __half2 *)&raw.x);
        const float2 hi = __half22float2(*(const __half2 *)&raw.y);
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    static __device__ __forceinline__ void store(void *base, unsigned long long v, const float4 &x,
                                                 const OptParams &op) {
        uint2 raw;
        if (op.stochastic) {
            // add 13 random bits below the fp16 mantissa, then truncate (the usual fp32->fp16
            // stochastic rounding; like fbgemm it ignores the fp16 subnormal range)
            const unsigned long long r = mix64(op.sr_seed ^ (v * 0x9E3779B97F4A7C15ull));
            const float a = __uint_as_float(__float_as_uint(x.x) + (unsigned)(r & 0x1fff));
            const float b = __uint_as_float(__float_as_uint(x.y) + (unsigned)((r >> 13) & 0x1fff));
            const float c = __uint_as_float(__float_as_uint(x.z) + (unsigned)((r >> 26) & 0x1fff));
            const float d = __uint_as_float(__float_as_uint(x.w) + (unsigned)((r >> 39) & 0x1fff));
            const __half2 lo = __halves2half2(__float2half_rz(a), __float2half_rz(b));
            const __half2 hi = __halves2half2(__float2half_rz(c), __float2half_rz(d));
            raw.x = *(const unsigned *)&lo;
            raw.y = *(const unsigned *)&hi;
        } else {
            const __half2 lo = __floats2half2_rn(x.x, x.y);
            const __half2 hi = __floats2half2_rn(x.z, x.w);
            raw.x = *(const unsigned *)&lo;
            raw.y = *(const unsigned *)&hi;
        }
        *((uint2 *)base + v) = raw;
    }
};

// The optimizer step of ONE row whose complete gradient g (acc, spread over the G lanes of a lane
// group) is known.  Called with the lanes of one group converged; gmask names exactly those lanes.
template <typename WT, int G, int C>
__device__ __forceinline__ void apply_row_update(const OptParams &op, unsigned long long row,
                                                 const float4 (&acc)[C], const bool (&col_ok)[C],
                                                 int lane_g, int vec4, int dim, unsigned gmask) {
    float mult = op.lr;
    if (op.optimizer == PB200_OPT_ROWWISE_ADAGRAD) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (col_ok[c]) {
                ss = fmaf(acc[c].x, acc[c].x, ss);
                ss = fmaf(acc[c].y, acc[c].y, ss);
                ss = fmaf(acc[c].z, acc[c].z, ss);
                ss = fmaf(acc[c].w, acc[c].w, ss);
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(gmask, ss, o, G);
        float m = 0.f;
        if (lane_g == 0) {
            m = op.state[row] + ss / (float)dim;
            op.state[row] = m;
        }
        m = __shfl_sync(gmask, m, 0, G);
        mult = op.lr / (sqrtf(m) + op.eps);
    }
    const unsigned long long v0 = row * (unsigned)vec4;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (col_ok[c]) {
            const unsigned long long v = v0 + (unsigned)(c * G + lane_g);
            float4 w = Row4<WT>::load(op.weights, v);
            w.x = fmaf(-mult, acc[c].x, w.x);
            w.y = fmaf(-mult, acc[c].y, w.y);
            w.z = fmaf(-mult, acc[c].z, w.z);
            w.w = fmaf(-mult, acc[c].w, w.w);
            Row4<WT>::store(op.weights, v, w, op);
        }
    }
}

// ---- E1 ----------------------------------------------------------------------------------------
// What a run's update needs besides its gradient sum — the Adagrad state of the row, and for fp16
// tables the old weights (they must be rounded once, so no red.add) — is PREFETCHED together with
// the gradient rows of the batch, for the entries that can end a run (the key changes after them, or
// they are the last entry of the batch).  A dependent load at the end of every run would serialise
// on DRAM latency when most rows are hit once (uniform indices: measured 16.5 ms vs 11.9 ms for the
// SORTED variant at 64 tables); fp32 rows are updated with one red.global.add.v4.f32 per lane, which
// needs no load at all and is still deterministic because every row is written exactly once.
template <typename WT>
struct OldRow;                       // the prefetched old weights of one row vector
template <>
struct OldRow<float> {
    static constexpr bool kNeeded = false;
    struct raw {};
    static __device__ __forceinline__ raw load(const void *, unsigned long long) { return raw{}; }
};
template <>
struct OldRow<__half> {
    static constexpr bool kNeeded = true;
    using raw = uint2;
    static __device__ __forceinline__ raw load(const void *base, unsigned long long v) {
        return *((const uint2 *)base + v);
    }
};

template <typename WT, int OPT, int G, int C, bool SIDE>
__device__ __forceinline__ void exact_reduce_body(const BwdParams &p, const OptParams &op, long long n,
                                                  long long chunk_row0,
                                                  const unsigned *__restrict__ keys,
                                                  const unsigned *__restrict__ vals,
                                                  const unsigned *__restrict__ goff_of,
                                                  const float *__restrict__ w_of,
                                                  float4 *__restrict__ partial, int seg_len) {
    constexpr int BPW = 32 / G;
    constexpr int U = (C == 1) ? 8 : (C == 2 ? 4 : 2);
    constexpr unsigned kNoKey = 0xffffffffu;   // no chunk-relative row has this id
    constexpr bool kAdagrad = OPT == PB200_OPT_ROWWISE_ADAGRAD;
    constexpr bool kOldW = OldRow<WT>::kNeeded;
    using OldW = typename OldRow<WT>::raw;
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int vec4 = p.dim >> 2;
    n = min(n, *p.n_dev);    // n is the host's capacity, the plan knows the count
    const long long seg = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BPW + grp;
    const long long s0 = seg * seg_len;
    const long long s1 = min(s0 + (long long)seg_len, n);
    const int my_n = (s0 < n) ? (int)(s1 - s0) : 0;
    const int max_n = (BPW == 1) ? my_n : __reduce_max_sync(0xffffffffu, my_n);

    // does my first run continue the previous segment's last row / my last run continue into the next?
    bool head_cont = false, tail_cont = false;
    if (my_n > 0) {
        head_cont = s0 > 0 && keys[s0 - 1] == keys[s0];
        tail_cont = s1 < n && keys[s1] == keys[s1 - 1];
    }
    bool first_run = true;

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
    unsigned cur_key = kNoKey;
    float cur_state = 0.f;      // Adagrad state of cur_key's row, as prefetched with its latest entry
    OldW cur_w[C];              // fp16: old weights of cur_key's row

    auto flush = [&](bool last) {
        if (cur_key != kNoKey) {
            const int which = (first_run && head_cont) ? 0 : ((last && tail_cont) ? 1 : -1);
            if (which >= 0) {
                float4 *pp = partial + ((unsigned long long)seg * 2 + which) * (unsigned)vec4;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (col_ok[c]) pp[c * G + lane_g] = acc[c];
            } else {
                const unsigned long long row = (unsigned long long)chunk_row0 + cur_key;
                float mult = op.lr;
                if (kAdagrad) {
                    float ss = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if (col_ok[c]) {
                            ss = fmaf(acc[c].x, acc[c].x, ss);
                            ss = fmaf(acc[c].y, acc[c].y, ss);
                            ss = fmaf(acc[c].z, acc[c].z, ss);
                            ss = fmaf(acc[c].w, acc[c].w, ss);
                        }
                    }
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(gmask, ss, o, G);
                    const float m = cur_state + ss / (float)p.dim;
                    if (lane_g == 0) op.state[row] = m;
                    mult = op.lr / (sqrtf(m) + op.eps);
                }
                const unsigned long long v0 = row * (unsigned)vec4;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (col_ok[c]) {
                        const unsigned long long v = v0 + (unsigned)(c * G + lane_g);
                        if constexpr (kOldW) {
                            const uint2 raw = cur_w[c];
                            const float2 lo = __half22float2(*(const __half2 *)&raw.x);
                            const float2 hi = __half22float2(*(const __half2 *)&raw.y);
                            float4 w = make_float4(lo.x, lo.y, hi.x, hi.y);
                            w.x = fmaf(-mult, acc[c].x, w.x);
                            w.y = fmaf(-mult, acc[c].y, w.y);
                            w.z = fmaf(-mult, acc[c].z, w.z);
                            w.w = fmaf(-mult, acc[c].w, w.w);
                            Row4<WT>::store(op.weights, v, w, op);
                        } else {
                            float4 d = acc[c];
                            d.x *= -mult; d.y *= -mult; d.z *= -mult; d.w *= -mult;
                            red_add_f4((float4 *)op.weights + v, d);
                        }
                    }
                }
            }
            first_run = false;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
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
            float st[U];
            OldW ow[U][C];
#pragma unroll
            for (int u = 0; u < U; ++u) kk[u] = __shfl_sync(0xffffffffu, my_key, (j0 + u) & (G - 1), G);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = (j0 + u) & (G - 1);
                const unsigned goff = __shfl_sync(0xffffffffu, my_goff, src, G);
                if (SIDE) ww[u] = __shfl_sync(0xffffffffu, my_w, src, G);
#pragma unroll
                for (int c = 0; c < C; ++c) v[u][c] = ld_row_f4(colp[c] + goff);
                st[u] = 0.f;
                if constexpr (kAdagrad || kOldW) {
                    // entry u can end a run if it is valid and the next entry is invalid, outside
                    // this batch / G-block, or has another key
                    const bool ok_u = (j0 + u < valid) && (j0 + u < G);
                    const bool ok_n = (u + 1 < U) && (j0 + u + 1 < valid) && (j0 + u + 1 < G);
                    const bool ends = ok_u && (!ok_n || kk[u + 1 < U ? u + 1 : u] != kk[u]);
                    const unsigned long long row = (unsigned long long)chunk_row0 + kk[u];
                    if (kAdagrad && ends) st[u] = op.state[row];
                    if constexpr (kOldW) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            ow[u][c] = OldW{};
                            if (ends && col_ok[c])
                                ow[u][c] = OldRow<WT>::load(op.weights, row * (unsigned)vec4 + (unsigned)(c * G + lane_g));
                        }
                    }
                }
            }
            const bool all_valid = (j0 + U <= valid) && (j0 + U <= G);
            const bool same = kk[U - 1] == kk[0];   // keys are fully sorted: first == last => all equal
            if (!SIDE && all_valid && same && (kk[0] == cur_key || cur_key == kNoKey)) {
                cur_key = kk[0];
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        add2b(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                        add2b(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                    }
                }
                if (kAdagrad) cur_state = st[U - 1];
                if constexpr (kOldW) {
#pragma unroll
                    for (int c = 0; c < C; ++c) cur_w[c] = ow[U - 1][c];
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (j0 + u < valid && j0 + u < G) {
                        if (kk[u] != cur_key) {
                            flush(false);
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
                        // the prefetch of this entry is the freshest view of cur_key's row
                        if (kAdagrad) cur_state = st[u];
                        if constexpr (kOldW) {
#pragma unroll
                            for (int c = 0; c < C; ++c) cur_w[c] = ow[u][c];
                        }
                    }
                }
            }
        }
    }
    flush(true);
}

template <typename WT, int OPT, int G, int C, bool SIDE>
__global__ void __launch_bounds__(256) exact_reduce_kernel(const BwdParams p, const OptParams op,
                                                           long long n, long long chunk_row0,
                                                           const unsigned *__restrict__ keys,
                                                           const unsigned *__restrict__ vals,
                                                           const unsigned *__restrict__ goff_of,
                                                           const float *__restrict__ w_of,
                                                           float4 *__restrict__ partial, int seg_len) {
    exact_reduce_body<WT, OPT, G, C, SIDE>(p, op, n, chunk_row0, keys, vals, goff_of, w_of, partial, seg_len);
}
// same body compiled for 3 resident CTAs/SM (<= 85 registers; the Adagrad / fp16 variants otherwise
// use 90-102 and fit only 2): selectable with PB200_EXACT_OCC3=1, dim <= 128 without side arrays
template <typename WT, int OPT>
__global__ void __launch_bounds__(256, 3) exact_reduce_kernel_occ3(const BwdParams p, const OptParams op,
                                                                   long long n, long long chunk_row0,
                                                                   const unsigned *__restrict__ keys,
                                                                   const unsigned *__restrict__ vals,
                                                                   float4 *__restrict__ partial,
                                                                   int seg_len) {
    exact_reduce_body<WT, OPT, 32, 1, false>(p, op, n, chunk_row0, keys, vals, nullptr, nullptr, partial,
                                             seg_len);
}

// (A "slim" E1 that classifies each batch from its keys BEFORE issuing the loads — uniform batches
// with one prefetch and no predicates, mixed batches with compile-time-indexed prefetches — executed
// 30 % fewer warp instructions but ran slower: 9.77 vs 8.41 ms Adagrad at 64 tables, 73 ms for fp16.
// Duplicating the load sequences and the inlined flush grew the kernel to 7 184 SASS instructions
// (instruction-cache misses, issue slots 30 % busy) and to 156-316 B of spills; removed, evidence in
// profiles/r01i_*.)

// ---- E2 / E3 -----------------------------------------------------------------------------------
// E2: one lane group per segment whose tail run STARTS a multi-segment run.  A run that ends within
// the next U segments (almost all of them) is finished on the spot: tail partial + head partials of
// the following segments, then the update.  A longer run — a hot Zipf row spans up to 256 segments
// at batch 65536 — would be a chain of dependent steps for one lane group, so its first segment is
// appended to a work list instead.
// E3: persistent CTAs take the work list; the 256/G lane groups of a CTA walk a run's segments
// interleaved (group w takes segments seg+1+w, seg+1+w+NW, ...; U head partials in flight each) and
// combine through shared memory in group order.  The order in which runs are listed varies from
// launch to launch, the summation order inside a run does not: results stay deterministic.
// (Measured, profiles/r01e-r01f: E2 alone left 6.3 ms vs 3.7 ms SORTED when every chunk held one
// 10 M-row table; one CTA per segment instead cost 1.3 ms per 64 tables in CTA launches.)
template <typename WT, int G, int C>
__global__ void __launch_bounds__(256) exact_boundary_kernel(const BwdParams p, const OptParams op,
                                                             long long n, long long chunk_row0,
                                                             const unsigned *__restrict__ keys,
                                                             const float4 *__restrict__ partial,
                                                             unsigned *__restrict__ worklist,
                                                             unsigned *__restrict__ work_count,
                                                             int seg_len) {
    constexpr int BPW = 32 / G;
    constexpr int U = (C == 1) ? 8 : (C == 2 ? 4 : 2);
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int vec4 = p.dim >> 2;
    n = min(n, *p.n_dev);
    const long long seg = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * BPW + grp;
    const long long s0 = seg * seg_len;
    if (s0 >= n) return;
    const long long s1 = min(s0 + (long long)seg_len, n);
    if (s1 >= n) return;                                   // the last segment has no successor
    const unsigned kl = keys[s1 - 1];
    if (keys[s1] != kl) return;                            // my last run ends with me
    // my last run is my only run AND a continuation from before: an earlier segment owns the row
    if (keys[s0] == kl && s0 > 0 && keys[s0 - 1] == kl) return;

    bool col_ok[C];
    float4 acc[C];
    const float4 *tail = partial + ((unsigned long long)seg * 2 + 1) * (unsigned)vec4;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int col = c * G + lane_g;
        col_ok[c] = col < vec4;
        acc[c] = col_ok[c] ? ld_stream_f4(tail + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 v[U][C];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long jj = seg + 1 + u;
        const long long pos = jj * seg_len;
        const bool in = pos < n;
        ok[u] = in && keys[in ? pos : 0] == kl;
        // the load is unconditional (keeps U requests in flight); a segment past the run reads a
        // head partial that is simply not used
        const float4 *head = partial + ((unsigned long long)(in ? jj : seg) * 2) * (unsigned)vec4;
#pragma unroll
        for (int c = 0; c < C; ++c)
            v[u][c] = col_ok[c] ? ld_stream_f4(head + c * G + lane_g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (ok[U - 1]) {
        // all U following segments begin with my row: a long run, left to a whole CTA (E3)
        if (lane_g == 0) worklist[atomicAdd(work_count, 1u)] = (unsigned)seg;
        return;
    }
    bool done = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!done) {
            if (ok[u]) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    add2b(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                    add2b(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                }
            } else {
                done = true;
            }
        }
    }
    apply_row_update<WT, G, C>(op, (unsigned long long)chunk_row0 + kl, acc, col_ok, lane_g, vec4,
                               p.dim, gmask);
}

template <typename WT, int G, int C>
__global__ void __launch_bounds__(256) exact_long_run_kernel(const BwdParams p, const OptParams op,
                                                             long long n, long long chunk_row0,
                                                             const unsigned *__restrict__ keys,
                                                             const float4 *__restrict__ partial,
                                                             const unsigned *__restrict__ worklist,
                                                             const unsigned *__restrict__ work_count,
                                                             int seg_len) {
    constexpr int BPW = 32 / G;
    constexpr int NW = 8 * BPW;          // lane groups per CTA
    constexpr int U = (C == 1) ? 8 : (C == 2 ? 4 : 2);
    __shared__ float4 s_part[256 * C];   // [NW][C*G]
    const int lane = threadIdx.x & 31;
    const int lane_g = lane & (G - 1);
    const int grp = lane / G;
    const int worker = (threadIdx.x >> 5) * BPW + grp;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int vec4 = p.dim >> 2;
    bool col_ok[C];
#pragma unroll
    for (int c = 0; c < C; ++c) col_ok[c] = c * G + lane_g < vec4;
    const unsigned count = *work_count;
    n = min(n, *p.n_dev);
    for (unsigned item = blockIdx.x; item < count; item += gridDim.x) {
        const long long seg = worklist[item];
        const long long s1 = min((seg + 1) * (long long)seg_len, n);
        const unsigned kl = keys[s1 - 1];
        float4 acc[C];
        const float4 *tail = partial + ((unsigned long long)seg * 2 + 1) * (unsigned)vec4;
#pragma unroll
        for (int c = 0; c < C; ++c)
            acc[c] = (worker == 0 && col_ok[c]) ? ld_stream_f4(tail + c * G + lane_g)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        bool done = false;
        for (long long j = seg + 1 + worker; !done; j += (long long)NW * U) {
            float4 v[U][C];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long jj = j + (long long)u * NW;
                const long long pos = jj * seg_len;
                const bool in = pos < n;
                ok[u] = in && keys[in ? pos : 0] == kl;
                const float4 *head = partial + ((unsigned long long)(in ? jj : seg) * 2) * (unsigned)vec4;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    v[u][c] = col_ok[c] ? ld_stream_f4(head + c * G + lane_g) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!done) {
                    if (ok[u]) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            add2b(acc[c].x, acc[c].y, v[u][c].x, v[u][c].y);
                            add2b(acc[c].z, acc[c].w, v[u][c].z, v[u][c].w);
                        }
                    } else {
                        done = true;
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) s_part[worker * (C * G) + c * G + lane_g] = acc[c];
        __syncthreads();
        if (worker == 0) {
#pragma unroll 1
            for (int w = 1; w < NW; ++w) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 x = s_part[w * (C * G) + c * G + lane_g];
                    add2b(acc[c].x, acc[c].y, x.x, x.y);
                    add2b(acc[c].z, acc[c].w, x.z, x.w);
                }
            }
            apply_row_update<WT, G, C>(op, (unsigned long long)chunk_row0 + kl, acc, col_ok, lane_g, vec4,
                                       p.dim, gmask);
        }
        __syncthreads();   // s_part is reused by the next item
    }
}

// per segment: head + tail partial sums ([2][dim] fp32) and 8 bytes of work-list space
static PlanLayout plan_exact(long long n_indices, int num_tables, long long batch, int dim, int seg_len) {
    return plan_layout(n_indices, num_tables, batch, (size_t)2 * (size_t)dim * 4 + 8, seg_len);
}

template <typename WT>
static int bwd_exact(const BwdParams &p, const OptParams &op, int idx_type, long long max_table_rows,
                     void *scratch, long long scratch_bytes, bool plan_ready, cudaStream_t s) {
    const bool side = (p.psw != nullptr) || p.mean;
    const int seg_len = seg_len_from_env();
    const PlanLayout L = plan_exact(p.n_indices, p.num_tables, p.batch, p.dim, seg_len);
    if (!scratch || scratch_bytes < (long long)L.total) return PB200_EINVAL;
    if (p.n_indices > 0x7fffffffll) return PB200_EUNSUPPORTED;
    if (!plan_ready) {
        const int rc = build_sort_plan(p, idx_type, max_table_rows, scratch, L, s);
        if (rc != PB200_OK) return rc;
    }
    const SortedView sv = sorted_view(scratch, L, p.n_indices);
    BwdParams pr = p;
    pr.n_dev = sv.count;
    const int vec4 = p.dim >> 2;
    const bool adagrad = op.optimizer == PB200_OPT_ROWWISE_ADAGRAD;
    // measured (profiles/r01f): 3 resident CTAs/SM speed the Adagrad / fp16 variants up by 6-13 %
    static const int occ3 = [] {
        const char *e = getenv("PB200_EXACT_OCC3");
        return e ? atoi(e) : 1;
    }();
    // ONE reduce over the whole request: the keys are arena rows, sorted over all tables
    const long long n = sv.n, row0 = 0, n_seg = sv.n_seg;
    const unsigned *ks = sv.keys, *vs = sv.vals;
    float4 *partial = (float4 *)sv.extra;
    // behind the partial sums: the long-run work list (one entry per segment at most) + its counter
    unsigned *worklist = (unsigned *)(sv.extra + (size_t)n_seg * 2 * (size_t)p.dim * 4);
    unsigned *work_count = worklist + n_seg;
    if (n_seg > 1) PB200_CUDA_TRY(cudaMemsetAsync(work_count, 0, 4, s));
    long long g3 = 4ll * sm_count();
    if (g3 > n_seg) g3 = n_seg;
#define PB200_EXACT_LAUNCH(G_, C_)                                                                 \
    do {                                                                                           \
        const long long per_block = 8ll * (32 / G_);                                               \
        const long long g2 = (n_seg + per_block - 1) / per_block;                                  \
        if (g2 > 0x7fffffffll) return PB200_EUNSUPPORTED;                                          \
        if (!side && occ3 && G_ == 32 && C_ == 1 && adagrad)                                       \
            exact_reduce_kernel_occ3<WT, PB200_OPT_ROWWISE_ADAGRAD>                                \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, partial, seg_len);           \
        else if (!side && occ3 && G_ == 32 && C_ == 1)                                             \
            exact_reduce_kernel_occ3<WT, PB200_OPT_SGD>                                            \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, partial, seg_len);           \
        else if (side && adagrad)                                                                  \
            exact_reduce_kernel<WT, PB200_OPT_ROWWISE_ADAGRAD, G_, C_, true>                       \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, sv.goff_of, sv.w_of, partial, seg_len); \
        else if (side)                                                                             \
            exact_reduce_kernel<WT, PB200_OPT_SGD, G_, C_, true>                                   \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, sv.goff_of, sv.w_of, partial, seg_len); \
        else if (adagrad)                                                                          \
            exact_reduce_kernel<WT, PB200_OPT_ROWWISE_ADAGRAD, G_, C_, false>                      \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, nullptr, nullptr, partial, seg_len); \
        else                                                                                       \
            exact_reduce_kernel<WT, PB200_OPT_SGD, G_, C_, false>                                  \
                <<<(unsigned)g2, 256, 0, s>>>(pr, op, n, row0, ks, vs, nullptr, nullptr, partial, seg_len); \
        if (n_seg > 1) {                                                                           \
            exact_boundary_kernel<WT, G_, C_><<<(unsigned)g2, 256, 0, s>>>(                        \
                pr, op, n, row0, ks, partial, worklist, work_count, seg_len);                       \
            exact_long_run_kernel<WT, G_, C_><<<(unsigned)g3, 256, 0, s>>>(                        \
                pr, op, n, row0, ks, partial, worklist, work_count, seg_len);                       \
        }                                                                                          \
    } while (0)
    if (vec4 <= 4) PB200_EXACT_LAUNCH(4, 1);
    else if (vec4 <= 8) PB200_EXACT_LAUNCH(8, 1);
    else if (vec4 <= 16) PB200_EXACT_LAUNCH(16, 1);
    else if (vec4 <= 32) PB200_EXACT_LAUNCH(32, 1);
    else if (vec4 <= 64) PB200_EXACT_LAUNCH(32, 2);
    else PB200_EXACT_LAUNCH(32, 4);
#undef PB200_EXACT_LAUNCH
    count_launch(n_seg > 1 ? 3 : 1);
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" int64_t pb200_tbe_bwd_fused_scratch_bytes(int64_t n_indices, int32_t num_tables,
                                                     int64_t batch, int32_t dim) {
    if (n_indices <= 0 || num_tables < 1 || batch < 1 || dim < 1) return 0;
    return (int64_t)plan_exact(n_indices, num_tables, batch, dim, seg_len_from_env()).total;
}

extern "C" int pb200_tbe_bwd_fused(void *weights, int32_t weights_type, float *state,
                                   const int64_t *table_row_offsets, int32_t num_tables, int32_t dim,
                                   const void *indices, int64_t n_indices, const void *offsets,
                                   int64_t batch, int32_t idx_type, const float *psw,
                                   int32_t pool_mode, const float *grad_out, int64_t go_stride_t,
                                   int64_t go_stride_b, int32_t optimizer, float lr, float eps,
                                   int32_t stochastic_rounding, uint64_t sr_seed,
                                   int64_t max_table_rows, void *scratch, int64_t scratch_bytes,
                                   int32_t plan_ready, void *stream) {
    if (!weights || !table_row_offsets || !offsets || !grad_out || (!indices && n_indices > 0))
        return PB200_EINVAL;
    if (num_tables < 1 || dim < 1 || batch < 0 || n_indices < 0) return PB200_EINVAL;
    if (pool_mode != PB200_POOL_SUM && pool_mode != PB200_POOL_MEAN) return PB200_EINVAL;
    if (weights_type != PB200_W_F32 && weights_type != PB200_W_F16) return PB200_EINVAL;
    if (optimizer != PB200_OPT_SGD && optimizer != PB200_OPT_ROWWISE_ADAGRAD) return PB200_EINVAL;
    if (optimizer == PB200_OPT_ROWWISE_ADAGRAD && !state) return PB200_EINVAL;
    if (idx_type != PB200_IDX_I64 && idx_type != PB200_IDX_I32) return PB200_EINVAL;
    if (batch == 0 || n_indices == 0) return PB200_OK;
    // rows move as 16 B (fp32) / 8 B (fp16) vectors of 4 elements
    if (dim % 4 != 0 || dim > 512) return PB200_EUNSUPPORTED;
    const uintptr_t w_align = weights_type == PB200_W_F32 ? 15 : 7;
    if (((uintptr_t)weights & w_align) || ((uintptr_t)grad_out & 15)) return PB200_EALIGN;
    if (go_stride_t % 4 != 0 || go_stride_b % 4 != 0) return PB200_EALIGN;

    BwdParams p{};
    p.dst = nullptr;
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
    p.scale = 1.f;
    p.num_tables = num_tables;
    p.dim = dim;
    p.mean = pool_mode == PB200_POOL_MEAN;
    OptParams op{};
    op.weights = weights;
    op.state = state;
    op.lr = lr;
    op.eps = eps;
    op.optimizer = optimizer;
    op.stochastic = (weights_type == PB200_W_F16) && stochastic_rounding;
    op.sr_seed = sr_seed;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ready = plan_ready != 0;
    if (weights_type == PB200_W_F32)
        return bwd_exact<float>(p, op, idx_type, max_table_rows, scratch, scratch_bytes, ready, st);
    return bwd_exact<__half>(p, op, idx_type, max_table_rows, scratch, scratch_bytes, ready, st);
}
