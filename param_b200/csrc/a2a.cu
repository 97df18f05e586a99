// a2a.cu — single-node all-to-all as direct peer-HBM writes over NVLink 5 / NVSwitch (sm_100a).
//
// Replaces, for the DLRM path only:
//   dist.all_to_all_single(out, in, out_splits, in_splits)
//       train/comms/pt/pytorch_dist_backend.py:330-357 (all_to_all_single), :262-328 (all_to_allv)
//   All2Allv_Req/All2Allv_Wait + torch.cat either side of it
//       train/comms/pt/dlrm.py:86-218, :1253
//
// Every rank maps every peer's "window" (data) and "pad" (flags).  ONE kernel per rank and per
// collective does, for each destination j, a 2-D strided copy
//     peer_window[j] + dst_off[j] + r*dst_stride[j] + c   <-   src[j] + r*src_stride[j] + c
// with 16 B loads from local HBM and 16 B stores straight into the peer's HBM, so the per-rank
// output permute ([T,N,E] -> [lN, T_global*E] and its transpose) costs no extra pass: it is the
// address calculation of the push.  all_to_all_single is the 1-row case.
//
// Protocol per call (epoch e, monotonically increasing, kept in device memory so that a captured
// CUDA graph replays correctly):
//   1. ready : rank j tells every source r "my window may be overwritten for epoch e" and where
//              r's block goes (pad[r].ready_payload[j], pad[r].ready_epoch[j], st.release.sys).
//              This is what posting the receive buffer is for NCCL: the kernel runs after all of
//              j's earlier stream work, so earlier readers of the window have finished.
//   2. push  : every CTA, for each destination (rotated by rank and CTA so that all W-1 NVSwitch
//              ports are busy at once), waits for that destination's ready flag (ld.acquire.sys
//              on its OWN pad — local memory), then copies its slice.
//   3. done  : per destination, the last CTA to finish (device-scope counter) executes
//              fence.acq_rel.sys and st.release.sys pad[j].done_epoch[me] = e.
//   4. wait  : the last CTA of the grid spins until done_epoch[r] >= e for all sources r, then
//              publishes the new epoch.  Kernel completion == c10d Work.wait() semantics.
// No CTA ever waits on another CTA of the same grid, so the grid need not be co-resident.
#include <stdlib.h>

#include "a2a_common.cuh"

namespace pb200 {

constexpr int kA2AThreads = 512;
constexpr int kA2AUnroll = 8;

// 256-bit global accesses (Blackwell LDG.E.256 / STG.E.256): 32 B per lane, 1 KB per warp request
struct __align__(32) V32 {
    unsigned long long a, b, c, d;
};
__device__ __forceinline__ V32 ld_src_v32(const V32 *p) {
    V32 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer_v32(V32 *p, const V32 &v) {
    asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a),
                 "l"(v.b), "l"(v.c), "l"(v.d)
                 : "memory");
}
__device__ int g_a2a_vec32 = 0;   // set from PB200_A2A_VEC32 (experiment knob)

// contiguous copy of n 16 B units: kA2AUnroll independent loads in flight per thread
__device__ __forceinline__ void copy_linear16(int4 *__restrict__ dst, const int4 *__restrict__ src,
                                              long long n) {
    const long long step = kA2AThreads;
    if (g_a2a_vec32 && ((((unsigned long long)dst | (unsigned long long)src) & 31ull) == 0)) {
        const long long n2 = n >> 1;
        V32 *d2 = (V32 *)dst;
        const V32 *s2 = (const V32 *)src;
        long long u = threadIdx.x;
        constexpr int UN = kA2AUnroll / 2;
        for (; u + (UN - 1) * step < n2; u += UN * step) {
            V32 v[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) v[k] = ld_src_v32(s2 + u + k * step);
#pragma unroll
            for (int k = 0; k < UN; ++k) st_peer_v32(d2 + u + k * step, v[k]);
        }
        for (; u < n2; u += step) st_peer_v32(d2 + u, ld_src_v32(s2 + u));
        if ((n & 1) && threadIdx.x == 0) st_peer_v4(dst + n - 1, ld_src_v4(src + n - 1));
        return;
    }
    long long u = threadIdx.x;
    for (; u + (kA2AUnroll - 1) * step < n; u += kA2AUnroll * step) {
        int4 v[kA2AUnroll];
#pragma unroll
        for (int k = 0; k < kA2AUnroll; ++k) v[k] = ld_src_v4(src + u + k * step);
#pragma unroll
        for (int k = 0; k < kA2AUnroll; ++k) st_peer_v4(dst + u + k * step, v[k]);
    }
    for (; u < n; u += step) st_peer_v4(dst + u, ld_src_v4(src + u));
}

// contiguous copy of n units of 8 / 4 / 1 bytes (blocks whose addresses or sizes are not 16 B
// aligned: int64 index blocks with uneven splits are the common case), 4 loads in flight per thread
template <typename T>
__device__ __forceinline__ void copy_linear_small(T *dst, const T *src, long long n) {
    const long long step = kA2AThreads;
    long long u = threadIdx.x;
    for (; u + 3 * step < n; u += 4 * step) {
        const T v0 = src[u], v1 = src[u + step], v2 = src[u + 2 * step], v3 = src[u + 3 * step];
        dst[u] = v0;
        dst[u + step] = v1;
        dst[u + 2 * step] = v2;
        dst[u + 3 * step] = v3;
    }
    for (; u < n; u += step) dst[u] = src[u];
}

// Copy units [u0, u1) of one destination's 2-D block; a unit is UNIT bytes, run_units = units per
// row.  No division in the steady state: rows are walked explicitly.
template <int UNIT>
__device__ __forceinline__ void copy_units(unsigned char *dst, long long dst_stride,
                                           const unsigned char *src, long long src_stride,
                                           long long run_units, long long rows, long long u0,
                                           long long u1) {
    if (UNIT == 16) {
        if (rows == 1 || (src_stride == run_units * 16 && dst_stride == run_units * 16)) {
            // one contiguous range (all_to_all_single, or rows that happen to be adjacent)
            copy_linear16((int4 *)dst + u0, (const int4 *)src + u0, u1 - u0);
            return;
        }
        long long r = u0 / run_units;           // one division per (CTA, destination)
        long long c = u0 - r * run_units;
        long long left = u1 - u0;
        if (run_units >= kA2AThreads && (run_units % kA2AThreads) == 0 && c == 0 &&
            (left % run_units) == 0) {
            // long rows (e.g. 64 tables x 128 floats = 2048 units): thread t owns columns
            // t + k*512; (row, k) advance incrementally so kA2AUnroll loads are always in flight,
            // also across row boundaries
            const int per_row = (int)(run_units / kA2AThreads);
            const long long iters = (left / run_units) * per_row;
            const unsigned char *sp = src + r * src_stride + (long long)threadIdx.x * 16;
            unsigned char *dp = dst + r * dst_stride + (long long)threadIdx.x * 16;
            int k = 0;
            long long it = 0;
            for (; it + kA2AUnroll <= iters; it += kA2AUnroll) {
                int4 v[kA2AUnroll];
                unsigned char *d[kA2AUnroll];
#pragma unroll
                for (int j = 0; j < kA2AUnroll; ++j) {
                    v[j] = ld_src_v4((const int4 *)(sp + (long long)k * kA2AThreads * 16));
                    d[j] = dp + (long long)k * kA2AThreads * 16;
                    if (++k == per_row) {
                        k = 0;
                        sp += src_stride;
                        dp += dst_stride;
                    }
                }
#pragma unroll
                for (int j = 0; j < kA2AUnroll; ++j) st_peer_v4((int4 *)d[j], v[j]);
            }
            for (; it < iters; ++it) {
                st_peer_v4((int4 *)(dp + (long long)k * kA2AThreads * 16),
                           ld_src_v4((const int4 *)(sp + (long long)k * kA2AThreads * 16)));
                if (++k == per_row) {
                    k = 0;
                    sp += src_stride;
                    dp += dst_stride;
                }
            }
        } else if (run_units >= kA2AThreads) {
            // long rows, general: the whole CTA streams one row segment at a time
            while (left > 0) {
                const long long n = min(run_units - c, left);
                copy_linear16((int4 *)(dst + r * dst_stride) + c, (const int4 *)(src + r * src_stride) + c, n);
                left -= n;
                c = 0;
                ++r;
            }
        } else {
            // short rows: a thread owns a column, rows advance by blockDim / run_units per step
            // (only valid if run_units divides the block; otherwise fall back to per-unit division)
            const int ru = (int)run_units;
            if ((kA2AThreads % ru) == 0 && c == 0 && (left % ru) == 0) {
                const int col = threadIdx.x % ru;
                const long long rstep = kA2AThreads / ru;
                const long long r_end = r + left / ru;     // u1 - u0 is a multiple of run_units here
                long long rr = r + threadIdx.x / ru;
                for (; rr + (kA2AUnroll - 1) * rstep < r_end; rr += kA2AUnroll * rstep) {
                    int4 v[kA2AUnroll];
#pragma unroll
                    for (int k = 0; k < kA2AUnroll; ++k)
                        v[k] = ld_src_v4((const int4 *)(src + (rr + k * rstep) * src_stride) + col);
#pragma unroll
                    for (int k = 0; k < kA2AUnroll; ++k)
                        st_peer_v4((int4 *)(dst + (rr + k * rstep) * dst_stride) + col, v[k]);
                }
                for (; rr < r_end; rr += rstep)
                    st_peer_v4((int4 *)(dst + rr * dst_stride) + col,
                               ld_src_v4((const int4 *)(src + rr * src_stride) + col));
            } else {
                for (long long u = u0 + threadIdx.x; u < u1; u += kA2AThreads) {
                    const long long rr = u / run_units, cc = u - rr * run_units;
                    st_peer_v4((int4 *)(dst + rr * dst_stride) + cc,
                               ld_src_v4((const int4 *)(src + rr * src_stride) + cc));
                }
            }
        }
    } else {
        if (rows == 1) {   // all_to_all_single: one contiguous range, no per-unit division
            if (UNIT == 8)
                copy_linear_small((long long *)dst + u0, (const long long *)src + u0, u1 - u0);
            else if (UNIT == 4)
                copy_linear_small((int *)dst + u0, (const int *)src + u0, u1 - u0);
            else
                copy_linear_small(dst + u0, src + u0, u1 - u0);
            return;
        }
        for (long long u = u0 + threadIdx.x; u < u1; u += kA2AThreads) {
            const long long r = u / run_units;
            const long long c = u - r * run_units;
            if (UNIT == 8)
                *(long long *)(dst + r * dst_stride + c * 8) =
                    ld_stream_i64((const long long *)(src + r * src_stride + c * 8));
            else if (UNIT == 4)
                *(int *)(dst + r * dst_stride + c * 4) = *(const int *)(src + r * src_stride + c * 4);
            else
                dst[r * dst_stride + c] = src[r * src_stride + c];
        }
    }
}

__global__ void __launch_bounds__(kA2AThreads) a2a_push_kernel(const A2AArgs a) {
    __shared__ unsigned long long s_epoch;
    __shared__ long long s_dst_off[PB200_A2A_MAX_RANKS];
    const int W = a.world;
    const int me = a.rank;
    SignalPad *my_pad = a.peer_pad[me];

    if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long *)a.epoch + 1ull;
    __syncthreads();
    const unsigned long long e = s_epoch;

    // 1. ready: CTA 0 posts my receive offsets to every source; every CTA then collects the
    //    destinations' offsets — one thread per destination, all flags polled concurrently
    //    (the flags live in this rank's own pad: local memory reads).
    if (threadIdx.x < W) {
        const int r = threadIdx.x;
        if (r == me) {
            s_dst_off[r] = a.recv_off[me];
        } else {
            if (blockIdx.x == 0) {
                SignalPad *pp = a.peer_pad[r];
                st_relaxed_sys(&pp->ready_payload[me], (unsigned long long)a.recv_off[r]);
                st_release_sys(&pp->ready_epoch[me], e);
            }
            const bool ok = wait_flag_ge(&my_pad->ready_epoch[r], e, a.spin_cycles, a.error);
            // on timeout: never write to an unready peer
            s_dst_off[r] = ok ? (long long)ld_relaxed_sys(&my_pad->ready_payload[r]) : -1;
        }
    }
    __syncthreads();

    // 2. push: all destinations back to back, no barrier in between.  Order: self first, then
    //    destinations staggered by rank and by CTA so every NVSwitch port is busy.
    for (int k = 0; k < W; ++k) {
        const int j = (k == 0) ? me : (me + 1 + ((k - 1) + blockIdx.x) % (W - 1)) % W;
        PeerCopy pc = a.copy[j];
        if (a.dev_counts) {
            // block sizes live in device memory (computed by an earlier kernel on this stream)
            long long before = 0;
            for (int k = 0; k < j; ++k) before += a.dev_counts[k];
            long long bytes = a.dev_counts[j] * a.dev_elem_bytes;
            if (bytes > a.dev_slot_bytes) {      // never write past the receiver's slot
                bytes = a.dev_slot_bytes;
                if (threadIdx.x == 0) atomicExch(a.error, 2u);
            }
            pc.src = a.dev_src + before * a.dev_elem_bytes;
            pc.src_stride = pc.dst_stride = 0;
            pc.run_bytes = bytes;
            pc.rows = bytes > 0 ? 1 : 0;
        }
        const long long dst_off = s_dst_off[j];
        if (pc.rows > 0 && pc.run_bytes > 0 && dst_off >= 0) {
            unsigned char *dst = a.peer_data[j] + dst_off;
            // copy unit from the alignment of everything that moves for THIS destination (the
            // destination offset is chosen by the receiver, so it is only known here)
            unsigned long long bits = (unsigned long long)pc.src | (unsigned long long)pc.run_bytes |
                                      (unsigned long long)dst;
            if (pc.rows > 1)
                bits |= (unsigned long long)pc.src_stride | (unsigned long long)pc.dst_stride;
            // (int64 payloads with uneven splits — the DLRM index exchange — are 8 B aligned)
            const int ul = (bits & 15ull) == 0 ? 4 : ((bits & 7ull) == 0 ? 3 : ((bits & 3ull) == 0 ? 2 : 0));
            const long long run_units = pc.run_bytes >> ul;
            const long long total_units = run_units * pc.rows;
            // contiguous slice per CTA (long store runs); whole rows when the block is 2-D
            long long u0, u1;
            if (pc.rows > 1 && pc.rows >= (long long)gridDim.x) {
                const long long rper = (pc.rows + gridDim.x - 1) / gridDim.x;
                u0 = min(rper * blockIdx.x, pc.rows) * run_units;
                u1 = min(rper * (blockIdx.x + 1), pc.rows) * run_units;
            } else {
                const long long per = (total_units + gridDim.x - 1) / gridDim.x;
                u0 = min(per * blockIdx.x, total_units);
                u1 = min(u0 + per, total_units);
            }
            if (u0 < u1) {
                if (ul == 4)
                    copy_units<16>(dst, pc.dst_stride, pc.src, pc.src_stride, run_units, pc.rows, u0, u1);
                else if (ul == 3)
                    copy_units<8>(dst, pc.dst_stride, pc.src, pc.src_stride, run_units, pc.rows, u0, u1);
                else if (ul == 2)
                    copy_units<4>(dst, pc.dst_stride, pc.src, pc.src_stride, run_units, pc.rows, u0, u1);
                else
                    copy_units<1>(dst, pc.dst_stride, pc.src, pc.src_stride, run_units, pc.rows, u0, u1);
            }
        }
    }

    // 3. done: one barrier, then W threads signal their destination in parallel — the last CTA to
    //    finish (device-scope counter) releases the flag; fence.acq_rel.sys orders this CTA's
    //    peer stores (all threads, via the barrier) before the counter / flag.
    __syncthreads();
    if (threadIdx.x < W && (int)threadIdx.x != me) {
        const int j = threadIdx.x;
        __threadfence_system();
        const unsigned prev = atomicAdd(&a.peer_cnt[j], 1u);
        if (prev == gridDim.x - 1) {
            a.peer_cnt[j] = 0;
            __threadfence_system();
            st_release_sys(&a.peer_pad[j]->done_epoch[me], e);
        }
    }

    // 4. wait: the last CTA of the grid waits for all sources (in parallel), then publishes the epoch
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(a.grid_cnt, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < W && (int)threadIdx.x != me)
            wait_flag_ge(&my_pad->done_epoch[threadIdx.x], e, a.spin_cycles, a.error);
        __syncthreads();
        if (threadIdx.x == 0) {
            *a.grid_cnt = 0;
            *(volatile unsigned long long *)a.epoch = e;
            __threadfence();
        }
    }
}

}  // namespace pb200

using namespace pb200;


extern "C" int pb200_a2a_comm_create(pb200_a2a_comm **comm, int32_t rank, int32_t world,
                                     void *const *peer_data, void *const *peer_signal,
                                     int64_t window_bytes) {
    if (!comm || !peer_data || !peer_signal) return PB200_EINVAL;
    if (world < 1 || world > PB200_A2A_MAX_RANKS || rank < 0 || rank >= world || window_bytes < 0)
        return PB200_EINVAL;
    pb200_a2a_comm *c = (pb200_a2a_comm *)calloc(1, sizeof(pb200_a2a_comm));
    if (!c) return PB200_EINVAL;
    c->rank = rank;
    c->world = world;
    c->window_bytes = window_bytes;
    for (int r = 0; r < world; ++r) {
        if (!peer_data[r] || !peer_signal[r]) {
            free(c);
            return PB200_EINVAL;
        }
        c->peer_data[r] = (unsigned char *)peer_data[r];
        c->peer_pad[r] = (SignalPad *)peer_signal[r];
    }
    unsigned char *blk = nullptr;
    cudaError_t e = cudaMalloc(&blk, 256);
    if (e != cudaSuccess) {
        free(c);
        return (int)e;
    }
    e = cudaMemset(blk, 0, 256);
    if (e != cudaSuccess) {
        cudaFree(blk);
        free(c);
        return (int)e;
    }
    c->d_epoch = (unsigned long long *)blk;
    c->d_grid_cnt = (unsigned *)(blk + 8);
    c->d_peer_cnt = (unsigned *)(blk + 64);
    c->d_error = (unsigned *)(blk + 16);
    c->spin_cycles = 20ll * 1000 * 1000 * 1000;  // ~10 s at 2 GHz
    const char *env = getenv("PB200_A2A_CTAS");
    c->max_ctas = env ? atoi(env) : 0;
    const char *v32 = getenv("PB200_A2A_VEC32");
    const int use32 = v32 ? atoi(v32) : 0;
    cudaMemcpyToSymbol(g_a2a_vec32, &use32, sizeof(int));
    *comm = c;
    return PB200_OK;
}

extern "C" int pb200_a2a_comm_config(pb200_a2a_comm *comm, int32_t max_ctas, double spin_timeout_s) {
    if (!comm) return PB200_EINVAL;
    if (max_ctas >= 0) comm->max_ctas = max_ctas;
    if (spin_timeout_s > 0) comm->spin_cycles = (long long)(spin_timeout_s * 2.0e9);
    return PB200_OK;
}

extern "C" int pb200_a2a_comm_error(pb200_a2a_comm *comm, int32_t *error_out) {
    if (!comm || !error_out) return PB200_EINVAL;
    unsigned v = 0;
    PB200_CUDA_TRY(cudaMemcpy(&v, comm->d_error, sizeof(v), cudaMemcpyDeviceToHost));
    *error_out = (int32_t)v;
    return PB200_OK;
}

extern "C" int pb200_a2a_comm_destroy(pb200_a2a_comm *comm) {
    if (!comm) return PB200_OK;
    if (comm->d_epoch) cudaFree(comm->d_epoch);
    free(comm);
    return PB200_OK;
}

int pb200::a2a_launch_args(pb200_a2a_comm *c, A2AArgs &a, long long max_peer_bytes, cudaStream_t st, int grid_cap) {
    for (int r = 0; r < c->world; ++r) {
        a.peer_data[r] = c->peer_data[r];
        a.peer_pad[r] = c->peer_pad[r];
    }
    a.epoch = c->d_epoch;
    a.peer_cnt = c->d_peer_cnt;
    a.grid_cnt = c->d_grid_cnt;
    a.error = c->d_error;
    a.spin_cycles = c->spin_cycles;
    a.rank = c->rank;
    a.world = c->world;
    // grid: one CTA per 16 KB of the largest per-peer block, at least 1, at most the SM count
    long long grid = (max_peer_bytes + (16ll << 10) - 1) / (16ll << 10);
    int cap = c->max_ctas > 0 ? c->max_ctas : sm_count();
    if (grid_cap > 0 && grid_cap < cap) cap = grid_cap;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    a2a_push_kernel<<<(unsigned)grid, kA2AThreads, 0, st>>>(a);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

extern "C" int pb200_a2a_single(pb200_a2a_comm *c, const void *in, int64_t total_in_bytes,
                                const int64_t *in_split_bytes, const int64_t *out_split_bytes,
                                int64_t out_window_off, void *out, void *stream) {
    if (!c || (!in && total_in_bytes > 0) || total_in_bytes < 0 || out_window_off < 0)
        return PB200_EINVAL;
    if ((in_split_bytes == nullptr) != (out_split_bytes == nullptr)) return PB200_EINVAL;
    const int W = c->world;
    if (!in_split_bytes && total_in_bytes % W != 0) return PB200_EINVAL;
    A2AArgs a{};
    long long in_off = 0, out_off = out_window_off, max_peer = 0, total_out = 0;
    for (int r = 0; r < W; ++r) {
        const long long sb = in_split_bytes ? in_split_bytes[r] : total_in_bytes / W;
        const long long rb = out_split_bytes ? out_split_bytes[r] : total_in_bytes / W;
        if (sb < 0 || rb < 0) return PB200_EINVAL;
        a.copy[r].src = (const unsigned char *)in + in_off;
        a.copy[r].src_stride = 0;
        a.copy[r].dst_stride = 0;
        a.copy[r].run_bytes = sb;
        a.copy[r].rows = sb > 0 ? 1 : 0;
        a.recv_off[r] = out_off;
        in_off += sb;
        out_off += rb;
        total_out += rb;
        if (sb > max_peer) max_peer = sb;
    }
    if (in_off != total_in_bytes && in_split_bytes) return PB200_EINVAL;
    if (out_off > c->window_bytes) return PB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = a2a_launch_args(c, a, max_peer, st);
    if (rc != PB200_OK) return rc;
    if (out && total_out > 0)
        PB200_CUDA_TRY(cudaMemcpyAsync(out, c->peer_data[c->rank] + out_window_off, (size_t)total_out,
                                       cudaMemcpyDeviceToDevice, st));
    return PB200_OK;
}

// List form (dist.all_to_all(output_tensor_list, input_tensor_list)): W separate input blocks, W separate
// landing places.  Same kernel, same protocol: the per-destination source pointer and the per-source
// receive offset were already independent — nothing is packed before or unpacked after.
extern "C" int pb200_a2a_list(pb200_a2a_comm *c, const void *const *in_ptrs, const int64_t *in_bytes,
                              const int64_t *out_window_offs, const int64_t *out_bytes,
                              void *const *out_copy, void *stream) {
    if (!c || !in_ptrs || !in_bytes || !out_window_offs || !out_bytes) return PB200_EINVAL;
    const int W = c->world;
    A2AArgs a{};
    long long max_peer = 0;
    for (int r = 0; r < W; ++r) {
        if (in_bytes[r] < 0 || out_bytes[r] < 0 || out_window_offs[r] < 0) return PB200_EINVAL;
        if (in_bytes[r] > 0 && !in_ptrs[r]) return PB200_EINVAL;
        if (out_window_offs[r] + out_bytes[r] > c->window_bytes) return PB200_EINVAL;
        a.copy[r].src = (const unsigned char *)in_ptrs[r];
        a.copy[r].src_stride = 0;
        a.copy[r].dst_stride = 0;
        a.copy[r].run_bytes = in_bytes[r];
        a.copy[r].rows = in_bytes[r] > 0 ? 1 : 0;
        a.recv_off[r] = out_window_offs[r];
        if (in_bytes[r] > max_peer) max_peer = in_bytes[r];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = a2a_launch_args(c, a, max_peer, st);
    if (rc != PB200_OK) return rc;
    if (out_copy)
        for (int r = 0; r < W; ++r)
            if (out_copy[r] && out_bytes[r] > 0)
                PB200_CUDA_TRY(cudaMemcpyAsync(out_copy[r], c->peer_data[c->rank] + out_window_offs[r],
                                               (size_t)out_bytes[r], cudaMemcpyDeviceToDevice, st));
    return PB200_OK;
}

extern "C" int pb200_a2a_pooled_fwd(pb200_a2a_comm *c, const float *in, int64_t in_stride_t,
                                    int64_t in_stride_n, int32_t emb_dim,
                                    const int64_t *batch_split, const int64_t *tables_split,
                                    int64_t out_window_off, void *stream) {
    if (!c || !in || !batch_split || !tables_split || emb_dim < 1 || out_window_off < 0)
        return PB200_EINVAL;
    const int W = c->world, me = c->rank;
    long long T_global = 0, table_base[PB200_A2A_MAX_RANKS], n_base[PB200_A2A_MAX_RANKS], N = 0;
    for (int r = 0; r < W; ++r) {
        if (batch_split[r] < 0 || tables_split[r] < 0) return PB200_EINVAL;
        table_base[r] = T_global;
        T_global += tables_split[r];
        n_base[r] = N;
        N += batch_split[r];
    }
    const long long T_local = tables_split[me];
    const long long E = emb_dim;
    // the pushed run is one (row, table) cell of E floats unless the local layout already has the
    // tables adjacent inside a row (in_stride_t == E), in which case it is the whole T_local*E run
    // NB: the choice must be the same on every rank (it fixes the number of collectives), so it
    // depends on the layout alone, never on this rank's table count.
    const bool tables_adjacent = (in_stride_t == E);
    if (!tables_adjacent) {
        // [T, N, E]-style input (dlrm.py torch.stack layout): one 2-D push per local table.
        // Every rank must run the same number of rounds (epochs advance in lock step), so the
        // round count is max_r tables_split[r]; ranks with fewer tables send empty rounds.
        // (The TBE forward can write [N, T*E] directly, which is the single-launch fast path.)
        cudaStream_t st = (cudaStream_t)stream;
        long long T_max = 0;
        for (int r = 0; r < W; ++r) T_max = tables_split[r] > T_max ? tables_split[r] : T_max;
        if (out_window_off + batch_split[me] * T_global * E * 4 > c->window_bytes)
            return PB200_EINVAL;
        for (long long t = 0; t < T_max; ++t) {
            A2AArgs a{};
            long long max_peer = 0;
            for (int j = 0; j < W; ++j) {
                const bool have = t < T_local;
                a.copy[j].src =
                    (const unsigned char *)(in + (have ? t : 0) * in_stride_t + n_base[j] * in_stride_n);
                a.copy[j].src_stride = in_stride_n * 4;
                a.copy[j].dst_stride = T_global * E * 4;
                a.copy[j].run_bytes = E * 4;
                a.copy[j].rows = have ? batch_split[j] : 0;
                // where source j's t-th table goes in MY window
                a.recv_off[j] = out_window_off + (table_base[j] + t) * E * 4;
                const long long bytes = E * 4 * a.copy[j].rows;
                if (bytes > max_peer) max_peer = bytes;
            }
            int rc = a2a_launch_args(c, a, max_peer, st);
            if (rc != PB200_OK) return rc;
        }
        return PB200_OK;
    }
    A2AArgs a{};
    long long max_peer = 0;
    for (int j = 0; j < W; ++j) {
        a.copy[j].src = (const unsigned char *)(in + n_base[j] * in_stride_n);
        a.copy[j].src_stride = in_stride_n * 4;
        a.copy[j].dst_stride = T_global * E * 4;
        a.copy[j].run_bytes = T_local * E * 4;
        a.copy[j].rows = batch_split[j];
        a.recv_off[j] = out_window_off + table_base[j] * E * 4;  // source j's columns in my rows
        const long long bytes = a.copy[j].run_bytes * a.copy[j].rows;
        if (bytes > max_peer) max_peer = bytes;
    }
    if (out_window_off + batch_split[me] * T_global * E * 4 > c->window_bytes) return PB200_EINVAL;
    return a2a_launch_args(c, a, max_peer, (cudaStream_t)stream);
}

// tables [lo, hi) of a rank that owns n tables, for part `part` of `parts` (contiguous, remainder to the low parts)
static void part_range(long long n, int part, int parts, long long &lo, long long &hi) {
    const long long k = n / parts, m = n % parts;
    lo = part * k + (part < m ? part : m);
    hi = lo + k + (part < m ? 1 : 0);
}

extern "C" int pb200_a2a_pooled_bwd(pb200_a2a_comm *c, const float *grad, int32_t emb_dim,
                                    const int64_t *batch_split, const int64_t *tables_split,
                                    int64_t out_window_off, void *stream) {
    return pb200_a2a_pooled_bwd_part(c, grad, emb_dim, batch_split, tables_split, out_window_off, 0, 1, stream);
}

extern "C" int pb200_a2a_pooled_bwd_part(pb200_a2a_comm *c, const float *grad, int32_t emb_dim,
                                         const int64_t *batch_split, const int64_t *tables_split,
                                         int64_t out_window_off, int32_t part, int32_t parts, void *stream) {
    if (!c || !grad || !batch_split || !tables_split || emb_dim < 1 || out_window_off < 0)
        return PB200_EINVAL;
    if (parts < 1 || part < 0 || part >= parts) return PB200_EINVAL;
    const int W = c->world, me = c->rank;
    long long T_global = 0, table_base[PB200_A2A_MAX_RANKS], n_base[PB200_A2A_MAX_RANKS], N = 0;
    for (int r = 0; r < W; ++r) {
        if (batch_split[r] < 0 || tables_split[r] < 0) return PB200_EINVAL;
        table_base[r] = T_global;
        T_global += tables_split[r];
        n_base[r] = N;
        N += batch_split[r];
    }
    const long long E = emb_dim;
    const long long T_local = tables_split[me];
    // my grad [lN_me, T_global*E]: owner j gets columns [table_base[j]*E, +T_j*E) of every row,
    // landing in its window as rows n_base[me] .. of a [N, T_j*E] tensor
    // with parts > 1 only the columns of every owner's tables [lo, hi) of this part move: the owner can start
    // reducing part g while part g + 1 is still on the wire (same layout in the window, same epoch protocol)
    long long my_lo, my_hi;
    part_range(T_local, part, parts, my_lo, my_hi);
    A2AArgs a{};
    long long max_peer = 0;
    for (int j = 0; j < W; ++j) {
        long long lo, hi;
        part_range(tables_split[j], part, parts, lo, hi);
        a.copy[j].src = (const unsigned char *)(grad + (table_base[j] + lo) * E);
        a.copy[j].src_stride = T_global * E * 4;
        a.copy[j].dst_stride = tables_split[j] * E * 4;
        a.copy[j].run_bytes = (hi - lo) * E * 4;
        a.copy[j].rows = (hi > lo) ? batch_split[me] : 0;
        // source j's rows start at n_base[j] in MY [N, T_local*E] window tensor, my part's columns at my_lo
        a.recv_off[j] = out_window_off + n_base[j] * T_local * E * 4 + my_lo * E * 4;
        const long long bytes = a.copy[j].run_bytes * a.copy[j].rows;
        if (bytes > max_peer) max_peer = bytes;
    }
    if (out_window_off + N * T_local * E * 4 > c->window_bytes) return PB200_EINVAL;
    // a partial exchange exists to run UNDER the reduce of the previous part: a full grid of 512-thread CTAs
    // (28 K registers each) would leave room for one reduce CTA per SM instead of three.  The exchange is
    // NVLink-bound — 32 CTAs reach 90 % of the full grid's bandwidth (profiles/r01_a2a_cta_sweep_n8.log) — so the
    // partial pushes are capped (PB200_A2A_PART_CTAS, read per call; 0 = no cap).
    int part_cap = 0;
    if (parts > 1) {
        const char *e = getenv("PB200_A2A_PART_CTAS");
        part_cap = e ? atoi(e) : 32;
    }
    return a2a_launch_args(c, a, max_peer, (cudaStream_t)stream, part_cap);
}
