// a2a_common.cuh — signal pad, system-scope flag helpers and the communicator record shared by the
// push kernel (a2a.cu) and the fused lookup + all-to-all kernel (fused_fwd_a2a.cu).
#pragma once
#include "common.cuh"

struct pb200_a2a_comm;

namespace pb200 {

struct SignalPad {
    unsigned long long ready_epoch[PB200_A2A_MAX_RANKS];
    unsigned long long ready_payload[PB200_A2A_MAX_RANKS];
    unsigned long long done_epoch[PB200_A2A_MAX_RANKS];
};
static_assert(sizeof(SignalPad) <= PB200_A2A_SIGNAL_BYTES, "signal pad too small");

struct PeerCopy {
    const unsigned char *src;   // local
    long long src_stride;       // bytes between rows
    long long dst_stride;       // bytes between rows in the destination window
    long long run_bytes;        // contiguous bytes per row
    long long rows;
};

struct A2AArgs {
    unsigned char *peer_data[PB200_A2A_MAX_RANKS];
    SignalPad *peer_pad[PB200_A2A_MAX_RANKS];
    PeerCopy copy[PB200_A2A_MAX_RANKS];            // what I send to each destination
    long long recv_off[PB200_A2A_MAX_RANKS];       // where source r must write inside MY window
    unsigned long long *epoch;                      // device: last completed epoch
    unsigned *peer_cnt;                             // device [W]: CTAs finished per destination
    unsigned *grid_cnt;                             // device: CTAs finished overall
    unsigned *error;                                // device: set to 1 when a spin wait timed out
    long long spin_cycles;                          // give up a flag wait after this many clocks
    int rank;
    int world;
    // device-resident send counts (sparse input redistribution without a host round trip): when
    // dev_counts != nullptr the block for destination j is dev_counts[j] elements starting at
    // dev_src + (sum_{k<j} dev_counts[k]) * dev_elem_bytes, and copy[] is ignored.  The receiver
    // gives every source a fixed slot of dev_slot_bytes (recv_off[r] = base + r * slot).
    const long long *dev_counts;
    const unsigned char *dev_src;
    long long dev_elem_bytes;
    long long dev_slot_bytes;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int4 ld_src_v4(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer_v4(int4 *p, const int4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// Bounded spin on a flag in local memory: a peer that never arrives (crashed rank, mismatched
// call sequence) must not hang the GPU — after spin_cycles the kernel records an error and moves on.
__device__ __forceinline__ bool wait_flag_ge(const unsigned long long *flag, unsigned long long e,
                                             long long budget, unsigned *err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < e) {
        if (clock64() - t0 > budget) {
            atomicExch(err, 1u);
            return false;
        }
    }
    return true;
}

}  // namespace pb200

namespace pb200 {
// fills the communicator fields of `a` and launches the push kernel (a2a.cu); shared with the
// sparse-input redistribution entry (sparse_dist.cu)
// grid_cap > 0: at most that many CTAs for this launch (a push that runs under a compute kernel)
int a2a_launch_args(pb200_a2a_comm *c, A2AArgs &a, long long max_peer_bytes, cudaStream_t st, int grid_cap = 0);
}  // namespace pb200

struct pb200_a2a_comm {
    int rank;
    int world;
    long long window_bytes;
    unsigned char *peer_data[PB200_A2A_MAX_RANKS];
    pb200::SignalPad *peer_pad[PB200_A2A_MAX_RANKS];
    unsigned long long *d_epoch;
    unsigned *d_peer_cnt;
    unsigned *d_grid_cnt;
    unsigned *d_error;
    long long spin_cycles;
    int max_ctas;
};
