// sparse_dist.cu — device-side regroup of the redistributed sparse inputs (integer, bit-exact).
//
// Replaces paramDLRM_Net.splitPerTable + lengthsToOffsets, train/comms/pt/dlrm.py:430-504,
// :245-251: an O(W*T_l) python loop of slice + torch.cat (quadratic copying) and a .to("cpu")
// of the segment sums.  Here: 4 launches, no host round trip.
//
// in : lengths_in [W][T_l][b]   (rank-major, then local table, then sample; dlrm.py:449-451)
//      indices_in concatenated in the same (rank, table, sample) order
// out: lengths_out [T_l][W*b]   (table f = concat over ranks r of lengths[r][f][:])
//      offsets_out [T_l*W*b + 1] exclusive cumsum over the table-major concatenation
//                  (TBE layout; table f's nn.EmbeddingBag offsets are
//                   offsets_out[f*W*b : (f+1)*W*b] - offsets_out[f*W*b])
//      indices_out table-major permutation of indices_in
//
// pb200_sparse_data_dist is the whole SparseDataDist step (dlrm.py:744-855) as one asynchronous
// call: lengths all-to-all -> per-destination index counts (device) -> indices all-to-all whose
// block sizes are read from device memory by the push kernel, each source writing into a fixed
// slot of the receiver's window -> regroup.  No .item(), no D2H: the reference syncs the host
// twice here (dlrm.py:801-818).
#include <cub/device/device_scan.cuh>

#include "a2a_common.cuh"

namespace pb200 {

// flat over all W*T*b lengths: permuted copy + per-segment sums (warp-aggregated atomics).
// seg_sum must be zeroed beforehand.
__global__ void __launch_bounds__(256) seg_sum_permute_kernel(const long long *__restrict__ lengths_in,
                                                              int W, int T, long long b,
                                                              long long *__restrict__ lengths_out,
                                                              unsigned long long *__restrict__ seg_sum) {
    const long long n = (long long)W * T * b;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; i0 < n; i0 += stride) {
        const long long i = i0 + (threadIdx.x & 31);
        long long v = 0, seg = -1;
        if (i < n) {
            seg = i / b;
            const long long s = i - seg * b;
            const int r = (int)(seg / T), t = (int)(seg - (long long)r * T);
            v = ld_stream_i64(lengths_in + i);
            lengths_out[((long long)t * W + r) * b + s] = v;
        }
        const long long seg0 = __shfl_sync(0xffffffffu, seg, 0);
        const bool uniform = __all_sync(0xffffffffu, seg == seg0);
        if (uniform) {
            long long sum = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if ((threadIdx.x & 31) == 0 && seg0 >= 0) atomicAdd(seg_sum + seg0, (unsigned long long)sum);
        } else if (seg >= 0 && v != 0) {
            atomicAdd(seg_sum + seg, (unsigned long long)v);
        }
    }
}

// single CTA: exclusive prefix of seg_sum in (r,t) order -> in_start / in_end, in (t,r) order ->
// out_start.  slot_elems > 0: source r's indices start at r*slot_elems (fixed slots) instead of
// right behind source r-1's.
__global__ void __launch_bounds__(64) seg_starts_kernel(const long long *seg_sum, int W, int T,
                                                        long long slot_elems, long long *in_start,
                                                        long long *in_end, long long *out_start) {
    // W*T is at most a few thousand: serial scan by one thread per ordering is ~microseconds
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int r = 0; r < W; ++r) {
            if (slot_elems > 0) acc = (long long)r * slot_elems;
            for (int t = 0; t < T; ++t) {
                const int s = r * T + t;
                in_start[s] = acc;
                acc += seg_sum[s];
                in_end[s] = acc;
            }
        }
    } else if (threadIdx.x == 32) {
        long long acc = 0;
        for (int t = 0; t < T; ++t)
            for (int r = 0; r < W; ++r) {
                out_start[r * T + t] = acc;
                acc += seg_sum[r * T + t];
            }
    }
}

// flat over the input index space [0, n_indices) (packed: all received indices; slots: W*slot_elems,
// with unused tails): element i belongs to the segment s with in_start[s] <= i < in_end[s] (binary
// search in a shared-memory copy of in_start, re-done only when the running segment is left), and
// moves to out_start[s] + (i - in_start[s]); positions in no segment (slot tails) are skipped.
constexpr int kSegCopyChunk = 256 * 8;
__global__ void __launch_bounds__(256) seg_copy_kernel(const long long *__restrict__ indices_in,
                                                       long long n_indices,
                                                       const long long *__restrict__ in_start,
                                                       const long long *__restrict__ in_end,
                                                       const long long *__restrict__ out_start,
                                                       int n_seg, long long *__restrict__ indices_out,
                                                       long long out_cap) {
    extern __shared__ long long s_start[];   // [n_seg] in_start, [n_seg] in_end, [n_seg] out_start
    long long *s_end = s_start + n_seg;
    long long *s_out = s_end + n_seg;
    for (int k = threadIdx.x; k < n_seg; k += blockDim.x) {
        s_start[k] = in_start[k];
        s_end[k] = in_end[k];
        s_out[k] = out_start[k];
    }
    __syncthreads();
    const long long n_chunks = (n_indices + kSegCopyChunk - 1) / kSegCopyChunk;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long base = chunk * kSegCopyChunk;
        long long v[8];
        long long idx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            idx[k] = base + threadIdx.x + k * 256;
            v[k] = idx[k] < n_indices ? ld_stream_i64(indices_in + idx[k]) : 0;
        }
        int seg = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long i = idx[k];
            if (i >= n_indices) break;
            if (seg < 0 || i >= s_end[seg]) {
                int lo = 0, hi = n_seg - 1;   // last s with in_start[s] <= i (the non-empty one of equals)
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_start[mid] <= i)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                seg = lo;
            }
            if (i >= s_start[seg] && i < s_end[seg]) {
                // out_cap only bites when the lengths promise more indices than arrived (a
                // truncated slot, already flagged): never write outside the output buffer
                const long long o = s_out[seg] + (i - s_start[seg]);
                if (o < out_cap) indices_out[o] = v[k];
            }
        }
    }
}

__global__ void write_total_kernel(const long long *offsets_out, const long long *lengths_out,
                                   long long n, long long *dst) {
    // offsets_out[n] = offsets_out[n-1] + lengths_out[n-1]
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = n > 0 ? offsets_out[n - 1] + lengths_out[n - 1] : 0;
}

// counts[j] = sum of my lengths over the tables rank j owns = how many indices I send to j
// (dlrm.py:801-809).  kCountSlices CTAs per destination, each summing a slice and adding it with one
// integer atomic (one CTA per destination took 0.2 ms for 64 tables x 8192 bags at W = 2: a single
// CTA streams only ~20 GB/s).  counts must be zeroed beforehand.
struct TableBases {
    long long base[PB200_A2A_MAX_RANKS + 1];
};
constexpr int kCountSlices = 64;
__global__ void __launch_bounds__(256) dest_counts_kernel(const long long *__restrict__ lengths,
                                                          const TableBases tb, long long b,
                                                          unsigned long long *__restrict__ counts) {
    __shared__ long long s_part[8];
    const int j = blockIdx.y;
    const long long lo = tb.base[j] * b, hi = tb.base[j + 1] * b;
    long long acc = 0;
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
         i += (long long)gridDim.x * blockDim.x)
        acc += ld_stream_i64(lengths + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        if (t != 0) atomicAdd(counts + j, (unsigned long long)t);
    }
}

static size_t scan_tmp_bytes(long long n) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const long long *)nullptr, (long long *)nullptr,
                                  (int)(n > 0x7fffffffll ? 0x7fffffff : n));
    return (tmp + 255) & ~(size_t)255;
}

// scratch layout: [cub scan tmp | seg_sum | in_start | in_end | out_start | counts(256 B)]
static size_t seg_array_bytes(int world, int tables_local) {
    return ((((size_t)world * tables_local + 1) * 8) + 255) & ~(size_t)255;
}

static int regroup_impl(const long long *lengths_in, const long long *indices_in, long long n_space,
                        long long slot_elems, int world, int tables_local, long long local_batch,
                        long long *lengths_out, long long *offsets_out, long long *indices_out,
                        unsigned char *base, cudaStream_t st) {
    const long long n = (long long)world * tables_local * local_batch;
    const int n_seg = world * tables_local;
    const size_t tmp_bytes = scan_tmp_bytes(n);
    const size_t seg = seg_array_bytes(world, tables_local);
    long long *seg_sum = (long long *)(base + tmp_bytes);
    long long *in_start = (long long *)(base + tmp_bytes + seg);
    long long *in_end = (long long *)(base + tmp_bytes + 2 * seg);
    long long *out_start = (long long *)(base + tmp_bytes + 3 * seg);

    PB200_CUDA_TRY(cudaMemsetAsync(seg_sum, 0, (size_t)n_seg * 8, st));
    {
        long long g = (n + 255) / 256;
        if (g > (long long)sm_count() * 8) g = (long long)sm_count() * 8;
        seg_sum_permute_kernel<<<(unsigned)g, 256, 0, st>>>(lengths_in, world, tables_local, local_batch,
                                                            lengths_out, (unsigned long long *)seg_sum);
    }
    count_launch();
    PB200_LAUNCH_CHECK();
    seg_starts_kernel<<<1, 64, 0, st>>>(seg_sum, world, tables_local, slot_elems, in_start, in_end,
                                        out_start);
    count_launch();
    PB200_LAUNCH_CHECK();
    size_t tmp = tmp_bytes;
    PB200_CUDA_TRY(cub::DeviceScan::ExclusiveSum(base, tmp, (const long long *)lengths_out, offsets_out,
                                                 (int)n, st));
    count_launch(2);
    write_total_kernel<<<1, 32, 0, st>>>(offsets_out, lengths_out, n, offsets_out + n);
    count_launch();
    PB200_LAUNCH_CHECK();
    if (n_space > 0) {
        const long long n_chunks = (n_space + kSegCopyChunk - 1) / kSegCopyChunk;
        long long grid = n_chunks < (long long)sm_count() * 8 ? n_chunks : (long long)sm_count() * 8;
        const size_t smem = (size_t)3 * n_seg * 8;
        if (smem > 200 * 1024) return PB200_EUNSUPPORTED;
        if (smem > 48 * 1024)
            PB200_CUDA_TRY(cudaFuncSetAttribute(seg_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
        seg_copy_kernel<<<(unsigned)grid, 256, smem, st>>>(indices_in, n_space, in_start, in_end, out_start,
                                                          n_seg, indices_out, n_space);
        count_launch();
        PB200_LAUNCH_CHECK();
    }
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" int64_t pb200_regroup_scratch_bytes(int32_t world, int32_t tables_local,
                                               int64_t local_batch) {
    if (world < 1 || tables_local < 0 || local_batch < 0) return 0;
    const long long n = (long long)world * tables_local * local_batch;
    return (int64_t)(scan_tmp_bytes(n) + 4 * seg_array_bytes(world, tables_local) + 512);
}

extern "C" int pb200_regroup_sparse(const int64_t *lengths_in, const int64_t *indices_in,
                                    int64_t n_indices, int32_t world, int32_t tables_local,
                                    int64_t local_batch, int64_t *lengths_out, int64_t *offsets_out,
                                    int64_t *indices_out, void *scratch, int64_t scratch_bytes,
                                    void *stream) {
    if (!lengths_in || !lengths_out || !offsets_out || (!indices_in && n_indices > 0) ||
        (!indices_out && n_indices > 0) || !scratch)
        return PB200_EINVAL;
    if (world < 1 || tables_local < 0 || local_batch < 0 || n_indices < 0) return PB200_EINVAL;
    const long long n = (long long)world * tables_local * local_batch;
    if (n > 0x7fffffffll) return PB200_EUNSUPPORTED;
    if (scratch_bytes < pb200_regroup_scratch_bytes(world, tables_local, local_batch))
        return PB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (world * tables_local == 0 || local_batch == 0) {
        PB200_CUDA_TRY(cudaMemsetAsync(offsets_out, 0, 8, st));
        return PB200_OK;
    }
    return regroup_impl((const long long *)lengths_in, (const long long *)indices_in, n_indices, 0, world,
                        tables_local, local_batch, (long long *)lengths_out, (long long *)offsets_out,
                        (long long *)indices_out, (unsigned char *)scratch, st);
}

extern "C" int pb200_sparse_data_dist(pb200_a2a_comm *c, const int64_t *lengths, const int64_t *indices,
                                      int64_t n_indices_local, const int64_t *tables_split,
                                      int64_t local_batch, int64_t lengths_window_off,
                                      int64_t indices_window_off, int64_t slot_elems,
                                      int64_t *lengths_out, int64_t *offsets_out, int64_t *indices_out,
                                      void *scratch, int64_t scratch_bytes, void *stream) {
    if (!c || !lengths || !tables_split || !lengths_out || !offsets_out || !indices_out || !scratch)
        return PB200_EINVAL;
    if ((!indices && n_indices_local > 0) || n_indices_local < 0 || local_batch < 1 || slot_elems < 1 ||
        lengths_window_off < 0 || indices_window_off < 0)
        return PB200_EINVAL;
    const int W = c->world, me = c->rank;
    TableBases tb{};
    for (int r = 0; r < W; ++r) {
        if (tables_split[r] < 0) return PB200_EINVAL;
        tb.base[r + 1] = tb.base[r] + tables_split[r];
    }
    const int T_l = (int)tables_split[me];
    const long long b = local_batch;
    const long long n_len = (long long)W * T_l * b;
    if (n_len > 0x7fffffffll || T_l < 1) return PB200_EUNSUPPORTED;
    if (scratch_bytes < pb200_regroup_scratch_bytes(W, T_l, b)) return PB200_EINVAL;
    // window carve-up: lengths [W][T_l][b] int64, then W index slots; 16 B aligned, not overlapping
    if ((lengths_window_off | indices_window_off) & 15) return PB200_EALIGN;
    const long long len_end = lengths_window_off + n_len * 8;
    const long long idx_end = indices_window_off + (long long)W * slot_elems * 8;
    if (len_end > c->window_bytes || idx_end > c->window_bytes) return PB200_EINVAL;
    if (lengths_window_off < idx_end && indices_window_off < len_end) return PB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *sbase = (unsigned char *)scratch;
    long long *counts = (long long *)(sbase + pb200_regroup_scratch_bytes(W, T_l, b) - 256);

    // Load every kernel of this call BEFORE the first push kernel is queued.  With CUDA's lazy module
    // loading the first launch of a kernel may synchronise the context; if that happened behind a
    // push kernel that is spinning on a peer which lives in the same process (virtual ranks on one
    // GPU, tests), the peer's launches would never be issued.
    {
        static bool loaded = false;
        if (!loaded) {
            cudaFuncAttributes fa;
            PB200_CUDA_TRY(cudaFuncGetAttributes(&fa, dest_counts_kernel));
            PB200_CUDA_TRY(cudaFuncGetAttributes(&fa, seg_sum_permute_kernel));
            PB200_CUDA_TRY(cudaFuncGetAttributes(&fa, seg_starts_kernel));
            PB200_CUDA_TRY(cudaFuncGetAttributes(&fa, seg_copy_kernel));
            PB200_CUDA_TRY(cudaFuncGetAttributes(&fa, write_total_kernel));
            size_t tmp = scan_tmp_bytes(1);
            PB200_CUDA_TRY(cub::DeviceScan::ExclusiveSum(sbase, tmp, (const long long *)counts,
                                                         (long long *)counts, 1, st));   // cub's kernels
            PB200_CUDA_TRY(cudaStreamSynchronize(st));
            loaded = true;
        }
    }

    // 1. lengths: rank j receives the lengths of ITS tables from everyone (dlrm.py:768-785)
    long long in_split[PB200_A2A_MAX_RANKS], out_split[PB200_A2A_MAX_RANKS];
    for (int r = 0; r < W; ++r) {
        in_split[r] = tables_split[r] * b * 8;
        out_split[r] = (long long)T_l * b * 8;
    }
    int rc = pb200_a2a_single(c, lengths, tb.base[W] * b * 8, (const int64_t *)in_split,
                              (const int64_t *)out_split, lengths_window_off, nullptr, stream);
    if (rc != PB200_OK) return rc;

    // 2. how many indices go to each destination — stays on the device
    PB200_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)W * 8, st));
    dest_counts_kernel<<<dim3(kCountSlices, W), 256, 0, st>>>((const long long *)lengths, tb, b,
                                                             (unsigned long long *)counts);
    count_launch();
    PB200_LAUNCH_CHECK();

    // 3. indices: block sizes read from `counts` by the push kernel, fixed slot per source
    {
        A2AArgs a{};
        for (int r = 0; r < W; ++r) a.recv_off[r] = indices_window_off + (long long)r * slot_elems * 8;
        a.dev_counts = counts;
        a.dev_src = (const unsigned char *)indices;
        a.dev_elem_bytes = 8;
        a.dev_slot_bytes = slot_elems * 8;
        // the grid is sized from what this rank could send to one peer at most (host-known bounds)
        long long bound = n_indices_local < slot_elems ? n_indices_local : slot_elems;
        rc = a2a_launch_args(c, a, bound * 8, st);
        if (rc != PB200_OK) return rc;
    }

    // 4. regroup to table-major + TBE offsets (splitPerTable / lengthsToOffsets, dlrm.py:430-504, :245-251)
    const unsigned char *win = c->peer_data[me];
    return regroup_impl((const long long *)(win + lengths_window_off),
                        (const long long *)(win + indices_window_off), (long long)W * slot_elems,
                        slot_elems, W, T_l, b, (long long *)lengths_out, (long long *)offsets_out,
                        (long long *)indices_out, sbase, st);
}
