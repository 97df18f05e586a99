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
#include <cub/device/device_scan.cuh>

#include "common.cuh"

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

// single CTA: exclusive prefix of seg_sum in (r,t) order -> in_start, in (t,r) order -> out_start
__global__ void __launch_bounds__(64) seg_starts_kernel(const long long *seg_sum, int W, int T,
                                                        long long *in_start, long long *out_start) {
    // W*T is at most a few thousand: serial scan by one thread per ordering is ~microseconds
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int s = 0; s < W * T; ++s) {
            in_start[s] = acc;
            acc += seg_sum[s];
        }
        in_start[W * T] = acc;
    } else if (threadIdx.x == 32) {
        long long acc = 0;
        for (int t = 0; t < T; ++t)
            for (int r = 0; r < W; ++r) {
                out_start[r * T + t] = acc;
                acc += seg_sum[r * T + t];
            }
    }
}

// flat over all indices (input order): element i belongs to the segment s with
// in_start[s] <= i < in_start[s+1] (binary search in a shared-memory copy of in_start, re-done only
// when the running segment is left), and moves to out_start[s] + (i - in_start[s]).
constexpr int kSegCopyChunk = 256 * 8;
__global__ void __launch_bounds__(256) seg_copy_kernel(const long long *__restrict__ indices_in,
                                                       long long n_indices,
                                                       const long long *__restrict__ in_start,
                                                       const long long *__restrict__ out_start,
                                                       int n_seg, long long *__restrict__ indices_out) {
    extern __shared__ long long s_start[];   // [n_seg + 1] in_start, then [n_seg] out_start
    long long *s_out = s_start + n_seg + 1;
    for (int k = threadIdx.x; k <= n_seg; k += blockDim.x) s_start[k] = in_start[k];
    for (int k = threadIdx.x; k < n_seg; k += blockDim.x) s_out[k] = out_start[k];
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
            if (seg < 0 || i >= s_start[seg + 1]) {
                int lo = 0, hi = n_seg - 1;   // last s with in_start[s] <= i and a non-empty range
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_start[mid] <= i)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                seg = lo;
            }
            indices_out[s_out[seg] + (i - s_start[seg])] = v[k];
        }
    }
}

__global__ void write_total_kernel(const long long *offsets_out, const long long *lengths_out,
                                   long long n, long long *dst) {
    // offsets_out[n] = offsets_out[n-1] + lengths_out[n-1]
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = n > 0 ? offsets_out[n - 1] + lengths_out[n - 1] : 0;
}

static size_t scan_tmp_bytes(long long n) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const long long *)nullptr, (long long *)nullptr,
                                  (int)(n > 0x7fffffffll ? 0x7fffffff : n));
    return (tmp + 255) & ~(size_t)255;
}

}  // namespace pb200

using namespace pb200;

extern "C" int64_t pb200_regroup_scratch_bytes(int32_t world, int32_t tables_local,
                                               int64_t local_batch) {
    if (world < 1 || tables_local < 0 || local_batch < 0) return 0;
    const long long n = (long long)world * tables_local * local_batch;
    const size_t seg = ((((size_t)world * tables_local + 1) * 8) + 255) & ~(size_t)255;
    return (int64_t)(scan_tmp_bytes(n) + 3 * seg + 256);
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
    const int n_seg = world * tables_local;
    if (n_seg == 0 || local_batch == 0) {
        PB200_CUDA_TRY(cudaMemsetAsync(offsets_out, 0, 8, st));
        return PB200_OK;
    }
    unsigned char *base = (unsigned char *)scratch;
    const size_t tmp_bytes = scan_tmp_bytes(n);
    const size_t seg = ((((size_t)n_seg + 1) * 8) + 255) & ~(size_t)255;
    long long *seg_sum = (long long *)(base + tmp_bytes);
    long long *in_start = (long long *)(base + tmp_bytes + seg);
    long long *out_start = (long long *)(base + tmp_bytes + 2 * seg);

    PB200_CUDA_TRY(cudaMemsetAsync(seg_sum, 0, (size_t)n_seg * 8, st));
    {
        long long g = (n + 255) / 256;
        if (g > (long long)sm_count() * 8) g = (long long)sm_count() * 8;
        seg_sum_permute_kernel<<<(unsigned)g, 256, 0, st>>>((const long long *)lengths_in, world,
                                                            tables_local, local_batch,
                                                            (long long *)lengths_out,
                                                            (unsigned long long *)seg_sum);
    }
    count_launch();
    PB200_LAUNCH_CHECK();
    seg_starts_kernel<<<1, 64, 0, st>>>(seg_sum, world, tables_local, in_start, out_start);
    count_launch();
    PB200_LAUNCH_CHECK();
    size_t tmp = tmp_bytes;
    PB200_CUDA_TRY(cub::DeviceScan::ExclusiveSum(base, tmp, (const long long *)lengths_out,
                                                 (long long *)offsets_out, (int)n, st));
    count_launch(2);
    write_total_kernel<<<1, 32, 0, st>>>((const long long *)offsets_out,
                                         (const long long *)lengths_out, n,
                                         (long long *)offsets_out + n);
    count_launch();
    PB200_LAUNCH_CHECK();
    if (n_indices > 0) {
        const long long n_chunks = (n_indices + kSegCopyChunk - 1) / kSegCopyChunk;
        long long grid = n_chunks < (long long)sm_count() * 8 ? n_chunks : (long long)sm_count() * 8;
        const size_t smem = ((size_t)2 * n_seg + 1) * 8;
        if (smem > 200 * 1024) return PB200_EUNSUPPORTED;
        if (smem > 48 * 1024)
            PB200_CUDA_TRY(cudaFuncSetAttribute(seg_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
        seg_copy_kernel<<<(unsigned)grid, 256, smem, st>>>((const long long *)indices_in, n_indices, in_start,
                                                          out_start, n_seg, (long long *)indices_out);
        count_launch();
        PB200_LAUNCH_CHECK();
    }
    return PB200_OK;
}
