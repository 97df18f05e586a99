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

// one warp per (r, t) segment: sum of b lengths; also writes the permuted lengths
__global__ void __launch_bounds__(256) seg_sum_permute_kernel(const long long *__restrict__ lengths_in,
                                                              int W, int T, long long b,
                                                              long long *__restrict__ lengths_out,
                                                              long long *__restrict__ seg_sum) {
    const int lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= (long long)W * T) return;
    const int r = (int)(seg / T), t = (int)(seg % T);
    const long long *src = lengths_in + seg * b;
    long long *dst = lengths_out + ((long long)t * W + r) * b;
    long long s = 0;
    for (long long i = lane; i < b; i += 32) {
        const long long v = src[i];
        dst[i] = v;
        s += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) seg_sum[seg] = s;
}

// single CTA: exclusive prefix of seg_sum in (r,t) order -> in_start, in (t,r) order -> out_start
__global__ void __launch_bounds__(1024) seg_starts_kernel(const long long *seg_sum, int W, int T,
                                                          long long *in_start, long long *out_start) {
    // W*T is at most a few thousand: serial scan by one thread per ordering is ~microseconds
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int s = 0; s < W * T; ++s) {
            in_start[s] = acc;
            acc += seg_sum[s];
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

// one CTA per (r, t) segment (grid-strided): contiguous copy of the segment's indices
__global__ void __launch_bounds__(256) seg_copy_kernel(const long long *__restrict__ indices_in,
                                                       const long long *__restrict__ seg_sum,
                                                       const long long *__restrict__ in_start,
                                                       const long long *__restrict__ out_start,
                                                       int n_seg, long long *__restrict__ indices_out) {
    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const long long n = seg_sum[seg];
        const long long *src = indices_in + in_start[seg];
        long long *dst = indices_out + out_start[seg];
        for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = ld_stream_i64(src + i);
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
    const size_t seg = (((size_t)world * tables_local * 8) + 255) & ~(size_t)255;
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
    const size_t seg = (((size_t)n_seg * 8) + 255) & ~(size_t)255;
    long long *seg_sum = (long long *)(base + tmp_bytes);
    long long *in_start = (long long *)(base + tmp_bytes + seg);
    long long *out_start = (long long *)(base + tmp_bytes + 2 * seg);

    seg_sum_permute_kernel<<<(n_seg + 7) / 8, 256, 0, st>>>((const long long *)lengths_in, world,
                                                            tables_local, local_batch,
                                                            (long long *)lengths_out, seg_sum);
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
        int grid = n_seg < sm_count() * 8 ? n_seg : sm_count() * 8;
        seg_copy_kernel<<<grid, 256, 0, st>>>((const long long *)indices_in, seg_sum, in_start,
                                              out_start, n_seg, (long long *)indices_out);
        count_launch();
        PB200_LAUNCH_CHECK();
    }
    return PB200_OK;
}
