// lib.cu — library-level entry points: version, error strings, launch accounting, device info,
// and the counter-based synthetic-data generators used by bench.py (C-ABI §8 in param_b200.h).
#include <math.h>

#include "common.cuh"

namespace pb200 {

int64_t g_launch_count = 0;

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// dst[i] = lo + (hi - lo) * u,  u = top 24 bits of mix64(seed ^ i) / 2^24  (exact in fp32)
__global__ void __launch_bounds__(256) fill_uniform_kernel(float *dst, long long n, float lo,
                                                           float span, unsigned long long seed) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v;
        const unsigned long long e = (unsigned long long)i << 2;
        v.x = lo + span * ((float)(mix64(seed ^ (e + 0)) >> 40) * (1.0f / 16777216.0f));
        v.y = lo + span * ((float)(mix64(seed ^ (e + 1)) >> 40) * (1.0f / 16777216.0f));
        v.z = lo + span * ((float)(mix64(seed ^ (e + 2)) >> 40) * (1.0f / 16777216.0f));
        v.w = lo + span * ((float)(mix64(seed ^ (e + 3)) >> 40) * (1.0f / 16777216.0f));
        ((float4 *)dst)[i] = v;
    }
    // tail
    const long long tail0 = n4 << 2;
    for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = lo + span * ((float)(mix64(seed ^ (unsigned long long)i) >> 40) *
                              (1.0f / 16777216.0f));
}

// Inverse-CDF sampling of a truncated Zipf: u in [0,1) with 53 random bits, binary search for the
// first k with cdf[k] > u.  cdf is the normalised inclusive prefix sum of k^-alpha, k = 1..num_rows
// (train/compute/pt/pytorch_emb.py:143-144 builds the same pmf).
__device__ __forceinline__ long long zipf_draw(const double *__restrict__ cdf, long long num_rows,
                                               unsigned long long key) {
    const double u = (double)(mix64(key) >> 11) * (1.0 / 9007199254740992.0);
    long long lo = 0, hi = num_rows - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cdf[mid] > u)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

// One thread per bag.  dedupe != 0 reproduces the reference's "sampling without replacement per
// bag" (pytorch_emb.py:146-157: oversample, keep the first nnz distinct values) — but keeps drawing
// past 2*nnz instead of failing when a bag has fewer than nnz distinct values among them.
__global__ void __launch_bounds__(128) fill_zipf_bags_kernel(long long *dst, long long n_bags,
                                                             int nnz,
                                                             const double *__restrict__ cdf,
                                                             long long num_rows, int dedupe,
                                                             unsigned long long seed) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < n_bags; b += stride) {
        long long *out = dst + b * nnz;
        const unsigned long long base = seed ^ ((unsigned long long)b * 0xD1B54A32D192ED03ull);
        int have = 0;
        const int max_draws = 64 * nnz + 64;
        for (int k = 0; have < nnz && k < max_draws; ++k) {
            const long long v = zipf_draw(cdf, num_rows, base + (unsigned long long)k);
            bool dup = false;
            if (dedupe)
                for (int j = 0; j < have; ++j) dup |= (out[j] == v);
            if (!dup) out[have++] = v;
        }
        // practically unreachable: top up with the smallest unused row ids
        for (long long v = 0; have < nnz && v < num_rows; ++v) {
            bool dup = false;
            for (int j = 0; j < have; ++j) dup |= (out[j] == v);
            if (!dup) out[have++] = v;
        }
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_abi_version(void) { return PB200_ABI_VERSION; }

extern "C" const char *pb200_error_string(int code) {
    switch (code) {
        case PB200_OK: return "ok";
        case PB200_EINVAL: return "pb200: invalid argument";
        case PB200_EUNSUPPORTED: return "pb200: unsupported shape or dtype";
        case PB200_EALIGN: return "pb200: pointer or stride not aligned for the vector path";
        case PB200_EBOUNDS: return "pb200: index out of range";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pb200: unknown error";
}

extern "C" int64_t pb200_launch_count(void) {
    return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED);
}

extern "C" int pb200_device_info(int *sm, int *smem_optin, int *cc_major, int *cc_minor) {
    int dev = 0;
    PB200_CUDA_TRY(cudaGetDevice(&dev));
    int v = 0;
    if (sm) {
        PB200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm = v;
    }
    if (smem_optin) {
        PB200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        *smem_optin = v;
    }
    if (cc_major) {
        PB200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        PB200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return PB200_OK;
}

extern "C" int pb200_fill_uniform(float *dst, int64_t n, float lo, float hi, uint64_t seed,
                                  void *stream) {
    if (!dst || n < 0) return PB200_EINVAL;
    if (((uintptr_t)dst & 15) != 0) return PB200_EALIGN;
    if (n == 0) return PB200_OK;
    const int grid = sm_count() * 8;
    fill_uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, n, lo, hi - lo, seed);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}

extern "C" int pb200_fill_zipf_indices(int64_t *dst, int64_t n_bags, int32_t nnz,
                                       const double *cdf_dev, int64_t num_rows, int32_t dedupe,
                                       uint64_t seed, void *stream) {
    if (!dst || !cdf_dev || n_bags < 0 || nnz < 1 || num_rows < 1) return PB200_EINVAL;
    if (dedupe && nnz > num_rows) return PB200_EINVAL;
    if (n_bags == 0) return PB200_OK;
    long long grid = (n_bags + 127) / 128;
    if (grid > (long long)sm_count() * 16) grid = (long long)sm_count() * 16;
    fill_zipf_bags_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(
        (long long *)dst, n_bags, nnz, cdf_dev, num_rows, dedupe, seed);
    count_launch();
    PB200_LAUNCH_CHECK();
    return PB200_OK;
}
