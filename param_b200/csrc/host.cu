// host.cu — host-buffer entry point (C-ABI §7): the end-to-end call a non-torch host makes.
//
// indices / offsets / out live in HOST memory (pinned for full PCIe rate); the weight arena stays
// resident in HBM.  Tables are processed in groups; the H2D of group g+1, the lookup kernel of
// group g and the D2H of group g-1 run on three streams with double-buffered device staging.
#include <stdlib.h>

#include "common.cuh"

struct pb200_host_ctx {
    long long max_idx;   // index elements per staging buffer
    long long max_bags;  // bags per staging buffer
    int dim;
    long long *d_idx[2];
    long long *d_off[2];
    float *d_out[2];
    cudaStream_t s_h2d, s_k, s_d2h;
    cudaEvent_t ev_h2d[2], ev_k[2], ev_d2h[2], ev_bwd[2];
    void *bwd_scratch;          // sort plan + reducer scratch of the backward (grown on first use)
    long long bwd_scratch_bytes;
    double *d_loss;             // loss form: [tables] sums + [tables][kLossChunks] partial sums (grown on first use)
    long long loss_tables;
};

using namespace pb200;

// ---- loss form: per-table sum of the pooled vectors, the scalar(s) a training step hands back --------
// Two fixed-shape stages so that the value does not depend on scheduling: CTA (c, t) sums slice c of table t's
// [B, dim] pooled block in double (strided float4 reads, xor-shuffle tree, warps combined in order), then one
// thread per table adds the kLossChunks partial sums in order.
constexpr int kLossChunks = 32;
constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads) pooled_sum_kernel(const float *__restrict__ pooled,
                                                                   long long per_table, double *partial) {
    __shared__ double s_w[kLossThreads / 32];
    const int t = blockIdx.y, c = blockIdx.x;
    const long long n4 = per_table >> 2;
    const long long per = (n4 + kLossChunks - 1) / kLossChunks;
    const long long lo = (long long)c * per;
    const long long hi = lo + per < n4 ? lo + per : n4;
    const float4 *src = (const float4 *)(pooled + (long long)t * per_table);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long i = lo + threadIdx.x;
    for (; i + 3 * kLossThreads < hi; i += 4 * kLossThreads) {
        const float4 v0 = __ldg(src + i), v1 = __ldg(src + i + kLossThreads);
        const float4 v2 = __ldg(src + i + 2 * kLossThreads), v3 = __ldg(src + i + 3 * kLossThreads);
        a0 += ((double)v0.x + (double)v0.y) + ((double)v0.z + (double)v0.w);
        a1 += ((double)v1.x + (double)v1.y) + ((double)v1.z + (double)v1.w);
        a2 += ((double)v2.x + (double)v2.y) + ((double)v2.z + (double)v2.w);
        a3 += ((double)v3.x + (double)v3.y) + ((double)v3.z + (double)v3.w);
    }
    for (; i < hi; i += kLossThreads) {
        const float4 v = __ldg(src + i);
        a0 += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) sum += s_w[w];
        partial[(long long)t * kLossChunks + c] = sum;
    }
}

__global__ void loss_final_kernel(const double *partial, double *loss, int num_tables) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tables) return;
    double sum = 0.0;
#pragma unroll
    for (int c = 0; c < kLossChunks; ++c) sum += partial[(long long)t * kLossChunks + c];
    loss[t] = sum;
}

extern "C" int pb200_host_ctx_destroy(pb200_host_ctx *c);

extern "C" int pb200_host_ctx_create(pb200_host_ctx **ctx, int64_t max_indices_per_group,
                                     int64_t max_bags_per_group, int32_t dim) {
    if (!ctx || max_indices_per_group < 1 || max_bags_per_group < 1 || dim < 1) return PB200_EINVAL;
    pb200_host_ctx *c = (pb200_host_ctx *)calloc(1, sizeof(pb200_host_ctx));
    if (!c) return PB200_EINVAL;
    c->max_idx = max_indices_per_group;
    c->max_bags = max_bags_per_group;
    c->dim = dim;
    // any failure below releases what was created so far (calloc zeroed every handle)
    auto create = [&]() -> int {
        for (int i = 0; i < 2; ++i) {
            PB200_CUDA_TRY(cudaMalloc(&c->d_idx[i], (size_t)(c->max_idx + 4) * 8));
            PB200_CUDA_TRY(cudaMalloc(&c->d_off[i], (size_t)(c->max_bags + 4) * 8));
            PB200_CUDA_TRY(cudaMalloc(&c->d_out[i], (size_t)c->max_bags * dim * 4));
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
            PB200_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_bwd[i], cudaEventDisableTiming));
        }
        PB200_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        PB200_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
        PB200_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        return PB200_OK;
    };
    const int rc = create();
    if (rc != PB200_OK) {
        pb200_host_ctx_destroy(c);
        return rc;
    }
    *ctx = c;
    return PB200_OK;
}

extern "C" int pb200_host_ctx_destroy(pb200_host_ctx *c) {
    if (!c) return PB200_OK;
    for (int i = 0; i < 2; ++i) {
        if (c->d_idx[i]) cudaFree(c->d_idx[i]);
        if (c->d_off[i]) cudaFree(c->d_off[i]);
        if (c->d_out[i]) cudaFree(c->d_out[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
        if (c->ev_bwd[i]) cudaEventDestroy(c->ev_bwd[i]);
    }
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_k) cudaStreamDestroy(c->s_k);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    if (c->bwd_scratch) cudaFree(c->bwd_scratch);
    if (c->d_loss) cudaFree(c->d_loss);
    free(c);
    return PB200_OK;
}

// out_host != nullptr: the pooled vectors go back to the host (layout out_layout); loss_host != nullptr: only the
// per-table sums do ([num_tables] doubles) and the pooled vectors never leave the device
static int step_host_impl(pb200_host_ctx *c, float *weights_dev, const int64_t *table_row_offsets_dev,
                          const int64_t *table_row_offsets_host, int32_t num_tables, int32_t dim,
                          const int64_t *indices_host, int64_t n_indices, const int64_t *offsets_host,
                          int64_t batch, int32_t pool_mode, float *out_host, int32_t out_layout,
                          double *loss_host, int32_t tables_per_group, int32_t do_bwd, float bwd_scale) {
    if (!c || !weights_dev || !table_row_offsets_dev || !indices_host || !offsets_host) return PB200_EINVAL;
    if ((out_host == nullptr) == (loss_host == nullptr)) return PB200_EINVAL;
    if (num_tables < 1 || dim != c->dim || batch < 1 || tables_per_group < 1) return PB200_EINVAL;
    if (out_layout != 0 && out_layout != 1) return PB200_EINVAL;
    if ((long long)tables_per_group * batch > c->max_bags) return PB200_EINVAL;
    if (loss_host) {
        if (dim % 4 != 0) return PB200_EUNSUPPORTED;
        out_layout = 1;                              // table-major staging: a table's pooled block is contiguous
        if (num_tables > c->loss_tables) {
            PB200_CUDA_TRY(cudaStreamSynchronize(c->s_k));
            if (c->d_loss) cudaFree(c->d_loss);
            c->d_loss = nullptr;
            c->loss_tables = 0;
            PB200_CUDA_TRY(cudaMalloc(&c->d_loss, (size_t)num_tables * (kLossChunks + 1) * sizeof(double)));
            c->loss_tables = num_tables;
        }
    }
    double *d_partial = c->d_loss ? c->d_loss + num_tables : nullptr;
    // bulk staging wants absolute even index positions 16 B-aligned: keep batch even or fall back
    const int algo = PB200_FWD_AUTO;
    if (do_bwd) {
        // the backward is the same SORTED pipeline the device-resident path uses (sort plan + one
        // segmented reduce); its scratch belongs to the context and is sized for the largest group
        long long need = 0;
        for (int t0 = 0; t0 < num_tables; t0 += tables_per_group) {
            const int tg = (t0 + tables_per_group <= num_tables) ? tables_per_group : num_tables - t0;
            const long long n = offsets_host[(long long)(t0 + tg) * batch] - offsets_host[(long long)t0 * batch];
            const long long b = pb200_tbe_bwd_scratch_bytes(n, tg, batch, 0, PB200_BWD_SORTED);
            if (b > need) need = b;
        }
        if (need > c->bwd_scratch_bytes) {
            PB200_CUDA_TRY(cudaStreamSynchronize(c->s_k));
            if (c->bwd_scratch) cudaFree(c->bwd_scratch);
            c->bwd_scratch = nullptr;
            c->bwd_scratch_bytes = 0;
            PB200_CUDA_TRY(cudaMalloc(&c->bwd_scratch, (size_t)need));
            c->bwd_scratch_bytes = need;
        }
    }
    int g = 0;
    for (int t0 = 0; t0 < num_tables; t0 += tables_per_group, ++g) {
        const int b = g & 1;
        const int tg = (t0 + tables_per_group <= num_tables) ? tables_per_group : num_tables - t0;
        const long long bag0 = (long long)t0 * batch, nb = (long long)tg * batch;
        const long long i_lo = offsets_host[bag0], i_hi = offsets_host[bag0 + nb];
        const long long n = i_hi - i_lo;
        if (n < 0 || i_hi > n_indices || n > c->max_idx) return PB200_EINVAL;
        long long max_rows = 0;   // bound of the group's largest table (0 = unknown: full 32-bit sort)
        if (table_row_offsets_host)
            for (int t = t0; t < t0 + tg; ++t) {
                const long long r = table_row_offsets_host[t + 1] - table_row_offsets_host[t];
                if (r > max_rows) max_rows = r;
            }
        // staging buffer b must be free: its previous D2H (group g-2) finished
        if (g >= 2) {
            PB200_CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_d2h[b], 0));
            if (do_bwd) PB200_CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_bwd[b], 0));
        }
        // element i_lo sits at d_idx[pad] with pad = i_lo & 1, so that the pointer handed to the
        // kernel (d_idx + pad - i_lo) keeps every even absolute position 16 B-aligned
        const long long pad = i_lo & 1;
        if (n > 0)
            PB200_CUDA_TRY(cudaMemcpyAsync(c->d_idx[b] + pad, indices_host + i_lo, (size_t)n * 8,
                                           cudaMemcpyHostToDevice, c->s_h2d));
        PB200_CUDA_TRY(cudaMemcpyAsync(c->d_off[b], offsets_host + bag0, (size_t)(nb + 1) * 8,
                                       cudaMemcpyHostToDevice, c->s_h2d));
        PB200_CUDA_TRY(cudaEventRecord(c->ev_h2d[b], c->s_h2d));

        PB200_CUDA_TRY(cudaStreamWaitEvent(c->s_k, c->ev_h2d[b], 0));
        const long long st_t = out_layout == 0 ? dim : batch * dim;
        const long long st_b = out_layout == 0 ? (long long)tg * dim : dim;
        int rc = pb200_tbe_fwd(weights_dev, table_row_offsets_dev + t0, tg, dim,
                               c->d_idx[b] + pad - i_lo, i_hi, c->d_off[b], batch, PB200_IDX_I64,
                               nullptr, pool_mode, c->d_out[b], st_t, st_b, algo, c->s_k);
        if (rc != PB200_OK) return rc;
        if (loss_host) {
            pooled_sum_kernel<<<dim3(kLossChunks, (unsigned)tg), kLossThreads, 0, c->s_k>>>(
                c->d_out[b], (long long)batch * dim, d_partial + (long long)t0 * kLossChunks);
            PB200_LAUNCH_CHECK();
            count_launch(1);
        }
        PB200_CUDA_TRY(cudaEventRecord(c->ev_k[b], c->s_k));
        if (do_bwd) {
            // training step: the pooled vectors double as the incoming gradient (in a real model
            // dOut is produced on the device by the layers above); scatter-add into the arena
            rc = pb200_tbe_bwd(weights_dev, table_row_offsets_dev + t0, tg, dim,
                               c->d_idx[b] + pad - i_lo, n, c->d_off[b], batch, PB200_IDX_I64,
                               nullptr, pool_mode, c->d_out[b], st_t, st_b, bwd_scale,
                               n > 0 ? PB200_BWD_SORTED : PB200_BWD_ATOMIC, max_rows,
                               n > 0 ? c->bwd_scratch : nullptr, n > 0 ? c->bwd_scratch_bytes : 0, 0,
                               c->s_k);
            if (rc != PB200_OK) return rc;
            PB200_CUDA_TRY(cudaEventRecord(c->ev_bwd[b], c->s_k));
        }

        PB200_CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_k[b], 0));
        if (loss_host) {
            // nothing to copy back per group: the event below only marks d_out[b] as consumed
        } else if (out_layout == 0) {
            // device [B, tg*dim] -> host columns [t0*dim, (t0+tg)*dim) of [B, T*dim]
            PB200_CUDA_TRY(cudaMemcpy2DAsync(out_host + (long long)t0 * dim,
                                             (size_t)num_tables * dim * 4, c->d_out[b],
                                             (size_t)tg * dim * 4, (size_t)tg * dim * 4,
                                             (size_t)batch, cudaMemcpyDeviceToHost, c->s_d2h));
        } else {
            PB200_CUDA_TRY(cudaMemcpyAsync(out_host + (long long)t0 * batch * dim, c->d_out[b],
                                           (size_t)nb * dim * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        }
        PB200_CUDA_TRY(cudaEventRecord(c->ev_d2h[b], c->s_d2h));
        // buffer reuse: H2D of group g+2 waits ev_d2h[b] (top of the loop); the kernel of group
        // g+2 waits that H2D, hence transitively this D2H — no extra edge needed, and kernel g+1
        // is free to overlap this D2H.
    }
    if (loss_host) {
        loss_final_kernel<<<(num_tables + 127) / 128, 128, 0, c->s_k>>>(d_partial, c->d_loss, num_tables);
        PB200_LAUNCH_CHECK();
        count_launch(1);
        PB200_CUDA_TRY(cudaMemcpyAsync(loss_host, c->d_loss, (size_t)num_tables * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->s_k));
    }
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_k));
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    return PB200_OK;
}

extern "C" int64_t pb200_pooled_sum_scratch_bytes(int64_t n_blocks) {
    return n_blocks > 0 ? n_blocks * kLossChunks * (int64_t)sizeof(double) : 0;
}

extern "C" int pb200_pooled_sum(const float *pooled_dev, int64_t n_blocks, int64_t block_elems,
                                double *sums_dev, void *scratch_dev, int64_t scratch_bytes, void *stream) {
    if (!pooled_dev || !sums_dev || n_blocks < 1 || block_elems < 1) return PB200_EINVAL;
    if (block_elems % 4 != 0 || ((uintptr_t)pooled_dev & 15) != 0) return PB200_EUNSUPPORTED;
    if (n_blocks > 65535 * 1024ll || n_blocks > 0x7fffffffll) return PB200_EUNSUPPORTED;
    if (!scratch_dev || scratch_bytes < pb200_pooled_sum_scratch_bytes(n_blocks)) return PB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = (double *)scratch_dev;
    for (long long b0 = 0; b0 < n_blocks; b0 += 65535) {          // grid.y limit
        const long long nb = n_blocks - b0 < 65535 ? n_blocks - b0 : 65535;
        pooled_sum_kernel<<<dim3(kLossChunks, (unsigned)nb), kLossThreads, 0, st>>>(
            pooled_dev + b0 * block_elems, block_elems, partial + b0 * kLossChunks);
        PB200_LAUNCH_CHECK();
        count_launch(1);
    }
    loss_final_kernel<<<(unsigned)((n_blocks + 127) / 128), 128, 0, st>>>(partial, sums_dev, (int)n_blocks);
    PB200_LAUNCH_CHECK();
    count_launch(1);
    return PB200_OK;
}

extern "C" int pb200_tbe_step_host(pb200_host_ctx *c, float *weights_dev,
                                   const int64_t *table_row_offsets_dev,
                                   const int64_t *table_row_offsets_host, int32_t num_tables,
                                   int32_t dim, const int64_t *indices_host, int64_t n_indices,
                                   const int64_t *offsets_host, int64_t batch, int32_t pool_mode,
                                   float *out_host, int32_t out_layout, int32_t tables_per_group,
                                   int32_t do_bwd, float bwd_scale) {
    if (!out_host) return PB200_EINVAL;
    return step_host_impl(c, weights_dev, table_row_offsets_dev, table_row_offsets_host, num_tables, dim,
                          indices_host, n_indices, offsets_host, batch, pool_mode, out_host, out_layout,
                          nullptr, tables_per_group, do_bwd, bwd_scale);
}

extern "C" int pb200_tbe_step_host_loss(pb200_host_ctx *c, float *weights_dev,
                                        const int64_t *table_row_offsets_dev,
                                        const int64_t *table_row_offsets_host, int32_t num_tables,
                                        int32_t dim, const int64_t *indices_host, int64_t n_indices,
                                        const int64_t *offsets_host, int64_t batch, int32_t pool_mode,
                                        double *loss_host, int32_t tables_per_group, int32_t do_bwd,
                                        float bwd_scale) {
    if (!loss_host) return PB200_EINVAL;
    return step_host_impl(c, weights_dev, table_row_offsets_dev, table_row_offsets_host, num_tables, dim,
                          indices_host, n_indices, offsets_host, batch, pool_mode, nullptr, 1, loss_host,
                          tables_per_group, do_bwd, bwd_scale);
}

extern "C" int pb200_tbe_fwd_host(pb200_host_ctx *c, const float *weights_dev,
                                  const int64_t *table_row_offsets_dev,
                                  const int64_t *table_row_offsets_host, int32_t num_tables,
                                  int32_t dim, const int64_t *indices_host, int64_t n_indices,
                                  const int64_t *offsets_host, int64_t batch, int32_t pool_mode,
                                  float *out_host, int32_t out_layout, int32_t tables_per_group) {
    return pb200_tbe_step_host(c, (float *)weights_dev, table_row_offsets_dev,
                               table_row_offsets_host, num_tables, dim, indices_host, n_indices,
                               offsets_host, batch, pool_mode, out_host, out_layout,
                               tables_per_group, 0, 0.f);
}
