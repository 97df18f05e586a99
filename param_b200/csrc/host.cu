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
};

using namespace pb200;

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
    free(c);
    return PB200_OK;
}

extern "C" int pb200_tbe_step_host(pb200_host_ctx *c, float *weights_dev,
                                   const int64_t *table_row_offsets_dev,
                                   const int64_t *table_row_offsets_host, int32_t num_tables,
                                   int32_t dim, const int64_t *indices_host, int64_t n_indices,
                                   const int64_t *offsets_host, int64_t batch, int32_t pool_mode,
                                   float *out_host, int32_t out_layout, int32_t tables_per_group,
                                   int32_t do_bwd, float bwd_scale) {
    if (!c || !weights_dev || !table_row_offsets_dev || !indices_host || !offsets_host || !out_host)
        return PB200_EINVAL;
    if (num_tables < 1 || dim != c->dim || batch < 1 || tables_per_group < 1) return PB200_EINVAL;
    if (out_layout != 0 && out_layout != 1) return PB200_EINVAL;
    if ((long long)tables_per_group * batch > c->max_bags) return PB200_EINVAL;
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
        if (out_layout == 0) {
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
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_k));
    PB200_CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    return PB200_OK;
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
