/*
 * param_oracle.c — CPU restatement of the reference algorithm for the EmbeddingBag + DLRM
 * all-to-all hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under param_b200/ imports, links or calls it; the product path
 * fails loudly when libparam_b200.so is missing instead of routing through here.
 *
 * The reference (facebookresearch/param @ 1e115ff) is 100 % Python and delegates the arithmetic of
 * this path to third-party torch (requirements.txt:1, unpinned; installed here 2.11.0+cu128):
 *   aten::_embedding_bag   via nn.EmbeddingBag(features, embdim, mode="sum")
 *                          train/compute/pt/pytorch_emb.py:179, train/comms/pt/dlrm.py:379-380
 *   c10d all_to_all_single train/comms/pt/pytorch_dist_backend.py:336-351
 * so this file restates the PUBLISHED semantics of those two ops plus the reference's own
 * index/permute logic, and is pinned (tests/test_oracle_golden.py) against golden vectors
 * generated in the build container by running the reference's own call sites — torch CPU
 * nn.EmbeddingBag, the reference's init_indices / calculateLengths / splitPerTable /
 * All2Allv_Req+Wait on 3 gloo ranks — with tests/golden/make_golden.py (committed).
 *
 * Build: gcc -O2 -std=c11 -fPIC -shared -fopenmp -ffp-contract=off (see param_b200/build.py).
 * -ffp-contract=off keeps "acc += w" a plain fp32 add so SUM mode is bit-identical to a
 * sequential in-order accumulation, which is what torch's CPU path produces (SURVEY §8c probe).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_POOL_SUM 0
#define ORACLE_POOL_MEAN 1

/* ---------------------------------------------------------------------------------------------
 * EmbeddingBag forward — semantics of nn.EmbeddingBag.__call__(indices, offsets)
 * (train/compute/pt/pytorch_emb.py:40,61; offsets built at :171-174 = start of each bag, no
 * trailing entry, last bag runs to len(indices); empty bag -> zeros).
 * mode MEAN divides the sum by the bag length; per_sample_weights multiplies each row (torch only
 * allows it with mode="sum").  fp32, accumulated in index order.
 * threads > 1 parallelises over bags (each bag is still sequential), used for the CPU baseline.
 * ------------------------------------------------------------------------------------------- */
void oracle_embbag_fwd(const float *weight, int64_t num_rows, int32_t dim, const int64_t *indices,
                       int64_t n_indices, const int64_t *offsets, int64_t n_bags,
                       int32_t include_last_offset, const float *psw, int32_t mode, float *out,
                       int64_t out_row_stride, int32_t threads) {
    (void)num_rows;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int64_t b = 0; b < n_bags; ++b) {
        const int64_t begin = offsets[b];
        const int64_t end =
            (b + 1 < n_bags || include_last_offset) ? offsets[b + 1] : n_indices;
        float *o = out + b * out_row_stride;
        for (int32_t d = 0; d < dim; ++d) o[d] = 0.0f;
        for (int64_t i = begin; i < end; ++i) {
            const float *row = weight + indices[i] * (int64_t)dim;
            if (psw) {
                const float w = psw[i];
                for (int32_t d = 0; d < dim; ++d) o[d] = o[d] + w * row[d];
            } else {
                for (int32_t d = 0; d < dim; ++d) o[d] = o[d] + row[d];
            }
        }
        if (mode == ORACLE_POOL_MEAN && end > begin) {
            const float cnt = (float)(end - begin);
            for (int32_t d = 0; d < dim; ++d) o[d] = o[d] / cnt;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * EmbeddingBag backward, dense-equivalent: dW[indices[i], :] += scale * w_i * grad[bag(i), :]
 * (autograd of the op above; LookupOut.backward, train/comms/pt/pytorch_dist_backend.py:849-857;
 * for sparse=True tables the reference gets the same values as an uncoalesced COO tensor whose
 * to_dense() equals this).  Accumulates in lookup order into double when
 * `dst64` is given (used as the high-precision comparison target), else into fp32 dst.
 * ------------------------------------------------------------------------------------------- */
void oracle_embbag_bwd(float *dst, double *dst64, int32_t dim, const int64_t *indices,
                       int64_t n_indices, const int64_t *offsets, int64_t n_bags,
                       int32_t include_last_offset, const float *psw, int32_t mode,
                       const float *grad_out, int64_t grad_row_stride, float scale) {
    for (int64_t b = 0; b < n_bags; ++b) {
        const int64_t begin = offsets[b];
        const int64_t end =
            (b + 1 < n_bags || include_last_offset) ? offsets[b + 1] : n_indices;
        const float *g = grad_out + b * grad_row_stride;
        const float inv = (mode == ORACLE_POOL_MEAN && end > begin) ? 1.0f / (float)(end - begin) : 1.0f;
        for (int64_t i = begin; i < end; ++i) {
            const float w = (psw ? psw[i] : 1.0f) * inv * scale;
            const int64_t r = indices[i] * (int64_t)dim;
            if (dst64) {
                for (int32_t d = 0; d < dim; ++d) dst64[r + d] += (double)w * (double)g[d];
            } else {
                for (int32_t d = 0; d < dim; ++d) dst[r + d] = dst[r + d] + w * g[d];
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Batched multi-table forward over the TBE request layout
 * (train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:93-135,191-213:
 * indices = cat over tables, offsets int64[T*B+1] cumulative over the concatenation, output
 * [B, sum_t D_t]).  The arithmetic lives in fbgemm_gpu (absent) => parity unpinned at that
 * boundary; restated as the per-table loop the reference itself uses in dlrm.py:363-388.
 * out element (t, b, d) at out[t*out_stride_t + b*out_stride_b + d].
 * ------------------------------------------------------------------------------------------- */
void oracle_tbe_fwd(const float *weights, const int64_t *table_row_offsets, int32_t num_tables,
                    int32_t dim, const int64_t *indices, const int64_t *offsets, int64_t batch,
                    const float *psw, int32_t mode, float *out, int64_t out_stride_t,
                    int64_t out_stride_b, int32_t threads) {
    for (int32_t t = 0; t < num_tables; ++t) {
        const float *w = weights + table_row_offsets[t] * (int64_t)dim;
        oracle_embbag_fwd(w, table_row_offsets[t + 1] - table_row_offsets[t], dim, indices, 0,
                          offsets + (int64_t)t * batch, batch, 1, psw, mode,
                          out + (int64_t)t * out_stride_t, out_stride_b, threads);
    }
}

void oracle_tbe_bwd(float *dst, double *dst64, const int64_t *table_row_offsets,
                    int32_t num_tables, int32_t dim, const int64_t *indices,
                    const int64_t *offsets, int64_t batch, const float *psw, int32_t mode,
                    const float *grad_out, int64_t go_stride_t, int64_t go_stride_b, float scale) {
    for (int32_t t = 0; t < num_tables; ++t) {
        const int64_t base = table_row_offsets[t] * (int64_t)dim;
        oracle_embbag_bwd(dst ? dst + base : NULL, dst64 ? dst64 + base : NULL, dim, indices, 0,
                          offsets + (int64_t)t * batch, batch, 1, psw, mode,
                          grad_out + (int64_t)t * go_stride_t, go_stride_b, scale);
    }
}

/* ---------------------------------------------------------------------------------------------
 * c10d all_to_all_single over W simulated ranks (train/comms/pt/pytorch_dist_backend.py:336-351;
 * semantics pinned by a live 3/4-rank gloo run): rank d's output is the
 * source-rank-major concatenation of the blocks every rank s addressed to d.
 *   in_all  : all ranks' input buffers back to back, rank s starts at byte in_rank_off[s]
 *   splits  : W*W matrix, splits[s*W + d] = BYTES rank s sends to rank d
 *   out_all : all ranks' outputs back to back, rank d starts at out_rank_off[d]
 * ------------------------------------------------------------------------------------------- */
void oracle_all_to_all_single(int32_t world, const uint8_t *in_all, const int64_t *in_rank_off,
                              const int64_t *splits, uint8_t *out_all,
                              const int64_t *out_rank_off) {
    for (int32_t d = 0; d < world; ++d) {
        int64_t o = out_rank_off[d];
        for (int32_t s = 0; s < world; ++s) {
            int64_t src = in_rank_off[s];
            for (int32_t k = 0; k < d; ++k) src += splits[(int64_t)s * world + k];
            const int64_t n = splits[(int64_t)s * world + d];
            memcpy(out_all + o, in_all + src, (size_t)n);
            o += n;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * DLRM pooled-embedding exchange, forward (train/comms/pt/dlrm.py:86-134 All2Allv_Req.forward:
 * input = cat(inputs, dim=1).view(-1) with in-splits gNS_j*sum(E) and out-splits lN*gSS_j;
 * :157-177 All2Allv_Wait.forward: per-source views [lN, T_src*E]; :1253 torch.cat(B, dim=1)).
 * Rank r holds pooled[r] = [T_r, N, E] (dlrm.py:387 torch.stack); result on rank j is
 * [lN_j, T_global*E] with tables in global order.
 *   pooled_all : rank r's [T_r, N, E] block starts at element pooled_off[r]
 *   out_all    : rank j's [lN_j, T_global*E] block starts at element out_off[j]
 * ------------------------------------------------------------------------------------------- */
void oracle_pooled_a2a_fwd(int32_t world, int32_t emb_dim, const int64_t *batch_split,
                           const int64_t *tables_split, const float *pooled_all,
                           const int64_t *pooled_off, float *out_all, const int64_t *out_off) {
    int64_t N = 0, Tg = 0;
    for (int32_t r = 0; r < world; ++r) {
        N += batch_split[r];
        Tg += tables_split[r];
    }
    int64_t nbase = 0;
    for (int32_t j = 0; j < world; ++j) {
        float *out = out_all + out_off[j];
        int64_t tbase = 0;
        for (int32_t r = 0; r < world; ++r) {
            const float *p = pooled_all + pooled_off[r];
            for (int64_t t = 0; t < tables_split[r]; ++t)
                for (int64_t n = 0; n < batch_split[j]; ++n)
                    memcpy(out + n * Tg * emb_dim + (tbase + t) * emb_dim,
                           p + (t * N + nbase + n) * emb_dim, sizeof(float) * (size_t)emb_dim);
            tbase += tables_split[r];
        }
        nbase += batch_split[j];
    }
}

/* Backward of the exchange (dlrm.py:180-218 All2Allv_Wait.backward, :137-154
 * All2Allv_Req.backward): rank j's grad [lN_j, T_global*E] -> owner r gets, per local table,
 * [N, E] (grad_inputs = view([N, -1]).split(E, dim=1)); written here as [T_r, N, E]. */
void oracle_pooled_a2a_bwd(int32_t world, int32_t emb_dim, const int64_t *batch_split,
                           const int64_t *tables_split, const float *grad_all,
                           const int64_t *grad_off, float *out_all, const int64_t *out_off) {
    int64_t N = 0, Tg = 0;
    for (int32_t r = 0; r < world; ++r) {
        N += batch_split[r];
        Tg += tables_split[r];
    }
    int64_t tbase = 0;
    for (int32_t r = 0; r < world; ++r) {
        float *out = out_all + out_off[r];
        int64_t nbase = 0;
        for (int32_t j = 0; j < world; ++j) {
            const float *g = grad_all + grad_off[j];
            for (int64_t t = 0; t < tables_split[r]; ++t)
                for (int64_t n = 0; n < batch_split[j]; ++n)
                    memcpy(out + (t * N + nbase + n) * emb_dim,
                           g + n * Tg * emb_dim + (tbase + t) * emb_dim,
                           sizeof(float) * (size_t)emb_dim);
            nbase += batch_split[j];
        }
        tbase += tables_split[r];
    }
}

/* ---------------------------------------------------------------------------------------------
 * Sparse-input regroup after the lengths / indices all-to-all
 * (train/comms/pt/dlrm.py:430-504 splitPerTable, :245-251 lengthsToOffsets).
 *   lengths_in [W][T_l][b]; indices_in in the same (rank, table, sample) order.
 *   lengths_out [T_l][W*b]: table f = concat_r lengths[r][f][:]          (:449-470)
 *   indices_out : table-major; within a table rank-major                 (:478-491)
 *   offsets_out [T_l*W*b + 1]: exclusive cumsum over lengths_out flattened; table f's
 *       per-table offsets (lengthsToOffsets, :245-251) are offsets_out[f*W*b + k] -
 *       offsets_out[f*W*b].
 * ------------------------------------------------------------------------------------------- */
void oracle_split_per_table(const int64_t *lengths_in, const int64_t *indices_in, int32_t world,
                            int32_t tables_local, int64_t local_batch, int64_t *lengths_out,
                            int64_t *offsets_out, int64_t *indices_out) {
    const int64_t b = local_batch;
    int64_t *seg_in = (int64_t *)malloc(sizeof(int64_t) * (size_t)(world * tables_local + 1));
    int64_t acc = 0;
    for (int32_t r = 0; r < world; ++r)
        for (int32_t f = 0; f < tables_local; ++f) {
            seg_in[r * tables_local + f] = acc;
            for (int64_t s = 0; s < b; ++s) acc += lengths_in[((int64_t)r * tables_local + f) * b + s];
        }
    int64_t o = 0, io = 0;
    for (int32_t f = 0; f < tables_local; ++f)
        for (int32_t r = 0; r < world; ++r) {
            const int64_t *src = lengths_in + ((int64_t)r * tables_local + f) * b;
            int64_t n = 0;
            for (int64_t s = 0; s < b; ++s) {
                lengths_out[o] = src[s];
                offsets_out[o] = io + n;
                n += src[s];
                ++o;
            }
            memcpy(indices_out + io, indices_in + seg_in[r * tables_local + f],
                   sizeof(int64_t) * (size_t)n);
            io += n;
        }
    offsets_out[o] = io;
    free(seg_in);
}

/* offsets -> lengths per feature, as calculateLengths does before the exchange
 * (train/comms/pt/dlrm.py:226-242: roll(-1), first differences, last = len(indices) - off[-1]). */
void oracle_calculate_lengths(const int64_t *offsets, int64_t n_bags, int64_t n_indices,
                              int64_t *lengths) {
    for (int64_t i = 0; i < n_bags; ++i)
        lengths[i] = (i + 1 < n_bags ? offsets[i + 1] : n_indices) - offsets[i];
}

/* ---------------------------------------------------------------------------------------------
 * Fused "exact" optimizers of the batched op the reference builds for its emb_lookup kernel
 * (train/comms/pt/comms_utils.py:2015 optimizer=OptimType.EXACT_ROWWISE_ADAGRAD;
 * train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:279-301 lr / eps).
 * fbgemm_gpu itself is not in the reference tree: this restates its published update — `grad` is
 * the dense-equivalent gradient of ONE step (oracle_tbe_bwd), every row gets one update:
 *   optimizer 1 (exact_sgd)              w -= lr * g
 *   optimizer 2 (exact_row_wise_adagrad) m[row] += mean_d(g[row,d]^2);
 *                                        w[row] -= lr / (sqrt(m[row]) + eps) * g[row]
 * float64 throughout.  Pinned against torch.optim.SGD / torch.optim.Adagrad on row-constant
 * gradients (tests/test_oracle_golden.py, tests/golden/tbe_optim_torch.npz).
 * ------------------------------------------------------------------------------------------- */
#include <math.h>

void oracle_fused_optimizer_step(double *weights, double *state, const double *grad, int64_t rows,
                                 int32_t dim, int32_t optimizer, double lr, double eps) {
    for (int64_t r = 0; r < rows; ++r) {
        const double *g = grad + r * dim;
        double *w = weights + r * dim;
        double mult = lr;
        if (optimizer == 2) {
            double ss = 0.0;
            for (int32_t d = 0; d < dim; ++d) ss += g[d] * g[d];
            state[r] += ss / (double)dim;
            mult = lr / (sqrt(state[r]) + eps);
        }
        for (int32_t d = 0; d < dim; ++d) w[d] -= mult * g[d];
    }
}
