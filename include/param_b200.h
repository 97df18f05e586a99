/*
 * param_b200.h — C-ABI of libparam_b200.so
 *
 * B200-native (sm_100a) replacement for the ONE hot path PARAM exercises for
 * recommendation workloads: EmbeddingBag forward/backward and the DLRM sparse
 * all-to-all.  Every entry point is `extern "C"`, takes plain pointers, sizes
 * and an opaque `cudaStream_t` (as void*), never a torch type, and is
 * asynchronous on the given stream unless its name ends in `_host`.
 *
 * All reference citations are relative to facebookresearch/param @ 1e115ff.
 *
 * Return value: 0 (PB200_OK) on success, a negative PB200_E* code for argument
 * errors, or a positive cudaError_t value when the CUDA runtime reported an
 * error.  pb200_error_string() turns either into text.  The Python host layer
 * (param_b200/_cabi.py) raises on any non-zero return.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device every
 * compute entry returns a cudaError_t.
 */
#ifndef PARAM_B200_H_
#define PARAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------- */
#define PB200_OK 0
#define PB200_EINVAL -1      /* bad argument (null pointer, negative size …)   */
#define PB200_EUNSUPPORTED -2 /* shape/dtype outside what the kernels cover     */
#define PB200_EALIGN -3      /* pointer not aligned as the vector path needs    */
#define PB200_EBOUNDS -4     /* index outside [0, num_rows) (checked entry only) */

/* ---- enums --------------------------------------------------------------- */
/* pooling mode: nn.EmbeddingBag(mode=...) — train/compute/pt/pytorch_emb.py:179 */
#define PB200_POOL_SUM 0
#define PB200_POOL_MEAN 1

/* index element type at the boundary: the reference is int64 everywhere
 * (pytorch_emb.py:158,172; dlrm.py:467,489,803); int32 is accepted too. */
#define PB200_IDX_I64 0
#define PB200_IDX_I32 1

/* forward kernel variants (all produce identical bits) */
#define PB200_FWD_AUTO 0
#define PB200_FWD_DIRECT 1 /* one lane-group per bag, indices read straight from HBM  */
#define PB200_FWD_STAGED 2 /* persistent CTAs, cp.async.bulk (TMA) index staging      */
#define PB200_FWD_PIPELINED 3 /* persistent CTAs, register software pipeline over bags  */

/* backward kernel variants */
#define PB200_BWD_AUTO 0
#define PB200_BWD_ATOMIC 1 /* red.global.add.v4.f32 per lookup (order not fixed)       */
#define PB200_BWD_SORTED 2 /* radix sort by row, segmented reduce, one RMW per row     */
#define PB200_BWD_EXACT 3  /* same sort; every touched row written exactly once, no atomics
                            * (deterministic; section 3b)                                */

/* table element type (fbgemm SparseType weights_precision,
 * split_table_batched_embeddings_ops.py:291) */
#define PB200_W_F32 0
#define PB200_W_F16 1

/* optimizer fused into the backward (fbgemm OptimType, comms_utils.py:2015,
 * split_table_batched_embeddings_ops.py:290) */
#define PB200_OPT_SGD 1             /* exact_sgd:            w -= lr * g                          */
#define PB200_OPT_ROWWISE_ADAGRAD 2 /* exact_row_wise_adagrad: m += mean_d(g_d^2);
                                     *                         w -= lr / (sqrt(m) + eps) * g     */

/* ---- library info -------------------------------------------------------- */
int pb200_abi_version(void);
const char *pb200_error_string(int code);
/* number of kernels this library has launched in this process (all entries);
 * bench.py reports the delta over the timed region as "gpu_launches". */
int64_t pb200_launch_count(void);
/* SM count / shared memory per block (optin) of the current device. */
int pb200_device_info(int *sm_count, int *smem_optin_bytes, int *cc_major, int *cc_minor);

/* =========================================================================
 * 1. EmbeddingBag forward
 * =========================================================================
 * Replaces: torch.nn.EmbeddingBag(features, embdim, mode="sum").__call__
 *   (indices, offsets)              train/compute/pt/pytorch_emb.py:61,179
 *   E(sparse_index_group_batch, sparse_offset_group_batch)
 *                                   train/comms/pt/dlrm.py:379-380
 *
 * out[b, :] = sum_{i = offsets[b]}^{end(b)-1} psw[i] * weight[indices[i], :]
 *   end(b) = offsets[b+1] for b < n_bags-1, else n_indices   (nn.EmbeddingBag:
 *   offsets has n_bags entries, last bag runs to the end of `indices`).
 *   If include_last_offset != 0, offsets has n_bags+1 entries.
 *   Empty bag -> zeros.  mode MEAN divides by the bag length (0 -> zeros).
 *   fp32 accumulation in index order (bit-identical to the CPU reference for
 *   SUM without per-sample weights).
 *
 * weight  : fp32 [num_rows, dim] row-major, device
 * indices : int64/int32 [n_indices], device
 * offsets : same integer type as indices, device
 * psw     : fp32 [n_indices] per_sample_weights or NULL
 * out     : fp32, row b at out + b*out_row_stride (elements); out_row_stride >= dim
 */
int pb200_embbag_fwd(const float *weight, int64_t num_rows, int32_t dim,
                     const void *indices, int64_t n_indices,
                     const void *offsets, int64_t n_bags, int32_t include_last_offset,
                     int32_t idx_type, const float *psw, int32_t pool_mode,
                     float *out, int64_t out_row_stride, int32_t algo, void *stream);

/* =========================================================================
 * 2. Batched multi-table EmbeddingBag forward (TBE request layout)
 * =========================================================================
 * Replaces: SplitTableBatchedEmbeddingBagsCodegen.forward(indices, offsets,
 *   per_sample_weights)   train/compute/python/workloads/pytorch/
 *   split_table_batched_embeddings_ops.py:311-313 (layout :93-135,191-213),
 *   backendFunctions.emb_lookup fwd   train/comms/pt/pytorch_dist_backend.py:832-848,
 *   and the per-table loop paramDLRM_Net.apply_emb   train/comms/pt/dlrm.py:363-388.
 *
 * T tables share one fp32 arena: table t occupies rows
 * [table_row_offsets[t], table_row_offsets[t+1]) of `weights` ([sum rows, dim]).
 * The arena must hold fewer than 2^32 rows (row ids are 32-bit inside the kernels).
 * indices = cat_t(indices_t) (table-major), offsets int[T*B + 1] cumulative over
 * the concatenation, bag (t, b) = offsets[t*B + b .. t*B + b + 1).
 * Output element (t, b, d) is written at out[t*out_stride_t + b*out_stride_b + d]:
 *   TBE / a2a-ready layout [B, T*dim]:  out_stride_t = dim,  out_stride_b = T*dim
 *   dlrm.py torch.stack layout [T, B, dim]: out_stride_t = B*dim, out_stride_b = dim
 */
int pb200_tbe_fwd(const float *weights, const int64_t *table_row_offsets /* device, T+1 */,
                  int32_t num_tables, int32_t dim,
                  const void *indices, int64_t n_indices,
                  const void *offsets /* T*B+1 */, int64_t batch, int32_t idx_type,
                  const float *psw, int32_t pool_mode,
                  float *out, int64_t out_stride_t, int64_t out_stride_b,
                  int32_t algo, void *stream);

/* Same, fp16 tables (weights_precision = fp16): weights_f16 is a __half arena [sum rows, dim],
 * dim % 4 == 0, 8 B aligned.  Rows are converted to fp32 on load; accumulation and output are
 * fp32, so the result equals pb200_tbe_fwd on the table converted to fp32, bit for bit. */
int pb200_tbe_fwd_f16(const void *weights_f16, const int64_t *table_row_offsets,
                      int32_t num_tables, int32_t dim,
                      const void *indices, int64_t n_indices,
                      const void *offsets, int64_t batch, int32_t idx_type,
                      const float *psw, int32_t pool_mode,
                      float *out, int64_t out_stride_t, int64_t out_stride_b, void *stream);

/* Debug-mode bounds check (ATen raises on an out-of-range index; the fast path
 * does not check).  Writes the number of offending lookups to *bad_count_dev
 * (device int64, zeroed by this call).  Table-relative indices, same layout as
 * pb200_tbe_fwd (num_tables = 1, batch = n_bags for the single-table op). */
int pb200_check_indices(const int64_t *table_row_offsets, int32_t num_tables,
                        const void *indices, int64_t n_indices,
                        const void *offsets, int64_t batch, int32_t idx_type,
                        int64_t *bad_count_dev, void *stream);

/* =========================================================================
 * 3. EmbeddingBag backward (scatter-add of the pooled gradient into the table)
 * =========================================================================
 * Replaces: autograd of nn.EmbeddingBag — LookupOut.backward(grad_output)
 *   train/comms/pt/pytorch_dist_backend.py:849-857, tempB.backward(C)
 *   train/comms/pt/dlrm.py:1296, …CodegenOp.backward
 *   split_table_batched_embeddings_ops.py:318-324.
 *
 * dst[row(t, indices[i]), :] += scale * psw[i] * grad_out[(t, bag(i)), :]
 *   dst == a dense grad buffer shaped like the arena (scale = 1)  -> dW, or
 *   dst == the weight arena itself (scale = -lr)                  -> fused SGD.
 * grad_out element (t, b, d) at grad_out[t*go_stride_t + b*go_stride_b + d].
 * MEAN mode divides by the bag length.
 *
 * PB200_BWD_SORTED / PB200_BWD_EXACT need scratch: query the size with
 * pb200_tbe_bwd_scratch_bytes() and pass a device buffer of that size.
 * PB200_BWD_EXACT is pb200_tbe_bwd_fused (section 3b) with SGD, lr = -scale, fp32 rows.
 *
 * max_table_rows: an upper bound of the row count of the largest table (host value; the caller
 *   built table_row_offsets, so it knows).  It fixes the number of radix passes of the sort
 *   (20 key bits for 1 M rows = two 10-bit passes); <= 0 = unknown, the sort then covers all 32
 *   bits of a row id.  Nothing is read back from the device: the call never synchronises and is
 *   CUDA-graph capturable.
 * plan_ready != 0: `scratch` already holds the sort plan that pb200_tbe_plan_build (section 3c)
 *   produced for exactly these indices / offsets / psw / pool_mode / gradient strides; the call
 *   then launches the segmented reduce only.
 */
int64_t pb200_tbe_bwd_scratch_bytes(int64_t n_indices, int32_t num_tables, int64_t batch,
                                    int64_t total_rows, int32_t algo);
int pb200_tbe_bwd(float *dst, const int64_t *table_row_offsets, int32_t num_tables, int32_t dim,
                  const void *indices, int64_t n_indices,
                  const void *offsets, int64_t batch, int32_t idx_type,
                  const float *psw, int32_t pool_mode,
                  const float *grad_out, int64_t go_stride_t, int64_t go_stride_b,
                  float scale, int32_t algo, int64_t max_table_rows,
                  void *scratch, int64_t scratch_bytes, int32_t plan_ready, void *stream);

/* Sparse form of the single-table backward: the VALUES of the uncoalesced COO gradient that
 * nn.EmbeddingBag(sparse=True) hands to autograd (the reference allocates its tables that way,
 * train/comms/pt/pytorch_dist_backend.py:923-934; ATen: _embedding_bag_sparse_backward).
 *   values[i, :] = psw[i] * grad_out[bag(i), :]   (MEAN: divided by the bag length), i < n_indices
 * The COO indices are the lookup indices themselves, so the gradient of a table costs
 * n_indices * dim * 4 bytes instead of a dense zero-filled [num_rows, dim] buffer.
 * grad_out: fp32, row b at grad_out + b*go_row_stride; offsets as in pb200_embbag_fwd. */
int pb200_embbag_bwd_sparse(const float *grad_out, int64_t go_row_stride, int32_t dim,
                            const void *offsets, int64_t n_bags, int32_t include_last_offset,
                            int64_t n_indices, int32_t idx_type, const float *psw, int32_t pool_mode,
                            float *values /* [n_indices, dim] */, void *stream);

/* The segmented reduce of pb200_tbe_bwd(PB200_BWD_SORTED, plan_ready = 1) restricted to tables
 * [table_lo, table_hi) of the request: the sorted range of those tables is read from `offsets` on the device
 * (no host knowledge of lookup counts needed).  Lets the backward of a table group start as soon as ITS
 * gradient columns have arrived (pb200_a2a_pooled_bwd_part).  The plan must have been built for the whole
 * request (pb200_tbe_plan_build); unweighted or weighted, SUM or MEAN. */
int pb200_tbe_bwd_tables(float *dst, const int64_t *table_row_offsets, int32_t num_tables, int32_t dim,
                         const void *indices, int64_t n_indices, const void *offsets, int64_t batch,
                         int32_t idx_type, const float *psw, int32_t pool_mode, const float *grad_out,
                         int64_t go_stride_t, int64_t go_stride_b, float scale, int32_t table_lo,
                         int32_t table_hi, void *plan, int64_t plan_bytes, void *stream);

/* =========================================================================
 * 3b. Backward with the optimizer fused in ("exact": one update per touched row)
 * =========================================================================
 * Replaces: the fused backward + optimizer of fbgemm's
 *   SplitTableBatchedEmbeddingBagsCodegen(optimizer=OptimType.EXACT_ROWWISE_ADAGRAD)
 *   as built at train/comms/pt/comms_utils.py:1995-2017 and
 *   split_table_batched_embeddings_ops.py:279-301 (optimizer, weights_precision, lr, eps,
 *   stochastic_rounding), driven by LookupOut.backward(grad_output)
 *   pytorch_dist_backend.py:849-857 / ...CodegenOp.backward :318-324.
 *
 * g[row] = sum over ALL lookups i of the request with row(t, indices[i]) == row of
 *          psw[i] * grad_out[(t, bag(i)), :]            (MEAN: divided by the bag length)
 * then ONE update per touched row:
 *   PB200_OPT_SGD             w[row] -= lr * g[row]       (lr = -1 on a zeroed fp32 buffer
 *                                                         gives the dense gradient itself)
 *   PB200_OPT_ROWWISE_ADAGRAD state[row] += mean_d(g[row][d]^2);
 *                             w[row] -= lr / (sqrt(state[row]) + eps) * g[row]
 * No atomics: results are run-to-run identical.  weights: fp32 (16 B aligned) or fp16 (8 B
 * aligned) arena, dim % 4 == 0, dim <= 512.  state: fp32 [sum rows] (ROWWISE_ADAGRAD only).
 * stochastic_rounding != 0 (fp16 tables only): the updated weight is rounded to fp16
 * stochastically with a counter-based generator keyed on (sr_seed, element) — pass a new
 * sr_seed per step; 0 = round to nearest even.
 * scratch: pb200_tbe_bwd_fused_scratch_bytes() bytes of device memory.
 * max_table_rows, plan_ready: as for pb200_tbe_bwd.
 */
int64_t pb200_tbe_bwd_fused_scratch_bytes(int64_t n_indices, int32_t num_tables, int64_t batch,
                                          int32_t dim);
int pb200_tbe_bwd_fused(void *weights, int32_t weights_type, float *state,
                        const int64_t *table_row_offsets, int32_t num_tables, int32_t dim,
                        const void *indices, int64_t n_indices,
                        const void *offsets, int64_t batch, int32_t idx_type,
                        const float *psw, int32_t pool_mode,
                        const float *grad_out, int64_t go_stride_t, int64_t go_stride_b,
                        int32_t optimizer, float lr, float eps,
                        int32_t stochastic_rounding, uint64_t sr_seed, int64_t max_table_rows,
                        void *scratch, int64_t scratch_bytes, int32_t plan_ready, void *stream);

/* =========================================================================
 * 3c. Sort plan: the index-only half of the sort-based backward, ahead of the gradient
 * =========================================================================
 * Replaces: the radix sort + index bookkeeping inside the autograd backward of
 *   nn.EmbeddingBag / fbgemm TBE (train/comms/pt/pytorch_dist_backend.py:849-857,
 *   dlrm.py:1296), which the reference can only start once the gradient has arrived.
 *
 * The sort of the lookups by row depends on indices / offsets only.  This call builds it into
 * `scratch` (same buffer and size as the backward call that will consume it, SORTED or EXACT)
 * on `stream` — typically a side stream, while the forward lookup of the same request runs —
 * and the backward is then called with plan_ready = 1 and launches the segmented reduce alone.
 * Hand-written per-table LSD radix sort (param_b200/csrc/radix_sort.cu): the table is known from the
 * position, so only the table-relative row is sorted (ceil(bits(max_table_rows) / 10) passes);
 * the first pass reads the indices and derives the gradient-row offset of each lookup from the
 * offsets, the last pass emits arena rows: one globally sorted array.  Stable: equal rows keep
 * request order, so the reducers' summation order is fixed.  psw / pool_mode / go_stride_* must
 * be the ones the backward call will use (they are folded into the plan's values).
 */
int pb200_tbe_plan_build(void *scratch, int64_t scratch_bytes,
                         const int64_t *table_row_offsets, int64_t max_table_rows,
                         int32_t num_tables, int32_t dim,
                         const void *indices, int64_t n_indices,
                         const void *offsets, int64_t batch, int32_t idx_type,
                         const float *psw, int32_t pool_mode,
                         int64_t go_stride_t, int64_t go_stride_b, void *stream);
/* Host-side query, no CUDA call: the digit plan and tiling pb200_tbe_plan_build uses for a request of this
 * shape — passes, per-pass digit width and shift (arrays of 4), bags per tile, tiles per table — and the byte
 * offsets of the sorted keys / values inside the plan buffer.  Everything the host decides about a plan is a
 * function of these by-value arguments (which is what makes building and consuming it free of device
 * read-backs); exposed so that it can be checked without a GPU. */
int pb200_sort_plan_geometry(int64_t n_indices, int32_t num_tables, int64_t batch, int64_t max_table_rows,
                             int32_t *passes, int32_t *bits, int32_t *shifts, int32_t *tile_bags,
                             int32_t *tiles_per_table, int64_t *keys_offset, int64_t *vals_offset);

/* =========================================================================
 * 4. Peer-memory all-to-all (single node, NVLink 5 / NVSwitch)
 * =========================================================================
 * Replaces: backendFunctions.all_to_all_single / all_to_allv
 *   train/comms/pt/pytorch_dist_backend.py:330-357, :262-328
 *   (dist.all_to_all_single(out, in, out_splits, in_splits)),
 *   et_replay/comm/backend/pytorch_dist_backend.py:317-379.
 *
 * A communicator is a table of peer-mapped base pointers: for every rank r a
 * data window (peer_data[r], window_bytes each) and a signal pad
 * (peer_signal[r], >= PB200_A2A_SIGNAL_BYTES, zero-initialised once).  The
 * host layer obtains them from torch.distributed._symmetric_memory (or any
 * CUDA-IPC / VMM mapping) — this library only needs the addresses.
 */
#define PB200_A2A_MAX_RANKS 16
#define PB200_A2A_SIGNAL_BYTES 4096

typedef struct pb200_a2a_comm pb200_a2a_comm; /* opaque */

int pb200_a2a_comm_create(pb200_a2a_comm **comm, int32_t rank, int32_t world,
                          void *const *peer_data, void *const *peer_signal,
                          int64_t window_bytes);
int pb200_a2a_comm_destroy(pb200_a2a_comm *comm);
/* max_ctas: cap on the grid of the push kernel (0 = SM count; < 0 leaves it unchanged);
 * spin_timeout_s: a flag wait that lasts longer (a peer that never arrives) records an error
 * instead of hanging the GPU (<= 0 leaves the ~10 s default). */
int pb200_a2a_comm_config(pb200_a2a_comm *comm, int32_t max_ctas, double spin_timeout_s);
/* Synchronous: *error_out != 0 if any collective on this communicator timed out. */
int pb200_a2a_comm_error(pb200_a2a_comm *comm, int32_t *error_out);

/* Byte-granular all_to_all_single.  `in` is any local device buffer (it need
 * not be in the window).  The result lands in THIS rank's data window at byte
 * offset out_window_off, laid out source-rank-major exactly like c10d:
 *   out = concat_{src r} in_r[ block destined to me ].
 * in_split_bytes[w], out_split_bytes[w]: host arrays of W byte counts
 * (c10d input_split_sizes / output_split_sizes times the element size).
 * Equal splits: pass NULL for both and total_in_bytes (a multiple of W).
 * If `out` is non-NULL the window result is copied there on the stream.
 * Graph-capturable: no host synchronisation inside. */
int pb200_a2a_single(pb200_a2a_comm *comm, const void *in, int64_t total_in_bytes,
                     const int64_t *in_split_bytes, const int64_t *out_split_bytes,
                     int64_t out_window_off, void *out, void *stream);

/* List form — dist.all_to_all(output_tensor_list, input_tensor_list).
 * Replaces: backendFunctions.all_to_all   train/comms/pt/pytorch_dist_backend.py:207-260.
 * in_ptrs[w] / in_bytes[w]: host arrays, the device block sent to rank w.  The block received from
 * source r lands at byte offset out_window_offs[r] of THIS rank's window (out_bytes[r] bytes; an
 * output tensor that lives in the window is written in place by the peer).  out_copy: NULL, or a
 * host array of W device pointers — block r is then copied from the window to out_copy[r] on the
 * stream (entries may be NULL).  No packing pass before, no unpacking pass after. */
int pb200_a2a_list(pb200_a2a_comm *comm, const void *const *in_ptrs, const int64_t *in_bytes,
                   const int64_t *out_window_offs, const int64_t *out_bytes,
                   void *const *out_copy, void *stream);

/* =========================================================================
 * 5. DLRM pooled-embedding exchange fused with the output permute
 * =========================================================================
 * Replaces: All2Allv_Req.forward + All2Allv_Wait.forward + torch.cat(B, dim=1)
 *   train/comms/pt/dlrm.py:86-134, :157-177, :1253 (forward) and
 *   All2Allv_Wait.backward + All2Allv_Req.backward   dlrm.py:180-218, :137-154.
 *
 * forward : local pooled `in` (element (t, n, e) at in[t*in_stride_t +
 *   n*in_stride_n + e], t < T_local, n < N = sum(batch_split)) is written
 *   straight into every destination rank j's window as the final
 *   [lN_j, T_global*E] tensor: rows n in j's batch slice, columns
 *   (table_base + t)*E + e.  No cat before, none after.
 * backward: the transpose — rank j's grad [lN_j, T_global*E] (local buffer) is
 *   scattered to the owners of each table block, landing in the owner's window
 *   as [T_local, N, E] (table-major, what the lookup backward consumes).
 * batch_split[w]  : host, local batch of each rank (a2ai.gNS, dlrm.py:874-876)
 * tables_split[w] : host, tables owned by each rank (n_emb_per_rank)
 */
int pb200_a2a_pooled_fwd(pb200_a2a_comm *comm, const float *in,
                         int64_t in_stride_t, int64_t in_stride_n,
                         int32_t emb_dim, const int64_t *batch_split,
                         const int64_t *tables_split, int64_t out_window_off,
                         void *stream);
int pb200_a2a_pooled_bwd(pb200_a2a_comm *comm, const float *grad /* [lN, T_global*E] */,
                         int32_t emb_dim, const int64_t *batch_split,
                         const int64_t *tables_split, int64_t out_window_off,
                         void *stream);

/* The backward exchange in `parts` pieces: piece `part` moves, for every owner, the gradient columns of its tables
 * [lo, hi) of that piece (tables split contiguously, remainder to the low pieces).  Same window layout as the
 * whole exchange, one epoch per piece: the owner reduces piece g (pb200_tbe_bwd with table_lo / table_hi ... see
 * pb200_tbe_bwd_tables) while piece g + 1 is still on the wire.  parts = 1 is pb200_a2a_pooled_bwd. */
int pb200_a2a_pooled_bwd_part(pb200_a2a_comm *comm, const float *grad, int32_t emb_dim,
                              const int64_t *batch_split, const int64_t *tables_split,
                              int64_t out_window_off, int32_t part, int32_t parts, void *stream);

/* Fused lookup + exchange, ONE kernel per rank: the batched EmbeddingBag forward over this
 * rank's tables for the GLOBAL batch (TBE request: offsets[T_local*N + 1]) whose epilogue stores
 * every pooled row directly into the owning rank's final [lN, T_global*E] tensor in its window —
 * apply_emb + All2Allv_Req/Wait + both torch.cat of dlrm.py:363-388, :86-177, :1253 in one launch,
 * with the NVLink transfer of one bag overlapping the gather of the next.  Same epoch protocol as
 * pb200_a2a_single (the calls can be mixed on one communicator). */
int pb200_tbe_fwd_a2a(pb200_a2a_comm *comm, const float *weights,
                      const int64_t *table_row_offsets, int32_t num_tables_local, int32_t dim,
                      const void *indices, int64_t n_indices, const void *offsets,
                      int32_t idx_type, int32_t pool_mode,
                      const int64_t *batch_split, const int64_t *tables_split,
                      int64_t out_window_off, void *stream);

/* =========================================================================
 * 6. Sparse-input regroup (integer, bit-exact)
 * =========================================================================
 * Replaces: paramDLRM_Net.splitPerTable + lengthsToOffsets
 *   train/comms/pt/dlrm.py:430-504, :245-251.
 *
 * After the lengths/indices all-to-all a rank holds, for its T_local tables,
 *   lengths [W][T_local][b]  and  indices in the same (rank, table, sample) order.
 * Output: per-table lengths regrouped to [T_local][W*b] (global batch in rank
 * order), TBE offsets int64[T_local*W*b + 1] (exclusive cumsum over the
 * table-major concatenation) and indices permuted to table-major order.
 * scratch: pb200_regroup_scratch_bytes() bytes of device memory.
 */
int64_t pb200_regroup_scratch_bytes(int32_t world, int32_t tables_local, int64_t local_batch);
int pb200_regroup_sparse(const int64_t *lengths_in, const int64_t *indices_in, int64_t n_indices,
                         int32_t world, int32_t tables_local, int64_t local_batch,
                         int64_t *lengths_out, int64_t *offsets_out, int64_t *indices_out,
                         void *scratch, int64_t scratch_bytes, void *stream);

/* SparseDataDist as ONE asynchronous call, no host round trip.
 * Replaces: SparseDataDist.forward  train/comms/pt/dlrm.py:744-855 (lengths all_to_all, .item() /
 *   .numpy() syncs for the index counts :801-818, indices all_to_all, splitPerTable).
 *
 * lengths : int64 [T_global * b]  this rank's LOCAL batch, all global tables (table-major)
 * indices : int64 [n_indices_local] in the same order (n_indices_local = host-known tensor size)
 * Steps on `stream`: (1) lengths all-to-all into the window at lengths_window_off
 *   ([W][T_local][b]); (2) per-destination index counts, kept in device memory; (3) indices
 *   all-to-all whose block sizes the push kernel reads from device memory — source r writes into
 *   the fixed slot [indices_window_off + r*slot_elems*8, +slot_elems*8) of the receiver's window
 *   (slot_elems >= the most indices one rank can send to another: T_local_max * b * max bag);
 *   (4) regroup as pb200_regroup_sparse.  A block larger than its slot is truncated and
 *   pb200_a2a_comm_error() reports 2.
 * Outputs (device): lengths_out int64[T_local*W*b], offsets_out int64[T_local*W*b + 1],
 *   indices_out int64 capacity W*slot_elems (valid prefix: offsets_out[last]).
 * scratch: pb200_regroup_scratch_bytes(world, T_local, b) bytes.  Graph-capturable.
 */
int pb200_sparse_data_dist(pb200_a2a_comm *comm, const int64_t *lengths, const int64_t *indices,
                           int64_t n_indices_local, const int64_t *tables_split /* host [W] */,
                           int64_t local_batch, int64_t lengths_window_off,
                           int64_t indices_window_off, int64_t slot_elems,
                           int64_t *lengths_out, int64_t *offsets_out, int64_t *indices_out,
                           void *scratch, int64_t scratch_bytes, void *stream);

/* =========================================================================
 * 7. Host-buffer entry (end-to-end path: H2D + kernels + D2H inside the call)
 * =========================================================================
 * The call a non-torch host makes: indices/offsets/out are HOST pointers
 * (pinned for full PCIe rate), the weight arena stays resident in HBM.
 * Tables are processed in groups so that the H2D of group g+1, the kernel of
 * group g and the D2H of group g-1 overlap on three streams.  Synchronous.
 * out_host receives [B, T*dim] (out_layout 0) or [T, B, dim] (out_layout 1).
 */
typedef struct pb200_host_ctx pb200_host_ctx; /* opaque: staging buffers + streams */
int pb200_host_ctx_create(pb200_host_ctx **ctx, int64_t max_indices_per_group,
                          int64_t max_bags_per_group, int32_t dim);
int pb200_host_ctx_destroy(pb200_host_ctx *ctx);
int pb200_tbe_fwd_host(pb200_host_ctx *ctx, const float *weights_dev,
                       const int64_t *table_row_offsets_dev, const int64_t *table_row_offsets_host,
                       int32_t num_tables, int32_t dim,
                       const int64_t *indices_host, int64_t n_indices,
                       const int64_t *offsets_host, int64_t batch,
                       int32_t pool_mode, float *out_host, int32_t out_layout,
                       int32_t tables_per_group);
/* Training-step form: as above, then (do_bwd != 0) the pooled vectors of each group are used as
 * the incoming gradient and scattered into the arena with scale bwd_scale (= -lr): the D2H of the
 * pooled result overlaps the backward kernel. */
int pb200_tbe_step_host(pb200_host_ctx *ctx, float *weights_dev,
                        const int64_t *table_row_offsets_dev, const int64_t *table_row_offsets_host,
                        int32_t num_tables, int32_t dim,
                        const int64_t *indices_host, int64_t n_indices,
                        const int64_t *offsets_host, int64_t batch,
                        int32_t pool_mode, float *out_host, int32_t out_layout,
                        int32_t tables_per_group, int32_t do_bwd, float bwd_scale);
/* Loss form of the training step: what leaves the device is the step's scalar result, not the pooled
 * vectors — loss_host[t] = sum over the batch and the embedding dimension of table t's pooled vectors,
 * accumulated in double in a fixed order ([num_tables] doubles, host memory).  The pooled vectors stay
 * in HBM, as they do in the reference's GPU loop (train/compute/pt/pytorch_emb.py:48-69 measure_gpu
 * leaves `results` on the device) and in its DLRM step, where they feed the all-to-all
 * (train/comms/pt/dlrm.py:1239-1253).  H2D of indices + offsets, lookup, per-table sum, (do_bwd) the
 * scatter-add, and the D2H of the sums all happen inside the call.  dim must be a multiple of 4. */
int pb200_tbe_step_host_loss(pb200_host_ctx *ctx, float *weights_dev,
                             const int64_t *table_row_offsets_dev, const int64_t *table_row_offsets_host,
                             int32_t num_tables, int32_t dim,
                             const int64_t *indices_host, int64_t n_indices,
                             const int64_t *offsets_host, int64_t batch,
                             int32_t pool_mode, double *loss_host,
                             int32_t tables_per_group, int32_t do_bwd, float bwd_scale);

/* The reduction the loss form uses, on device pointers: sums_dev[i] = sum of the block_elems floats of block i
 * of pooled_dev (n_blocks contiguous blocks), accumulated in double in a fixed order (32 slices per block, an
 * xor-shuffle tree per slice, slices added in order) — run to run identical.  block_elems must be a multiple of
 * 4 and pooled_dev 16 B-aligned.  Asynchronous on `stream`. */
int64_t pb200_pooled_sum_scratch_bytes(int64_t n_blocks);
int pb200_pooled_sum(const float *pooled_dev, int64_t n_blocks, int64_t block_elems, double *sums_dev,
                     void *scratch_dev, int64_t scratch_bytes, void *stream);

/* =========================================================================
 * 8. Synthetic data on the device (benchmark support, not on the parity path)
 * =========================================================================
 * Counter-based generators so that a 100+ GB arena can be filled in place and
 * re-derived on the CPU for spot checks (value = f(seed, element index)).
 * pb200_fill_uniform: dst[i] = lo + (hi-lo) * u(seed, i)     (U(±sqrt(1/n)) init,
 *   pytorch_dist_backend.py:926-928)
 * pb200_fill_zipf_indices: n_bags bags of nnz indices, truncated Zipf(alpha) over
 *   [0, num_rows) by inverse CDF on a device CDF table (pytorch_emb.py:143-146 pmf); dedupe != 0
 *   keeps the indices of a bag distinct, as the reference's per-bag dedupe does (:146-157).
 */
int pb200_fill_uniform(float *dst, int64_t n, float lo, float hi, uint64_t seed, void *stream);
int pb200_fill_zipf_indices(int64_t *dst, int64_t n_bags, int32_t nnz, const double *cdf_dev,
                            int64_t num_rows, int32_t dedupe, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PARAM_B200_H_ */
