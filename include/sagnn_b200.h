/*
 * sagnn_b200 -- C ABI of the B200-native SelfGNN interval-graph propagation.
 *
 * Drop-in boundary for ONE path of LIU-YUXI/SA-GNN (paths below are in that repo):
 * the per-time-interval user<->item propagation of model.py:118-134 (forward) and
 * its TF1-autodiff backward (model.py:250), fed by the adjacency lists of
 * DataHandler.transToLsts / transpose (DataHandler.py:9-11,47-69; call sites
 * model.py:227-237).  The reference has no FFI of its own (pure Python on TF 1.14);
 * these entry points are what a binding for that path would call -- see
 * INTEGRATION.md for the ctypes stub and the model.py patch.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.
 *   - every function returns a sagnn_status (0 = OK); sagnn_last_error() returns a
 *     thread-local message for the last non-zero status.  Nothing throws.
 *   - "_dev" pointers are device memory on the plan's device, "_host" host memory.
 *   - the caller owns every tensor and the workspace; the library owns the plan.
 *   - propagate_* enqueue work on the given stream and never synchronise or
 *     allocate, so they can be captured into a CUDA graph.
 *   - embeddings are fp32, row-major, contiguous: user tables [T, U, d], item
 *     tables [T, I, d]; d in {32, 64, 128, 256}.
 *   - there is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef SAGNN_B200_H_
#define SAGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sagnn_plan sagnn_plan;
typedef void* sagnn_stream_t; /* a cudaStream_t (NULL = legacy default stream) */

typedef enum sagnn_status {
  SAGNN_OK = 0,
  SAGNN_INVALID_ARG = 1,
  SAGNN_UNSORTED_INPUT = 2,      /* segment ids not non-decreasing (TF SegmentSum rejects these too) */
  SAGNN_CUDA_ERROR = 3,
  SAGNN_WORKSPACE_TOO_SMALL = 4,
  SAGNN_NOT_FINALIZED = 5,
  SAGNN_OUT_OF_RANGE = 6         /* an edge id outside [0,U) x [0,I) */
} sagnn_status;

/* side of an interval graph: the A_k CSR (rows = users, model.py:122 'user' call) or
 * the A_k^T CSR (rows = items, model.py:123 'item' call). */
enum { SAGNN_SIDE_USER = 0, SAGNN_SIDE_ITEM = 1 };

/* edge-weight modes for sagnn_plan_finalize */
enum {
  SAGNN_WEIGHTS_NONE = 0,     /* reference-exact: binary structure, stored values ignored (model.py:84-86) */
  SAGNN_WEIGHTS_LIGHTGCN = 1, /* w = rowD*colD, D = 1/(sqrt(deg+1e-8)+1e-8) on structural degrees
                                 (the formula of DataHandler.py:54-55, which the reference then truncates away) */
  SAGNN_WEIGHTS_CUSTOM = 2    /* per-edge fp32 weights passed to sagnn_plan_set_interval */
};

const char* sagnn_last_error(void);
const char* sagnn_version(void);

/* ---- plan: device-side CSR (A_k) + CSC (= CSR of A_k^T) for T interval graphs ----------
 * Replaces the 2T transToLsts() calls + tf.sparse.SparseTensor wrapping of
 * model.py:227-237 and the transpose() of DataHandler.py:9-11. */

/* Producer of the interval graphs from raw events (preprocess_to_trnmat.ipynb cell 7, `trans_sub`):
 * users_dev / items_dev int32 [n], times_dev int64 [n] in the notebook's visiting order (user
 * ascending, then the user's items, then the pair's timestamps); minn / maxx as `trans` (cell 13)
 * leaves them.  Event e goes to interval int((t - minn) / ((maxx - minn) / T)) clamped to T-1; per
 * (interval, user, item) the FIRST event is kept.  Writes the adjacency lists of all intervals back
 * to back, interval 0 first, each row-major sorted -- rows / cols / stored values (timestamps, intc) --
 * into row_out / col_out / val_out (capacity n) and the T list lengths into nnz_host: interval k is
 * the slice [sum_{j<k} nnz_j, +nnz_k), which is what sagnn_plan_set_interval takes.  An empty
 * interval has nnz 0 (pass the reference's (0,0) fallback edge to the plan, DataHandler.py:66-68).
 * Synchronises the stream.  (timeMat, which the model never reads, is not produced.) */
int sagnn_bucket_events(const int32_t* users_dev, const int32_t* items_dev, const int64_t* times_dev, int64_t n,
                        int U, int I, int T, int64_t minn, int64_t maxx, int32_t* row_out_dev, int32_t* col_out_dev,
                        int32_t* val_out_dev, int64_t* nnz_host, sagnn_stream_t stream);

/* nnz_host[T]: edge count of every interval (an empty interval must be passed as the
 * reference's single fallback edge (0,0), DataHandler.py:66-68 -- the Python shim does).
 * Size limits (checked, SAGNN_INVALID_ARG otherwise; all far above BASELINE's largest config, 10 M x 2 M rows,
 * 1 B edges over 8 intervals, which is split over 8 plans anyway):
 *   1 <= nnz[k] < 2^31                 per interval (edge ids inside a segment are 32-bit);
 *   T * (U + I) < 2^31                 global row ids are 32-bit (checked here);
 *   2 * sum_k nnz[k] < 2^32            edge entries of both orientations of all intervals (checked by finalize);
 *   packed task stream < 2^36 bytes    its directory addresses 16-byte units with 32 bits (checked by finalize);
 *   device SM count <= 160             size of the by-value CTA table of the packet-stream kernel (checked at launch);
 *   latdim d in {32, 64, 128, 256}     at every propagate / message_propagate call. */
int sagnn_plan_create(int T, int U, int I, const int64_t* nnz_host, sagnn_plan** out);

/* Adjacency list of interval k in the order transToLsts emits it (row-major COO of the
 * canonical CSR): row_dev/col_dev int32 [nnz] (user id, item id).  val_dev (nullable)
 * are the stored int32 values (timestamps) -- only used for the value-sum degrees and
 * sagnn_plan_norm_data.  w_dev (nullable) are custom fp32 edge weights (same order).
 * Builds the A_k CSR, the stable transposed CSR and structural degrees on the device.
 * Returns SAGNN_UNSORTED_INPUT if rows are not non-decreasing. Synchronises the stream. */
int sagnn_plan_set_interval(sagnn_plan* plan, int k, const int32_t* row_dev, const int32_t* col_dev,
                            const int32_t* val_dev, const float* w_dev, int64_t nnz,
                            sagnn_stream_t stream);

/* Optional, before finalize -- ROW SHARDING of the graphs over several GPUs (second way):
 * this plan only schedules user rows [u_begin,u_end) of every A_k and item rows [i_begin,i_end) of
 * every A_k^T (it still holds the complete adjacency, since the rows it owns gather from every row of
 * the other side).  propagate_* then read whole tables and write only the owned rows; between layers
 * the caller all-gathers the freshly written table (sagnn_propagate_fwd_layers / _bwd_levels /
 * sagnn_workspace_table below). */
int sagnn_plan_set_row_block(sagnn_plan* plan, int u_begin, int u_end, int i_begin, int i_end);

/* Optional, before finalize: the latdim the plan will mostly be run with (default 64).  It picks the
 * schedule the plan carries: the packed task stream of the packet-stream kernel below 128, the task records of
 * the round-1 kernel from 128 on (measured faster there); every latdim in {32, 64, 128, 256} works with either. */
int sagnn_plan_set_latdim_hint(sagnn_plan* plan, int d);

/* Optional, before finalize: stage the n highest-degree rows of every source table in shared memory (TMA bulk
 * copies at kernel start; every task's edge codes list its hot edges first, so the staged rows are read in a
 * loop of their own without a per-edge test).  Default 0: on B200 an L1 hit costs the same L1TEX cycles as a
 * shared-memory read and the staged copy takes the L1's capacity, so the measured step is slower with it
 * (DESIGN.md section 4); kept for tables whose hot set does not stay in L1.  n is capped by what fits next to the
 * packet rings at the hinted latdim (sagnn_plan_stats reports the number used); plans hinted latdim >= 128 ignore
 * it.  The hot-first order changes the summation order inside a row, so results differ in the last bits between
 * plans with different hot sets (still deterministic, still within 1e-5 of the reference).
 * (north_star "TMA/shared-memory staging of hot item rows"; the reference has no counterpart.) */
int sagnn_plan_set_hot_rows(sagnn_plan* plan, int n);

/* Row pointers over all intervals, optional edge weights, degree-binned schedule.
 * Must be called once after every interval has been set. */
int sagnn_plan_finalize(sagnn_plan* plan, int weight_mode, sagnn_stream_t stream);

/* Parity hooks (bit-exact against transToLsts / transpose):
 * indptr_dev int32 [R+1], indices_dev int32 [nnz_k] of the CSR of side `side`. */
int sagnn_plan_get_csr(const sagnn_plan* plan, int k, int side, int32_t* indptr_dev,
                       int32_t* indices_dev, sagnn_stream_t stream);
/* deg_dev int32 [R] structural degrees; valsum_dev (nullable) int64 [R] sums of stored
 * values (np.sum(mat, axis=...) of DataHandler.py:54-55; needs val_dev at set_interval). */
int sagnn_plan_get_degrees(const sagnn_plan* plan, int k, int side, int32_t* deg_dev,
                           int64_t* valsum_dev, sagnn_stream_t stream);
/* data_dev int32 [nnz_k]: the reference's "normalised" values, i.e.
 * (int32)((double)val * rowD[row] * colD[col]) of DataHandler.py:56-59 (all zeros on
 * real data).  Needs val_dev at set_interval. */
int sagnn_plan_norm_data(const sagnn_plan* plan, int k, int side, int32_t* data_dev,
                         sagnn_stream_t stream);
/* w_dev fp32 [nnz_k]: the edge weights in use for side `side` (CSR order). */
int sagnn_plan_get_weights(const sagnn_plan* plan, int k, int side, float* w_dev,
                           sagnn_stream_t stream);

/* schedule statistics: out[0]=rows, out[1]=short rows, out[2]=long rows, out[3]=chunks,
 * out[4]=max degree, out[5]=sum of edges over both sides, out[6]=hot slots per table in use, out[7]=SMs */
int sagnn_plan_stats(const sagnn_plan* plan, int64_t* out8);

/* Load balance.  Persistent CTAs are dealt to the 2T segments (seg = 2k + side) from a static cost
 * model; sagnn_plan_rebalance re-deals them from measured per-segment work (e.g. mean CTA time x
 * CTAs of a traced launch, see sagnn_debug_trace) -- the device tables are updated in place, so
 * captured CUDA graphs stay valid.  sagnn_plan_get_split returns the current CTAs per segment. */
int sagnn_plan_rebalance(sagnn_plan* plan, const double* seg_work /* [2T] */, sagnn_stream_t stream);
int sagnn_plan_get_split(const sagnn_plan* plan, int* ctas_per_segment /* [2T] */);

/* Diagnostics: when trace_dev != NULL every following layer launch (up to capacity_launches)
 * writes, per persistent CTA, 4 x uint64 {segment, t_start, t_hot_rows_staged, t_end} in ns
 * (%globaltimer) at trace_dev + launch * SMs * 4.  Pass NULL to switch tracing off. */
int sagnn_debug_trace(sagnn_plan* plan, uint64_t* trace_dev, int capacity_launches);

int sagnn_plan_destroy(sagnn_plan* plan);

/* ---- propagation ------------------------------------------------------------------------
 * Replaces the T x L x 2 messagePropagate() calls, residual adds and add_n of
 * model.py:118-129 (forward) and their autodiff (backward). */

/* Bytes of: forward scratch, saved activation-sign masks (forward output consumed by
 * backward), backward scratch. */
int sagnn_workspace_bytes(const sagnn_plan* plan, int n_layers, int d, size_t* fwd_bytes,
                          size_t* mask_bytes, size_t* bwd_bytes);

/* user_out[T,U,d], item_out[T,I,d] = sum_l E^l with E0^{l+1} = E0^l + lrelu(A E1^l),
 * E1^{l+1} = E1^l + lrelu(A^T E0^l), lrelu(x) = max(leaky*x, x).
 * masks_dev (mask_bytes, nullable when no backward will follow) receives the sign bits. */
int sagnn_propagate_fwd(const sagnn_plan* plan, const float* u_embed_dev, const float* i_embed_dev,
                        float* user_out_dev, float* item_out_dev, int n_layers, int d, float leaky,
                        void* masks_dev, void* workspace_dev, size_t workspace_bytes,
                        sagnn_stream_t stream);

/* d_u_embed[T,U,d], d_i_embed[T,I,d] from dense upstream grads g_user[T,U,d], g_item[T,I,d]
 * and the masks written by the matching forward. */
int sagnn_propagate_bwd(const sagnn_plan* plan, const float* g_user_dev, const float* g_item_dev,
                        float* d_u_embed_dev, float* d_i_embed_dev, int n_layers, int d, float leaky,
                        const void* masks_dev, void* workspace_dev, size_t workspace_bytes,
                        sagnn_stream_t stream);

/* Layout flags of the _ex entry points.  SAGNN_LAYOUT_RTD: the tensors handed to / taken from the
 * interval-fusion stage use the transposed layout of model.py:133-134 (tf.transpose(.., [1,0,2])):
 * user_out [U,T,d], item_out [I,T,d] in the forward, g_user [U,T,d], g_item [I,T,d] in the
 * backward.  The embedding tables and their gradients stay [T,U,d] / [T,I,d] (model.py:108-109).
 * The transpose is fused into the epilogue (row stride T*d), so the two full-tensor copies of
 * model.py:133-134 (and of their autodiff) disappear; results are bitwise those of the default
 * layout, transposed. */
enum { SAGNN_LAYOUT_TRD = 0, SAGNN_LAYOUT_RTD = 1 };
int sagnn_propagate_fwd_ex(const sagnn_plan* plan, const float* u_embed_dev, const float* i_embed_dev,
                           float* user_out_dev, float* item_out_dev, int n_layers, int d, float leaky,
                           void* masks_dev, void* workspace_dev, size_t workspace_bytes, unsigned flags,
                           sagnn_stream_t stream);
int sagnn_propagate_bwd_ex(const sagnn_plan* plan, const float* g_user_dev, const float* g_item_dev,
                           float* d_u_embed_dev, float* d_i_embed_dev, int n_layers, int d, float leaky,
                           const void* masks_dev, void* workspace_dev, size_t workspace_bytes,
                           unsigned flags, sagnn_stream_t stream);

/* Forward with the hand-off to a ROW-SHARDED consumer fused into the epilogue (interval sharding,
 * `world` ranks, one process per GPU): the last layer writes row r of its layer sums not to
 * user_out / item_out but straight into the receive buffer of rank r / blk (blk = ceil(rows / world)),
 * at [rank][r % blk][k][:] of that rank's fp32 buffer [world, blk, T, d] -- peer-memory stores over
 * NVLink / NVSwitch (buffers mapped on this device, e.g. CUDA IPC or torch symmetric memory; the
 * tables user_recv_host[world] / item_recv_host[world] are HOST arrays of those device pointers, this
 * rank's own buffer included).  No collective kernel, no pack copy: the transfer overlaps the math row
 * by row.  The caller synchronises the ranks (any barrier after the call's stream work) before
 * consumers read.  user_out / item_out ([U,T,d] / [I,T,d]) only hold the partial layer sums of
 * n_layers >= 3.  Replaces model.py:131-134 + the exchange a data-parallel consumer needs. */
int sagnn_propagate_fwd_scatter(const sagnn_plan* plan, const float* u_embed_dev, const float* i_embed_dev,
                                float* user_out_dev, float* item_out_dev, int n_layers, int d, float leaky,
                                void* masks_dev, void* workspace_dev, size_t workspace_bytes, int world,
                                int rank, const void* const* user_recv_host, const void* const* item_recv_host,
                                sagnn_stream_t stream);

/* Row-sharded execution (plans with sagnn_plan_set_row_block): the L-layer forward / backward one
 * stage at a time, so that the caller can all-gather the table each stage wrote (owned rows only)
 * before the next stage gathers from it.  Same tensors, masks and workspace in every call of a step.
 *   forward : layers [l_begin, l_end) of n_layers; after layer l < n_layers-1 exchange table
 *             sagnn_workspace_table(which=0, index=l)  (E^{l+1}); u_embed / i_embed must be complete.
 *   backward: phases [ph_begin, ph_end) of n_layers+1; phase 0 = sigma'(Z^{L-1}) (.) G of the owned rows
 *             (g_user / g_item must be complete), phase j >= 1 = level kernel j-1; after phase
 *             j < n_layers exchange table sagnn_workspace_table(which=1, index=j).
 * user_out / item_out / d_u_embed / d_i_embed receive the owned rows only. */
int sagnn_propagate_fwd_layers(const sagnn_plan* plan, int l_begin, int l_end, const float* u_embed_dev,
                               const float* i_embed_dev, float* user_out_dev, float* item_out_dev,
                               int n_layers, int d, float leaky, void* masks_dev, void* workspace_dev,
                               size_t workspace_bytes, sagnn_stream_t stream);
int sagnn_propagate_bwd_levels(const sagnn_plan* plan, int ph_begin, int ph_end, const float* g_user_dev,
                               const float* g_item_dev, float* d_u_embed_dev, float* d_i_embed_dev,
                               int n_layers, int d, float leaky, const void* masks_dev, void* workspace_dev,
                               size_t workspace_bytes, sagnn_stream_t stream);
/* Byte offset inside the workspace of the table to exchange (layout [T,U,d] then [T,I,d], fp32):
 * which=0: written by forward layer `index` (0 <= index < n_layers-1);
 * which=1: the gather source of backward level `index`, written by backward phase `index`
 *          (0 <= index < n_layers). */
int sagnn_workspace_table(const sagnn_plan* plan, int n_layers, int d, int which, int index,
                          size_t* offset_bytes, size_t* user_bytes, size_t* item_bytes);

/* The same restricted to interval k (rows of the other intervals are not touched): all SMs work on
 * that interval's two CSRs, so a caller can pipeline per-interval copies with compute.  Tensors,
 * masks and workspace are the full [T, ...] buffers; calls must be stream-ordered. */
int sagnn_propagate_fwd_interval(const sagnn_plan* plan, int k, const float* u_embed_dev,
                                 const float* i_embed_dev, float* user_out_dev, float* item_out_dev,
                                 int n_layers, int d, float leaky, void* masks_dev, void* workspace_dev,
                                 size_t workspace_bytes, sagnn_stream_t stream);
int sagnn_propagate_bwd_interval(const sagnn_plan* plan, int k, const float* g_user_dev,
                                 const float* g_item_dev, float* d_u_embed_dev, float* d_i_embed_dev,
                                 int n_layers, int d, float leaky, const void* masks_dev,
                                 void* workspace_dev, size_t workspace_bytes, sagnn_stream_t stream);

/* One messagePropagate() call (model.py:80-92): out[R,d] = lrelu(B_k,side * src[C,d]).
 * workspace: forward scratch size. */
int sagnn_message_propagate(const sagnn_plan* plan, int k, int side, const float* src_dev,
                            float* out_dev, int d, float leaky, void* workspace_dev,
                            size_t workspace_bytes, sagnn_stream_t stream);

/* ---- sampled pair scores over the outputs ------------------------------------
 * The second consumer of user_vector / item_vector and the producer of the sparse part of their
 * upstream gradient:  scores[s] = sum_c act(u_rows[uids[s], c] * i_rows[iids[s], c]),
 * activation 1 = lrelu (model.py:196-198, the SSL scores preds_one of interval k), 0 = none
 * (model.py:171-173, the plain prediction dot product).  Replaces two tf.nn.embedding_lookup gathers,
 * Mul, Maximum and reduce_sum.  A table is (base pointer, row stride in floats): interval k of a
 * [T,R,d] tensor = (t + k*R*d, d), of a [R,T,d] tensor = (t + k*d, T*d).  ids are int32 device arrays
 * and are not range-checked (like embedding_lookup on a GPU). */
int sagnn_pair_scores_fwd(const float* u_rows_dev, int64_t u_stride, const float* i_rows_dev, int64_t i_stride,
                          const int32_t* uids_dev, const int32_t* iids_dev, int64_t n, int d, int activation,
                          float leaky, float* scores_dev, sagnn_stream_t stream);
/* Backward: ADDS g_scores[s] * act'(x) * other_row into d_u_rows[uids[s]] / d_i_rows[iids[s]] (either
 * may be NULL) -- i.e. accumulates the sparse gradient straight into the dense upstream tables that
 * sagnn_propagate_bwd takes (TF: IndexedSlices -> unsorted_segment_sum -> AddN).  Uses float atomics
 * (samples repeat rows), so the summation order, not the set of terms, may vary between runs. */
int sagnn_pair_scores_bwd(const float* u_rows_dev, int64_t u_stride, const float* i_rows_dev, int64_t i_stride,
                          const int32_t* uids_dev, const int32_t* iids_dev, int64_t n, int d, int activation,
                          float leaky, const float* g_scores_dev, float* d_u_rows_dev, int64_t du_stride,
                          float* d_i_rows_dev, int64_t di_stride, sagnn_stream_t stream);

/* The same accumulation WITHOUT atomics: the samples are radix-sorted by the row they scatter into (stable, so
 * sample order inside a row is kept), the warp at the head of each run adds the run's terms in that order and
 * updates the gradient row once -- bit-identical results run to run.  ws: device scratch of
 * sagnn_pair_scores_bwd_ws_bytes(n) bytes.  Gradient tables 16-byte aligned, strides multiples of 4.
 * Negative or out-of-range ids are not checked (as above). */
int sagnn_pair_scores_bwd_ws_bytes(int64_t n, size_t* bytes);
int sagnn_pair_scores_bwd_det(const float* u_rows_dev, int64_t u_stride, const float* i_rows_dev, int64_t i_stride,
                              const int32_t* uids_dev, const int32_t* iids_dev, int64_t n, int d, int activation,
                              float leaky, const float* g_scores_dev, float* d_u_rows_dev, int64_t du_stride,
                              float* d_i_rows_dev, int64_t di_stride, void* ws_dev, size_t ws_bytes,
                              sagnn_stream_t stream);

/* ---- device-side sampleSslBatch (model.py:304-339) ---------------------------
 * For interval k and the batch users bat_ids_dev int32 [batch]: posset(u) = the items of user u in
 * A_k (read from the plan's CSR instead of densifying subMat[k][batIds].toarray()), s = min(ssl_num,
 * |posset| / 2), all = 2*s uniform draws with replacement from posset; emits, users in batch order,
 * i_locs[cur] = all[j], i_locs[cur+1] = all[s+j], u_locs[cur] = u_locs[cur+1] = u,
 * u_locs_seq[cur] = u_locs_seq[cur+1] = batch position, cur += 2 (j < s) -- the reference's
 * suids[k] / siids[k] / suLocs_seq[k] feed.  Output capacity batch*2*ssl_num int32 each; the number
 * of entries written goes to *n_out_host.  Counter-based generator keyed by (seed, k, batch
 * position, draw): same seed, same samples; the stream is not numpy's.  Synchronises the stream. */
int sagnn_sample_ssl_batch(const sagnn_plan* plan, int k, const int32_t* bat_ids_dev, int batch, int ssl_num,
                           uint64_t seed, int32_t* u_locs_dev, int32_t* i_locs_dev, int32_t* u_locs_seq_dev,
                           int64_t* n_out_host, sagnn_stream_t stream);

/* The same for EVERY interval of the plan in one call (what a training step needs: model.py:313 loops k over
 * graphNum): outputs are [T, cap] int32 with cap = batch*2*ssl_num (interval k's entries start at k*cap),
 * n_out_host int64 [T].  Same draws as T calls of sagnn_sample_ssl_batch with the same seed; two kernels, one scan
 * and one stream synchronisation per step, scratch kept in the plan (one sampler call at a time per plan). */
int sagnn_sample_ssl_batch_all(const sagnn_plan* plan, const int32_t* bat_ids_dev, int batch, int ssl_num,
                               uint64_t seed, int32_t* u_locs_dev, int32_t* i_locs_dev, int32_t* u_locs_seq_dev,
                               int64_t* n_out_host, sagnn_stream_t stream);

/* ---- device-side sampleTrainBatch + negSamp (model.py:252-302, DataHandler.py:28-41) ----------
 * seq_ptr_dev int64 [U+1] / seq_items_dev int32: handler.sequence as CSR (a user's interactions in time order);
 * tst_int_dev int32 [U] (held-out item, -1 = None; may be NULL); bat_ids_dev int32 [batch].  Per batch position b,
 * u = bat_ids[b], seq = sequence[u], posset = seq[:-1], sampNum = min(train_sample_num, len(posset)) (0: nothing
 * emitted), choose = randint(1, max(min(pred_num + 1, len(posset) - 3), 1)): the positive posset[-choose], sampNum
 * times, and sampNum negatives = uniform items without a training interaction of u in ANY interval of the plan (the
 * reference tests the dense row of trnMat; here a binary search in the plan's CSR rows) and different from seq[-1]
 * and tst_int[u].  Outputs: u_locs / i_locs / u_locs_seq int32 (capacity 2*batch*train_sample_num): the positives of
 * all users in batch order, then their negatives in the same order (*n_out_host = entries written);
 * sequence int32 / mask float [batch_pad, pos_length]: the last pos_length items of posset[:-choose], right-aligned,
 * mask 1 on them, rows >= batch zero (the reference pads to args.batch); choose_out int32 [batch] (nullable).
 * Counter-based generator keyed by (seed, batch position, draw): same seed, same samples; not numpy's stream.
 * Synchronises the stream. */
int sagnn_sample_train_batch(const sagnn_plan* plan, const int64_t* seq_ptr_dev, const int32_t* seq_items_dev,
                             const int32_t* tst_int_dev, const int32_t* bat_ids_dev, int batch, int batch_pad,
                             int train_sample_num, int pred_num, int pos_length, uint64_t seed, int32_t* u_locs_dev,
                             int32_t* i_locs_dev, int32_t* u_locs_seq_dev, int32_t* sequence_dev, float* mask_dev,
                             int32_t* choose_out_dev, int64_t* n_out_host, sagnn_stream_t stream);

/* ---- stream-exact samplers, parity mode (model.py:252-339, DataHandler.py:28-41, main.py:21-22) ----
 * The reference draws its samples from numpy's global RandomState and from CPython's `random` module, both
 * MT19937, both seeded in main.py.  These HOST functions restate the two libraries' bounded-integer algorithms over
 * caller-owned generator states, so that for the same seeds (or a state taken over from np.random.get_state() /
 * random.getstate()) they emit the reference's samples bit for bit and leave the streams where the reference would.
 * They read the caller's host CSR arrays (scipy's indptr / indices, canonical format: sorted, no duplicates) instead
 * of densifying rows; no GPU is involved.  The device samplers above are the throughput mode.
 *
 * sagnn_mt19937: key + position, the layout of np.random.get_state()[1:3] and of random.getstate()[1]. */
typedef struct sagnn_mt19937 { uint32_t key[624]; int32_t pos; } sagnn_mt19937;
void sagnn_mt19937_seed_numpy(sagnn_mt19937* state, uint32_t seed);    /* np.random.seed(seed)   */
void sagnn_mt19937_seed_python(sagnn_mt19937* state, uint64_t seed);   /* random.seed(seed), a non-negative int */
uint32_t sagnn_mt19937_next32(sagnn_mt19937* state);
/* out[i] = np.random.randint(low, high) (n draws; np.random.choice(m) == randint(0, m)) */
int sagnn_np_randint(sagnn_mt19937* np_state, int64_t low, int64_t high, int64_t n, int64_t* out);
/* out = np.random.permutation(n)   (model.py:342, the epoch's user order) */
int sagnn_np_permutation(sagnn_mt19937* np_state, int64_t n, int64_t* out);
/* *out = random.randint(a, b) */
int sagnn_py_randint(sagnn_mt19937* py_state, int64_t a, int64_t b, int64_t* out);

/* Recommender.sampleSslBatch(batIds, handler.subMat) (model.py:304-339).  indptr[k] / indices[k]: interval k's CSR
 * (int32, [n_user+1] / [nnz_k]); nonzero (nullable, per-interval entries nullable): 1 where the stored value != 0
 * (the reference tests `temLabel != 0`).  Outputs [T, cap] int32 with cap = batch*2*ssl_num (interval k starts at
 * k*cap), n_out int64 [T]; entries (pos, neg) interleaved exactly like the reference's lists. */
int sagnn_np_sample_ssl_batch(sagnn_mt19937* np_state, int T, const int32_t* const* indptr,
                              const int32_t* const* indices, const uint8_t* const* nonzero, const int32_t* bat_ids,
                              int batch, int ssl_num, int n_user, int n_item, int32_t* u_locs, int32_t* i_locs,
                              int32_t* u_locs_seq, int64_t* n_out);

/* Recommender.sampleTrainBatch(batIds, handler.trnMat, ...) + negSamp (model.py:252-302, DataHandler.py:28-41).
 * seq_ptr int64 [n_user+1] / seq_items int32: handler.sequence as CSR; tst_int int32 [n_user], -1 = None (nullable);
 * label_*: the CSR of labelMat (handler.trnMat) with the same `nonzero` convention.  `choose` comes from py_state
 * (random.randint), the negatives from np_state (np.random.choice).  Outputs: u_locs / i_locs / u_locs_seq int32
 * [2*batch*train_sample_num] (positives, then negatives; *n_out entries), sequence int64 / mask double
 * [batch_pad, pos_length] (the reference's np.zeros(..., dtype=int) / np.zeros(...) rows), choose_out int32 [batch]
 * (nullable).  A user with fewer than 3 interactions makes the reference raise (model.py:293): SAGNN_INVALID_ARG. */
int sagnn_np_sample_train_batch(sagnn_mt19937* np_state, sagnn_mt19937* py_state, const int64_t* seq_ptr,
                                const int32_t* seq_items, const int32_t* tst_int, const int32_t* label_indptr,
                                const int32_t* label_indices, const uint8_t* label_nonzero, const int32_t* bat_ids,
                                int batch, int batch_pad, int train_sample_num, int pred_num, int pos_length,
                                int n_user, int n_item, int32_t* u_locs, int32_t* i_locs, int32_t* u_locs_seq,
                                int64_t* sequence, double* mask, int32_t* choose_out, int64_t* n_out);

/* Host-buffer entry point (what a non-torch caller binds): copies the embeddings (and,
 * when g_*_host != NULL, the upstream gradients) to the device, runs forward (+ backward),
 * copies the results back and synchronises.  Device buffers are cached inside the plan.
 * Pass pinned host memory for full PCIe bandwidth. */
int sagnn_propagate_host(sagnn_plan* plan, const float* u_embed_host, const float* i_embed_host,
                         const float* g_user_host, const float* g_item_host, float* user_out_host,
                         float* item_out_host, float* d_u_embed_host, float* d_i_embed_host,
                         int n_layers, int d, float leaky);

/* The same, split for callers whose autodiff runs forward and backward at different times
 * (a TF1 py_func pair): forward keeps the sign masks inside the plan when
 * keep_masks != 0; backward uses them and fails with SAGNN_INVALID_ARG if there are none. */
int sagnn_host_forward(sagnn_plan* plan, const float* u_embed_host, const float* i_embed_host,
                       float* user_out_host, float* item_out_host, int n_layers, int d, float leaky,
                       int keep_masks);
int sagnn_host_backward(sagnn_plan* plan, const float* g_user_host, const float* g_item_host,
                        float* d_u_embed_host, float* d_i_embed_host, int n_layers, int d, float leaky);

#ifdef __cplusplus
}
#endif
#endif /* SAGNN_B200_H_ */
