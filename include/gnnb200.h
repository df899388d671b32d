/* gnnb200.h — C ABI of libgnnb200.so: the B200 (sm_100a) message-passing hot path.
 *
 * The reference (alonbebchuk/GNN-Pretraining) has no FFI: its seam is the Python nn.Module /
 * function surface over PyTorch-Geometric.  Each entry point below names the reference call
 * site (paths relative to the reference root) or the PyG primitive whose
 * device work it replaces.  The Python binding a maintainer adds is in INTEGRATION.md.
 *
 * Conventions (all functions):
 *   - plain pointers to DEVICE memory, sizes as int64_t, no torch types;
 *   - returns 0 on success, a negative GNNB200_E* code, or a positive cudaError_t;
 *   - never allocates, never synchronises, no mutable global state: re-entrant across streams;
 *   - scratch is caller-provided: calling with workspace == NULL writes the required size to
 *     *workspace_bytes and returns 0 without launching anything;
 *   - row-major matrices with an explicit leading dimension in ELEMENTS;
 *   - reductions have a fixed order: same inputs -> same bits, no float atomics anywhere.
 */
#ifndef GNNB200_H_
#define GNNB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gnnb200_stream_t; /* == cudaStream_t */

#define GNNB200_OK 0
#define GNNB200_EINVAL (-1)     /* bad argument (null pointer, negative size, bad mode)      */
#define GNNB200_ERANGE (-2)     /* N or E does not fit the int32 CSR (>= 2^31)               */
#define GNNB200_EWORKSPACE (-3) /* workspace too small                                        */
#define GNNB200_EUNSUPPORTED (-4) /* shape/alignment not supported by this kernel              */
#define GNNB200_ETMA (-5)       /* TMA tensor-map encode failed (driver entry point / arguments) */

int gnnb200_version(void);
const char* gnnb200_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * CSR / CSC build.  Replaces the COO gather + scatter_add indexing of PyG GINConv.propagate
 * (src/models/gnn.py:41; SURVEY §8a note): bit-exact with
 *     perm   = torch.sort(key_row, stable=True).indices      -> eid
 *     col    = other_row[perm]
 *     rowptr = cat([0, cumsum(bincount(key_row, minlength=N))])
 * edge_index is [2, E] int64 row-major (row 0 = src, row 1 = dst).  by_src = 0 groups by dst
 * (forward aggregation), by_src = 1 groups by src (transposed graph, backward pass).
 * eid may be NULL.  Indices must lie in [0, N).
 * ------------------------------------------------------------------------------------------ */
int gnnb200_csr_build_i64(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                          int by_src, int32_t* rowptr, int32_t* col, int32_t* eid,
                          void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);

/* Segment offsets of a SORTED int64 id vector (PyG Batch.batch -> ptr, App. A.6):
 * ptr[g] = first position with ids[pos] >= g, ptr[num_segments] = n. */
int gnnb200_segment_ptr_i64(const int64_t* ids, int64_t n, int64_t num_segments, int32_t* ptr,
                            gnnb200_stream_t stream);

/* to_undirected()'s coalesce (src/pretrain/tasks.py:108; App. A.4): sort the E columns by
 * row*N+col ascending and drop duplicates.  out is [2, E] (capacity), *out_count (device)
 * receives the number of kept columns; unused tail columns are left untouched. */
int gnnb200_coalesce_i64(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                         int64_t* out, int64_t* out_count, void* workspace, size_t* workspace_bytes,
                         gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-destination aggregation over a CSR (K1/K2/K1^T of SURVEY §2.4; GINConv of
 * src/models/gnn.py:29-37,41; modes of App. A.7):
 *   SUM : out[i] = sum_{e in row i} x[col[e]]                       (edge order, from 0.0f)
 *   MEAN: out[i] = SUM / max(deg_i, 1)
 *   GCN : out[i] = sum_e (dinv[col[e]]*dinv[i]) * x[col[e]] + (dinv[i]*dinv[i]) * x_self[i]
 * and, when self_x != NULL (SUM/MEAN): out[i] += (1 + *eps) * self_x[i]   (eps NULL -> 0).
 * The backward pass is the same call on the by_src CSR with x = grad (transposed gather).
 * feat is the row width; x/self_x/out have leading dimensions ldx/lds/ldo (elements).
 * ------------------------------------------------------------------------------------------ */
#define GNNB200_AGG_SUM 0
#define GNNB200_AGG_MEAN 1
#define GNNB200_AGG_GCN 2
#define GNNB200_AGG_ACCUMULATE 8 /* OR into SUM: start each row's sum from the value already in out (chunked halo passes) */
#define GNNB200_AGG_SKIP_LONG 16 /* OR into SUM: leave rows with more than GNNB200_AGG_LONG_ROW neighbours untouched; the caller covers
                                    them with gnnb200_aggregate_long_rows_f32 (128-bit layouts only, else GNNB200_EUNSUPPORTED) */
#define GNNB200_AGG_LONG_ROW 1024
int gnnb200_aggregate_f32(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                          int64_t num_rows, int64_t feat, int mode, const float* self_x, int64_t lds,
                          const float* eps, const float* dinv, float* out, int64_t ldo,
                          gnnb200_stream_t stream);

/* The rows listed in `rows` (int64 [num_long_rows]; the hubs of a skewed degree distribution) of the same aggregation, one
 * thread block per row: 8 warps sum 8 contiguous slices of the neighbour list in edge order and the partial rows are added
 * in slice order — deterministic, equal to the sequential sum up to fp32 re-association.  mode: SUM [| ACCUMULATE]. */
int gnnb200_aggregate_long_rows_f32(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                                    const int64_t* rows, int64_t num_long_rows, int64_t feat, int mode,
                                    const float* self_x, int64_t lds, const float* eps, float* out, int64_t ldo,
                                    gnnb200_stream_t stream);

/* Deterministic dot product sum(a*b) over n elements -> *out (device).  d(eps) of GINConv. */
int gnnb200_dot_f32(const float* a, const float* b, int64_t n, float* out, void* workspace,
                    size_t* workspace_bytes, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Global graph pooling over contiguous segments (global_mean_pool / global_max_pool,
 * src/models/finetune_model.py:75, src/pretrain/tasks.py:241-246,299,331; App. A.2/A.3).
 * ptr is int32 [num_segments+1].  Empty segments give 0 rows.
 *   fwd: out [num_segments, feat]
 *   bwd: MEAN/SUM dx[r] = g[seg(r)] (/ max(cnt,1));  MAX follows torch's native amax rule:
 *        dx[r,c] = g[s,c] * [x[r,c]==out[s,c]] / (ties + (out[s,c]==0 ? 1 : 0)).
 * ------------------------------------------------------------------------------------------ */
#define GNNB200_POOL_SUM 0
#define GNNB200_POOL_MEAN 1
#define GNNB200_POOL_MAX 2
int gnnb200_segment_pool_fwd_f32(const float* x, int64_t ldx, const int32_t* ptr, int64_t num_rows,
                                 int64_t num_segments, int64_t feat, int mode, float* out, int64_t ldo,
                                 void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);
int gnnb200_segment_pool_bwd_f32(const float* grad_out, int64_t ldg, const float* x, int64_t ldx,
                                 const float* out, int64_t ldo, const int32_t* ptr, int64_t num_rows,
                                 int64_t num_segments, int64_t feat, int mode, float* grad_x, int64_t ldgx,
                                 gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row gather / scatter (K12: NFM mask-token write and target gather,
 * src/models/pretrain_model.py:84-86; h[mask_indices] at src/pretrain/tasks.py:82).
 *   gather : out[i]      = x[idx[i]]
 *   scatter: out[idx[i]] = src[i]  (or the single broadcast row src when broadcast != 0)
 *   gather_bwd: grad_x[r] = sum of grad_out rows i with idx[i] == r, in ascending i (deterministic).
 * ------------------------------------------------------------------------------------------ */
int gnnb200_rows_gather_f32(const float* x, int64_t ldx, const int64_t* idx, int64_t num_idx,
                            int64_t feat, float* out, int64_t ldo, gnnb200_stream_t stream);
int gnnb200_rows_scatter_f32(const float* src, int64_t lds, int broadcast, const int64_t* idx,
                             int64_t num_idx, int64_t feat, float* out, int64_t ldo,
                             gnnb200_stream_t stream);
/* rowptr/eid: CSR over idx built by gnnb200_csr_build_i64 on [idx; arange] with by_src = 1
 * (rows = x rows, eid = positions in grad_out, ascending). */
int gnnb200_rows_gather_bwd_f32(const float* grad_out, int64_t ldg, const int32_t* rowptr,
                                const int32_t* eid, int64_t num_rows, int64_t feat, float* grad_x,
                                int64_t ldgx, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense transforms (K3: nn.Linear of src/models/gnn.py:13,31,34 and src/models/heads.py:41).
 *   C[M,N] = op(A) * op(B) (+ bias[N]) (+ residual[M,N]) (ReLU)        op = identity or transpose
 *   (residual fuses GINLayer's `gin_conv(h) + h`, src/models/gnn.py:41; may be NULL)
 *   col_sum / col_m2 (may be NULL): per-column sum and centred second moment of the written C, i.e. the
 *   batch statistics of the BatchNorm that follows (src/models/gnn.py:15,32,38), computed in the epilogue
 *   on the tensor path (no extra pass over C) and by a second pass on the FFMA path; not with ReLU.
 *   A is [M,K] (transa=0) or [K,M] (transa=1); B is [K,N] (transb=0) or [N,K] (transb=1).
 * precision: F32 = fp32 FFMA (1e-5 class); TF32 = tcgen05.mma kind::tf32 with TMA-fed
 * shared-memory tiles and TMEM fp32 accumulators (2e-2 class); TF32X3 = the same pipeline with every operand
 * split in shared memory into hi + lo tf32 parts and three MMAs per K-step (fp32 class).  TF32 needs a TMA-legal
 * layout (16-byte aligned bases, ld % 4 == 0); otherwise GNNB200_EUNSUPPORTED.
 * ------------------------------------------------------------------------------------------ */
#define GNNB200_GEMM_F32 0
#define GNNB200_GEMM_TF32 1
#define GNNB200_GEMM_AUTO 2 /* TF32 tensor path when the layout allows it, else the fp32 FFMA kernel */
#define GNNB200_GEMM_TF32X3 3 /* error-compensated 3xTF32 on the tensor pipe: fp32-class (1e-5) accuracy */
#define GNNB200_GEMM_AUTO_X3 4 /* TF32X3 when the layout allows it, else the fp32 FFMA kernel */
#define GNNB200_GEMM_AUTO_FWD3 5 /* the default of the modules: nn.Linear FORWARD passes go through gnnb200_linear_x3w_f32 (3xTF32,
                                    weights pre-split), every other GEMM (dX, dW, similarity matrices) is GNNB200_GEMM_AUTO;
                                    passed to gnnb200_gemm_f32 itself it means GNNB200_GEMM_AUTO */
#define GNNB200_EPI_NONE 0
#define GNNB200_EPI_RELU 1
int gnnb200_gemm_f32(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb,
                     float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias,
                     const float* residual, int64_t ldr, int epilogue, int precision, float* col_sum,
                     float* col_m2, void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);

/* nn.Linear forward  Y = X W^T (+bias)(+residual)(ReLU)  in error-compensated 3xTF32 with the weight matrix split ONCE
 * (gnnb200_split_tf32_f32: hi = w with the 13 low mantissa bits cleared, lo = w - hi; redo it after every optimizer
 * step) instead of inside every tile: fp32-class activations from the tf32 tensor pipe.  This is what the forward passes of
 * src/models/gnn.py:13,31,34 and src/models/heads.py:41 run under the default precision: the ReLU masks of the backward
 * pass are decided by these outputs, and plain tf32 there flips enough of them to put the end-to-end gradients outside
 * the 2e-2 class (tests/test_gpu_models.py), while plain tf32 in the backward GEMMs does not.
 * X [M,K] (ldx), W / W_hi / W_lo [N,K] (ldw), Y [M,N].  raw_hi != 0: W itself is read as the hi operand (kind::tf32 ignores
 * the 13 low mantissa bits; W_hi may be NULL).  Layouts the tensor-core kernel cannot take run the fp32 FFMA kernel on W. */
int gnnb200_split_tf32_f32(const float* x, int64_t n, float* hi, float* lo, gnnb200_stream_t stream);
int gnnb200_linear_x3w_f32(const float* X, int64_t ldx, const float* W, const float* W_hi, const float* W_lo,
                           int64_t ldw, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, const float* bias,
                           const float* residual, int64_t ldr, int epilogue, int raw_hi, float* col_sum,
                           float* col_m2, void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);

/* Column statistics over rows (K4 BatchNorm1d batch stats of src/models/gnn.py:15,32,38 and bias
 * gradients): sum[c] = sum_r x[r,c]; when sumsq != NULL also the centred second moment
 * m2[c] = sum_r (x[r,c]-mean_c)^2 (Chan merge of per-block Welford partials). */
int gnnb200_colstats_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* sum,
                         float* m2, void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused BatchNorm1d (+ReLU)(+dropout) over node rows (K4/K5: src/models/gnn.py:19-22,31-32,41-43).
 *   finalize: mean = sum/rows, invstd = rsqrt(m2/rows + eps); running stats (may be NULL) updated
 *             like torch (momentum, unbiased variance).  sum/m2 come from gnnb200_colstats_f32.
 *   fwd     : y = drop(relu((x - mean) * invstd * gamma + beta)); dropout is Bernoulli(1-p) with
 *             1/(1-p) scaling from a Philox4x32-10 stream keyed by (seed, element index).
 *   bwd     : recomputes x_hat / ReLU sign / dropout mask from x (the saved pre-BN activation);
 *             writes grad_x, dgamma, dbeta (two-stage fixed-order column reductions).
 *             training = 0 treats mean/invstd as constants (eval mode).
 *             phase 0 = reduce + apply on one device; phase 1 = reduce only (local dgamma/dbeta);
 *             phase 2 = apply only with caller-provided (all-reduced) dgamma/dbeta and the global
 *             row count rows_total — the node-partitioned multi-GPU path.
 * cols % 4 == 0 and 16-byte aligned rows are required (GNNB200_EUNSUPPORTED otherwise).
 * ------------------------------------------------------------------------------------------ */
int gnnb200_bn_finalize_f32(const float* sum, const float* m2, int64_t rows, int64_t cols, float eps,
                            float momentum, float* running_mean, float* running_var, float* mean,
                            float* invstd, gnnb200_stream_t stream);
/* Node-partitioned batches (BASELINE config 5): moments [parts][3][cols] = every rank's (row count, column sums, centred
 * second moments) after an all-gather; merged in rank order (Chan) and finalised as above in one launch. */
int gnnb200_bn_merge_finalize_f32(const float* moments, int64_t parts, int64_t cols, float eps, float momentum,
                                  float* running_mean, float* running_var, float* mean, float* invstd,
                                  gnnb200_stream_t stream);
int gnnb200_bn_act_fwd_f32(const float* x, int64_t ldx, const float* mean, const float* invstd,
                           const float* gamma, const float* beta, int relu, float drop_p, uint64_t seed,
                           int64_t rows, int64_t cols, float* y, int64_t ldy, gnnb200_stream_t stream);
int gnnb200_bn_act_bwd_f32(const float* grad_y, int64_t ldg, const float* x, int64_t ldx, const float* mean,
                           const float* invstd, const float* gamma, const float* beta, int relu, float drop_p,
                           uint64_t seed, int training, int phase, int64_t rows, int64_t rows_total, int64_t cols,
                           float* grad_x, int64_t ldgx, float* dgamma, float* dbeta, void* workspace,
                           size_t* workspace_bytes, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Link-prediction decoder input (K8, src/models/heads.py:59-65): for each edge (u,v)
 *   feat[e] = [h[u]+h[v], h[u]*h[v], |h[u]-h[v]|]   -> [E, 3*H]
 * and its backward into grad_h (deterministic: edges grouped per node by the caller's CSRs).
 * ------------------------------------------------------------------------------------------ */
int gnnb200_lp_features_f32(const float* h, int64_t ldh, const int64_t* edges, int64_t num_edges,
                            int64_t hidden, float* feat, int64_t ldf, gnnb200_stream_t stream);
/* u_ptr/u_eid: CSR of the edge list grouped by edges[0] (by_src = 1), v_ptr/v_eid grouped by edges[1]
 * (by_src = 0), both from gnnb200_csr_build_i64 with eid requested. */
int gnnb200_lp_features_bwd_f32(const float* h, int64_t ldh, const int64_t* edges, int64_t num_edges,
                                int64_t hidden, const float* grad_feat, int64_t ldf, const int32_t* u_ptr,
                                const int32_t* u_eid, const int32_t* v_ptr, const int32_t* v_eid,
                                int64_t num_nodes, float* grad_h, int64_t ldgh, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * NT-Xent / InfoNCE (K9, src/pretrain/tasks.py:192-213,265-287): z is [2M, D] (already the
 * concatenation of both views, NOT yet normalised).  Computes row-normalised z, the row-wise
 * log-sum-exp of z z^T / T with the diagonal excluded, and loss = sum_i (lse_i - sim_{i,pos(i)}).
 * Never materialises the [2M,2M] matrix.  fwd writes loss (1 float), lse [2M], norm [2M] (= ||z_i||) and
 * zn [2M,D] (dense); bwd writes grad_z [2M,D] given d(loss) (1 float, device).  bwd: D % 16 == 0, D <= 128.
 * ------------------------------------------------------------------------------------------ */
int gnnb200_ntxent_fwd_f32(const float* z, int64_t ldz, int64_t two_m, int64_t dim, float temperature,
                           float* zn, float* lse, float* norm, float* loss, void* workspace,
                           size_t* workspace_bytes, gnnb200_stream_t stream);
int gnnb200_ntxent_bwd_f32(const float* zn, const float* lse, const float* norm, const float* grad_loss,
                           int64_t two_m, int64_t dim, float temperature, float* grad_z, int64_t ldgz,
                           gnnb200_stream_t stream);

/* Tensor-core variant of the same loss for large contrastive batches: the caller forms sim = zn zn^T [2M, 2M] with
 * gnnb200_gemm_f32 (tcgen05), sim_fwd reduces it row-wise (lse, per-row loss, total loss), sim_bwd overwrites it IN PLACE
 * with dL/dsim, and grad_zn = (dsim + dsim^T) zn is two more GEMMs; normalize_rows(_bwd) are F.normalize and its Jacobian. */
int gnnb200_normalize_rows_f32(const float* z, int64_t ldz, int64_t rows, int64_t dim, float* zn, float* norm,
                               gnnb200_stream_t stream);
int gnnb200_normalize_rows_bwd_f32(const float* zn, const float* grad_zn, int64_t ldg, const float* norm, int64_t rows,
                                   int64_t dim, float* grad_z, int64_t ldz, gnnb200_stream_t stream);
int gnnb200_ntxent_sim_fwd_f32(const float* sim, int64_t lds, int64_t two_m, float temperature, float* lse,
                               float* row_loss, float* loss, gnnb200_stream_t stream);
int gnnb200_ntxent_sim_bwd_f32(float* sim, int64_t lds, int64_t two_m, float temperature, const float* lse,
                               const float* grad_loss, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batched negative sampling, deterministic branch (src/pretrain/tasks.py:107-111 -> PyG
 * batched_negative_sampling, App. A.5): for every graph whose sampler would draw arange(population) — all TU-sized graphs at
 * the reference's call site — the negatives are ALL its non-edges in ascending code order, truncated to `quota`.
 *   edge_index int64 [2, E] with columns grouped by graph (to_undirected order); node_ptr int32 [G+1] (Batch.ptr);
 *   edge_ptr int32 [G+1] = column ranges per graph; last_graph -> int64 on the device = graph of the last column (graphs
 *   behind it are skipped, like upstream); quota = num_neg_samples (the batch's edge count at the call site).
 *   count: counts [G] int64 = negatives per graph; *needs_host (int32, caller zeroes it) is raised when some graph takes
 *          upstream's random.sample branch or has more than 724 nodes — the caller then uses the host sampler instead.
 *   write: out int64 [2, total] (total = sum(counts), offsets = exclusive scan of counts), node ids with the batch offsets.
 * Integer work: bit-exact with the oracle.
 * ------------------------------------------------------------------------------------------ */
int gnnb200_negsample_count_i64(const int64_t* edge_index, int64_t num_edges, const int32_t* node_ptr,
                                const int32_t* edge_ptr, int64_t num_graphs, const int64_t* last_graph, int64_t quota,
                                int64_t* counts, int32_t* needs_host, gnnb200_stream_t stream);
int gnnb200_negsample_write_i64(const int64_t* edge_index, int64_t num_edges, const int32_t* node_ptr,
                                const int32_t* edge_ptr, int64_t num_graphs, const int64_t* counts, const int64_t* offsets,
                                int64_t total, int64_t* out, gnnb200_stream_t stream);

/* Host helper, no device work: Python's random.sample(range(n), k) on a copy of the interpreter's Mersenne Twister state
 * (random.getstate()[1] = 624 words + position), advanced in place — the random branch of PyG negative_sampling draws exactly
 * upstream's numbers (src/pretrain/tasks.py:109-111) without interpreter time per draw.  GNNB200_ERANGE for n >= 2^32. */
int gnnb200_host_py_sample_range(uint32_t* mt624, int32_t* pos, int64_t n, int64_t k, int64_t* out);

/* ------------------------------------------------------------------------------------------
 * Head tails and loss sums (K10: src/models/heads.py:16-24,35-50; src/pretrain/tasks.py:84,120,305,336;
 * src/finetune/finetune.py's cross_entropy).  One launch each (a second, fixed-order finish launch beyond 2^19
 * elements), deterministic.  All tensors dense (contiguous) unless a leading dimension is given.
 *   act_dropout_fwd : y = dropout(relu(x)) (relu != 0; Philox stream of gnnb200_bn_act_fwd_f32 keyed by (seed, i / 4))
 *   act_dropout_bwd : grad_x = grad_y / (1 - p) where y > 0, else 0 (the kept-and-positive mask is read off y)
 *   scale           : y = alpha * x (gradient reversal backward: alpha = -lambda)
 *   sqdiff_sum/_bwd : out[0] = sum (a - b)^2 = F.mse_loss(a, b, reduction='sum'); grad_a = 2 grad_loss[0] (a - b)
 *   sigmoid_bce     : probs = sigmoid(logits); loss[0] = F.binary_cross_entropy(probs, labels, reduction='sum') incl. its
 *                     clamp of the logarithms at -100; bwd = torch's BCE backward times sigmoid'
 *   ce_sum          : loss[0] = F.cross_entropy(logits, target, reduction='sum'), lse [rows] saved for the backward;
 *                     bwd: grad_logits = grad_loss[0] (softmax - onehot)
 * ------------------------------------------------------------------------------------------ */
int gnnb200_act_dropout_fwd_f32(const float* x, int64_t n, int relu, float drop_p, uint64_t seed, float* y,
                                gnnb200_stream_t stream);
int gnnb200_act_dropout_bwd_f32(const float* grad_y, const float* y, int64_t n, float drop_p, float* grad_x,
                                gnnb200_stream_t stream);
int gnnb200_scale_f32(const float* x, int64_t n, float alpha, float* y, gnnb200_stream_t stream);
int gnnb200_sqdiff_sum_f32(const float* a, const float* b, int64_t n, float* out, void* workspace,
                           size_t* workspace_bytes, gnnb200_stream_t stream);
int gnnb200_sqdiff_bwd_f32(const float* a, const float* b, const float* grad_loss, int64_t n, float* grad_a,
                           gnnb200_stream_t stream);
int gnnb200_sigmoid_bce_fwd_f32(const float* logits, const float* labels, int64_t n, float* probs, float* loss,
                                void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);
int gnnb200_sigmoid_bce_bwd_f32(const float* probs, const float* labels, const float* grad_loss, int64_t n,
                                float* grad_logits, gnnb200_stream_t stream);
int gnnb200_ce_sum_fwd_f32(const float* logits, int64_t ldz, const int64_t* target, int64_t rows, int64_t cols, float* lse,
                           float* loss, void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream);
int gnnb200_ce_bwd_f32(const float* logits, int64_t ldz, const int64_t* target, const float* lse, const float* grad_loss,
                       int64_t rows, int64_t cols, float* grad_logits, int64_t ldg, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Gradient surgery on a flat buffer (src/pretrain/gradient_surgery.py:41-101).
 * task_grads [T, P] holds each task's gradient of every parameter (zeros where absent); the P
 * parameters are cut into S tensors by seg_offsets [S+1]; present [T, S] says which task produced a
 * gradient for which tensor; order [T] is the shuffled task order and rank_of [T] its inverse.
 * Per tensor, task i is projected against the ORIGINAL gradients of the tasks before it in `order`
 * (skip when either norm is 0; project when the dot product is negative), then out [P] receives, for the
 * tensors present in order[0], the mean over the tasks that have them (has_out [S] marks those).
 * work [T, P] is scratch; counters [T, S, 2] = (conflicts, projections) per task and tensor.
 * ------------------------------------------------------------------------------------------ */
int gnnb200_pcgrad_f32(const float* task_grads, float* work, int64_t num_tasks, int64_t num_params,
                       const int64_t* seg_offsets, int64_t num_segments, const uint8_t* present,
                       const int32_t* order, const int32_t* rank_of, float* out, uint8_t* has_out,
                       int32_t* counters, gnnb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Node-partitioned aggregation over NVLink peer memory (BASELINE config 5; SURVEY §8e — the reference is
 * single-device, src/models/gnn.py:41 is the op being sharded).  One process per GPU; every rank publishes its
 * row shard in a buffer obtained from gnnb200_peer_alloc and maps the other ranks' buffers with
 * gnnb200_peer_open (CUDA IPC handles, exchanged by the caller over its own channel).  The gather kernel then
 * reads neighbour rows directly from the owning GPU (SUM mode with the optional (1+eps) self term, same
 * edge-order accumulation as gnnb200_aggregate_f32), so halo transfer and sums overlap inside one kernel.
 *   peer_x   : DEVICE array of num_peers device pointers (slot s = base of the row buffer of the rank in slot s,
 *              all 16-byte aligned with leading dimension ldx); 1 <= num_peers <= GNNB200_MAX_PEERS
 *   col      : int32 [E_local], encoded (slot << 28) | row_in_that_buffer   (rows < 2^28)
 * The caller orders the kernel after every rank's publish (a barrier on the stream) and must not overwrite a
 * published buffer while a peer may still read it (double buffering + one barrier per pass is enough).
 * gnnb200_peer_alloc / _open / _close / _free are the only entry points that allocate or map memory; they act
 * on the current device and synchronise like cudaMalloc / cudaFree.
 * ------------------------------------------------------------------------------------------ */
#define GNNB200_MAX_PEERS 16
#define GNNB200_PEER_HANDLE_BYTES 64
int gnnb200_aggregate_peer_f32(const float* const* peer_x, int num_peers, int64_t ldx, const int32_t* rowptr,
                               const int32_t* col, int64_t num_rows, int64_t feat, const float* self_x,
                               int64_t lds, const float* eps, float* out, int64_t ldo, gnnb200_stream_t stream);
/* stream-ordered copy of this rank's [rows, feat] block into its published buffer */
int gnnb200_peer_publish_f32(const float* src, int64_t lds, int64_t rows, int64_t feat, float* dst, int64_t ldd,
                             gnnb200_stream_t stream);
/* stream-ordered copy of `count` floats from a (peer-mapped or local) buffer: the copy-engine all-gather of 'peercopy' */
int gnnb200_peer_copy_f32(float* dst, const float* src, int64_t count, gnnb200_stream_t stream);
int gnnb200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle /* [GNNB200_PEER_HANDLE_BYTES] */);
int gnnb200_peer_open(const unsigned char* handle, void** ptr);
int gnnb200_peer_close(void* ptr);
int gnnb200_peer_free(void* ptr);

/* ------------------------------------------------------------------------------------------
 * One call per GIN layer pass (GINLayer.forward of src/models/gnn.py:26-43 and its backward): the launch
 * sequence gather(+self term) -> Linear -> BatchNorm+ReLU -> Linear(+h) -> BatchNorm+ReLU+dropout issued from
 * C++ on the caller's stream.  Every step is one of the entry points above with the arguments documented
 * there; the only difference to calling them one by one is host time (one FFI crossing per layer pass).
 *   fwd: reads h, eps, the parameters and (eval mode, training == 0) mean1..invstd2; writes z, a1, r1, s, out and
 *        (training) mean1..invstd2 + the running buffers (may be NULL).
 *   bwd: training mode only (GNNB200_EUNSUPPORTED otherwise); rowptr/col are the by-SOURCE CSR; reads grad_out and
 *        what fwd saved; writes ds (which becomes dh when need_dh: the transposed gather accumulates into it), dr1,
 *        da1, dz, dw1 [mid, hidden], dw2 [hidden, mid], dgamma1..dbeta2, and deps (NULL = not wanted; needs ldh == hidden).
 *        The bias gradients are identically zero in training mode (a bias feeding BatchNorm) and are not written.
 * All activations are dense row-major (leading dimension = width) except h (ldh).  Workspace query as everywhere.
 * ------------------------------------------------------------------------------------------ */
typedef struct gnnb200_gin_layer {
  int64_t num_rows, hidden, mid;
  const int32_t* rowptr;
  const int32_t* col;
  const float* h;
  int64_t ldh;
  const float* eps;
  const float *w1, *b1, *gamma1, *beta1, *w2, *b2, *gamma2, *beta2;
  float *running_mean1, *running_var1, *running_mean2, *running_var2;
  float *mean1, *invstd1, *mean2, *invstd2;
  float *z, *a1, *r1, *s, *out;
  const float* grad_out;
  float *ds, *dr1, *da1, *dz;
  float *dw1, *dw2, *dgamma1, *dbeta1, *dgamma2, *dbeta2, *deps;
  uint64_t seed;
  float drop_p, momentum1, bn_eps1, momentum2, bn_eps2;
  int training, precision, need_dh;
  /* precision == GNNB200_GEMM_AUTO_FWD3: the two forward GEMMs are gnnb200_linear_x3w_f32 calls on these pre-split weights
   * (gnnb200_split_tf32_f32; w*_hi may be NULL when x3w_raw_hi != 0); the backward GEMMs stay gnnb200_gemm_f32 */
  int x3w_raw_hi;
  const float *w1_hi, *w1_lo, *w2_hi, *w2_lo;
} gnnb200_gin_layer_t;
int gnnb200_gin_layer_fwd_f32(const gnnb200_gin_layer_t* layer, void* workspace, size_t* workspace_bytes,
                              gnnb200_stream_t stream);
int gnnb200_gin_layer_bwd_f32(const gnnb200_gin_layer_t* layer, void* workspace, size_t* workspace_bytes,
                              gnnb200_stream_t stream);
/* Development hook (per thread): between _begin and _end the two composites record the calls they would make as
 * 64-bit words (function id, argument count, arguments) instead of making them; _end returns the word count
 * (-1 = buffer too small).  Function ids: 0 aggregate, 1 gemm, 2 colstats, 3 bn_finalize, 4 bn_act_fwd, 5 bn_act_bwd, 6 dot,
 * 7 linear_x3w. */
int gnnb200_dev_trace_begin(uint64_t* buf, size_t capacity_words);
long long gnnb200_dev_trace_end(void);

#ifdef __cplusplus
}
#endif
#endif /* GNNB200_H_ */
