/*
 * seesaw_b200 — C ABI of the B200-native SeeSaw vector-search hot path.
 *
 * This is the drop-in boundary.  The reference (orm011/seesaw 1.3.0) has no FFI: its hot path is
 * numpy/pandas inside three Python plug points.  Every entry point below names
 * the reference code it replaces (paths relative to /root/reference/seesaw); INTEGRATION.md
 * shows the ctypes stub a reference maintainer would add at each plug point.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * non-zero ssw_status otherwise, with a message available from ssw_last_error() (thread local).
 * The caller owns every host buffer; the library owns device memory until ssw_db_destroy().
 * The host-buffer entry points of a handle (ssw_scan_topk, ssw_scan_topk_sharded, ssw_rescore, ssw_score_all,
 * ssw_topk_from_scores, ssw_db_set_boxes*, ssw_db_attach_exact) serialise on a per-handle mutex and may be
 * called from several threads; the *_device entry points enqueue on the caller's stream and share the handle's
 * workspace, so use them from one thread / stream at a time.  There is NO CPU fallback:
 * without a CUDA device of compute capability 10.x every compute entry point fails with
 * SSW_ERR_NO_DEVICE.
 *
 * Tie-breaking definition (the reference's np.argsort is unstable, so this is imposed):
 *   rows are ranked by (score desc, original row asc); an image is represented by its best row;
 *   images are ranked by (best score desc, best row asc).  kNN columns by (fp32(1-dot) asc, column asc).
 */
#ifndef SEESAW_B200_H_
#define SEESAW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssw_db ssw_db;

enum ssw_dtype { SSW_F32 = 0, SSW_F16 = 1 };
enum ssw_synth_kind { SSW_SYNTH_TRI = 0, SSW_SYNTH_LATTICE = 1 };
enum ssw_status {
  SSW_OK = 0,
  SSW_ERR_INVALID = 1,     /* bad argument (null pointer, unsupported dim, k out of range ...) */
  SSW_ERR_NO_DEVICE = 2,   /* no sm_100 device / CUDA driver: there is no CPU path            */
  SSW_ERR_CUDA = 3,        /* a CUDA runtime call or kernel failed                            */
  SSW_ERR_OOM = 4
};

#define SSW_MAX_TOPK 2048      /* largest k of the fused scan epilogue                        */
#define SSW_MAX_BATCH 64       /* queries per tensor-core batched pass (larger nq is looped)  */
#define SSW_MAX_KNN_K1 64      /* largest n_neighbors+1 of the fused kNN epilogue             */
#define SSW_MAX_WORLD 8        /* GPUs of one NVSwitch box a database can be sharded over     */

const char* ssw_last_error(void);
int ssw_version(void);
/* number of usable sm_100 devices; 0 (and SSW_OK) when none */
int ssw_device_count(int* out_count);

/* ---- database --------------------------------------------------------------------------
 * Replaces the in-RAM arrays of MultiscaleIndex.__init__ / CoarseIndex.__init__
 * (indices/multiscale/multiscale_index.py:203-231, indices/coarse/coarse_index.py:22-30):
 * `vectors` [n_rows, dim] row-major in ORIGINAL order plus vector_meta.dbidx per row.
 * The device copy is stably grouped by dbidx (CSR over images); results always refer to
 * original row positions.  dtype_in is the host buffer's type, dtype_store the HBM type
 * (fp16 storage of fp32 input rounds to nearest: the "looser bound" case of the north star).
 * global_row_base: original index of row 0 when this is one shard of a row-sharded database
 * (tie-breaking and returned rows are global).  dim must be a multiple of 256 and <= 1024. */
int ssw_db_create(ssw_db** out, int device, const void* vectors, int dtype_in, int dtype_store,
                  int64_t n_rows, int dim, const int32_t* dbidx_per_row, int64_t global_row_base);
/* Same, with the vectors generated in HBM by the counter-based generator of
 * seesaw_b200/synth.py (row r of this shard is global row global_row_base + r). */
int ssw_db_create_synthetic(ssw_db** out, int device, int dtype_store, int64_t n_rows, int dim,
                            const int32_t* dbidx_per_row, int64_t global_row_base,
                            uint64_t seed, int kind);
/* Same as ssw_db_create with the vectors already in the memory of `device` (e.g. embeddings an encoder just
 * produced there): [n_rows, dim] of dtype_in, rows grouped by ascending dbidx; copied / converted device to device. */
int ssw_db_create_device(ssw_db** out, int device, const void* d_vectors, int dtype_in, int dtype_store,
                         int64_t n_rows, int dim, const int32_t* dbidx_per_row, int64_t global_row_base);
int ssw_db_destroy(ssw_db* db);
int ssw_db_info(const ssw_db* db, int64_t* n_rows, int64_t* n_images, int* dim, int* dtype_store,
                int* device);
/* Exact mode.  The reference stores and scans float32 vectors (multiscale_tools.py:200, multiscale_index.py:171);
 * fp16 storage halves the bytes of the scan but rounds every score.  ssw_db_attach_exact keeps the caller's
 * float32 rows [n_rows, dim] (ORIGINAL order, host pointer) in HBM next to the fp16 copy.  From then on
 * ssw_scan_topk / ssw_scan_topk_sharded return the float32 top-k — the fp16 scan proposes more candidates than
 * asked, they are re-scored from the float32 rows and the first k are accepted only when an error bound proves
 * no other image can rank among them; otherwise that query is re-scanned over the float32 rows —
 * and ssw_rescore / ssw_score_all read the float32 rows.  Costs 2x the fp16 bytes of HBM; fp16 storage only.
 * ssw_db_exact_info: rho = max_i ||v_i - fp16(v_i)||, vmax = max_i ||v_i||, and the number of queries answered
 * in exact mode / of those re-scanned in float32 since creation (any pointer may be NULL). */
int ssw_db_attach_exact(ssw_db* db, const float* vectors_f32);
int ssw_db_exact_info(ssw_db* db, int* attached, double* rho, double* vmax, int64_t* queries, int64_t* rescans);
/* device pointer to the stored vectors (grouped order) — for zero-copy consumers (kNN build) */
int ssw_db_vectors_device(const ssw_db* db, void** dev_ptr);

/* ---- stage-1 scan ----------------------------------------------------------------------
 * Replaces _get_top_exact + _get_top_dbidxs as called by MultiscaleIndex._query_prelim
 * (multiscale_index.py:170-175, 189-199, 291-312) and the mask/matvec/argsort of
 * CoarseIndex.query (coarse_index.py:57-96); also serves VectorIndex.query
 * (vector_index.py:55-60) when every row is its own image.
 * For each of nq fp32 queries [nq, dim]: s = V.q per row, per-image max over rows, images whose
 * dbidx is in that query's exclude list dropped, best k images returned best-first.
 * exclude_dbidx / exclude_offsets [nq+1] form per-query id lists (any order, ids absent from
 * the database are ignored); both may be NULL for "no exclusion".
 * Outputs are [nq, k] row-major; out_count[q] = min(k, #eligible images); slots past the count
 * are filled with dbidx -1, score -inf, row -1.  Any output pointer may be NULL. */
int ssw_scan_topk(ssw_db* db, const float* queries, int nq, int k,
                  const int32_t* exclude_dbidx, const int64_t* exclude_offsets,
                  int32_t* out_dbidx, float* out_score, int64_t* out_row, int32_t* out_count);

/* Device-resident variant used for the HBM-resident benchmark leg and the multi-GPU path.
 * All pointers are DEVICE pointers on db's device; `stream` is a cudaStream_t (0 = default).
 * exclude_bits: [nq, ssw_exclude_words(db)] uint32 bitmaps over LOCAL image indices, or NULL.
 * out_key [nq,k] uint64 = (order-preserving score bits << 32) | ~global_row, 0 for empty slots,
 * out_dbidx [nq,k] int32; optional decoded outputs d_out_score / d_out_row [nq,k], d_out_count [nq]
 * (each may be NULL).  Asynchronous: returns after enqueueing. */
int ssw_scan_topk_device(ssw_db* db, const float* d_queries, int nq, int k,
                         const uint32_t* d_exclude_bits, uint64_t* d_out_key, int32_t* d_out_dbidx,
                         float* d_out_score, int64_t* d_out_row, int32_t* d_out_count, void* stream);
int ssw_exclude_words(const ssw_db* db, int64_t* words_per_query);
/* Build the bitmaps on the device from id lists already in device memory. */
int ssw_exclude_build_device(ssw_db* db, const int32_t* d_exclude_dbidx, const int64_t* d_exclude_offsets,
                             int nq, int64_t n_ids_total, uint32_t* d_bits_out, void* stream);
/* Merge n_lists candidate lists per query ([n_lists, nq, k] keys + dbidx, e.g. the all-gathered
 * per-shard outputs of ssw_scan_topk_device) into the global top-k, and decode it.
 * d_out_* are device pointers [nq,k] (each may be NULL). */
int ssw_merge_topk_device(int device, const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists,
                          int nq, int k, uint64_t* d_out_key, int32_t* d_out_dbidx,
                          float* d_out_score, int64_t* d_out_row, int32_t* d_out_count, void* stream);

/* ---- row-sharded database: one process per GPU ------------------------------------------
 * The reference has no multi-device path: images are split into contiguous ranges,
 * one shard (ssw_db with global_row_base) per GPU.  ssw_scan_topk_sharded_device scans this rank's
 * shard and finishes the step in ONE kernel: merge of the shard's lists, stores of the shard's top-k
 * into every rank's exchange buffer over NVLink, flag hand-shake, merge of the world's lists — every
 * rank ends with the same global top-k (outputs as ssw_merge_topk_device).  All ranks must call it
 * with the same nq, k, capacities and `epoch` (non-zero, +1 per call).  nq <= min(nq_cap, #SMs).
 * Exchange buffers: ssw_xchg_create allocates and zeroes this rank's buffer and returns its 64-byte
 * CUDA IPC handle; the host layer exchanges handles (any transport) and maps each peer's buffer with
 * ssw_xchg_open; peer_bufs[r] is rank r's buffer as seen from this device (peer_bufs[rank] = own). */
int ssw_xchg_create(int device, int world, int nq_cap, int k_cap, void** d_buf, void* ipc_handle_out,
                    int64_t* bytes);
int ssw_xchg_open(int device, const void* ipc_handle, void** d_peer);
int ssw_xchg_close(int device, void* d_peer);
int ssw_xchg_destroy(int device, void* d_buf);
int ssw_scan_topk_sharded_device(ssw_db* db, const float* d_queries, int nq, int k,
                                 const uint32_t* d_exclude_bits, void* const* peer_bufs, int world, int rank,
                                 int nq_cap, int k_cap, uint32_t epoch, uint64_t* d_out_key,
                                 int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                                 int32_t* d_out_count, void* stream);

/* Pipelined form of the same step for a stream of batches (a serving loop): the scan of this call runs on `stream`,
 * its exchange + merge on an internal stream UNDER THE SCAN OF THE NEXT CALL (two sets of scan workspaces alternate).
 * On small shards (< 2.5 GB) the pipelined scan runs on (SM count - 4) CTAs and the exchange on 4 blocks too large to
 * share an SM with a scan CTA, so the two never compete for an SM; on larger shards one small exchange block per query
 * runs NEXT TO the scan CTAs.  ssw_scan_pipeline_side_sms(db, s) fixes the number of SMs left to the exchange (0 = none,
 * -1 = the automatic choice; call it between steps, after a drain), ..._in_use reports the choice in force.
 * The outputs of a call are complete on `stream`
 * once the NEXT pipelined call on this handle has been enqueued, or after ssw_scan_pipeline_drain(db, stream) —
 * pass distinct output buffers to consecutive calls.  Same arguments and collective contract as above; nq <= 64,
 * k <= 64, fp16 storage without an exact copy (the batched kernel).  Throughput at 8 GPUs no longer pays the
 * exchange latency and the skew between ranks per step. */
int ssw_scan_topk_sharded_pipelined_device(ssw_db* db, const float* d_queries, int nq, int k,
                                           const uint32_t* d_exclude_bits, void* const* peer_bufs, int world, int rank,
                                           int nq_cap, int k_cap, uint32_t epoch, uint64_t* d_out_key,
                                           int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                                           int32_t* d_out_count, void* stream);
int ssw_scan_pipeline_drain(ssw_db* db, void* stream);
int ssw_scan_pipeline_side_sms(ssw_db* db, int side_sms);
int ssw_scan_pipeline_side_sms_in_use(const ssw_db* db, int* side_sms);

/* Host-buffer form of the sharded step (arguments as ssw_scan_topk + the exchange arguments above):
 * one H2D of queries and id lists, bitmap build, scan, fused exchange + merge, one D2H of the results. */
int ssw_scan_topk_sharded(ssw_db* db, const float* queries, int nq, int k, const int32_t* exclude_dbidx,
                          const int64_t* exclude_offsets, void* const* peer_bufs, int world, int rank,
                          int nq_cap, int k_cap, uint32_t epoch, int32_t* out_dbidx, float* out_score,
                          int64_t* out_row, int32_t* out_count);

/* Kernel selection for ssw_scan_topk*: 0 = auto (streaming SIMT kernel for a single query or k > 64,
 * tcgen05 batched kernel otherwise), 1 = force streaming kernel, 2 = force tcgen05 kernel. */
int ssw_set_scan_mode(ssw_db* db, int mode);

/* ---- full score vector -----------------------------------------------------------------
 * Replaces MultiscaleIndex.score / CoarseIndex.score (multiscale_index.py:284-285,
 * coarse_index.py:37-38): out_scores[n_rows] fp32 in ORIGINAL row order (host pointer). */
int ssw_score_all(ssw_db* db, const float* query, float* out_scores);
int ssw_score_all_device(ssw_db* db, const float* d_query, float* d_out_scores, void* stream);

/* ---- stage 2: rescoring of the shortlisted images ----------------------------------------
 * Replaces the gather + matvec of MultiscaleIndex.query (multiscale_index.py:341-349) and
 * rescore_candidates / score_frame2 (:379-403, :112-150) for the candidate images of stage 1.
 * ssw_db_set_boxes uploads vector_meta's x1,y1,x2,y2,zoom_level per ORIGINAL row (needed by
 * 'avg_score', the reference's default aggregation).  ssw_rescore: for every candidate dbidx the
 * aggregated score of its best patch (float64) and that patch's ORIGINAL row; scores are
 * vectors.q [- vectors.query2].  agg_method 0 = plain_score, 1 = avg_score; aug_larger 0 = all,
 * 1 = greater, 2 = adjacent.  The caller orders the images (ascending dbidx, stable by score) and
 * takes topk, as rescore_candidates does (:388-399).  Ids not in the database give row -1. */
int ssw_db_set_boxes(ssw_db* db, const int32_t* x1, const int32_t* y1, const int32_t* x2, const int32_t* y2,
                     const int32_t* zoom_level);
/* Same with the box columns in their own type.  The reference computes the IoU of the self-join in the dtype of
 * vector_meta's x1,y1,x2,y2 (box_utils.py:336-350 via torchvision's _box_inter_union): float32 for indices
 * written by the tiling pipeline (multiscale_tools.py:111: pixel / scale_factor as float32), exact integers with
 * a float32 quotient for integer columns, float64 for float64 columns; the kernel repeats those operations in
 * that type and order, so the per-level best-IoU patch is the reference's.  x1..y2: [n_rows] of box_dtype. */
enum ssw_box_dtype { SSW_BOX_I32 = 0, SSW_BOX_F32 = 1, SSW_BOX_F64 = 2 };
int ssw_db_set_boxes_typed(ssw_db* db, int box_dtype, const void* x1, const void* y1, const void* x2, const void* y2,
                           const int32_t* zoom_level);
int ssw_rescore(ssw_db* db, const float* query, const float* query2, const int32_t* cand_dbidx, int n_cand,
                int agg_method, int aug_larger, double* out_score, int64_t* out_row);

/* ---- top-k images from a caller-supplied score per row ---------------------------------
 * Replaces _get_top_dbidxs when the scores are not a dot product with the stored vectors: label
 * propagation over the kNN graph (KnnProp2.next_batch, loops/graph_based.py:97-99, calls
 * _get_top_dbidxs(vec_idxs, scores, ...) of multiscale_index.py:189-199 on ALL sorted rows).
 * scores[n_rows] fp32 in ORIGINAL row order; row_mask[n_rows] (0 = row absent, e.g. labeled rows;
 * NULL = all present); exclude_dbidx[n_exclude] image ids to drop.  Same ranking and outputs as
 * ssw_scan_topk for one query: per-image max, ties to the lower row, best k images. */
int ssw_topk_from_scores(ssw_db* db, const float* scores, const uint8_t* row_mask, int k,
                         const int32_t* exclude_dbidx, int64_t n_exclude, int32_t* out_dbidx,
                         float* out_score, int64_t* out_row, int32_t* out_count);

/* The same when the caller already holds the rows in ranked order — exactly _get_top_dbidxs' own input
 * (multiscale_index.py:189-199 takes vec_idxs, the rows best first): row_order[n_order] ORIGINAL row numbers best
 * first (rows not listed take no part).  An image is represented by its earliest-listed row and images are ranked
 * by that position, so the result follows the caller's order whatever the dtype of the scores behind it (label
 * propagation scores are float64; ties keep the caller's order).  out_pos[k] = position in row_order of each image's
 * best row (the caller reads its score there), out_row[k] = that row. */
int ssw_topk_from_order(ssw_db* db, const int64_t* row_order, int64_t n_order, int k, const int32_t* exclude_dbidx,
                        int64_t n_exclude, int32_t* out_dbidx, int64_t* out_pos, int64_t* out_row, int32_t* out_count);

/* ---- exact kNN graph -------------------------------------------------------------------
 * Replaces the matmul + row argsort of compute_exact_knn (knn_graph.py:170-182):
 * for rows [row_begin,row_end) of vectors [n, dim]: the k1 = min(n_neighbors+1, n) columns j
 * minimising (fp32(1 - dot(i,j)), j), self included when it ranks.  out_idx/out_dist are
 * [(row_end-row_begin), k1] host buffers.  The caller runs post_process_graph_df unchanged
 * (knn_graph.py:142-168).  The tensor cores multiply fp16 values.  For float32 input that is not
 * fp16-representable they only PROPOSE candidates (k1 + max(16, k1), at most SSW_MAX_KNN_K1); the distances are then
 * recomputed from the float32 rows, the first k1 accepted when an error bound proves no other column can rank among
 * them, and rows that cannot be certified (near-duplicates) re-scanned against all columns in float32 — the result
 * is the float32 neighbour list the reference computes, under column tie-breaking (ssw_knn_exact.cu). */
int ssw_knn_build(int device, const void* vectors, int dtype_in, int64_t n, int dim, int k1,
                  int64_t row_begin, int64_t row_end, int32_t* out_idx, float* out_dist);
/* After ssw_knn_build / ssw_knn_graph on this thread: rows whose candidates were re-ranked in float32, rows of those
 * that needed the full float32 re-scan, and rho = max_i ||v_i - fp16(v_i)|| (0: the fp16 pass was already exact). */
int ssw_knn_exact_stats(int64_t* rows_refined, int64_t* rows_rescanned, double* rho);
/* d_vectors_f16: [n, dim] fp16 already in HBM; outputs device [(row_end-row_begin), k1]. */
int ssw_knn_build_device(int device, const void* d_vectors_f16, int64_t n, int dim, int k1,
                         int64_t row_begin, int64_t row_end, int32_t* d_out_idx, float* d_out_dist,
                         void* stream);

/* The whole compute_exact_knn (knn_graph.py:170-191) in one call: candidates (above) + post_process_graph_df
 * (:142-168) on the device.  Outputs are the four columns of the reference's edge table in its order
 * (src_vertex, dst_rank): int32 src / dst, float32 distance (clipped at 0), int32 dst_rank; one rank-0
 * self edge per vertex; *out_edges rows are valid.  capacity >= n * (min(n_neighbors+1, n) + 1). */
int ssw_knn_graph(int device, const void* vectors, int dtype_in, int64_t n, int dim, int n_neighbors,
                  int32_t* out_src, int32_t* out_dst, float* out_distance, int32_t* out_rank, int64_t capacity,
                  int64_t* out_edges);
/* Device-resident post-processing of a candidate table [rows, k1] (rows are vertices src_offset ..);
 * d_workspace: ssw_knn_edges_workspace_bytes(rows); outputs sized rows * (k1 + 1); *d_total = edges. */
int ssw_knn_edges_workspace_bytes(int64_t rows, int64_t* bytes);
int ssw_knn_edges_device(int device, const int32_t* d_idx, const float* d_dist, int64_t rows, int k1,
                         int64_t src_offset, int32_t* d_src, int32_t* d_dst, float* d_distance, int32_t* d_rank,
                         int64_t* d_total, void* d_workspace, void* stream);

/* ---- weight matrix of the kNN graph --------------------------------------------------------
 * Replaces get_weight_matrix(df, kfun=, self_edges=False, normalized=False, symmetric=True) (knn_graph.py:31-104), the
 * matrix KnnProp2 propagates over (loops/graph_based.py:36-43).  src/dst/weight[n_edges]: the edge table sorted by
 * src_vertex (vertices 0 .. n-1, one self edge each, at most one edge per ordered pair) with weight = kfun(distance)
 * evaluated by the caller (the reference's own numpy function: values stay bit-identical).  Output: CSR with sorted
 * indices — both directions of every listed edge, value = (sum of the listed positive weights) / (number of
 * listings), explicit zeros on the diagonal — exactly the arrays the reference builds; capacity >= 2 * n_edges
 * entries, *out_nnz valid.  out_weight_sum[n] (may be NULL) = W.sum(0) as LabelPropagation computes it
 * (label_propagation.py:24).  Fails with SSW_ERR_INVALID when a vertex ends with zero degree (:77). */
int ssw_weight_matrix(int device, const int32_t* src, const int32_t* dst, const double* weight, int64_t n_edges,
                      int64_t n, int64_t* out_indptr, int32_t* out_indices, double* out_data, int64_t capacity,
                      int64_t* out_nnz, double* out_weight_sum);

/* ---- label propagation over the kNN graph ----------------------------------------------
 * Replaces the iteration of LabelPropagation.fit_transform / _step (label_propagation.py:30-83):
 * new = (W @ old + reg_lambda * reg_values) / (weight_sum + reg_lambda); new[label_ids] = label_values;
 * stop when max((new - old)^2) < epsilon.  W (CSR, float64, sorted indices) is what get_weight_matrix
 * (knn_graph.py:31-104) builds from the kNN edge table; weight_sum[n] = W.sum(0) as the reference
 * computes it.  All arithmetic is IEEE float64 in scipy's order: iterates are bit-identical.
 * ssw_lp_fit returns the iterate fit_transform returns (the previous one on convergence), the number
 * of iterations run and whether it converged.  reg_values / start_value may be NULL as in the reference
 * (reg_values NULL requires reg_lambda == 0). */
typedef struct ssw_lp ssw_lp;
int ssw_lp_create(ssw_lp** out, int device, int64_t n, const int64_t* indptr, const int32_t* indices,
                  const double* data, const double* weight_sum, double reg_lambda);
int ssw_lp_destroy(ssw_lp* lp);
int ssw_lp_fit(ssw_lp* lp, const int64_t* label_ids, const double* label_values, int64_t n_labels,
               const double* reg_values, const double* start_value, int max_iter, double epsilon,
               double* out_values, int* out_iterations, int* out_converged);

/* Same iteration with the prior term handed over ready-made: lambda_reg_values[n] = reg_lambda * reg_values AS THE
 * CALLER'S ARITHMETIC PRODUCED IT (the reference multiplies in the dtype of reg_values — float32 priors from
 * index.score give a float32 product, label_propagation.py:31 — and only then adds in float64), and x0[n] the start
 * iterate (start_value, else reg_values, else zeros; labels are clamped by the library).  NULL prior = zeros. */
int ssw_lp_fit_scaled(ssw_lp* lp, const int64_t* label_ids, const double* label_values, int64_t n_labels,
                      const double* lambda_reg_values, const double* x0, int max_iter, double epsilon,
                      double* out_values, int* out_iterations, int* out_converged);

/* ---- introspection for benchmarks / tests ---------------------------------------------- */
/* With profiling on, launches of the dominant scan kernel (K1 streaming or K2 tcgen05) are bracketed by CUDA
 * events on their own stream: every launch for on == 1, every on-th launch for on > 1 (an event between two
 * kernels rules out their programmatic dependent launch, so a sampled step runs without that overlap and the
 * others with it).  ssw_profile_read synchronises those events and
 * returns the summed kernel time and launch count since the last read, then resets them. */
int ssw_profile_enable(ssw_db* db, int on);
int ssw_profile_read(ssw_db* db, double* scan_kernel_ms, int64_t* scan_kernel_launches);
/* Work counters of the fused top-k epilogue, summed over all CTAs and queries since the last call: images that
 * passed the cheap threshold vote and were offered to a list, and list updates (appends + replacements) — the
 * data-dependent part of the scan (pooled thresholds).  enable != 0 switches counting on (the call
 * returns the counts so far and resets them), 0 off.  Synchronises the device. */
int ssw_scan_stats(ssw_db* db, int enable, int64_t* list_updates, int64_t* images_offered);
/* number of kernels this library has launched since load (all handles) */
int64_t ssw_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SEESAW_B200_H_ */
