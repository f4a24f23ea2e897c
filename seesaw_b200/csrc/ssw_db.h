// Internal definition of the database handle and kernel launch entry points.
#pragma once
#include <mutex>
#include <utility>
#include <vector>

#include "ssw_common.cuh"

struct ssw_db {
  int device = 0;
  int dim = 0;
  int dtype = SSW_F16;          // storage type in HBM
  int64_t n_rows = 0;
  int64_t n_images = 0;
  int64_t row_base = 0;         // global index of this shard's original row 0
  int sm_count = 0;
  int scan_mode = 0;

  // HBM layout (SURVEY.md §7 "row order is not guaranteed grouped by image"):
  void* d_vecs = nullptr;          // [n_rows, dim] stored type, rows STABLY grouped by image
  int64_t* d_row_ptr = nullptr;    // [n_images + 1] CSR over device rows
  int32_t* d_img_dbidx = nullptr;  // [n_images] dbidx of each local image, ascending
  int64_t* d_orig_row = nullptr;   // [n_rows] device row -> original local row; NULL when identity
  int32_t* d_part = nullptr;       // [scan_warps + 1] image range of every scan warp
  int scan_grid = 0;               // CTAs of the streaming scan (one per SM)
  int64_t max_cta_images = 0;      // most images any CTA's range holds
  uint32_t* d_last_bits = nullptr; // [n_rows/32 + pad] bit r set <=> device row r is the last row of its image
  void* d_boxes = nullptr;         // [n_rows][4] x1,y1,x2,y2 per device row, element type box_kind (stage 2; optional)
  int32_t* d_zoom = nullptr;       // [n_rows] zoom level per device row
  int box_kind = 0;                // SSW_BOX_I32 / SSW_BOX_F32 / SSW_BOX_F64
  // exact mode (ssw_db_attach_exact): the reference's own fp32 values next to the fp16 scan copy
  float* d_exact = nullptr;        // [n_rows, dim] fp32, same (grouped) row order as d_vecs
  double exact_rho = 0.0;          // max over rows of ||v - fp16(v)||_2
  double exact_vmax = 0.0;         // max over rows of ||v||_2
  int64_t exact_queries = 0;       // queries answered in exact mode / of those, re-scanned in fp32
  int64_t exact_rescans = 0;
  std::mutex mu;                   // serialises the host-buffer entry points (they share the staging blocks)
  std::vector<int32_t> h_img_dbidx; // host copy of d_img_dbidx (candidate id -> image index), filled on first use
  unsigned long long* d_scan_stats = nullptr;  // [2] list updates / images offered, counted by the scan kernels (ssw_scan_stats)
  int* d_xchg_timed_out = nullptr; // pinned, device-visible host word: set by the fused exchange kernel when a peer never answered
  void* d_tc_ws = nullptr;         // prepared A operand of the tcgen05 batched scan (one batch)

  int64_t excl_words = 0;          // uint32 words of one exclusion bitmap (n_images bits, padded)

  // per-handle workspace (grown on demand), all on `stream` unless the caller passes one
  cudaStream_t stream = nullptr;
  uint64_t* d_list_keys = nullptr; // [nq][lists][k] per-CTA candidate lists
  int32_t* d_list_dbidx = nullptr;
  uint64_t* d_gthr = nullptr;      // [nq] shared lower bound on the k-th best key, followed by
  uint32_t* d_pub1 = nullptr;      // [nq][grid] best score per CTA of the streaming scan (same allocation)
  int32_t* d_cand_cnt = nullptr;   // [nq] compacted candidates per query of the batched scan (same allocation)
  size_t list_capacity = 0;        // entries (per pipeline parity: two such blocks are allocated)
  // pipelined sharded step: the exchange of step i runs on `xstream` under the scan of step i+1
  cudaStream_t xstream = nullptr;
  cudaEvent_t ev_scan = nullptr, ev_xdone = nullptr;
  bool pipe_pending = false;
  // ... on `side_sms` SMs of its own: the pipelined scan runs on a grid of sm_count - side_sms CTAs (its own partition)
  int side_sms = -1;               // -1 = automatic (by shard size), 0 = the exchange blocks co-reside with the scan CTAs
  int side_grid = 0;               // grid d_part_side was built for (0 = not built)
  int32_t* d_part_side = nullptr;  // [side_grid * kScanWarps + 1]
  int64_t side_max_cta_images = 0;
  struct ScanWs {                  // the second set of scan workspaces (swapped with the fields above per pipelined call)
    uint64_t* keys = nullptr;
    int32_t* dbidx = nullptr;
    uint64_t* gthr = nullptr;
    uint32_t* pub1 = nullptr;
    int32_t* cand_cnt = nullptr;
    size_t list_capacity = 0;
    int gthr_capacity = 0;
  } ws_alt;
  int gthr_capacity = 0;
  // staging for the host-pointer API
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;
  void* h_stage = nullptr;         // pinned
  size_t h_stage_bytes = 0;
  // optional per-launch timing of the scan kernel (ssw_profile_enable)
  bool profiling = false;
  int prof_stride = 1;             // bracket every prof_stride-th scan-kernel launch
  int64_t prof_counter = 0;
  bool prof_sampled = false;       // the launch being enqueued is bracketed (no programmatic dependent launch around it)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_free;
};

namespace ssw {

#ifdef SSW_TRACE
extern unsigned long long* g_trace_block;   // development timeline block of the handle that enabled ssw_scan_stats
#endif

constexpr int kScanWarps = 8;             // warps per CTA of the streaming scan
constexpr int kMergeCap = 8192;           // candidates the merge kernel sorts in shared memory

int ensure_device(int device, int* sm_count);
int ensure_stage(ssw_db* db, size_t dev_bytes, size_t host_bytes);
// event pair around the scan kernel when profiling is on (no-ops otherwise)
void prof_begin(ssw_db* db, cudaStream_t st);
void prof_end(ssw_db* db, cudaStream_t st);

// streaming single-query scan (K1); MODE 0 = fused segmented max + exclusion + top-k lists,
// MODE 1 = plain score vector (index.score)
// `exact`: scan the fp32 copy (db->d_exact) instead of the stored vectors
int launch_scan1(ssw_db* db, const float* d_query, int k, const uint32_t* d_excl, uint64_t* d_list_keys,
                 int32_t* d_list_dbidx, uint64_t* d_gthr, uint32_t* d_pub, cudaStream_t st, bool exact = false);
int launch_score_all(ssw_db* db, const float* d_query, float* d_out, cudaStream_t st);
// exact mode: fp32 re-scoring of the candidate images of a scan, and the certified final top-k
int launch_exact_rescore(ssw_db* db, const float* d_queries, int nq, int kc, const int32_t* d_cand_dbidx,
                         uint64_t* d_keys32, cudaStream_t st);
int launch_exact_finish(const float* d_queries, int dim, int nq, int kc, int k, const uint64_t* d_keys16,
                        const uint64_t* d_keys32, const int32_t* d_cand_dbidx, double err_per_unit_q,
                        int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                        uint64_t* d_out_key, int32_t* d_certified, cudaStream_t st);
int launch_row_error_stats(const float* d_rows_f32, int64_t n_rows, int dim, float* d_stats2, cudaStream_t st);
// tcgen05 batched scan (K2): one pass for up to 64 queries; fp16 storage, dim 256/512/768, k <= 64
bool scan_tc_supported(const ssw_db* db, int k);
size_t scan_tc_workspace_bytes(int dim, int grid);
// d_cand_keys / d_cand_dbidx: [nq][grid * (k + kScanTcSlack)] compacted candidates, d_cand_cnt [nq] their number per query
constexpr int kScanTcSlack = 8;
// `grid`: CTAs of this launch with their partition (null = the database's own: one CTA per SM)
struct ScanTcGrid {
  int grid;
  const int32_t* part;             // [grid * kScanWarps + 1]
  int64_t max_cta_images;
};
int launch_scan_tc(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_excl, uint64_t* d_cand_keys,
                   int32_t* d_cand_dbidx, int32_t* d_cand_cnt, uint64_t* d_gthr, void* workspace, cudaStream_t st,
                   const ScanTcGrid* grid = nullptr);
// image partition for `grid` CTAs computed on the device (same rule as the database's own); d_max_cta: most images per CTA
int launch_part_build(ssw_db* db, int grid, int32_t* d_part, int* d_max_cta, cudaStream_t st);
// merge kernel (K4)
int launch_merge(const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists, int64_t list_stride,
                 int64_t query_stride, int nq, int k, const uint64_t* d_thr, uint64_t* d_out_key,
                 int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                 cudaStream_t st, bool pdl = false, const int32_t* d_counts = nullptr);
// fused shard merge + peer exchange + world merge (multi-GPU); peers[] are peer-mapped exchange buffers
size_t xchg_bytes(int world, int nq_cap, int k_cap);
int launch_exchange_merge(const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists, int64_t list_stride,
                          int64_t query_stride, int nq, int k, const uint64_t* d_thr, void* const* peers, int world,
                          int rank, int nq_cap, int k_cap, uint32_t epoch, int* d_timed_out, uint64_t* d_out_key,
                          int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                          cudaStream_t st, bool pdl = false, const int32_t* d_counts = nullptr);
// slim exchange for the pipelined sharded step: side_blocks > 0 = that many 1024-thread blocks on SMs of their own,
// 0 = one 128-thread block per query next to the scan CTAs
int launch_exchange_slim(const uint64_t* d_keys, const int32_t* d_dbidx, int64_t query_stride, int nq, int k,
                         const int32_t* d_counts, void* const* peers, int world, int rank, int nq_cap, int k_cap, uint32_t epoch,
                         int* d_timed_out, uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                         int32_t* d_out_count, cudaStream_t st, int side_blocks);
int launch_image_max(ssw_db* db, const float* d_scores, const uint8_t* d_row_mask, const uint32_t* d_excl,
                     int64_t n_padded, uint64_t* d_keys, int32_t* d_dbidx, cudaStream_t st, const uint32_t* d_pos = nullptr);
int launch_order_to_pos(const int64_t* d_order, int64_t n_order, int64_t n_rows, uint32_t* d_pos, cudaStream_t st);
int launch_exclude_build(ssw_db* db, const int32_t* d_ids, const int64_t* d_offsets, int nq, uint32_t* d_bits,
                         cudaStream_t st);
int launch_synth(void* d_out, int dtype, int64_t n_rows, int dim, int64_t global_row0, uint64_t seed, int kind,
                 cudaStream_t st);
int launch_convert_rows(const void* d_src, int dtype_src, void* d_dst, int dtype_dst, int64_t n_elems,
                        cudaStream_t st);

}  // namespace ssw
