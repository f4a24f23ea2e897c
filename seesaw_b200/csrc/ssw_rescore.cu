// K7: stage 2 of the multiscale query on the device — rescoring of the <= shortlist_size candidate images.
//
// Replaces, for the candidates of stage 1, the host work of MultiscaleIndex.query (seesaw/indices/multiscale/
// multiscale_index.py:341-351: gather all rows of the shortlisted images, scores = vectors @ q [- vectors @ q2])
// and rescore_candidates / score_frame2 (:379-403, :112-150): per image either the first patch attaining the
// maximum score ('plain_score') or, for every patch, the mean over zoom levels of the score of the patch of
// that level it overlaps most (IoU > 0; 'avg_score', the reference's default — basic_types.py:58), then the
// best patch.  One CTA per candidate image; the image's rows are contiguous in HBM (rows are grouped by
// image), boxes and zoom levels sit next to them in device-row order.
//
// Arithmetic follows the reference operation by operation.  Patch scores are fp32 dot products (canon_dot;
// from the fp32 copy when one is attached).  The IoU is torchvision's _box_inter_union as box_iou calls it
// (seesaw/box_utils.py:336-350) IN THE BOX COLUMNS' OWN TYPE: the tiling pipeline writes float32 boxes
// (multiscale_tools.py:111), so area, intersection, union and the quotient are float32 operations in that
// order; integer boxes give exact integer intersection / union and a float32 quotient (torch's true division
// of integer tensors); float64 boxes stay float64.  The per-level best is the FIRST maximum of that IoU
// (groupby(...).iou.idxmax(), multiscale_index.py:140); the mean over levels is pandas' group_mean on a float32
// column — Kahan-compensated float32 summation in ascending level order, float32 division by the count (:142;
// established against pandas on 60k random groups, tests/test_oracle.py) — and the image's best patch the
// first maximum of that (:149-150).
#include <algorithm>
#include <vector>

#include "ssw_db.h"

namespace ssw {

struct RescoreArgs {
  const void* vecs;            // [n_rows, dim] stored type (or the fp32 copy)
  int dtype, dim;
  const void* boxes;           // [n_rows][4] x1,y1,x2,y2 in device-row order (null for plain_score)
  const int32_t* zoom;         // [n_rows]
  const int64_t* row_ptr;
  const int64_t* orig_row;     // may be null
  const float* q;              // [dim]
  const float* q2;             // [dim] or null
  const int32_t* cand_img;     // [n_cand] local image index, -1 = id not in the database
  int agg;                     // 0 plain_score, 1 avg_score
  int aug;                     // 0 all, 1 greater, 2 adjacent
  int max_rows;                // rows per image the shared arrays hold
  double* out_score;           // [n_cand]
  int64_t* out_row;            // [n_cand] ORIGINAL row of the winning patch (-1: no such image)
  int32_t* out_status;         // [n_cand] 0 ok, 1 image too large for the kernel
};

constexpr int kRescoreThreads = 128;

// IoU of two boxes the way the reference computes it for boxes of type BT; the result type is what the
// reference compares (float for integer and float32 boxes, double for float64 boxes).
template <typename BT> struct IouT { using type = float; };
template <> struct IouT<double> { using type = double; };

__device__ __forceinline__ float box_iou_ref(const int32_t* l, const int32_t* r) {
  const long long la = (long long)(l[2] - l[0]) * (l[3] - l[1]), ra = (long long)(r[2] - r[0]) * (r[3] - r[1]);
  const long long w = (long long)min(l[2], r[2]) - max(l[0], r[0]), h = (long long)min(l[3], r[3]) - max(l[1], r[1]);
  const long long inter = (w > 0 ? w : 0) * (h > 0 ? h : 0);
  const long long uni = la + ra - inter;
  return __fdiv_rn(__ll2float_rn(inter), __ll2float_rn(uni));
}
__device__ __forceinline__ float box_iou_ref(const float* l, const float* r) {
  const float la = __fmul_rn(__fsub_rn(l[2], l[0]), __fsub_rn(l[3], l[1]));
  const float ra = __fmul_rn(__fsub_rn(r[2], r[0]), __fsub_rn(r[3], r[1]));
  const float w = fmaxf(__fsub_rn(fminf(l[2], r[2]), fmaxf(l[0], r[0])), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(l[3], r[3]), fmaxf(l[1], r[1])), 0.f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(la, ra), inter));
}
__device__ __forceinline__ double box_iou_ref(const double* l, const double* r) {
  const double la = __dmul_rn(__dsub_rn(l[2], l[0]), __dsub_rn(l[3], l[1]));
  const double ra = __dmul_rn(__dsub_rn(r[2], r[0]), __dsub_rn(r[3], r[1]));
  const double w = fmax(__dsub_rn(fmin(l[2], r[2]), fmax(l[0], r[0])), 0.0);
  const double h = fmax(__dsub_rn(fmin(l[3], r[3]), fmax(l[1], r[1])), 0.0);
  const double inter = __dmul_rn(w, h);
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(la, ra), inter));
}

template <typename BT>
__global__ void __launch_bounds__(kRescoreThreads) rescore_kernel(const RescoreArgs a) {
  using IT = typename IouT<BT>::type;
  extern __shared__ __align__(16) uint8_t rsmem[];
  BT* s_box = reinterpret_cast<BT*>(rsmem);                                          // [max_rows][4]
  float* s_score = reinterpret_cast<float*>(s_box + 4 * (size_t)a.max_rows);          // [max_rows]
  float* s_agg = s_score + a.max_rows;                                                // [max_rows]
  int32_t* s_z = reinterpret_cast<int32_t*>(s_agg + a.max_rows);                      // [max_rows]
  __shared__ int s_zoom[32];
  __shared__ int s_nzoom;
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = a.cand_img[c];
  if (img < 0) {
    if (tid == 0) {
      a.out_score[c] = -INFINITY;
      a.out_row[c] = -1;
      a.out_status[c] = 0;
    }
    return;
  }
  const int64_t r0 = a.row_ptr[img];
  const int P = (int)(a.row_ptr[img + 1] - r0);
  if (P > a.max_rows) {
    if (tid == 0) a.out_status[c] = 1;
    return;
  }
  // ---- patch scores: one warp per row
  for (int p = warp; p < P; p += kRescoreThreads / 32) {
    float s;
    if (a.dtype == SSW_F16) {
      const __half* row = static_cast<const __half*>(a.vecs) + (r0 + p) * (int64_t)a.dim;
      s = canon_dot<__half>(row, a.q, a.dim, lane);
      if (a.q2) s = __fsub_rn(s, canon_dot<__half>(row, a.q2, a.dim, lane));
    } else {
      const float* row = static_cast<const float*>(a.vecs) + (r0 + p) * (int64_t)a.dim;
      s = canon_dot<float>(row, a.q, a.dim, lane);
      if (a.q2) s = __fsub_rn(s, canon_dot<float>(row, a.q2, a.dim, lane));
    }
    if (lane == 0) s_score[p] = s;
  }
  if (a.agg == 1) {
    const BT* gb = static_cast<const BT*>(a.boxes) + r0 * 4;
    for (int i = tid; i < 4 * P; i += kRescoreThreads) s_box[i] = gb[i];
    for (int i = tid; i < P; i += kRescoreThreads) s_z[i] = a.zoom[r0 + i];
  }
  __syncthreads();
  if (a.agg == 1 && tid == 0) {          // distinct zoom levels, ascending (a handful)
    int nz = 0;
    for (int p = 0; p < P; ++p) {
      const int z = s_z[p];
      int j = 0;
      while (j < nz && s_zoom[j] != z) ++j;
      if (j == nz && nz < 32) s_zoom[nz++] = z;
    }
    for (int i = 1; i < nz; ++i) {
      const int z = s_zoom[i];
      int j = i - 1;
      while (j >= 0 && s_zoom[j] > z) {
        s_zoom[j + 1] = s_zoom[j];
        --j;
      }
      s_zoom[j + 1] = z;
    }
    s_nzoom = nz;
  }
  __syncthreads();
  // ---- aggregated score of every (left) patch
  for (int l = tid; l < P; l += kRescoreThreads) {
    float agg;
    if (a.agg == 0) {
      agg = s_score[l];
    } else {
      const BT* lb = s_box + 4 * l;
      const int lz = s_z[l];
      float sumx = 0.f, comp = 0.f;      // pandas' group_mean: Kahan summation in the column's float32
      int levels = 0;
      for (int zi = 0; zi < s_nzoom; ++zi) {
        const int z = s_zoom[zi];
        if ((a.aug == 1 && z < lz) || (a.aug == 2 && z != lz)) continue;
        IT best_iou = 0;
        int best_r = -1;
        for (int r = 0; r < P; ++r) {
          if (s_z[r] != z) continue;
          const IT iou = box_iou_ref(lb, s_box + 4 * r);
          if (iou > best_iou) {            // strict: the first (lowest position) maximum wins; NaN never joins
            best_iou = iou;
            best_r = r;
          }
        }
        if (best_r >= 0) {
          const float y = __fsub_rn(s_score[best_r], comp);
          const float t = __fadd_rn(sumx, y);
          comp = __fsub_rn(__fsub_rn(t, sumx), y);
          sumx = t;
          ++levels;
        }
      }
      agg = levels > 0 ? __fdiv_rn(sumx, (float)levels) : __int_as_float(0x7fc00000);   // NaN: no overlap at all
    }
    s_agg[l] = agg;
  }
  __syncthreads();
  // ---- best patch: max aggregated score, first position on ties; NaN rows never win
  if (tid == 0) {
    float m = -INFINITY;
    int pos = -1;
    for (int p = 0; p < P; ++p)
      if (s_agg[p] > m || (pos < 0 && s_agg[p] == m)) {
        m = s_agg[p];
        pos = p;
      }
    if (pos < 0) pos = 0;
    const int64_t drow = r0 + pos;
    a.out_score[c] = (double)m;
    a.out_row[c] = a.orig_row ? a.orig_row[drow] : drow;
    a.out_status[c] = 0;
  }
}

}  // namespace ssw

using namespace ssw;

extern "C" {

int ssw_db_set_boxes_typed(ssw_db* db, int box_dtype, const void* x1, const void* y1, const void* x2, const void* y2,
                           const int32_t* zoom) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_REQUIRE(box_dtype == SSW_BOX_I32 || box_dtype == SSW_BOX_F32 || box_dtype == SSW_BOX_F64, "unknown box dtype");
  SSW_REQUIRE(db->n_rows == 0 || (x1 && y1 && x2 && y2 && zoom), "null argument");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  const int64_t n = db->n_rows;
  const size_t es = box_dtype == SSW_BOX_F64 ? 8 : 4;
  std::vector<int64_t> perm;
  if (db->d_orig_row) {
    perm.resize(n);
    SSW_CUDA(cudaMemcpy(perm.data(), db->d_orig_row, (size_t)n * 8, cudaMemcpyDeviceToHost));
  }
  std::vector<uint8_t> packed((size_t)n * 4 * es);
  std::vector<int32_t> z((size_t)n);
  const uint8_t* cols[4] = {static_cast<const uint8_t*>(x1), static_cast<const uint8_t*>(y1),
                            static_cast<const uint8_t*>(x2), static_cast<const uint8_t*>(y2)};
  for (int64_t r = 0; r < n; ++r) {
    const int64_t o = perm.empty() ? r : perm[r];
    for (int j = 0; j < 4; ++j) memcpy(&packed[((size_t)r * 4 + j) * es], cols[j] + (size_t)o * es, es);
    z[r] = zoom[o];
  }
  cudaFree(db->d_boxes);
  cudaFree(db->d_zoom);
  db->d_boxes = nullptr;
  db->d_zoom = nullptr;
  SSW_CUDA(cudaMalloc(&db->d_boxes, std::max<size_t>(packed.size(), 16)));
  SSW_CUDA(cudaMalloc((void**)&db->d_zoom, std::max<size_t>(z.size() * 4, 16)));
  if (n) {
    SSW_CUDA(cudaMemcpy(db->d_boxes, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    SSW_CUDA(cudaMemcpy(db->d_zoom, z.data(), z.size() * 4, cudaMemcpyHostToDevice));
  }
  db->box_kind = box_dtype;
  return SSW_OK;
}

int ssw_db_set_boxes(ssw_db* db, const int32_t* x1, const int32_t* y1, const int32_t* x2, const int32_t* y2,
                     const int32_t* zoom) {
  return ssw_db_set_boxes_typed(db, SSW_BOX_I32, x1, y1, x2, y2, zoom);
}

int ssw_rescore(ssw_db* db, const float* query, const float* query2, const int32_t* cand_dbidx, int n_cand, int agg_method,
                int aug_larger, double* out_score, int64_t* out_row) {
  SSW_REQUIRE(db != nullptr && query != nullptr && out_score != nullptr && out_row != nullptr, "null argument");
  SSW_REQUIRE(n_cand >= 0 && (n_cand == 0 || cand_dbidx != nullptr), "bad candidate list");
  SSW_REQUIRE(agg_method == 0 || agg_method == 1, "agg_method: 0 = plain_score, 1 = avg_score");
  SSW_REQUIRE(aug_larger >= 0 && aug_larger <= 2, "aug_larger: 0 = all, 1 = greater, 2 = adjacent");
  SSW_REQUIRE(agg_method == 0 || db->d_boxes != nullptr, "avg_score needs the patch boxes: call ssw_db_set_boxes first");
  if (n_cand == 0) return SSW_OK;
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  // host: dbidx -> local image index needs the id table; binary search on a host copy kept with the handle
  if (db->h_img_dbidx.size() != (size_t)db->n_images) {
    db->h_img_dbidx.resize(db->n_images);
    if (db->n_images)
      SSW_CUDA(cudaMemcpy(db->h_img_dbidx.data(), db->d_img_dbidx, (size_t)db->n_images * 4, cudaMemcpyDeviceToHost));
  }
  std::vector<int32_t> img(n_cand);
  for (int i = 0; i < n_cand; ++i) {
    auto it = std::lower_bound(db->h_img_dbidx.begin(), db->h_img_dbidx.end(), cand_dbidx[i]);
    img[i] = (it != db->h_img_dbidx.end() && *it == cand_dbidx[i]) ? (int32_t)(it - db->h_img_dbidx.begin()) : -1;
  }
  auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
  const size_t q_b = up16((size_t)db->dim * 4), c_b = up16((size_t)n_cand * 4);
  const size_t os_b = up16((size_t)n_cand * 8), or_b = up16((size_t)n_cand * 8), st_b = up16((size_t)n_cand * 4);
  const size_t in_b = 2 * q_b + c_b, out_b = os_b + or_b + st_b;
  int rc = ensure_stage(db, in_b + out_b, std::max(in_b, out_b));
  if (rc) return rc;
  uint8_t* h = static_cast<uint8_t*>(db->h_stage);
  uint8_t* d = static_cast<uint8_t*>(db->d_stage);
  memcpy(h, query, (size_t)db->dim * 4);
  if (query2) memcpy(h + q_b, query2, (size_t)db->dim * 4);
  memcpy(h + 2 * q_b, img.data(), (size_t)n_cand * 4);
  cudaStream_t st = db->stream;
  SSW_CUDA(cudaMemcpyAsync(d, h, in_b, cudaMemcpyHostToDevice, st));
  RescoreArgs a{};
  a.vecs = db->d_exact ? static_cast<const void*>(db->d_exact) : db->d_vecs;   // the reference's fp32 values when attached
  a.dtype = db->d_exact ? (int)SSW_F32 : db->dtype;
  a.dim = db->dim;
  a.boxes = db->d_boxes;
  a.zoom = db->d_zoom;
  a.row_ptr = db->d_row_ptr;
  a.orig_row = db->d_orig_row;
  a.q = reinterpret_cast<const float*>(d);
  a.q2 = query2 ? reinterpret_cast<const float*>(d + q_b) : nullptr;
  a.cand_img = reinterpret_cast<const int32_t*>(d + 2 * q_b);
  a.agg = agg_method;
  a.aug = aug_larger;
  a.max_rows = 4096;
  a.out_score = reinterpret_cast<double*>(d + in_b);
  a.out_row = reinterpret_cast<int64_t*>(d + in_b + os_b);
  a.out_status = reinterpret_cast<int32_t*>(d + in_b + os_b + or_b);
  const size_t bes = db->box_kind == SSW_BOX_F64 ? 8 : 4;
  const size_t smem = (size_t)a.max_rows * (4 * bes + 4 + 4 + 4);
  auto launch = [&](auto kern) -> int {
    SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_cand, kRescoreThreads, smem, st>>>(a);
    return SSW_OK;
  };
  if (db->box_kind == SSW_BOX_F64) rc = launch(rescore_kernel<double>);
  else if (db->box_kind == SSW_BOX_F32) rc = launch(rescore_kernel<float>);
  else rc = launch(rescore_kernel<int32_t>);
  if (rc) return rc;
  SSW_LAUNCHED();
  SSW_CUDA(cudaMemcpyAsync(h, d + in_b, out_b, cudaMemcpyDeviceToHost, st));
  SSW_CUDA(cudaStreamSynchronize(st));
  const int32_t* status = reinterpret_cast<const int32_t*>(h + os_b + or_b);
  for (int i = 0; i < n_cand; ++i)
    if (status[i] != 0) {
      set_error("an image has more than 4096 patches: beyond the device rescoring kernel");
      return SSW_ERR_INVALID;
    }
  memcpy(out_score, h, (size_t)n_cand * 8);
  memcpy(out_row, h + os_b, (size_t)n_cand * 8);
  return SSW_OK;
}

}  // extern "C"
