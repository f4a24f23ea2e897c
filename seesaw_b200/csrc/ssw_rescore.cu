// K7: stage 2 of the multiscale query on the device — rescoring of the <= shortlist_size candidate images.
//
// Replaces, for the candidates of stage 1, the host work of MultiscaleIndex.query (seesaw/indices/multiscale/
// multiscale_index.py:341-351: gather all rows of the shortlisted images, scores = vectors @ q [- vectors @ q2])
// and rescore_candidates / score_frame2 (:379-403, :112-150): per image either the first patch attaining the
// maximum score ('plain_score') or, for every patch, the mean over zoom levels of the score of the patch of
// that level it overlaps most (IoU > 0; 'avg_score', the reference's default — basic_types.py:58), then the
// best patch.  One CTA per candidate image; the image's rows are contiguous in HBM (rows are grouped by
// image), boxes and zoom levels sit next to them in device-row order.  IoU and the averages are float64 like
// the host statement (seesaw_b200/rescore.py); patch scores are fp32 dot products.
#include <algorithm>
#include <vector>

#include "ssw_db.h"

namespace ssw {

struct RescoreArgs {
  const void* vecs;            // [n_rows, dim] stored type
  int dtype, dim;
  const int32_t* boxes;        // [n_rows][5] x1,y1,x2,y2,zoom in device-row order (null for plain_score)
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

template <typename T>
__device__ __forceinline__ float row_dot(const T* __restrict__ row, const float* __restrict__ q, int dim, int lane) {
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) {
    float v;
    if constexpr (sizeof(T) == 2) v = __half2float(row[i]); else v = row[i];
    s = fmaf(v, q[i], s);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

__global__ void __launch_bounds__(kRescoreThreads) rescore_kernel(const RescoreArgs a) {
  extern __shared__ __align__(16) uint8_t rsmem[];
  float* s_score = reinterpret_cast<float*>(rsmem);                       // [max_rows]
  int32_t* s_box = reinterpret_cast<int32_t*>(s_score + a.max_rows);       // [max_rows][5]
  double* s_agg = reinterpret_cast<double*>(s_box + 5 * (size_t)a.max_rows + ((a.max_rows & 1) ? 1 : 0) + 2);
  __shared__ int s_zoom[32];
  __shared__ int s_nzoom;
  __shared__ unsigned long long s_best;
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
      s = row_dot(row, a.q, a.dim, lane);
      if (a.q2) s -= row_dot(row, a.q2, a.dim, lane);
    } else {
      const float* row = static_cast<const float*>(a.vecs) + (r0 + p) * (int64_t)a.dim;
      s = row_dot(row, a.q, a.dim, lane);
      if (a.q2) s -= row_dot(row, a.q2, a.dim, lane);
    }
    if (lane == 0) s_score[p] = s;
  }
  if (a.agg == 1)
    for (int i = tid; i < 5 * P; i += kRescoreThreads) s_box[i] = a.boxes[r0 * 5 + i];
  __syncthreads();
  if (a.agg == 1 && tid == 0) {          // distinct zoom levels, ascending (a handful)
    int nz = 0;
    for (int p = 0; p < P; ++p) {
      const int z = s_box[5 * p + 4];
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
  if (tid == 0) s_best = 0ull;
  __syncthreads();
  // ---- aggregated score of every (left) patch
  for (int l = tid; l < P; l += kRescoreThreads) {
    double agg;
    if (a.agg == 0) {
      agg = (double)s_score[l];
    } else {
      const double lx1 = s_box[5 * l], ly1 = s_box[5 * l + 1], lx2 = s_box[5 * l + 2], ly2 = s_box[5 * l + 3];
      const int lz = s_box[5 * l + 4];
      const double larea = (lx2 - lx1) * (ly2 - ly1);
      double total = 0.0;
      int levels = 0;
      for (int zi = 0; zi < s_nzoom; ++zi) {
        const int z = s_zoom[zi];
        if ((a.aug == 1 && z < lz) || (a.aug == 2 && z != lz)) continue;
        double best_iou = 0.0;
        int best_r = -1;
        for (int r = 0; r < P; ++r) {
          if (s_box[5 * r + 4] != z) continue;
          const double rx1 = s_box[5 * r], ry1 = s_box[5 * r + 1], rx2 = s_box[5 * r + 2], ry2 = s_box[5 * r + 3];
          const double w = fmin(lx2, rx2) - fmax(lx1, rx1), h = fmin(ly2, ry2) - fmax(ly1, ry1);
          const double inter = fmax(w, 0.0) * fmax(h, 0.0);
          const double iou = inter / (larea + (rx2 - rx1) * (ry2 - ry1) - inter);
          if (iou > best_iou) {            // strict: the first (lowest position) maximum wins
            best_iou = iou;
            best_r = r;
          }
        }
        if (best_r >= 0) {
          total += (double)s_score[best_r];
          ++levels;
        }
      }
      agg = levels > 0 ? total / levels : -INFINITY;
    }
    s_agg[l] = agg;
  }
  __syncthreads();
  // ---- best patch: max aggregated score, first position on ties (position fits 20 bits: max_rows <= 2^20)
  if (tid == 0) {
    double m = -INFINITY;
    int pos = 0;
    for (int p = 0; p < P; ++p)
      if (s_agg[p] > m) {
        m = s_agg[p];
        pos = p;
      }
    const int64_t drow = r0 + pos;
    a.out_score[c] = m;
    a.out_row[c] = a.orig_row ? a.orig_row[drow] : drow;
    a.out_status[c] = 0;
  }
}

}  // namespace ssw

using namespace ssw;

extern "C" {

int ssw_db_set_boxes(ssw_db* db, const int32_t* x1, const int32_t* y1, const int32_t* x2, const int32_t* y2,
                     const int32_t* zoom) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_REQUIRE(db->n_rows == 0 || (x1 && y1 && x2 && y2 && zoom), "null argument");
  SSW_CUDA(cudaSetDevice(db->device));
  const int64_t n = db->n_rows;
  std::vector<int64_t> perm;
  if (db->d_orig_row) {
    perm.resize(n);
    SSW_CUDA(cudaMemcpy(perm.data(), db->d_orig_row, (size_t)n * 8, cudaMemcpyDeviceToHost));
  }
  std::vector<int32_t> packed((size_t)n * 5);
  for (int64_t r = 0; r < n; ++r) {
    const int64_t o = perm.empty() ? r : perm[r];
    packed[5 * r] = x1[o];
    packed[5 * r + 1] = y1[o];
    packed[5 * r + 2] = x2[o];
    packed[5 * r + 3] = y2[o];
    packed[5 * r + 4] = zoom[o];
  }
  if (db->d_boxes) cudaFree(db->d_boxes);
  db->d_boxes = nullptr;
  SSW_CUDA(cudaMalloc((void**)&db->d_boxes, std::max<size_t>(packed.size() * 4, 16)));
  if (n) SSW_CUDA(cudaMemcpy(db->d_boxes, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
  return SSW_OK;
}

int ssw_rescore(ssw_db* db, const float* query, const float* query2, const int32_t* cand_dbidx, int n_cand, int agg_method,
                int aug_larger, double* out_score, int64_t* out_row) {
  SSW_REQUIRE(db != nullptr && query != nullptr && out_score != nullptr && out_row != nullptr, "null argument");
  SSW_REQUIRE(n_cand >= 0 && (n_cand == 0 || cand_dbidx != nullptr), "bad candidate list");
  SSW_REQUIRE(agg_method == 0 || agg_method == 1, "agg_method: 0 = plain_score, 1 = avg_score");
  SSW_REQUIRE(aug_larger >= 0 && aug_larger <= 2, "aug_larger: 0 = all, 1 = greater, 2 = adjacent");
  SSW_REQUIRE(agg_method == 0 || db->d_boxes != nullptr, "avg_score needs the patch boxes: call ssw_db_set_boxes first");
  if (n_cand == 0) return SSW_OK;
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
  a.vecs = db->d_vecs;
  a.dtype = db->dtype;
  a.dim = db->dim;
  a.boxes = db->d_boxes;
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
  const size_t smem = (size_t)a.max_rows * 4 + ((size_t)a.max_rows * 5 + 3) * 4 + (size_t)a.max_rows * 8 + 16;
  SSW_CUDA(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rescore_kernel<<<n_cand, kRescoreThreads, smem, st>>>(a);
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
