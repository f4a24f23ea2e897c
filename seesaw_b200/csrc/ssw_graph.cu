// Symmetric weight matrix of a kNN edge table on the device — the input of label propagation.
//
// Replaces get_weight_matrix(df, kfun=, self_edges=False, normalized=False, symmetric=True) (seesaw/knn_graph.py:31-104)
// as KnnProp2 calls it (seesaw/loops/graph_based.py:36-43) for the edge table of compute_exact_knn:
//   * structure  = every listed edge (i, j) and its mirror (j, i)                      (adjacency_m.T + adjacency_m, :41-42)
//   * value(i,j) = (w_ji [if listed and > 0] + w_ij [if listed and > 0]) / #listings   (:55-63; w = kfun(distance) from the
//                  CALLER, so the transcendental is the reference's own numpy exp, not a device approximation)
//   * the diagonal is kept as explicit zeros (setdiag(0.) on stored entries, :72)
//   * CSR with sorted column indices (:98-100); every vertex must keep a positive degree (:77).
// weight_sum = W.sum(0) as LabelPropagation computes it (label_propagation.py:24): for this symmetric matrix the
// column sum accumulates rows in ascending order, i.e. it is the sequential float64 sum of the vertex's own row.
// The edge table must be sorted by src_vertex (it is: post_process_graph_df sorts by (src, rank)), hold one self
// edge per vertex and at most one edge per ordered pair.
#include <algorithm>
#include <vector>

#include "ssw_db.h"

namespace ssw {

constexpr int kWmBlock = 256;

__global__ void wm_rowptr_kernel(const int32_t* __restrict__ src, int64_t n_edges, int64_t n, int64_t* __restrict__ eptr, int* __restrict__ bad) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = src[e];
    if (s < 0 || s >= n || (e > 0 && src[e - 1] > s)) {
      *bad = 1;                       // out of range or not sorted by source
      continue;
    }
    if (e == 0 || src[e - 1] != s) eptr[s] = e;
    if (e == n_edges - 1) eptr[n] = n_edges;
  }
}

__global__ void wm_check_rowptr_kernel(const int64_t* __restrict__ eptr, int64_t n, int* __restrict__ bad) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v <= n; v += (int64_t)gridDim.x * blockDim.x)
    if (eptr[v] < 0 || (v > 0 && eptr[v] <= eptr[v - 1])) *bad = 1;      // a vertex without edges (no self edge)
}

// position of edge (a -> b) in a's list, or -1
__device__ __forceinline__ int64_t wm_find(const int64_t* eptr, const int32_t* dst, int32_t a, int32_t b) {
  for (int64_t e = eptr[a]; e < eptr[a + 1]; ++e)
    if (dst[e] == b) return e;
  return -1;
}

// one thread per listed edge (i -> j): an edge whose mirror is not listed adds one entry to row j
__global__ void wm_count_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n_edges, int64_t n,
                                const int64_t* __restrict__ eptr, int32_t* __restrict__ extra, int* __restrict__ bad) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = src[e], j = dst[e];
    if (j < 0 || j >= n) {
      *bad = 1;
      continue;
    }
    if (i != j && wm_find(eptr, dst, j, i) < 0) atomicAdd(extra + j, 1);
  }
}

// row sizes -> CSR row pointer: block-local inclusive scan, then a serial pass over the block totals (n / 256 values)
__global__ void wm_scan_local_kernel(const int64_t* __restrict__ eptr, const int32_t* __restrict__ extra, int64_t n,
                                     int64_t* __restrict__ indptr, int64_t* __restrict__ block_total) {
  __shared__ int64_t s[kWmBlock];
  const int64_t v = blockIdx.x * (int64_t)kWmBlock + threadIdx.x;
  int64_t c = v < n ? (eptr[v + 1] - eptr[v]) + extra[v] : 0;
  s[threadIdx.x] = c;
  __syncthreads();
  for (int off = 1; off < kWmBlock; off <<= 1) {
    const int64_t t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (v < n) indptr[v] = s[threadIdx.x] - c;      // exclusive, block-local
  if (threadIdx.x == kWmBlock - 1) block_total[blockIdx.x] = s[threadIdx.x];
}

__global__ void wm_scan_totals_kernel(int64_t* block_total, int64_t n_blocks, int64_t* indptr, int64_t n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t run = 0;
    for (int64_t b = 0; b < n_blocks; ++b) {
      const int64_t t = block_total[b];
      block_total[b] = run;
      run += t;
    }
    indptr[n] = run;
  }
}

__global__ void wm_add_base_kernel(int64_t* __restrict__ indptr, const int64_t* __restrict__ block_base, int64_t n) {
  const int64_t v = blockIdx.x * (int64_t)kWmBlock + threadIdx.x;
  if (v < n) indptr[v] += block_base[blockIdx.x];
}

// one thread per listed edge: its own entry in row i, and — when the mirror is not listed — the mirror entry in row j
__global__ void wm_fill_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const double* __restrict__ w,
                               int64_t n_edges, const int64_t* __restrict__ eptr, const int64_t* __restrict__ indptr,
                               int32_t* __restrict__ cursor, int32_t* __restrict__ indices, double* __restrict__ data) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = src[e], j = dst[e];
    const double wij = w[e] > 0.0 ? w[e] : 0.0;               // zero weights are masked out of the value (:49-53), not of the structure
    const int64_t at = indptr[i] + (e - eptr[i]);
    indices[at] = j;
    if (i == j) {
      data[at] = 0.0;                                          // setdiag(0.)
      continue;
    }
    const int64_t r = wm_find(eptr, dst, j, i);
    if (r >= 0) {
      const double wji = w[r] > 0.0 ? w[r] : 0.0;
      data[at] = __ddiv_rn(__dadd_rn(wji, wij), 2.0);          // (weight_mat.T + weight_mat)[i, j] / 2 listings
    } else {
      data[at] = wij;                                          // / 1 listing
      const int64_t m = indptr[j] + (eptr[j + 1] - eptr[j]) + atomicAdd(cursor + j, 1);
      indices[m] = i;
      data[m] = wij;
    }
  }
}

// one thread per row: order the entries by column (rows hold tens of entries), then the row sum in that order
__global__ void wm_sort_rows_kernel(const int64_t* __restrict__ indptr, int64_t n, int32_t* __restrict__ indices,
                                    double* __restrict__ data, double* __restrict__ row_sum, int* __restrict__ bad) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = indptr[v], b = indptr[v + 1];
    for (int64_t x = a + 1; x < b; ++x) {
      const int32_t c = indices[x];
      const double d = data[x];
      int64_t y = x - 1;
      while (y >= a && indices[y] > c) {
        indices[y + 1] = indices[y];
        data[y + 1] = data[y];
        --y;
      }
      indices[y + 1] = c;
      data[y + 1] = d;
    }
    double s = 0.0;
    for (int64_t x = a; x < b; ++x) {
      s = __dadd_rn(s, data[x]);
      if (x > a && indices[x] == indices[x - 1]) *bad = 2;    // an ordered pair listed twice
    }
    if (row_sum) row_sum[v] = s;
    if (!(s > 0.0)) *bad = 3;                                  // 'no zero degree nodes allowed' (:77)
  }
}

}  // namespace ssw

using namespace ssw;

extern "C" {

int ssw_weight_matrix(int device, const int32_t* src, const int32_t* dst, const double* weight, int64_t n_edges, int64_t n,
                      int64_t* out_indptr, int32_t* out_indices, double* out_data, int64_t capacity, int64_t* out_nnz,
                      double* out_weight_sum) {
  SSW_REQUIRE(src && dst && weight && out_indptr && out_indices && out_data && out_nnz, "null argument");
  SSW_REQUIRE(n > 0 && n_edges >= n && n < (int64_t)0x7FFFFFFF, "bad shape (one self edge per vertex is required)");
  SSW_REQUIRE(capacity >= 2 * n_edges, "output capacity must be 2 * n_edges entries");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  int32_t *d_src = nullptr, *d_dst = nullptr, *d_extra = nullptr, *d_cursor = nullptr, *d_indices = nullptr;
  double *d_w = nullptr, *d_data = nullptr, *d_sum = nullptr;
  int64_t *d_eptr = nullptr, *d_indptr = nullptr, *d_btot = nullptr;
  int* d_bad = nullptr;
  const int64_t n_blocks = (n + kWmBlock - 1) / kWmBlock;
  auto cleanup = [&]() {
    for (void* p : {(void*)d_src, (void*)d_dst, (void*)d_extra, (void*)d_cursor, (void*)d_indices, (void*)d_w, (void*)d_data,
                    (void*)d_sum, (void*)d_eptr, (void*)d_indptr, (void*)d_btot, (void*)d_bad})
      cudaFree(p);
  };
  auto chk = [&](cudaError_t e, const char* what) -> int {
    if (e == cudaSuccess) return SSW_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cleanup();
    return e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA;
  };
#define WM_TRY(expr) \
  if ((rc = chk((expr), #expr))) return rc
  WM_TRY(cudaMalloc((void**)&d_src, (size_t)n_edges * 4));
  WM_TRY(cudaMalloc((void**)&d_dst, (size_t)n_edges * 4));
  WM_TRY(cudaMalloc((void**)&d_w, (size_t)n_edges * 8));
  WM_TRY(cudaMalloc((void**)&d_eptr, (size_t)(n + 1) * 8));
  WM_TRY(cudaMalloc((void**)&d_indptr, (size_t)(n + 1) * 8));
  WM_TRY(cudaMalloc((void**)&d_btot, (size_t)n_blocks * 8));
  WM_TRY(cudaMalloc((void**)&d_extra, (size_t)n * 4));
  WM_TRY(cudaMalloc((void**)&d_cursor, (size_t)n * 4));
  WM_TRY(cudaMalloc((void**)&d_indices, (size_t)2 * n_edges * 4));
  WM_TRY(cudaMalloc((void**)&d_data, (size_t)2 * n_edges * 8));
  WM_TRY(cudaMalloc((void**)&d_sum, (size_t)n * 8));
  WM_TRY(cudaMalloc((void**)&d_bad, 4));
  WM_TRY(cudaMemcpy(d_src, src, (size_t)n_edges * 4, cudaMemcpyHostToDevice));
  WM_TRY(cudaMemcpy(d_dst, dst, (size_t)n_edges * 4, cudaMemcpyHostToDevice));
  WM_TRY(cudaMemcpy(d_w, weight, (size_t)n_edges * 8, cudaMemcpyHostToDevice));
  WM_TRY(cudaMemset(d_extra, 0, (size_t)n * 4));
  WM_TRY(cudaMemset(d_cursor, 0, (size_t)n * 4));
  WM_TRY(cudaMemset(d_bad, 0, 4));
  WM_TRY(cudaMemset(d_eptr, 0xFF, (size_t)(n + 1) * 8));
  const int grid_e = (int)std::min<int64_t>((n_edges + kWmBlock - 1) / kWmBlock, 148 * 32);
  const int grid_v = (int)std::min<int64_t>(n_blocks, 148 * 32);
  auto launched = [&]() -> int {
    ++g_launch_count;
    return chk(cudaGetLastError(), "kernel launch");
  };
  wm_rowptr_kernel<<<grid_e, kWmBlock>>>(d_src, n_edges, n, d_eptr, d_bad);
  if ((rc = launched())) return rc;
  wm_check_rowptr_kernel<<<grid_v, kWmBlock>>>(d_eptr, n, d_bad);
  if ((rc = launched())) return rc;
  int bad = 0;
  WM_TRY(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  std::vector<int64_t> probe(2);
  WM_TRY(cudaMemcpy(probe.data(), d_eptr, 8, cudaMemcpyDeviceToHost));
  if (bad || probe[0] != 0) {
    cleanup();
    set_error("edge table must be sorted by src_vertex, vertices numbered 0 .. n-1, every vertex with its self edge");
    return SSW_ERR_INVALID;
  }
  wm_count_kernel<<<grid_e, kWmBlock>>>(d_src, d_dst, n_edges, n, d_eptr, d_extra, d_bad);
  if ((rc = launched())) return rc;
  wm_scan_local_kernel<<<(int)n_blocks, kWmBlock>>>(d_eptr, d_extra, n, d_indptr, d_btot);
  if ((rc = launched())) return rc;
  wm_scan_totals_kernel<<<1, 32>>>(d_btot, n_blocks, d_indptr, n);
  if ((rc = launched())) return rc;
  wm_add_base_kernel<<<(int)n_blocks, kWmBlock>>>(d_indptr, d_btot, n);
  if ((rc = launched())) return rc;
  wm_fill_kernel<<<grid_e, kWmBlock>>>(d_src, d_dst, d_w, n_edges, d_eptr, d_indptr, d_cursor, d_indices, d_data);
  if ((rc = launched())) return rc;
  wm_sort_rows_kernel<<<grid_v, kWmBlock>>>(d_indptr, n, d_indices, d_data, d_sum, d_bad);
  if ((rc = launched())) return rc;
  WM_TRY(cudaDeviceSynchronize());
  WM_TRY(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  if (bad) {
    cleanup();
    set_error(bad == 3 ? "no zero degree nodes allowed" : bad == 2 ? "an ordered vertex pair is listed more than once" : "vertex id out of range");
    return SSW_ERR_INVALID;
  }
  WM_TRY(cudaMemcpy(out_indptr, d_indptr, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
  const int64_t nnz = out_indptr[n];
  WM_TRY(cudaMemcpy(out_indices, d_indices, (size_t)nnz * 4, cudaMemcpyDeviceToHost));
  WM_TRY(cudaMemcpy(out_data, d_data, (size_t)nnz * 8, cudaMemcpyDeviceToHost));
  if (out_weight_sum) WM_TRY(cudaMemcpy(out_weight_sum, d_sum, (size_t)n * 8, cudaMemcpyDeviceToHost));
#undef WM_TRY
  *out_nnz = nnz;
  cleanup();
  return SSW_OK;
}

}  // extern "C"
