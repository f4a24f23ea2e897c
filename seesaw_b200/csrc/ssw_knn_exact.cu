// Exact kNN candidates for float32 vectors.
//
// The tensor-core build (ssw_knn.cu) multiplies fp16 roundings of the vectors; the reference multiplies the
// float32 values themselves (seesaw/knn_graph.py:173).  For vectors that are not fp16-representable the rounded
// distances differ by up to ~2 * ||v - fp16(v)|| * ||v||, enough to reorder neighbours near the k-th place.
// Same recipe as the scan's exact mode (ssw_exact.cu):
//   1. the tensor-core kernel proposes kc > k1 candidate columns per row;
//   2. knn_refine_kernel (one warp per row) recomputes d = fl(1 - dot) for the candidates from the float32 rows
//      with the canonical dot product (canon_dot: symmetric in its two rows bit for bit), orders them by (d, column)
//      and writes the first k1 — CERTIFIED when the k1-th distance plus the error bound E stays below the fp16
//      distance of the last candidate: every column outside the candidates has an fp16 distance at least that large
//      and a float32 distance within E of it;
//   3. rows that cannot be certified (near-duplicates packed tighter than E) are recomputed against ALL columns in
//      float32 by knn_exact_rows_kernel, eight rows per CTA pass over V, same arithmetic.
// E = 2 * rho * vmax + vmax^2 * dim * 1.8e-7 + 3e-7  (rho = max_i ||v_i - fp16(v_i)||, vmax = max_i ||v_i||; the
// second term bounds both accumulations, the last the two roundings of 1 - dot).
#include <algorithm>
#include <vector>

#include "ssw_db.h"

namespace ssw {

constexpr int kRefineWarps = 8;

struct KnnRefineArgs {
  const float* v;          // [n, dim] float32
  int64_t n;
  int dim, k1, kc;
  int64_t row_begin, rows;
  const int32_t* cand_idx; // [rows, kc], -1 = empty, ordered by (fp16 distance, column)
  const float* cand_dist;  // [rows, kc] fp16-arithmetic distances
  float err;               // E
  int32_t* out_idx;        // [rows, k1]
  float* out_dist;
  int32_t* fail_rows;      // rows (relative) that could not be certified
  int* n_fail;
};

__device__ __forceinline__ uint64_t dist_key(float d, int32_t col) {     // smaller key = nearer, ties to the lower column
  return ((uint64_t)f32_ordered(d) << 32) | (uint32_t)col;
}

__global__ void __launch_bounds__(kRefineWarps * 32) knn_refine_kernel(const KnnRefineArgs a) {
  extern __shared__ __align__(16) uint8_t kr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_row = reinterpret_cast<float*>(kr_smem) + (size_t)warp * a.dim;                                   // this warp's row
  uint64_t* s_key = reinterpret_cast<uint64_t*>(kr_smem + (size_t)kRefineWarps * a.dim * 4) + warp * 64;     // kc <= 64
  const int64_t r = blockIdx.x * (int64_t)kRefineWarps + warp;
  if (r >= a.rows) return;
  const int64_t i = a.row_begin + r;
  for (int e = lane; e < a.dim; e += 32) s_row[e] = a.v[i * a.dim + e];
  __syncwarp();
  int n_cand = 0;
  for (int c = 0; c < a.kc; ++c) {
    const int32_t j = a.cand_idx[r * a.kc + c];
    uint64_t key = ~0ull;
    if (j >= 0) {
      const float dot = canon_dot<float>(a.v + (int64_t)j * a.dim, s_row, a.dim, lane);
      key = dist_key(__fsub_rn(1.0f, dot), j);
      ++n_cand;
    }
    if (lane == 0) s_key[c] = key;
  }
  for (int c = a.kc + lane; c < 64; c += 32) s_key[c] = ~0ull;
  __syncwarp();
  // rank by counting (keys are distinct: they embed the column); lane l owns slots l and l + 32
  const int want = min(a.k1, n_cand);
  uint64_t kth = 0;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c = lane + 32 * u;
    const uint64_t mine = s_key[c];
    int rank = 0;
    for (int x = 0; x < 64; ++x) rank += s_key[x] < mine;
    if (mine != ~0ull && rank < a.k1) {
      a.out_idx[r * a.k1 + rank] = (int32_t)(mine & 0xFFFFFFFFu);
      a.out_dist[r * a.k1 + rank] = f32_from_ordered((uint32_t)(mine >> 32));
    }
    if (mine != ~0ull && rank == want - 1) kth = mine;
  }
  for (int c = want + lane; c < a.k1; c += 32) {
    a.out_idx[r * a.k1 + c] = -1;
    a.out_dist[r * a.k1 + c] = INFINITY;
  }
  kth |= __shfl_xor_sync(0xffffffffu, kth, 16);      // exactly one lane holds it
  kth |= __shfl_xor_sync(0xffffffffu, kth, 8);
  kth |= __shfl_xor_sync(0xffffffffu, kth, 4);
  kth |= __shfl_xor_sync(0xffffffffu, kth, 2);
  kth |= __shfl_xor_sync(0xffffffffu, kth, 1);
  if (lane == 0) {
    bool ok = true;
    if (n_cand == a.kc && (int64_t)a.kc < a.n) {     // columns outside the candidates exist
      const float d_k = f32_from_ordered((uint32_t)(kth >> 32));
      const float d16_last = a.cand_dist[r * a.kc + a.kc - 1];
      ok = a.kc > a.k1 && (double)d_k + (double)a.err < (double)d16_last;
    }
    if (!ok) a.fail_rows[atomicAdd(a.n_fail, 1)] = (int32_t)r;
  }
}

// ---- step 3: eight uncertified rows per CTA against all columns, float32 canonical arithmetic ----
constexpr int kExactRows = 8;

struct KnnExactArgs {
  const float* v;
  int64_t n;
  int dim, k1;
  int64_t row_begin;
  const int32_t* fail_rows;
  int n_fail;
  int32_t* out_idx;
  float* out_dist;
};

template <int C>     // dim = C * 128
__global__ void __launch_bounds__(256) knn_exact_rows_kernel(const KnnExactArgs a) {
  extern __shared__ __align__(16) uint8_t ke_smem[];
  constexpr int DIM = C * 128;
  float* s_q = reinterpret_cast<float*>(ke_smem);                                        // [8][DIM]
  uint64_t* s_list = reinterpret_cast<uint64_t*>(ke_smem + (size_t)kExactRows * DIM * 4);  // [8 warps][8 rows][k1]
  __shared__ uint64_t s_worst[8][kExactRows];
  __shared__ int s_cnt[8][kExactRows], s_wpos[8][kExactRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = blockIdx.x * kExactRows;
  const int nr = min(kExactRows, a.n_fail - g0);
  for (int e = threadIdx.x; e < kExactRows * DIM; e += 256) {
    const int rr = e / DIM;
    s_q[e] = rr < nr ? a.v[(a.row_begin + a.fail_rows[g0 + rr]) * (int64_t)DIM + (e % DIM)] : 0.f;
  }
  if (lane < kExactRows) {
    s_cnt[warp][lane] = 0;
    s_wpos[warp][lane] = 0;
    s_worst[warp][lane] = ~0ull;
  }
  __syncthreads();
  uint64_t* mylists = s_list + (size_t)warp * kExactRows * a.k1;
  for (int64_t j = warp; j < a.n; j += 8) {
    // this lane's elements of column j: chunk c covers elements (c*32 + lane)*4 .. +3 (canon_dot's layout)
    float x[C][4];
    const float4* src = reinterpret_cast<const float4*>(a.v + j * (int64_t)DIM);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 t = __ldg(src + c * 32 + lane);
      x[c][0] = t.x; x[c][1] = t.y; x[c][2] = t.z; x[c][3] = t.w;
    }
    float mine = 0.f;     // lane rr keeps row rr's distance
#pragma unroll
    for (int rr = 0; rr < kExactRows; ++rr) {
      const float4* q = reinterpret_cast<const float4*>(s_q + rr * DIM);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 t = q[c * 32 + lane];
        s = fmaf(x[c][0], t.x, s);
        s = fmaf(x[c][1], t.y, s);
        s = fmaf(x[c][2], t.z, s);
        s = fmaf(x[c][3], t.w, s);
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
      if (lane == rr) mine = __fsub_rn(1.0f, s);
    }
    if (lane < nr) {      // offer (d, j) to this warp's list of the lane's row: keep the k1 smallest keys
      const uint64_t key = dist_key(mine, (int32_t)j);
      uint64_t* L = mylists + (size_t)lane * a.k1;
      int cnt = s_cnt[warp][lane];
      bool rescan = false;
      if (cnt < a.k1) {
        L[cnt] = key;
        s_cnt[warp][lane] = ++cnt;
        rescan = cnt == a.k1;
      } else if (key < s_worst[warp][lane]) {
        L[s_wpos[warp][lane]] = key;
        rescan = true;
      }
      if (rescan) {
        uint64_t mk = 0;
        int mp = 0;
        for (int s2 = 0; s2 < a.k1; ++s2)
          if (L[s2] >= mk) {
            mk = L[s2];
            mp = s2;
          }
        s_worst[warp][lane] = mk;
        s_wpos[warp][lane] = mp;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // merge the eight warps' lists of every row: warp rr ranks row rr's <= 8*k1 keys by counting
  if (warp < nr) {
    const int rr = warp;
    const int64_t out_row = a.fail_rows[g0 + rr];
    const int total = 8 * a.k1;
    for (int e = lane; e < total; e += 32) {
      const int w = e / a.k1, s2 = e % a.k1;
      if (s2 >= s_cnt[w][rr]) continue;
      const uint64_t mine = s_list[((size_t)w * kExactRows + rr) * a.k1 + s2];
      int rank = 0;
      for (int w2 = 0; w2 < 8; ++w2)
        for (int s3 = 0; s3 < s_cnt[w2][rr]; ++s3) rank += s_list[((size_t)w2 * kExactRows + rr) * a.k1 + s3] < mine;
      if (rank < a.k1) {
        a.out_idx[out_row * a.k1 + rank] = (int32_t)(mine & 0xFFFFFFFFu);
        a.out_dist[out_row * a.k1 + rank] = f32_from_ordered((uint32_t)(mine >> 32));
      }
    }
  }
}

// Refines the candidate table of rows [row_begin, row_begin + rows) in place of (d_out_idx, d_out_dist) [rows, k1].
// *rows_rescanned receives the number of rows that went through step 3.  Synchronises the stream once.
int knn_exact_refine(const float* d_v32, int64_t n, int dim, int k1, int kc, int64_t row_begin, int64_t rows,
                     const int32_t* d_cand_idx, const float* d_cand_dist, double err, int32_t* d_out_idx, float* d_out_dist,
                     cudaStream_t st, int64_t* rows_rescanned) {
  if (rows_rescanned) *rows_rescanned = 0;
  if (rows == 0) return SSW_OK;
  int32_t* d_fail = nullptr;
  int* d_nfail = nullptr;
  SSW_CUDA(cudaMalloc((void**)&d_fail, (size_t)rows * 4));
  SSW_CUDA(cudaMalloc((void**)&d_nfail, 4));
  SSW_CUDA(cudaMemsetAsync(d_nfail, 0, 4, st));
  KnnRefineArgs a{d_v32, n, dim, k1, kc, row_begin, rows, d_cand_idx, d_cand_dist, (float)err, d_out_idx, d_out_dist, d_fail, d_nfail};
  const size_t smem = (size_t)kRefineWarps * dim * 4 + (size_t)kRefineWarps * 64 * 8;
  SSW_CUDA(cudaFuncSetAttribute(knn_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_refine_kernel<<<(int)((rows + kRefineWarps - 1) / kRefineWarps), kRefineWarps * 32, smem, st>>>(a);
  SSW_LAUNCHED();
  int n_fail = 0;
  SSW_CUDA(cudaMemcpyAsync(&n_fail, d_nfail, 4, cudaMemcpyDeviceToHost, st));
  SSW_CUDA(cudaStreamSynchronize(st));
  int rc = SSW_OK;
  if (n_fail > 0) {
    KnnExactArgs e{d_v32, n, dim, k1, row_begin, d_fail, n_fail, d_out_idx, d_out_dist};
    const size_t smem2 = (size_t)kExactRows * dim * 4 + (size_t)8 * kExactRows * k1 * 8;
    const int grid = (n_fail + kExactRows - 1) / kExactRows;
    auto launch = [&](auto kern) -> int {
      SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      kern<<<grid, 256, smem2, st>>>(e);
      SSW_LAUNCHED();
      return SSW_OK;
    };
    switch (dim) {
      case 256: rc = launch(knn_exact_rows_kernel<2>); break;
      case 512: rc = launch(knn_exact_rows_kernel<4>); break;
      case 768: rc = launch(knn_exact_rows_kernel<6>); break;
      default: set_error("exact kNN refinement supports dim 256, 512 or 768"); rc = SSW_ERR_INVALID;
    }
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
      set_error("exact kNN re-scan failed");
      rc = SSW_ERR_CUDA;
    }
  }
  cudaFree(d_fail);
  cudaFree(d_nfail);
  if (rows_rescanned) *rows_rescanned = n_fail;
  return rc;
}

}  // namespace ssw
