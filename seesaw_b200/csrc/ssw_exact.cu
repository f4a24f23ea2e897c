// Exact mode: reference-exact (fp32) results from an fp16 scan.
//
// The reference stores and scans fp32 vectors (seesaw/indices/multiscale/multiscale_tools.py:200,
// multiscale_index.py:170-175).  Streaming them as fp16 halves the HBM bytes of the scan, but rounding the
// database moves every score by up to ||v - fp16(v)|| * ||q||, enough to reorder images near the k-th place.
// With an fp32 copy attached (ssw_db_attach_exact) the scan answers in three steps, all on the device:
//   1. the fp16 scan (K1 / K2) returns kc > k candidate images per query;
//   2. exact_rescore_kernel re-scores every row of those images from the fp32 copy with the canonical
//      dot product (canon_dot, ssw_common.cuh: bit-identical to what the streaming scan computes on fp32
//      rows) and keeps each image's best (score, row);
//   3. exact_finish_kernel ranks the candidates by their fp32 keys and CERTIFIES the first k: every image
//      that is not a candidate has an fp16 score <= c_min (the last candidate's), and an fp32 score within
//      E of its fp16 score, so if the k-th fp32 key's score exceeds c_min + E no image outside the
//      candidates can belong to the top k, ties included.  E = ||q|| * (rho + vmax * dim * 1.8e-7): rho =
//      max_i ||v_i - fp16(v_i)|| (measured at attach time), the second term bounds the fp32 summation error of
//      both scans (worst-case (dim-1) * 2^-24 for the SIMT chain, dim * 2^-23 for the tensor-core accumulate,
//      2^-22 for the hi/lo fp16 split of the query).
// A query that cannot be certified (scores packed tighter than E around the k-th place, e.g. duplicated
// images) is re-scanned by the streaming kernel over the fp32 copy — same arithmetic as step 2, so either
// way the answer is THE fp32 top-k under index tie-breaking.
#include <algorithm>
#include <vector>

#include "ssw_db.h"

namespace ssw {

// ---- max_i ||v_i - fp16(v_i)||^2 and max_i ||v_i||^2 over the fp32 rows: one warp per row ----
__global__ void row_error_stats_kernel(const float* __restrict__ rows, int64_t n_rows, int dim, float* __restrict__ stats2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float m_err = 0.f, m_norm = 0.f;
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const float* row = rows + r * dim;
    double e2 = 0.0, n2 = 0.0;
    for (int i = lane; i < dim; i += 32) {
      const float v = row[i];
      const float d = v - __half2float(__float2half_rn(v));
      e2 += (double)d * d;
      n2 += (double)v * v;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      e2 += __shfl_xor_sync(0xffffffffu, e2, m);
      n2 += __shfl_xor_sync(0xffffffffu, n2, m);
    }
    // round up: these feed an error BOUND
    m_err = fmaxf(m_err, __double2float_ru(e2));
    m_norm = fmaxf(m_norm, __double2float_ru(n2));
  }
  if (lane == 0) {      // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(stats2), __float_as_int(m_err));
    atomicMax(reinterpret_cast<int*>(stats2 + 1), __float_as_int(m_norm));
  }
}

int launch_row_error_stats(const float* d_rows_f32, int64_t n_rows, int dim, float* d_stats2, cudaStream_t st) {
  SSW_CUDA(cudaMemsetAsync(d_stats2, 0, 8, st));
  if (n_rows == 0) return SSW_OK;
  const int grid = (int)std::min<int64_t>((n_rows + 7) / 8, 148 * 8);
  row_error_stats_kernel<<<grid, 256, 0, st>>>(d_rows_f32, n_rows, dim, d_stats2);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ---- step 2: block (c, q) re-scores candidate image c of query q from the fp32 rows ----
struct ExactRescoreArgs {
  const float* exact;        // [n_rows, dim]
  const int64_t* row_ptr;
  const int64_t* orig_row;   // may be null
  const int32_t* img_dbidx;
  int64_t n_images, row_base;
  int dim, kc;
  const float* q;            // [nq, dim]
  const int32_t* cand_dbidx; // [nq, kc], -1 = empty slot
  uint64_t* keys32;          // [nq, kc] best fp32 key of the image (0 = empty)
};

constexpr int kExactThreads = 128;

__global__ void __launch_bounds__(kExactThreads) exact_rescore_kernel(const ExactRescoreArgs a) {
  extern __shared__ __align__(16) float s_q[];
  __shared__ unsigned long long s_best[kExactThreads / 32];
  const int c = blockIdx.x, qi = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t slot = (int64_t)qi * a.kc + c;
  const int32_t id = a.cand_dbidx[slot];
  if (id < 0) {
    if (tid == 0) a.keys32[slot] = 0ull;
    return;
  }
  for (int i = tid; i < a.dim; i += kExactThreads) s_q[i] = a.q[(int64_t)qi * a.dim + i];
  int64_t lo = 0, hi = a.n_images;          // the image of this dbidx
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a.img_dbidx[mid] < id) lo = mid + 1; else hi = mid;
  }
  __syncthreads();
  const int64_t r0 = a.row_ptr[lo], r1 = a.row_ptr[lo + 1];
  uint64_t best = 0;
  for (int64_t r = r0 + warp; r < r1; r += kExactThreads / 32) {
    const float s = canon_dot<float>(a.exact + r * a.dim, s_q, a.dim, lane);
    const int64_t orow = a.orig_row ? a.orig_row[r] : r;
    const uint64_t key = s == s ? make_key(s, (uint32_t)(a.row_base + orow)) : 0ull;   // NaN never ranks
    best = key > best ? key : best;
  }
  if (lane == 0) s_best[warp] = best;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int w = 1; w < kExactThreads / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
    a.keys32[slot] = best;
  }
}

int launch_exact_rescore(ssw_db* db, const float* d_queries, int nq, int kc, const int32_t* d_cand_dbidx,
                         uint64_t* d_keys32, cudaStream_t st) {
  ExactRescoreArgs a{};
  a.exact = db->d_exact;
  a.row_ptr = db->d_row_ptr;
  a.orig_row = db->d_orig_row;
  a.img_dbidx = db->d_img_dbidx;
  a.n_images = db->n_images;
  a.row_base = db->row_base;
  a.dim = db->dim;
  a.kc = kc;
  a.q = d_queries;
  a.cand_dbidx = d_cand_dbidx;
  a.keys32 = d_keys32;
  dim3 grid(kc, nq);
  exact_rescore_kernel<<<grid, kExactThreads, (size_t)db->dim * 4, st>>>(a);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ---- step 3: one block per query ranks the kc fp32 keys, writes the first k and the certificate ----
struct ExactFinishArgs {
  const float* q;
  int dim, kc, k;
  const uint64_t* keys16;     // [nq, kc] the scan's keys, best first, 0 = empty
  const uint64_t* keys32;     // [nq, kc]
  const int32_t* cand_dbidx;  // [nq, kc]
  double err_per_unit_q;      // E = err_per_unit_q * ||q||
  int32_t* out_dbidx;
  float* out_score;
  int64_t* out_row;
  int32_t* out_count;
  uint64_t* out_key;
  int32_t* certified;
};

constexpr int kFinishThreads = 256;

__global__ void __launch_bounds__(kFinishThreads) exact_finish_kernel(const ExactFinishArgs a) {
  extern __shared__ __align__(16) uint8_t fsmem[];
  uint64_t* s_key = reinterpret_cast<uint64_t*>(fsmem);                 // [kc]
  uint64_t* s_top = s_key + a.kc;                                        // [k] ranked
  __shared__ double s_red[kFinishThreads / 32];
  __shared__ int s_valid;
  const int qi = blockIdx.x, tid = threadIdx.x;
  const uint64_t* k16 = a.keys16 + (int64_t)qi * a.kc;
  const uint64_t* k32 = a.keys32 + (int64_t)qi * a.kc;
  if (tid == 0) s_valid = 0;
  for (int i = tid; i < a.k; i += kFinishThreads) s_top[i] = 0ull;
  __syncthreads();
  int local_valid = 0;
  for (int i = tid; i < a.kc; i += kFinishThreads) {
    s_key[i] = k32[i];
    local_valid += k16[i] != 0ull;
  }
  if (local_valid) atomicAdd(&s_valid, local_valid);
  double n2 = 0.0;
  for (int i = tid; i < a.dim; i += kFinishThreads) {
    const double x = a.q[(int64_t)qi * a.dim + i];
    n2 += x * x;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  if ((tid & 31) == 0) s_red[tid >> 5] = n2;
  __syncthreads();
  // rank by counting: keys embed distinct rows, so the rank is the output slot
  for (int i = tid; i < a.kc; i += kFinishThreads) {
    const uint64_t mine = s_key[i];
    if (mine == 0ull) continue;
    int rank = 0;
    for (int j = 0; j < a.kc; ++j) rank += s_key[j] > mine;
    if (rank < a.k) {
      s_top[rank] = mine;
      const int64_t o = (int64_t)qi * a.k + rank;
      if (a.out_key) a.out_key[o] = mine;
      if (a.out_dbidx) a.out_dbidx[o] = a.cand_dbidx[(int64_t)qi * a.kc + i];
      if (a.out_score) a.out_score[o] = key_score(mine);
      if (a.out_row) a.out_row[o] = (int64_t)key_row(mine);
    }
  }
  __syncthreads();
  const int n_valid = s_valid;
  const int m = min(n_valid, a.k);
  for (int i = m + tid; i < a.k; i += kFinishThreads) {
    const int64_t o = (int64_t)qi * a.k + i;
    if (a.out_key) a.out_key[o] = 0ull;
    if (a.out_dbidx) a.out_dbidx[o] = -1;
    if (a.out_score) a.out_score[o] = -INFINITY;
    if (a.out_row) a.out_row[o] = -1;
  }
  if (tid == 0) {
    if (a.out_count) a.out_count[qi] = m;
    int ok = 1;
    if (n_valid == a.kc) {        // the scan filled every candidate slot: images outside it exist (or may)
      double qn2 = 0.0;
      for (int w = 0; w < kFinishThreads / 32; ++w) qn2 += s_red[w];
      const double E = a.err_per_unit_q * sqrt(qn2) * (1.0 + 1e-6);
      const double c_min = (double)key_score(k16[a.kc - 1]);
      const double t = (double)key_score(s_top[a.k - 1]);
      ok = (t - c_min > E) ? 1 : 0;
      if (a.kc <= a.k) ok = 0;   // no margin at all
    }
    a.certified[qi] = ok;
  }
}

int launch_exact_finish(const float* d_queries, int dim, int nq, int kc, int k, const uint64_t* d_keys16,
                        const uint64_t* d_keys32, const int32_t* d_cand_dbidx, double err_per_unit_q,
                        int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                        uint64_t* d_out_key, int32_t* d_certified, cudaStream_t st) {
  ExactFinishArgs a{d_queries, dim, kc, k, d_keys16, d_keys32, d_cand_dbidx, err_per_unit_q,
                    d_out_dbidx, d_out_score, d_out_row, d_out_count, d_out_key, d_certified};
  exact_finish_kernel<<<nq, kFinishThreads, (size_t)(kc + k) * 8, st>>>(a);
  SSW_LAUNCHED();
  return SSW_OK;
}

}  // namespace ssw
