// K6: label propagation over the kNN graph on the device.
//
// Replaces the iteration of LabelPropagation.fit_transform / _step (seesaw/label_propagation.py:30-43, 45-83):
//     new = (W @ old + reg_lambda * reg_values) / (weight_sum + reg_lambda) ; new[label_ids] = label_values
//     stop when max((new - old)^2) < epsilon
// W is the CSR weight matrix get_weight_matrix builds from K3's edge table (seesaw/knn_graph.py:31-104).
// One thread per row walks the row's entries IN INDEX ORDER with separate IEEE multiply and add (no
// FMA contraction), exactly like scipy's csr_matvec, so every iterate is bit-identical to the
// reference's float64 arithmetic.  HBM-bound: nnz * (8 + 4 + 8 gathered) bytes per iteration.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ssw_db.h"

struct ssw_lp {
  int device = 0;
  int64_t n = 0, nnz = 0;
  double reg_lambda = 0.0;
  int64_t* d_indptr = nullptr;
  int32_t* d_indices = nullptr;
  double* d_data = nullptr;
  double* d_wsum = nullptr;     // weight_matrix.sum(0), computed by the caller like the reference does
  double* d_reg = nullptr;
  double* d_x[2] = {nullptr, nullptr};
  int32_t* d_slot = nullptr;    // label slot of every vertex, -1 = unlabeled
  double* d_label_values = nullptr;
  int64_t label_capacity = 0;
  unsigned long long* d_diff = nullptr;
  cudaStream_t stream = nullptr;
};

namespace ssw {

__global__ void lp_step_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                               const double* __restrict__ data, const double* __restrict__ x_old,
                               const double* __restrict__ lreg, const double* __restrict__ wsum, double lambda,
                               const int32_t* __restrict__ slot, const double* __restrict__ label_values,
                               double* __restrict__ x_new, unsigned long long* diff_max) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (i < n) {
    double y = 0.0;
    for (int64_t j = indptr[i]; j < indptr[i + 1]; ++j) y = __dadd_rn(y, __dmul_rn(data[j], x_old[indices[j]]));
    const double w = __dadd_rn(y, lreg[i]);        // lreg = reg_lambda * reg_values, multiplied by the caller
    double v = __ddiv_rn(w, __dadd_rn(wsum[i], lambda));
    const int32_t s = slot[i];
    if (s >= 0) v = label_values[s];
    x_new[i] = v;
    const double d = __dsub_rn(v, x_old[i]);
    d2 = __dmul_rn(d, d);
  }
  // block maximum of the squared change (non-negative doubles order like their bit patterns)
  unsigned long long b = (unsigned long long)__double_as_longlong(d2);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const unsigned long long o = shfl_xor_u64(b, m);
    b = o > b ? o : b;
  }
  __shared__ unsigned long long s_max[8];
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = s_max[w] > b ? s_max[w] : b;
    if (b) atomicMax(diff_max, b);
  }
}

}  // namespace ssw

using namespace ssw;

extern "C" {

int ssw_lp_destroy(ssw_lp* lp) {
  if (!lp) return SSW_OK;
  cudaSetDevice(lp->device);
  if (lp->stream) cudaStreamSynchronize(lp->stream);
  cudaFree(lp->d_indptr);
  cudaFree(lp->d_indices);
  cudaFree(lp->d_data);
  cudaFree(lp->d_wsum);
  cudaFree(lp->d_reg);
  cudaFree(lp->d_x[0]);
  cudaFree(lp->d_x[1]);
  cudaFree(lp->d_slot);
  cudaFree(lp->d_label_values);
  cudaFree(lp->d_diff);
  if (lp->stream) cudaStreamDestroy(lp->stream);
  delete lp;
  return SSW_OK;
}

int ssw_lp_create(ssw_lp** out, int device, int64_t n, const int64_t* indptr, const int32_t* indices,
                  const double* data, const double* weight_sum, double reg_lambda) {
  SSW_REQUIRE(out != nullptr, "out handle is null");
  *out = nullptr;
  SSW_REQUIRE(n > 0 && indptr != nullptr && weight_sum != nullptr, "null argument");
  SSW_REQUIRE(reg_lambda >= 0.0, "reg_lambda must be non-negative");
  const int64_t nnz = indptr[n];
  SSW_REQUIRE(nnz >= 0 && (nnz == 0 || (indices != nullptr && data != nullptr)), "bad CSR arrays");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  ssw_lp* lp = new ssw_lp();
  lp->device = device;
  lp->n = n;
  lp->nnz = nnz;
  lp->reg_lambda = reg_lambda;
  auto fail = [&](cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    ssw_lp_destroy(lp);
    return e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA;
  };
  cudaError_t e;
#define LP_TRY(expr)                      \
  if ((e = (expr)) != cudaSuccess) return fail(e, #expr)
  LP_TRY(cudaStreamCreateWithFlags(&lp->stream, cudaStreamNonBlocking));
  LP_TRY(cudaMalloc((void**)&lp->d_indptr, (size_t)(n + 1) * 8));
  LP_TRY(cudaMalloc((void**)&lp->d_indices, std::max<size_t>((size_t)nnz * 4, 16)));
  LP_TRY(cudaMalloc((void**)&lp->d_data, std::max<size_t>((size_t)nnz * 8, 16)));
  LP_TRY(cudaMalloc((void**)&lp->d_wsum, (size_t)n * 8));
  LP_TRY(cudaMalloc((void**)&lp->d_reg, (size_t)n * 8));
  LP_TRY(cudaMalloc((void**)&lp->d_x[0], (size_t)n * 8));
  LP_TRY(cudaMalloc((void**)&lp->d_x[1], (size_t)n * 8));
  LP_TRY(cudaMalloc((void**)&lp->d_slot, (size_t)n * 4));
  LP_TRY(cudaMalloc((void**)&lp->d_diff, 8));
  LP_TRY(cudaMemcpy(lp->d_indptr, indptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice));
  if (nnz) {
    LP_TRY(cudaMemcpy(lp->d_indices, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    LP_TRY(cudaMemcpy(lp->d_data, data, (size_t)nnz * 8, cudaMemcpyHostToDevice));
  }
  LP_TRY(cudaMemcpy(lp->d_wsum, weight_sum, (size_t)n * 8, cudaMemcpyHostToDevice));
#undef LP_TRY
  *out = lp;
  return SSW_OK;
}

int ssw_lp_fit(ssw_lp* lp, const int64_t* label_ids, const double* label_values, int64_t n_labels,
               const double* reg_values, const double* start_value, int max_iter, double epsilon, double* out_values,
               int* out_iterations, int* out_converged) {
  SSW_REQUIRE(lp != nullptr && out_values != nullptr, "null argument");
  SSW_REQUIRE(reg_values != nullptr || lp->reg_lambda == 0.0, "reg_values is required when reg_lambda > 0");
  // host-side setup exactly as fit_transform (label_propagation.py:45-62), float64 throughout
  const int64_t n = lp->n;
  std::vector<double> x0(n, 0.0), lreg(n, 0.0);
  if (reg_values)
    for (int64_t i = 0; i < n; ++i) lreg[i] = lp->reg_lambda * reg_values[i];
  if (start_value) memcpy(x0.data(), start_value, (size_t)n * 8);
  else if (reg_values) memcpy(x0.data(), reg_values, (size_t)n * 8);
  return ssw_lp_fit_scaled(lp, label_ids, label_values, n_labels, lreg.data(), x0.data(), max_iter, epsilon, out_values,
                           out_iterations, out_converged);
}

int ssw_lp_fit_scaled(ssw_lp* lp, const int64_t* label_ids, const double* label_values, int64_t n_labels,
                      const double* lambda_reg_values, const double* x0_in, int max_iter, double epsilon,
                      double* out_values, int* out_iterations, int* out_converged) {
  SSW_REQUIRE(lp != nullptr && out_values != nullptr && x0_in != nullptr, "null argument");
  SSW_REQUIRE(n_labels >= 0 && (n_labels == 0 || (label_ids != nullptr && label_values != nullptr)), "bad labels");
  SSW_REQUIRE(lambda_reg_values != nullptr || lp->reg_lambda == 0.0, "the prior term is required when reg_lambda > 0");
  SSW_REQUIRE(max_iter >= 0, "max_iter must be non-negative");
  SSW_CUDA(cudaSetDevice(lp->device));
  const int64_t n = lp->n;
  std::vector<double> x0(x0_in, x0_in + n), reg(n, 0.0);
  if (lambda_reg_values) memcpy(reg.data(), lambda_reg_values, (size_t)n * 8);
  std::vector<int32_t> slot(n, -1);
  for (int64_t t = 0; t < n_labels; ++t) {
    SSW_REQUIRE(label_ids[t] >= 0 && label_ids[t] < n, "label id out of range");
    slot[label_ids[t]] = (int32_t)t;          // later duplicates win, like numpy's fancy assignment
  }
  for (int64_t i = 0; i < n; ++i)
    if (slot[i] >= 0) x0[i] = label_values[slot[i]];
  if (n_labels > lp->label_capacity) {
    cudaFree(lp->d_label_values);
    lp->d_label_values = nullptr;
    lp->label_capacity = 0;
    SSW_CUDA(cudaMalloc((void**)&lp->d_label_values, (size_t)n_labels * 8));
    lp->label_capacity = n_labels;
  }
  cudaStream_t st = lp->stream;
  SSW_CUDA(cudaMemcpyAsync(lp->d_x[0], x0.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
  SSW_CUDA(cudaMemcpyAsync(lp->d_reg, reg.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
  SSW_CUDA(cudaMemcpyAsync(lp->d_slot, slot.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  if (n_labels) SSW_CUDA(cudaMemcpyAsync(lp->d_label_values, label_values, (size_t)n_labels * 8, cudaMemcpyHostToDevice, st));
  SSW_CUDA(cudaStreamSynchronize(st));
  int cur = 0, it = 0;
  bool converged = false;
  const int grid = (int)((n + 255) / 256);
  for (it = 1; it <= max_iter; ++it) {
    SSW_CUDA(cudaMemsetAsync(lp->d_diff, 0, 8, st));
    lp_step_kernel<<<grid, 256, 0, st>>>(n, lp->d_indptr, lp->d_indices, lp->d_data, lp->d_x[cur], lp->d_reg, lp->d_wsum,
                                         lp->reg_lambda, lp->d_slot, lp->d_label_values, lp->d_x[cur ^ 1], lp->d_diff);
    SSW_LAUNCHED();
    unsigned long long bits = 0;
    SSW_CUDA(cudaMemcpyAsync(&bits, lp->d_diff, 8, cudaMemcpyDeviceToHost, st));
    SSW_CUDA(cudaStreamSynchronize(st));
    double diff;
    memcpy(&diff, &bits, 8);
    if (diff < epsilon) {         // :66-70 — converged: the PREVIOUS iterate is what fit_transform returns
      converged = true;
      break;
    }
    cur ^= 1;
  }
  if (it > max_iter) it = max_iter;
  SSW_CUDA(cudaMemcpy(out_values, lp->d_x[cur], (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (out_iterations) *out_iterations = it;
  if (out_converged) *out_converged = converged ? 1 : 0;
  return SSW_OK;
}

}  // extern "C"
