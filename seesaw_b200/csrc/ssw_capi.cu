// C ABI of libseesaw_b200: handle management, uploads, and the host-pointer wrappers around the
// device entry points.  See include/seesaw_b200.h for the contract and the reference citations.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "ssw_db.h"

namespace ssw {

static thread_local std::string t_error;
int64_t g_launch_count = 0;

void set_error(const std::string& msg) { t_error = msg; }
bool debug_sync() {
  static const bool on = [] {
    const char* e = getenv("SSW_DEBUG_SYNC");
    return e && e[0] == '1';
  }();
  return on;
}

// cudaGetDeviceProperties costs milliseconds per call; the answer per device never changes.
int ensure_device(int device, int* sm_count) {
  constexpr int kMaxDev = 64;
  static int n_dev = -1;
  static int sm_of[kMaxDev];          // 0 = not probed yet, -1 = not sm_100
  if (n_dev < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      set_error("no CUDA device available: seesaw_b200 has no CPU path");
      return SSW_ERR_NO_DEVICE;
    }
    n_dev = n < kMaxDev ? n : kMaxDev;
  }
  if (device < 0 || device >= n_dev) {
    set_error("device index out of range");
    return SSW_ERR_INVALID;
  }
  if (sm_of[device] == 0) {
    int major = 0, sms = 0;
    SSW_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    SSW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    sm_of[device] = major == 10 ? sms : -1;
  }
  if (sm_of[device] < 0) {
    set_error("device " + std::to_string(device) + " is not sm_100 (Blackwell B200); kernels are sm_100a only");
    return SSW_ERR_NO_DEVICE;
  }
  SSW_CUDA(cudaSetDevice(device));
  if (sm_count) *sm_count = sm_of[device];
  return SSW_OK;
}

void prof_begin(ssw_db* db, cudaStream_t st) {
  db->prof_sampled = false;
  if (!db->profiling) return;
  if (db->prof_stride > 1 && (db->prof_counter++ % db->prof_stride) != 0) return;   // sample every prof_stride-th launch
  db->prof_sampled = true;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!db->prof_free.empty()) {
    ev = db->prof_free.back();
    db->prof_free.pop_back();
  } else {
    cudaEventCreate(&ev.first);
    cudaEventCreate(&ev.second);
  }
  cudaEventRecord(ev.first, st);
  db->prof_pending.push_back(ev);
}
void prof_end(ssw_db* db, cudaStream_t st) {
  if (!db->prof_sampled || db->prof_pending.empty()) return;
  cudaEventRecord(db->prof_pending.back().second, st);
}

template <typename T>
static int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  SSW_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return SSW_OK;
}

static int ensure_lists(ssw_db* db, int nq, int lists, int k) {
  const size_t need = (size_t)nq * lists * k;
  if (need > db->list_capacity) {
    if (db->d_list_keys) cudaFree(db->d_list_keys);
    if (db->d_list_dbidx) cudaFree(db->d_list_dbidx);
    db->list_capacity = 0;
    int rc = dev_alloc(&db->d_list_keys, need);
    if (rc) return rc;
    rc = dev_alloc(&db->d_list_dbidx, need);
    if (rc) return rc;
    db->list_capacity = need;
  }
  if (nq > db->gthr_capacity) {
    if (db->d_gthr) cudaFree(db->d_gthr);
    db->gthr_capacity = 0;
    // one allocation: nq thresholds (8 B), nq x grid published bests (4 B) — zeroed together per call of the
    // streaming scan — then nq candidate counters of the batched scan (zeroed by its preparation kernel)
    uint8_t* p = nullptr;
    SSW_CUDA(cudaMalloc((void**)&p, (size_t)nq * 8 + (size_t)nq * db->scan_grid * 4 + (size_t)nq * 4));
    db->d_gthr = reinterpret_cast<uint64_t*>(p);
    db->d_pub1 = reinterpret_cast<uint32_t*>(p + (size_t)nq * 8);
    db->d_cand_cnt = reinterpret_cast<int32_t*>(p + (size_t)nq * 8 + (size_t)nq * db->scan_grid * 4);
    db->gthr_capacity = nq;
  }
  return SSW_OK;
}

// the pipelined sharded step alternates between two sets of scan workspaces: the exchange of step i still reads
// one set while the scan of step i+1 fills the other
static void swap_scan_ws(ssw_db* db) {
  std::swap(db->d_list_keys, db->ws_alt.keys);
  std::swap(db->d_list_dbidx, db->ws_alt.dbidx);
  std::swap(db->d_gthr, db->ws_alt.gthr);
  std::swap(db->d_pub1, db->ws_alt.pub1);
  std::swap(db->d_cand_cnt, db->ws_alt.cand_cnt);
  std::swap(db->list_capacity, db->ws_alt.list_capacity);
  std::swap(db->gthr_capacity, db->ws_alt.gthr_capacity);
}

int ensure_stage(ssw_db* db, size_t dev_bytes, size_t host_bytes) {
  if (dev_bytes > db->d_stage_bytes) {
    if (db->d_stage) cudaFree(db->d_stage);
    db->d_stage_bytes = 0;
    SSW_CUDA(cudaMalloc(&db->d_stage, dev_bytes));
    db->d_stage_bytes = dev_bytes;
  }
  if (host_bytes > db->h_stage_bytes) {
    if (db->h_stage) cudaFreeHost(db->h_stage);
    db->h_stage_bytes = 0;
    SSW_CUDA(cudaMallocHost(&db->h_stage, host_bytes));
    db->h_stage_bytes = host_bytes;
  }
  return SSW_OK;
}

// Builds the image CSR, the (optional) permutation and the scan partition; uploads the metadata.
static int build_layout(ssw_db* db, const int32_t* dbidx_per_row, std::vector<int64_t>* perm_out) {
  const int64_t n = db->n_rows;
  bool sorted = true;
  for (int64_t i = 1; i < n; ++i)
    if (dbidx_per_row[i] < dbidx_per_row[i - 1]) {
      sorted = false;
      break;
    }
  std::vector<int64_t> perm;   // device row -> original row
  if (!sorted) {
    perm.resize(n);
    std::iota(perm.begin(), perm.end(), (int64_t)0);
    std::stable_sort(perm.begin(), perm.end(),
                     [&](int64_t x, int64_t y) { return dbidx_per_row[x] < dbidx_per_row[y]; });
  }
  std::vector<int32_t> img_dbidx;
  std::vector<int64_t> row_ptr;
  int32_t prev = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t d = dbidx_per_row[sorted ? i : perm[i]];
    if (i == 0 || d != prev) {
      img_dbidx.push_back(d);
      row_ptr.push_back(i);
      prev = d;
    }
  }
  row_ptr.push_back(n);
  // one bit per device row: last row of its image (read by the batched scan instead of the 4-byte ids)
  std::vector<uint32_t> last_bits((size_t)(n / 32) + 8, 0u);
  for (size_t i = 1; i < row_ptr.size(); ++i) {
    const int64_t r = row_ptr[i] - 1;
    if (r >= 0) last_bits[r >> 5] |= 1u << (r & 31);
  }
  db->n_images = (int64_t)img_dbidx.size();
  db->excl_words = ((db->n_images + 31) / 32 + 3) / 4 * 4;
  if (db->excl_words == 0) db->excl_words = 4;

  // scan partition: warp g gets images [part[g], part[g+1]) with ~n/G rows each
  const int G = db->scan_grid * kScanWarps;
  std::vector<int32_t> part(G + 1);
  for (int g = 0; g <= G; ++g) {
    const int64_t target = (int64_t)((__int128)n * g / G);
    part[g] = (int32_t)(std::lower_bound(row_ptr.begin(), row_ptr.end(), target) - row_ptr.begin());
    if (part[g] > db->n_images) part[g] = (int32_t)db->n_images;
  }
  part[0] = 0;
  part[G] = (int32_t)db->n_images;
  db->max_cta_images = 0;
  for (int c = 0; c < db->scan_grid; ++c)
    db->max_cta_images = std::max<int64_t>(db->max_cta_images, part[(c + 1) * kScanWarps] - part[c * kScanWarps]);

  int rc;
  if ((rc = dev_alloc(&db->d_row_ptr, row_ptr.size()))) return rc;
  if ((rc = dev_alloc(&db->d_img_dbidx, img_dbidx.size()))) return rc;
  if ((rc = dev_alloc(&db->d_part, part.size()))) return rc;
  if ((rc = dev_alloc(&db->d_last_bits, last_bits.size()))) return rc;
  SSW_CUDA(cudaMemcpy(db->d_last_bits, last_bits.data(), last_bits.size() * 4, cudaMemcpyHostToDevice));
  SSW_CUDA(cudaMemcpy(db->d_row_ptr, row_ptr.data(), row_ptr.size() * 8, cudaMemcpyHostToDevice));
  if (!img_dbidx.empty())
    SSW_CUDA(cudaMemcpy(db->d_img_dbidx, img_dbidx.data(), img_dbidx.size() * 4, cudaMemcpyHostToDevice));
  SSW_CUDA(cudaMemcpy(db->d_part, part.data(), part.size() * 4, cudaMemcpyHostToDevice));
  if (!sorted) {
    if ((rc = dev_alloc(&db->d_orig_row, (size_t)n))) return rc;
    SSW_CUDA(cudaMemcpy(db->d_orig_row, perm.data(), n * 8, cudaMemcpyHostToDevice));
  }
  if (perm_out) *perm_out = std::move(perm);
  return SSW_OK;
}

// Host rows [n_rows, dim] of dtype_in in ORIGINAL order -> d_dst in the database's grouped order as dtype_dst.
// perm: device row -> original row (empty = identity).
static int upload_rows(ssw_db* db, const void* vectors, int dtype_in, void* d_dst, int dtype_dst,
                       const std::vector<int64_t>& perm) {
  const int64_t n_rows = db->n_rows;
  const int dim = db->dim;
  const size_t es_in = dtype_in == SSW_F16 ? 2 : 4, es_st = dtype_dst == SSW_F16 ? 2 : 4;
  const size_t row_in = (size_t)dim * es_in;
  const bool identity = perm.empty();
  if (identity && dtype_in == dtype_dst) {
    SSW_CUDA(cudaMemcpy(d_dst, vectors, (size_t)n_rows * row_in, cudaMemcpyHostToDevice));
    return SSW_OK;
  }
  // chunked: gather rows on the host into pinned staging, copy, convert/place on the device
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(32u << 20) / (int64_t)row_in);
  int rc = ensure_stage(db, (size_t)chunk_rows * row_in, (size_t)chunk_rows * row_in);
  if (rc) return rc;
  const uint8_t* src = static_cast<const uint8_t*>(vectors);
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk_rows) {
    const int64_t nr = std::min(chunk_rows, n_rows - r0);
    uint8_t* h = static_cast<uint8_t*>(db->h_stage);
    if (identity) {
      memcpy(h, src + (size_t)r0 * row_in, (size_t)nr * row_in);
    } else {
      for (int64_t i = 0; i < nr; ++i) memcpy(h + (size_t)i * row_in, src + (size_t)perm[r0 + i] * row_in, row_in);
    }
    SSW_CUDA(cudaMemcpyAsync(db->d_stage, h, (size_t)nr * row_in, cudaMemcpyHostToDevice, db->stream));
    if ((rc = launch_convert_rows(db->d_stage, dtype_in, static_cast<uint8_t*>(d_dst) + (size_t)r0 * dim * es_st, dtype_dst,
                                  nr * dim, db->stream)))
      return rc;
    SSW_CUDA(cudaStreamSynchronize(db->stream));
  }
  return SSW_OK;
}

static int db_new(ssw_db** out, int device, int dtype_store, int64_t n_rows, int dim, int64_t row_base) {
  SSW_REQUIRE(out != nullptr, "out handle is null");
  *out = nullptr;
  SSW_REQUIRE(dtype_store == SSW_F16 || dtype_store == SSW_F32, "dtype_store must be SSW_F32 or SSW_F16");
  SSW_REQUIRE(n_rows >= 0 && n_rows < (int64_t)0xFFFFFFFFll, "n_rows out of range");
  SSW_REQUIRE(row_base >= 0 && row_base + n_rows < (int64_t)0xFFFFFFFFll, "global row index must fit 32 bits");
  SSW_REQUIRE(dim == 256 || dim == 512 || dim == 768 || dim == 1024, "dim must be 256, 512, 768 or 1024");
  int sms = 0;
  int rc = ensure_device(device, &sms);
  if (rc) return rc;
  ssw_db* db = new ssw_db();
  db->device = device;
  db->dim = dim;
  db->dtype = dtype_store;
  db->n_rows = n_rows;
  db->row_base = row_base;
  db->sm_count = sms;
  db->scan_grid = sms;
  cudaError_t e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete db;
    set_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    return SSW_ERR_CUDA;
  }
  *out = db;
  return SSW_OK;
}

}  // namespace ssw

using namespace ssw;

extern "C" {

const char* ssw_last_error(void) { return t_error.c_str(); }
int ssw_version(void) { return 100; }
int64_t ssw_kernel_launch_count(void) { return g_launch_count; }

int ssw_device_count(int* out_count) {
  SSW_REQUIRE(out_count != nullptr, "out_count is null");
  *out_count = 0;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return SSW_OK;
  }
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++*out_count;
  }
  return SSW_OK;
}

int ssw_db_destroy(ssw_db* db) {
  if (!db) return SSW_OK;
  cudaSetDevice(db->device);
  if (db->stream) cudaStreamSynchronize(db->stream);
  cudaFree(db->d_vecs);
  cudaFree(db->d_row_ptr);
  cudaFree(db->d_img_dbidx);
  cudaFree(db->d_orig_row);
  cudaFree(db->d_part);
  cudaFree(db->d_part_side);
  cudaFree(db->d_last_bits);
  cudaFree(db->d_boxes);
  cudaFree(db->d_zoom);
  cudaFree(db->d_exact);
  if (db->d_xchg_timed_out) cudaFreeHost(db->d_xchg_timed_out);
  cudaFree(db->d_scan_stats);
  cudaFree(db->d_tc_ws);
  cudaFree(db->d_list_keys);
  cudaFree(db->d_list_dbidx);
  cudaFree(db->d_gthr);
  cudaFree(db->ws_alt.keys);
  cudaFree(db->ws_alt.dbidx);
  cudaFree(db->ws_alt.gthr);
  if (db->xstream) {
    cudaStreamSynchronize(db->xstream);
    cudaStreamDestroy(db->xstream);
  }
  if (db->ev_scan) cudaEventDestroy(db->ev_scan);
  if (db->ev_xdone) cudaEventDestroy(db->ev_xdone);
  cudaFree(db->d_stage);
  if (db->h_stage) cudaFreeHost(db->h_stage);
  for (auto* v : {&db->prof_pending, &db->prof_free})
    for (auto& ev : *v) {
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
  if (db->stream) cudaStreamDestroy(db->stream);
  delete db;
  return SSW_OK;
}

int ssw_db_create(ssw_db** out, int device, const void* vectors, int dtype_in, int dtype_store,
                  int64_t n_rows, int dim, const int32_t* dbidx_per_row, int64_t global_row_base) {
  SSW_REQUIRE(vectors != nullptr || n_rows == 0, "vectors is null");
  SSW_REQUIRE(dbidx_per_row != nullptr || n_rows == 0, "dbidx_per_row is null");
  SSW_REQUIRE(dtype_in == SSW_F16 || dtype_in == SSW_F32, "dtype_in must be SSW_F32 or SSW_F16");
  ssw_db* db = nullptr;
  int rc = db_new(&db, device, dtype_store, n_rows, dim, global_row_base);
  if (rc) return rc;
  std::vector<int64_t> perm;
  auto fail = [&](int code) {
    ssw_db_destroy(db);
    return code;
  };
  if ((rc = build_layout(db, dbidx_per_row, &perm))) return fail(rc);
  const size_t es_st = dtype_store == SSW_F16 ? 2 : 4;
  cudaError_t e = cudaMalloc(&db->d_vecs, std::max<size_t>((size_t)n_rows * dim * es_st, 16));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(vectors): ") + cudaGetErrorString(e));
    return fail(e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA);
  }
  if ((rc = upload_rows(db, vectors, dtype_in, db->d_vecs, dtype_store, perm))) return fail(rc);
  *out = db;
  return SSW_OK;
}

int ssw_db_create_device(ssw_db** out, int device, const void* d_vectors, int dtype_in, int dtype_store, int64_t n_rows,
                         int dim, const int32_t* dbidx_per_row, int64_t global_row_base) {
  SSW_REQUIRE(d_vectors != nullptr || n_rows == 0, "d_vectors is null");
  SSW_REQUIRE(dbidx_per_row != nullptr || n_rows == 0, "dbidx_per_row is null");
  SSW_REQUIRE(dtype_in == SSW_F16 || dtype_in == SSW_F32, "dtype_in must be SSW_F32 or SSW_F16");
  ssw_db* db = nullptr;
  int rc = db_new(&db, device, dtype_store, n_rows, dim, global_row_base);
  if (rc) return rc;
  auto fail = [&](int code) {
    ssw_db_destroy(db);
    return code;
  };
  for (int64_t i = 1; i < n_rows; ++i)
    if (dbidx_per_row[i] < dbidx_per_row[i - 1]) {
      set_error("device-resident vectors must be grouped by ascending dbidx");
      return fail(SSW_ERR_INVALID);
    }
  if ((rc = build_layout(db, dbidx_per_row, nullptr))) return fail(rc);
  const size_t es = dtype_store == SSW_F16 ? 2 : 4;
  cudaError_t e = cudaMalloc(&db->d_vecs, std::max<size_t>((size_t)n_rows * dim * es, 16));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(vectors): ") + cudaGetErrorString(e));
    return fail(e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA);
  }
  if ((rc = launch_convert_rows(d_vectors, dtype_in, db->d_vecs, dtype_store, n_rows * dim, db->stream))) return fail(rc);
  e = cudaStreamSynchronize(db->stream);
  if (e != cudaSuccess) {
    set_error(std::string("device copy: ") + cudaGetErrorString(e));
    return fail(SSW_ERR_CUDA);
  }
  *out = db;
  return SSW_OK;
}

int ssw_db_create_synthetic(ssw_db** out, int device, int dtype_store, int64_t n_rows, int dim,
                            const int32_t* dbidx_per_row, int64_t global_row_base, uint64_t seed, int kind) {
  SSW_REQUIRE(dbidx_per_row != nullptr || n_rows == 0, "dbidx_per_row is null");
  SSW_REQUIRE(kind == SSW_SYNTH_TRI || kind == SSW_SYNTH_LATTICE, "unknown synthetic kind");
  ssw_db* db = nullptr;
  int rc = db_new(&db, device, dtype_store, n_rows, dim, global_row_base);
  if (rc) return rc;
  auto fail = [&](int code) {
    ssw_db_destroy(db);
    return code;
  };
  for (int64_t i = 1; i < n_rows; ++i)
    if (dbidx_per_row[i] < dbidx_per_row[i - 1]) {
      set_error("synthetic databases need rows grouped by ascending dbidx");
      return fail(SSW_ERR_INVALID);
    }
  if ((rc = build_layout(db, dbidx_per_row, nullptr))) return fail(rc);
  const size_t es = dtype_store == SSW_F16 ? 2 : 4;
  cudaError_t e = cudaMalloc(&db->d_vecs, std::max<size_t>((size_t)n_rows * dim * es, 16));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(vectors): ") + cudaGetErrorString(e));
    return fail(e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA);
  }
  if ((rc = launch_synth(db->d_vecs, dtype_store, n_rows, dim, global_row_base, seed, kind, db->stream))) return fail(rc);
  e = cudaStreamSynchronize(db->stream);
  if (e != cudaSuccess) {
    set_error(std::string("synthetic fill: ") + cudaGetErrorString(e));
    return fail(SSW_ERR_CUDA);
  }
  *out = db;
  return SSW_OK;
}

int ssw_db_info(const ssw_db* db, int64_t* n_rows, int64_t* n_images, int* dim, int* dtype_store, int* device) {
  SSW_REQUIRE(db != nullptr, "db is null");
  if (n_rows) *n_rows = db->n_rows;
  if (n_images) *n_images = db->n_images;
  if (dim) *dim = db->dim;
  if (dtype_store) *dtype_store = db->dtype;
  if (device) *device = db->device;
  return SSW_OK;
}

int ssw_db_vectors_device(const ssw_db* db, void** dev_ptr) {
  SSW_REQUIRE(db != nullptr && dev_ptr != nullptr, "null argument");
  *dev_ptr = db->d_vecs;
  return SSW_OK;
}

int ssw_profile_enable(ssw_db* db, int on) {
  SSW_REQUIRE(db != nullptr, "db is null");
  db->profiling = on != 0;
  db->prof_stride = on > 1 ? on : 1;
  db->prof_counter = 0;
  db->prof_sampled = false;
  return SSW_OK;
}

int ssw_profile_read(ssw_db* db, double* scan_kernel_ms, int64_t* scan_kernel_launches) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_CUDA(cudaSetDevice(db->device));
  double total = 0.0;
  int64_t n = 0;
  for (auto& ev : db->prof_pending) {
    SSW_CUDA(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    SSW_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    total += ms;
    ++n;
    db->prof_free.push_back(ev);
  }
  db->prof_pending.clear();
  if (scan_kernel_ms) *scan_kernel_ms = total;
  if (scan_kernel_launches) *scan_kernel_launches = n;
  return SSW_OK;
}

int ssw_scan_stats(ssw_db* db, int enable, int64_t* list_updates, int64_t* images_offered) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_CUDA(cudaSetDevice(db->device));
  unsigned long long v[2] = {0, 0};
  if (db->d_scan_stats) {
    SSW_CUDA(cudaDeviceSynchronize());
    SSW_CUDA(cudaMemcpy(v, db->d_scan_stats, 16, cudaMemcpyDeviceToHost));
    SSW_CUDA(cudaMemset(db->d_scan_stats, 0, 16));
  }
  if (enable && !db->d_scan_stats) {
#ifdef SSW_TRACE
    const size_t stats_bytes = 128 + 192 * 16 * 8;      // + the development timeline of every CTA
#else
    const size_t stats_bytes = 16;
#endif
    SSW_CUDA(cudaMalloc((void**)&db->d_scan_stats, stats_bytes));
    SSW_CUDA(cudaMemset(db->d_scan_stats, 0, stats_bytes));
#ifdef SSW_TRACE
    g_trace_block = db->d_scan_stats;
#endif
  } else if (!enable && db->d_scan_stats) {
    cudaFree(db->d_scan_stats);
    db->d_scan_stats = nullptr;
#ifdef SSW_TRACE
    g_trace_block = nullptr;
#endif
  }
  if (list_updates) *list_updates = (int64_t)v[0];
  if (images_offered) *images_offered = (int64_t)v[1];
  return SSW_OK;
}

#ifdef SSW_TRACE
// development only: copy the CTA timelines (clock64 stamps, 16 per CTA) of the last traced launch
int ssw_scan_trace_read(ssw_db* db, long long* out, int n_ctas) {
  SSW_REQUIRE(db != nullptr && db->d_scan_stats != nullptr && n_ctas <= 192, "tracing is off");
  SSW_CUDA(cudaDeviceSynchronize());
  SSW_CUDA(cudaMemcpy(out, reinterpret_cast<long long*>(db->d_scan_stats) + 16, (size_t)n_ctas * 16 * 8, cudaMemcpyDeviceToHost));
  return SSW_OK;
}
#endif

int ssw_set_scan_mode(ssw_db* db, int mode) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  db->scan_mode = mode;
  return SSW_OK;
}

int ssw_exclude_words(const ssw_db* db, int64_t* words_per_query) {
  SSW_REQUIRE(db != nullptr && words_per_query != nullptr, "null argument");
  *words_per_query = db->excl_words;
  return SSW_OK;
}

int ssw_exclude_build_device(ssw_db* db, const int32_t* d_exclude_dbidx, const int64_t* d_exclude_offsets,
                             int nq, int64_t n_ids_total, uint32_t* d_bits_out, void* stream) {
  SSW_REQUIRE(db != nullptr && d_bits_out != nullptr && d_exclude_offsets != nullptr, "null argument");
  SSW_REQUIRE(nq > 0, "nq must be positive");
  (void)n_ids_total;
  SSW_CUDA(cudaSetDevice(db->device));
  return launch_exclude_build(db, d_exclude_dbidx, d_exclude_offsets, nq, d_bits_out, (cudaStream_t)stream);
}

struct XchgCtx {
  void* const* peers;
  int world, rank, nq_cap, k_cap;
  uint32_t epoch;
};

static int scan_topk_impl(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_exclude_bits,
                          uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                          int32_t* d_out_count, cudaStream_t st, const XchgCtx* xc = nullptr, bool exact_rows = false) {
  if (db->pipe_pending) {      // a pipelined step's exchange may still read the scan workspace: order this call behind it
    SSW_CUDA(cudaStreamWaitEvent(st, db->ev_xdone, 0));
    db->pipe_pending = false;
  }
  const int lists = db->scan_grid;
  const int kl = k + kScanTcSlack;           // slots per (query, CTA): the batched scan publishes up to k + slack entries
  int rc = ensure_lists(db, nq, lists, kl);
  if (rc) return rc;
  const bool tc_ok = scan_tc_supported(db, k) && !exact_rows;
  if (db->scan_mode == 2 && !tc_ok && !exact_rows) {
    set_error("tcgen05 batched scan needs fp16 storage, dim 256/512/768 and k <= 64");
    return SSW_ERR_INVALID;
  }
  const bool use_tc = !exact_rows && (db->scan_mode == 2 || (db->scan_mode == 0 && tc_ok && nq >= 2));   // one batched pass costs about one streaming pass
  if (use_tc) {
    if (!db->d_tc_ws) SSW_CUDA(cudaMalloc(&db->d_tc_ws, scan_tc_workspace_bytes(db->dim, db->scan_grid)));
    for (int q0 = 0; q0 < nq; q0 += SSW_MAX_BATCH) {
      const int nb = std::min(SSW_MAX_BATCH, nq - q0);
      rc = launch_scan_tc(db, d_queries + (size_t)q0 * db->dim, nb, k,
                          d_exclude_bits ? d_exclude_bits + (size_t)q0 * db->excl_words : nullptr,
                          db->d_list_keys + (size_t)q0 * lists * kl, db->d_list_dbidx + (size_t)q0 * lists * kl,
                          db->d_cand_cnt + q0, db->d_gthr + q0, db->d_tc_ws, st);
      if (rc) return rc;
    }
  } else {
    SSW_CUDA(cudaMemsetAsync(db->d_gthr, 0, (size_t)db->gthr_capacity * 8 + (size_t)db->gthr_capacity * db->scan_grid * 4, st));
    for (int q = 0; q < nq; ++q) {
      prof_begin(db, st);
      rc = launch_scan1(db, d_queries + (size_t)q * db->dim, k,
                        d_exclude_bits ? d_exclude_bits + (size_t)q * db->excl_words : nullptr,
                        db->d_list_keys + (size_t)q * lists * k, db->d_list_dbidx + (size_t)q * lists * k,
                        db->d_gthr + q, db->d_pub1 + (size_t)q * db->scan_grid, st, exact_rows);
      prof_end(db, st);
      if (rc) return rc;
    }
  }
  if (xc) {
    if (!db->d_xchg_timed_out) {      // pinned host word the kernel writes directly (UVA): read without a CUDA call
      SSW_CUDA(cudaHostAlloc((void**)&db->d_xchg_timed_out, 4, cudaHostAllocMapped));
      *db->d_xchg_timed_out = 0;
    }
    return launch_exchange_merge(db->d_list_keys, db->d_list_dbidx, lists, k, (int64_t)lists * (use_tc ? kl : k), nq, k, db->d_gthr,
                                 xc->peers, xc->world, xc->rank, xc->nq_cap, xc->k_cap, xc->epoch, db->d_xchg_timed_out,
                                 d_out_key, d_out_dbidx, d_out_score, d_out_row, d_out_count, st, use_tc && !db->prof_sampled,
                                 use_tc ? db->d_cand_cnt : nullptr);
  }
  // after the batched scan the merge is a programmatic dependent launch (the scan triggers it early)
  return launch_merge(db->d_list_keys, db->d_list_dbidx, lists, k, (int64_t)lists * (use_tc ? kl : k), nq, k, db->d_gthr,
                      d_out_key, d_out_dbidx, d_out_score, d_out_row, d_out_count, st, use_tc && !db->prof_sampled,
                      use_tc ? db->d_cand_cnt : nullptr);
}

int ssw_scan_topk_sharded_device(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_exclude_bits,
                                 void* const* peer_bufs, int world, int rank, int nq_cap, int k_cap, uint32_t epoch,
                                 uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                                 int32_t* d_out_count, void* stream) {
  SSW_REQUIRE(db != nullptr && d_queries != nullptr && peer_bufs != nullptr, "null argument");
  SSW_REQUIRE(world >= 1 && world <= SSW_MAX_WORLD && rank >= 0 && rank < world, "bad world/rank");
  SSW_REQUIRE(nq > 0 && nq <= nq_cap && nq <= db->sm_count, "nq must be in [1, min(nq_cap, SM count)]");
  SSW_REQUIRE(k > 0 && k <= k_cap && k <= SSW_MAX_TOPK, "k must be in [1, min(k_cap, SSW_MAX_TOPK)]");
  SSW_REQUIRE(epoch != 0, "epoch must be non-zero and increase by one per call");
  for (int i = 0; i < world; ++i) SSW_REQUIRE(peer_bufs[i] != nullptr, "peer buffer is null");
  SSW_CUDA(cudaSetDevice(db->device));
  XchgCtx xc{peer_bufs, world, rank, nq_cap, k_cap, epoch};
  return scan_topk_impl(db, d_queries, nq, k, d_exclude_bits, d_out_key, d_out_dbidx, d_out_score, d_out_row,
                        d_out_count, (cudaStream_t)stream, &xc);
}

// SMs the pipelined step leaves to its exchange blocks.  Automatic choice (side_sms < 0): giving up 4 of 148 SMs costs
// the scan up to 2.7 % of its time (it follows the SM count once the boards power-cap their clock), exchange blocks
// NEXT TO the scan CTAs cost it a roughly constant ~12 us (scan CTAs of the next step find their SM occupied) — so the
// exchange gets SMs of its own only on small shards (< 2.5 GB: under ~0.4 ms of scan; the 8-GPU shard of 10M x 512).
static int pipeline_side_sms(const ssw_db* db) {
  int side = db->side_sms;
  if (side < 0) side = (double)db->n_rows * db->dim * 2.0 < 2.5e9 ? 4 : 0;
  return db->sm_count - side >= 64 ? side : 0;
}

int ssw_scan_topk_sharded_pipelined_device(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_exclude_bits,
                                           void* const* peer_bufs, int world, int rank, int nq_cap, int k_cap, uint32_t epoch,
                                           uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                                           int32_t* d_out_count, void* stream) {
  SSW_REQUIRE(db != nullptr && d_queries != nullptr && peer_bufs != nullptr, "null argument");
  SSW_REQUIRE(world >= 1 && world <= SSW_MAX_WORLD && rank >= 0 && rank < world, "bad world/rank");
  SSW_REQUIRE(nq >= 1 && nq <= nq_cap && nq <= SSW_MAX_BATCH, "pipelined step: 1 <= nq <= min(nq_cap, 64)");
  SSW_REQUIRE(k > 0 && k <= k_cap && scan_tc_supported(db, k) && world * k <= 1024,
              "pipelined step needs the batched kernel (fp16 storage, dim 256/512/768, k <= 64)");
  SSW_REQUIRE(epoch != 0, "epoch must be non-zero and increase by one per call");
  SSW_REQUIRE(db->d_exact == nullptr, "the pipelined step serves plain fp16 databases (exact mode certifies on the host path)");
  for (int i = 0; i < world; ++i) SSW_REQUIRE(peer_bufs[i] != nullptr, "peer buffer is null");
  SSW_CUDA(cudaSetDevice(db->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!db->xstream) {
    SSW_CUDA(cudaStreamCreateWithFlags(&db->xstream, cudaStreamNonBlocking));
    SSW_CUDA(cudaEventCreateWithFlags(&db->ev_scan, cudaEventDisableTiming));
    SSW_CUDA(cudaEventCreateWithFlags(&db->ev_xdone, cudaEventDisableTiming));
  }
  if (!db->d_xchg_timed_out) {
    SSW_CUDA(cudaHostAlloc((void**)&db->d_xchg_timed_out, 4, cudaHostAllocMapped));
    *db->d_xchg_timed_out = 0;
  }
  // the scan of a pipelined step leaves `side_sms` SMs to the exchange blocks: its own grid and image partition
  const int side = pipeline_side_sms(db);
  const int grid = db->sm_count - side;
  ScanTcGrid sg{grid, db->d_part, db->max_cta_images};
  if (side > 0) {
    if (db->side_grid != grid) {
      int32_t* part = nullptr;
      SSW_CUDA(cudaMalloc((void**)&part, ((size_t)grid * kScanWarps + 1) * 4 + 4));
      int* d_max = reinterpret_cast<int*>(part + (size_t)grid * kScanWarps + 1);
      int rc0 = launch_part_build(db, grid, part, d_max, st);
      int h_max = 0;
      cudaError_t e = rc0 ? cudaSuccess : cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, st);
      if (!rc0 && e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (rc0 || e != cudaSuccess) {
        cudaFree(part);
        if (rc0) return rc0;
        set_error(std::string("side partition: ") + cudaGetErrorString(e));
        return SSW_ERR_CUDA;
      }
      if (db->d_part_side) cudaFree(db->d_part_side);
      db->d_part_side = part;
      db->side_grid = grid;
      db->side_max_cta_images = h_max;
    }
    sg.part = db->d_part_side;
    sg.max_cta_images = db->side_max_cta_images;
  }
  swap_scan_ws(db);       // this set was last read by the exchange two steps back, which `stream` already waited for
  const int lists = grid, kl = k + kScanTcSlack;
  int rc = ensure_lists(db, nq, db->scan_grid, kl);
  if (rc) return rc;
  if (!db->d_tc_ws) SSW_CUDA(cudaMalloc(&db->d_tc_ws, scan_tc_workspace_bytes(db->dim, db->scan_grid)));
  rc = launch_scan_tc(db, d_queries, nq, k, d_exclude_bits, db->d_list_keys, db->d_list_dbidx, db->d_cand_cnt, db->d_gthr,
                      db->d_tc_ws, st, &sg);
  if (rc) return rc;
  // from here on `stream` holds the PREVIOUS call's results (and frees its workspace for the next call)
  if (db->pipe_pending) SSW_CUDA(cudaStreamWaitEvent(st, db->ev_xdone, 0));
  SSW_CUDA(cudaEventRecord(db->ev_scan, st));
  SSW_CUDA(cudaStreamWaitEvent(db->xstream, db->ev_scan, 0));
  rc = launch_exchange_slim(db->d_list_keys, db->d_list_dbidx, (int64_t)lists * kl, nq, k, db->d_cand_cnt, peer_bufs, world, rank,
                            nq_cap, k_cap, epoch, db->d_xchg_timed_out, d_out_key, d_out_dbidx, d_out_score, d_out_row,
                            d_out_count, db->xstream, side);
  if (rc) return rc;
  SSW_CUDA(cudaEventRecord(db->ev_xdone, db->xstream));
  db->pipe_pending = true;
  return SSW_OK;
}

int ssw_scan_pipeline_side_sms(ssw_db* db, int side_sms) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_REQUIRE(side_sms >= -1 && side_sms <= 16, "side_sms must be in [-1, 16] (-1 = automatic)");
  db->side_sms = side_sms;
  return SSW_OK;
}

int ssw_scan_pipeline_side_sms_in_use(const ssw_db* db, int* side_sms) {
  SSW_REQUIRE(db != nullptr && side_sms != nullptr, "null argument");
  *side_sms = pipeline_side_sms(db);
  return SSW_OK;
}

int ssw_scan_pipeline_drain(ssw_db* db, void* stream) {
  SSW_REQUIRE(db != nullptr, "db is null");
  SSW_CUDA(cudaSetDevice(db->device));
  if (db->pipe_pending) {
    SSW_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, db->ev_xdone, 0));
    db->pipe_pending = false;
  }
  return SSW_OK;
}

int ssw_xchg_create(int device, int world, int nq_cap, int k_cap, void** d_buf, void* ipc_handle_out, int64_t* bytes) {
  SSW_REQUIRE(d_buf != nullptr, "d_buf is null");
  SSW_REQUIRE(world >= 1 && world <= SSW_MAX_WORLD && nq_cap >= 1 && k_cap >= 1 && k_cap <= SSW_MAX_TOPK, "bad capacity");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  const size_t n = xchg_bytes(world, nq_cap, k_cap);
  SSW_CUDA(cudaMalloc(d_buf, n));
  SSW_CUDA(cudaMemset(*d_buf, 0, n));
  SSW_CUDA(cudaDeviceSynchronize());
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SSW_CUDA(cudaIpcGetMemHandle(&h, *d_buf));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (bytes) *bytes = (int64_t)n;
  return SSW_OK;
}

int ssw_xchg_open(int device, const void* ipc_handle, void** d_peer) {
  SSW_REQUIRE(ipc_handle != nullptr && d_peer != nullptr, "null argument");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  SSW_CUDA(cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess));
  return SSW_OK;
}

int ssw_xchg_close(int device, void* d_peer) {
  if (!d_peer) return SSW_OK;
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  SSW_CUDA(cudaIpcCloseMemHandle(d_peer));
  return SSW_OK;
}

int ssw_xchg_destroy(int device, void* d_buf) {
  if (!d_buf) return SSW_OK;
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  SSW_CUDA(cudaFree(d_buf));
  return SSW_OK;
}

int ssw_scan_topk_device(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_exclude_bits,
                         uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                         int32_t* d_out_count, void* stream) {
  SSW_REQUIRE(db != nullptr && d_queries != nullptr, "null argument");
  SSW_REQUIRE(nq > 0, "nq must be positive");
  SSW_REQUIRE(k > 0 && k <= SSW_MAX_TOPK, "k must be in [1, SSW_MAX_TOPK]");
  SSW_CUDA(cudaSetDevice(db->device));
  return scan_topk_impl(db, d_queries, nq, k, d_exclude_bits, d_out_key, d_out_dbidx, d_out_score, d_out_row,
                        d_out_count, (cudaStream_t)stream);
}

int ssw_merge_topk_device(int device, const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists, int nq, int k,
                          uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                          int32_t* d_out_count, void* stream) {
  SSW_REQUIRE(d_keys != nullptr && d_dbidx != nullptr, "null argument");
  SSW_REQUIRE(n_lists > 0 && nq > 0, "n_lists and nq must be positive");
  SSW_REQUIRE(k > 0 && k <= SSW_MAX_TOPK, "k must be in [1, SSW_MAX_TOPK]");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  return launch_merge(d_keys, d_dbidx, n_lists, (int64_t)nq * k, k, nq, k, nullptr, d_out_key, d_out_dbidx,
                      d_out_score, d_out_row, d_out_count, (cudaStream_t)stream);
}

// Candidates the fp16 scan is asked for in exact mode (see ssw_exact.cu): the tensor-core kernel keeps at
// most 64 per query, the streaming kernel up to SSW_MAX_TOPK.
static int exact_candidates(const ssw_db* db, int nq, int k) {
  const int64_t cap = std::max<int64_t>(k, std::min<int64_t>(db->n_images, SSW_MAX_TOPK));
  const bool tc = db->scan_mode != 1 && nq >= 2 && scan_tc_supported(db, 64) && k <= 56;
  if (db->scan_mode == 2) return (int)std::min<int64_t>(cap, 64);
  if (tc) return (int)std::min<int64_t>(cap, 64);
  return (int)std::min<int64_t>(cap, (int64_t)k + std::max(16, k / 4));
}

static int scan_topk_host(ssw_db* db, const float* queries, int nq, int k, const int32_t* exclude_dbidx,
                          const int64_t* exclude_offsets, int32_t* out_dbidx, float* out_score, int64_t* out_row,
                          int32_t* out_count, const XchgCtx* xc) {
  SSW_REQUIRE(db != nullptr && queries != nullptr, "null argument");
  SSW_REQUIRE(nq > 0, "nq must be positive");
  SSW_REQUIRE(k > 0 && k <= SSW_MAX_TOPK, "k must be in [1, SSW_MAX_TOPK]");
  SSW_REQUIRE((exclude_dbidx == nullptr) == (exclude_offsets == nullptr) || exclude_offsets != nullptr,
              "exclude_offsets is required with exclude_dbidx");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  if (exclude_offsets) {
    SSW_REQUIRE(exclude_offsets[0] == 0, "exclude_offsets[0] must be 0");
    for (int i = 0; i < nq; ++i) SSW_REQUIRE(exclude_offsets[i + 1] >= exclude_offsets[i], "exclude_offsets must be non-decreasing");
  }
  const int64_t n_ids = exclude_offsets ? exclude_offsets[nq] : 0;
  SSW_REQUIRE(n_ids == 0 || exclude_dbidx != nullptr, "exclude_dbidx is null");
  const bool has_excl = exclude_offsets != nullptr && n_ids > 0;
  const bool exact = db->d_exact != nullptr;
  const int kc = exact ? exact_candidates(db, nq, k) : k;
  // one staging block each side:  queries | offsets | ids   ->   + bitmaps | keys | [exact scratch] | dbidx | score | row | count [| cert]
  auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
  const size_t q_bytes = up16((size_t)nq * db->dim * 4);
  const size_t off_bytes = up16((size_t)(nq + 1) * 8);
  const size_t ids_bytes = up16((size_t)n_ids * 4);
  const size_t in_bytes = q_bytes + off_bytes + ids_bytes;
  const size_t bits_bytes = has_excl ? up16((size_t)nq * db->excl_words * 4) : 0;
  const size_t nk = (size_t)nq * k, nkc = (size_t)nq * kc;
  const size_t key_b = up16(nk * 8), db_b = up16(nk * 4), sc_b = up16(nk * 4), row_b = up16(nk * 8), cnt_b = up16((size_t)nq * 4);
  const size_t x16_b = exact ? up16(nkc * 8) : 0, xdb_b = exact ? up16(nkc * 4) : 0, x32_b = exact ? up16(nkc * 8) : 0;
  const size_t cert_b = exact ? cnt_b : 0;
  const size_t out_bytes = db_b + sc_b + row_b + cnt_b + cert_b;
  const size_t dev_total = in_bytes + bits_bytes + key_b + x16_b + xdb_b + x32_b + out_bytes;
  int rc = ensure_stage(db, dev_total, std::max(in_bytes, out_bytes));
  if (rc) return rc;
  uint8_t* h = static_cast<uint8_t*>(db->h_stage);
  uint8_t* d = static_cast<uint8_t*>(db->d_stage);
  memcpy(h, queries, (size_t)nq * db->dim * 4);
  if (has_excl) {
    memcpy(h + q_bytes, exclude_offsets, (size_t)(nq + 1) * 8);
    memcpy(h + q_bytes + off_bytes, exclude_dbidx, (size_t)n_ids * 4);
  }
  cudaStream_t st = db->stream;
  SSW_CUDA(cudaMemcpyAsync(d, h, has_excl ? in_bytes : q_bytes, cudaMemcpyHostToDevice, st));
  const float* d_q = reinterpret_cast<const float*>(d);
  uint32_t* d_bits = nullptr;
  if (has_excl) {
    d_bits = reinterpret_cast<uint32_t*>(d + in_bytes);
    rc = launch_exclude_build(db, reinterpret_cast<const int32_t*>(d + q_bytes + off_bytes),
                              reinterpret_cast<const int64_t*>(d + q_bytes), nq, d_bits, st);
    if (rc) return rc;
  }
  uint8_t* p = d + in_bytes + bits_bytes;
  uint64_t* d_key = reinterpret_cast<uint64_t*>(p);
  p += key_b;
  uint64_t* d_key16 = reinterpret_cast<uint64_t*>(p);
  p += x16_b;
  int32_t* d_cand = reinterpret_cast<int32_t*>(p);
  p += xdb_b;
  uint64_t* d_key32 = reinterpret_cast<uint64_t*>(p);
  p += x32_b;
  uint8_t* d_out = p;
  int32_t* d_dbidx = reinterpret_cast<int32_t*>(d_out);
  float* d_score = reinterpret_cast<float*>(d_out + db_b);
  int64_t* d_row = reinterpret_cast<int64_t*>(d_out + db_b + sc_b);
  int32_t* d_cnt = reinterpret_cast<int32_t*>(d_out + db_b + sc_b + row_b);
  int32_t* d_cert = reinterpret_cast<int32_t*>(d_out + db_b + sc_b + row_b + cnt_b);
  auto check_xchg = [&]() -> int {
    if (xc && db->d_xchg_timed_out) {      // read after the stream synchronised
      const int flag = *static_cast<volatile int*>(db->d_xchg_timed_out);
      if (flag) {
        *db->d_xchg_timed_out = 0;
        set_error("fused exchange: a peer rank did not deliver its lists within 10 s");
        return SSW_ERR_CUDA;
      }
    }
    return SSW_OK;
  };
  if (!exact) {
    rc = scan_topk_impl(db, d_q, nq, k, d_bits, d_key, d_dbidx, d_score, d_row, d_cnt, st, xc);
    if (rc) return rc;
    SSW_CUDA(cudaMemcpyAsync(h, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    SSW_CUDA(cudaStreamSynchronize(st));
    if ((rc = check_xchg())) return rc;
  } else {
    // 1. fp16 scan for kc candidates  2. fp32 re-scoring  3. ranking + certificate   (ssw_exact.cu)
    rc = scan_topk_impl(db, d_q, nq, kc, d_bits, d_key16, d_cand, nullptr, nullptr, nullptr, st, nullptr);
    if (rc) return rc;
    if ((rc = launch_exact_rescore(db, d_q, nq, kc, d_cand, d_key32, st))) return rc;
    const double err_unit = db->exact_rho + db->exact_vmax * (double)db->dim * 1.8e-7;
    if ((rc = launch_exact_finish(d_q, db->dim, nq, kc, k, d_key16, d_key32, d_cand, err_unit, d_dbidx, d_score, d_row,
                                  d_cnt, d_key, d_cert, st)))
      return rc;
    SSW_CUDA(cudaMemcpyAsync(h, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    SSW_CUDA(cudaStreamSynchronize(st));
    const int32_t* cert = reinterpret_cast<const int32_t*>(h + db_b + sc_b + row_b + cnt_b);
    int n_rescan = 0;
    for (int q = 0; q < nq; ++q)
      if (!cert[q]) {      // scores too close to call from fp16: the streaming kernel over the fp32 rows decides
        rc = scan_topk_impl(db, d_q + (size_t)q * db->dim, 1, k, d_bits ? d_bits + (size_t)q * db->excl_words : nullptr,
                            d_key + (size_t)q * k, d_dbidx + (size_t)q * k, d_score + (size_t)q * k, d_row + (size_t)q * k,
                            d_cnt + q, st, nullptr, true);
        if (rc) return rc;
        ++n_rescan;
      }
    db->exact_queries += nq;
    db->exact_rescans += n_rescan;
    if (xc) {        // this shard's exact top-k is ONE list per query for the fused exchange
      if (!db->d_xchg_timed_out) {      // pinned host word the kernel writes directly (UVA): read without a CUDA call
        SSW_CUDA(cudaHostAlloc((void**)&db->d_xchg_timed_out, 4, cudaHostAllocMapped));
        *db->d_xchg_timed_out = 0;
      }
      // the exchange writes the world's result over d_out while reading d_key / d_dbidx: hand it copies
      rc = ensure_lists(db, nq, 1, k);
      if (rc) return rc;
      SSW_CUDA(cudaMemcpyAsync(db->d_list_keys, d_key, nk * 8, cudaMemcpyDeviceToDevice, st));
      SSW_CUDA(cudaMemcpyAsync(db->d_list_dbidx, d_dbidx, nk * 4, cudaMemcpyDeviceToDevice, st));
      rc = launch_exchange_merge(db->d_list_keys, db->d_list_dbidx, 1, k, k, nq, k, nullptr, xc->peers, xc->world, xc->rank,
                                 xc->nq_cap, xc->k_cap, xc->epoch, db->d_xchg_timed_out, d_key, d_dbidx, d_score, d_row,
                                 d_cnt, st);
      if (rc) return rc;
    }
    if (n_rescan || xc) {
      SSW_CUDA(cudaMemcpyAsync(h, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
      SSW_CUDA(cudaStreamSynchronize(st));
      if ((rc = check_xchg())) return rc;
    }
  }
  if (out_dbidx) memcpy(out_dbidx, h, nk * 4);
  if (out_score) memcpy(out_score, h + db_b, nk * 4);
  if (out_row) memcpy(out_row, h + db_b + sc_b, nk * 8);
  if (out_count) memcpy(out_count, h + db_b + sc_b + row_b, (size_t)nq * 4);
  return SSW_OK;
}

int ssw_db_attach_exact(ssw_db* db, const float* vectors) {
  SSW_REQUIRE(db != nullptr && (vectors != nullptr || db->n_rows == 0), "null argument");
  SSW_REQUIRE(db->dtype == SSW_F16, "an exact fp32 copy only makes sense next to fp16 storage");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  if (db->d_exact) {
    cudaFree(db->d_exact);
    db->d_exact = nullptr;
  }
  std::vector<int64_t> perm;
  if (db->d_orig_row) {
    perm.resize(db->n_rows);
    SSW_CUDA(cudaMemcpy(perm.data(), db->d_orig_row, (size_t)db->n_rows * 8, cudaMemcpyDeviceToHost));
  }
  float* d = nullptr;
  cudaError_t e = cudaMalloc((void**)&d, std::max<size_t>((size_t)db->n_rows * db->dim * 4, 16));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(exact fp32 copy): ") + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA;
  }
  int rc = upload_rows(db, vectors, SSW_F32, d, SSW_F32, perm);
  float* d_stats = nullptr;
  float stats[2] = {0.f, 0.f};
  if (!rc && cudaMalloc((void**)&d_stats, 8) != cudaSuccess) rc = SSW_ERR_CUDA;
  if (!rc) rc = launch_row_error_stats(d, db->n_rows, db->dim, d_stats, db->stream);
  if (!rc && cudaMemcpyAsync(stats, d_stats, 8, cudaMemcpyDeviceToHost, db->stream) != cudaSuccess) rc = SSW_ERR_CUDA;
  if (!rc && cudaStreamSynchronize(db->stream) != cudaSuccess) rc = SSW_ERR_CUDA;
  cudaFree(d_stats);
  if (rc) {
    cudaFree(d);
    if (rc == SSW_ERR_CUDA) set_error("attaching the exact copy failed");
    return rc;
  }
  db->d_exact = d;
  db->exact_rho = std::sqrt((double)stats[0]) * (1.0 + 1e-6);
  db->exact_vmax = std::sqrt((double)stats[1]) * (1.0 + 1e-6);
  return SSW_OK;
}

int ssw_db_exact_info(ssw_db* db, int* attached, double* rho, double* vmax, int64_t* queries, int64_t* rescans) {
  SSW_REQUIRE(db != nullptr, "db is null");
  if (attached) *attached = db->d_exact != nullptr;
  if (rho) *rho = db->exact_rho;
  if (vmax) *vmax = db->exact_vmax;
  if (queries) *queries = db->exact_queries;
  if (rescans) *rescans = db->exact_rescans;
  return SSW_OK;
}

int ssw_topk_from_scores(ssw_db* db, const float* scores, const uint8_t* row_mask, int k, const int32_t* exclude_dbidx,
                         int64_t n_exclude, int32_t* out_dbidx, float* out_score, int64_t* out_row, int32_t* out_count) {
  SSW_REQUIRE(db != nullptr && (scores != nullptr || db->n_rows == 0), "null argument");
  SSW_REQUIRE(k > 0 && k <= SSW_MAX_TOPK, "k must be in [1, SSW_MAX_TOPK]");
  SSW_REQUIRE(n_exclude >= 0 && (n_exclude == 0 || exclude_dbidx != nullptr), "bad exclude list");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
  const int64_t n = db->n_rows;
  const int64_t n_lists = std::max<int64_t>(1, (db->n_images + k - 1) / k), n_padded = n_lists * k;
  // device block: scores | mask | offsets | ids | bitmap | image keys | image dbidx | outputs
  const size_t sc_b = up16((size_t)n * 4), mk_b = row_mask ? up16((size_t)n) : 0, off_b = 16, ids_b = up16((size_t)n_exclude * 4);
  const size_t bits_b = n_exclude ? up16((size_t)db->excl_words * 4) : 0;
  const size_t key_b = up16((size_t)n_padded * 8), id_b = up16((size_t)n_padded * 4);
  const size_t o_db = up16((size_t)k * 4), o_sc = up16((size_t)k * 4), o_row = up16((size_t)k * 8), o_cnt = 16;
  const size_t in_b = sc_b + mk_b + off_b + ids_b, out_b = o_db + o_sc + o_row + o_cnt;
  int rc = ensure_stage(db, in_b + bits_b + key_b + id_b + out_b, std::max(off_b + ids_b, out_b));
  if (rc) return rc;
  uint8_t* d = static_cast<uint8_t*>(db->d_stage);
  uint8_t* h = static_cast<uint8_t*>(db->h_stage);
  cudaStream_t st = db->stream;
  if (n) SSW_CUDA(cudaMemcpyAsync(d, scores, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  if (row_mask && n) SSW_CUDA(cudaMemcpyAsync(d + sc_b, row_mask, (size_t)n, cudaMemcpyHostToDevice, st));
  uint32_t* d_bits = nullptr;
  if (n_exclude) {
    int64_t* off = reinterpret_cast<int64_t*>(h);
    off[0] = 0;
    off[1] = n_exclude;
    memcpy(h + off_b, exclude_dbidx, (size_t)n_exclude * 4);
    SSW_CUDA(cudaMemcpyAsync(d + sc_b + mk_b, h, off_b + ids_b, cudaMemcpyHostToDevice, st));
    d_bits = reinterpret_cast<uint32_t*>(d + in_b);
    rc = launch_exclude_build(db, reinterpret_cast<const int32_t*>(d + sc_b + mk_b + off_b),
                              reinterpret_cast<const int64_t*>(d + sc_b + mk_b), 1, d_bits, st);
    if (rc) return rc;
  }
  uint64_t* d_keys = reinterpret_cast<uint64_t*>(d + in_b + bits_b);
  int32_t* d_ids = reinterpret_cast<int32_t*>(d + in_b + bits_b + key_b);
  rc = launch_image_max(db, reinterpret_cast<const float*>(d), row_mask ? d + sc_b : nullptr, d_bits, n_padded, d_keys, d_ids, st);
  if (rc) return rc;
  uint8_t* d_out = d + in_b + bits_b + key_b + id_b;
  rc = launch_merge(d_keys, d_ids, (int)n_lists, k, n_padded, 1, k, nullptr, nullptr, reinterpret_cast<int32_t*>(d_out),
                    reinterpret_cast<float*>(d_out + o_db), reinterpret_cast<int64_t*>(d_out + o_db + o_sc),
                    reinterpret_cast<int32_t*>(d_out + o_db + o_sc + o_row), st);
  if (rc) return rc;
  SSW_CUDA(cudaStreamSynchronize(st));      // the uploads above read pageable host memory: finish before h is reused
  SSW_CUDA(cudaMemcpyAsync(h, d_out, out_b, cudaMemcpyDeviceToHost, st));
  SSW_CUDA(cudaStreamSynchronize(st));
  if (out_dbidx) memcpy(out_dbidx, h, (size_t)k * 4);
  if (out_score) memcpy(out_score, h + o_db, (size_t)k * 4);
  if (out_row) memcpy(out_row, h + o_db + o_sc, (size_t)k * 8);
  if (out_count) memcpy(out_count, h + o_db + o_sc + o_row, 4);
  return SSW_OK;
}

int ssw_topk_from_order(ssw_db* db, const int64_t* row_order, int64_t n_order, int k, const int32_t* exclude_dbidx,
                        int64_t n_exclude, int32_t* out_dbidx, int64_t* out_pos, int64_t* out_row, int32_t* out_count) {
  SSW_REQUIRE(db != nullptr && (row_order != nullptr || n_order == 0), "null argument");
  SSW_REQUIRE(n_order >= 0 && n_order < (int64_t)0xFFFFFFFEll, "row list too long");
  SSW_REQUIRE(k > 0 && k <= SSW_MAX_TOPK, "k must be in [1, SSW_MAX_TOPK]");
  SSW_REQUIRE(n_exclude >= 0 && (n_exclude == 0 || exclude_dbidx != nullptr), "bad exclude list");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
  const int64_t n = db->n_rows;
  const int64_t n_lists = std::max<int64_t>(1, (db->n_images + k - 1) / k), n_padded = n_lists * k;
  // device block: order | pos | offsets | ids | bitmap | image keys | image dbidx | outputs (dbidx, key, count)
  const size_t ord_b = up16((size_t)n_order * 8), pos_b = up16((size_t)std::max<int64_t>(n, 1) * 4), off_b = 16;
  const size_t ids_b = up16((size_t)n_exclude * 4), bits_b = n_exclude ? up16((size_t)db->excl_words * 4) : 0;
  const size_t key_b = up16((size_t)n_padded * 8), id_b = up16((size_t)n_padded * 4);
  const size_t o_db = up16((size_t)k * 4), o_key = up16((size_t)k * 8), o_cnt = 16;
  const size_t out_b = o_db + o_key + o_cnt;
  int rc = ensure_stage(db, ord_b + pos_b + off_b + ids_b + bits_b + key_b + id_b + out_b, std::max(off_b + ids_b, out_b));
  if (rc) return rc;
  uint8_t* d = static_cast<uint8_t*>(db->d_stage);
  uint8_t* h = static_cast<uint8_t*>(db->h_stage);
  cudaStream_t st = db->stream;
  if (n_order) SSW_CUDA(cudaMemcpyAsync(d, row_order, (size_t)n_order * 8, cudaMemcpyHostToDevice, st));
  uint32_t* d_pos = reinterpret_cast<uint32_t*>(d + ord_b);
  if ((rc = launch_order_to_pos(reinterpret_cast<const int64_t*>(d), n_order, n, d_pos, st))) return rc;
  uint32_t* d_bits = nullptr;
  uint8_t* p = d + ord_b + pos_b;
  if (n_exclude) {
    int64_t* off = reinterpret_cast<int64_t*>(h);
    off[0] = 0;
    off[1] = n_exclude;
    memcpy(h + off_b, exclude_dbidx, (size_t)n_exclude * 4);
    SSW_CUDA(cudaMemcpyAsync(p, h, off_b + ids_b, cudaMemcpyHostToDevice, st));
    d_bits = reinterpret_cast<uint32_t*>(p + off_b + ids_b);
    rc = launch_exclude_build(db, reinterpret_cast<const int32_t*>(p + off_b), reinterpret_cast<const int64_t*>(p), 1, d_bits, st);
    if (rc) return rc;
  }
  p += off_b + ids_b + bits_b;
  uint64_t* d_keys = reinterpret_cast<uint64_t*>(p);
  int32_t* d_ids = reinterpret_cast<int32_t*>(p + key_b);
  rc = launch_image_max(db, nullptr, nullptr, d_bits, n_padded, d_keys, d_ids, st, d_pos);
  if (rc) return rc;
  uint8_t* d_out = p + key_b + id_b;
  rc = launch_merge(d_keys, d_ids, (int)n_lists, k, n_padded, 1, k, nullptr, reinterpret_cast<uint64_t*>(d_out + o_db),
                    reinterpret_cast<int32_t*>(d_out), nullptr, nullptr, reinterpret_cast<int32_t*>(d_out + o_db + o_key), st);
  if (rc) return rc;
  SSW_CUDA(cudaStreamSynchronize(st));      // the uploads above read pageable host memory: finish before h is reused
  SSW_CUDA(cudaMemcpyAsync(h, d_out, out_b, cudaMemcpyDeviceToHost, st));
  SSW_CUDA(cudaStreamSynchronize(st));
  int cnt = 0;
  memcpy(&cnt, h + o_db + o_key, 4);
  const uint64_t* keys = reinterpret_cast<const uint64_t*>(h + o_db);
  for (int i = 0; i < k; ++i) {
    const bool ok = i < cnt;
    if (out_dbidx) out_dbidx[i] = ok ? reinterpret_cast<const int32_t*>(h)[i] : -1;
    if (out_pos) out_pos[i] = ok ? (int64_t)(0xFFFFFFFFull - (keys[i] >> 32)) : -1;
    if (out_row) out_row[i] = ok ? (int64_t)key_row(keys[i]) : -1;
  }
  if (out_count) *out_count = cnt;
  return SSW_OK;
}

int ssw_scan_topk(ssw_db* db, const float* queries, int nq, int k, const int32_t* exclude_dbidx,
                  const int64_t* exclude_offsets, int32_t* out_dbidx, float* out_score, int64_t* out_row,
                  int32_t* out_count) {
  return scan_topk_host(db, queries, nq, k, exclude_dbidx, exclude_offsets, out_dbidx, out_score, out_row, out_count,
                        nullptr);
}

int ssw_scan_topk_sharded(ssw_db* db, const float* queries, int nq, int k, const int32_t* exclude_dbidx,
                          const int64_t* exclude_offsets, void* const* peer_bufs, int world, int rank, int nq_cap,
                          int k_cap, uint32_t epoch, int32_t* out_dbidx, float* out_score, int64_t* out_row,
                          int32_t* out_count) {
  SSW_REQUIRE(db != nullptr && peer_bufs != nullptr, "null argument");
  SSW_REQUIRE(world >= 1 && world <= SSW_MAX_WORLD && rank >= 0 && rank < world, "bad world/rank");
  SSW_REQUIRE(nq > 0 && nq <= nq_cap && nq <= db->sm_count, "nq must be in [1, min(nq_cap, SM count)]");
  SSW_REQUIRE(k > 0 && k <= k_cap, "k must be in [1, k_cap]");
  SSW_REQUIRE(epoch != 0, "epoch must be non-zero and increase by one per call");
  XchgCtx xc{peer_bufs, world, rank, nq_cap, k_cap, epoch};
  return scan_topk_host(db, queries, nq, k, exclude_dbidx, exclude_offsets, out_dbidx, out_score, out_row, out_count, &xc);
}

int ssw_score_all_device(ssw_db* db, const float* d_query, float* d_out_scores, void* stream) {
  SSW_REQUIRE(db != nullptr && d_query != nullptr && d_out_scores != nullptr, "null argument");
  SSW_CUDA(cudaSetDevice(db->device));
  return launch_score_all(db, d_query, d_out_scores, (cudaStream_t)stream);
}

int ssw_score_all(ssw_db* db, const float* query, float* out_scores) {
  SSW_REQUIRE(db != nullptr && query != nullptr && out_scores != nullptr, "null argument");
  std::lock_guard<std::mutex> guard(db->mu);
  SSW_CUDA(cudaSetDevice(db->device));
  const size_t q_bytes = ((size_t)db->dim * 4 + 15) / 16 * 16;
  const size_t s_bytes = (size_t)db->n_rows * 4;
  int rc = ensure_stage(db, q_bytes + s_bytes, q_bytes);
  if (rc) return rc;
  memcpy(db->h_stage, query, (size_t)db->dim * 4);
  uint8_t* d = static_cast<uint8_t*>(db->d_stage);
  SSW_CUDA(cudaMemcpyAsync(d, db->h_stage, q_bytes, cudaMemcpyHostToDevice, db->stream));
  rc = launch_score_all(db, reinterpret_cast<const float*>(d), reinterpret_cast<float*>(d + q_bytes), db->stream);
  if (rc) return rc;
  SSW_CUDA(cudaMemcpyAsync(out_scores, d + q_bytes, s_bytes, cudaMemcpyDeviceToHost, db->stream));
  SSW_CUDA(cudaStreamSynchronize(db->stream));
  return SSW_OK;
}

}  // extern "C"
