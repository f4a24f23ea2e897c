// K1: streaming single-query patch scan with a fused epilogue (sm_100a).
//
// Replaces, for one query vector, the reference's
//     scores = vectors @ q ; argsort(-scores) ; dbidx.isin(exclude) ; np.unique first-occurrence ; head(k)
// (seesaw/indices/multiscale/multiscale_index.py:170-199, called from _query_prelim :291-312) and
// the mask/matvec/argsort of CoarseIndex.query (seesaw/indices/coarse/coarse_index.py:57-96).
//
// Data movement: the database is one contiguous [n_rows, dim] array, rows grouped by image.
// Every warp owns a contiguous, image-aligned row range and streams it with 1-D bulk async
// copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) into a private 3-stage ring in
// shared memory: 8 warps x 3 stages x 8 KB = 192 KB in flight per SM, far above the ~45 KB
// Little's law needs at 6.5 TB/s.  No thread ever issues a global load for vector data.
// Compute: conflict-free 128-bit LDS, fp32 FMA against the query held in registers, a
// transposing butterfly so 8 row sums cost 9 shuffles, then a segmented max driven by the
// image-boundary bitmap (per-lane running maxima, combined only when an image ends AND can still
// matter), the exclusion bitmap test and a threshold-filtered insert into the CTA's top-k list.
// A global lower bound on the k-th best key is shared through one L2 word so late candidates are
// rejected with one compare: max over CTAs of their k-th best, and the pooled group-min bound over
// the CTAs' published bests (pooled_group_min, ssw_common.cuh).
// HBM bytes per row: dim*sizeof(T) + 1 bit (image boundary)  -> the roofline in DESIGN.md.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "ssw_db.h"

namespace ssw {

struct Scan1Args {
  const void* vecs;
  const uint32_t* last_bits;  // bit r set <=> device row r is the last row of its image
  const int64_t* row_ptr;
  const int32_t* img_dbidx;
  const int64_t* orig_row;   // may be null (identity)
  const int32_t* part;       // [warps+1]
  const float* q;
  const uint32_t* excl;      // may be null
  uint64_t* list_keys;       // [grid][k]
  int32_t* list_dbidx;       // [grid][k]
  uint64_t* g_thr;
  uint32_t* pub;             // [grid] best score every CTA holds (order-preserving bits), pooled into g_thr
  float* scores_out;
  int64_t row_base;
  int k;
  unsigned long long* stats;  // [2] list updates / images offered (null = not counted)
};

constexpr int kStages = 3;
constexpr int kStageBytesMax = 8192;

template <typename T, int C>
struct Scan1Cfg {
  static constexpr int EPC = 16 / sizeof(T);            // elements per 16-byte chunk
  static constexpr int DIM = C * 32 * EPC;
  static constexpr int ROW_BYTES = DIM * sizeof(T);
  static constexpr int RT_RAW = kStageBytesMax / ROW_BYTES;
  static constexpr int RT = RT_RAW >= 8 ? 8 : (RT_RAW >= 4 ? 4 : 2);   // rows per tile (power of two)
  static constexpr int STAGE_BYTES = RT * ROW_BYTES;
  static constexpr int LANES_PER_ROW = 32 / RT;         // lanes holding one row's total after the butterfly
  static_assert(STAGE_BYTES <= kStageBytesMax, "stage too large");
};

__device__ __forceinline__ void cvt8(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// Sum RT per-lane partials across the warp so that lanes [r*LPR, (r+1)*LPR) all hold row r's total.
template <int RT>
__device__ __forceinline__ float transpose_reduce(float* acc, int lane) {
  float v;
  if constexpr (RT == 8) {
    float a4[4], a2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = b4 ? acc[i] : acc[i + 4];
      float keep = b4 ? acc[i + 4] : acc[i];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = b3 ? a4[i] : a4[i + 2];
      float keep = b3 ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
      float send = b2 ? a2[0] : a2[1];
      float keep = b2 ? a2[1] : a2[0];
      v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
  } else if constexpr (RT == 4) {
    float a2[2];
    const bool b4 = lane & 16, b3 = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = b4 ? acc[i] : acc[i + 2];
      float keep = b4 ? acc[i + 2] : acc[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    {
      float send = b3 ? a2[0] : a2[1];
      float keep = b3 ? a2[1] : a2[0];
      v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
  } else {
    const bool b4 = lane & 16;
    float send = b4 ? acc[0] : acc[1];
    float keep = b4 ? acc[1] : acc[0];
    v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
  }
  return v;
}

// CTA-wide top-k list in shared memory, guarded by a spin lock; warps insert cooperatively.
struct TopkList {
  uint64_t* keys;
  int32_t* dbidx;
  volatile uint64_t* thr;    // k-th best key once the list is full, else 0
  volatile int* cnt;
  volatile int* minpos;
  int* lock;
  uint32_t* best;            // best score held (order-preserving bits)
  uint32_t* pub;             // this CTA's slot of the published bests
};

__device__ __forceinline__ void list_recompute_min(const TopkList& L, int k, int lane, uint64_t* g_thr) {
  uint64_t mn = ~0ull;
  int pos = 0;
  for (int i = lane; i < k; i += 32) {
    uint64_t v = L.keys[i];
    if (v < mn) {
      mn = v;
      pos = i;
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    uint64_t o = shfl_xor_u64(mn, m);
    int op = __shfl_xor_sync(0xffffffffu, pos, m);
    if (o < mn) {
      mn = o;
      pos = op;
    }
  }
  if (lane == 0) {
    *L.minpos = pos;
    *L.thr = mn;
    atomicMax(reinterpret_cast<unsigned long long*>(g_thr), (unsigned long long)mn);
  }
}

// All 32 lanes call this with identical arguments.
__device__ __forceinline__ void list_insert(const TopkList& L, int k, int lane, uint64_t key, int32_t dbidx,
                                            uint64_t* g_thr) {
  if (lane == 0) {
    while (atomicCAS(L.lock, 0, 1) != 0) {
    }
  }
  __syncwarp();
  __threadfence_block();
  // One view of the list state for the whole warp: lane 0's, taken at the shuffle (a sync point),
  // i.e. before lane 0 can run ahead and modify it.  Lanes reading the volatile words on their own
  // could see cnt before/after lane 0's update and take different branches.
  const int cnt = __shfl_sync(0xffffffffu, (int)*L.cnt, 0);
  const uint64_t lthr = shfl_u64(*L.thr, 0);
  const int mpos = __shfl_sync(0xffffffffu, (int)*L.minpos, 0);
  if (lane == 0 && (uint32_t)(key >> 32) > *L.best) {      // a new best of this CTA: let the pooled bound see it
    *L.best = (uint32_t)(key >> 32);
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(L.pub), "r"((uint32_t)(key >> 32)) : "memory");
  }
  if (cnt < k) {
    if (lane == 0) {
      L.keys[cnt] = key;
      L.dbidx[cnt] = dbidx;
      *L.cnt = cnt + 1;
    }
    __syncwarp();
    if (cnt + 1 == k) list_recompute_min(L, k, lane, g_thr);
  } else if (key > lthr) {
    if (lane == 0) {
      L.keys[mpos] = key;
      L.dbidx[mpos] = dbidx;
    }
    __syncwarp();
    list_recompute_min(L, k, lane, g_thr);
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicExch(L.lock, 0);
  }
}

template <typename T, int C, int MODE>
__global__ void __launch_bounds__(kScanWarps * 32, 1) scan1_kernel(const Scan1Args a) {
  using Cfg = Scan1Cfg<T, C>;
  constexpr int RT = Cfg::RT;
  constexpr int EPC = Cfg::EPC;
  constexpr int NQ = C * EPC;   // query values held per lane

  extern __shared__ __align__(128) uint8_t smem[];
  // [warps][stages][STAGE_BYTES] | keys[k] | dbidx[k] | mbar[warps*stages] | ctrl
  uint8_t* ring_base = smem;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem + kScanWarps * kStages * Cfg::STAGE_BYTES);
  int32_t* s_dbidx = reinterpret_cast<int32_t*>(s_keys + a.k);
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(
      smem + kScanWarps * kStages * Cfg::STAGE_BYTES + ((a.k * 12 + 15) / 16) * 16);
  uint64_t* s_thr = s_mbar + kScanWarps * kStages;
  int* s_ctrl = reinterpret_cast<int*>(s_thr + 1);   // cnt, minpos, lock, best

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    *s_thr = 0;
    s_ctrl[0] = 0;
    s_ctrl[1] = 0;
    s_ctrl[2] = 0;
    s_ctrl[3] = 0;
  }
  const uint32_t ring = smem_u32(ring_base + warp * kStages * Cfg::STAGE_BYTES);
  const uint32_t bar0 = smem_u32(s_mbar + warp * kStages);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  TopkList L{s_keys, s_dbidx, s_thr, s_ctrl, s_ctrl + 1, s_ctrl + 2, reinterpret_cast<uint32_t*>(s_ctrl + 3),
             a.pub + blockIdx.x};

  // query slice of this lane: chunk c covers elements (c*32 + lane)*EPC ...
  float qr[NQ];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int e = 0; e < EPC; ++e) qr[c * EPC + e] = __ldg(a.q + (c * 32 + lane) * EPC + e);

  const int gw = blockIdx.x * kScanWarps + warp;
  const int img0 = a.part[gw], img1 = a.part[gw + 1];
  const int64_t r_begin = a.row_ptr[img0], r_end = a.row_ptr[img1];
  const int64_t nrows = r_end - r_begin;
  const int64_t ntiles = (nrows + RT - 1) / RT;
  const uint8_t* gbase = reinterpret_cast<const uint8_t*>(a.vecs) + r_begin * (int64_t)Cfg::ROW_BYTES;

  auto issue = [&](int64_t t) {
    const int s = (int)(t % kStages);
    const int64_t rows = (nrows - t * RT < (int64_t)RT ? nrows - t * RT : (int64_t)RT);
    const uint32_t bytes = (uint32_t)rows * Cfg::ROW_BYTES;
    mbar_expect_tx(bar0 + 8 * s, bytes);
    bulk_g2s(ring + s * Cfg::STAGE_BYTES, gbase + t * (int64_t)Cfg::STAGE_BYTES, bytes, bar0 + 8 * s);
  };
  if (lane == 0) {
    for (int64_t t = 0; t < (ntiles < (int64_t)kStages ? ntiles : (int64_t)kStages); ++t) issue(t);
  }

  // Segmented max without per-row broadcasts: after the butterfly the 32/RT lanes of group r all hold
  // row r's score; every lane keeps the running (max, row) of ITS group's rows of the image being
  // walked, and only when an image ends (warp-uniform, from the boundary bitmap) and some lane's
  // running max reaches the threshold are the groups combined (3 shuffle rounds) and the image offered.
  constexpr int LPR = Cfg::LANES_PER_ROW;
  const int my_r = lane / LPR;                 // row of the tile this lane's group holds
  float run_m = -INFINITY;
  uint32_t run_row = 0;                        // device row of run_m
  int cur_img = img0;
  uint64_t g_cached = 0;

  auto emit = [&](int img, uint64_t key) {
    // key carries the LOCAL device row; quick reject on the score half first.
    // *L.thr changes under other warps' inserts: take lane 0's view so the whole warp makes ONE
    // decision before the warp-collective insert.
    uint64_t thr = *L.thr;
    thr = thr > g_cached ? thr : g_cached;
    thr = shfl_u64(thr, 0);
    if (a.stats && lane == 0) atomicAdd(a.stats + 1, 1ull);
    if ((key >> 32) < (thr >> 32)) return;
    const uint32_t drow = key_row(key);
    const int64_t orow = a.orig_row ? a.orig_row[drow] : (int64_t)drow;
    key = (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(a.row_base + orow));
    if (key <= thr) return;
    if (a.excl && ((a.excl[img >> 5] >> (img & 31)) & 1u)) return;
    if (a.stats && lane == 0) atomicAdd(a.stats, 1ull);
    list_insert(L, a.k, lane, key, a.img_dbidx[img], a.g_thr);
  };

  constexpr int kPoolV = 6;                    // up to 192 CTAs
  int pool_ngp = 1;
  while (pool_ngp < a.k) pool_ngp <<= 1;
  const bool pool_here = MODE == 0 && blockIdx.x == 0 && warp == 0 && pool_ngp <= 128 && (int)gridDim.x >= pool_ngp &&
                         (int)gridDim.x <= 32 * kPoolV;
  uint32_t pool_v[kPoolV] = {0, 0, 0, 0, 0, 0};
  // boundary bits of the next tile, fetched one tile ahead (2 words cover any alignment of 8 rows)
  uint32_t wb0 = 0, wb1 = 0;
  if (MODE == 0 && ntiles > 0) {
    const uint32_t* p = a.last_bits + (r_begin >> 5);
    wb0 = __ldg(p);
    wb1 = __ldg(p + 1);
  }

  for (int64_t t = 0; t < ntiles; ++t) {
    const int s = (int)(t % kStages);
    const uint32_t parity = (uint32_t)((t / kStages) & 1);
    const int64_t row0 = r_begin + t * RT;                 // device row of the tile's first row
    const int rows_here = (int)(nrows - t * RT < (int64_t)RT ? nrows - t * RT : (int64_t)RT);
    uint32_t ends = 0;
    float thr_f = -INFINITY;
    if (MODE == 0) {
      ends = __funnelshift_r(wb0, wb1, (int)(row0 & 31)) & ((1u << rows_here) - 1u);
      if (t + 1 < ntiles) {
        const uint32_t* p = a.last_bits + ((row0 + RT) >> 5);
        wb0 = __ldg(p);
        wb1 = __ldg(p + 1);
      }
      if ((t & 7) == 0) g_cached = ld_relaxed_u64(a.g_thr);
      if (pool_here) {          // one warp of the grid turns the published bests into the shared bound (see pooled_group_min)
        if ((t & 7) == 0) {
#pragma unroll
          for (int m = 0; m < kPoolV; ++m) {
            const int c = lane + 32 * m;
            pool_v[m] = 0u;
            if (c < (int)gridDim.x) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pool_v[m]) : "l"(a.pub + c) : "memory");
          }
        } else if ((t & 7) == 1) {
          const uint32_t tp = pooled_group_min<kPoolV>(pool_v, pool_ngp, lane);
          if (lane == 0 && tp != 0) atomicMax(reinterpret_cast<unsigned long long*>(a.g_thr), (unsigned long long)tp << 32);
        }
      }
      // score an image must reach to matter (possibly stale, i.e. low: the exact test is in emit)
      uint64_t th = *L.thr;
      th = th > g_cached ? th : g_cached;
      thr_f = th == 0 ? -INFINITY : key_score(th);
    }
    mbar_wait(bar0 + 8 * s, parity);

    uint4 raw[RT][C];
    const uint32_t sbase = ring + s * Cfg::STAGE_BYTES + lane * 16;
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < C; ++c) raw[r][c] = lds128(sbase + r * Cfg::ROW_BYTES + c * 512);
    __syncwarp();
    if (lane == 0 && t + kStages < ntiles) issue(t + kStages);
    __syncwarp();

    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      float s0 = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if constexpr (sizeof(T) == 2) {
          float f[8];
          cvt8(raw[r][c], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) s0 = fmaf(f[e], qr[c * 8 + e], s0);
        } else {
          s0 = fmaf(__uint_as_float(raw[r][c].x), qr[c * 4 + 0], s0);
          s0 = fmaf(__uint_as_float(raw[r][c].y), qr[c * 4 + 1], s0);
          s0 = fmaf(__uint_as_float(raw[r][c].z), qr[c * 4 + 2], s0);
          s0 = fmaf(__uint_as_float(raw[r][c].w), qr[c * 4 + 3], s0);
        }
      }
      acc[r] = s0;
    }
    const float total = transpose_reduce<RT>(acc, lane);   // lanes [r*LPR,(r+1)*LPR) hold row r

    if (MODE == 1) {
      const int r = lane / Cfg::LANES_PER_ROW;
      if ((lane % Cfg::LANES_PER_ROW) == 0 && r < rows_here) {
        const int64_t drow = row0 + r;
        const int64_t orow = a.orig_row ? a.orig_row[drow] : drow;
        a.scores_out[orow] = total;
      }
    } else {
      const bool valid = my_r < rows_here;
      const uint32_t my_row = (uint32_t)(row0 + my_r);
      if (ends == 0) {            // warp-uniform fast path: no image ends inside this tile
        if (valid && total > run_m) {
          run_m = total;
          run_row = my_row;
        }
      } else {
        int lo = 0;
        uint32_t e = ends;
        do {                      // warp-uniform walk over the image ends of the tile
          const int p = __ffs(e) - 1;
          e &= e - 1;
          const float x = (valid && my_r >= lo && my_r <= p) ? total : -INFINITY;
          if (x > run_m) {
            run_m = x;
            run_row = my_row;
          }
          if (__any_sync(0xffffffffu, run_m >= thr_f)) {
            // (score desc, row asc) max over the RT groups; empty groups carry key 0
            uint64_t key = run_m > -INFINITY ? make_key(run_m, run_row) : 0ull;
#pragma unroll
            for (int m = LPR; m < 32; m <<= 1) {
              const uint64_t o = shfl_xor_u64(key, m);
              key = o > key ? o : key;
            }
            if (key != 0) emit(cur_img, key);
          }
          run_m = -INFINITY;
          ++cur_img;
          lo = p + 1;
        } while (e);
        if (valid && my_r >= lo && total > run_m) {
          run_m = total;
          run_row = my_row;
        }
      }
    }
  }
  if (MODE == 0) {
    __syncthreads();
    const int cnt = s_ctrl[0];
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) {
      a.list_keys[(int64_t)blockIdx.x * a.k + i] = i < cnt ? s_keys[i] : 0ull;
      a.list_dbidx[(int64_t)blockIdx.x * a.k + i] = i < cnt ? s_dbidx[i] : -1;
    }
  }
}

template <typename T, int C, int MODE>
static int launch_scan1_t(ssw_db* db, const Scan1Args& a, cudaStream_t st) {
  using Cfg = Scan1Cfg<T, C>;
  const size_t smem = (size_t)kScanWarps * kStages * Cfg::STAGE_BYTES + ((a.k * 12 + 15) / 16) * 16 +
                      kScanWarps * kStages * 8 + 8 + 16;      // ... | mbar | thr | cnt, minpos, lock, best
  auto kern = scan1_kernel<T, C, MODE>;
  SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<db->scan_grid, kScanWarps * 32, smem, st>>>(a);
  SSW_LAUNCHED();
  return SSW_OK;
}

template <int MODE>
static int dispatch_scan1(ssw_db* db, const Scan1Args& a, cudaStream_t st, int dtype) {
  const int d = db->dim;
  if (dtype == SSW_F16) {
    switch (d) {
      case 256: return launch_scan1_t<__half, 1, MODE>(db, a, st);
      case 512: return launch_scan1_t<__half, 2, MODE>(db, a, st);
      case 768: return launch_scan1_t<__half, 3, MODE>(db, a, st);
      case 1024: return launch_scan1_t<__half, 4, MODE>(db, a, st);
    }
  } else {
    switch (d) {
      case 256: return launch_scan1_t<float, 2, MODE>(db, a, st);
      case 512: return launch_scan1_t<float, 4, MODE>(db, a, st);
      case 768: return launch_scan1_t<float, 6, MODE>(db, a, st);
      case 1024: return launch_scan1_t<float, 8, MODE>(db, a, st);
    }
  }
  set_error("unsupported dim for the streaming scan (need 256, 512, 768 or 1024)");
  return SSW_ERR_INVALID;
}

int launch_scan1(ssw_db* db, const float* d_query, int k, const uint32_t* d_excl, uint64_t* d_list_keys,
                 int32_t* d_list_dbidx, uint64_t* d_gthr, uint32_t* d_pub, cudaStream_t st, bool exact) {
  Scan1Args a{};
  a.vecs = exact ? static_cast<const void*>(db->d_exact) : db->d_vecs;
  a.last_bits = db->d_last_bits;
  a.row_ptr = db->d_row_ptr;
  a.img_dbidx = db->d_img_dbidx;
  a.orig_row = db->d_orig_row;
  a.part = db->d_part;
  a.q = d_query;
  a.excl = d_excl;
  a.list_keys = d_list_keys;
  a.list_dbidx = d_list_dbidx;
  a.g_thr = d_gthr;
  a.pub = d_pub;
  a.scores_out = nullptr;
  a.row_base = db->row_base;
  a.k = k;
  a.stats = db->d_scan_stats;
  return dispatch_scan1<0>(db, a, st, exact ? (int)SSW_F32 : db->dtype);
}

// With an exact copy attached the full score vector comes from the fp32 rows (index.score feeds label
// propagation priors and active search: the reference's values, not their fp16 roundings).
int launch_score_all(ssw_db* db, const float* d_query, float* d_out, cudaStream_t st) {
  Scan1Args a{};
  a.vecs = db->d_exact ? static_cast<const void*>(db->d_exact) : db->d_vecs;
  a.last_bits = db->d_last_bits;
  a.row_ptr = db->d_row_ptr;
  a.img_dbidx = db->d_img_dbidx;
  a.orig_row = db->d_orig_row;
  a.part = db->d_part;
  a.q = d_query;
  a.scores_out = d_out;
  a.row_base = db->row_base;
  a.k = 0;
  return dispatch_scan1<1>(db, a, st, db->d_exact ? (int)SSW_F32 : db->dtype);
}

// ------------------------------------------------------------------------------------------
// K4: merge candidate lists -> exact top-k (one CTA per query).  Used after the per-CTA lists
// of K1/K2 and after the NCCL all-gather of per-shard lists.
// ------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 1024;
constexpr int kMergeOut = SSW_MAX_TOPK;     // selected candidates are sorted in a second, small buffer

struct MergeArgs {
  const uint64_t* keys;
  const int32_t* dbidx;
  int n_lists;
  int64_t list_stride, query_stride;
  int k;
  const int32_t* counts;     // non-null: query q's candidates are ONE compacted list of counts[q] entries (n_lists = 1)
  const uint64_t* thr;
  uint64_t* out_key;
  int32_t* out_dbidx;
  float* out_score;
  int64_t* out_row;
  int32_t* out_count;
};

// One MSB-first radix-select step over 8 bits: given the histogram of byte d of the keys that match
// the prefix above it, warp 0 finds the bin holding the `remaining`-th largest key.
// `tid`: thread index within the block (or within the 128-thread group of the slim exchange kernel).
__device__ __forceinline__ void merge_pick_bin(const int* hist, int d, uint64_t* s_prefix, int* s_remaining, int tid) {
  if (tid < 32) {
    const int lane = tid;
    // lane l owns bins 255-8l .. 248-8l (descending order of key value)
    int c[8], tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c[i] = hist[255 - 8 * lane - i];
      tot += c[i];
    }
    int incl = tot;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, m);
      if (lane >= m) incl += o;
    }
    const int rem = *s_remaining;
    const int excl = incl - tot;
    __syncwarp();
    if (excl < rem && rem <= incl) {     // exactly one lane
      int need = rem - excl, b = 255 - 8 * lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (need > 0 && c[i] >= need) {
          b = 255 - 8 * lane - i;
          *s_remaining = need;
          need = 0;
        } else if (need > 0) {
          need -= c[i];
        }
      }
      *s_prefix = *s_prefix | ((uint64_t)b << (8 * d));
    }
  }
}

// Scratch of one merge block (dynamic shared memory + a few words of static state).
struct MergeSmem {
  uint64_t* sk;   // [kMergeCap] survivors
  int32_t* sd;
  uint64_t* ok;   // [kMergeOut] selected, then sorted, top-k
  int32_t* od;
  int* cnt;
  int* hist;      // [256]
  uint64_t* prefix;
  int* remaining;
};

// Candidate lists (n_lists x k, list l at keys + l*list_stride) -> the entries with key >= thr gathered in
// shared memory; returns their number (<= kMergeCap).  When more survive than fit, the exact k-th largest
// key is found by radix select over the lists in global memory and only keys >= it are gathered (exactly k:
// every key embeds a distinct row).  `cg`: read through L2 (lists written by a peer GPU).
// `len` = entries per list (the per-CTA lists hold k; a compacted list holds what its counter says), `k` = how
// many the caller wants.
template <bool CG>
__device__ __forceinline__ int merge_gather(const MergeSmem& M, const uint64_t* kq, const int32_t* dq, int n_lists,
                                            int64_t list_stride, uint32_t len, uint32_t k, uint64_t thr) {
  const int tid = threadIdx.x;
  const uint32_t total = (uint32_t)n_lists * len;
  if (thr == 0) thr = 1;   // key 0 == empty slot
  auto ldk = [&](int64_t off) { return CG ? __ldcg(kq + off) : kq[off]; };
  auto ldd = [&](int64_t off) { return CG ? __ldcg(dq + off) : dq[off]; };
  const bool dense = list_stride == (int64_t)len || n_lists == 1;      // lists back to back: entry e sits at offset e
  auto gather = [&](uint64_t lo) {
    if (tid == 0) *M.cnt = 0;
    __syncthreads();
    // four independent loads in flight per thread (the lists sit in L2 / peer-written memory)
    for (uint32_t e0 = tid; e0 < total; e0 += 4 * kMergeThreads) {
      uint64_t key[4];
      int64_t off[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t e = e0 + u * kMergeThreads;
        off[u] = dense ? (int64_t)e : (int64_t)(e / len) * list_stride + (e % len);
        key[u] = e < total ? ldk(off[u]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (key[u] >= lo) {
          const int p = atomicAdd(M.cnt, 1);
          if (p < kMergeCap) {
            M.sk[p] = key[u];
            M.sd[p] = ldd(off[u]);
          }
        }
      }
    }
    __syncthreads();
    return *M.cnt;
  };
  int n = gather(thr);
  if (n > kMergeCap) {
    if (tid == 0) {
      *M.prefix = 0;
      *M.remaining = (int)k;
    }
    for (int d = 7; d >= 0; --d) {
      for (int i = tid; i < 256; i += kMergeThreads) M.hist[i] = 0;
      __syncthreads();
      const uint64_t prefix = *M.prefix;
      for (uint32_t e = tid; e < total; e += kMergeThreads) {
        const int64_t off = dense ? (int64_t)e : (int64_t)(e / len) * list_stride + (e % len);
        const uint64_t key = ldk(off);
        if (key < thr) continue;
        const bool match = (d == 7) || ((key >> (8 * (d + 1))) == (prefix >> (8 * (d + 1))));
        if (match) atomicAdd(&M.hist[(int)((key >> (8 * d)) & 255)], 1);
      }
      __syncthreads();
      merge_pick_bin(M.hist, d, M.prefix, M.remaining, threadIdx.x);
      __syncthreads();
    }
    n = gather(*M.prefix);
  }
  return min(n, kMergeCap);
}

// Few candidates (total <= kMergeThreads): copy every slot into shared memory as it lies — no filter, no
// shared-memory atomics (hundreds of atomicAdds on one counter serialise and cost more than the whole rest of
// the merge); empty slots carry key 0 and are skipped by the ranking.  Returns total.
template <bool CG>
__device__ __forceinline__ int merge_load_direct(const MergeSmem& M, const uint64_t* kq, const int32_t* dq, int n_lists,
                                                 int64_t list_stride, uint32_t len) {
  const uint32_t total = (uint32_t)n_lists * len;
  const bool dense = list_stride == (int64_t)len || n_lists == 1;
  for (uint32_t e = threadIdx.x; e < total; e += kMergeThreads) {
    const int64_t off = dense ? (int64_t)e : (int64_t)(e / len) * list_stride + (e % len);
    M.sk[e] = CG ? __ldcg(kq + off) : kq[off];
    M.sd[e] = CG ? __ldcg(dq + off) : dq[off];
  }
  __syncthreads();
  return (int)total;
}

// The n gathered survivors -> their best min(n, k), sorted best-first in (ok, od); returns that count.
// Keys equal to 0 (empty slots of a direct load) never rank.
__device__ __forceinline__ int merge_select_sort(const MergeSmem& M, int n, int k) {
  const int tid = threadIdx.x;
  int m;
  if (n <= kMergeThreads) {
    // few survivors (the usual case once the pooled thresholds have pruned the lists): every thread ranks
    // one key by counting the larger ones — keys are unique, so the rank IS the output position
    for (int i = tid; i < k && i < kMergeOut; i += kMergeThreads) {
      M.ok[i] = 0;
      M.od[i] = -1;
    }
    __syncthreads();
    const uint64_t mine = tid < n ? M.sk[tid] : 0ull;
    if (mine != 0ull) {
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < n; ++j) rank += M.sk[j] > mine;
      if (rank < k) {
        M.ok[rank] = mine;
        M.od[rank] = M.sd[tid];
      }
    }
    const int valid = __syncthreads_count(mine != 0ull);
    return min(valid, k);
  }
  if (n <= k) {
    m = n;
    for (int i = tid; i < n; i += kMergeThreads) {
      M.ok[i] = M.sk[i];
      M.od[i] = M.sd[i];
    }
    __syncthreads();
  } else {
    if (tid == 0) {
      *M.prefix = 0;
      *M.remaining = k;
    }
    for (int d = 7; d >= 0; --d) {
      for (int i = tid; i < 256; i += kMergeThreads) M.hist[i] = 0;
      __syncthreads();
      const uint64_t prefix = *M.prefix;
      for (int i = tid; i < n; i += kMergeThreads) {
        const uint64_t key = M.sk[i];
        const bool match = (d == 7) || ((key >> (8 * (d + 1))) == (prefix >> (8 * (d + 1))));
        if (match) atomicAdd(&M.hist[(int)((key >> (8 * d)) & 255)], 1);
      }
      __syncthreads();
      merge_pick_bin(M.hist, d, M.prefix, M.remaining, threadIdx.x);
      __syncthreads();
    }
    const uint64_t T = *M.prefix;     // the k-th largest key
    if (tid == 0) *M.cnt = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kMergeThreads) {
      const uint64_t key = M.sk[i];
      if (key >= T) {
        const int p = atomicAdd(M.cnt, 1);
        if (p < kMergeOut) {
          M.ok[p] = key;
          M.od[p] = M.sd[i];
        }
      }
    }
    __syncthreads();
    m = min(*M.cnt, k);
  }
  int P = 1;
  while (P < m) P <<= 1;
  for (int i = m + tid; i < P; i += kMergeThreads) {
    M.ok[i] = 0;
    M.od[i] = -1;
  }
  __syncthreads();
  // bitonic sort, descending (P <= 2048)
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (P >> 1); i += kMergeThreads) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t x = M.ok[lo], y = M.ok[hi];
        if ((x < y) == desc) {
          M.ok[lo] = y;
          M.ok[hi] = x;
          const int32_t t = M.od[lo];
          M.od[lo] = M.od[hi];
          M.od[hi] = t;
        }
      }
      __syncthreads();
    }
  }
  return m;
}

__device__ __forceinline__ void merge_write(const MergeSmem& M, const MergeArgs& a, int q, int cnt) {
  const int tid = threadIdx.x;
  for (int i = tid; i < a.k; i += kMergeThreads) {
    const bool valid = i < cnt;
    const uint64_t key = valid ? M.ok[i] : 0ull;
    const int64_t o = (int64_t)q * a.k + i;
    if (a.out_key) a.out_key[o] = key;
    if (a.out_dbidx) a.out_dbidx[o] = valid ? M.od[i] : -1;
    if (a.out_score) a.out_score[o] = valid ? key_score(key) : -INFINITY;
    if (a.out_row) a.out_row[o] = valid ? (int64_t)key_row(key) : -1;
  }
  if (tid == 0 && a.out_count) a.out_count[q] = cnt;
}

#define SSW_MERGE_SMEM(M)                                                                   \
  extern __shared__ __align__(16) uint8_t msmem[];                                          \
  __shared__ int s_cnt, s_remaining, s_hist[256];                                           \
  __shared__ uint64_t s_prefix;                                                             \
  MergeSmem M;                                                                              \
  M.sk = reinterpret_cast<uint64_t*>(msmem);                                                \
  M.ok = M.sk + kMergeCap;                                                                  \
  M.sd = reinterpret_cast<int32_t*>(M.ok + kMergeOut);                                      \
  M.od = M.sd + kMergeCap;                                                                  \
  M.cnt = &s_cnt;                                                                           \
  M.hist = s_hist;                                                                          \
  M.prefix = &s_prefix;                                                                     \
  M.remaining = &s_remaining

__global__ void __launch_bounds__(kMergeThreads, 1) merge_topk_kernel(const MergeArgs a) {
  SSW_MERGE_SMEM(M);
  pdl_wait();        // launched as a programmatic dependent of the batched scan: its lists are complete from here
  const int q = blockIdx.x;
  const int nl = a.counts ? 1 : a.n_lists;
  const uint32_t len = a.counts ? (uint32_t)a.counts[q] : (uint32_t)a.k;
  const uint64_t* kq = a.keys + (int64_t)q * a.query_stride;
  const int32_t* dq = a.dbidx + (int64_t)q * a.query_stride;
  const int n = (uint64_t)nl * len <= (uint64_t)kMergeThreads
                    ? merge_load_direct<false>(M, kq, dq, nl, a.list_stride, len)
                    : merge_gather<false>(M, kq, dq, nl, a.list_stride, len, (uint32_t)a.k, a.thr ? a.thr[q] : 0ull);
  const int m = merge_select_sort(M, n, a.k);
  merge_write(M, a, q, m);
}

int launch_merge(const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists, int64_t list_stride,
                 int64_t query_stride, int nq, int k, const uint64_t* d_thr, uint64_t* d_out_key,
                 int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                 cudaStream_t st, bool pdl, const int32_t* d_counts) {
  MergeArgs a{d_keys, d_dbidx, n_lists, list_stride, query_stride, k, d_counts, d_thr,
              d_out_key, d_out_dbidx, d_out_score, d_out_row, d_out_count};
  const size_t smem = (size_t)(kMergeCap + kMergeOut) * 12;
  SSW_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SSW_CUDA(launch_kernel(merge_topk_kernel, dim3(nq), dim3(kMergeThreads), smem, st, pdl, a));
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// K4x: the multi-GPU step in ONE kernel — merge this shard's per-CTA lists, store the shard's top-k
// into every peer's exchange buffer over NVLink (peer-mapped pointers), signal, wait for the peers'
// signals, merge the world's lists.  One block per query; block q of a rank depends only on block q
// of the other ranks, and nq <= 148 blocks are all co-resident, so the spin-wait cannot deadlock.
// Exchange buffer of a rank (identical layout everywhere; two parities so a fast rank's next step
// never overwrites what a slow rank still reads):
//   keys  [2][world][nq_cap][k_cap] u64 | dbidx [2][world][nq_cap][k_cap] i32 | flags [2][world][nq_cap] u32
// ------------------------------------------------------------------------------------------
struct XchgArgs {
  MergeArgs m;                 // local lists in, final outputs out
  void* peers[8];              // exchange buffer of every rank, as mapped on this device
  int world, rank, nq_cap, k_cap;
  uint32_t epoch;              // strictly increasing per call, same on all ranks
  int nq;                      // queries of the step (the slim kernel's groups loop over them)
#ifdef SSW_TRACE
  unsigned long long* trace;   // development timeline (ssw_scan_stats block), first 8 blocks: %globaltimer at entry / exit
                               // (CTA slot 160 + parity), end of the push phase and ns spent waiting for flags (slot 162 + parity)
#endif
  int* timed_out;              // set to 1 when a peer's flag did not arrive within kXchgTimeoutNs
};

constexpr unsigned long long kXchgTimeoutNs = 10ull * 1000 * 1000 * 1000;   // a dead peer must not hang the GPU

__host__ __device__ inline size_t xchg_keys_bytes(int world, int nq_cap, int k_cap) { return (size_t)2 * world * nq_cap * k_cap * 8; }
__host__ __device__ inline size_t xchg_dbidx_bytes(int world, int nq_cap, int k_cap) { return (size_t)2 * world * nq_cap * k_cap * 4; }
size_t xchg_bytes(int world, int nq_cap, int k_cap) {
  return xchg_keys_bytes(world, nq_cap, k_cap) + xchg_dbidx_bytes(world, nq_cap, k_cap) + (size_t)2 * world * nq_cap * 4;
}

__global__ void __launch_bounds__(kMergeThreads, 1) exchange_merge_kernel(const XchgArgs x) {
  SSW_MERGE_SMEM(M);
  const MergeArgs& a = x.m;
  const int q = blockIdx.x, tid = threadIdx.x;
  const int par = (int)(x.epoch & 1u);
  pdl_wait();
  // ---- 1. this shard's top-k
  const int nl = a.counts ? 1 : a.n_lists;
  const uint32_t len = a.counts ? (uint32_t)a.counts[q] : (uint32_t)a.k;
  const uint64_t* kq = a.keys + (int64_t)q * a.query_stride;
  const int32_t* dq = a.dbidx + (int64_t)q * a.query_stride;
  int n = (uint64_t)nl * len <= (uint64_t)kMergeThreads
              ? merge_load_direct<false>(M, kq, dq, nl, a.list_stride, len)
              : merge_gather<false>(M, kq, dq, nl, a.list_stride, len, (uint32_t)a.k, a.thr ? a.thr[q] : 0ull);
  int m = merge_select_sort(M, n, a.k);
  // ---- 2. store it into slot `rank` of every rank's buffer (own included), then raise the flags
  const size_t slot = (((size_t)par * x.world + x.rank) * x.nq_cap + q);
  for (int p = 0; p < x.world; ++p) {
    uint8_t* base = static_cast<uint8_t*>(x.peers[p]);
    uint64_t* pk = reinterpret_cast<uint64_t*>(base) + slot * x.k_cap;
    int32_t* pd = reinterpret_cast<int32_t*>(base + xchg_keys_bytes(x.world, x.nq_cap, x.k_cap)) + slot * x.k_cap;
    for (int i = tid; i < a.k; i += kMergeThreads) {
      pk[i] = i < m ? M.ok[i] : 0ull;
      pd[i] = i < m ? M.od[i] : -1;
    }
  }
  // Publication: the block's stores are ordered before the barrier (CTA scope); the few threads that raise the flags
  // then fence at system scope — fences are cumulative, so the whole block's stores are visible to a peer that
  // acquires the flag.  (A system-scope fence in all 1024 threads cost a third of this kernel: ncu, ERRBAR stalls.)
  __syncthreads();
  if (tid < x.world) {
    uint8_t* base = static_cast<uint8_t*>(x.peers[tid]);
    uint32_t* flag = reinterpret_cast<uint32_t*>(base + xchg_keys_bytes(x.world, x.nq_cap, x.k_cap) +
                                                 xchg_dbidx_bytes(x.world, x.nq_cap, x.k_cap)) + slot;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(x.epoch) : "memory");
  }
  // ---- 3. wait until every rank's slot of THIS rank's buffer carries this epoch
  uint8_t* mine = static_cast<uint8_t*>(x.peers[x.rank]);
  if (tid < x.world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + xchg_keys_bytes(x.world, x.nq_cap, x.k_cap) +
                                                             xchg_dbidx_bytes(x.world, x.nq_cap, x.k_cap)) +
                           (((size_t)par * x.world + tid) * x.nq_cap + q);
    uint32_t v;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v != x.epoch) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > kXchgTimeoutNs) {
          if (x.timed_out) *x.timed_out = 1;
          break;
        }
      }
    } while (v != x.epoch);
  }
  __syncthreads();
  // ---- 4. merge the world's lists (read through L2: they were written by other GPUs)
  const size_t q0 = ((size_t)par * x.world) * x.nq_cap + q;
  const uint64_t* wk = reinterpret_cast<const uint64_t*>(mine) + q0 * x.k_cap;
  const int32_t* wd = reinterpret_cast<const int32_t*>(mine + xchg_keys_bytes(x.world, x.nq_cap, x.k_cap)) + q0 * x.k_cap;
  n = (uint64_t)x.world * a.k <= (uint64_t)kMergeThreads
          ? merge_load_direct<true>(M, wk, wd, x.world, (int64_t)x.nq_cap * x.k_cap, (uint32_t)a.k)
          : merge_gather<true>(M, wk, wd, x.world, (int64_t)x.nq_cap * x.k_cap, (uint32_t)a.k, (uint32_t)a.k, 0ull);
  m = merge_select_sort(M, n, a.k);
  merge_write(M, a, q, m);
}

int launch_exchange_merge(const uint64_t* d_keys, const int32_t* d_dbidx, int n_lists, int64_t list_stride,
                          int64_t query_stride, int nq, int k, const uint64_t* d_thr, void* const* peers, int world,
                          int rank, int nq_cap, int k_cap, uint32_t epoch, int* d_timed_out, uint64_t* d_out_key,
                          int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row, int32_t* d_out_count,
                          cudaStream_t st, bool pdl, const int32_t* d_counts) {
  XchgArgs x{};
  x.timed_out = d_timed_out;
  x.m = MergeArgs{d_keys, d_dbidx, n_lists, list_stride, query_stride, k, d_counts, d_thr,
                  d_out_key, d_out_dbidx, d_out_score, d_out_row, d_out_count};
  for (int i = 0; i < world; ++i) x.peers[i] = peers[i];
  x.world = world;
  x.rank = rank;
  x.nq_cap = nq_cap;
  x.k_cap = k_cap;
  x.epoch = epoch;
  const size_t smem = (size_t)(kMergeCap + kMergeOut) * 12;
  SSW_CUDA(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SSW_CUDA(launch_kernel(exchange_merge_kernel, dim3(nq), dim3(kMergeThreads), smem, st, pdl, x));
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// K4x, slim form for the PIPELINED sharded step: the same shard merge + peer exchange + world merge, run on a second
// stream so that the exchange of step i — including the wait for the slowest peer — lies UNDER the scan of step i+1
// instead of after its own.  A query is served by a GROUP of 128 threads (its own named barrier and 2.6 KB of shared
// memory); input is the batched scan's compacted candidate array; k <= 64, world * k <= 1024.  Two launch shapes:
//  * side blocks <S blocks x 8 groups> (default): the pipelined scan runs on (SM count - S) CTAs and these S blocks of
//    1024 threads cannot share an SM with a scan CTA (it leaves 6.4k registers), so whichever kernel is placed first
//    the two end up on DISJOINT SMs: the exchange costs the scan nothing but the S SMs it gave up.
//  * co-resident blocks <nq blocks x 1 group> (side SMs = 0): 128 threads, <= 48 registers, next to a scan CTA on the
//    same SM.  Measured at 8 GPUs: the scan kernel slows from 216 to 243 us with these blocks next to it.
// ------------------------------------------------------------------------------------------
constexpr int kSlimThreads = 128;    // threads of one group

struct SlimSmem {                    // scratch of one group
  uint64_t k[64];                    // survivors of the select
  uint64_t ok[64];                   // the best k, sorted best first
  uint64_t prefix;
  int32_t d[64];
  int32_t od[64];
  int hist[256];
  int cnt, remaining;
};

// barrier of group g (named barriers 1..8; barrier 0 is left to __syncthreads)
__device__ __forceinline__ void slim_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kSlimThreads) : "memory"); }

// The slim kernel must execute few instructions (in its co-resident shape it shares SMs with a scan CTA whose epilogue
// is issue-latency bound): the k best of n keys are found by an 8-pass MSB radix select (n / 128 keys per thread and
// pass) and only those <= k survivors are ranked by counting (k^2 compares), instead of ranking all n (n^2).
// `at(e)` maps entry e to its offset in keys / dbs; tid = thread within the group g.  Leaves the best min(n_valid, k)
// entries, sorted best first, in S.ok / S.od (other slots 0 / -1).  Ends with a group barrier.
template <bool CG, typename At>
__device__ __forceinline__ void slim_topk(const uint64_t* keys, const int32_t* dbs, int n, At at, int k, SlimSmem& S, int tid,
                                          int g) {
  auto ldk = [&](int e) { return CG ? __ldcg(keys + at(e)) : keys[at(e)]; };
  if (tid == 0) {
    S.prefix = 0;
    S.remaining = k;
    S.cnt = 0;
  }
  for (int i = tid; i < k; i += kSlimThreads) {
    S.ok[i] = 0ull;
    S.od[i] = -1;
  }
  slim_sync(g);
  uint64_t T = 1;                        // fewer than k valid keys: keep them all
  int valid = 0;
  for (int e = tid; e < n; e += kSlimThreads) valid += ldk(e) != 0ull;
  if (valid) atomicAdd(&S.cnt, valid);
  slim_sync(g);
  const int n_valid = S.cnt;
  slim_sync(g);
  if (tid == 0) S.cnt = 0;
  if (n_valid > k) {
    for (int d = 7; d >= 0; --d) {
      for (int i = tid; i < 256; i += kSlimThreads) S.hist[i] = 0;
      slim_sync(g);
      const uint64_t prefix = S.prefix;
      for (int e = tid; e < n; e += kSlimThreads) {
        const uint64_t key = ldk(e);
        const bool match = (d == 7) || ((key >> (8 * (d + 1))) == (prefix >> (8 * (d + 1))));
        if (match && key != 0ull) atomicAdd(&S.hist[(int)((key >> (8 * d)) & 255)], 1);
      }
      slim_sync(g);
      merge_pick_bin(S.hist, d, &S.prefix, &S.remaining, tid);
      slim_sync(g);
    }
    T = S.prefix;                        // the k-th largest key (keys are unique)
  }
  slim_sync(g);
  for (int e = tid; e < n; e += kSlimThreads) {
    const uint64_t key = ldk(e);
    if (key != 0ull && key >= T) {
      const int p = atomicAdd(&S.cnt, 1);
      if (p < 64) {
        S.k[p] = key;
        S.d[p] = CG ? __ldcg(dbs + at(e)) : dbs[at(e)];
      }
    }
  }
  slim_sync(g);
  const int m = min(S.cnt, 64);
  for (int i = tid; i < m; i += kSlimThreads) {
    const uint64_t mine = S.k[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += S.k[j] > mine;
    if (rank < k) {
      S.ok[rank] = mine;
      S.od[rank] = S.d[i];
    }
  }
  slim_sync(g);
}

// Group (blockIdx.x, g) serves queries first, first + stride, ...  — the same order on every rank, and a group raises
// the flags of ALL its queries before it waits for any, so the waits cannot deadlock whatever is resident.
template <int GROUPS>
__global__ void __launch_bounds__(kSlimThreads * GROUPS, GROUPS == 1 ? 10 : 1) exchange_slim_kernel(const XchgArgs x) {
  __shared__ SlimSmem s_all[GROUPS];
  const int g = threadIdx.x / kSlimThreads, tid = threadIdx.x % kSlimThreads;
  SlimSmem& S = s_all[g];
  const MergeArgs& a = x.m;
  const int k = a.k, nq = x.nq;
  const int first = blockIdx.x * GROUPS + g, stride = gridDim.x * GROUPS;
  const int par = (int)(x.epoch & 1u);
  const size_t keys_bytes = xchg_keys_bytes(x.world, x.nq_cap, x.k_cap);
  const size_t flags_off = keys_bytes + xchg_dbidx_bytes(x.world, x.nq_cap, x.k_cap);
#ifdef SSW_TRACE
  unsigned long long* tr = (x.trace && threadIdx.x == 0 && blockIdx.x < 8) ? x.trace + 16 + (160 + par) * 16 + 2 * blockIdx.x : nullptr;
  if (tr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[0]));
#endif
  for (int q = first; q < nq; q += stride) {
    // ---- 1. this shard's top-k from its compacted candidates
    const uint64_t* kq = a.keys + (int64_t)q * a.query_stride;
    const int32_t* dq = a.dbidx + (int64_t)q * a.query_stride;
    slim_topk<false>(kq, dq, a.counts[q], [](int e) { return (int64_t)e; }, k, S, tid, g);
    // ---- 2. store it into slot `rank` of every rank's buffer (own included), then raise the flags
    const size_t slot = (((size_t)par * x.world + x.rank) * x.nq_cap + q);
    for (int p = 0; p < x.world; ++p) {
      uint8_t* base = static_cast<uint8_t*>(x.peers[p]);
      uint64_t* pk = reinterpret_cast<uint64_t*>(base) + slot * x.k_cap;
      int32_t* pd = reinterpret_cast<int32_t*>(base + keys_bytes) + slot * x.k_cap;
      for (int i = tid; i < k; i += kSlimThreads) {
        pk[i] = S.ok[i];
        pd[i] = S.od[i];
      }
    }
    slim_sync(g);       // the group's stores are ordered before the flag threads' system fence (fences are cumulative)
    if (tid < x.world) {
      uint32_t* flag = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(x.peers[tid]) + flags_off) + slot;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(x.epoch) : "memory");
    }
  }
  uint8_t* mine = static_cast<uint8_t*>(x.peers[x.rank]);
#ifdef SSW_TRACE
  if (tr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[32]));
  unsigned long long waited = 0;
#endif
  for (int q = first; q < nq; q += stride) {
    // ---- 3. wait until every rank's slot of THIS rank's buffer carries this epoch
#ifdef SSW_TRACE
    unsigned long long w0 = 0, w1 = 0;
    if (tr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w0));
#endif
    if (tid < x.world) {
      const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + flags_off) + (((size_t)par * x.world + tid) * x.nq_cap + q);
      uint32_t v;
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v != x.epoch) {
          __nanosleep(500);       // co-resident shape: the next step's scan shares this SM, do not burn its issue slots
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > kXchgTimeoutNs) {
            if (x.timed_out) *x.timed_out = 1;
            break;
          }
        }
      } while (v != x.epoch);
    }
    slim_sync(g);
#ifdef SSW_TRACE
    if (tr) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w1));
      waited += w1 - w0;
    }
#endif
    // ---- 4. merge the world's lists (read through L2: they were written by other GPUs)
    const size_t q0 = ((size_t)par * x.world) * x.nq_cap + q;
    const uint64_t* wk = reinterpret_cast<const uint64_t*>(mine) + q0 * x.k_cap;
    const int32_t* wd = reinterpret_cast<const int32_t*>(mine + keys_bytes) + q0 * x.k_cap;
    const int64_t lstride = (int64_t)x.nq_cap * x.k_cap;
    slim_topk<true>(wk, wd, x.world * k, [=](int e) { return (int64_t)(e / k) * lstride + (e % k); }, k, S, tid, g);
    int cnt = 0;
    for (int i = tid; i < k; i += kSlimThreads) {
      const uint64_t key = S.ok[i];
      const bool valid = key != 0ull;
      cnt += valid;
      const int64_t o = (int64_t)q * k + i;
      if (a.out_key) a.out_key[o] = key;
      if (a.out_dbidx) a.out_dbidx[o] = valid ? S.od[i] : -1;
      if (a.out_score) a.out_score[o] = valid ? key_score(key) : -INFINITY;
      if (a.out_row) a.out_row[o] = valid ? (int64_t)key_row(key) : -1;
    }
    // count of valid slots: group-wide sum of the per-thread partials
    if (tid == 0) S.cnt = 0;
    slim_sync(g);
    if (cnt) atomicAdd(&S.cnt, cnt);
    slim_sync(g);
    if (tid == 0 && a.out_count) a.out_count[q] = S.cnt;
    slim_sync(g);       // S is reused by the group's next query
  }
#ifdef SSW_TRACE
  if (tr) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[1]));
    tr[33] = waited;
  }
#endif
}

#ifdef SSW_TRACE
unsigned long long* g_trace_block = nullptr;      // set by ssw_scan_stats
#endif

constexpr int kSideGroups = 8;       // groups per side block: 1024 threads, too many to fit next to a scan CTA

// side_blocks > 0: that many blocks of kSideGroups groups (disjoint SMs from the pipelined scan); 0: one co-resident
// 128-thread block per query.
int launch_exchange_slim(const uint64_t* d_keys, const int32_t* d_dbidx, int64_t query_stride, int nq, int k,
                         const int32_t* d_counts, void* const* peers, int world, int rank, int nq_cap, int k_cap, uint32_t epoch,
                         int* d_timed_out, uint64_t* d_out_key, int32_t* d_out_dbidx, float* d_out_score, int64_t* d_out_row,
                         int32_t* d_out_count, cudaStream_t st, int side_blocks) {
  XchgArgs x{};
  x.timed_out = d_timed_out;
  x.m = MergeArgs{d_keys, d_dbidx, 1, query_stride, query_stride, k, d_counts, nullptr,
                  d_out_key, d_out_dbidx, d_out_score, d_out_row, d_out_count};
  for (int i = 0; i < world; ++i) x.peers[i] = peers[i];
  x.world = world;
  x.rank = rank;
  x.nq_cap = nq_cap;
  x.k_cap = k_cap;
  x.epoch = epoch;
  x.nq = nq;
#ifdef SSW_TRACE
  x.trace = g_trace_block;
  if (getenv("SSW_SIDE_CARVEOUT")) {     // experiment: same shared-memory configuration as the scan kernel
    cudaFuncSetAttribute(exchange_slim_kernel<kSideGroups>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("SSW_SIDE_CARVEOUT")));
    cudaFuncSetAttribute(exchange_slim_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("SSW_SIDE_CARVEOUT")));
  }
#endif
  if (side_blocks > 0)
    exchange_slim_kernel<kSideGroups><<<side_blocks, kSlimThreads * kSideGroups, 0, st>>>(x);
  else
    exchange_slim_kernel<1><<<nq, kSlimThreads, 0, st>>>(x);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// scan partition for a grid other than the database's own (the pipelined sharded step scans on SM count - S CTAs):
// the same rule as build_layout — range g of `ranges` starts at the first image whose first row is >= n_rows * g / ranges —
// evaluated on the device from the CSR.  out_max_cta: most images any CTA (kScanWarps consecutive ranges) holds.
// ------------------------------------------------------------------------------------------
__global__ void part_build_kernel(const int64_t* __restrict__ row_ptr, int64_t n_images, int64_t n_rows, int ranges,
                                  int32_t* __restrict__ part, int* __restrict__ out_max_cta) {
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  for (int g = threadIdx.x; g <= ranges; g += blockDim.x) {
    const int64_t target = n_rows * g / ranges;          // n_rows < 2^32, ranges < 2^12
    int64_t lo = 0, hi = n_images + 1;                   // first position of row_ptr[0 .. n_images] with value >= target
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (row_ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (lo > n_images) lo = n_images;
    if (g == 0) lo = 0;
    if (g == ranges) lo = n_images;
    part[g] = (int32_t)lo;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ranges / kScanWarps; c += blockDim.x)
    atomicMax(&s_max, part[(c + 1) * kScanWarps] - part[c * kScanWarps]);
  __syncthreads();
  if (threadIdx.x == 0) *out_max_cta = s_max;
}

int launch_part_build(ssw_db* db, int grid, int32_t* d_part, int* d_max_cta, cudaStream_t st) {
  part_build_kernel<<<1, 1024, 0, st>>>(db->d_row_ptr, db->n_images, db->n_rows, grid * kScanWarps, d_part, d_max_cta);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// exclusion bitmaps: per query, bit i set <=> local image i is excluded
// (replaces pr.BitMap difference / DataFrame.isin, multiscale_index.py:192-193, 295)
// ------------------------------------------------------------------------------------------
__global__ void exclude_build_kernel(const int32_t* ids, const int64_t* offsets, const int32_t* img_dbidx,
                                     int64_t n_images, int64_t words, uint32_t* bits) {
  const int q = blockIdx.y;
  const int64_t b = offsets[q], e = offsets[q + 1];
  for (int64_t i = b + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t id = ids[i];
    // image ids are usually consecutive (dbidx = position in the dataset): try that slot first — one load instead
    // of log2(n_images) dependent ones (the search was 8 of the ~55 us the host-buffer path adds to a step)
    int64_t lo = (int64_t)id - (int64_t)img_dbidx[0];
    if (lo < 0 || lo >= n_images || img_dbidx[lo] != id) {
      lo = 0;
      int64_t hi = n_images;            // first position with img_dbidx >= id
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (img_dbidx[mid] < id) lo = mid + 1; else hi = mid;
      }
    }
    if (lo < n_images && img_dbidx[lo] == id) atomicOr(bits + (int64_t)q * words + (lo >> 5), 1u << (lo & 31));
  }
}

int launch_exclude_build(ssw_db* db, const int32_t* d_ids, const int64_t* d_offsets, int nq, uint32_t* d_bits,
                         cudaStream_t st) {
  SSW_CUDA(cudaMemsetAsync(d_bits, 0, (size_t)nq * db->excl_words * 4, st));
  if (db->n_images == 0) return SSW_OK;
  dim3 grid(8, nq);
  exclude_build_kernel<<<grid, 256, 0, st>>>(d_ids, d_offsets, db->d_img_dbidx, db->n_images, db->excl_words, d_bits);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// K5: per-image best key from a caller-supplied score per row (scores that are not a dot product:
// label propagation, KnnProp2.next_batch -> _get_top_dbidxs, seesaw/loops/graph_based.py:97-99).
// One thread per image walks the image's rows; the top-k over the image keys is K4's job.
// ------------------------------------------------------------------------------------------
// `pos` non-null: rank by the caller's ORDER instead of a score — pos[o] = position of original row o in the caller's
// best-first row list (0xFFFFFFFF = not listed); an image's best row is its earliest-listed one.
__global__ void image_max_kernel(const float* __restrict__ scores, const uint32_t* __restrict__ pos,
                                 const uint8_t* __restrict__ row_mask,
                                 const int64_t* __restrict__ row_ptr, const int64_t* __restrict__ orig_row,
                                 const int32_t* __restrict__ img_dbidx, const uint32_t* __restrict__ excl,
                                 int64_t row_base, int64_t n_images, int64_t n_padded, uint64_t* __restrict__ keys_out,
                                 int32_t* __restrict__ dbidx_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_padded; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t best = 0;
    int32_t id = -1;
    if (i < n_images && !(excl && ((excl[i >> 5] >> (i & 31)) & 1u))) {
      id = img_dbidx[i];
      for (int64_t r = row_ptr[i]; r < row_ptr[i + 1]; ++r) {
        const int64_t o = orig_row ? orig_row[r] : r;
        if (row_mask && !row_mask[o]) continue;
        uint64_t key;
        if (pos) {
          const uint32_t p = pos[o];
          if (p == 0xFFFFFFFFu) continue;
          key = ((uint64_t)(0xFFFFFFFEu - p) + 1ull) << 32 | (uint64_t)(0xFFFFFFFFu - (uint32_t)(row_base + o));
        } else {
          const float sc = scores[o];
          if (sc != sc) continue;                       // NaN never ranks
          key = make_key(sc, (uint32_t)(row_base + o));
        }
        best = key > best ? key : best;
      }
    }
    keys_out[i] = best;
    dbidx_out[i] = best ? id : -1;
  }
}

__global__ void order_to_pos_kernel(const int64_t* __restrict__ order, int64_t n_order, int64_t n_rows, uint32_t* __restrict__ pos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_order; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = order[p];
    if (o >= 0 && o < n_rows) atomicMin(pos + o, (uint32_t)p);      // a row listed twice counts at its first position
  }
}

int launch_order_to_pos(const int64_t* d_order, int64_t n_order, int64_t n_rows, uint32_t* d_pos, cudaStream_t st) {
  SSW_CUDA(cudaMemsetAsync(d_pos, 0xFF, (size_t)std::max<int64_t>(n_rows, 1) * 4, st));
  if (n_order == 0) return SSW_OK;
  const int grid = (int)std::min<int64_t>((n_order + 255) / 256, 148 * 16);
  order_to_pos_kernel<<<grid, 256, 0, st>>>(d_order, n_order, n_rows, d_pos);
  SSW_LAUNCHED();
  return SSW_OK;
}

int launch_image_max(ssw_db* db, const float* d_scores, const uint8_t* d_row_mask, const uint32_t* d_excl,
                     int64_t n_padded, uint64_t* d_keys, int32_t* d_dbidx, cudaStream_t st, const uint32_t* d_pos) {
  if (n_padded == 0) return SSW_OK;
  const int grid = (int)std::min<int64_t>((n_padded + 255) / 256, 148 * 16);
  image_max_kernel<<<grid, 256, 0, st>>>(d_scores, d_pos, d_row_mask, db->d_row_ptr, db->d_orig_row, db->d_img_dbidx, d_excl,
                                         db->row_base, db->n_images, n_padded, d_keys, d_dbidx);
  SSW_LAUNCHED();
  return SSW_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic rows (bit-identical to seesaw_b200/synth.py) and dtype conversion
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T>
__global__ void synth_kernel(T* out, int64_t n_quads, int quads_per_row, int64_t global_row0, uint64_t seed, int kind) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_quads; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = (uint64_t)(global_row0 * quads_per_row + i) + seed * 0x9E3779B97F4A7C15ull;
    const uint64_t h = splitmix(ctr);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pair = (int)((h >> (16 * j)) & 255) + (int)((h >> (16 * j + 8)) & 255);
      v[j] = kind == SSW_SYNTH_TRI ? (float)(pair - 255) * (1.0f / 2048.0f) : (float)(pair % 9 - 4) * 0.125f;
    }
    if constexpr (sizeof(T) == 2) {
      __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(out)[i] = w;
    } else {
      reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int launch_synth(void* d_out, int dtype, int64_t n_rows, int dim, int64_t global_row0, uint64_t seed, int kind,
                 cudaStream_t st) {
  const int qpr = dim / 4;
  const int64_t n_quads = n_rows * qpr;
  const int grid = (int)std::min<int64_t>((n_quads + 255) / 256, 148 * 32);
  if (n_quads == 0) return SSW_OK;
  if (dtype == SSW_F16)
    synth_kernel<__half><<<grid, 256, 0, st>>>((__half*)d_out, n_quads, qpr, global_row0, seed, kind);
  else
    synth_kernel<float><<<grid, 256, 0, st>>>((float*)d_out, n_quads, qpr, global_row0, seed, kind);
  SSW_LAUNCHED();
  return SSW_OK;
}

template <typename S, typename D>
__global__ void convert_kernel(const S* src, D* dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if constexpr (sizeof(S) == 4 && sizeof(D) == 2) dst[i] = __float2half_rn(src[i]);
    else if constexpr (sizeof(S) == 2 && sizeof(D) == 4) dst[i] = __half2float(src[i]);
    else dst[i] = src[i];
  }
}

int launch_convert_rows(const void* d_src, int dtype_src, void* d_dst, int dtype_dst, int64_t n, cudaStream_t st) {
  if (n == 0) return SSW_OK;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 32);
  if (dtype_src == SSW_F32 && dtype_dst == SSW_F16)
    convert_kernel<float, __half><<<grid, 256, 0, st>>>((const float*)d_src, (__half*)d_dst, n);
  else if (dtype_src == SSW_F16 && dtype_dst == SSW_F32)
    convert_kernel<__half, float><<<grid, 256, 0, st>>>((const __half*)d_src, (float*)d_dst, n);
  else if (dtype_src == SSW_F32)
    convert_kernel<float, float><<<grid, 256, 0, st>>>((const float*)d_src, (float*)d_dst, n);
  else
    convert_kernel<__half, __half><<<grid, 256, 0, st>>>((const __half*)d_src, (__half*)d_dst, n);
  SSW_LAUNCHED();
  return SSW_OK;
}

}  // namespace ssw
