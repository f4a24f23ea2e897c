// K2: batched patch scan on tcgen05/TMEM — up to 64 concurrent session queries per pass.
//
// Same semantics as K1 (ssw_scan.cu) for every query of the batch; replaces 64 sequential runs of
//   scores = vectors @ q ; argsort ; isin(exclude) ; first occurrence per dbidx ; head(k)
// (seesaw/indices/multiscale/multiscale_index.py:170-199, 291-312 — the reference has no batching,
// each session process scans on its own: SURVEY.md §8b).
//
// The database is read from HBM ONCE per batch (arithmetic intensity 64 flop/B would make 64 CUDA-core
// passes compute-bound; on the tensor pipe the pass stays HBM-bound).  Orientation: queries are the
// resident A operand in TMEM, database rows are the streamed B operand, so accumulator lane = query
// and accumulator column = database row: every epilogue thread owns one query and walks the rows of
// the tile in order — the per-image segmented max is a running max, image boundaries are warp-uniform.
// fp32 queries are split into hi + lo fp16 parts after scaling by a power of two (keeps ~22 mantissa
// bits; exact for fp16-representable queries): lane 32w+j holds the hi part of query 16w+j, lane
// 32w+16+j its lo part, and the two partial dot products meet with one shfl_xor(16).
#include <algorithm>

#include "ssw_db.h"
#include "ssw_tc.cuh"

namespace ssw {

struct ScanTcArgs {
  const float* q;            // [nq, DIM] fp32
  int nq;
  int k;
  const uint32_t* excl;      // [nq, excl_words] or null
  int64_t excl_words;
  const int32_t* img_of_row; // [n_rows + 1], sentinel -1
  const int64_t* row_ptr;
  const int32_t* img_dbidx;
  const int64_t* orig_row;   // null when identity
  const int32_t* part;       // [grid*8 + 1] image partition of the streaming scan, 8 entries per CTA
  uint64_t* list_keys;       // [nq][grid][k]
  int32_t* list_dbidx;
  uint64_t* g_thr;           // [nq]
  int64_t row_base;
};

constexpr int kTcMaxK = 64;   // per-query list length the batched epilogue keeps in shared memory

struct QList {
  uint64_t* keys;   // slot s of query qi at keys[s * 64 + qi]
  int32_t* dbidx;
};

template <int DIM, int NT, int NS>
__global__ void __launch_bounds__(kTcThreads, 1) scan_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                 const ScanTcArgs a) {
  using Cfg = TcCfg<DIM, NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem;
  const TcSmem S = tc_carve(smem_raw, NS, Cfg::STAGE_BYTES, &smem);
  uint8_t* after = smem + NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16;
  QList L{reinterpret_cast<uint64_t*>(after), reinterpret_cast<int32_t*>(after + (size_t)a.k * 64 * 8)};
  const uint32_t tmem = tc_setup(S, NS, Cfg::TMEM_ALLOC, &tmap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int img0 = a.part[blockIdx.x * kScanWarps], img1 = a.part[(blockIdx.x + 1) * kScanWarps];
  const int64_t r_begin = a.row_ptr[img0], r_end = a.row_ptr[img1];
  const int64_t nrows = r_end - r_begin;
  const int ntiles = (int)((nrows + NT - 1) / NT);

  if (warp == 0) {
    if (lane == 0) {
      TcPipe p(NS);
      for (int t = 0; t < ntiles; ++t) {
        for (int kc = 0; kc < Cfg::KC; ++kc) {
          mbar_wait(S.empty + 8 * p.stage, p.phase ^ 1);
          mbar_expect_tx(S.full + 8 * p.stage, Cfg::STAGE_BYTES);
          tma_load_2d(S.stages + p.stage * Cfg::STAGE_BYTES, &tmap, kc * kTcKChunk, (int)(r_begin + (int64_t)t * NT),
                      S.full + 8 * p.stage);
          p.advance();
        }
      }
    }
  } else if (warp == 1) {
    TcPipe p(NS);
    mbar_wait(S.a_ready, 0);
    tc_fence_after();
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = t & 1;
      mbar_wait(S.tmem_empty + 8 * as, ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < Cfg::KC; ++kc) {
        mbar_wait(S.full + 8 * p.stage, p.phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t bdesc = make_bdesc_sw128(S.stages + p.stage * Cfg::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ts(tmem + Cfg::ACC_BASE + as * NT, tmem + Cfg::A_BASE + kc * 32 + k * 8, bdesc + 2 * k, Cfg::IDESC,
                       (kc | k) != 0);
          tc_commit(S.empty + 8 * p.stage);
        }
        __syncwarp();
        p.advance();
      }
      if (lane == 0) tc_commit(S.tmem_full + 8 * as);
      __syncwarp();
    }
  } else {
    // ===== epilogue: lane j<16 of quarter q4 owns query 16*q4 + j (hi part), lane j+16 its lo part
    const int q4 = warp & 3;
    const int qi = q4 * 16 + (lane & 15);
    const bool is_lo = lane >= 16;
    const bool q_ok = qi < a.nq;
    const bool owner = q_ok && !is_lo;                  // the thread that keeps this query's list
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    const int k = a.k;

    // ---- A operand: scale by 2^e so the largest |q| lands in [1024, 2048), split hi/lo, store
    float inv_scale = 1.0f;
    {
      const float* qp = a.q + (size_t)(q_ok ? qi : 0) * DIM;
      float mx = 0.f;
      if (q_ok)
        for (int i = 0; i < DIM; ++i) mx = fmaxf(mx, fabsf(__ldg(qp + i)));
      int e = 0;
      if (mx > 0.f && mx < INFINITY) {
        int ex;
        frexpf(mx, &ex);          // mx = m * 2^ex, m in [0.5, 1)
        e = 11 - ex;              // mx * 2^e in [1024, 2048)
      }
      const float scale = ldexpf(1.0f, e);
      inv_scale = ldexpf(1.0f, -e);
#pragma unroll 1
      for (int c = 0; c < Cfg::A_COLS / 32; ++c) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x0 = 0.f, x1 = 0.f;
          if (q_ok) {
            x0 = __ldg(qp + c * 64 + 2 * j) * scale;
            x1 = __ldg(qp + c * 64 + 2 * j + 1) * scale;
          }
          __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
          if (is_lo) {
            h0 = __float2half_rn(x0 - __half2float(h0));
            h1 = __float2half_rn(x1 - __half2float(h1));
          }
          r[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        }
        tmem_st32(lane_addr + Cfg::A_BASE + c * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(S.a_ready);
    }

    int cnt = 0, minpos = 0;
    uint64_t minkey = 0;            // valid once cnt == k
    uint64_t thr = 0;               // reject keys <= thr
    float run_max = -INFINITY;
    int run_col = 0;                // row (relative to r_begin) of the running max
    const uint32_t* my_excl = (a.excl && q_ok) ? a.excl + (size_t)qi * a.excl_words : nullptr;

    // candidate of one finished image: (score, device row) -> key; insert into this query's list
    auto flush = [&](float smax, int64_t drow) {
      if (!owner) return;
      const float sc = smax * inv_scale;
      uint64_t key = make_key(sc, (uint32_t)drow);
      if ((key >> 32) < (thr >> 32)) return;
      const int64_t orow = a.orig_row ? a.orig_row[drow] : drow;
      key = (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(a.row_base + orow));
      if (key <= thr) return;
      const int img = a.img_of_row[drow];
      if (my_excl && ((my_excl[img >> 5] >> (img & 31)) & 1u)) return;
      const int32_t dbi = a.img_dbidx[img];
      if (cnt < k) {
        L.keys[cnt * 64 + qi] = key;
        L.dbidx[cnt * 64 + qi] = dbi;
        if (++cnt < k) return;
      } else {
        L.keys[minpos * 64 + qi] = key;
        L.dbidx[minpos * 64 + qi] = dbi;
      }
      uint64_t mk = ~0ull;
      int mp = 0;
      for (int s = 0; s < k; ++s) {
        const uint64_t x = L.keys[s * 64 + qi];
        if (x < mk) {
          mk = x;
          mp = s;
        }
      }
      minkey = mk;
      minpos = mp;
      if (mk > thr) thr = mk;
      atomicMax(reinterpret_cast<unsigned long long*>(a.g_thr + qi), (unsigned long long)mk);
    };

    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = t & 1;
      const int64_t row0 = r_begin + (int64_t)t * NT;
      const int valid = (int)min((int64_t)NT, r_end - row0);
      // image-boundary bits of this tile: bit c set <=> row0+c is the last row of its image
      uint32_t endmask[NT / 32];
#pragma unroll
      for (int g = 0; g < NT / 32; ++g) {
        const int c = g * 32 + lane;
        bool e = false;
        if (c < valid) e = a.img_of_row[row0 + c] != a.img_of_row[row0 + c + 1];
        endmask[g] = __ballot_sync(0xffffffffu, e);
      }
      if (owner) {
        const uint64_t g = ld_relaxed_u64(a.g_thr + qi);
        if (g > thr) thr = g;
      }
      mbar_wait(S.tmem_full + 8 * as, (t >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < NT / 32; ++g) {
        uint32_t v[32];
        tmem_ld32(lane_addr + Cfg::ACC_BASE + as * NT + g * 32, v);
        tmem_ld_wait();
        const uint32_t em = endmask[g];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float mine = __uint_as_float(v[i]);
          const float s = mine + __shfl_xor_sync(0xffffffffu, mine, 16);     // hi + lo partial dots
          if (s > run_max) {      // strict: the first (lowest) row wins ties
            run_max = s;
            run_col = (int)(row0 - r_begin) + g * 32 + i;
          }
          if ((em >> i) & 1u) {   // warp-uniform
            flush(run_max, r_begin + run_col);
            run_max = -INFINITY;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(S.tmem_empty + 8 * as);
    }
    // ---- publish this CTA's list of every query
    if (owner) {
      const int64_t o = ((int64_t)qi * gridDim.x + blockIdx.x) * k;
      for (int s = 0; s < k; ++s) {
        a.list_keys[o + s] = s < cnt ? L.keys[s * 64 + qi] : 0ull;
        a.list_dbidx[o + s] = s < cnt ? L.dbidx[s * 64 + qi] : -1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_ALLOC);
}

template <int DIM, int NT, int NS>
static int launch_scan_tc_t(ssw_db* db, const ScanTcArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<DIM, NT>;
  CUtensorMap tmap;
  int rc = make_tmap_f16_rows(&tmap, db->d_vecs, db->n_rows, DIM, NT);
  if (rc) return rc;
  const size_t smem = (size_t)NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16 + (size_t)a.k * 64 * 12 +
                      tc_smem_slack;
  auto kern = scan_tc_kernel<DIM, NT, NS>;
  SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<db->scan_grid, kTcThreads, smem, st>>>(tmap, a);
  SSW_LAUNCHED();
  return SSW_OK;
}

bool scan_tc_supported(const ssw_db* db, int k) {
  return db->dtype == SSW_F16 && (db->dim == 256 || db->dim == 512 || db->dim == 768) && k <= kTcMaxK;
}

// One pass over the database for queries [0, nq), nq <= 64.
int launch_scan_tc(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_excl, uint64_t* d_list_keys,
                   int32_t* d_list_dbidx, uint64_t* d_gthr, cudaStream_t st) {
  ScanTcArgs a{};
  a.q = d_queries;
  a.nq = nq;
  a.k = k;
  a.excl = d_excl;
  a.excl_words = db->excl_words;
  a.img_of_row = db->d_img_of_row;
  a.row_ptr = db->d_row_ptr;
  a.img_dbidx = db->d_img_dbidx;
  a.orig_row = db->d_orig_row;
  a.part = db->d_part;
  a.list_keys = d_list_keys;
  a.list_dbidx = d_list_dbidx;
  a.g_thr = d_gthr;
  a.row_base = db->row_base;
  // shared memory: NS stages of NT*128 B + 64 lists of k (key, dbidx) pairs (k <= 64 -> <= 48 KB)
  switch (db->dim) {
    case 256: return launch_scan_tc_t<256, 128, 10>(db, a, st);
    case 512: return launch_scan_tc_t<512, 128, 10>(db, a, st);
    case 768: return launch_scan_tc_t<768, 64, 20>(db, a, st);
  }
  set_error("batched scan supports dim 256, 512 or 768");
  return SSW_ERR_INVALID;
}

}  // namespace ssw
