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

// Per-query epilogue state in SHARED memory (structure of arrays over the 64 query slots).  It used
// to be a struct in local memory: with 214 KB of the SM carved out as shared memory the L1 left for
// local memory thrashes and every access cost an L2 round trip (ncu: LDL-dependent stalls, 2.7 us per
// image boundary).
struct QShared {
  uint64_t* keys;       // [k][64]   candidate keys, slot s of query q at keys[s*64 + q]
  int32_t* img;         // [k][64]   local image index of the candidate
  uint64_t* thr;        // [64] reject keys <= thr: max(own k-th best once full, shared lower bound)
  uint64_t* pend_key;   // [64] candidate whose exclusion word is still on its way (0 = none)
  int32_t* pend_img;    // [64]
  int32_t* cnt;         // [64]
  int32_t* minpos;      // [64]
  float* inv_scale;     // [64]
};
__host__ __device__ constexpr size_t qshared_bytes(int k) { return (size_t)k * 64 * 12 + 64 * (8 + 8 + 4 + 4 + 4 + 4); }

__device__ __forceinline__ QShared qshared_carve(uint8_t* base, int k) {
  QShared q;
  q.keys = reinterpret_cast<uint64_t*>(base);
  q.thr = q.keys + (size_t)k * 64;
  q.pend_key = q.thr + 64;
  q.img = reinterpret_cast<int32_t*>(q.pend_key + 64);
  q.pend_img = q.img + (size_t)k * 64;
  q.cnt = q.pend_img + 64;
  q.minpos = q.cnt + 64;
  q.inv_scale = reinterpret_cast<float*>(q.minpos + 64);
  return q;
}

__device__ __forceinline__ void qlist_insert(const QShared& Q, const ScanTcArgs& a, int q, int k, uint64_t key, int img) {
  if (key <= Q.thr[q]) return;
  const int cnt = Q.cnt[q];
  if (cnt < k) {
    Q.keys[cnt * 64 + q] = key;
    Q.img[cnt * 64 + q] = img;
    Q.cnt[q] = cnt + 1;
    if (cnt + 1 < k) return;
  } else {
    const int mp = Q.minpos[q];
    Q.keys[mp * 64 + q] = key;
    Q.img[mp * 64 + q] = img;
  }
  uint64_t mk = ~0ull;
  int mp = 0;
  for (int s = 0; s < k; ++s) {
    const uint64_t x = Q.keys[s * 64 + q];
    if (x < mk) {
      mk = x;
      mp = s;
    }
  }
  Q.minpos[q] = mp;
  if (mk > Q.thr[q]) Q.thr[q] = mk;
  atomicMax(reinterpret_cast<unsigned long long*>(a.g_thr + q), (unsigned long long)mk);
}

__device__ __forceinline__ void qlist_resolve_pending(const QShared& Q, const ScanTcArgs& a, int q, int k) {
  const uint64_t pk = Q.pend_key[q];
  if (pk == 0) return;
  const int img = Q.pend_img[q];
  const uint32_t w = a.excl[(size_t)q * a.excl_words + (img >> 5)];    // prefetched when the candidate was parked
  Q.pend_key[q] = 0;
  if (!((w >> (img & 31)) & 1u)) qlist_insert(Q, a, q, k, pk, img);
}

// Candidate of one finished image (max score, device row holding it, image index) -> the owner's
// top-k list.  The exclusion test needs one word of a 2 MB table: instead of stalling on it, the
// candidate is parked with a prefetch and settled at the next call.  Out of line: once per image.
__device__ __noinline__ void scan_tc_flush(uint8_t* qbase, const ScanTcArgs& a, int q, float smax, int64_t drow, int img) {
  const int k = a.k;
  const QShared Q = qshared_carve(qbase, k);
  if (a.excl) qlist_resolve_pending(Q, a, q, k);
  const float sc = smax * Q.inv_scale[q];
  uint64_t key = make_key(sc, (uint32_t)drow);
  const uint64_t thr = Q.thr[q];
  if ((key >> 32) < (thr >> 32)) return;
  const int64_t orow = a.orig_row ? a.orig_row[drow] : drow;
  key = (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(a.row_base + orow));
  if (key <= thr) return;
  if (a.excl) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.excl + (size_t)q * a.excl_words + (img >> 5)));
    Q.pend_key[q] = key;
    Q.pend_img[q] = img;
  } else {
    qlist_insert(Q, a, q, k, key, img);
  }
}

// Running (max score, column) of the two queries a thread sees through the 16x256b loads; passed and
// returned BY VALUE so it stays in registers across the out-of-line call.
struct Run2 {
  float m0, m1;
  int c0, c1;
};

// One 8-column group that contains at least one image boundary (about 1 group in 5): walk the columns
// in order; column p belongs to the thread with lane%4 == p/2; at a boundary the four threads of a
// quad combine their running maxima and the owner flushes the image.  eb: boundary bits of the group,
// sXY: this thread's column Y of half X, col8: column (relative to r_begin) of p = 0, img_lane: image
// id of column (32g + lane) held by each lane, sub: 8-column group index within the 32 columns,
// own_q: query slot this thread owns for half own_h (own_h < 0: none).
// misc packs sub (bits 0-3), own_h + 1 (bits 4-7) and own_q (bits 8..) to keep the call in registers.
__device__ __noinline__ Run2 scan_tc_slow_group(Run2 run, uint8_t* qbase, const ScanTcArgs& a, uint32_t eb, float s00,
                                                float s01, float s10, float s11, int col8, int img_lane, int misc,
                                                int64_t r_begin) {
  const int lane = threadIdx.x & 31, j = lane & 3;
  const int sub = misc & 15, own_h = ((misc >> 4) & 15) - 1, own_q = misc >> 8;
#pragma unroll 1
  for (int p = 0; p < 8; ++p) {
    if ((p >> 1) == j) {
      const float x0 = (p & 1) ? s01 : s00, x1 = (p & 1) ? s11 : s10;
      if (x0 > run.m0) {
        run.m0 = x0;
        run.c0 = col8 + p;
      }
      if (x1 > run.m1) {
        run.m1 = x1;
        run.c1 = col8 + p;
      }
    }
    if ((eb >> p) & 1u) {      // warp-uniform
      const int img = __shfl_sync(0xffffffffu, img_lane, sub * 8 + p);
      // (score desc, column asc) max over the quad, both halves
      uint64_t k0 = ((uint64_t)f32_ordered(run.m0) << 32) | (uint32_t)(0x7FFFFFFF - run.c0);
      uint64_t k1 = ((uint64_t)f32_ordered(run.m1) << 32) | (uint32_t)(0x7FFFFFFF - run.c1);
      uint64_t o0 = shfl_xor_u64(k0, 1), o1 = shfl_xor_u64(k1, 1);
      k0 = o0 > k0 ? o0 : k0;
      k1 = o1 > k1 ? o1 : k1;
      o0 = shfl_xor_u64(k0, 2);
      o1 = shfl_xor_u64(k1, 2);
      k0 = o0 > k0 ? o0 : k0;
      k1 = o1 > k1 ? o1 : k1;
      if (own_h >= 0) {
        const uint64_t kk = own_h ? k1 : k0;
        const float best = f32_from_ordered((uint32_t)(kk >> 32));
        const int col = 0x7FFFFFFF - (int)(uint32_t)(kk & 0xFFFFFFFFu);
        if (best > -INFINITY) scan_tc_flush(qbase, a, own_q, best, r_begin + col, img);
      }
      run.m0 = run.m1 = -INFINITY;
      run.c0 = run.c1 = 0;
    }
  }
  return run;
}

template <int DIM, int NT, int NS>
__global__ void __launch_bounds__(kTcThreads, 1) scan_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                 const ScanTcArgs a) {
  using Cfg = TcCfg<DIM, NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem;
  const TcSmem S = tc_carve(smem_raw, NS, Cfg::STAGE_BYTES, &smem);
  uint8_t* after = smem + NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16;
  const QShared Q = qshared_carve(after, a.k);
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
          mbar_wait_parked(S.empty + 8 * p.stage, p.phase ^ 1);
          mbar_expect_tx(S.full + 8 * p.stage, Cfg::STAGE_BYTES);
          tma_load_2d(S.stages + p.stage * Cfg::STAGE_BYTES, &tmap, kc * kTcKChunk, (int)(r_begin + (int64_t)t * NT),
                      S.full + 8 * p.stage);
          p.advance();
        }
      }
    }
  } else if (warp == 1) {
    TcPipe p(NS);
    mbar_wait_parked(S.a_ready, 0);
    tc_fence_after();
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = t & 1;
      mbar_wait_parked(S.tmem_empty + 8 * as, ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < Cfg::KC; ++kc) {
        mbar_wait_parked(S.full + 8 * p.stage, p.phase);
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
    // ===== epilogue warps.  TMEM lane layout of quarter q4 (lanes 32*q4 ..):
    //   lane 16h + r      (r < 8): hi part of query 16*q4 + 8h + r
    //   lane 16h + 8 + r          : lo part of the same query
    // so a 16x256b load of half h hands thread (r = lane/4, j = lane%4) BOTH parts of query 16*q4+8h+r
    // for columns 8i+2j, 8i+2j+1: the hi+lo sum needs no shuffle and no lane is redundant.
    const int q4 = warp & 3;
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    const int k = a.k;

    // ---- A operand (thread t writes TMEM lane 32*q4 + t)
    {
      const int h = lane >> 4, r = lane & 7;
      const bool is_lo = (lane >> 3) & 1;
      const int qa = q4 * 16 + h * 8 + r;
      const bool qa_ok = qa < a.nq;
      const float* qp = a.q + (size_t)(qa_ok ? qa : 0) * DIM;
      float mx = 0.f;
      if (qa_ok)
        for (int i = 0; i < DIM; ++i) mx = fmaxf(mx, fabsf(__ldg(qp + i)));
      int e = 0;
      if (mx > 0.f && mx < INFINITY) {
        int ex;
        frexpf(mx, &ex);          // mx = m * 2^ex, m in [0.5, 1)
        e = 11 - ex;              // mx * 2^e in [1024, 2048): hi keeps 11 bits, lo the next 11
      }
      const float scale = ldexpf(1.0f, e);
      if (!is_lo) {
        Q.inv_scale[qa] = ldexpf(1.0f, -e);
        Q.thr[qa] = 0;
        Q.pend_key[qa] = 0;
        Q.pend_img[qa] = 0;
        Q.cnt[qa] = 0;
        Q.minpos[qa] = 0;
      }
#pragma unroll 1
      for (int c = 0; c < Cfg::A_COLS / 32; ++c) {
        uint32_t rr[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          float x0 = 0.f, x1 = 0.f;
          if (qa_ok) {
            x0 = __ldg(qp + c * 64 + 2 * jj) * scale;
            x1 = __ldg(qp + c * 64 + 2 * jj + 1) * scale;
          }
          __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
          if (is_lo) {
            h0 = __float2half_rn(x0 - __half2float(h0));
            h1 = __float2half_rn(x1 - __half2float(h1));
          }
          rr[jj] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        }
        tmem_st32(lane_addr + Cfg::A_BASE + c * 32, rr);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(S.a_ready);
      __syncwarp();            // the state of this warp's 16 query slots is visible below
    }

    const int j = lane & 3, r = lane >> 2;
    const int own_h = j < 2 ? j : -1;                  // thread j==h of a quad owns the quad's half-h query
    const int qi = q4 * 16 + (own_h < 0 ? 0 : own_h) * 8 + r;
    const bool owner = own_h >= 0 && qi < a.nq;
    const int my_own_h = owner ? own_h : -1;
    Run2 run{-INFINITY, -INFINITY, 0, 0};

    // Image ids of this lane's columns (lane, lane+32, ...) and of their right neighbours, fetched one
    // tile AHEAD (consumed right after the load they cost a loaded-HBM latency per 32 columns).
    constexpr int NG = NT / 32;
    int img_a[NG], img_b[NG];
    auto fetch_imgs = [&](int t) {
      const int64_t row0 = r_begin + (int64_t)t * NT;
      const int valid = (int)min((int64_t)NT, r_end - row0);
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int c = g * 32 + lane;
        img_a[g] = c < valid ? a.img_of_row[row0 + c] : 0;
        img_b[g] = c < valid ? a.img_of_row[row0 + c + 1] : 0;
      }
    };
    if (ntiles > 0) fetch_imgs(0);
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = t & 1;
      const int64_t row0 = r_begin + (int64_t)t * NT;
      // bit i of ends[g] set <=> row0 + 32g + i is the last row of its image (0 for columns past r_end)
      uint32_t ends[NG];
      int img_now[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        ends[g] = __ballot_sync(0xffffffffu, img_a[g] != img_b[g]);
        img_now[g] = img_a[g];
      }
      if (t + 1 < ntiles) fetch_imgs(t + 1);
      uint64_t gshared = 0;
      if (owner) gshared = ld_relaxed_u64(a.g_thr + qi);
      mbar_wait(S.tmem_full + 8 * as, (t >> 1) & 1);
      tc_fence_after();
      if (owner && gshared > Q.thr[qi]) Q.thr[qi] = gshared;
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        uint32_t em = ends[0];
        int img_lane = img_now[0];
#pragma unroll
        for (int gg = 1; gg < NG; ++gg) {
          em = (g == gg) ? ends[gg] : em;
          img_lane = (g == gg) ? img_now[gg] : img_lane;
        }
        uint32_t v0[16], v1[16];
        tmem_ld_16x256b_x4(lane_addr + Cfg::ACC_BASE + as * NT + g * 32, v0);
        tmem_ld_16x256b_x4(lane_addr + (16u << 16) + Cfg::ACC_BASE + as * NT + g * 32, v1);
        tmem_ld_wait();
        const int colbase = (int)(row0 - r_begin) + g * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float s00 = __uint_as_float(v0[4 * i]) + __uint_as_float(v0[4 * i + 2]);
          const float s01 = __uint_as_float(v0[4 * i + 1]) + __uint_as_float(v0[4 * i + 3]);
          const float s10 = __uint_as_float(v1[4 * i]) + __uint_as_float(v1[4 * i + 2]);
          const float s11 = __uint_as_float(v1[4 * i + 1]) + __uint_as_float(v1[4 * i + 3]);
          const uint32_t eb = (em >> (8 * i)) & 0xFFu;
          const int c0 = colbase + 8 * i + 2 * j;
          if (eb == 0) {          // warp-uniform fast path: no image ends inside these 8 columns
            if (s00 > run.m0) { run.m0 = s00; run.c0 = c0; }
            if (s01 > run.m0) { run.m0 = s01; run.c0 = c0 + 1; }
            if (s10 > run.m1) { run.m1 = s10; run.c1 = c0; }
            if (s11 > run.m1) { run.m1 = s11; run.c1 = c0 + 1; }
          } else {
            run = scan_tc_slow_group(run, after, a, eb, s00, s01, s10, s11, colbase + 8 * i, img_lane,
                                     i | ((my_own_h + 1) << 4) | (qi << 8), r_begin);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(S.tmem_empty + 8 * as);
    }
    // ---- settle the parked candidate and publish this CTA's list of every query
    if (owner) {
      if (a.excl) qlist_resolve_pending(Q, a, qi, k);
      const int cnt = Q.cnt[qi];
      const int64_t o = ((int64_t)qi * gridDim.x + blockIdx.x) * k;
      for (int s = 0; s < k; ++s) {
        a.list_keys[o + s] = s < cnt ? Q.keys[s * 64 + qi] : 0ull;
        a.list_dbidx[o + s] = s < cnt ? a.img_dbidx[Q.img[s * 64 + qi]] : -1;
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
  const size_t smem = (size_t)NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16 + qshared_bytes(a.k) +
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
