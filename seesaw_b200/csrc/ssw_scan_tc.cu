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
// and accumulator column = database row: the per-image segmented max is a running max along the
// columns and image boundaries are warp-uniform.
// fp32 queries are split into hi + lo fp16 parts after scaling by a power of two (keeps ~22 mantissa
// bits; exact for fp16-representable queries) by a small preparation kernel that writes the TMEM image
// of the A operand.  TMEM lane layout of a 32-lane quarter q4 (h = 0,1; r = 0..7):
//   lane 16h + r     : hi part of query 16*q4 + 8h + r
//   lane 16h + 8 + r : lo part of the same query
// so a 16x256b tcgen05.ld of half h hands thread (r = lane/4, j = lane%4) BOTH parts of one query for
// columns 8i+2j, 8i+2j+1 (i = 0..3): hi + lo needs no shuffle and no lane is redundant.
//
// Epilogue cost model: eight epilogue warps, two per 32-lane quarter of tensor memory
// (warp q4 takes the quarter's half 0 = queries 16*q4 + r, warp q4 + 4 its half 1 = queries 16*q4 + 8 + r), so
// a thread carries ONE query and every scheduler has two epilogue warps to alternate (the epilogue is bound by
// dependent-issue latency).  Per 128-row tile a thread folds 32 scores (FADD hi + lo, then a depth-3 argmax
// tree per 8 columns); an image boundary costs one vote, and only when some lane's partial max reaches its
// query's threshold a quad reduction (4 SHFL) plus the owner's threshold / exclusion / list work.  Everything
// lives in registers and shared memory: no local memory, no out-of-line calls.  One more warp pools the CTAs'
// published bests into tighter thresholds for the whole life of the kernel.
#include <algorithm>
#include <cstdlib>

#include "ssw_db.h"
#include "ssw_tc.cuh"

namespace ssw {

struct ScanTcArgs {
  const uint32_t* a_img;     // [DIM/8][128][4] packed fp16 pairs: image of the A operand in 16-byte pieces, piece p of
                             // TMEM lane L at ((p * 128 + L) * 4): a warp copying one piece of 32 lanes reads 512 contiguous bytes
  const float* inv_scale;    // [64] 2^-e of every query slot (scores = accumulator * inv_scale)
  int nq;
  int k;
  const uint32_t* excl;      // [nq, excl_words] or null
  int64_t excl_words;
  int excl_slice_words;      // > 0: every CTA keeps its image range's slice of the bitmaps in shared memory
  const uint32_t* last_bits; // bit r set <=> device row r is the last row of its image
  const int64_t* row_ptr;
  const int32_t* img_dbidx;
  const int64_t* orig_row;   // null when identity
  const int32_t* part;       // [grid*8 + 1] image partition of the streaming scan, 8 entries per CTA
  uint64_t* cand_keys;       // [nq][grid * k] compacted candidates of every query: each CTA appends the entries of
  int32_t* cand_dbidx;       //   its list that still reach the shared bound when it finishes
  int32_t* cand_cnt;         // [nq] entries appended so far
  uint64_t* g_thr;           // [nq]
  uint32_t* pub;             // [64][grid] best score (order-preserving bits) every CTA holds per query
  int64_t row_base;
  unsigned long long* stats; // [2] list updates / images offered to a list, summed over the grid (null = not counted)
};

// Development-only timeline (make EXTRA=-DSSW_TRACE): clock64 stamps of one CTA's phases into a.stats[16 + ...];
// EXTRA=-DSSW_TRACE=2 stamps %globaltimer instead (ns, comparable across SMs and kernels: scripts/trace_pipe.py)
#ifdef SSW_TRACE
__device__ __forceinline__ long long ssw_trace_now() {
#if SSW_TRACE == 2
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
#else
  return clock64();
#endif
}
#define SSW_TR(slot, cond)                                                                                        \
  do {                                                                                                            \
    if (a.stats && (cond)) reinterpret_cast<long long*>(a.stats)[16 + blockIdx.x * 16 + (slot)] = ssw_trace_now(); \
  } while (0)
// ... and cycles spent inside a wait, summed into `acc` and stored by SSW_TRV: which of memory / MMA / epilogue a CTA waits for
#define SSW_TW(acc, stmt)          \
  do {                             \
    const long long _t = clock64(); \
    stmt;                          \
    acc += clock64() - _t;         \
  } while (0)
#define SSW_TRV(slot, cond, val)                                                                          \
  do {                                                                                                    \
    if (a.stats && (cond)) reinterpret_cast<long long*>(a.stats)[16 + blockIdx.x * 16 + (slot)] = (val); \
  } while (0)
#else
#define SSW_TR(slot, cond) do { } while (0)
#define SSW_TW(acc, stmt) do { stmt; } while (0)
#define SSW_TRV(slot, cond, val) do { } while (0)
#endif

constexpr int kTcMaxK = 64;   // per-query list length the batched epilogue keeps in shared memory
constexpr int kTcSlack = 8;   // extra list slots: entries are appended and the list is cut back to k when it overflows

// ------------------------------------------------------------------------------------------
// query preparation: fp32 [nq, DIM] -> hi/lo fp16 image of the A operand, scales, zeroed thresholds
// one block per TMEM lane (128), DIM/2 threads, thread t converts elements 2t, 2t+1
// ------------------------------------------------------------------------------------------
__global__ void scan_tc_prep_kernel(const float* __restrict__ q, int nq, int dim, uint32_t* __restrict__ a_img,
                                    float* __restrict__ inv_scale, uint64_t* __restrict__ g_thr,
                                    int32_t* __restrict__ cand_cnt, uint32_t* __restrict__ pub, int n_pub) {
  pdl_launch_dependents();      // the scan kernel may set itself up (barriers, TMEM, tensor map) while this runs
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pub; i += gridDim.x * blockDim.x) pub[i] = 0u;
  const int L = blockIdx.x;                       // TMEM lane
  const int q4 = L >> 5, h = (L >> 4) & 1, r = L & 7;
  const bool is_lo = (L >> 3) & 1;
  const int qa = q4 * 16 + h * 8 + r;
  const bool ok = qa < nq;
  const int t = threadIdx.x;
  float x0 = 0.f, x1 = 0.f;
  if (ok) {
    const float2 v = reinterpret_cast<const float2*>(q + (size_t)qa * dim)[t];
    x0 = v.x;
    x1 = v.y;
  }
  __shared__ float s_max[32];
  float mx = fmaxf(fabsf(x0), fabsf(x1));
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  if ((t & 31) == 0) s_max[t >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) mx = fmaxf(mx, s_max[w]);
  int e = 0;
  if (mx > 0.f && mx < INFINITY) {
    int ex;
    frexpf(mx, &ex);          // mx = m * 2^ex, m in [0.5, 1)
    e = 11 - ex;              // mx * 2^e in [1024, 2048): hi keeps 11 bits, lo the next 11
  }
  const float scale = ldexpf(1.0f, e);
  x0 *= scale;
  x1 *= scale;
  __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  if (is_lo) {
    h0 = __float2half_rn(x0 - __half2float(h0));
    h1 = __float2half_rn(x1 - __half2float(h1));
  }
  a_img[((size_t)(t >> 2) * 128 + L) * 4 + (t & 3)] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
  if (t == 0 && !is_lo) {
    inv_scale[qa] = ldexpf(1.0f, -e);
    if (ok) {
      g_thr[qa] = 0ull;
      cand_cnt[qa] = 0;
    }
  }
}

// Per-query epilogue state in SHARED memory (structure of arrays over the 64 query slots).
struct QShared {
  uint64_t* keys;       // [k + kTcSlack][64]   candidate keys, slot s of query q at keys[s*64 + q]
  int32_t* img;         // [k + kTcSlack][64]   local image index of the candidate
  uint64_t* thr;        // [64] reject keys <= thr: max(own k-th best after a compaction, shared lower bound)
  int32_t* cnt;         // [64] entries held (<= k + kTcSlack)
  int32_t* minpos;      // [64] (unused)
  uint32_t* best;       // [64] order-preserving score bits of the best candidate held (published to the other CTAs)
  int* done;            // epilogue warps that have finished (the threshold warp leaves at 8)
  uint32_t* upd;        // [64] list updates of this CTA (appends + replacements), [64] images offered (past the vote)
  uint32_t* excl;       // [64][slice_words] this CTA's slice of the exclusion bitmaps (optional)
};
__host__ __device__ constexpr size_t qshared_bytes(int k) { return (size_t)(k + kTcSlack) * 64 * 12 + 64 * (8 + 4 + 4 + 4 + 8) + 16; }

__device__ __forceinline__ QShared qshared_carve(uint8_t* base, int k) {
  QShared q;
  q.keys = reinterpret_cast<uint64_t*>(base);
  q.thr = q.keys + (size_t)(k + kTcSlack) * 64;
  q.img = reinterpret_cast<int32_t*>(q.thr + 64);
  q.cnt = q.img + (size_t)(k + kTcSlack) * 64;
  q.minpos = q.cnt + 64;
  q.best = reinterpret_cast<uint32_t*>(q.minpos + 64);
  q.done = reinterpret_cast<int*>(q.best + 64);
  q.upd = reinterpret_cast<uint32_t*>(q.done + 4);
  q.excl = q.upd + 128;
  return q;
}

// threshold key -> the accumulator-space score a partial maximum must reach to matter
__device__ __forceinline__ float thr_to_acc(uint64_t thr, float scale) {
  return thr == 0 ? -INFINITY : key_score(thr) * scale;
}

// Owner thread of query q: the finished image `img` peaked at accumulator value `best` in device row
// `drow`.  Applies the threshold and the exclusion bitmap and APPENDS the candidate to the query's list; sets
// `full` when the list has reached its k + kTcSlack slots (the warp then cuts it back to k: scan_tc_compact).
// A list is never searched for its minimum on insertion: on data whose scores rise along a CTA's range every
// image replaces the weakest entry, and a k-entry rescan per image by one thread made that CTA the straggler of
// the whole launch (10M rows sorted by one query's score: 3.3 ms instead of 1.55 ms).
__device__ __forceinline__ void scan_tc_offer(const QShared& Q, const ScanTcArgs& a, int q, float best,
                                              float inv_scale, int64_t drow, int img, int slice_base, bool& full) {
  // every shared-memory word the path needs is loaded up front: one LDS latency instead of a chain of dependent ones
  const uint64_t thr = Q.thr[q];
  const int cnt = Q.cnt[q];
  const uint32_t best_seen = Q.best[q];
  uint32_t xw = 0;
  if (a.excl && a.excl_slice_words > 0) xw = Q.excl[q * a.excl_slice_words + (img >> 5) - slice_base];
  uint64_t key = make_key(best * inv_scale, (uint32_t)drow);
  ++Q.upd[64 + q];
  if ((key >> 32) < (thr >> 32)) return;
  const int64_t orow = a.orig_row ? a.orig_row[drow] : drow;
  key = (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(a.row_base + orow));
  if (key <= thr) return;
  if (a.excl) {
    if (a.excl_slice_words == 0) xw = __ldg(a.excl + (size_t)q * a.excl_words + (img >> 5));
    if ((xw >> (img & 31)) & 1u) return;
  }
  if ((uint32_t)(key >> 32) > best_seen) {      // a new best of this CTA: let the other CTAs see it
    Q.best[q] = (uint32_t)(key >> 32);
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.pub + (size_t)q * gridDim.x + blockIdx.x),
                 "r"((uint32_t)(key >> 32)) : "memory");
  }
  ++Q.upd[q];
  Q.keys[cnt * 64 + q] = key;
  Q.img[cnt * 64 + q] = img;
  Q.cnt[q] = cnt + 1;
  full = cnt + 1 >= a.k + kTcSlack;
}

// All 32 lanes: cut query q's full list (k + kTcSlack entries) back to its k best.  Every lane ranks up to three
// entries by counting the larger ones (keys are unique, the rank is the slot: the list comes out sorted best
// first); the k-th key becomes the query's own threshold and is folded into the shared bound.  Returns that key.
__device__ __forceinline__ uint64_t scan_tc_compact(const QShared& Q, const ScanTcArgs& a, int q, int lane) {
  const int k = a.k, C = a.k + kTcSlack;        // C <= 72: three slots per lane
  uint64_t mine[3];
  int mimg[3], rank[3] = {0, 0, 0};
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int e = lane + 32 * u;
    mine[u] = e < C ? Q.keys[e * 64 + q] : 0ull;
    mimg[u] = e < C ? Q.img[e * 64 + q] : 0;
  }
#pragma unroll 4
  for (int j = 0; j < C; ++j) {
    const uint64_t x = Q.keys[j * 64 + q];
#pragma unroll
    for (int u = 0; u < 3; ++u) rank[u] += x > mine[u];
  }
  __syncwarp();
  uint64_t kth = 0;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    if (lane + 32 * u < C && rank[u] < k) {
      Q.keys[rank[u] * 64 + q] = mine[u];
      Q.img[rank[u] * 64 + q] = mimg[u];
      if (rank[u] == k - 1) kth = mine[u];
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) kth |= shfl_xor_u64(kth, m);      // exactly one lane holds it
  if (lane == 0) {
    Q.cnt[q] = k;
    atomicMax(reinterpret_cast<unsigned long long*>(Q.thr + q), (unsigned long long)kth);    // the threshold warp writes too
    atomicMax(reinterpret_cast<unsigned long long*>(a.g_thr + q), (unsigned long long)kth);
  }
  __syncwarp();
  return kth;
}

// Threshold warp.  Every CTA publishes the best score it holds per query (a.pub).  Split the G CTAs into
// NGP >= k groups (CTA c in group c mod NGP): the smallest of the group maxima is a score that at least
// k distinct images reach, hence a valid lower bound of the final k-th best — far tighter than a single
// CTA's own k-th best early in the scan, when almost every image would otherwise enter its CTA's list
// (expected list updates per query and CTA drop from k*ln(n/k) to a handful).
// What a list update costs is serial work of one epilogue thread, and how many there are depends on how STALE
// the bound is (a CTA timeline showed the first half of a 66-tile range running 25 % slower than the second:
// every CTA used to re-pool all 64 queries from the 38 KB of published bests, ~12 us per sweep under a
// saturated memory system).  So the pooling is split: CTA b pools only the 8 queries of group (b mod 8) — one
// batch of loads, ~1.5 us — and folds the result into the shared bound g_thr; EVERY CTA's threshold warp then
// just reads the 64 shared bounds (two coalesced 8-byte loads per lane) each round and raises the thresholds
// the epilogue warps read.  The warp runs for the whole life of the kernel, off the epilogue's critical path.
__device__ __forceinline__ void scan_tc_threshold_warp(const QShared& Q, const ScanTcArgs& a, int lane, int n_epi) {
  constexpr int VMAX = 6;                       // up to 192 CTAs
  constexpr int QB = 8;                         // queries pooled per CTA (loads in flight together)
  const int G = gridDim.x;
  int ngp = 1;
  while (ngp < a.k) ngp <<= 1;                  // k <= 64
  const bool pooled = G >= ngp && G <= 32 * VMAX;
  const int n_groups = (a.nq + QB - 1) / QB;
  const int q0 = ((int)blockIdx.x % n_groups) * QB;
  volatile int* done = Q.done;
  uint64_t seen0 = 0, seen1 = 0;                // what this lane last folded into Q.thr[lane], Q.thr[lane + 32]
  while (*done < n_epi) {
    if (pooled) {
      uint32_t v[QB][VMAX];
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int q = q0 + u;
#pragma unroll
        for (int m = 0; m < VMAX; ++m) {
          const int c = lane + 32 * m;
          v[u][m] = 0u;
          if (q < a.nq && c < G)
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[u][m]) : "l"(a.pub + (size_t)q * G + c) : "memory");
        }
      }
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int q = q0 + u;
        if (q >= a.nq) break;
        const uint32_t t = pooled_group_min<VMAX>(v[u], ngp, lane);
        if (lane == 0 && t != 0) atomicMax(reinterpret_cast<unsigned long long*>(a.g_thr + q), (unsigned long long)t << 32);
      }
    }
    const uint64_t g0 = lane < a.nq ? ld_relaxed_u64(a.g_thr + lane) : 0ull;
    const uint64_t g1 = lane + 32 < a.nq ? ld_relaxed_u64(a.g_thr + lane + 32) : 0ull;
    if (g0 > seen0) {
      atomicMax(reinterpret_cast<unsigned long long*>(Q.thr + lane), (unsigned long long)g0);
      seen0 = g0;
    }
    if (g1 > seen1) {
      atomicMax(reinterpret_cast<unsigned long long*>(Q.thr + lane + 32), (unsigned long long)g1);
      seen1 = g1;
    }
    __syncwarp();
    if (*done < n_epi) __nanosleep(200);
  }
}

// (max, LOWEST index attaining it) of 8 values as a depth-3 tree: the epilogue warps run alone on their
// schedulers, so dependent-issue latency — not instruction count — is what a tile costs (ncu: stall_wait
// is the top stall reason); a sequential running max would be a chain of 8 compare/select pairs.
__device__ __forceinline__ void argmax8(const float* x, float& m, int& idx) {
  const bool p0 = x[1] > x[0], p1 = x[3] > x[2], p2 = x[5] > x[4], p3 = x[7] > x[6];
  const float m0 = p0 ? x[1] : x[0], m1 = p1 ? x[3] : x[2], m2 = p2 ? x[5] : x[4], m3 = p3 ? x[7] : x[6];
  const int i0 = p0 ? 1 : 0, i1 = p1 ? 3 : 2, i2 = p2 ? 5 : 4, i3 = p3 ? 7 : 6;
  const bool q0 = m1 > m0, q1 = m3 > m2;
  const float n0 = q0 ? m1 : m0, n1 = q1 ? m3 : m2;
  const int j0 = q0 ? i1 : i0, j1 = q1 ? i3 : i2;
  const bool r = n1 > n0;
  m = r ? n1 : n0;
  idx = r ? j1 : j0;
}

struct Epi1State {
  float m;            // running max of this thread's columns for its quad's query
  int c;
  float thr;
  int cur_img;
};
struct Epi1Ctx {
  int j;
  bool owner;         // j == 0 owns the quad's query
  int own_q;
  float inv, scale;
  int64_t r_begin;
  int slice_base;
};

__device__ __forceinline__ void scan_tc8_boundary(Epi1State& st, const Epi1Ctx& cx, const QShared& Q, const ScanTcArgs& a) {
  if (__any_sync(0xffffffffu, st.m >= st.thr)) {
    uint64_t k0 = ((uint64_t)f32_ordered(st.m) << 32) | (uint32_t)(0x7FFFFFFF - st.c);
    uint64_t o0 = shfl_xor_u64(k0, 1);
    k0 = o0 > k0 ? o0 : k0;
    o0 = shfl_xor_u64(k0, 2);
    k0 = o0 > k0 ? o0 : k0;
    bool full = false;
    if (cx.owner) {
      const float best = f32_from_ordered((uint32_t)(k0 >> 32));
      if (best >= st.thr) {
        const int col = 0x7FFFFFFF - (int)(uint32_t)(k0 & 0xFFFFFFFFu);
        scan_tc_offer(Q, a, cx.own_q, best, cx.inv, cx.r_begin + col, st.cur_img, cx.slice_base, full);
      }
    }
    uint32_t need = __ballot_sync(0xffffffffu, full);
    while (need) {        // rare: a list ran over its slack — the whole warp cuts it back to k
      const int src = __ffs(need) - 1;
      need &= need - 1;
      const int q = __shfl_sync(0xffffffffu, cx.own_q, src);
      const uint64_t kth = scan_tc_compact(Q, a, q, threadIdx.x & 31);
      if ((int)(threadIdx.x & 31) == src) {
        const float t = thr_to_acc(kth, cx.scale);
        st.thr = t > st.thr ? t : st.thr;
      }
    }
    __syncwarp();
  }
  st.m = -INFINITY;
  st.c = 0;
  ++st.cur_img;
}

__device__ __forceinline__ void scan_tc8_group(Epi1State& st, const Epi1Ctx& cx, const QShared& Q, const ScanTcArgs& a,
                                               const uint32_t* v, uint32_t em, int colbase) {
  float sa[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sa[2 * i] = __uint_as_float(v[4 * i]) + __uint_as_float(v[4 * i + 2]);
    sa[2 * i + 1] = __uint_as_float(v[4 * i + 1]) + __uint_as_float(v[4 * i + 3]);
  }
  // Fast path (almost every group once the thresholds are up): if none of the warp's 256 scores reaches its
  // query's threshold, no image can become a candidate THROUGH these columns — a running maximum below the
  // threshold is never looked at — so nothing is folded and only the image boundaries are walked (one vote each).
  // An image whose maximum does reach the threshold has its arg-max column in a group that takes the full path
  // below, with the exact (max, lowest column) bookkeeping.
  {
    const float mx = fmaxf(fmaxf(fmaxf(sa[0], sa[1]), fmaxf(sa[2], sa[3])), fmaxf(fmaxf(sa[4], sa[5]), fmaxf(sa[6], sa[7])));
    if (!__any_sync(0xffffffffu, mx >= st.thr)) {
      while (em) {
        em &= em - 1;
        scan_tc8_boundary(st, cx, Q, a);
      }
      return;
    }
  }
  const int cb = colbase + 2 * cx.j;
  float ma;
  int ia;
  if (em == 0) {
    argmax8(sa, ma, ia);
    if (ma > st.m) { st.m = ma; st.c = cb + ((ia >> 1) << 3) + (ia & 1); }
    return;
  }
  int lo = 0;
  for (;;) {
    const int p = em ? __ffs(em) - 1 : 31;
    const uint32_t seg = (0xFFFFFFFFu >> (31 - p)) & (0xFFFFFFFFu << lo);
    const uint32_t mine = seg >> (2 * cx.j);
    float xa[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) xa[2 * i + e] = ((mine >> (8 * i + e)) & 1u) ? sa[2 * i + e] : -INFINITY;
    argmax8(xa, ma, ia);
    if (ma > st.m) { st.m = ma; st.c = cb + ((ia >> 1) << 3) + (ia & 1); }
    if (em == 0) break;
    em &= em - 1;
    scan_tc8_boundary(st, cx, Q, a);
    lo = p + 1;
    if (lo == 32) break;
  }
}

__device__ __forceinline__ void tmem_ld_wait_regs16(uint32_t* x) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]),
                 "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
               :
               : "memory");
}

// ... and for two register blocks at once
__device__ __forceinline__ void tmem_ld_wait_regs32(uint32_t* x, uint32_t* y) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]),
                 "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15]),
                 "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]),
                 "+r"(y[8]), "+r"(y[9]), "+r"(y[10]), "+r"(y[11]), "+r"(y[12]), "+r"(y[13]), "+r"(y[14]), "+r"(y[15])
               :
               : "memory");
}

constexpr int kScanTc8Threads = 64 + 8 * 32 + 32;   // producer, MMA, 8 epilogue warps, threshold warp

template <int DIM, int NT, int NS, int NACC>
__global__ void __launch_bounds__(kScanTc8Threads, 1) scan_tc8_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const ScanTcArgs a) {
  using Cfg = TcCfg<DIM, NT, NACC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem;
  const TcSmem S = tc_carve(smem_raw, NS, Cfg::STAGE_BYTES, &smem);
  uint8_t* after = smem + NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16;
  const QShared Q = qshared_carve(after, a.k);
  SSW_TR(0, threadIdx.x == 0);
  if (threadIdx.x < 64) {
    Q.thr[threadIdx.x] = 0;
    Q.cnt[threadIdx.x] = 0;
    Q.minpos[threadIdx.x] = 0;
    Q.best[threadIdx.x] = 0;
    Q.upd[threadIdx.x] = 0;
    Q.upd[64 + threadIdx.x] = 0;
    if (threadIdx.x == 0) *Q.done = 0;
  }
  const uint32_t tmem = tc_setup(S, NS, Cfg::TMEM_ALLOC, &tmap, 256, 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // everything above overlapped the tail of the query-preparation kernel (programmatic dependent launch); from
  // here on its outputs (A image, scales, zeroed thresholds / published bests) are read
  SSW_TR(1, threadIdx.x == 0);
  pdl_launch_dependents();
  pdl_wait();
  SSW_TR(2, threadIdx.x == 0);

  const int img0 = a.part[blockIdx.x * kScanWarps], img1 = a.part[(blockIdx.x + 1) * kScanWarps];
  const int64_t r_begin = a.row_ptr[img0], r_end = a.row_ptr[img1];
  const int64_t nrows = r_end - r_begin;
  const int ntiles = (int)((nrows + NT - 1) / NT);

  if (warp == 0) {
    if (lane == 0) {
      TcPipe p(NS);
      SSW_TR(3, true);
      [[maybe_unused]] long long w_ring = 0;        // trace build: cycles the producer waited for a free stage
      for (int t = 0; t < ntiles; ++t) {
        for (int kc = 0; kc < Cfg::KC; ++kc) {
          SSW_TW(w_ring, mbar_wait_parked(S.empty + 8 * p.stage, p.phase ^ 1));
          mbar_expect_tx(S.full + 8 * p.stage, Cfg::STAGE_BYTES);
          tma_load_2d(S.stages + p.stage * Cfg::STAGE_BYTES, &tmap, kc * kTcKChunk, (int)(r_begin + (int64_t)t * NT),
                      S.full + 8 * p.stage);
          p.advance();
        }
      }
      SSW_TRV(4, true, w_ring);
      SSW_TR(5, true);
    }
    __syncwarp();       // the idle lanes must not reach the closing __syncthreads ahead of lane 0
  } else if (warp == 1) {
    TcPipe p(NS);
    mbar_wait_parked(S.a_ready, 0);
    tc_fence_after();
    SSW_TR(6, lane == 0);
    [[maybe_unused]] long long w_acc = 0, w_full = 0;   // trace build: cycles waited for a free accumulator / for a loaded stage
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = NACC == 2 ? (t & 1) : 0;
      SSW_TW(w_acc, mbar_wait_parked(S.tmem_empty + 8 * as, ((NACC == 2 ? (t >> 1) : t) & 1) ^ 1));
      tc_fence_after();
      for (int kc = 0; kc < Cfg::KC; ++kc) {
        SSW_TW(w_full, mbar_wait_parked(S.full + 8 * p.stage, p.phase));
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
    SSW_TRV(7, lane == 0, w_acc);
    SSW_TRV(14, lane == 0, w_full);
    SSW_TR(8, lane == 0);
  } else if (warp == 10) {
    scan_tc_threshold_warp(Q, a, lane, 8);
  } else {
    // ===== epilogue warps 2..9: quarter = warp % 4, half = (warp - 2) / 4
    const int q4 = warp & 3, h = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    const int k = a.k;
    // ---- this warp's 8 exclusion bitmaps, CTA slice: loads issued now, stored after the A operand copy (their
    //      latency hides under it instead of adding 8 dependent round trips before the first tile)
    const int slice_base = img0 >> 5;
    const int nw = (a.excl && a.excl_slice_words > 0 && img1 > img0) ? ((img1 - 1) >> 5) - slice_base + 1 : 0;
    uint32_t xs[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int qs = q4 * 16 + h * 8 + i, w = lane + 32 * u;
        xs[i][u] = (qs < a.nq && w < nw) ? __ldg(a.excl + (size_t)qs * a.excl_words + slice_base + w) : 0u;
      }
    {      // A operand: the two warps of a quarter copy half of the columns of its 32 TMEM lanes each; all of a
           // warp's 16-byte pieces are requested at once (coalesced: a piece of 32 lanes is 512 contiguous bytes)
      constexpr int NCH = Cfg::A_COLS / 32;           // 32-column chunks of the operand
      constexpr int PER = NCH / 2;                    // chunks per warp
      constexpr int RND = PER > 4 ? 3 : PER;          // chunks per round (registers: 32 per chunk)
      static_assert(NCH % 2 == 0 && PER % RND == 0, "chunks must split evenly over two warps and the rounds");
      const uint4* src = reinterpret_cast<const uint4*>(a.a_img) + q4 * 32 + lane;
#pragma unroll 1
      for (int c = h * PER; c < (h + 1) * PER; c += RND) {
        uint32_t rr[RND][32];
#pragma unroll
        for (int u = 0; u < RND; ++u)
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const uint4 w = __ldg(src + (size_t)((c + u) * 8 + x) * 128);
            rr[u][4 * x] = w.x;
            rr[u][4 * x + 1] = w.y;
            rr[u][4 * x + 2] = w.z;
            rr[u][4 * x + 3] = w.w;
          }
#pragma unroll
        for (int u = 0; u < RND; ++u) tmem_st32(lane_addr + Cfg::A_BASE + (c + u) * 32, rr[u]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(S.a_ready);
    }
    if (nw > 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int qs = q4 * 16 + h * 8 + i;
        if (qs < a.nq) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (lane + 32 * u < nw) Q.excl[qs * a.excl_slice_words + lane + 32 * u] = xs[i][u];
          for (int w = lane + 64; w < nw; w += 32)      // slices beyond 2048 images per CTA: plain loop
            Q.excl[qs * a.excl_slice_words + w] = __ldg(a.excl + (size_t)qs * a.excl_words + slice_base + w);
        }
      }
      __syncwarp();
    }
    Epi1Ctx cx;
    cx.slice_base = slice_base;
    cx.j = lane & 3;
    const int r = lane >> 2;
    const int qA = q4 * 16 + h * 8 + r;
    cx.own_q = qA;
    cx.owner = cx.j == 0 && qA < a.nq;
    cx.inv = __ldg(a.inv_scale + qA);
    cx.scale = 1.0f / cx.inv;
    cx.r_begin = r_begin;
    Epi1State st;
    st.m = -INFINITY;
    st.c = 0;
    st.thr = -INFINITY;
    st.cur_img = img0;

    constexpr int NG = NT / 32;
    static_assert(NG == 2 || NG == 4, "tile must be 64 or 128 rows");
    uint32_t wb[NG + 1];
    auto fetch_bits = [&](int t) {
      const uint32_t* p = a.last_bits + ((r_begin + (int64_t)t * NT) >> 5);
#pragma unroll
      for (int i = 0; i <= NG; ++i) wb[i] = __ldg(p + i);
    };
    if (ntiles > 0) fetch_bits(0);
    const uint32_t half_addr = (uint32_t)(h * 16) << 16;
    [[maybe_unused]] long long w_tile = 0;          // trace build: cycles this epilogue warp waited for a finished accumulator
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = NACC == 2 ? (t & 1) : 0;
      const int64_t row0 = r_begin + (int64_t)t * NT;
      const int sh = (int)(row0 & 31);
      const int valid = (int)min((int64_t)NT, r_end - row0);
      uint32_t ends[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int rem = valid - 32 * g;
        const uint32_t keep = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
        ends[g] = __funnelshift_r(wb[g], wb[g + 1], sh) & keep;
      }
      if (t + 1 < ntiles) fetch_bits(t + 1);
      st.thr = thr_to_acc(Q.thr[qA], cx.scale);
      SSW_TW(w_tile, mbar_wait(S.tmem_full + 8 * as, (NACC == 2 ? (t >> 1) : t) & 1));
      tc_fence_after();
      SSW_TR(9, warp == 2 && lane == 0 && t == 0);
      SSW_TR(12, warp == 2 && lane == 0 && t == ntiles / 2);
      const uint32_t acc = lane_addr + half_addr + Cfg::ACC_BASE + as * NT;
      const int colbase = (int)(row0 - r_begin);
      if constexpr (NACC == 1) {
        // One accumulator (dim 768: the A operand takes 384 of the 512 TMEM columns).  Drain the whole tile into
        // registers first and hand the accumulator back at once, so the next tile's MMAs run under this tile's
        // epilogue arithmetic instead of after it.
        static_assert(NG == 4, "single-accumulator variant is written for 128-row tiles");
        uint32_t v[NG][16];
#pragma unroll
        for (int g = 0; g < NG; ++g) tmem_ld_16x256b_x4(acc + 32 * g, v[g]);
        tmem_ld_wait_regs32(v[0], v[1]);
        tmem_ld_wait_regs32(v[2], v[3]);
        tc_fence_before();
        mbar_arrive(S.tmem_empty + 8 * as);
#pragma unroll
        for (int g = 0; g < NG; ++g) scan_tc8_group(st, cx, Q, a, v[g], ends[g], colbase + 32 * g);
      } else {
        uint32_t va[16], vb[16];
        tmem_ld_16x256b_x4(acc, va);
        tmem_ld_wait_regs16(va);
#pragma unroll 1
        for (int gp = 0; gp < NG; gp += 2) {
          uint32_t eA = ends[0], eB = ends[1];
          if constexpr (NG == 4) {
            eA = gp ? ends[2] : eA;
            eB = gp ? ends[3] : eB;
          }
          tmem_ld_16x256b_x4(acc + 32 * (gp + 1), vb);
          scan_tc8_group(st, cx, Q, a, va, eA, colbase + 32 * gp);
          tmem_ld_wait_regs16(vb);
          if (gp + 2 < NG) tmem_ld_16x256b_x4(acc + 32 * (gp + 2), va);
          scan_tc8_group(st, cx, Q, a, vb, eB, colbase + 32 * (gp + 1));
          if (gp + 2 < NG) tmem_ld_wait_regs16(va);
        }
        tc_fence_before();
        mbar_arrive(S.tmem_empty + 8 * as);
      }
    }
    __syncwarp();
    SSW_TRV(15, warp == 2 && lane == 0, w_tile);
    SSW_TR(10, warp == 2 && lane == 0);
    if (lane == 0) atomicAdd(Q.done, 1);
    if (a.stats && lane < 8 && q4 * 16 + h * 8 + lane < a.nq) {
      atomicAdd(a.stats, (unsigned long long)Q.upd[q4 * 16 + h * 8 + lane]);
      atomicAdd(a.stats + 1, (unsigned long long)Q.upd[64 + q4 * 16 + h * 8 + lane]);
    }
    // ---- publish: of every query's list only the entries that still reach the shared bound, appended to the
    //      query's compacted candidate array — the merge reads tens to hundreds of keys per query instead of
    //      grid x k slots.  The warp's 8 queries reserve their ranges with 8 atomics in flight together (one
    //      round trip, not eight).
    {
      const int nqw = min(8, a.nq - (q4 * 16 + h * 8));
      const int64_t cap = (int64_t)gridDim.x * (k + kTcSlack);
      uint64_t key[8][3];
      uint32_t mask[8][3];
      int n_pass = 0;                                   // lane qq: entries of query qq that pass
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const int qi = q4 * 16 + h * 8 + qq;
        const int cntq = qq < nqw ? Q.cnt[qi] : 0;
        const uint64_t thr = qq < nqw ? ld_relaxed_u64(a.g_thr + qi) : ~0ull;
        int tot = 0;
#pragma unroll
        for (int u = 0; u < 3; ++u) {                   // k + kTcSlack <= 72: three slots per lane
          const int sl = lane + 32 * u;
          key[qq][u] = sl < cntq ? Q.keys[sl * 64 + qi] : 0ull;
          mask[qq][u] = __ballot_sync(0xffffffffu, key[qq][u] != 0ull && key[qq][u] >= thr);
          tot += __popc(mask[qq][u]);
        }
        if (lane == qq) n_pass = tot;
      }
      int base = 0;
      if (lane < nqw && n_pass > 0) base = atomicAdd(a.cand_cnt + q4 * 16 + h * 8 + lane, n_pass);
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const int qi = q4 * 16 + h * 8 + qq;
        int o = __shfl_sync(0xffffffffu, base, qq);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if ((mask[qq][u] >> lane) & 1u) {
            const int64_t at = (int64_t)qi * cap + o + __popc(mask[qq][u] & ((1u << lane) - 1u));
            a.cand_keys[at] = key[qq][u];
            a.cand_dbidx[at] = __ldg(a.img_dbidx + Q.img[(lane + 32 * u) * 64 + qi]);
          }
          o += __popc(mask[qq][u]);
        }
      }
    }
  }
  SSW_TR(11, warp == 2 && lane == 0);
  tc_fence_before();
  __syncthreads();
  SSW_TR(13, threadIdx.x == 0);
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_ALLOC);
}

template <int DIM, int NT, int NS, int NACC>
static int launch_scan_tc_t(ssw_db* db, const ScanTcArgs& a, cudaStream_t st, int grid, int64_t max_cta_images) {
  using Cfg = TcCfg<DIM, NT, NACC>;
  CUtensorMap tmap;
  int rc = make_tmap_f16_rows(&tmap, db->d_vecs, db->n_rows, DIM, NT);
  if (rc) return rc;
  size_t smem = (size_t)NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16 + qshared_bytes(a.k) + tc_smem_slack;
  ScanTcArgs a2 = a;
  a2.excl_slice_words = 0;
  if (a.excl) {   // keep the CTA's slice of the bitmaps in shared memory when it fits (227 KB per CTA)
    const int words = (int)(max_cta_images / 32) + 2;
    if (smem + (size_t)64 * words * 4 <= 232448) {
      a2.excl_slice_words = words;
      smem += (size_t)64 * words * 4;
    }
  }
  auto kern = scan_tc8_kernel<DIM, NT, NS, NACC>;
  SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(db, st);      // times the scan kernel alone (not the query preparation)
  // programmatic dependent launch: the kernel's set-up runs under the tail of the preparation kernel (an event
  // record between the two would serialise them, so not while profiling)
  SSW_CUDA(launch_kernel(kern, dim3(grid), dim3(kScanTc8Threads), smem, st, !db->prof_sampled, tmap, a2));
  prof_end(db, st);
  SSW_LAUNCHED();
  return SSW_OK;
}

bool scan_tc_supported(const ssw_db* db, int k) {
  return db->dtype == SSW_F16 && (db->dim == 256 || db->dim == 512 || db->dim == 768) && k <= kTcMaxK;
}

size_t scan_tc_workspace_bytes(int dim, int grid) { return (size_t)128 * (dim / 2) * 4 + 64 * 4 + (size_t)64 * grid * 4; }

// One pass over the database for queries [0, nq), nq <= 64.  `workspace` holds the prepared A operand
// (scan_tc_workspace_bytes); the preparation kernel also zeroes the queries' shared thresholds.
int launch_scan_tc(ssw_db* db, const float* d_queries, int nq, int k, const uint32_t* d_excl, uint64_t* d_cand_keys,
                   int32_t* d_cand_dbidx, int32_t* d_cand_cnt, uint64_t* d_gthr, void* workspace, cudaStream_t st,
                   const ScanTcGrid* sg) {
  const int grid = sg ? sg->grid : db->scan_grid;
  const int64_t max_cta_images = sg ? sg->max_cta_images : db->max_cta_images;
  uint32_t* a_img = static_cast<uint32_t*>(workspace);
  float* inv_scale = reinterpret_cast<float*>(a_img + (size_t)128 * (db->dim / 2));
  uint32_t* pub = reinterpret_cast<uint32_t*>(inv_scale + 64);
  scan_tc_prep_kernel<<<128, db->dim / 2, 0, st>>>(d_queries, nq, db->dim, a_img, inv_scale, d_gthr, d_cand_cnt, pub,
                                                     64 * grid);
  SSW_LAUNCHED();
  ScanTcArgs a{};
  a.a_img = a_img;
  a.inv_scale = inv_scale;
  a.nq = nq;
  a.k = k;
  a.excl = d_excl;
  a.excl_words = db->excl_words;
  a.last_bits = db->d_last_bits;
  a.row_ptr = db->d_row_ptr;
  a.img_dbidx = db->d_img_dbidx;
  a.orig_row = db->d_orig_row;
  a.part = sg ? sg->part : db->d_part;
  a.cand_keys = d_cand_keys;
  a.cand_dbidx = d_cand_dbidx;
  a.cand_cnt = d_cand_cnt;
  a.g_thr = d_gthr;
  a.pub = pub;
  a.row_base = db->row_base;
  a.stats = db->d_scan_stats;
  // shared memory: NS stages of NT*128 B + 64 lists of k (key, image) pairs (k <= 64 -> <= 48 KB)
  switch (db->dim) {
    case 256: return launch_scan_tc_t<256, 128, 10, 2>(db, a, st, grid, max_cta_images);
    case 512: return launch_scan_tc_t<512, 128, 10, 2>(db, a, st, grid, max_cta_images);
    // 768: A takes 384 of the 512 TMEM columns; one 128-column accumulator (MMA and epilogue alternate,
    // together well under the tile's HBM time) beats two 64-column ones (twice the per-tile overhead)
    case 768: return launch_scan_tc_t<768, 128, 10, 1>(db, a, st, grid, max_cta_images);
  }
  set_error("batched scan supports dim 256, 512 or 768");
  return SSW_ERR_INVALID;
}

}  // namespace ssw
