// K3: exact all-pairs kNN-graph build on tcgen05 tiles with a fused per-row top-(k+1).
//
// Replaces   all_pairs = 1. - V @ V.T ; np.argsort(all_pairs, axis=-1)[:, :k+1]
// (seesaw/knn_graph.py:170-182, compute_exact_knn).  The N x N matrix is never materialised:
// a CTA keeps a 128-row block of V resident as the A operand (shared memory in the CTA-pair N=256 kernel that is
// the default, tensor memory in the single-CTA TS kernel used for dim 768 and large k1), streams every row of V through shared
// memory as B (TMA, SWIZZLE_128B), accumulates the dot products in TMEM and the four epilogue warps
// (one thread per output row) turn each accumulator tile into d = fp32(1 - dot) and keep the k1
// smallest (d, column) pairs of their row — one max tree and one compare per 32 columns in the
// common case.  The edge table of post_process_graph_df is produced on the device too (bottom).
// Ranking is on d, not on the dot: the rounding of 1 - dot merges near-equal dots into exact ties,
// which break by ascending column (SURVEY.md §7 "Graph ties are created by the 1 - dot rounding").
// Work: 2*N^2*DIM flops; V is read from HBM about once per wave of row blocks (the CTAs walk it in
// step and share it through L2).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ssw_db.h"
#include "ssw_tc.cuh"

namespace ssw {

struct KnnArgs {
  const __half* v;      // [n, DIM] fp16
  int64_t n;
  int k1;
  int64_t row_begin, row_end;   // output rows [row_begin, row_end)
  int32_t* out_idx;     // [(row_end-row_begin), k1]
  float* out_dist;
  unsigned int* wave_sync;  // one counter: TMA producers of all CTAs meet here between row blocks
};

// Per-row list state in shared memory (one 16-byte record per epilogue thread), so the out-of-line
// insert takes scalar arguments only and nothing spills to local memory.
struct KnnRowState {
  uint64_t maxkey;   // worst (largest) key held once the list is full
  int cnt, maxpos;
};

// Rare path of the epilogue: offer (d, col) to the row's list of the k1 smallest keys.  Columns arrive
// in ascending order, so on a tie with the current worst entry the newcomer (larger column) loses
// through the key compare.  Returns the row's distance threshold (inf until the list is full).
__device__ __noinline__ float knn_offer(KnnRowState* st, uint64_t* mylist, int k1, float d, int64_t col, int64_t n) {
  int cnt = st->cnt;
  const float cur = cnt < k1 ? INFINITY : key_score(st->maxkey);
  if (col >= n) return cur;      // zero-filled out-of-range columns of the last tile
  const uint64_t key = ((uint64_t)f32_ordered(d) << 32) | (uint32_t)col;
  if (cnt < k1) {
    mylist[cnt * 128] = key;
    st->cnt = ++cnt;
    if (cnt < k1) return INFINITY;
  } else if (key < st->maxkey) {
    mylist[st->maxpos * 128] = key;
  } else {
    return cur;
  }
  uint64_t mk = 0;
  int mp = 0;
  for (int s = 0; s < k1; ++s) {
    const uint64_t x = mylist[s * 128];
    if (x >= mk) {
      mk = x;
      mp = s;
    }
  }
  st->maxkey = mk;
  st->maxpos = mp;
  return key_score(mk);
}

// Dot-product value a column must EXCEED to be able to enter a full list.  Columns arrive in ascending
// order, so a newcomer that only ties the list's worst distance loses on the column: it needs
// d = fl(1 - x) < thr.  For thr in [0.5, 2] the subtraction 1 - thr is exact (Sterbenz) and x <= 1 - thr
// implies d >= thr: the bound is tight, which matters for degenerate rows (an all-zero vector ties every
// column at d = 1).  Below 0.5 one ulp of margin keeps the cheap test conservative; the exact test on d
// follows in the rare path.
__device__ __forceinline__ float knn_thr_dot(float thr) {
  if (thr == INFINITY) return -INFINITY;
  const float u = 1.0f - thr;
  return thr >= 0.5f ? u : u - 1.2e-7f;
}

// 32 accumulator columns (dots of this thread's row with columns col0 .. col0+31).  Common case: one
// max tree (FMNMX3) and one compare for the whole group.
__device__ __forceinline__ void knn_group(const uint32_t* v, float& thr, float& thr_dot, KnnRowState* st,
                                          uint64_t* mylist, int k1, int64_t col0, int64_t n) {
  float m8[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float a = fmaxf(fmaxf(__uint_as_float(v[8 * s]), __uint_as_float(v[8 * s + 1])), __uint_as_float(v[8 * s + 2]));
    const float b = fmaxf(fmaxf(__uint_as_float(v[8 * s + 3]), __uint_as_float(v[8 * s + 4])), __uint_as_float(v[8 * s + 5]));
    m8[s] = fmaxf(fmaxf(a, b), fmaxf(__uint_as_float(v[8 * s + 6]), __uint_as_float(v[8 * s + 7])));
  }
  const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  if (mx > thr_dot) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (m8[s] > thr_dot) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = __uint_as_float(v[8 * s + e]);
          if (x > thr_dot) {
            const float d = __fsub_rn(1.0f, x);
            if (d < thr) {
              thr = knn_offer(st, mylist, k1, d, col0 + 8 * s + e, n);
              thr_dot = knn_thr_dot(thr);
            }
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void tmem_ld_wait_regs32(uint32_t* x) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]),
                 "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15]),
                 "+r"(x[16]), "+r"(x[17]), "+r"(x[18]), "+r"(x[19]), "+r"(x[20]), "+r"(x[21]), "+r"(x[22]), "+r"(x[23]),
                 "+r"(x[24]), "+r"(x[25]), "+r"(x[26]), "+r"(x[27]), "+r"(x[28]), "+r"(x[29]), "+r"(x[30]), "+r"(x[31])
               :
               : "memory");
}

template <int DIM, int NT, int NS, int NACC = 2>
__global__ void __launch_bounds__(kTcThreads, 1) knn_kernel(const __grid_constant__ CUtensorMap tmap, const KnnArgs a) {
  using Cfg = TcCfg<DIM, NT, NACC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem;
  const TcSmem S = tc_carve(smem_raw, NS, Cfg::STAGE_BYTES, &smem);
  // after the barrier block: per-row candidate lists keys[k1][128], then the 128 row-state records
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + NS * Cfg::STAGE_BYTES + ((tc_bar_bytes(NS) + 15) / 16) * 16);
  KnnRowState* states = reinterpret_cast<KnnRowState*>(lists + (size_t)a.k1 * 128);
  const uint32_t tmem = tc_setup(S, NS, Cfg::TMEM_ALLOC, &tmap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t nblocks = (a.row_end - a.row_begin + 127) / 128;
  const int ntiles = (int)((a.n + NT - 1) / NT);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      TcPipe p(NS);
      for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        for (int t = 0; t < ntiles; ++t) {
          for (int kc = 0; kc < Cfg::KC; ++kc) {
            mbar_wait_parked(S.empty + 8 * p.stage, p.phase ^ 1);
            mbar_expect_tx(S.full + 8 * p.stage, Cfg::STAGE_BYTES);
            tma_load_2d(S.stages + p.stage * Cfg::STAGE_BYTES, &tmap, kc * kTcKChunk, t * NT, S.full + 8 * p.stage);
            p.advance();
          }
        }
      }
    }
    __syncwarp();       // the idle lanes must not reach the closing barrier ahead of lane 0
  } else if (warp == 1) {
    // ===== MMA issuer =====
    TcPipe p(NS);
    uint32_t it = 0, blk_phase = 0;
    for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
      mbar_wait_parked(S.a_ready, blk_phase);
      blk_phase ^= 1;
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t, ++it) {
        const uint32_t as = NACC == 2 ? (it & 1) : 0;
        mbar_wait_parked(S.tmem_empty + 8 * as, ((NACC == 2 ? (it >> 1) : it) & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < Cfg::KC; ++kc) {
          mbar_wait_parked(S.full + 8 * p.stage, p.phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t bdesc = make_bdesc_sw128(S.stages + p.stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ts(tmem + Cfg::ACC_BASE + as * NT, tmem + Cfg::A_BASE + kc * 32 + k * 8, bdesc + 2 * k,
                         Cfg::IDESC, (kc | k) != 0);
            tc_commit(S.empty + 8 * p.stage);
          }
          __syncwarp();
          p.advance();
        }
        if (lane == 0) tc_commit(S.tmem_full + 8 * as);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: one thread per output row =====
    const int q4 = warp & 3;
    const int lrow = q4 * 32 + lane;                     // TMEM lane == row within the block
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    uint64_t* mylist = lists + lrow;                      // slot s at mylist[s * 128]
    KnnRowState* st = states + lrow;
    const int k1 = a.k1;
    uint32_t it = 0;
    for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
      const int64_t row = a.row_begin + b * 128 + lrow;
      const bool row_ok = row < a.row_end;
      // ---- A operand: this thread's row of V into its TMEM lane (fp16 pairs are already packed).
      // The previous block's MMAs have all completed: its last accumulator was drained below.
      {
        const uint4* src = reinterpret_cast<const uint4*>(a.v + (row_ok ? row : 0) * (int64_t)DIM);
#pragma unroll 1
        for (int c = 0; c < Cfg::A_COLS / 32; ++c) {
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 w = row_ok ? __ldg(src + c * 8 + j) : make_uint4(0, 0, 0, 0);
            r[4 * j] = w.x;
            r[4 * j + 1] = w.y;
            r[4 * j + 2] = w.z;
            r[4 * j + 3] = w.w;
          }
          tmem_st32(lane_addr + Cfg::A_BASE + c * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(S.a_ready);
      }
      st->cnt = 0;
      st->maxpos = 0;
      st->maxkey = ~0ull;
      float thr = INFINITY, thr_dot = row_ok ? -INFINITY : INFINITY;   // padding rows (zero vectors) skip everything
      for (int t = 0; t < ntiles; ++t, ++it) {
        const uint32_t as = NACC == 2 ? (it & 1) : 0;
        mbar_wait(S.tmem_full + 8 * as, (NACC == 2 ? (it >> 1) : it) & 1);
        tc_fence_after();
        const int64_t col0 = (int64_t)t * NT;
        const uint32_t acc = lane_addr + Cfg::ACC_BASE + as * NT;
        uint32_t va[32], vb[32];
        tmem_ld32(acc, va);
        tmem_ld_wait_regs32(va);
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 64) {
          tmem_ld32(acc + c0 + 32, vb);
          knn_group(va, thr, thr_dot, st, mylist, k1, col0 + c0, a.n);
          tmem_ld_wait_regs32(vb);
          if (c0 + 64 < NT) tmem_ld32(acc + c0 + 64, va);
          knn_group(vb, thr, thr_dot, st, mylist, k1, col0 + c0 + 32, a.n);
          if (c0 + 64 < NT) tmem_ld_wait_regs32(va);
        }
        tc_fence_before();
        mbar_arrive(S.tmem_empty + 8 * as);
      }
      const int cnt = st->cnt;
      // ---- sort the row's k1 candidates ascending by (d, col) and write them out
      if (row_ok) {
        for (int i = 1; i < cnt; ++i) {
          const uint64_t x = mylist[i * 128];
          int j = i - 1;
          while (j >= 0 && mylist[j * 128] > x) {
            mylist[(j + 1) * 128] = mylist[j * 128];
            --j;
          }
          mylist[(j + 1) * 128] = x;
        }
        const int64_t o = (row - a.row_begin) * k1;
        for (int i = 0; i < k1; ++i) {
          const uint64_t x = i < cnt ? mylist[i * 128] : 0;
          a.out_idx[o + i] = i < cnt ? (int32_t)(x & 0xFFFFFFFFu) : -1;
          a.out_dist[o + i] = i < cnt ? f32_from_ordered((uint32_t)(x >> 32)) : INFINITY;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_ALLOC);
}

// ------------------------------------------------------------------------------------------
// K3, full-rate variant: CTA pairs + N = 256 MMAs.  Measured with the epilogue switched off, N = 128
// instructions cap the tensor pipe at ~1200 TFLOP/s while N = 256 reach ~1480 (per-instruction
// overhead).  Two 256-column accumulators fill tensor memory, so A moves to shared memory (SS form):
// a pair owns 256 output rows; each CTA keeps its 128 rows of A resident (DIM*256 bytes, loaded by TMA
// once per row block) and streams its 128-row half of every 256-row B tile; per MMA a CTA reads 4 KB
// of A and 4 KB of B from shared memory in 128 cycles — half the shared-memory bandwidth.
// ------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    knn3_kernel(const __grid_constant__ CUtensorMap tmap, const KnnArgs a, const int NS) {
  constexpr int NT = 256;                         // B tile rows per MMA (128 per CTA)
  constexpr int KC = DIM / kTcKChunk;
  constexpr int CHUNK_BYTES = 128 * 128;          // 128 rows x 64 fp16: one TMA box, one SWIZZLE_128B operand tile
  constexpr int A_BYTES = KC * CHUNK_BYTES;
  constexpr uint32_t IDESC = make_idesc_f16(256, NT);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_smem = (raw + 1023u) & ~1023u;                 // resident A: KC chunks
  const uint32_t stages = a_smem + A_BYTES;                       // NS chunks of B
  const uint32_t bars = stages + (uint32_t)NS * CHUNK_BYTES;
  const uint32_t full = bars, empty = bars + 8 * NS, tmem_full = bars + 16 * NS, tmem_empty = tmem_full + 16;
  const uint32_t a_ready = tmem_empty + 16, a_empty = a_ready + 8, tmem_ptr = a_empty + 8;
  uint8_t* after = smem_raw + (a_smem - raw) + A_BYTES + NS * CHUNK_BYTES + ((16 * NS + 16 + 16 + 8 + 8 + 8 + 15) / 16) * 16;
  uint64_t* lists = reinterpret_cast<uint64_t*>(after);
  KnnRowState* states = reinterpret_cast<KnnRowState*>(lists + (size_t)a.k1 * 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < NS; ++i) {
      mbar_init(full + 8 * i, 1);
      mbar_init(empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 256);
    }
    mbar_init(a_ready, 1);
    mbar_init(a_empty, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_ptr));

  const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int64_t nblocks2 = (a.row_end - a.row_begin + 255) / 256;
  const int ntiles = (int)((a.n + NT - 1) / NT);

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, blk = 0;
      for (int64_t b = pair; b < nblocks2; b += npairs, ++blk) {
        // Every CTA streams ALL of V once per row block.  V (1 GB at 1M x 512) does not fit in L2, so the
        // CTAs only share each B tile through L2 while they walk V in step; left alone they drift apart
        // over tens of blocks and the build turns HBM-bound (measured on the full 1M build: 1085 ms without,
        // 920 ms with this meeting point).  The producers therefore meet between row blocks.
        if (blk > 0 && a.wave_sync) {
          const unsigned int want = blk * gridDim.x;
          unsigned int seen;
          unsigned long long t0, t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.wave_sync) : "memory");
            if (seen < want) {
              __nanosleep(500);
              // the meeting point only improves L2 sharing: if some CTA is not even resident yet (GPU shared
              // with another kernel) stop waiting instead of risking a dead lock
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
              if (t1 - t0 > 200000000ull) break;
            }
          } while (seen < want);
        }
        // this CTA's 128 rows of A, once the previous block's MMAs have all completed
        if (blk > 0) mbar_wait_parked(a_empty, (blk - 1) & 1);
        if (leader) mbar_expect_tx(a_ready, 2 * A_BYTES);
        const int arow = (int)(a.row_begin + b * 256 + (int64_t)rank * 128);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d_2sm(a_smem + kc * CHUNK_BYTES, &tmap, kc * kTcKChunk, arow, a_ready);
        for (int t = 0; t < ntiles; ++t) {
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait_parked(empty + 8 * stage, phase ^ 1);
            if (leader) mbar_expect_tx(full + 8 * stage, 2 * CHUNK_BYTES);
            tma_load_2d_2sm(stages + stage * CHUNK_BYTES, &tmap, kc * kTcKChunk, t * NT + (int)rank * (NT / 2),
                            full + 8 * stage);
            if (++stage == NS) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (a.wave_sync) atomicAdd(a.wave_sync, 1u);       // all loads of this row block are issued
      }
    }
    __syncwarp();       // the idle lanes must not reach the closing barrier ahead of lane 0
  } else if (warp == 1) {
    // ===== MMA issuer (even CTA only) =====
    if (leader) {
      int stage = 0;
      uint32_t phase = 0, it = 0, blk = 0;
      for (int64_t b = pair; b < nblocks2; b += npairs, ++blk) {
        mbar_wait_parked(a_ready, blk & 1);
        tc_fence_after();
        for (int t = 0; t < ntiles; ++t, ++it) {
          const uint32_t as = it & 1;
          mbar_wait_parked(tmem_empty + 8 * as, ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait_parked(full + 8 * stage, phase);
            tc_fence_after();
            if (lane == 0) {
              const uint64_t adesc = make_bdesc_sw128(a_smem + kc * CHUNK_BYTES);
              const uint64_t bdesc = make_bdesc_sw128(stages + stage * CHUNK_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_f16_ss_2sm(tmem + as * NT, adesc + 2 * k, bdesc + 2 * k, IDESC, (kc | k) != 0);
              tc_commit_2sm(empty + 8 * stage);
            }
            __syncwarp();
            if (++stage == NS) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (lane == 0) {
            tc_commit_2sm(tmem_full + 8 * as);
            if (t == ntiles - 1) tc_commit_2sm(a_empty);       // A may be replaced once these MMAs are done
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: one thread per output row (both CTAs) =====
    const int q4 = warp & 3;
    const int lrow = q4 * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    uint64_t* mylist = lists + lrow;
    KnnRowState* st = states + lrow;
    const int k1 = a.k1;
    uint32_t it = 0;
    for (int64_t b = pair; b < nblocks2; b += npairs) {
      const int64_t row = a.row_begin + b * 256 + (int64_t)rank * 128 + lrow;
      const bool row_ok = row < a.row_end;
      st->cnt = 0;
      st->maxpos = 0;
      st->maxkey = ~0ull;
      float thr = INFINITY, thr_dot = row_ok ? -INFINITY : INFINITY;   // padding rows (zero vectors) skip everything
      for (int t = 0; t < ntiles; ++t, ++it) {
        const uint32_t as = it & 1;
        mbar_wait(tmem_full + 8 * as, (it >> 1) & 1);
        tc_fence_after();
        const int64_t col0 = (int64_t)t * NT;
        const uint32_t acc = lane_addr + as * NT;
        uint32_t va[32], vb[32];
        tmem_ld32(acc, va);
        tmem_ld_wait_regs32(va);
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 64) {
          tmem_ld32(acc + c0 + 32, vb);
          knn_group(va, thr, thr_dot, st, mylist, k1, col0 + c0, a.n);
          tmem_ld_wait_regs32(vb);
          if (c0 + 64 < NT) tmem_ld32(acc + c0 + 64, va);
          knn_group(vb, thr, thr_dot, st, mylist, k1, col0 + c0 + 32, a.n);
          if (c0 + 64 < NT) tmem_ld_wait_regs32(va);
        }
        tc_fence_before();
        mbar_arrive_leader(tmem_empty + 8 * as);
      }
      const int cnt = st->cnt;
      if (row_ok) {
        for (int i = 1; i < cnt; ++i) {
          const uint64_t x = mylist[i * 128];
          int j = i - 1;
          while (j >= 0 && mylist[j * 128] > x) {
            mylist[(j + 1) * 128] = mylist[j * 128];
            --j;
          }
          mylist[(j + 1) * 128] = x;
        }
        const int64_t o = (row - a.row_begin) * k1;
        for (int i = 0; i < k1; ++i) {
          const uint64_t x = i < cnt ? mylist[i * 128] : 0;
          a.out_idx[o + i] = i < cnt ? (int32_t)(x & 0xFFFFFFFFu) : -1;
          a.out_dist[o + i] = i < cnt ? f32_from_ordered((uint32_t)(x >> 32)) : INFINITY;
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem, 512);
}

int make_tmap_f16_rows(CUtensorMap* out, const void* base, int64_t n_rows, int dim, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SSW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return SSW_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)std::max<int64_t>(n_rows, 1)};
  const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kTcKChunk, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return SSW_ERR_CUDA;
  }
  return SSW_OK;
}

template <int DIM, int NT>
static int launch_knn_t(int sm_count, const KnnArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<DIM, NT>;
  const size_t list_bytes = (size_t)a.k1 * 128 * 8 + 128 * sizeof(KnnRowState);
  CUtensorMap tmap;
  int rc = make_tmap_f16_rows(&tmap, a.v, a.n, DIM, NT);
  if (rc) return rc;
  const int64_t nblocks = (a.row_end - a.row_begin + 127) / 128;
  const int grid = (int)std::min<int64_t>(sm_count, nblocks);
  auto go = [&](auto kern, int NSv) -> int {
    const size_t smem = (size_t)NSv * Cfg::STAGE_BYTES + ((tc_bar_bytes(NSv) + 15) / 16) * 16 + list_bytes + tc_smem_slack;
    SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTcThreads, smem, st>>>(tmap, a);
    SSW_LAUNCHED();
    return SSW_OK;
  };
  // stage count by the shared memory left after the candidate lists (k1 <= 64 -> <= 64 KB)
  if (list_bytes <= 24 * 1024) return go(knn_kernel<DIM, NT, 12>, 12);
  return go(knn_kernel<DIM, NT, 8>, 8);
}

// SS / N = 256 variant: needs >= 3 B stages next to the resident A block and the candidate lists
template <int DIM>
static int launch_knn3_t(int sm_count, const KnnArgs& a, cudaStream_t st, bool* launched) {
  *launched = false;
  const size_t list_bytes = (size_t)a.k1 * 128 * 8 + 128 * sizeof(KnnRowState);
  const size_t fixed = (size_t)(DIM / 64) * 16384 + list_bytes + tc_smem_slack + 512;
  if (fixed + 3 * 16384 > 232448) return SSW_OK;
  const int NS = (int)std::min<size_t>((232448 - fixed) / 16384, 8);
  CUtensorMap tmap;
  int rc = make_tmap_f16_rows(&tmap, a.v, a.n, DIM, 128);      // box = 128 rows x 64 fp16 (A block chunk / B half tile)
  if (rc) return rc;
  const int64_t nblocks2 = (a.row_end - a.row_begin + 255) / 256;
  const int grid = 2 * (int)std::min<int64_t>(sm_count / 2, nblocks2);
  const size_t smem = fixed + (size_t)NS * 16384;
  auto kern = knn3_kernel<DIM>;
  SSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // inter-block meeting point of the producers: one word per launch, allocated and freed in stream order
  unsigned int* sync_word = nullptr;
  SSW_CUDA(cudaMallocAsync((void**)&sync_word, 128, st));
  SSW_CUDA(cudaMemsetAsync(sync_word, 0, 4, st));
  KnnArgs a2 = a;
  a2.wave_sync = sync_word;
  kern<<<grid, kTcThreads, smem, st>>>(tmap, a2, NS);
  ++ssw::g_launch_count;
  const cudaError_t le = cudaGetLastError();
  if (sync_word) cudaFreeAsync(sync_word, st);
  if (le != cudaSuccess) {
    set_error(std::string("knn3_kernel launch: ") + cudaGetErrorString(le));
    return SSW_ERR_CUDA;
  }
  *launched = true;
  return SSW_OK;
}

static int launch_knn(int sm_count, const KnnArgs& a, int dim, cudaStream_t st) {
  // CTA pairs with N = 256 SS MMAs (knn3_kernel) whenever the resident A block, three B stages and the candidate
  // lists fit in shared memory (dim 256 / 512, any k1 <= 64 at dim 256, k1 <= ~40 at dim 512); otherwise the
  // single-CTA TS kernel (dim 768, large k1)
  if (dim == 256 || dim == 512) {
    bool launched = false;
    const int rc = dim == 256 ? launch_knn3_t<256>(sm_count, a, st, &launched) : launch_knn3_t<512>(sm_count, a, st, &launched);
    if (rc || launched) return rc;
  }
  switch (dim) {
    case 256: return launch_knn_t<256, 128>(sm_count, a, st);
    case 512: return launch_knn_t<512, 128>(sm_count, a, st);
    case 768: return launch_knn_t<768, 64>(sm_count, a, st);
  }
  set_error("kNN build supports dim 256, 512 or 768");
  return SSW_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------
// Edge table on the device: post_process_graph_df (seesaw/knn_graph.py:142-168) applied to the candidate
// table [rows, k1] (already ordered by (distance, column)): distances clipped at 0, self edges dropped,
// dst_rank = 1.. in table order (== rank('first') of the clipped, non-decreasing distances), one rank-0
// zero-distance self edge per vertex, rows ordered by (src_vertex, dst_rank).  Three small kernels: per-row
// counts + block-local scan, scan of the block totals, scatter.
// ------------------------------------------------------------------------------------------
constexpr int kEdgeBlock = 1024;

__global__ void __launch_bounds__(kEdgeBlock) knn_edge_count_kernel(const int32_t* __restrict__ idx, int64_t rows, int k1,
                                                                    int64_t src_offset, int32_t* __restrict__ local_off,
                                                                    int64_t* __restrict__ block_total) {
  __shared__ int s_warp[32];
  const int64_t r = blockIdx.x * (int64_t)kEdgeBlock + threadIdx.x;
  int c = 0;
  if (r < rows) {
    const int32_t self = (int32_t)(r + src_offset);
    c = 1;
    for (int j = 0; j < k1; ++j) {
      const int32_t v = idx[r * k1 + j];
      c += (v >= 0 && v != self);
    }
  }
  int incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, m);
    if (lane >= m) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, w, m);
      if (lane >= m) w += o;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  const int before = (warp ? s_warp[warp - 1] : 0) + incl - c;
  if (r < rows) local_off[r] = before;
  if (threadIdx.x == kEdgeBlock - 1) block_total[blockIdx.x] = before + c;
}

__global__ void knn_edge_scan_kernel(int64_t* block_total, int64_t n_blocks, int64_t* total_out) {
  // single thread: n_blocks = rows / 1024 (~1000 for a million vertices)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t run = 0;
    for (int64_t b = 0; b < n_blocks; ++b) {
      const int64_t t = block_total[b];
      block_total[b] = run;
      run += t;
    }
    *total_out = run;
  }
}

__global__ void __launch_bounds__(kEdgeBlock) knn_edge_write_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist,
                                                                    int64_t rows, int k1, int64_t src_offset,
                                                                    const int32_t* __restrict__ local_off,
                                                                    const int64_t* __restrict__ block_base, int32_t* __restrict__ src,
                                                                    int32_t* __restrict__ dst, float* __restrict__ distance,
                                                                    int32_t* __restrict__ rank) {
  const int64_t r = blockIdx.x * (int64_t)kEdgeBlock + threadIdx.x;
  if (r >= rows) return;
  const int32_t self = (int32_t)(r + src_offset);
  int64_t o = block_base[blockIdx.x] + local_off[r];
  src[o] = self;
  dst[o] = self;
  distance[o] = 0.f;
  rank[o] = 0;
  int rk = 0;
  for (int j = 0; j < k1; ++j) {
    const int32_t v = idx[r * k1 + j];
    if (v < 0 || v == self) continue;
    ++o;
    ++rk;
    src[o] = self;
    dst[o] = v;
    distance[o] = fmaxf(dist[r * k1 + j], 0.f);
    rank[o] = rk;
  }
}

int knn_exact_refine(const float* d_v32, int64_t n, int dim, int k1, int kc, int64_t row_begin, int64_t rows,
                     const int32_t* d_cand_idx, const float* d_cand_dist, double err, int32_t* d_out_idx, float* d_out_dist,
                     cudaStream_t st, int64_t* rows_rescanned);

static thread_local int64_t t_knn_exact_rows = 0, t_knn_rescanned = 0;
static thread_local double t_knn_rho = 0.0;

}  // namespace ssw

using namespace ssw;

extern "C" {

int ssw_knn_build_device(int device, const void* d_vectors_f16, int64_t n, int dim, int k1, int64_t row_begin,
                         int64_t row_end, int32_t* d_out_idx, float* d_out_dist, void* stream);

// Candidates of rows [row_begin, row_end) with the vectors already on the device: d_v16 = the fp16 copy the tensor
// cores multiply, d_v32 = the caller's float32 values (or null).  With float32 values that are not fp16-representable
// the tensor-core pass proposes kc > k1 candidates and ssw_knn_exact.cu re-ranks / certifies them in float32.
static int knn_candidates_on_device(int device, const float* d_v32, const void* d_v16, int64_t n, int dim, int k1,
                                    int64_t row_begin, int64_t row_end, int32_t* d_idx, float* d_dist, cudaStream_t st) {
  t_knn_exact_rows = t_knn_rescanned = 0;
  t_knn_rho = 0.0;
  double rho = 0.0, vmax = 0.0;
  if (d_v32) {
    float* d_stats = nullptr;
    float stats[2] = {0.f, 0.f};
    SSW_CUDA(cudaMalloc((void**)&d_stats, 8));
    int rc = launch_row_error_stats(d_v32, n, dim, d_stats, st);
    if (!rc && cudaMemcpyAsync(stats, d_stats, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = SSW_ERR_CUDA;
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = SSW_ERR_CUDA;
    cudaFree(d_stats);
    if (rc) return rc;
    rho = std::sqrt((double)stats[0]) * (1.0 + 1e-6);
    vmax = std::sqrt((double)stats[1]) * (1.0 + 1e-6);
  }
  t_knn_rho = rho;
  if (rho == 0.0)       // fp16 input, or float32 values the fp16 copy holds exactly: the tensor-core result is the answer
    return ssw_knn_build_device(device, d_v16, n, dim, k1, row_begin, row_end, d_idx, d_dist, st);
  const int64_t rows = row_end - row_begin;
  const int kc = (int)std::min<int64_t>(std::min<int64_t>(n, SSW_MAX_KNN_K1), (int64_t)k1 + std::max(16, k1));
  int32_t* d_cidx = nullptr;
  float* d_cdist = nullptr;
  SSW_CUDA(cudaMalloc((void**)&d_cidx, std::max<size_t>((size_t)rows * kc, 1) * 4));
  cudaError_t e = cudaMalloc((void**)&d_cdist, std::max<size_t>((size_t)rows * kc, 1) * 4);
  if (e != cudaSuccess) {
    cudaFree(d_cidx);
    set_error(std::string("cudaMalloc(candidates): ") + cudaGetErrorString(e));
    return SSW_ERR_OOM;
  }
  int rc = ssw_knn_build_device(device, d_v16, n, dim, kc, row_begin, row_end, d_cidx, d_cdist, st);
  const double err = 2.0 * rho * vmax + vmax * vmax * (double)dim * 1.8e-7 + 3e-7;
  int64_t rescanned = 0;
  if (!rc) rc = knn_exact_refine(d_v32, n, dim, k1, kc, row_begin, rows, d_cidx, d_cdist, err, d_idx, d_dist, st, &rescanned);
  cudaFree(d_cidx);
  cudaFree(d_cdist);
  t_knn_exact_rows = rows;
  t_knn_rescanned = rescanned;
  return rc;
}

int ssw_knn_exact_stats(int64_t* rows_refined, int64_t* rows_rescanned, double* rho) {
  if (rows_refined) *rows_refined = t_knn_exact_rows;
  if (rows_rescanned) *rows_rescanned = t_knn_rescanned;
  if (rho) *rho = t_knn_rho;
  return SSW_OK;
}

int ssw_knn_build_device(int device, const void* d_vectors_f16, int64_t n, int dim, int k1, int64_t row_begin,
                         int64_t row_end, int32_t* d_out_idx, float* d_out_dist, void* stream) {
  SSW_REQUIRE(d_vectors_f16 != nullptr && d_out_idx != nullptr && d_out_dist != nullptr, "null argument");
  SSW_REQUIRE(n > 0 && n < (int64_t)0x7FFFFFFF, "n out of range");
  SSW_REQUIRE(k1 >= 1 && k1 <= SSW_MAX_KNN_K1 && k1 <= n, "k1 must be in [1, min(n, SSW_MAX_KNN_K1)]");
  SSW_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, "bad row range");
  int sms = 0;
  int rc = ensure_device(device, &sms);
  if (rc) return rc;
  if (row_begin == row_end) return SSW_OK;
  KnnArgs a{static_cast<const __half*>(d_vectors_f16), n, k1, row_begin, row_end, d_out_idx, d_out_dist, nullptr};
  return launch_knn(sms, a, dim, (cudaStream_t)stream);
}

int ssw_knn_build(int device, const void* vectors, int dtype_in, int64_t n, int dim, int k1, int64_t row_begin,
                  int64_t row_end, int32_t* out_idx, float* out_dist) {
  SSW_REQUIRE(vectors != nullptr && out_idx != nullptr && out_dist != nullptr, "null argument");
  SSW_REQUIRE(dtype_in == SSW_F16 || dtype_in == SSW_F32, "dtype_in must be SSW_F32 or SSW_F16");
  SSW_REQUIRE(n > 0 && n < (int64_t)0x7FFFFFFF, "n out of range");
  SSW_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, "bad row range");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  const size_t es = dtype_in == SSW_F16 ? 2 : 4;
  void *d_in = nullptr, *d_v = nullptr;
  int32_t* d_idx = nullptr;
  float* d_dist = nullptr;
  const size_t nout = (size_t)(row_end - row_begin) * k1;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_in);
    if (d_v != d_in) cudaFree(d_v);
    cudaFree(d_idx);
    cudaFree(d_dist);
  };
  auto chk = [&](cudaError_t e, const char* what) -> int {
    if (e == cudaSuccess) return SSW_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cleanup();
    return e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA;
  };
  if ((rc = chk(cudaMalloc(&d_in, (size_t)n * dim * es), "cudaMalloc(vectors)"))) return rc;
  if ((rc = chk(cudaMemcpy(d_in, vectors, (size_t)n * dim * es, cudaMemcpyHostToDevice), "cudaMemcpy(vectors)"))) return rc;
  if (dtype_in == SSW_F32) {
    if ((rc = chk(cudaMalloc(&d_v, (size_t)n * dim * 2), "cudaMalloc(fp16 vectors)"))) return rc;
    if ((rc = launch_convert_rows(d_in, SSW_F32, d_v, SSW_F16, n * dim, st))) {
      cleanup();
      return rc;
    }
  } else {
    d_v = d_in;
  }
  if ((rc = chk(cudaMalloc((void**)&d_idx, std::max<size_t>(nout, 1) * 4), "cudaMalloc(out_idx)"))) return rc;
  if ((rc = chk(cudaMalloc((void**)&d_dist, std::max<size_t>(nout, 1) * 4), "cudaMalloc(out_dist)"))) return rc;
  SSW_REQUIRE(k1 >= 1 && k1 <= SSW_MAX_KNN_K1 && k1 <= n, "k1 must be in [1, min(n, SSW_MAX_KNN_K1)]");
  rc = knn_candidates_on_device(device, dtype_in == SSW_F32 ? static_cast<const float*>(d_in) : nullptr, d_v, n, dim, k1, row_begin,
                                row_end, d_idx, d_dist, st);
  if (rc) {
    cleanup();
    return rc;
  }
  if ((rc = chk(cudaDeviceSynchronize(), "knn kernel"))) return rc;
  if ((rc = chk(cudaMemcpy(out_idx, d_idx, nout * 4, cudaMemcpyDeviceToHost), "cudaMemcpy(out_idx)"))) return rc;
  if ((rc = chk(cudaMemcpy(out_dist, d_dist, nout * 4, cudaMemcpyDeviceToHost), "cudaMemcpy(out_dist)"))) return rc;
  cleanup();
  return SSW_OK;
}

int ssw_knn_edges_device(int device, const int32_t* d_idx, const float* d_dist, int64_t rows, int k1, int64_t src_offset,
                         int32_t* d_src, int32_t* d_dst, float* d_distance, int32_t* d_rank, int64_t* d_total,
                         void* d_workspace, void* stream) {
  SSW_REQUIRE(d_idx && d_dist && d_src && d_dst && d_distance && d_rank && d_total && d_workspace, "null argument");
  SSW_REQUIRE(rows >= 0 && k1 >= 1, "bad shape");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_blocks = (rows + kEdgeBlock - 1) / kEdgeBlock;
  int64_t* block_total = static_cast<int64_t*>(d_workspace);                         // [n_blocks]
  int32_t* local_off = reinterpret_cast<int32_t*>(block_total + std::max<int64_t>(n_blocks, 1));   // [rows]
  if (rows == 0) {
    SSW_CUDA(cudaMemsetAsync(d_total, 0, 8, st));
    return SSW_OK;
  }
  knn_edge_count_kernel<<<(int)n_blocks, kEdgeBlock, 0, st>>>(d_idx, rows, k1, src_offset, local_off, block_total);
  SSW_LAUNCHED();
  knn_edge_scan_kernel<<<1, 32, 0, st>>>(block_total, n_blocks, d_total);
  SSW_LAUNCHED();
  knn_edge_write_kernel<<<(int)n_blocks, kEdgeBlock, 0, st>>>(d_idx, d_dist, rows, k1, src_offset, local_off, block_total, d_src,
                                                               d_dst, d_distance, d_rank);
  SSW_LAUNCHED();
  return SSW_OK;
}

int ssw_knn_edges_workspace_bytes(int64_t rows, int64_t* bytes) {
  SSW_REQUIRE(bytes != nullptr && rows >= 0, "bad argument");
  *bytes = (std::max<int64_t>((rows + kEdgeBlock - 1) / kEdgeBlock, 1)) * 8 + std::max<int64_t>(rows, 1) * 4;
  return SSW_OK;
}

int ssw_knn_graph(int device, const void* vectors, int dtype_in, int64_t n, int dim, int n_neighbors, int32_t* out_src,
                  int32_t* out_dst, float* out_distance, int32_t* out_rank, int64_t capacity, int64_t* out_edges) {
  SSW_REQUIRE(vectors && out_src && out_dst && out_distance && out_rank && out_edges, "null argument");
  SSW_REQUIRE(dtype_in == SSW_F16 || dtype_in == SSW_F32, "dtype_in must be SSW_F32 or SSW_F16");
  SSW_REQUIRE(n > 0 && n < (int64_t)0x7FFFFFFF && n_neighbors >= 0, "bad shape");
  const int k1 = (int)std::min<int64_t>((int64_t)n_neighbors + 1, n);
  SSW_REQUIRE(k1 <= SSW_MAX_KNN_K1, "n_neighbors + 1 exceeds SSW_MAX_KNN_K1");
  SSW_REQUIRE(capacity >= n * (k1 + 1), "output capacity must be n * (min(n_neighbors + 1, n) + 1) edges");
  int rc = ensure_device(device, nullptr);
  if (rc) return rc;
  const size_t es = dtype_in == SSW_F16 ? 2 : 4;
  const size_t cand = (size_t)n * k1, cap = (size_t)n * (k1 + 1);
  int64_t ws_bytes = 0;
  ssw_knn_edges_workspace_bytes(n, &ws_bytes);
  void *d_in = nullptr, *d_v = nullptr, *d_ws = nullptr;
  int32_t *d_idx = nullptr, *d_src = nullptr, *d_dst = nullptr, *d_rank = nullptr;
  float *d_dist = nullptr, *d_edist = nullptr;
  int64_t* d_total = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_in);
    if (d_v != d_in) cudaFree(d_v);
    cudaFree(d_ws);
    cudaFree(d_idx);
    cudaFree(d_dist);
    cudaFree(d_src);
    cudaFree(d_dst);
    cudaFree(d_edist);
    cudaFree(d_rank);
    cudaFree(d_total);
  };
  auto chk = [&](cudaError_t e, const char* what) -> int {
    if (e == cudaSuccess) return SSW_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cleanup();
    return e == cudaErrorMemoryAllocation ? SSW_ERR_OOM : SSW_ERR_CUDA;
  };
#define KG_TRY(expr)                   \
  if ((rc = chk((expr), #expr))) return rc
  KG_TRY(cudaMalloc(&d_in, (size_t)n * dim * es));
  KG_TRY(cudaMemcpy(d_in, vectors, (size_t)n * dim * es, cudaMemcpyHostToDevice));
  if (dtype_in == SSW_F32) {
    KG_TRY(cudaMalloc(&d_v, (size_t)n * dim * 2));
    if ((rc = launch_convert_rows(d_in, SSW_F32, d_v, SSW_F16, n * dim, nullptr))) {
      cleanup();
      return rc;
    }
  } else {
    d_v = d_in;
  }
  KG_TRY(cudaMalloc((void**)&d_idx, cand * 4));
  KG_TRY(cudaMalloc((void**)&d_dist, cand * 4));
  KG_TRY(cudaMalloc((void**)&d_src, cap * 4));
  KG_TRY(cudaMalloc((void**)&d_dst, cap * 4));
  KG_TRY(cudaMalloc((void**)&d_edist, cap * 4));
  KG_TRY(cudaMalloc((void**)&d_rank, cap * 4));
  KG_TRY(cudaMalloc((void**)&d_total, 8));
  KG_TRY(cudaMalloc(&d_ws, (size_t)ws_bytes));
  rc = knn_candidates_on_device(device, dtype_in == SSW_F32 ? static_cast<const float*>(d_in) : nullptr, d_v, n, dim, k1, 0, n,
                                d_idx, d_dist, nullptr);
  if (!rc) rc = ssw_knn_edges_device(device, d_idx, d_dist, n, k1, 0, d_src, d_dst, d_edist, d_rank, d_total, d_ws, nullptr);
  if (rc) {
    cleanup();
    return rc;
  }
  KG_TRY(cudaDeviceSynchronize());
  int64_t total = 0;
  KG_TRY(cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost));
  KG_TRY(cudaMemcpy(out_src, d_src, (size_t)total * 4, cudaMemcpyDeviceToHost));
  KG_TRY(cudaMemcpy(out_dst, d_dst, (size_t)total * 4, cudaMemcpyDeviceToHost));
  KG_TRY(cudaMemcpy(out_distance, d_edist, (size_t)total * 4, cudaMemcpyDeviceToHost));
  KG_TRY(cudaMemcpy(out_rank, d_rank, (size_t)total * 4, cudaMemcpyDeviceToHost));
#undef KG_TRY
  *out_edges = total;
  cleanup();
  return SSW_OK;
}

}  // extern "C"
