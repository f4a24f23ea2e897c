// tcgen05 / TMEM / TMA building blocks shared by the batched scan (K2) and the kNN-graph build (K3).
//
// Both kernels are "resident-A, streaming-B" contractions:
//   A  = 128 rows x DIM fp16 that stay put for the whole pass (64 queries as hi/lo fp16 pairs for K2,
//        a 128-row block of V for K3).  A lives in TENSOR MEMORY (TS-form tcgen05.mma), written once
//        with tcgen05.st, so shared memory holds nothing but the B pipeline and its bandwidth is
//        spent on B alone.
//   B  = the database, streamed by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of NT rows x 64
//        fp16) through an NS-stage mbarrier ring.
//   D  = [128 lanes x NT columns] fp32 accumulators in TMEM, double-buffered, drained by four
//        epilogue warps with tcgen05.ld while the next tile's MMAs run.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue
// (warp w may touch TMEM lanes 32*(w%4) .. +31).
#pragma once
#include <cuda.h>

#include "ssw_common.cuh"

namespace ssw {

constexpr int kTcThreads = 192;
constexpr int kTcKChunk = 64;   // fp16 elements per stage row = 128 bytes = one SWIZZLE_128B span

// ---------------------------------------------------------------------------- PTX
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]   (kind::f16: fp16 inputs, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05 op of this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster (same TPC) execute one MMA of M = 256: each keeps its own 128 rows of A and
// D in its own tensor memory and HALF of every B tile in its own shared memory, so the B bytes an SM
// pulls from L2 per flop are halved.  The even CTA (rank 0) issues the MMAs; barriers that collect
// arrivals from both CTAs live in the even CTA: in the shared::cluster window bit 24 of a shared
// address selects the odd CTA of the pair (cute/arch/copy_sm100_tma.hpp, Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// SS form: both operands from shared memory (each CTA supplies its own 128 rows of A and its half of B)
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// this CTA's half of a B tile into its own shared memory; the bytes are counted on the EVEN CTA's barrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(tmap), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

// 32 consecutive TMEM columns of this thread's lane -> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 32 columns, mma-accumulator style: with j = lane%4, r = lane/4 the thread receives, for
// every 8-column group i (0..3):  v[4i+0] = (lane r, col 8i+2j)   v[4i+1] = (lane r,   col 8i+2j+1)
//                                  v[4i+2] = (lane r+8, col 8i+2j) v[4i+3] = (lane r+8, col 8i+2j+1)
// (cute/atom/copy_traits_sm100.hpp, SM100_TMEM_LOAD_16dp256b4x).  taddr's lane field selects the
// 16-lane half of the warp's quarter.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------- descriptors
// Instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (0), both K-major (bits 15,16 = 0),
// N>>3 at bits 17-22, M>>4 at bits 24-28  (cute/arch/mma_sm100_desc.hpp, InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor of a K-major SWIZZLE_128B operand tile whose rows are 128 bytes:
// start address >> 4 (bits 0-13), LBO unused (0), SBO = 8 rows * 128 B = 1024 B >> 4 (bits 32-45),
// version 1 (bits 46-47), layout type 2 = SWIZZLE_128B (bits 61-63).  Stepping K by 16 fp16 inside
// the 128-byte span is a +32 B (>>4: +2) bump of the start address.
__device__ __forceinline__ uint64_t make_bdesc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int DIM, int NT, int NACC = 2>
struct TcCfg {
  static_assert(DIM % kTcKChunk == 0, "DIM must be a multiple of 64");
  static constexpr int KC = DIM / kTcKChunk;            // stages (TMA boxes) per tile
  static constexpr int A_COLS = DIM / 2;                // TMEM columns of the resident A operand
  static constexpr int ACC_BASE = 0;
  static constexpr int A_BASE = NACC * NT;          // NACC accumulator buffers of NT columns, then A
  static constexpr int TMEM_USED = A_BASE + A_COLS;
  static constexpr int TMEM_ALLOC = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128
                                    : TMEM_USED <= 256 ? 256 : 512;
  static_assert(TMEM_USED <= 512, "tensor memory overflow");
  static constexpr int STAGE_BYTES = NT * 128;
  static constexpr uint32_t IDESC = make_idesc_f16(128, NT);
};

struct TcPipe {
  int stage = 0;
  uint32_t phase = 0;
  int ns;
  __device__ explicit TcPipe(int n) : ns(n) {}
  __device__ __forceinline__ void advance() {
    if (++stage == ns) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// Barrier block laid out after the stages: full[ns] empty[ns] tmem_full[2] tmem_empty[2] a_ready tmem_ptr
struct TcSmem {
  uint32_t stages;      // shared address of stage 0 (1024-aligned)
  uint32_t full, empty, tmem_full, tmem_empty, a_ready, tmem_ptr;
};

// `smem` is the raw dynamic shared memory; stages start at the next 1024-byte boundary (SWIZZLE_128B
// atoms are 1024 B), so launches reserve tc_smem_slack extra bytes.  Returns the carve-up; *aligned
// receives the generic pointer matching S.stages.
constexpr int tc_smem_slack = 1024;
__device__ __forceinline__ TcSmem tc_carve(uint8_t* smem, int ns, int stage_bytes, uint8_t** aligned) {
  TcSmem s;
  const uint32_t raw = smem_u32(smem);
  s.stages = (raw + 1023u) & ~1023u;
  *aligned = smem + (s.stages - raw);
  const uint32_t bars = s.stages + (uint32_t)ns * stage_bytes;
  s.full = bars;
  s.empty = bars + 8 * ns;
  s.tmem_full = bars + 16 * ns;
  s.tmem_empty = s.tmem_full + 16;
  s.a_ready = s.tmem_empty + 16;
  s.tmem_ptr = s.a_ready + 8;
  return s;
}
__host__ __device__ constexpr int tc_bar_bytes(int ns) { return 16 * ns + 16 + 16 + 8 + 8; }

// Called by every thread at kernel start.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t tc_setup(const TcSmem& s, int ns, uint32_t tmem_cols, const CUtensorMap* tmap,
                                             uint32_t epilogue_threads = 128, uint32_t a_writer_threads = 128) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmap);
    for (int i = 0; i < ns; ++i) {
      mbar_init(s.full + 8 * i, 1);
      mbar_init(s.empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s.tmem_full + 8 * i, 1);
      mbar_init(s.tmem_empty + 8 * i, epilogue_threads);
    }
    mbar_init(s.a_ready, a_writer_threads);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(s.tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(base) : "r"(s.tmem_ptr));
  return base;
}

// Host: 2-D tensor map over a row-major [n_rows, dim] fp16 matrix, box = {64 elements, box_rows} rows,
// SWIZZLE_128B (matches make_bdesc_sw128).  Out-of-range rows are zero-filled.
int make_tmap_f16_rows(CUtensorMap* out, const void* base, int64_t n_rows, int dim, int box_rows);

}  // namespace ssw
