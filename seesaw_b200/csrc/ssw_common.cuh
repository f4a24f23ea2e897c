// Shared device/host helpers for libseesaw_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/seesaw_b200.h"

namespace ssw {

// ---------------------------------------------------------------------------- errors
void set_error(const std::string& msg);
extern int64_t g_launch_count;

#define SSW_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ssw::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return (_e == cudaErrorMemoryAllocation) ? SSW_ERR_OOM : SSW_ERR_CUDA;                 \
    }                                                                                        \
  } while (0)

#define SSW_REQUIRE(cond, msg)                                                               \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ssw::set_error(std::string(msg) + " (" #cond ")");                                     \
      return SSW_ERR_INVALID;                                                                \
    }                                                                                        \
  } while (0)

// After every kernel launch.  With SSW_DEBUG_SYNC=1 in the environment each launch is followed
// by a device synchronize so that an asynchronous fault is attributed to the right kernel.
bool debug_sync();
#define SSW_LAUNCHED()                                                                       \
  do {                                                                                       \
    ++ssw::g_launch_count;                                                                   \
    SSW_CUDA(cudaGetLastError());                                                            \
    if (ssw::debug_sync()) {                                                                 \
      cudaError_t _e = cudaDeviceSynchronize();                                              \
      if (_e != cudaSuccess) {                                                               \
        ssw::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) +              \
                       " kernel failed: " + cudaGetErrorString(_e));                         \
        return SSW_ERR_CUDA;                                                                 \
      }                                                                                      \
    }                                                                                        \
  } while (0)

// ---------------------------------------------------------------------------- keys
// A candidate is one 64-bit key: (order-preserving fp32 bits << 32) | ~global_row.
// Larger key == better candidate: higher score first, lower original row on equal score.
// Key 0 is reserved for "empty".
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f + 0.0f);   // -0.0 -> +0.0 so that they tie (numpy sorts them equal)
#else
  float g = f + 0.0f;
  uint32_t b;
  memcpy(&b, &g, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t global_row) {
  return ((uint64_t)f32_ordered(score) << 32) | (uint64_t)(0xFFFFFFFFu - global_row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) {
  return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
  return f32_from_ordered((uint32_t)(key >> 32));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------- programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in
// the stream still runs; pdl_wait() blocks until the predecessor has completed and its writes are visible
// (a no-op for a normally launched kernel), pdl_launch_dependents() lets the successor start launching.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch with (pdl = true) or without the programmatic-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Same, with a long suspend-time hint: the waiting warp sleeps in hardware until the phase completes
// instead of re-polling every few hundred cycles and stealing issue slots from the warps that share
// its scheduler (ncu on K2: 21M TRYWAIT executions from the producer/MMA warps).
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
  } while (!ok);
}
// 1-D bulk async copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}
// The canonical dot product of one stored row with a query: exactly the arithmetic of the streaming scan
// (scan1_kernel, ssw_scan.cu) for rows of type T — lane l owns the 16-byte chunks (c*32 + l), folds their
// elements in order into ONE fp32 FMA chain starting at 0, then the lanes are summed by the xor butterfly
// 16, 8, 4, 2, 1 (the transposed butterfly of the scan adds the same pairs in the same order).  Every kernel
// that re-scores a row (stage 2, exact-mode re-ranking) uses this, so a row has ONE fp32 score per storage
// type, whichever kernel computed it.  All 32 lanes receive the result.
template <typename T>
__device__ __forceinline__ float canon_dot(const T* __restrict__ row, const float* __restrict__ q, int dim, int lane) {
  constexpr int EPC = 16 / (int)sizeof(T);
  float s = 0.f;
  const int chunks = dim / (32 * EPC);
  for (int c = 0; c < chunks; ++c) {
    const int base = (c * 32 + lane) * EPC;
    const uint4 raw = *reinterpret_cast<const uint4*>(row + base);
    if constexpr (sizeof(T) == 2) {
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        s = fmaf(t.x, q[base + 2 * i], s);
        s = fmaf(t.y, q[base + 2 * i + 1], s);
      }
    } else {
      s = fmaf(__uint_as_float(raw.x), q[base + 0], s);
      s = fmaf(__uint_as_float(raw.y), q[base + 1], s);
      s = fmaf(__uint_as_float(raw.z), q[base + 2], s);
      s = fmaf(__uint_as_float(raw.w), q[base + 3], s);
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

// Pooled lower bound of a query's final k-th best score.  Every CTA publishes the best score it holds
// (order-preserving bits, 0 = none yet); lane l passes v[m] = value of CTA l + 32m.  The CTAs are split
// into ngp >= k groups (CTA c in group c mod ngp, ngp a power of two <= 128): the smallest group maximum
// is a score that ngp distinct images reach (distinct CTAs hold distinct images).  Returns 0 while some
// group has published nothing.
template <int VMAX>
__device__ __forceinline__ uint32_t pooled_group_min(const uint32_t* v, int ngp, int lane) {
  uint32_t t;
  if (ngp >= 32) {
    const int gpl = ngp >> 5;                 // groups per lane: 1, 2 or 4
    t = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t g = 0;
#pragma unroll
      for (int m = 0; m < VMAX; ++m) g = ((m & (gpl - 1)) == j && v[m] > g) ? v[m] : g;
      t = (j < gpl && g < t) ? g : t;
    }
  } else {
    t = 0;
#pragma unroll
    for (int m = 0; m < VMAX; ++m) t = v[m] > t ? v[m] : t;
    for (int sft = 16; sft >= ngp; sft >>= 1) {
      const uint32_t o = __shfl_xor_sync(0xffffffffu, t, sft);
      t = o > t ? o : t;
    }
  }
  return __reduce_min_sync(0xffffffffu, t);
}
#endif  // __CUDACC__

}  // namespace ssw
