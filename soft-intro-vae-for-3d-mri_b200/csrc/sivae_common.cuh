// sivae_common.cuh -- shared device/host helpers for libsivae.so (sm_100a only).
//
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// shared-memory / instruction descriptor encodings; error plumbing for the C ABI; Philox4x32-10.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sivae.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

// ------------------------------------------------------------------------------------------------
// host side: error handling, launch accounting
// ------------------------------------------------------------------------------------------------
namespace sivae {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

#define SIVAE_CHECK(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      sivae::set_error(__VA_ARGS__); \
      return -2;                    \
    }                               \
  } while (0)

#define SIVAE_LAUNCH_OK(name)                                              \
  do {                                                                     \
    sivae::g_launches.fetch_add(1, std::memory_order_relaxed);             \
    if (sivae::check_cuda(cudaGetLastError(), name) != 0) return -1;       \
  } while (0)

__host__ __device__ inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Dropout key: a per-call constant plus an optional device-side epoch counter (so a CUDA graph that bakes the
// constant still draws fresh masks on every replay).  sivae_set_seed_counter / sivae_advance_seed_counter.
struct SeedRef {
  unsigned long long seed;
  const unsigned long long* ctr;
};
SeedRef make_seed_ref(unsigned long long seed);

// Tensor-map encode entry point (driver API resolved at run time: no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// bf16 tensor map with 128B swizzle, zero OOB fill.  dims[0] is the contiguous dimension.
int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

}  // namespace sivae

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace sivae {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- explicit shared-memory accesses ---------------------------------------------------------------
// The kernels carve their dynamic shared memory from a 1024-byte-aligned pointer computed with integer arithmetic,
// which makes the compiler lose the address space and emit generic LD.E / ST.E.  Hot loops therefore use 32-bit
// shared-window addresses (smem_u32) and these wrappers (LDS / STS).
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128_f32(uint32_t addr) {   // addr 16-byte aligned
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (sticky launch failure) instead of hanging the GPU box.
// (Parking the waiter -- try_wait with a suspend-time hint, which ptxas turns into TRYWAIT + NANOSLEEP.SYNCS -- was
// measured: the spin loop's instructions disappear, but the step got 0.6 ms SLOWER (58.5 -> 59.1 ms, two A/B pairs) and
// no kernel faster: wake-up latency costs more than the issue slots the spinning warps take.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      printf("sivae: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Wait only until the bulk stores have READ their shared-memory source: enough before a CTA exits (its smem may be
// handed to the next CTA); the global writes themselves complete asynchronously and are ordered by kernel completion.
__device__ __forceinline__ void tma_store_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4         bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4    bits [46,48) version = 1      bits [61,64) layout = 2 (SWIZZLE_128B)
// K-major operand tile (rows of 64 bf16 = 128 B, 8-row swizzle atoms): LBO unused (1), SBO = 1024 B.
// MN-major operand tile (K-rows of 64 bf16 along MN): LBO = byte distance between 64-element MN chunks,
//   SBO = byte distance between groups of 8 K-rows (1024 B when rows are dense).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32.  a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- misc ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// BatchNorm statistics of one staged output tile (128 rows x 64 bf16 channels, SWIZZLE_128B): the calling warp sums
// rows [32*q, 32*q + 32) for the channel pair (2*lane, 2*lane + 1) -- conflict-free 4-byte LDS, all loads of a batch in
// flight before use.  rows_w = patch width (rows are (h, w) with w fastest); rows outside [0,wlim) x [0,hlim) skipped.
template <int ROWS_W>
__device__ __forceinline__ void tile_channel_sums(uint32_t tile_addr, int q, int lane, int wlim, int hlim, float& s1a,
                                                  float& s1b, float& s2a, float& s2b) {
  const uint32_t chunk = (uint32_t)lane >> 2, within = ((uint32_t)lane & 3u) * 4u;
  const uint32_t base = tile_addr + (uint32_t)q * 32u * 128u + within;
  const bool full = (wlim >= ROWS_W) && (hlim >= 128 / ROWS_W);
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 16) {
    uint32_t u[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) u[r] = lds32(base + (uint32_t)(r0 + r) * 128u + ((chunk ^ (uint32_t)((r0 + r) & 7)) << 4));
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int rr = r0 + r;                                   // row within the warp's quarter (32*q is a multiple of ROWS_W)
      const float m = (full || (rr % ROWS_W < wlim && (q * 32 + rr) / ROWS_W < hlim)) ? 1.f : 0.f;
      const float2 f = unpack_bf16x2(u[r]);
      const float fx = f.x * m, fy = f.y * m;
      s1a += fx; s1b += fy;
      s2a = fmaf(fx, fx, s2a); s2b = fmaf(fy, fy, s2b);
    }
  }
}

// Philox4x32-10 (Salmon et al.): counter-based, so forward and backward regenerate the same dropout mask.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ unsigned long long resolve_seed(const SeedRef& r) {
  return r.ctr ? r.seed + (*r.ctr) * 0x9E3779B97F4A7C15ull : r.seed;
}
// Dropout decisions use 16 random bits each (p is quantised to 1/65536): one Philox4x32-10 block serves 8 elements.
__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }
// keep-decision for element `idx` of a dropout call identified by `seed`: uniform16 >= p*65536
__device__ __forceinline__ bool philox_keep(unsigned long long seed, unsigned long long idx, float p) {
  const unsigned long long b = idx >> 3;
  uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), 0u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int k = (int)(idx & 7);
  const uint32_t w = (k >> 1) == 0 ? r.x : (k >> 1) == 1 ? r.y : (k >> 1) == 2 ? r.z : r.w;
  return ((w >> ((k & 1) * 16)) & 0xffffu) >= drop_threshold(p);
}

}  // namespace sivae
#endif  // __CUDACC__
