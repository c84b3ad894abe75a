// pointwise.cu -- HBM-bound kernels around the tensor-core convolutions (all NDHWC bf16, 16-byte vectors):
//   * BatchNorm3d (train) statistics / coefficients           models/models.py:18,22,56,60,93,119
//   * fused BN-apply + (Leaky)ReLU + residual + AvgPool3d(2) / Upsample(2) + Dropout, forward and backward
//                                                              models/models.py:15-22,39-41,53-60,94-95,121-122
// One thread owns a fixed group of 8 channels (one 16-byte vector) and walks voxels, so per-channel
// reductions need no atomics: registers -> shared-memory tree -> per-block partials -> fp64 finalize.
#include <cstdlib>

#include "sivae_common.cuh"

#include <cooperative_groups.h>

namespace sivae {

static constexpr int kBnThreads = 256;
static constexpr int kBnMaxBlocks = 148 * 8;

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ldf8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// keep-scale of dropout for the 8 channels of voxel-linear element base `e0` (= voxel*C + c0)
__device__ __forceinline__ void drop8(const uint8_t* mask, float p, unsigned long long seed, long long e0,
                                      float (&ks)[8]) {
  if (mask == nullptr && p <= 0.f) {
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = 1.f;
    return;
  }
  const float inv = 1.f / (1.f - p);
  if (mask != nullptr) {
    const uint2 m = *reinterpret_cast<const uint2*>(mask + e0);
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = (((k < 4 ? m.x : m.y) >> ((k & 3) * 8)) & 0xffu) ? inv : 0.f;
  } else {
    // e0 is a multiple of 8: one Philox block (8 x 16 random bits) covers the 8 elements, same stream as philox_keep
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long b = (unsigned long long)e0 >> 3;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), 0u, 0u), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const uint32_t thr = drop_threshold(p);
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = (((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu) >= thr) ? inv : 0.f;
  }
}

// Keep-bit store (dropout of the encoder stem, models/models.py:95): the forward pass draws the Philox decisions of the 8
// elements [8i, 8i+8) ONCE and stores them as byte i (bit k = element 8i+k kept); both backward passes read that byte
// instead of re-running Philox4x32-10 (which, not HBM, bounded them: 2.3 vs 3.9 TB/s).  1 bit per element = 1/16 of the
// bf16 tensor.  Same decisions as drop8 / philox_keep for the same (seed, element).
__device__ __forceinline__ uint32_t philox_keep_byte(float p, unsigned long long seed, unsigned long long item) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint4 r = philox4x32_10(make_uint4((uint32_t)item, (uint32_t)(item >> 32), 0u, 0u), key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const uint32_t thr = drop_threshold(p);
  uint32_t byte = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) byte |= ((((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu) >= thr) ? 1u : 0u) << k;
  return byte;
}
__device__ __forceinline__ void keep_scale_from_byte(uint32_t byte, float p, float (&ks)[8]) {
  const float inv = 1.f / (1.f - p);
#pragma unroll
  for (int k = 0; k < 8; ++k) ks[k] = ((byte >> k) & 1u) ? inv : 0.f;
}

// Block-level reduction of per-thread (s1[8], s2[8]) for threads sharing a channel chunk; result to
// partial[(blockIdx.x*2 + {0,1})*C + c].  Two passes over one 9 KB staging array (first the s1 sums, then the s2 sums;
// the additions and their order are those of a single 17 KB pass): a reduce block must fit into the ~16 KB of shared
// memory a persistent convolution CTA (209 KB) leaves on its SM, otherwise the BatchNorm-backward of one pass cannot
// run under the convolutions of the pass issued on the other stream (tools/overlap_probe.py: 0.6 % -> see profiles/).
__device__ __forceinline__ void block_reduce_channels(const float (&s1)[8], const float (&s2)[8], int C, int cpc,
                                                      float* __restrict__ partial) {
  __shared__ float sh[kBnThreads][9];
  const int lanes_v = kBnThreads / cpc;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (half) __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[threadIdx.x][k] = half ? s2[k] : s1[k];
    __syncthreads();
    for (int j = threadIdx.x; j < C; j += kBnThreads) {
      const int chunk = j >> 3, k = j & 7;
      float a = 0.f;
      for (int lv = 0; lv < lanes_v; ++lv) a += sh[lv * cpc + chunk][k];
      partial[((long long)blockIdx.x * 2 + half) * C + j] = a;
    }
  }
}

// ---- BN statistics ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long nvox,
                                                              int C, float* __restrict__ partial) {
  const int cpc = C >> 3;
  const int chunk = threadIdx.x % cpc;
  const int lanes_v = kBnThreads / cpc;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * lanes_v;
  for (long long v0 = (long long)blockIdx.x * lanes_v + threadIdx.x / cpc; v0 < nvox; v0 += stride * 4) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v0 + u * stride < nvox) raw[u] = *reinterpret_cast<const uint4*>(y + (v0 + u * stride) * C + chunk * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v0 + u * stride < nvox) {
        const float2 a = unpack_bf16x2(raw[u].x), b = unpack_bf16x2(raw[u].y), c = unpack_bf16x2(raw[u].z),
                     d = unpack_bf16x2(raw[u].w);
        const float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s1[k] += f[k];
          s2[k] += f[k] * f[k];
        }
      }
  }
  block_reduce_channels(s1, s2, C, cpc, partial);
}

// Finalize kernels: one block of 1024 threads per 8 consecutive channels (one 32-byte sector of a partial row): 128
// row-lanes x 8 channels stride over the per-block partials (coalesced sectors, <= 5 dependent loads for 592 blocks
// instead of 19 with a warp per channel), fp64 accumulation, fixed-order reduction -> deterministic.
static constexpr int kFinThreads = 1024;
__device__ __forceinline__ bool block_channel_sums8(const float* __restrict__ partial, int nblocks, int C, double& a,
                                                    double& b) {
  __shared__ double sh[kFinThreads / 32][8][2];
  const int ch = threadIdx.x & 7, r = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + ch;
  double sa = 0.0, sb = 0.0;
  if (c < C)
    for (int i = r; i < nblocks; i += kFinThreads / 8) {
      sa += (double)partial[((long long)i * 2 + 0) * C + c];
      sb += (double)partial[((long long)i * 2 + 1) * C + c];
    }
  sa += __shfl_xor_sync(0xffffffffu, sa, 8);
  sb += __shfl_xor_sync(0xffffffffu, sb, 8);
  sa += __shfl_xor_sync(0xffffffffu, sa, 16);
  sb += __shfl_xor_sync(0xffffffffu, sb, 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
    sh[warp][lane][0] = sa;
    sh[warp][lane][1] = sb;
  }
  __syncthreads();
  if (threadIdx.x >= 8) return false;
  a = 0.0;
  b = 0.0;
#pragma unroll 8
  for (int w = 0; w < kFinThreads / 32; ++w) {
    a += sh[w][threadIdx.x][0];
    b += sh[w][threadIdx.x][1];
  }
  return blockIdx.x * 8 + (int)threadIdx.x < C;
}

__global__ void __launch_bounds__(kFinThreads)
bn_stats_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, long long nvox,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                         float* running_var, long long* nbt, float momentum, float eps, float* mean_o,
                         float* invstd_o, float* scale_o, float* shift_o) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
  double a, b;
  if (!block_channel_sums8(partial, nblocks, C, a, b)) return;
  const int c = blockIdx.x * 8 + threadIdx.x;
  const double n = (double)nvox;
  const double mean = a / n;
  double var = b / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  mean_o[c] = (float)mean;
  invstd_o[c] = invstd;
  scale_o[c] = g * invstd;
  shift_o[c] = bt - (float)mean * g * invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unb = nvox > 1 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// ---- fused forward ------------------------------------------------------------------------------
// activated value of one input-voxel channel chunk
__device__ __forceinline__ void act_chunk(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ res,
                                          const float (&sc)[8], const float (&sh)[8], long long e0, float slope,
                                          const uint8_t* mask, float p, unsigned long long seed, float (&a)[8]) {
  float f[8], ks[8];
  ld8(y + e0, f);
  if (res != nullptr) {
    float r[8];
    ld8(res + e0, r);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]) + r[k];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
  }
  drop8(mask, p, seed, e0, ks);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = (f[k] > 0.f ? f[k] : slope * f[k]) * ks[k];
}

// The most frequent case on its own: no resampling, no residual, no dropout (first BatchNorm of every residual block).
// One tensor in, one out -- the access pattern of a plain copy -- so the bytes in flight per SM are what sets the
// bandwidth: 8 independent 16-byte loads per thread (the general kernel below issues 4 per tensor and carries the
// registers of the residual / dropout paths: 5.0 TB/s where a copy reaches 6.5).  Same arithmetic, bitwise equal results.
__global__ void __launch_bounds__(256, 2) bn_act_plain_fwd_kernel(const __nv_bfloat16* __restrict__ y,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift,
                                                                  __nv_bfloat16* __restrict__ out, unsigned nitems, int C,
                                                                  float slope) {
  constexpr int U = 8;
  const int chunk = (int)(threadIdx.x % (C >> 3));      // 256 threads and C/8 | 256: fixed across the grid-stride loop
  float sc[8], sh[8];
  ldf8(scale + chunk * 8, sc);
  ldf8(shift + chunk * 8, sh);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nitems; i0 += stride * U) {
    uint4 ry[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < nitems) ry[u] = *reinterpret_cast<const uint4*>(y + (size_t)i * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < nitems) {
        float f[8], a[8];
        unpack8(ry[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          f[k] = fmaf(f[k], sc[k], sh[k]);
          a[k] = f[k] > 0.f ? f[k] : slope * f[k];
        }
        st8(out + (size_t)i * 8, a);
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ y,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift,
                                                         const __nv_bfloat16* __restrict__ res,
                                                         __nv_bfloat16* __restrict__ out, int N, int D, int H, int W,
                                                         int C, float slope, const uint8_t* __restrict__ mask, float p,
                                                         const SeedRef sref, uint8_t* __restrict__ keep_bits) {
  const unsigned long long seed = resolve_seed(sref);
  const int cpc = C >> 3;
  const long long nvox_in = (long long)N * D * H * W;
  const long long items = (MODE == SIVAE_RESAMPLE_AVGPOOL2 ? nvox_in / 8 : nvox_in) * cpc;
  // 256 threads and C/8 | 256: a thread's channel chunk never changes across the grid-stride loop
  const int chunk = (int)(threadIdx.x % cpc);
  float sc[8], sh[8];
  ldf8(scale + chunk * 8, sc);
  ldf8(shift + chunk * 8, sh);
  if constexpr (MODE == SIVAE_RESAMPLE_NONE) {
    // streaming case: issue the loads of 4 voxels before consuming them
    // item i covers elements [8i, 8i+8): (i / cpc) * C + (i % cpc) * 8 == 8 * i, no division needed; 32-bit indices
    const unsigned stride = gridDim.x * blockDim.x, nitems = (unsigned)items;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nitems; i0 += stride * 4) {
      uint4 ry[4], rr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned i = i0 + u * stride;
        if (i < nitems) {
          const size_t e0 = (size_t)i * 8;
          ry[u] = *reinterpret_cast<const uint4*>(y + e0);
          if (res != nullptr) rr[u] = *reinterpret_cast<const uint4*>(res + e0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned i = i0 + u * stride;
        if (i < nitems) {
          const long long e0 = (long long)i * 8;
          float f[8], ks[8], a[8];
          unpack8(ry[u], f);
          if (res != nullptr) {
            float r[8];
            unpack8(rr[u], r);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]) + r[k];
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
          }
          if (keep_bits != nullptr) {
            const uint32_t byte = philox_keep_byte(p, seed, i);
            keep_bits[i] = (uint8_t)byte;
            keep_scale_from_byte(byte, p, ks);
          } else {
            drop8(mask, p, seed, e0, ks);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] = (f[k] > 0.f ? f[k] : slope * f[k]) * ks[k];
          st8(out + e0, a);
        }
      }
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / cpc;
    if (MODE == SIVAE_RESAMPLE_NONE) {
      float a[8];
      act_chunk(y, res, sc, sh, v * C + chunk * 8, slope, mask, p, seed, a);
      st8(out + v * C + chunk * 8, a);
    } else if (MODE == SIVAE_RESAMPLE_AVGPOOL2) {
      const unsigned Wo = W / 2, Ho = H / 2, Do = D / 2, vv = (unsigned)v;   // 32-bit: host checks item count < 2^31
      const unsigned wo = vv % Wo, t1 = vv / Wo;
      const unsigned ho = t1 % Ho, t2 = t1 / Ho;
      const unsigned dd = t2 % Do;
      const long long n = t2 / Do;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (res == nullptr && mask == nullptr && p == 0.f) {
        // common case (BuildingBlock, models/models.py:18-20): all eight 16-byte loads in flight before the first use
        // (one dependent load at a time ran this pass at 2.4 TB/s)
        uint4 raw[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const long long vi = ((n * D + (2 * dd + (q >> 2))) * H + (2 * ho + ((q >> 1) & 1))) * W + (2 * wo + (q & 1));
          raw[q] = *reinterpret_cast<const uint4*>(y + vi * C + chunk * 8);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
          unpack8(raw[q], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float t = fmaf(f[k], sc[k], sh[k]);
            acc[k] += t > 0.f ? t : slope * t;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const long long vi = ((n * D + (2 * dd + (q >> 2))) * H + (2 * ho + ((q >> 1) & 1))) * W + (2 * wo + (q & 1));
          float a[8];
          act_chunk(y, res, sc, sh, vi * C + chunk * 8, slope, mask, p, seed, a);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += a[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= 0.125f;
      st8(out + v * C + chunk * 8, acc);
    } else {
      const int w = (int)(v % W);
      const int h = (int)((v / W) % H);
      const int d = (int)((v / ((long long)W * H)) % D);
      const long long n = v / ((long long)W * H * D);
      float a[8];
      act_chunk(y, res, sc, sh, v * C + chunk * 8, slope, mask, p, seed, a);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const long long vo =
            ((n * (2 * D) + (2 * d + (q >> 2))) * (2 * H) + (2 * h + ((q >> 1) & 1))) * (2 * W) + (2 * w + (q & 1));
        st8(out + vo * C + chunk * 8, a);
      }
    }
  }
}

// ---- fused backward -----------------------------------------------------------------------------
// dt (gradient at the BN output, after activation/dropout/resample backward) and xhat for one input voxel chunk
template <int MODE>
__device__ __forceinline__ void bwd_chunk(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                          const __nv_bfloat16* __restrict__ res, const float (&mean)[8],
                                          const float (&invstd)[8], const float (&gam)[8], const float (&bet)[8],
                                          long long v, int chunk, int N, int D, int H, int W, int C, float slope,
                                          const uint8_t* mask, float p, unsigned long long seed, float (&dt)[8],
                                          float (&xh)[8]) {
  const long long e0 = v * C + chunk * 8;
  float f[8], gp[8], ks[8];
  ld8(y + e0, f);
  if (MODE == SIVAE_RESAMPLE_NONE) {
    ld8(g + e0, gp);
  } else {
    const int w = (int)(v % W);
    const int h = (int)((v / W) % H);
    const int d = (int)((v / ((long long)W * H)) % D);
    const long long n = v / ((long long)W * H * D);
    if (MODE == SIVAE_RESAMPLE_AVGPOOL2) {
      const long long vo = ((n * (D / 2) + d / 2) * (H / 2) + h / 2) * (W / 2) + w / 2;
      ld8(g + vo * C + chunk * 8, gp);
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] *= 0.125f;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const long long vo =
            ((n * (2 * D) + (2 * d + (q >> 2))) * (2 * H) + (2 * h + ((q >> 1) & 1))) * (2 * W) + (2 * w + (q & 1));
        float t[8];
        ld8(g + vo * C + chunk * 8, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) gp[k] += t[k];
      }
    }
  }
  drop8(mask, p, seed, e0, ks);
  float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (res != nullptr) ld8(res + e0, r);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    xh[k] = (f[k] - mean[k]) * invstd[k];
    const float t = fmaf(xh[k], gam[k], bet[k]) + r[k];
    dt[k] = gp[k] * ks[k] * (t > 0.f ? 1.f : slope);
  }
}

// Split load / compute form of bwd_chunk for the streaming modes (NONE, AVGPOOL2): a thread first issues the 16-byte
// loads of kBwdUnroll voxels, then consumes them, so ~4x more bytes are in flight per thread (HBM latency hiding).
static constexpr int kBwdUnroll = 2;
struct BwdRaw {
  uint4 y, g, r;
  uint2 m;
};

template <int MODE>
__device__ __forceinline__ void bwd_load(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                         const __nv_bfloat16* __restrict__ res, const uint8_t* __restrict__ mask,
                                         long long v, int chunk, int D, int H, int W, int C, BwdRaw& raw) {
  const long long e0 = v * C + chunk * 8;
  raw.y = *reinterpret_cast<const uint4*>(y + e0);
  if (MODE == SIVAE_RESAMPLE_NONE) {
    raw.g = *reinterpret_cast<const uint4*>(g + e0);
  } else {
    const unsigned vv = (unsigned)v;               // host guarantees nvox < 2^31
    const unsigned w = vv % (unsigned)W, t1 = vv / (unsigned)W;
    const unsigned h = t1 % (unsigned)H, t2 = t1 / (unsigned)H;
    const unsigned d = t2 % (unsigned)D, n = t2 / (unsigned)D;
    const size_t vo = (((size_t)n * (D / 2) + d / 2) * (H / 2) + h / 2) * (W / 2) + w / 2;
    raw.g = *reinterpret_cast<const uint4*>(g + vo * C + chunk * 8);
  }
  if (res != nullptr) raw.r = *reinterpret_cast<const uint4*>(res + e0);
  if (mask != nullptr) raw.m = *reinterpret_cast<const uint2*>(mask + e0);
}

template <int MODE>
__device__ __forceinline__ void bwd_compute(const BwdRaw& raw, bool has_res, const uint8_t* mask, float p,
                                            unsigned long long seed, long long e0, const float (&mean)[8],
                                            const float (&invstd)[8], const float (&gam)[8], const float (&bet)[8],
                                            float slope, float (&dt)[8], float (&xh)[8]) {
  float f[8], gp[8], r[8], ks[8];
  unpack8(raw.y, f);
  unpack8(raw.g, gp);
  if (has_res) {
    unpack8(raw.r, r);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = 0.f;
  }
  if (mask != nullptr) {
    const float inv = 1.f / (1.f - p);
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = (((k < 4 ? raw.m.x : raw.m.y) >> ((k & 3) * 8)) & 0xffu) ? inv : 0.f;
  } else {
    drop8(nullptr, p, seed, e0, ks);
  }
  const float gs = MODE == SIVAE_RESAMPLE_AVGPOOL2 ? 0.125f : 1.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    xh[k] = (f[k] - mean[k]) * invstd[k];
    const float t = fmaf(xh[k], gam[k], bet[k]) + r[k];
    dt[k] = gp[k] * gs * ks[k] * (t > 0.f ? 1.f : slope);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kBnThreads, 2)
bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                         const __nv_bfloat16* __restrict__ res, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int N, int D, int H, int W, int C, float slope,
                         const uint8_t* __restrict__ mask, float p, const SeedRef sref,
                         float* __restrict__ partial) {
  const unsigned long long seed = resolve_seed(sref);
  const int cpc = C >> 3;
  const int chunk = threadIdx.x % cpc;
  const int lanes_v = kBnThreads / cpc;
  const long long nvox = (long long)N * D * H * W;
  float mu[8], is[8], ga[8], be[8];
  ldf8(mean + chunk * 8, mu); ldf8(invstd + chunk * 8, is); ldf8(gamma + chunk * 8, ga); ldf8(beta + chunk * 8, be);
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * lanes_v;
  if constexpr (MODE == SIVAE_RESAMPLE_UPSAMPLE2) {
    for (long long v = (long long)blockIdx.x * lanes_v + threadIdx.x / cpc; v < nvox; v += stride) {
      float dt[8], xh[8];
      bwd_chunk<MODE>(g, y, res, mu, is, ga, be, v, chunk, N, D, H, W, C, slope, mask, p, seed, dt, xh);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1[k] += dt[k];
        s2[k] += dt[k] * xh[k];
      }
    }
  } else if (MODE == SIVAE_RESAMPLE_AVGPOOL2 && res == nullptr && mask == nullptr && p <= 0.f) {
    // walk POOLED voxels: one index decomposition and one g load serve the 8 children, whose y loads are all in flight
    const unsigned Wo = W / 2, Ho = H / 2, Do = D / 2;
    const unsigned nvo = (unsigned)(nvox / 8);
    for (unsigned vo = blockIdx.x * lanes_v + threadIdx.x / cpc; vo < nvo; vo += (unsigned)stride) {
      const unsigned wo = vo % Wo, t1 = vo / Wo;
      const unsigned ho = t1 % Ho, t2 = t1 / Ho;
      const unsigned dd = t2 % Do, n = t2 / Do;
      const size_t v000 = (((size_t)n * D + 2 * dd) * H + 2 * ho) * W + 2 * wo;
      uint4 ry[8];
      float gp[8];
      ld8(g + (size_t)vo * C + chunk * 8, gp);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        ry[q] = *reinterpret_cast<const uint4*>(y + (v000 + ((size_t)(q >> 2) * H + ((q >> 1) & 1)) * W + (q & 1)) * C + chunk * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float f[8];
        unpack8(ry[q], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (f[k] - mu[k]) * is[k];
          const float t = fmaf(xh, ga[k], be[k]);
          const float dt = gp[k] * 0.125f * (t > 0.f ? 1.f : slope);
          s1[k] += dt;
          s2[k] += dt * xh;
        }
      }
    }
  } else {
    for (long long v0 = (long long)blockIdx.x * lanes_v + threadIdx.x / cpc; v0 < nvox; v0 += stride * kBwdUnroll) {
      BwdRaw raw[kBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        const long long v = v0 + u * stride;
        if (v < nvox) bwd_load<MODE>(g, y, res, mask, v, chunk, D, H, W, C, raw[u]);
      }
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        const long long v = v0 + u * stride;
        if (v < nvox) {
          float dt[8], xh[8];
          bwd_compute<MODE>(raw[u], res != nullptr, mask, p, seed, v * C + chunk * 8, mu, is, ga, be, slope, dt, xh);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            s1[k] += dt[k];
            s2[k] += dt[k] * xh[k];
          }
        }
      }
    }
  }
  block_reduce_channels(s1, s2, C, cpc, partial);
}

// ---- lean kernels for the common case: no resample, no residual, no dropout --------------------------------------
// Folded per-channel constants keep the register count low enough for 4 voxels (128 B) in flight per thread at
// >= 2 CTAs/SM, which is what the HBM latency-bandwidth product needs (~64 KB in flight per SM).
//   t = y*S + T (S = gamma*invstd, T = beta - mean*S);  xhat = y*invstd + M2 (M2 = -mean*invstd);  dt = g * act'(t)
// DROP: 0 = no dropout, 1 = Philox regenerated in the kernel, 2 = keep-bit store written by the forward pass
template <int DROP>
__global__ void __launch_bounds__(kBnThreads, 2)
bn_bwd_reduce_plain_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                           const float* __restrict__ mean, const float* __restrict__ invstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, long long nvox, int C,
                           float slope, float p, const SeedRef sref, const uint8_t* __restrict__ keep_bits,
                           float* __restrict__ partial) {
  const unsigned long long seed = DROP == 1 ? resolve_seed(sref) : 0ull;
  const int cpc = C >> 3;
  const int chunk = threadIdx.x % cpc;
  const int lanes_v = kBnThreads / cpc;
  float S[8], T[8], IS[8], M2[8];
  {
    float mu[8], ga[8], be[8];
    ldf8(mean + chunk * 8, mu); ldf8(invstd + chunk * 8, IS); ldf8(gamma + chunk * 8, ga); ldf8(beta + chunk * 8, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      S[k] = ga[k] * IS[k];
      T[k] = be[k] - mu[k] * S[k];
      M2[k] = -mu[k] * IS[k];
    }
  }
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const unsigned stride = gridDim.x * lanes_v, nv = (unsigned)nvox;
  for (unsigned v0 = blockIdx.x * lanes_v + threadIdx.x / cpc; v0 < nv; v0 += stride * 4) {
    uint4 ry[4], rg[4];
    uint32_t kb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned v = v0 + u * stride;
      if (v < nv) {
        const size_t e0 = (size_t)v * C + chunk * 8;
        ry[u] = *reinterpret_cast<const uint4*>(y + e0);
        rg[u] = *reinterpret_cast<const uint4*>(g + e0);
        if (DROP == 2) kb[u] = keep_bits[(size_t)v * cpc + chunk];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v0 + u * stride < nv) {
        float f[8], gp[8];
        unpack8(ry[u], f);
        unpack8(rg[u], gp);
        if (DROP != 0) {
          float ks[8];
          if (DROP == 2) keep_scale_from_byte(kb[u], p, ks);
          else drop8(nullptr, p, seed, (long long)((size_t)(v0 + u * stride) * C + chunk * 8), ks);
#pragma unroll
          for (int k = 0; k < 8; ++k) gp[k] *= ks[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float t = fmaf(f[k], S[k], T[k]);
          const float dt = t > 0.f ? gp[k] : gp[k] * slope;
          s1[k] += dt;
          s2[k] = fmaf(dt, fmaf(f[k], IS[k], M2[k]), s2[k]);
        }
      }
    }
  }
  block_reduce_channels(s1, s2, C, cpc, partial);
}

// dconv = A*dt + y*U + V with A = gamma*invstd, U = -A*invstd*c2, V = A*(mean*invstd*c2 - c1)
template <int DROP>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_plain_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ coef, __nv_bfloat16* __restrict__ dconv, long long items, int C,
                          float slope, float p, const SeedRef sref, const uint8_t* __restrict__ keep_bits) {
  const unsigned long long seed = DROP == 1 ? resolve_seed(sref) : 0ull;
  const int cpc = C >> 3;
  const int chunk = (int)(threadIdx.x % cpc);
  float S[8], T[8], A[8], U[8], V[8];
  {
    float mu[8], is[8], ga[8], be[8], c1[8], c2[8];
    ldf8(mean + chunk * 8, mu); ldf8(invstd + chunk * 8, is); ldf8(gamma + chunk * 8, ga); ldf8(beta + chunk * 8, be);
    ldf8(coef + chunk * 8, c1); ldf8(coef + C + chunk * 8, c2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      S[k] = ga[k] * is[k];
      T[k] = be[k] - mu[k] * S[k];
      A[k] = S[k];
      U[k] = -A[k] * is[k] * c2[k];
      V[k] = A[k] * (mu[k] * is[k] * c2[k] - c1[k]);
    }
  }
  const unsigned stride = gridDim.x * blockDim.x, nitems = (unsigned)items;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nitems; i0 += stride * 4) {
    uint4 ry[4], rg[4];
    uint32_t kb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < nitems) {
        ry[u] = *reinterpret_cast<const uint4*>(y + (size_t)i * 8);
        rg[u] = *reinterpret_cast<const uint4*>(g + (size_t)i * 8);
        if (DROP == 2) kb[u] = keep_bits[i];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < nitems) {
        float f[8], gp[8], o[8];
        unpack8(ry[u], f);
        unpack8(rg[u], gp);
        if (DROP != 0) {
          float ks[8];
          if (DROP == 2) keep_scale_from_byte(kb[u], p, ks);
          else drop8(nullptr, p, seed, (long long)((size_t)i * 8), ks);
#pragma unroll
          for (int k = 0; k < 8; ++k) gp[k] *= ks[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float t = fmaf(f[k], S[k], T[k]);
          const float dt = t > 0.f ? gp[k] : gp[k] * slope;
          o[k] = fmaf(A[k], dt, fmaf(f[k], U[k], V[k]));
        }
        st8(dconv + (size_t)i * 8, o);
      }
    }
  }
}

// coef[0][c] = sum(dt)/n, coef[1][c] = sum(dt*xhat)/n ; dgamma = sum(dt*xhat), dbeta = sum(dt)
__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, long long nvox,
                       float* __restrict__ coef, float* dgamma, float* dbeta) {
  double a, b;
  if (!block_channel_sums8(partial, nblocks, C, a, b)) return;
  const int c = blockIdx.x * 8 + threadIdx.x;
  coef[c] = (float)(a / (double)nvox);
  coef[C + c] = (float)(b / (double)nvox);
  if (dbeta) dbeta[c] = (float)a;
  if (dgamma) dgamma[c] = (float)b;
}

template <int MODE>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                        const __nv_bfloat16* __restrict__ res, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ coef,
                        __nv_bfloat16* __restrict__ dconv, __nv_bfloat16* __restrict__ dres, int N, int D, int H, int W,
                        int C, float slope, const uint8_t* __restrict__ mask, float p, const SeedRef sref) {
  const unsigned long long seed = resolve_seed(sref);
  const int cpc = C >> 3;
  const long long items = (long long)N * D * H * W * cpc;
  const int chunk = (int)(threadIdx.x % cpc);   // loop-invariant (C/8 | 256)
  float mu[8], is[8], ga[8], be[8], c1[8], c2[8];
  ldf8(mean + chunk * 8, mu); ldf8(invstd + chunk * 8, is); ldf8(gamma + chunk * 8, ga); ldf8(beta + chunk * 8, be);
  ldf8(coef + chunk * 8, c1); ldf8(coef + C + chunk * 8, c2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if constexpr (MODE == SIVAE_RESAMPLE_UPSAMPLE2) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += stride) {
      const long long v = i / cpc;
      float dt[8], xh[8], o[8];
      bwd_chunk<MODE>(g, y, res, mu, is, ga, be, v, chunk, N, D, H, W, C, slope, mask, p, seed, dt, xh);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = ga[k] * is[k] * (dt[k] - c1[k] - xh[k] * c2[k]);
      st8(dconv + v * C + chunk * 8, o);
      if (dres != nullptr) st8(dres + v * C + chunk * 8, dt);
    }
  } else if (MODE == SIVAE_RESAMPLE_AVGPOOL2 && res == nullptr && mask == nullptr && p <= 0.f && dres == nullptr) {
    const unsigned Wo = W / 2, Ho = H / 2, Do = D / 2, ucpc = (unsigned)cpc;
    const unsigned npooled = (unsigned)(items / 8);
    for (unsigned io = blockIdx.x * blockDim.x + threadIdx.x; io < npooled; io += (unsigned)stride) {
      const unsigned vo = io / ucpc;
      const unsigned wo = vo % Wo, t1 = vo / Wo;
      const unsigned ho = t1 % Ho, t2 = t1 / Ho;
      const unsigned dd = t2 % Do, n = t2 / Do;
      const size_t v000 = (((size_t)n * D + 2 * dd) * H + 2 * ho) * W + 2 * wo;
      uint4 ry[8];
      float gp[8];
      ld8(g + (size_t)vo * C + chunk * 8, gp);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        ry[q] = *reinterpret_cast<const uint4*>(y + (v000 + ((size_t)(q >> 2) * H + ((q >> 1) & 1)) * W + (q & 1)) * C + chunk * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float f[8], o[8];
        unpack8(ry[q], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (f[k] - mu[k]) * is[k];
          const float t = fmaf(xh, ga[k], be[k]);
          const float dt = gp[k] * 0.125f * (t > 0.f ? 1.f : slope);
          o[k] = ga[k] * is[k] * (dt - c1[k] - xh * c2[k]);
        }
        st8(dconv + (v000 + ((size_t)(q >> 2) * H + ((q >> 1) & 1)) * W + (q & 1)) * C + chunk * 8, o);
      }
    }
  } else {
    const unsigned ucpc = (unsigned)cpc, ustride = (unsigned)stride, nitems = (unsigned)items;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nitems; i0 += ustride * kBwdUnroll) {
      BwdRaw raw[kBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        const unsigned i = i0 + u * ustride;
        if (i < nitems) bwd_load<MODE>(g, y, res, mask, (long long)(i / ucpc), chunk, D, H, W, C, raw[u]);
      }
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        const unsigned i = i0 + u * ustride;
        if (i < nitems) {
          const long long v = (long long)(i / ucpc);
          float dt[8], xh[8], o[8];
          bwd_compute<MODE>(raw[u], res != nullptr, mask, p, seed, v * C + chunk * 8, mu, is, ga, be, slope, dt, xh);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = ga[k] * is[k] * (dt[k] - c1[k] - xh[k] * c2[k]);
          st8(dconv + v * C + chunk * 8, o);
          if (dres != nullptr) st8(dres + v * C + chunk * 8, dt);
        }
      }
    }
  }
}

// ---- layout helpers -----------------------------------------------------------------------------
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int C,
                                      long long vox) {
  const long long total = (long long)N * C * vox;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long v = (i / C) % vox;
    const long long n = i / ((long long)C * vox);
    dst[i] = __float2bfloat16_rn(src[(n * C + c) * vox + v]);
  }
}
__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int N, int C,
                                      long long vox) {
  const long long total = (long long)N * C * vox;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % vox;
    const int c = (int)((i / vox) % C);
    const long long n = i / ((long long)C * vox);
    dst[i] = __bfloat162float(src[(n * vox + v) * C + c]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  static const long long cap = getenv("SIVAE_PW_BLOCKS") ? atoll(getenv("SIVAE_PW_BLOCKS")) : 148ll * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static bool channels_ok(int C) { return C >= 8 && (C % 8) == 0 && (kBnThreads % (C / 8)) == 0; }
// the fused forward / backward-apply kernels keep a thread's channel chunk fixed across their grid-stride loop

size_t bn_workspace_bytes(int C) { return ((size_t)kBnMaxBlocks * 2 * C + 2 * (size_t)C) * sizeof(float); }

static int reduce_blocks(long long nvox, int C) {
  // at least ~8 voxels per thread: small tensors get few blocks (cheap finalize), large ones 4 CTAs per SM
  const int lanes_v = kBnThreads / (C / 8);
  long long b = (nvox + lanes_v * 8 - 1) / (lanes_v * 8);
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  return (int)b;
}

int bn_train_coeffs(const void* y, long long nvox, int C, const float* gamma, const float* beta, float* rm, float* rv,
                    long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(channels_ok(C), "bn_train_coeffs: unsupported channel count %d", C);
  SIVAE_CHECK(nvox > 0, "bn_train_coeffs: empty tensor");
  SIVAE_CHECK(ws && ws_bytes >= bn_workspace_bytes(C), "bn_train_coeffs: workspace too small");
  const int blocks = reduce_blocks(nvox, C);
  bn_stats_kernel<<<blocks, kBnThreads, 0, st>>>((const __nv_bfloat16*)y, nvox, C, (float*)ws);
  SIVAE_LAUNCH_OK("bn_stats_kernel");
  bn_stats_finalize_kernel<<<cdiv(C, 8), kFinThreads, 0, st>>>((const float*)ws, blocks, C, nvox, gamma, beta, rm, rv, nbt,
                                                         momentum, eps, mean, invstd, scale, shift);
  SIVAE_LAUNCH_OK("bn_stats_finalize_kernel");
  return 0;
}

// =================================================================================================
// EXPERIMENT (opt-in, see small_path_ok).  Small tensors (the 256-channel layers at the latent resolution: 2.4 M
// elements, 4.9 MB): the three launches statistics -> finalize -> apply (and reduce -> finalize -> apply in backward)
// cost ~20 us, mostly launch latency and tails.  One thread-block CLUSTER of 16 CTAs does all three phases in a single
// launch: every CTA reduces its slice of the
// voxels, the per-channel partials are exchanged through distributed shared memory (cluster.map_shared_rank) between two
// cluster barriers, and the second pass over the slice hits L2.  Deterministic (fixed rank order, fp64 finalize).
// =================================================================================================
namespace cg = cooperative_groups;
static constexpr int kSmallCluster = 16, kSmallThreads = 512, kSmallMaxC = 256, kSmallUnroll = 4;
static constexpr long long kSmallMaxElems = 3ll << 20;     // 3 Mi elements (6 MB of bf16)

struct SmallSlice { long long v0, v1; };
__device__ __forceinline__ SmallSlice small_slice(long long nvox, unsigned rank) {
  const long long per = (nvox + kSmallCluster - 1) / kSmallCluster;
  SmallSlice s;
  s.v0 = per * rank;
  s.v1 = s.v0 + per < nvox ? s.v0 + per : nvox;
  if (s.v0 > nvox) s.v0 = nvox;
  return s;
}

// block-level reduction of per-thread (a[8], b[8]) over the threads that share a channel chunk -> part[0][c], part[1][c]
__device__ __forceinline__ void small_block_reduce(const float (&a)[8], const float (&b)[8], int C, int cpc,
                                                   float (*red)[17], float (*part)[kSmallMaxC]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[threadIdx.x][k] = a[k];
    red[threadIdx.x][8 + k] = b[k];
  }
  __syncthreads();
  const int lanes_v = kSmallThreads / cpc;
  for (int j = threadIdx.x; j < C; j += kSmallThreads) {
    const int chunk = j >> 3, k = j & 7;
    float x = 0.f, y = 0.f;
    for (int lv = 0; lv < lanes_v; ++lv) {
      x += red[lv * cpc + chunk][k];
      y += red[lv * cpc + chunk][8 + k];
    }
    part[0][j] = x;
    part[1][j] = y;
  }
}

__global__ void __launch_bounds__(kSmallThreads)
bn_small_fwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ res,
                    __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                    float* __restrict__ mean_o, float* __restrict__ invstd_o, float* __restrict__ scale_o,
                    float* __restrict__ shift_o, long long nvox, int C, float slope) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  __shared__ float red[kSmallThreads][17];
  __shared__ float part[2][kSmallMaxC];
  __shared__ float sc_s[kSmallMaxC], sh_s[kSmallMaxC];
  const int cpc = C >> 3, chunk = threadIdx.x % cpc, lanes_v = kSmallThreads / cpc;
  const SmallSlice sl = small_slice(nvox, rank);
  // ---- phase 1: per-channel sums of this CTA's voxel slice
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long v = sl.v0 + threadIdx.x / cpc; v < sl.v1; v += (long long)lanes_v * kSmallUnroll) {
    uint4 raw[kSmallUnroll];
#pragma unroll
    for (int u = 0; u < kSmallUnroll; ++u)
      if (v + (long long)u * lanes_v < sl.v1) raw[u] = *reinterpret_cast<const uint4*>(y + (v + (long long)u * lanes_v) * C + chunk * 8);
#pragma unroll
    for (int u = 0; u < kSmallUnroll; ++u)
      if (v + (long long)u * lanes_v < sl.v1) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s1[k] += f[k];
          s2[k] = fmaf(f[k], f[k], s2[k]);
        }
      }
  }
  small_block_reduce(s1, s2, C, cpc, red, part);
  cluster.sync();
  // ---- phase 2: every CTA reduces the 16 partial rows (fixed rank order) and derives scale / shift
  for (int c = threadIdx.x; c < C; c += kSmallThreads) {
    double a = 0.0, b = 0.0;
    for (unsigned r = 0; r < kSmallCluster; ++r) {
      const float* rp = cluster.map_shared_rank(&part[0][0], r);
      a += (double)rp[c];
      b += (double)rp[kSmallMaxC + c];
    }
    const double n = (double)nvox, mean = a / n;
    double var = b / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    sc_s[c] = g * invstd;
    sh_s[c] = bt - (float)mean * g * invstd;
    if (rank == 0) {
      mean_o[c] = (float)mean;
      invstd_o[c] = invstd;
      scale_o[c] = sc_s[c];
      shift_o[c] = sh_s[c];
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      if (running_var) {
        const double unb = nvox > 1 ? var * n / (n - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
      }
    }
  }
  if (rank == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
  cluster.sync();     // remote reads are finished before any CTA may exit; sc_s / sh_s visible block-wide
  // ---- phase 3: apply (second pass over the slice: L2 hits)
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = sc_s[chunk * 8 + k]; sh[k] = sh_s[chunk * 8 + k]; }
  for (long long v = sl.v0 + threadIdx.x / cpc; v < sl.v1; v += (long long)lanes_v * kSmallUnroll) {
    uint4 ry[kSmallUnroll], rr[kSmallUnroll];
#pragma unroll
    for (int u = 0; u < kSmallUnroll; ++u) {
      const long long vv = v + (long long)u * lanes_v;
      if (vv < sl.v1) {
        ry[u] = *reinterpret_cast<const uint4*>(y + vv * C + chunk * 8);
        if (res != nullptr) rr[u] = *reinterpret_cast<const uint4*>(res + vv * C + chunk * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kSmallUnroll; ++u) {
      const long long vv = v + (long long)u * lanes_v;
      if (vv < sl.v1) {
        float f[8], a[8];
        unpack8(ry[u], f);
        if (res != nullptr) {
          float r[8];
          unpack8(rr[u], r);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]) + r[k];
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = f[k] > 0.f ? f[k] : slope * f[k];
        st8(out + vv * C + chunk * 8, a);
      }
    }
  }
}

__global__ void __launch_bounds__(kSmallThreads)
bn_small_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                    const __nv_bfloat16* __restrict__ res, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    __nv_bfloat16* __restrict__ dconv, __nv_bfloat16* __restrict__ dres, float* dgamma, float* dbeta,
                    long long nvox, int C, float slope) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  __shared__ float red[kSmallThreads][17];
  __shared__ float part[2][kSmallMaxC];
  __shared__ float c1_s[kSmallMaxC], c2_s[kSmallMaxC];
  const int cpc = C >> 3, chunk = threadIdx.x % cpc, lanes_v = kSmallThreads / cpc;
  const SmallSlice sl = small_slice(nvox, rank);
  float mu[8], is[8], ga[8], be[8];
  ldf8(mean + chunk * 8, mu); ldf8(invstd + chunk * 8, is); ldf8(gamma + chunk * 8, ga); ldf8(beta + chunk * 8, be);
  // dt = g * act'(bn(y) (+ res)),  xhat = (y - mean) * invstd
  auto elem = [&](const uint4& ry, const uint4& rg, const uint4& rr, float (&dt)[8], float (&xh)[8]) {
    float f[8], gp[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unpack8(ry, f);
    unpack8(rg, gp);
    if (res != nullptr) unpack8(rr, r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xh[k] = (f[k] - mu[k]) * is[k];
      const float t = fmaf(xh[k], ga[k], be[k]) + r[k];
      dt[k] = gp[k] * (t > 0.f ? 1.f : slope);
    }
  };
  // ---- phase 1: sum(dt), sum(dt * xhat)
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long v = sl.v0 + threadIdx.x / cpc; v < sl.v1; v += (long long)lanes_v * 2) {
    uint4 ry[2], rg[2], rr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vv = v + (long long)u * lanes_v;
      if (vv < sl.v1) {
        ry[u] = *reinterpret_cast<const uint4*>(y + vv * C + chunk * 8);
        rg[u] = *reinterpret_cast<const uint4*>(g + vv * C + chunk * 8);
        if (res != nullptr) rr[u] = *reinterpret_cast<const uint4*>(res + vv * C + chunk * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (v + (long long)u * lanes_v < sl.v1) {
        float dt[8], xh[8];
        elem(ry[u], rg[u], rr[u], dt, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s1[k] += dt[k];
          s2[k] = fmaf(dt[k], xh[k], s2[k]);
        }
      }
  }
  small_block_reduce(s1, s2, C, cpc, red, part);
  cluster.sync();
  for (int c = threadIdx.x; c < C; c += kSmallThreads) {
    double a = 0.0, b = 0.0;
    for (unsigned r = 0; r < kSmallCluster; ++r) {
      const float* rp = cluster.map_shared_rank(&part[0][0], r);
      a += (double)rp[c];
      b += (double)rp[kSmallMaxC + c];
    }
    c1_s[c] = (float)(a / (double)nvox);
    c2_s[c] = (float)(b / (double)nvox);
    if (rank == 0) {
      if (dbeta) dbeta[c] = (float)a;
      if (dgamma) dgamma[c] = (float)b;
    }
  }
  cluster.sync();
  // ---- phase 3: dconv = gamma * invstd * (dt - mean(dt) - xhat * mean(dt * xhat)),  dres = dt
  float c1[8], c2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c1[k] = c1_s[chunk * 8 + k]; c2[k] = c2_s[chunk * 8 + k]; }
  for (long long v = sl.v0 + threadIdx.x / cpc; v < sl.v1; v += (long long)lanes_v * 2) {
    uint4 ry[2], rg[2], rr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vv = v + (long long)u * lanes_v;
      if (vv < sl.v1) {
        ry[u] = *reinterpret_cast<const uint4*>(y + vv * C + chunk * 8);
        rg[u] = *reinterpret_cast<const uint4*>(g + vv * C + chunk * 8);
        if (res != nullptr) rr[u] = *reinterpret_cast<const uint4*>(res + vv * C + chunk * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vv = v + (long long)u * lanes_v;
      if (vv < sl.v1) {
        float dt[8], xh[8], o[8];
        elem(ry[u], rg[u], rr[u], dt, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ga[k] * is[k] * (dt[k] - c1[k] - xh[k] * c2[k]);
        st8(dconv + vv * C + chunk * 8, o);
        if (dres != nullptr) st8(dres + vv * C + chunk * 8, dt);
      }
    }
  }
}

static bool small_path_ok(long long nvox, int C) {
  // OPT-IN (SIVAE_BN_CLUSTER=1): parity-tested, but measured SLOWER in the step (68.5 vs 67.5 ms, A/B on one box): one
  // 16-CTA cluster cannot keep enough loads in flight to beat three short launches that spread over all 148 SMs.
  const bool enabled = getenv("SIVAE_BN_CLUSTER") != nullptr;
  return enabled && C <= kSmallMaxC && C >= 8 && (kSmallThreads % (C / 8)) == 0 && nvox * C <= kSmallMaxElems &&
         nvox >= kSmallCluster;
}

template <class Kernel, class... Args>
static int launch_small_cluster(Kernel kernel, const char* name, cudaStream_t st, Args... args) {
  static bool attr_set[2] = {false, false};
  (void)attr_set;
  if (check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), name)) return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kSmallCluster, 1, 1);
  cfg.blockDim = dim3(kSmallThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSmallCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (check_cuda(cudaLaunchKernelEx(&cfg, kernel, args...), name)) return -1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Train-mode BatchNorm statistics + coefficients + apply (+ residual, (Leaky)ReLU) in one call; no resampling, no
// dropout.  Small tensors take the single-launch cluster kernel, larger ones the statistics / finalize / apply kernels.
int bn_act_fwd(const void* y, const float* scale, const float* shift, const void* res, void* out, int N, int D, int H,
               int W, int C, float slope, int resample, const uint8_t* mask, float p, unsigned long long seed,
               cudaStream_t st);
int bn_train_coeffs(const void* y, long long nvox, int C, const float* gamma, const float* beta, float* rm, float* rv,
                    long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                    void* ws, size_t ws_bytes, cudaStream_t st);
int bn_train_act_fwd(const void* y, const void* res, void* out, int N, int D, int H, int W, int C, const float* gamma,
                     const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps, float slope,
                     float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(channels_ok(C), "bn_train_act_fwd: unsupported channel count %d", C);
  const long long nvox = (long long)N * D * H * W;
  SIVAE_CHECK(nvox > 0, "bn_train_act_fwd: empty tensor");
  if (small_path_ok(nvox, C))
    return launch_small_cluster(bn_small_fwd_kernel, "bn_small_fwd_kernel", st, (const __nv_bfloat16*)y,
                                (const __nv_bfloat16*)res, (__nv_bfloat16*)out, gamma, beta, rm, rv, nbt, momentum, eps,
                                mean, invstd, scale, shift, nvox, C, slope);
  if (int rc = bn_train_coeffs(y, nvox, C, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale, shift, ws,
                               ws_bytes, st))
    return rc;
  return bn_act_fwd(y, scale, shift, res, out, N, D, H, W, C, slope, SIVAE_RESAMPLE_NONE, nullptr, 0.f, 0ull, st);
}

int bn_coeffs_from_partials(const float* partial, int nblocks, long long nvox, int C, const float* gamma,
                            const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                            float* mean, float* invstd, float* scale, float* shift, cudaStream_t st) {
  SIVAE_CHECK(channels_ok(C) && nblocks > 0 && nblocks <= kBnMaxBlocks && nvox > 0, "bn_coeffs_from_partials: bad arguments");
  bn_stats_finalize_kernel<<<cdiv(C, 8), kFinThreads, 0, st>>>(partial, nblocks, C, nvox, gamma, beta, rm, rv, nbt,
                                                                momentum, eps, mean, invstd, scale, shift);
  SIVAE_LAUNCH_OK("bn_stats_finalize_kernel");
  return 0;
}

static int check_resample(const char* who, int D, int H, int W, int resample) {
  SIVAE_CHECK(resample >= 0 && resample <= 2, "%s: bad resample mode %d", who, resample);
  if (resample == SIVAE_RESAMPLE_AVGPOOL2)
    SIVAE_CHECK(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "%s: AvgPool3d(2) needs even extents (%d,%d,%d)", who, D, H, W);
  return 0;
}

int bn_act_fwd(const void* y, const float* scale, const float* shift, const void* res, void* out, int N, int D, int H,
               int W, int C, float slope, int resample, const uint8_t* mask, float p, unsigned long long seed,
               cudaStream_t st) {
  SIVAE_CHECK(channels_ok(C), "bn_act_fwd: unsupported channel count %d", C);
  SIVAE_CHECK(p >= 0.f && p < 1.f, "bn_act_fwd: dropout p=%f out of range", p);
  if (check_resample("bn_act_fwd", D, H, W, resample)) return -2;
  const long long nvox = (long long)N * D * H * W;
  SIVAE_CHECK(nvox > 0, "bn_act_fwd: empty tensor");
  SIVAE_CHECK(nvox * (C / 8) < (1ll << 31), "bn_act_fwd: tensor too large for 32-bit item indices");
  const long long items = (resample == SIVAE_RESAMPLE_AVGPOOL2 ? nvox / 8 : nvox) * (C / 8);
  const int blocks = grid_for(items, 256);
  const __nv_bfloat16 *yy = (const __nv_bfloat16*)y, *rr = (const __nv_bfloat16*)res;
  __nv_bfloat16* oo = (__nv_bfloat16*)out;
  // `mask` with seed == 0: caller-provided byte keep-mask (read).  `mask` with seed != 0 and p > 0: keep-bit STORE --
  // the kernel draws the Philox decisions and writes one bit per element for the backward passes (see philox_keep_byte)
  uint8_t* keep_bits = nullptr;
  if (mask != nullptr && seed != 0ull && p > 0.f) {
    SIVAE_CHECK(resample == SIVAE_RESAMPLE_NONE, "bn_act_fwd: the keep-bit store needs resample = none");
    keep_bits = const_cast<uint8_t*>(mask);
    mask = nullptr;
  }
  if (resample == 0 && rr == nullptr && mask == nullptr && keep_bits == nullptr && p <= 0.f)
    bn_act_plain_fwd_kernel<<<blocks, 256, 0, st>>>(yy, scale, shift, oo, (unsigned)items, C, slope);
  else if (resample == 0)
    bn_act_fwd_kernel<0><<<blocks, 256, 0, st>>>(yy, scale, shift, rr, oo, N, D, H, W, C, slope, mask, p, make_seed_ref(seed), keep_bits);
  else if (resample == 1)
    bn_act_fwd_kernel<1><<<blocks, 256, 0, st>>>(yy, scale, shift, rr, oo, N, D, H, W, C, slope, mask, p, make_seed_ref(seed), nullptr);
  else
    bn_act_fwd_kernel<2><<<blocks, 256, 0, st>>>(yy, scale, shift, rr, oo, N, D, H, W, C, slope, mask, p, make_seed_ref(seed), nullptr);
  SIVAE_LAUNCH_OK("bn_act_fwd_kernel");
  return 0;
}

int bn_act_bwd(const void* g, const void* y, const void* res, const float* mean, const float* invstd,
               const float* gamma, const float* beta, void* dconv, void* dres, float* dgamma, float* dbeta, int N,
               int D, int H, int W, int C, float slope, int resample, const uint8_t* mask, float p,
               unsigned long long seed, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(channels_ok(C), "bn_act_bwd: unsupported channel count %d", C);
  if (check_resample("bn_act_bwd", D, H, W, resample)) return -2;
  SIVAE_CHECK(ws && ws_bytes >= bn_workspace_bytes(C), "bn_act_bwd: workspace too small");
  const long long nvox = (long long)N * D * H * W;
  SIVAE_CHECK(nvox > 0, "bn_act_bwd: empty tensor");
  SIVAE_CHECK(nvox * (C / 8) < (1ll << 31), "bn_act_bwd: tensor too large for 32-bit item indices");
  const int blocks = reduce_blocks(nvox, C);
  float* partial = (float*)ws;
  float* coef = partial + (size_t)kBnMaxBlocks * 2 * C;
  const __nv_bfloat16 *gg = (const __nv_bfloat16*)g, *yy = (const __nv_bfloat16*)y, *rr = (const __nv_bfloat16*)res;
  if (resample == SIVAE_RESAMPLE_NONE && mask == nullptr && p <= 0.f && small_path_ok(nvox, C))
    return launch_small_cluster(bn_small_bwd_kernel, "bn_small_bwd_kernel", st, (const __nv_bfloat16*)g,
                                (const __nv_bfloat16*)y, (const __nv_bfloat16*)res, mean, invstd, gamma, beta,
                                (__nv_bfloat16*)dconv, (__nv_bfloat16*)dres, dgamma, dbeta, nvox, C, slope);
  // `mask` with seed != 0 and p > 0 is the keep-bit store written by bn_act_fwd (1 bit per element)
  const uint8_t* keep_bits = (mask != nullptr && seed != 0ull && p > 0.f) ? mask : nullptr;
  SIVAE_CHECK(keep_bits == nullptr || (resample == SIVAE_RESAMPLE_NONE && res == nullptr && dres == nullptr),
              "bn_act_bwd: the keep-bit store needs resample = none and no residual");
  const bool plain = resample == SIVAE_RESAMPLE_NONE && res == nullptr && (mask == nullptr || keep_bits != nullptr) &&
                     dres == nullptr;
  if (plain) {
    const SeedRef sr = make_seed_ref(seed);
    const long long items = nvox * (C / 8);
    const int ablk = grid_for(items, 256 * 4);
    if (keep_bits) bn_bwd_reduce_plain_kernel<2><<<blocks, kBnThreads, 0, st>>>(gg, yy, mean, invstd, gamma, beta, nvox, C, slope, p, sr, keep_bits, partial);
    else if (p > 0.f) bn_bwd_reduce_plain_kernel<1><<<blocks, kBnThreads, 0, st>>>(gg, yy, mean, invstd, gamma, beta, nvox, C, slope, p, sr, nullptr, partial);
    else bn_bwd_reduce_plain_kernel<0><<<blocks, kBnThreads, 0, st>>>(gg, yy, mean, invstd, gamma, beta, nvox, C, slope, p, sr, nullptr, partial);
    SIVAE_LAUNCH_OK("bn_bwd_reduce_plain_kernel");
    bn_bwd_finalize_kernel<<<cdiv(C, 8), kFinThreads, 0, st>>>(partial, blocks, C, nvox, coef, dgamma, dbeta);
    SIVAE_LAUNCH_OK("bn_bwd_finalize_kernel");
    if (keep_bits) bn_bwd_apply_plain_kernel<2><<<ablk, 256, 0, st>>>(gg, yy, mean, invstd, gamma, beta, coef, (__nv_bfloat16*)dconv, items, C, slope, p, sr, keep_bits);
    else if (p > 0.f) bn_bwd_apply_plain_kernel<1><<<ablk, 256, 0, st>>>(gg, yy, mean, invstd, gamma, beta, coef, (__nv_bfloat16*)dconv, items, C, slope, p, sr, nullptr);
    else bn_bwd_apply_plain_kernel<0><<<ablk, 256, 0, st>>>(gg, yy, mean, invstd, gamma, beta, coef, (__nv_bfloat16*)dconv, items, C, slope, p, sr, nullptr);
    SIVAE_LAUNCH_OK("bn_bwd_apply_plain_kernel");
    return 0;
  }
#define SIVAE_BWD_REDUCE(M)                                                                                      \
  bn_act_bwd_reduce_kernel<M><<<blocks, kBnThreads, 0, st>>>(gg, yy, rr, mean, invstd, gamma, beta, N, D, H, W, C, \
                                                             slope, mask, p, make_seed_ref(seed), partial)
  if (resample == 0) SIVAE_BWD_REDUCE(0);
  else if (resample == 1) SIVAE_BWD_REDUCE(1);
  else SIVAE_BWD_REDUCE(2);
#undef SIVAE_BWD_REDUCE
  SIVAE_LAUNCH_OK("bn_act_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<cdiv(C, 8), kFinThreads, 0, st>>>(partial, blocks, C, nvox, coef, dgamma, dbeta);
  SIVAE_LAUNCH_OK("bn_bwd_finalize_kernel");
  const int ablocks = grid_for(nvox * (C / 8), 256);
#define SIVAE_BWD_APPLY(M)                                                                                      \
  bn_act_bwd_apply_kernel<M><<<ablocks, 256, 0, st>>>(gg, yy, rr, mean, invstd, gamma, beta, coef,               \
                                                      (__nv_bfloat16*)dconv, (__nv_bfloat16*)dres, N, D, H, W, C, \
                                                      slope, mask, p, make_seed_ref(seed))
  if (resample == 0) SIVAE_BWD_APPLY(0);
  else if (resample == 1) SIVAE_BWD_APPLY(1);
  else SIVAE_BWD_APPLY(2);
#undef SIVAE_BWD_APPLY
  SIVAE_LAUNCH_OK("bn_act_bwd_apply_kernel");
  return 0;
}

int ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, long long vox, cudaStream_t st) {
  const long long total = (long long)N * C * vox;
  SIVAE_CHECK(total > 0, "layout: empty tensor");
  ncdhw_to_ndhwc_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, N, C, vox);
  SIVAE_LAUNCH_OK("ncdhw_to_ndhwc_kernel");
  return 0;
}
int ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, long long vox, cudaStream_t st) {
  const long long total = (long long)N * C * vox;
  SIVAE_CHECK(total > 0, "layout: empty tensor");
  ndhwc_to_ncdhw_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)src, dst, N, C, vox);
  SIVAE_LAUNCH_OK("ndhwc_to_ncdhw_kernel");
  return 0;
}

}  // namespace sivae
