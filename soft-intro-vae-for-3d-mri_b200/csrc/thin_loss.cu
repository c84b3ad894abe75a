// thin_loss.cu -- (a) convolutions with one channel on one side (direct, HBM/L2-bound, no tensor cores):
//   encoder stem Conv3d(1,C,3), decoder tail Conv3d(C,1,3)+ReLU+Dropout, mu/var heads Conv3d(C,1,1),
//   decoder stem Conv3d(1,C,1)   (models/models.py:92,118,137-140,216-217) and their gradients;
// (b) latent / loss kernels: reparameterize (models/models.py:263-271), calc_kl (utils/my_trainer.py:38-48),
//   calc_reconstruction_loss (utils/my_trainer.py:62-78) forward + backward, float4 + warp shuffles.
#include "sivae_common.cuh"

namespace sivae {

// neighbour voxel for tap t (T == 27: 3x3x3 pad 1, T == 1: centre); returns -1 when out of bounds
__device__ __forceinline__ long long tap_voxel(int T, int t, long long n, int d, int h, int w, int D, int H, int W) {
  if (T == 1) return ((n * D + d) * H + h) * W + w;
  const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
  if ((unsigned)dd >= (unsigned)D || (unsigned)hh >= (unsigned)H || (unsigned)ww >= (unsigned)W) return -1;
  return ((n * D + dd) * H + hh) * W + ww;
}

// ---- 1 -> C -------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256) c1_to_cn_kernel(const float* __restrict__ x1, const float* __restrict__ w,
                                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                       int N, int D, int H, int W, int C, int flip, int accumulate) {
  extern __shared__ float w_s[];  // [T][C] (tap-major so a thread's 8 channels are contiguous), then bias [C]
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i % T;
    w_s[(flip ? (T - 1 - t) : t) * C + c] = w[i];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) w_s[T * C + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cpc = C >> 3;
  const long long items = (long long)N * D * H * W * cpc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i % cpc);
    const long long v = i / cpc;
    const int wq = (int)(v % W), hq = (int)((v / W) % H), dq = (int)((v / ((long long)W * H)) % D);
    const long long n = v / ((long long)W * H * D);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = w_s[T * C + chunk * 8 + k];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long long vn = tap_voxel(T, t, n, dq, hq, wq, D, H, W);
      const float xv = vn >= 0 ? __ldg(x1 + vn) : 0.f;
      const float* wr = w_s + t * C + chunk * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wr[k], xv, acc[k]);
    }
    __nv_bfloat16* dst = y + v * C + chunk * 8;
    if (accumulate) {
      const uint4 u = *reinterpret_cast<const uint4*>(dst);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// 3x3x3 variant, brick-tiled: a block owns a 4 x 4 x 16 (d,h,w) brick of voxels, stages the 6 x 6 x 18 fp32 halo of the
// one-channel input in shared memory, and each thread produces all C channels of one voxel (27 taps in registers,
// weights broadcast from shared memory).  1728 FMA per voxel for C = 64: FMA-pipe bound, not latency bound.
static constexpr int kBrickD = 4, kBrickH = 4, kBrickW = 16;
static constexpr int kHaloElems = (kBrickD + 2) * (kBrickH + 2) * (kBrickW + 2);

__device__ __forceinline__ void load_halo(const float* __restrict__ x1, float* xs, long long n, int d0, int h0, int w0,
                                          int D, int H, int W) {
  for (int i = threadIdx.x; i < kHaloElems; i += blockDim.x) {
    const int wx = i % (kBrickW + 2), hy = (i / (kBrickW + 2)) % (kBrickH + 2), dz = i / ((kBrickW + 2) * (kBrickH + 2));
    const int d = d0 - 1 + dz, h = h0 - 1 + hy, w = w0 - 1 + wx;
    float v = 0.f;
    if ((unsigned)d < (unsigned)D && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
      v = __ldg(x1 + ((n * D + d) * H + h) * W + w);
    xs[i] = v;
  }
}

__global__ void __launch_bounds__(256) c1_to_cn27_kernel(const float* __restrict__ x1, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ y, int N, int D, int H, int W,
                                                         int C, int flip, int accumulate) {
  extern __shared__ float sm[];
  float* w_s = sm;                 // [27][C]
  float* b_s = sm + 27 * C;        // [C]
  float* xs = b_s + C;             // halo
  for (int i = threadIdx.x; i < C * 27; i += blockDim.x) {
    const int c = i / 27, t = i % 27;
    w_s[(flip ? (26 - t) : t) * C + c] = w[i];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) b_s[i] = bias ? bias[i] : 0.f;
  const int bw = cdiv(W, kBrickW), bh = cdiv(H, kBrickH), bd = cdiv(D, kBrickD);
  const long long bricks = (long long)N * bd * bh * bw;
  const int wq = threadIdx.x % kBrickW, hq = (threadIdx.x / kBrickW) % kBrickH, dq = threadIdx.x / (kBrickW * kBrickH);
  for (long long b = blockIdx.x; b < bricks; b += gridDim.x) {
    const int tw = (int)(b % bw), th = (int)((b / bw) % bh), td = (int)((b / ((long long)bw * bh)) % bd);
    const long long n = b / ((long long)bw * bh * bd);
    const int d0 = td * kBrickD, h0 = th * kBrickH, w0 = tw * kBrickW;
    __syncthreads();
    load_halo(x1, xs, n, d0, h0, w0, D, H, W);
    __syncthreads();
    const int d = d0 + dq, h = h0 + hq, ww = w0 + wq;
    if (d < D && h < H && ww < W) {
      float xv[27];
#pragma unroll
      for (int t = 0; t < 27; ++t)
        xv[t] = xs[((dq + t / 9) * (kBrickH + 2) + hq + (t / 3) % 3) * (kBrickW + 2) + wq + t % 3];
      __nv_bfloat16* dst = y + (((n * D + d) * H + h) * W + ww) * C;
      for (int cg = 0; cg < C; cg += 8) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = b_s[cg + k];
#pragma unroll
        for (int t = 0; t < 27; ++t) {
          const float4 wa = *reinterpret_cast<const float4*>(w_s + t * C + cg);
          const float4 wb = *reinterpret_cast<const float4*>(w_s + t * C + cg + 4);
          acc[0] = fmaf(wa.x, xv[t], acc[0]); acc[1] = fmaf(wa.y, xv[t], acc[1]);
          acc[2] = fmaf(wa.z, xv[t], acc[2]); acc[3] = fmaf(wa.w, xv[t], acc[3]);
          acc[4] = fmaf(wb.x, xv[t], acc[4]); acc[5] = fmaf(wb.y, xv[t], acc[5]);
          acc[6] = fmaf(wb.z, xv[t], acc[6]); acc[7] = fmaf(wb.w, xv[t], acc[7]);
        }
        if (accumulate) {
          const uint4 u = *reinterpret_cast<const uint4*>(dst + cg);
          const float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
          acc[0] += a.x; acc[1] += a.y; acc[2] += bq.x; acc[3] += bq.y; acc[4] += c.x; acc[5] += c.y; acc[6] += e.x; acc[7] += e.y;
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(dst + cg) = o;
      }
    }
  }
}

// ---- C -> 1 -------------------------------------------------------------------------------------
// One warp per output voxel; lane owns CPL = C/32 consecutive channels (coalesced 64..512-byte reads per tap).
// A block covers a 2 x 4 x 32 (d,h,w) brick so the 27-tap neighbourhood mostly hits L1.
template <int CPL>
__device__ __forceinline__ void ld_cpl(const __nv_bfloat16* p, float (&f)[CPL]) {
  if (CPL == 2) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p));
    f[0] = a.x; f[1] = a.y;
  } else if (CPL == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    f[0] = a.x; f[1] = a.y; f[2 % CPL] = b.x; f[3 % CPL] = b.y;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2 % CPL] = b.x; f[3 % CPL] = b.y;
    f[4 % CPL] = c.x; f[5 % CPL] = c.y; f[6 % CPL] = d.x; f[7 % CPL] = d.y;
  }
}

template <int CPL, int T>
__global__ void __launch_bounds__(256) cn_to_c1_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, int N,
                                                       int D, int H, int W, int flip, int act,
                                                       const uint8_t* __restrict__ mask, float p,
                                                       const SeedRef sref) {
  const unsigned long long seed = resolve_seed(sref);
  constexpr int C = CPL * 32;
  constexpr int BD = 2, BH = 4, BW = 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps: warp <-> (d,h) row of the brick
  const int bw = cdiv(W, BW), bh = cdiv(H, BH), bd = cdiv(D, BD);
  float wr[T][CPL];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int k = 0; k < CPL; ++k) wr[t][k] = w[(lane * CPL + k) * T + (flip ? (T - 1 - t) : t)];
  const float b0 = bias ? bias[0] : 0.f;
  const float inv_keep = 1.f / (1.f - p);
  const long long bricks = (long long)N * bd * bh * bw;
  for (long long b = blockIdx.x; b < bricks; b += gridDim.x) {
    const int tw = (int)(b % bw), th = (int)((b / bw) % bh), td = (int)((b / ((long long)bw * bh)) % bd);
    const long long n = b / ((long long)bw * bh * bd);
    const int d = td * BD + (warp >> 2), h = th * BH + (warp & 3);
    if (d >= D || h >= H) continue;
    float mine = 0.f;
    const int w_lo = tw * BW, w_hi = min(W, w_lo + BW);
    for (int wq = w_lo; wq < w_hi; ++wq) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const long long vn = tap_voxel(T, t, n, d, h, wq, D, H, W);
        if (vn >= 0) {
          float f[CPL];
          ld_cpl<CPL>(x + vn * C + lane * CPL, f);
#pragma unroll
          for (int k = 0; k < CPL; ++k) acc = fmaf(wr[t][k], f[k], acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == wq - w_lo) mine = acc;
    }
    const int wq = w_lo + lane;
    if (wq < w_hi) {
      const long long v = ((n * D + d) * H + h) * W + wq;
      float r = mine + b0;
      if (act == 1) {
        r = fmaxf(r, 0.f);
        if (mask != nullptr) r = mask[v] ? r * inv_keep : 0.f;
        else if (p > 0.f) r = philox_keep(seed, (unsigned long long)v, p) ? r * inv_keep : 0.f;
      }
      y[v] = r;
    }
  }
}

// ---- weight gradient of the thin convolutions ---------------------------------------------------
// dw[c][t] = sum_v xc[v][c] * x1[v + delta(t')]   (t' = t, or T-1-t when flip)
template <int CPL, int T>
__global__ void __launch_bounds__(256) wgrad_c1_kernel(const __nv_bfloat16* __restrict__ xc,
                                                       const float* __restrict__ x1, int N, int D, int H, int W,
                                                       int flip, float* __restrict__ partial) {
  constexpr int C = CPL * 32;
  constexpr int PER = T * C + C + 1;  // floats per block partial: dw[t][c], sum_c[c], sum_1
  __shared__ float red[PER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long nvox = (long long)N * D * H * W;
  float acc[T][CPL], sc[CPL], s1 = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k) sc[k] = 0.f;
  for (long long v = (long long)blockIdx.x * nwarps + warp; v < nvox; v += (long long)gridDim.x * nwarps) {
    const int wq = (int)(v % W), hq = (int)((v / W) % H), dq = (int)((v / ((long long)W * H)) % D);
    const long long n = v / ((long long)W * H * D);
    float f[CPL];
    ld_cpl<CPL>(xc + v * C + lane * CPL, f);
#pragma unroll
    for (int k = 0; k < CPL; ++k) sc[k] += f[k];
    s1 += __ldg(x1 + v);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long long vn = tap_voxel(T, flip ? (T - 1 - t) : t, n, dq, hq, wq, D, H, W);
      const float xv = vn >= 0 ? __ldg(x1 + vn) : 0.f;
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[t][k] = fmaf(f[k], xv, acc[t][k]);
    }
  }
  // deterministic cross-warp accumulation in shared memory
  for (int wsel = 0; wsel < nwarps; ++wsel) {
    if (warp == wsel) {
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int idx = t * C + lane * CPL + k;
          red[idx] = (wsel == 0 ? 0.f : red[idx]) + acc[t][k];
        }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int idx = T * C + lane * CPL + k;
        red[idx] = (wsel == 0 ? 0.f : red[idx]) + sc[k];
      }
      if (lane == 0) red[T * C + C] = (wsel == 0 ? 0.f : red[T * C + C]) + s1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < PER; i += blockDim.x) partial[(long long)blockIdx.x * PER + i] = red[i];
}

// 3x3x3 variant, brick-tiled like c1_to_cn27_kernel: the one-channel operand's halo lives in shared memory (27 broadcast
// LDS per voxel), the C-channel operand streams through coalesced 4/8-byte-per-lane loads, four voxels in flight per warp.
template <int CPL>
__global__ void __launch_bounds__(256) wgrad_c1_27_kernel(const __nv_bfloat16* __restrict__ xc,
                                                          const float* __restrict__ x1, int N, int D, int H, int W,
                                                          int flip, float* __restrict__ partial) {
  constexpr int C = CPL * 32, T = 27;
  constexpr int PER = T * C + C + 1;
  __shared__ float red[PER];
  __shared__ float xs[kHaloElems];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[T][CPL], sc[CPL], s1 = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k) sc[k] = 0.f;
  const int bw = cdiv(W, kBrickW), bh = cdiv(H, kBrickH), bd = cdiv(D, kBrickD);
  const long long bricks = (long long)N * bd * bh * bw;
  for (long long b = blockIdx.x; b < bricks; b += gridDim.x) {
    const int tw = (int)(b % bw), th = (int)((b / bw) % bh), td = (int)((b / ((long long)bw * bh)) % bd);
    const long long n = b / ((long long)bw * bh * bd);
    const int d0 = td * kBrickD, h0 = th * kBrickH, w0 = tw * kBrickW;
    __syncthreads();
    load_halo(x1, xs, n, d0, h0, w0, D, H, W);
    __syncthreads();
    // warp `warp` owns voxels [warp*32, warp*32+32) of the brick: 2 h-rows of 16 w at depth warp/2
    const int dq = warp >> 1;
    const int d = d0 + dq;
    if (d >= D) continue;
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += 4) {
      float f[4][CPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const int hq = (warp & 1) * 2 + (j >> 4), wq = j & 15;
        const int h = h0 + hq, ww = w0 + wq;
        if (h < H && ww < W) {
          ld_cpl<CPL>(xc + (((n * D + d) * H + h) * W + ww) * C + lane * CPL, f[u]);
        } else {
#pragma unroll
          for (int k = 0; k < CPL; ++k) f[u][k] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const int hq = (warp & 1) * 2 + (j >> 4), wq = j & 15;
        const float* xb = xs + (dq * (kBrickH + 2) + hq) * (kBrickW + 2) + wq;
#pragma unroll
        for (int k = 0; k < CPL; ++k) sc[k] += f[u][k];
        s1 += xb[(kBrickH + 2) * (kBrickW + 2) + (kBrickW + 2) + 1];   // centre tap (zero outside the volume)
#pragma unroll
        for (int t = 0; t < T; ++t) {
          // offsets are compile-time; the tap flip is applied when the accumulators are written out
          const float xv = xb[(t / 9) * (kBrickH + 2) * (kBrickW + 2) + ((t / 3) % 3) * (kBrickW + 2) + t % 3];
#pragma unroll
          for (int k = 0; k < CPL; ++k) acc[t][k] = fmaf(f[u][k], xv, acc[t][k]);
        }
      }
    }
  }
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  for (int wsel = 0; wsel < nwarps; ++wsel) {
    if (warp == wsel) {
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int idx = (flip ? (T - 1 - t) : t) * C + lane * CPL + k;
          red[idx] = (wsel == 0 ? 0.f : red[idx]) + acc[t][k];
        }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int idx = T * C + lane * CPL + k;
        red[idx] = (wsel == 0 ? 0.f : red[idx]) + sc[k];
      }
      if (lane == 0) red[T * C + C] = (wsel == 0 ? 0.f : red[T * C + C]) + s1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < PER; i += blockDim.x) partial[(long long)blockIdx.x * PER + i] = red[i];
}

// one warp per output element: lanes stride over the per-block partials (fp64, fixed order -> deterministic)
__global__ void wgrad_c1_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, int T, float* dw,
                                         float* sum_c, float* sum_1) {
  const int per = T * C + C + 1;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= per) return;
  double a = 0.0;
  for (int b = lane; b < nblocks; b += 32) a += (double)partial[(long long)b * per + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane != 0) return;
  if (i < T * C) {
    const int t = i / C, c = i % C;
    dw[c * T + t] = (float)a;
  } else if (i < T * C + C) {
    if (sum_c) sum_c[i - T * C] = (float)a;
  } else if (sum_1) {
    // every warp visits every voxel exactly once, so s1 is the plain sum of x1
    sum_1[0] = (float)a;
  }
}

__global__ void relu_drop_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out,
                                     float* __restrict__ dy, long long n, float inv_keep) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dy[i] = out[i] > 0.f ? g[i] * inv_keep : 0.f;
}

// ---- latent / loss ------------------------------------------------------------------------------
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ eps, float eps_c, float* __restrict__ z, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float std = expf(__fmul_rn(0.5f, lv[i]));          // torch.exp(0.5 * logvar)
    const float t = __fmul_rn(eps ? eps[i] : eps_c, std);     // eps * std      (no FMA contraction:
    z[i] = __fadd_rn(mu[i], t);                               // mu + eps*std    three rounded ops, as torch)
  }
}
// Sampler of the reparameterisation (models/models.py:265-266: eps = torch.randn_like(std)): eps ~ N(0,1) drawn in the
// kernel from Philox4x32-10 (one block = 4 uniforms = 2 Box-Muller pairs = 4 normals for elements 4i..4i+3), written
// out for the backward pass, and z = mu + eps*exp(0.5*logvar) with the same three separately rounded operations.
__global__ void reparam_draw_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                        float* __restrict__ eps_out, float* __restrict__ z, long long n,
                                        const SeedRef sref) {
  const unsigned long long seed = resolve_seed(sref);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const long long quads = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)((unsigned long long)q >> 32), 0x5EED0001u, 0u), key);
    // uniforms in (0,1): (x + 0.5) * 2^-32, so the logarithm stays finite
    const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    const float r0 = sqrtf(-2.f * logf(fminf(u0, 0.99999994f))), r1 = sqrtf(-2.f * logf(fminf(u2, 0.99999994f)));
    const float e[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = q * 4 + k;
      if (i < n) {
        const float std = expf(__fmul_rn(0.5f, lv[i]));
        eps_out[i] = e[k];
        z[i] = __fadd_rn(mu[i], __fmul_rn(e[k], std));
      }
    }
  }
}
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ lv,
                                   const float* __restrict__ eps, float eps_c, float* dmu, float* dlv, long long n,
                                   int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dz[i];
    const float std = expf(0.5f * lv[i]);
    const float a = g, b = g * (eps ? eps[i] : eps_c) * std * 0.5f;
    if (dmu) dmu[i] = accumulate ? dmu[i] + a : a;
    if (dlv) dlv[i] = accumulate ? dlv[i] + b : b;
  }
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) v = warp_sum(v);
  return v;  // valid in warp 0
}

__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                     float* __restrict__ kl, long long n) {
  const long long base = (long long)blockIdx.x * n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float m = mu[base + i], l = lv[base + i];
    acc += 1.f + l - m * m - expf(l);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) kl[blockIdx.x] = -0.5f * acc;
}
__global__ void kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ g,
                              float* dmu, float* dlv, int B, long long n, int accumulate) {
  const long long total = (long long)B * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float gb = g[i / n];
    const float a = gb * mu[i], b = gb * 0.5f * (expf(lv[i]) - 1.f);
    if (dmu) dmu[i] = accumulate ? dmu[i] + a : a;
    if (dlv) dlv[i] = accumulate ? dlv[i] + b : b;
  }
}

static constexpr int kMseBlocksPerSample = 64;
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      long long n, float* __restrict__ partial) {
  const long long base = (long long)blockIdx.y * n;
  const long long n4 = n >> 2;
  float acc = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(x + base);
  const float4* y4 = reinterpret_cast<const float4*>(y + base);
  if (((uintptr_t)(x + base) & 15) == 0 && ((uintptr_t)(y + base) & 15) == 0) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 a = x4[i], b = y4[i];
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float d = x[base + i] - y[base + i];
      acc += d * d;
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float d = x[base + i] - y[base + i];
      acc += d * d;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
__global__ void mse_finalize_kernel(const float* __restrict__ partial, int per, float* __restrict__ r) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < per; i += 32) acc += partial[blockIdx.x * per + i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) r[blockIdx.x] = acc;
}
__global__ void mse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g,
                               float* dx, float* dy, int B, long long n) {
  const long long total = (long long)B * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = 2.f * (x[i] - y[i]) * g[i / n];
    if (dx) dx[i] = v;
    if (dy) dy[i] = -v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  if (b > 148ll * 16) b = 148ll * 16;
  if (b < 1) b = 1;
  return (int)b;
}
static constexpr int kWgradC1Blocks = 148 * 4;

int c1_to_cn(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int C, int T,
             int flip, int accumulate, cudaStream_t st) {
  SIVAE_CHECK(C >= 8 && C % 8 == 0, "c1_to_cn: C=%d must be a multiple of 8", C);
  SIVAE_CHECK(T == 1 || T == 27, "c1_to_cn: T=%d must be 1 or 27", T);
  const long long items = (long long)N * D * H * W * (C / 8);
  SIVAE_CHECK(items > 0, "c1_to_cn: empty tensor");
  const size_t smem = (size_t)(T + 1) * C * sizeof(float);
  SIVAE_CHECK(smem <= 48 * 1024, "c1_to_cn: C=%d too large", C);
  const int blocks = grid_for(items, 256);
  if (T == 27) {
    const long long bricks = (long long)N * cdiv(D, kBrickD) * cdiv(H, kBrickH) * cdiv(W, kBrickW);
    const size_t smem27 = ((size_t)28 * C + kHaloElems) * sizeof(float);
    SIVAE_CHECK(smem27 <= 48 * 1024, "c1_to_cn: C=%d too large", C);
    const int b27 = (int)(bricks < 148ll * 8 ? bricks : 148ll * 8);
    c1_to_cn27_kernel<<<b27, 256, smem27, st>>>(x1, w, bias, (__nv_bfloat16*)y, N, D, H, W, C, flip, accumulate);
  } else
    c1_to_cn_kernel<1><<<blocks, 256, smem, st>>>(x1, w, bias, (__nv_bfloat16*)y, N, D, H, W, C, flip, accumulate);
  SIVAE_LAUNCH_OK("c1_to_cn_kernel");
  return 0;
}

int cn_to_c1(const void* x, const float* w, const float* bias, float* y, int N, int D, int H, int W, int C, int T,
             int flip, int act, const uint8_t* mask, float p, unsigned long long seed, cudaStream_t st) {
  SIVAE_CHECK(T == 1 || T == 27, "cn_to_c1: T=%d must be 1 or 27", T);
  SIVAE_CHECK(C == 64 || C == 128 || C == 256, "cn_to_c1: C=%d must be 64, 128 or 256", C);
  SIVAE_CHECK(!(T == 27 && C == 256), "cn_to_c1: (C=256, T=27) is not instantiated");
  SIVAE_CHECK(p >= 0.f && p < 1.f, "cn_to_c1: dropout p=%f out of range", p);
  const long long bricks = (long long)N * cdiv(D, 2) * cdiv(H, 4) * cdiv(W, 32);
  SIVAE_CHECK(bricks > 0, "cn_to_c1: empty tensor");
  const int blocks = (int)(bricks < 148ll * 16 ? bricks : 148ll * 16);
  const __nv_bfloat16* xx = (const __nv_bfloat16*)x;
#define SIVAE_CN1(CPL, TT) \
  cn_to_c1_kernel<CPL, TT><<<blocks, 256, 0, st>>>(xx, w, bias, y, N, D, H, W, flip, act, mask, p, make_seed_ref(seed))
  if (C == 64 && T == 27) SIVAE_CN1(2, 27);
  else if (C == 64) SIVAE_CN1(2, 1);
  else if (C == 128 && T == 27) SIVAE_CN1(4, 27);
  else if (C == 128) SIVAE_CN1(4, 1);
  else SIVAE_CN1(8, 1);
#undef SIVAE_CN1
  SIVAE_LAUNCH_OK("cn_to_c1_kernel");
  return 0;
}

size_t wgrad_c1_workspace_bytes(int N, int D, int H, int W, int C, int T) {
  (void)N; (void)D; (void)H; (void)W;
  return (size_t)kWgradC1Blocks * ((size_t)T * C + C + 1) * sizeof(float);
}

int wgrad_c1(const void* xc, const float* x1, float* dw, float* sum_c, float* sum_1, int N, int D, int H, int W, int C,
             int T, int flip, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(T == 1 || T == 27, "wgrad_c1: T=%d must be 1 or 27", T);
  SIVAE_CHECK(C == 64 || C == 128 || C == 256, "wgrad_c1: C=%d must be 64, 128 or 256", C);
  SIVAE_CHECK(!(T == 27 && C == 256), "wgrad_c1: (C=256, T=27) is not instantiated");
  SIVAE_CHECK(ws && ws_bytes >= wgrad_c1_workspace_bytes(N, D, H, W, C, T), "wgrad_c1: workspace too small");
  const long long nvox = (long long)N * D * H * W;
  SIVAE_CHECK(nvox > 0, "wgrad_c1: empty tensor");
  int blocks = (int)((nvox + 7) / 8);
  if (blocks > kWgradC1Blocks) blocks = kWgradC1Blocks;
  const __nv_bfloat16* xx = (const __nv_bfloat16*)xc;
  float* partial = (float*)ws;
#define SIVAE_WG1(CPL, TT) wgrad_c1_kernel<CPL, TT><<<blocks, 256, 0, st>>>(xx, x1, N, D, H, W, flip, partial)
  if (T == 27) {
    const long long bricks = (long long)N * cdiv(D, kBrickD) * cdiv(H, kBrickH) * cdiv(W, kBrickW);
    blocks = (int)(bricks < kWgradC1Blocks ? bricks : kWgradC1Blocks);
    if (C == 64) wgrad_c1_27_kernel<2><<<blocks, 256, 0, st>>>(xx, x1, N, D, H, W, flip, partial);
    else wgrad_c1_27_kernel<4><<<blocks, 256, 0, st>>>(xx, x1, N, D, H, W, flip, partial);
  } else if (C == 64) SIVAE_WG1(2, 1);
  else if (C == 128) SIVAE_WG1(4, 1);
  else SIVAE_WG1(8, 1);
#undef SIVAE_WG1
  SIVAE_LAUNCH_OK("wgrad_c1_kernel");
  const int per = T * C + C + 1;
  wgrad_c1_finalize_kernel<<<cdiv(per, 8), 256, 0, st>>>(partial, blocks, C, T, dw, sum_c, sum_1);
  SIVAE_LAUNCH_OK("wgrad_c1_finalize_kernel");
  return 0;
}

int relu_drop_bwd(const float* g, const float* out, float* dy, long long n, float p, cudaStream_t st) {
  SIVAE_CHECK(n > 0 && p >= 0.f && p < 1.f, "relu_drop_bwd: bad arguments");
  relu_drop_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(g, out, dy, n, 1.f / (1.f - p));
  SIVAE_LAUNCH_OK("relu_drop_bwd_kernel");
  return 0;
}

int reparam_fwd(const float* mu, const float* lv, const float* eps, float eps_c, float* z, long long n,
                cudaStream_t st) {
  SIVAE_CHECK(n > 0, "reparam_fwd: empty tensor");
  reparam_fwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(mu, lv, eps, eps_c, z, n);
  SIVAE_LAUNCH_OK("reparam_fwd_kernel");
  return 0;
}
int reparam_draw_fwd(const float* mu, const float* lv, float* eps_out, float* z, long long n, unsigned long long seed,
                     cudaStream_t st) {
  SIVAE_CHECK(n > 0, "reparam_draw_fwd: empty tensor");
  SIVAE_CHECK(eps_out != nullptr, "reparam_draw_fwd: eps_out is required (the backward pass reads it)");
  reparam_draw_fwd_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(mu, lv, eps_out, z, n, make_seed_ref(seed));
  SIVAE_LAUNCH_OK("reparam_draw_fwd_kernel");
  return 0;
}
int reparam_bwd(const float* dz, const float* lv, const float* eps, float eps_c, float* dmu, float* dlv, long long n,
                int accumulate, cudaStream_t st) {
  SIVAE_CHECK(n > 0, "reparam_bwd: empty tensor");
  reparam_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(dz, lv, eps, eps_c, dmu, dlv, n, accumulate);
  SIVAE_LAUNCH_OK("reparam_bwd_kernel");
  return 0;
}
int kl_persample_fwd(const float* mu, const float* lv, float* kl, int B, long long n, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && n > 0, "kl_persample_fwd: empty tensor");
  kl_fwd_kernel<<<B, 256, 0, st>>>(mu, lv, kl, n);
  SIVAE_LAUNCH_OK("kl_fwd_kernel");
  return 0;
}
int kl_persample_bwd(const float* mu, const float* lv, const float* g, float* dmu, float* dlv, int B, long long n,
                     int accumulate, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && n > 0, "kl_persample_bwd: empty tensor");
  kl_bwd_kernel<<<grid_for((long long)B * n, 256), 256, 0, st>>>(mu, lv, g, dmu, dlv, B, n, accumulate);
  SIVAE_LAUNCH_OK("kl_bwd_kernel");
  return 0;
}
size_t mse_workspace_bytes(int B, long long n) {
  (void)n;
  return (size_t)B * kMseBlocksPerSample * sizeof(float);
}
int mse_persample_fwd(const float* x, const float* y, float* r, int B, long long n, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  SIVAE_CHECK(B > 0 && n > 0, "mse_persample_fwd: empty tensor");
  SIVAE_CHECK(ws && ws_bytes >= mse_workspace_bytes(B, n), "mse_persample_fwd: workspace too small");
  dim3 grid(kMseBlocksPerSample, B);
  mse_fwd_kernel<<<grid, 256, 0, st>>>(x, y, n, (float*)ws);
  SIVAE_LAUNCH_OK("mse_fwd_kernel");
  mse_finalize_kernel<<<B, 32, 0, st>>>((const float*)ws, kMseBlocksPerSample, r);
  SIVAE_LAUNCH_OK("mse_finalize_kernel");
  return 0;
}
int mse_persample_bwd(const float* x, const float* y, const float* g, float* dx, float* dy, int B, long long n,
                      cudaStream_t st) {
  SIVAE_CHECK(B > 0 && n > 0, "mse_persample_bwd: empty tensor");
  mse_bwd_kernel<<<grid_for((long long)B * n, 256), 256, 0, st>>>(x, y, g, dx, dy, B, n);
  SIVAE_LAUNCH_OK("mse_bwd_kernel");
  return 0;
}

}  // namespace sivae
