// pipeline.cu -- on-GPU input pipeline (NEXT-3).
//
//  * BrainDataset._preprocess (utils/data_load.py:25-30): per volume  v = clip(v, 0, 4*std(v));  v = (v - min) / (max - min)
//    -> one reduction pass (sum, sum of squares, min, max per volume; fp64 accumulation) + one elementwise pass.
//  * tio.RandomAffine(degrees=10) applied with probability 0.35 (aug-z-1200main.py:106-121): trilinear resampling of
//    every volume through its own 3x4 output-voxel -> input-voxel matrix (identity rows for the volumes that are not
//    augmented), samples outside the volume take the pad value (torchio default_pad_value='minimum').
// Both are memory-bound: 4 B read + 4 B written per voxel (the resampler's 8 gathers hit L1/L2).
#include "sivae_common.cuh"

namespace sivae {

static constexpr int kStatBlocks = 64;   // partial blocks per volume

__global__ void __launch_bounds__(256)
volume_stats_kernel(const float* __restrict__ x, long long n, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * n;
  double s = 0.0, s2 = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = xb[i];
    s += v;
    s2 += (double)v * v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  __shared__ double sh[4][256];
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s2; sh[2][threadIdx.x] = mn; sh[3][threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] = fmin(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
      sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[((size_t)b * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = sh[threadIdx.x][0];
}

// stats[b] = {mean, std (population), min, max}
__global__ void volume_stats_finalize_kernel(const double* __restrict__ partial, int nblocks, long long n,
                                             float* __restrict__ stats) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int i = 0; i < nblocks; ++i) {
    const double* p = partial + ((size_t)b * nblocks + i) * 4;
    s += p[0]; s2 += p[1]; mn = fmin(mn, p[2]); mx = fmax(mx, p[3]);
  }
  const double mean = s / (double)n;
  double var = s2 / (double)n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[b * 4 + 0] = (float)mean;
  stats[b * 4 + 1] = (float)sqrt(var);
  stats[b * 4 + 2] = (float)mn;
  stats[b * 4 + 3] = (float)mx;
}

__global__ void __launch_bounds__(256)
clip_minmax_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, const float* __restrict__ stats,
                   float cut_range) {
  const int b = blockIdx.y;
  const float hi = cut_range * stats[b * 4 + 1];
  // min / max of the clipped volume follow from the raw extrema (clip is monotone)
  const float lo_c = fminf(fmaxf(stats[b * 4 + 2], 0.f), hi), hi_c = fminf(fmaxf(stats[b * 4 + 3], 0.f), hi);
  const float inv = 1.f / (hi_c - lo_c);          // inf / NaN for a constant volume, exactly as the reference's division
  const float* xb = x + (size_t)b * n;
  float* yb = y + (size_t)b * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    yb[i] = (fminf(fmaxf(xb[i], 0.f), hi) - lo_c) * inv;
}

// y[b][d][h][w] = trilinear sample of x[b] at  M_b * (d, h, w, 1)  (M_b: 3x4, rows = input d, h, w coordinate)
__global__ void __launch_bounds__(256)
affine_resample_kernel(const float* __restrict__ x, float* __restrict__ y, int D, int H, int W,
                       const float* __restrict__ mats, const float* __restrict__ pad, const float* __restrict__ stats) {
  const int b = blockIdx.y;
  const long long n = (long long)D * H * W;
  const float* xb = x + (size_t)b * n;
  float* yb = y + (size_t)b * n;
  float m[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = mats[b * 12 + i];
  const float padv = pad ? pad[b] : (stats ? stats[b * 4 + 2] : 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / ((long long)W * H));
    const float fd = m[0] * d + m[1] * h + m[2] * w + m[3];
    const float fh = m[4] * d + m[5] * h + m[6] * w + m[7];
    const float fw = m[8] * d + m[9] * h + m[10] * w + m[11];
    const float d0f = floorf(fd), h0f = floorf(fh), w0f = floorf(fw);
    const int d0 = (int)d0f, h0 = (int)h0f, w0 = (int)w0f;
    const float td = fd - d0f, th = fh - h0f, tw = fw - w0f;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dd = d0 + (c >> 2), hh = h0 + ((c >> 1) & 1), ww = w0 + (c & 1);
      const float wgt = ((c >> 2) ? td : 1.f - td) * (((c >> 1) & 1) ? th : 1.f - th) * ((c & 1) ? tw : 1.f - tw);
      const bool in = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
      const float v = in ? __ldg(xb + ((long long)dd * H + hh) * W + ww) : padv;
      acc = fmaf(wgt, v, acc);
    }
    yb[i] = acc;
  }
}

size_t volume_stats_workspace_bytes(int B) { return B > 0 ? (size_t)B * kStatBlocks * 4 * sizeof(double) : 0; }

int volume_stats(const float* x, int B, long long n, float* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && n > 0, "volume_stats: empty input");
  SIVAE_CHECK(ws && ws_bytes >= volume_stats_workspace_bytes(B), "volume_stats: workspace too small");
  volume_stats_kernel<<<dim3(kStatBlocks, (unsigned)B), 256, 0, st>>>(x, n, (double*)ws);
  SIVAE_LAUNCH_OK("volume_stats_kernel");
  volume_stats_finalize_kernel<<<B, 32, 0, st>>>((const double*)ws, kStatBlocks, n, stats);
  SIVAE_LAUNCH_OK("volume_stats_finalize_kernel");
  return 0;
}

int preprocess_clip_minmax(const float* x, float* y, int B, long long n, float cut_range, float* stats, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(cut_range > 0.f && stats != nullptr, "preprocess_clip_minmax: bad arguments");
  if (int rc = volume_stats(x, B, n, stats, ws, ws_bytes, st)) return rc;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  clip_minmax_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, st>>>(x, y, n, stats, cut_range);
  SIVAE_LAUNCH_OK("clip_minmax_kernel");
  return 0;
}

int affine_resample(const float* x, float* y, int B, int D, int H, int W, const float* mats, const float* pad,
                    const float* stats, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && D > 0 && H > 0 && W > 0 && mats != nullptr, "affine_resample: bad arguments");
  SIVAE_CHECK(x != y, "affine_resample: in-place resampling is not supported");
  const long long n = (long long)D * H * W;
  long long blocks = (n + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  affine_resample_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, st>>>(x, y, D, H, W, mats, pad, stats);
  SIVAE_LAUNCH_OK("affine_resample_kernel");
  return 0;
}

}  // namespace sivae
