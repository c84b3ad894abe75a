// linear.cu -- the fully-connected latent heads of the FC-latent Soft-IntroVAE variant (SURVEY.md section 8f NEXT-1).
//
// Reference: models/mymodel.py:125 ``fc = Linear(forth_ch*5*6*5, 2*z_ch)`` applied to the NCDHW-flattened feature map
// (:140-142), and :150-153 ``dfc = Linear(z_ch, forth_ch*150) -> ReLU`` reshaped to [B, forth_ch, 5, 6, 5] (:219).
// With batch <= 8 these are weight-streaming products: every weight is used once per sample, so all three passes
// (forward, data gradient, weight gradient) are HBM-bound on the fp32 weight matrix W [J][K] (184 MB for 38400 x 1200).
// Nothing here is GEMM-shaped enough for the tensor cores to matter (AI = B/2 FLOP per byte); the kernels are written
// for coalesced 16-byte weight loads with enough of them in flight, fp32 FMA accumulation, and deterministic two-pass
// split reductions (no atomics).
//
//   linear_fwd   y[b][j]  = act(bias[j] + sum_k x[b][k] W[j][k])      rows of W streamed once per 8 samples
//   linear_dgrad dx[b][k] = sum_j dy[b][j] W[j][k]                    columns of W, coalesced along k
//   linear_wgrad dW[j][k] = sum_b dy[b][j] x[b][k];  db[j] = sum_b dy[b][j]      write-bound
//
// plus the two layout changes between the convolutional trunk (NDHWC bf16, channels padded to 64) and the Linear heads
// (NCDHW-flattened fp32, feature index c*S + s) and the element-wise ``LeakyReLU(a + b)`` of mymodel.py:136 whose
// branch already carries its own activation (so it is not the fused BN-residual form).
#include "sivae_common.cuh"

namespace sivae {

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static constexpr int kLinB = 8;          // samples per pass; larger batches are processed in chunks of 8
static constexpr int kLinKC = 1024;      // floats of x per sample staged in shared memory per chunk
static constexpr int kLinRowsPerWarp = 4;
static constexpr int kLinRowsPerCta = 32;  // 8 warps x 4 rows

__device__ __forceinline__ void vload(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void vload(const float* p, float (&v)[1]) { v[0] = __ldg(p); }

// partial[(split*B + b)*J + j] = sum over this split's k-range of x[b][k] * W[j][k]
template <int VEC>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int B, int K, int J, int k_per_split,
                  float* __restrict__ partial) {
  __shared__ __align__(16) float xs[kLinB][kLinKC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * kLinRowsPerCta + warp * kLinRowsPerWarp;
  const int k_lo = blockIdx.y * k_per_split, k_hi = min(K, k_lo + k_per_split);
  const float* wrow[kLinRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kLinRowsPerWarp; ++r) wrow[r] = W + (size_t)min(j0 + r, J - 1) * K;   // clamped rows are discarded
  float acc[kLinRowsPerWarp][kLinB] = {};
  for (int kc = k_lo; kc < k_hi; kc += kLinKC) {
    const int len = min(kLinKC, k_hi - kc);
    __syncthreads();
    for (int i = threadIdx.x; i < kLinB * kLinKC; i += 256) {
      const int b = i / kLinKC, c = i % kLinKC;
      xs[b][c] = (b < B && c < len) ? x[(size_t)b * K + kc + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = lane; i < len / VEC; i += 32) {
      float w[kLinRowsPerWarp][VEC];
#pragma unroll
      for (int r = 0; r < kLinRowsPerWarp; ++r) vload(wrow[r] + kc + i * VEC, w[r]);
#pragma unroll
      for (int b = 0; b < kLinB; ++b) {
        float xv[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) xv[v] = xs[b][i * VEC + v];
#pragma unroll
        for (int r = 0; r < kLinRowsPerWarp; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[r][b] = fmaf(w[r][v], xv[v], acc[r][b]);
      }
    }
  }
  // 32 partial sums per lane, 32 lanes: butterfly that halves the live values each step (31 shuffles instead of 160);
  // lane l ends up holding the warp total of value l = r * 8 + b
  float v[32];
#pragma unroll
  for (int r = 0; r < kLinRowsPerWarp; ++r)
#pragma unroll
    for (int b = 0; b < kLinB; ++b) v[r * kLinB + b] = acc[r][b];
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  {
    const int r = lane / kLinB, b = lane % kLinB;
    if (j0 + r < J && b < B) partial[((size_t)blockIdx.y * B + b) * J + j0 + r] = v[0];
  }
}

// out[b][i] = act(bias[i] + sum_s partial[(s*B + b)*n + i]);  act: 0 none, 1 ReLU
__global__ void linear_finalize_kernel(const float* __restrict__ partial, int splits, int B, int n,
                                       const float* __restrict__ bias, int act, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * n) return;
  const int b = (int)(idx / n), i = (int)(idx % n);
  float s = bias ? bias[i] : 0.f;
  for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)sp * B + b) * n + i];
  out[idx] = (act == 1) ? fmaxf(s, 0.f) : s;
}

// partial[(split*B + b)*K + k] = sum over this split's rows of dy[b][j] * W[j][k]
template <int VEC>
__global__ void __launch_bounds__(256)
linear_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ W, int B, int K, int J, int j_per_split,
                    float* __restrict__ partial) {
  constexpr int kJC = 128;
  __shared__ __align__(16) float dys[kJC][kLinB];
  __shared__ __align__(16) float red[3][kLinB][64 * VEC];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int col = (blockIdx.x * 64 + tx) * VEC;
  const bool live = col < K;
  const int j_lo = blockIdx.y * j_per_split, j_hi = min(J, j_lo + j_per_split);
  float acc[kLinB][VEC] = {};
  for (int jc = j_lo; jc < j_hi; jc += kJC) {
    const int len = min(kJC, j_hi - jc);
    __syncthreads();
    for (int i = threadIdx.x; i < kJC * kLinB; i += 256) {
      const int b = i / kJC, jj = i % kJC;                     // consecutive threads read consecutive j: coalesced
      dys[jj][b] = (b < B && jj < len) ? dy[(size_t)b * J + jc + jj] : 0.f;
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int jj = ty; jj < len; jj += 4) {
        float w[VEC];
        vload(W + (size_t)(jc + jj) * K + col, w);
        const float4 d0 = *reinterpret_cast<const float4*>(&dys[jj][0]);
        const float4 d1 = *reinterpret_cast<const float4*>(&dys[jj][4]);
        const float d[kLinB] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int b = 0; b < kLinB; ++b)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[b][v] = fmaf(d[b], w[v], acc[b][v]);
      }
    }
  }
  if (ty > 0) {
#pragma unroll
    for (int b = 0; b < kLinB; ++b)
#pragma unroll
      for (int v = 0; v < VEC; ++v) red[ty - 1][b][tx * VEC + v] = acc[b][v];
  }
  __syncthreads();
  if (ty == 0 && live) {
#pragma unroll
    for (int b = 0; b < kLinB; ++b) {
      if (b >= B) break;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float s = acc[b][v] + red[0][b][tx * VEC + v] + red[1][b][tx * VEC + v] + red[2][b][tx * VEC + v];
        partial[((size_t)blockIdx.y * B + b) * K + col + v] = s;
      }
    }
  }
}

// dW[j][k] = sum_b dy[b][j] x[b][k]  (any B);  db[j] = sum_b dy[b][j]
template <int VEC>
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int K, int J,
                    float* __restrict__ dW, float* __restrict__ db) {
  constexpr int kRows = 8;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int col = (blockIdx.x * 64 + tx) * VEC;
  const int j0 = blockIdx.y * (4 * kRows) + ty * kRows;
  if (blockIdx.x == 0 && tx < kRows && db != nullptr) {
    const int j = j0 + tx;
    if (j < J) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dy[(size_t)b * J + j];
      db[j] = s;
    }
  }
  if (col >= K) return;
  float acc[kRows][VEC] = {};
  for (int b = 0; b < B; ++b) {
    float xv[VEC];
    vload(x + (size_t)b * K + col, xv);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float d = (j0 + r < J) ? __ldg(dy + (size_t)b * J + j0 + r) : 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = fmaf(d, xv[v], acc[r][v]);
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    if (j0 + r >= J) break;
    float* dst = dW + (size_t)(j0 + r) * K + col;
    if (VEC == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][VEC > 1 ? 1 : 0], acc[r][VEC > 2 ? 2 : 0],
                                                    acc[r][VEC > 3 ? 3 : 0]);
    } else {
      dst[0] = acc[r][0];
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
// Split factors are chosen so that the grid fills a whole number of waves of resident CTAs (a 304-CTA grid on 296 slots
// runs for two CTA lifetimes): slots = occupancy x SM count, queried once per kernel.
template <typename Kern>
static int resident_slots(Kern kern) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ <= 0) occ = 2;
  return occ * num_sms();
}
static int fwd_splits(int K, int J, int vec) {
  static int slots4 = 0, slots1 = 0;
  if (!slots4) { slots4 = resident_slots(linear_fwd_kernel<4>); slots1 = resident_slots(linear_fwd_kernel<1>); }
  const int slots = vec == 4 ? slots4 : slots1;
  const int row_blocks = cdiv(J, kLinRowsPerCta);
  // cost model: waves x (k-range of one CTA + a fixed per-CTA overhead worth ~512 columns)
  int best = 1;
  long long best_cost = -1;
  for (int ks = 1; ks <= cdiv(K, kLinKC); ++ks) {
    const long long cost = (long long)cdiv((long long)row_blocks * ks, slots) * (cdiv(K, ks) + 512);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ks; }
  }
  return best;
}
static int dgrad_splits(int B, int K, int J, int vec) {
  static int slots4 = 0, slots1 = 0;
  if (!slots4) { slots4 = resident_slots(linear_dgrad_kernel<4>); slots1 = resident_slots(linear_dgrad_kernel<1>); }
  const int slots = vec == 4 ? slots4 : slots1;
  const int col_blocks = cdiv(K, 64 * vec);
  // cost model in units of weight rows per CTA: waves x (rows of one CTA + ~24 rows of fixed overhead) plus the partial
  // buffer's write + read traffic (2 B js rows per column block, spread over the resident CTAs)
  int best = 1;
  double best_cost = -1.0;
  const int js_max = min(cdiv(J, 32), 512);
  for (int js = 1; js <= js_max; ++js) {
    const double waves = (double)cdiv((long long)col_blocks * js, slots);
    const double cost = waves * (cdiv(J, js) + 24) + (double)col_blocks * 2.0 * B * js / slots;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = js; }
  }
  return best;
}

size_t linear_workspace_bytes(int B, int K, int J) {
  if (B <= 0 || K <= 0 || J <= 0) return 0;
  const int vec = (K % 4 == 0) ? 4 : 1;
  const size_t bb = (size_t)min(B, kLinB);
  const size_t f = (size_t)fwd_splits(K, J, vec) * bb * J, d = (size_t)dgrad_splits((int)bb, K, J, vec) * bb * K;
  return (f > d ? f : d) * sizeof(float);
}

int linear_fwd(const float* x, const float* W, const float* bias, float* y, int B, int K, int J, int act, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && K > 0 && J > 0, "linear_fwd: empty problem (B=%d K=%d J=%d)", B, K, J);
  SIVAE_CHECK(act == 0 || act == 1, "linear_fwd: act must be 0 (none) or 1 (ReLU)");
  SIVAE_CHECK(ws != nullptr && ws_bytes >= linear_workspace_bytes(B, K, J), "linear_fwd: workspace too small");
  const int vec = (K % 4 == 0) ? 4 : 1;
  const int ks = fwd_splits(K, J, vec);
  int kps = cdiv(K, ks);
  kps = (kps + 3) / 4 * 4;
  float* partial = static_cast<float*>(ws);
  for (int b0 = 0; b0 < B; b0 += kLinB) {
    const int bb = min(kLinB, B - b0);
    const dim3 grid(cdiv(J, kLinRowsPerCta), cdiv(K, kps));
    if (vec == 4)
      linear_fwd_kernel<4><<<grid, 256, 0, st>>>(x + (size_t)b0 * K, W, bb, K, J, kps, partial);
    else
      linear_fwd_kernel<1><<<grid, 256, 0, st>>>(x + (size_t)b0 * K, W, bb, K, J, kps, partial);
    SIVAE_LAUNCH_OK("linear_fwd_kernel");
    linear_finalize_kernel<<<cdiv((long long)bb * J, 256), 256, 0, st>>>(partial, grid.y, bb, J, bias, act,
                                                                         y + (size_t)b0 * J);
    SIVAE_LAUNCH_OK("linear_finalize_kernel");
  }
  return 0;
}

int linear_dgrad(const float* dy, const float* W, float* dx, int B, int K, int J, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  SIVAE_CHECK(B > 0 && K > 0 && J > 0, "linear_dgrad: empty problem (B=%d K=%d J=%d)", B, K, J);
  SIVAE_CHECK(ws != nullptr && ws_bytes >= linear_workspace_bytes(B, K, J), "linear_dgrad: workspace too small");
  const int vec = (K % 4 == 0) ? 4 : 1;
  const int js = dgrad_splits(min(B, kLinB), K, J, vec);
  const int jps = cdiv(J, js);
  float* partial = static_cast<float*>(ws);
  for (int b0 = 0; b0 < B; b0 += kLinB) {
    const int bb = min(kLinB, B - b0);
    const dim3 grid(cdiv(K, 64 * vec), cdiv(J, jps));
    if (vec == 4)
      linear_dgrad_kernel<4><<<grid, 256, 0, st>>>(dy + (size_t)b0 * J, W, bb, K, J, jps, partial);
    else
      linear_dgrad_kernel<1><<<grid, 256, 0, st>>>(dy + (size_t)b0 * J, W, bb, K, J, jps, partial);
    SIVAE_LAUNCH_OK("linear_dgrad_kernel");
    linear_finalize_kernel<<<cdiv((long long)bb * K, 256), 256, 0, st>>>(partial, grid.y, bb, K, nullptr, 0,
                                                                         dx + (size_t)b0 * K);
    SIVAE_LAUNCH_OK("linear_finalize_kernel");
  }
  return 0;
}

int linear_wgrad(const float* x, const float* dy, float* dW, float* db, int B, int K, int J, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && K > 0 && J > 0, "linear_wgrad: empty problem (B=%d K=%d J=%d)", B, K, J);
  const int vec = (K % 4 == 0) ? 4 : 1;
  const dim3 grid(cdiv(K, 64 * vec), cdiv(J, 32));
  if (vec == 4)
    linear_wgrad_kernel<4><<<grid, 256, 0, st>>>(x, dy, B, K, J, dW, db);
  else
    linear_wgrad_kernel<1><<<grid, 256, 0, st>>>(x, dy, B, K, J, dW, db);
  SIVAE_LAUNCH_OK("linear_wgrad_kernel");
  return 0;
}

// ---- layout changes between the NDHWC bf16 trunk and the NCDHW-flattened fp32 heads --------------------------------
// dst[b][c*S + s] = src[b][s][c] * (gate == nullptr || gate[b][c*S + s] > 0)
__global__ void ndhwc_to_flat_kernel(const __nv_bfloat16* __restrict__ src, int B, int S, int C, int Cp,
                                     const float* __restrict__ gate, float* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * C * S) return;
  const int b = (int)(idx / ((long long)C * S));
  const int r = (int)(idx % ((long long)C * S));
  const int c = r / S, s = r % S;
  float v = __bfloat162float(src[((size_t)b * S + s) * Cp + c]);
  if (gate != nullptr && !(gate[idx] > 0.f)) v = 0.f;
  dst[idx] = v;
}
// dst[b][s][c] = c < C ? src[b][c*S + s] : 0
__global__ void flat_to_ndhwc_kernel(const float* __restrict__ src, int B, int S, int C, int Cp,
                                     __nv_bfloat16* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * S * Cp) return;
  const int c = (int)(idx % Cp);
  const long long bs = idx / Cp;
  const int s = (int)(bs % S), b = (int)(bs / S);
  dst[idx] = __float2bfloat16(c < C ? src[((size_t)b * C + c) * S + s] : 0.f);
}

int ndhwc_to_flat(const void* src, float* dst, int B, int S, int C, int Cp, const float* gate, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && S > 0 && C > 0 && Cp >= C, "ndhwc_to_flat: bad shape (B=%d S=%d C=%d Cp=%d)", B, S, C, Cp);
  ndhwc_to_flat_kernel<<<cdiv((long long)B * C * S, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), B, S, C,
                                                                        Cp, gate, dst);
  SIVAE_LAUNCH_OK("ndhwc_to_flat_kernel");
  return 0;
}
int flat_to_ndhwc(const float* src, void* dst, int B, int S, int C, int Cp, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && S > 0 && C > 0 && Cp >= C, "flat_to_ndhwc: bad shape (B=%d S=%d C=%d Cp=%d)", B, S, C, Cp);
  flat_to_ndhwc_kernel<<<cdiv((long long)B * S * Cp, 256), 256, 0, st>>>(src, B, S, C, Cp,
                                                                         static_cast<__nv_bfloat16*>(dst));
  SIVAE_LAUNCH_OK("flat_to_ndhwc_kernel");
  return 0;
}

// ---- out = LeakyReLU(a + b)  (mymodel.py:136) and its gradient, bf16, 8 elements per thread -------------------------
__global__ void add_act_fwd_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long n8, float slope,
                                   uint4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 va = a[i], vb = b[i];
  const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, ub[4] = {vb.x, vb.y, vb.z, vb.w};
  uint32_t r[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 fa = unpack_bf16x2(ua[q]), fb = unpack_bf16x2(ub[q]);
    float s0 = fa.x + fb.x, s1 = fa.y + fb.y;
    s0 = s0 > 0.f ? s0 : s0 * slope;
    s1 = s1 > 0.f ? s1 : s1 * slope;
    r[q] = pack_bf16x2(s0, s1);
  }
  out[i] = make_uint4(r[0], r[1], r[2], r[3]);
}
// d(a) = d(b) = g * (out > 0 ? 1 : slope)
__global__ void add_act_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ out, long long n8, float slope,
                                   uint4* __restrict__ dz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 vg = g[i], vo = out[i];
  const uint32_t ug[4] = {vg.x, vg.y, vg.z, vg.w}, uo[4] = {vo.x, vo.y, vo.z, vo.w};
  uint32_t r[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 fg = unpack_bf16x2(ug[q]), fo = unpack_bf16x2(uo[q]);
    r[q] = pack_bf16x2(fo.x > 0.f ? fg.x : fg.x * slope, fo.y > 0.f ? fg.y : fg.y * slope);
  }
  dz[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

int add_act_fwd(const void* a, const void* b, void* out, long long n, float slope, cudaStream_t st) {
  SIVAE_CHECK(n > 0 && n % 8 == 0, "add_act_fwd: element count %lld must be a positive multiple of 8", n);
  add_act_fwd_kernel<<<cdiv(n / 8, 256), 256, 0, st>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b), n / 8,
                                                       slope, static_cast<uint4*>(out));
  SIVAE_LAUNCH_OK("add_act_fwd_kernel");
  return 0;
}
int add_act_bwd(const void* g, const void* out, void* dz, long long n, float slope, cudaStream_t st) {
  SIVAE_CHECK(n > 0 && n % 8 == 0, "add_act_bwd: element count %lld must be a positive multiple of 8", n);
  add_act_bwd_kernel<<<cdiv(n / 8, 256), 256, 0, st>>>(static_cast<const uint4*>(g), static_cast<const uint4*>(out), n / 8,
                                                       slope, static_cast<uint4*>(dz));
  SIVAE_LAUNCH_OK("add_act_bwd_kernel");
  return 0;
}

}  // namespace sivae
