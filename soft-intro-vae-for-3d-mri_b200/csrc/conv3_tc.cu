// conv3_tc.cu -- 3x3x3 convolution (stride 1, pad 1) as implicit GEMM on tcgen05 / TMEM, fed by TMA.
//
// Replaces the cuDNN calls behind nn.Conv3d(k=3,p=1,bias=False) in the reference's
// BuildingBlock / UpsampleBuildingkBlock (models/models.py:17,21,55,59) for forward, data gradient
// (same kernel on repacked weights) and weight gradient (split-K kernel below).
//
// Data layout: activations NDHWC bf16.  A 5-D TMA tensor map {C, W, H, D, N} with a {64, wt, ht, dt, 1}
// box and 128B swizzle drops a (wt*ht*dt) x 64-channel tile into shared memory as dense 128-byte rows
// -- exactly the canonical K-major SWIZZLE_128B UMMA operand.  The conv halo / zero padding comes for
// free from TMA out-of-bounds zero fill with (possibly negative) shifted box coordinates.
//
//   fprop / dgrad : D[128 voxels x BLOCK_N] += A[voxels x 64ci](tap-shifted box) * B[BLOCK_N x 64ci](tap slab)
//                   K loop = 27 taps x Cin/64;   A, B K-major.
//   wgrad         : D[2 units x 64ci, NT co]  += A[voxels x 64ci]^T (tap-shifted box) * B[voxels x 64co] (dy box)
//                   K loop = voxel boxes;        A, B MN-major (voxel rows are the K dimension).
#include "sivae_common.cuh"

namespace sivae {

static constexpr int kTileRows = 128;            // UMMA M (fprop) / max voxel rows per box
static constexpr int kTileBytes = kTileRows * 128;  // one 128-row x 64-channel bf16 tile

struct ConvGeom {
  int N, D, H, W;
  int wt, ht, dt;  // box extents of one tile of output voxels
  int tiles_w, tiles_h, tiles_d;
  int rows;        // wt*ht*dt  (<= 128)
  int cin_blocks;  // Cin / 64
};

template <class Geom>
__device__ __forceinline__ void decode_tile(const Geom& g, long long id, int& w0, int& h0, int& d0, int& n) {
  int tw = (int)(id % g.tiles_w);
  id /= g.tiles_w;
  int th = (int)(id % g.tiles_h);
  id /= g.tiles_h;
  int td = (int)(id % g.tiles_d);
  n = (int)(id / g.tiles_d);
  w0 = tw * g.wt;
  h0 = th * g.ht;
  d0 = td * g.dt;
}

// =================================================================================================
// fprop / dgrad
// =================================================================================================
// Epilogue of the single-output-channel variant (decoder tail Conv3d(C,1,3)+ReLU+Dropout, models/models.py:137-140,
// and the input gradient of the encoder stem): only accumulator column 0 is meaningful; it is written as fp32
// [N][D][H][W] with bias / ReLU / dropout fused.
struct ToOneEpilogue {
  float* y;
  const float* bias;   // 1 element or NULL
  const uint8_t* mask; // [N][D][H][W] keep-mask or NULL
  SeedRef seed;
  float p;
  int act;             // 0 identity, 1 ReLU + dropout
};

template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(192)
conv3_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const ConvGeom g, const ToOneEpilogue ep) {
  constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = kTileBytes + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int w0, h0, d0, n;
  decode_tile(g, blockIdx.x, w0, h0, d0, n);
  const int nb = blockIdx.y;
  const int num_kb = 27 * g.cin_blocks;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)g.rows * 128u + (uint32_t)B_BYTES;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], tx_bytes);
        const int tap = kb / g.cin_blocks, cb = kb - tap * g.cin_blocks;
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        uint8_t* a_dst = smem + s * STAGE_BYTES;
        tma_load_5d(a_dst, &tmA, &full_bar[s], cb * 64, w0 + kw - 1, h0 + kh - 1, d0 + kd - 1, n);
        tma_load_3d(a_dst + kTileBytes, &tmB, &full_bar[s], cb * 64, nb * BLOCK_N, tap);
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 64-channel K block = 4 x UMMA_K(16)
          umma_bf16(tmem_base, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                    idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bf16 -> swizzled smem -> TMA store =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_id & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    if constexpr (EPI == 1) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      const int w = w0 + row % g.wt, h = h0 + (row / g.wt) % g.ht, d = d0 + row / (g.wt * g.ht);
      if (row < g.rows && w < g.W && h < g.H && d < g.D) {
        const long long vox = (((long long)n * g.D + d) * g.H + h) * g.W + w;
        float r = __uint_as_float(v[0]) + (ep.bias ? ep.bias[0] : 0.f);
        if (ep.act == 1) {
          r = fmaxf(r, 0.f);
          const float inv_keep = 1.f / (1.f - ep.p);
          if (ep.mask != nullptr) r = ep.mask[vox] ? r * inv_keep : 0.f;
          else if (ep.p > 0.f) r = philox_keep(resolve_seed(ep.seed), (unsigned long long)vox, ep.p) ? r * inv_keep : 0.f;
        }
        ep.y[vox] = r;
      }
    } else {
    uint8_t* out_stage = smem;  // pipeline buffers are idle once tmem_full has fired
#pragma unroll 1
    for (int j = 0; j < BLOCK_N / 32; ++j) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
      tmem_ld_wait();
      uint8_t* tile = out_stage + (j >> 1) * kTileBytes + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
        pk.y = pack_bf16x2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
        pk.z = pack_bf16x2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
        pk.w = pack_bf16x2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
        const int chunk = (j & 1) * 4 + c;
        *reinterpret_cast<uint4*>(tile + ((chunk ^ (row & 7)) << 4)) = pk;
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp_id == 2 && lane == 0) {
#pragma unroll
      for (int jb = 0; jb < BLOCK_N / 64; ++jb)
        tma_store_5d(&tmC, out_stage + jb * kTileBytes, nb * BLOCK_N + jb * 64, w0, h0, d0, n);
      tma_store_commit();
      tma_store_wait_all();
    }
    }  // EPI == 0
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================================================================================
// wgrad
// =================================================================================================
struct WgradGeom {
  int N, D, H, W;
  int wt, ht, dt, tiles_w, tiles_h, tiles_d;
  int rows;             // voxel rows per box (<= 128); consumed in 16-row K steps, tail rows zeroed
  int cin_blocks;       // Cin/64
  int units;            // 27 * cin_blocks   (one unit = one tap x one 64-wide Cin block)
  int pairs_total;      // ceil(units/2)
  int ppc;              // unit pairs accumulated per CTA (TMEM: ppc * NT columns)
  int groups;           // ceil(pairs_total / ppc)
  int Cout;
  long long total_kb;   // voxel boxes in the whole tensor
  long long kb_per_split;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // descriptor strides (bytes), see make_smem_desc
};

template <int NT, int A_STAGES>
__global__ void __launch_bounds__(192, 1)
conv3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                   const WgradGeom g, float* __restrict__ partial) {
  constexpr int B_STAGES = 2;
  constexpr int B_STAGE_BYTES = (NT / 64) * kTileBytes;
  constexpr int A_STAGE_BYTES = 2 * kTileBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + B_STAGES * B_STAGE_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_a + A_STAGES * A_STAGE_BYTES);
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* tmem_full_bar = b_empty + B_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = blockIdx.x % g.groups;
  const int ntile = blockIdx.x / g.groups;
  const int pair0 = grp * g.ppc;
  const int npairs = min(g.ppc, g.pairs_total - pair0);
  const long long kb0 = (long long)blockIdx.y * g.kb_per_split;
  const long long kb1 = min(g.total_kb, kb0 + g.kb_per_split);

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 512);
  // The voxel rows are the GEMM K dimension and are consumed 16 at a time: rows [rows, roundup16(rows)) of
  // every tile are never written by TMA, so zero them once (generic proxy -> async proxy fence).
  const int rows_pad = (g.rows + 15) & ~15;
  if (rows_pad != g.rows) {
    constexpr int kTiles = B_STAGES * (NT / 64) + A_STAGES * 2;   // tiles are contiguous from smem
    const int tail_vec = (rows_pad - g.rows) * 8;                  // uint4 per tile tail
    for (int i = threadIdx.x; i < kTiles * tail_vec; i += blockDim.x) {
      const int tile = i / tail_vec, off = i - tile * tail_vec;
      reinterpret_cast<uint4*>(smem + (size_t)tile * kTileBytes + (size_t)g.rows * 128)[off] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tile_tx = (uint32_t)g.rows * 128u;

  if (warp_id == 0) {
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (long long kb = kb0; kb < kb1; ++kb) {
        int w0, h0, d0, n;
        decode_tile(g, kb, w0, h0, d0, n);
        {
          const int s = b_it % B_STAGES;
          mbar_wait(&b_empty[s], ((b_it / B_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(&b_full[s], tile_tx * (NT / 64));
#pragma unroll
          for (int jc = 0; jc < NT / 64; ++jc)
            tma_load_5d(smem_b + s * B_STAGE_BYTES + jc * kTileBytes, &tmDY, &b_full[s], ntile * NT + jc * 64, w0, h0,
                        d0, n);
          ++b_it;
        }
        for (int lp = 0; lp < npairs; ++lp) {
          const int s = a_it % A_STAGES;
          mbar_wait(&a_empty[s], ((a_it / A_STAGES) & 1u) ^ 1u);
          const int u0 = 2 * (pair0 + lp);
          const int nu = (u0 + 1 < g.units) ? 2 : 1;
          mbar_expect_tx(&a_full[s], tile_tx * nu);
          for (int uu = 0; uu < nu; ++uu) {
            const int u = u0 + uu;
            const int tap = u / g.cin_blocks, cb = u - tap * g.cin_blocks;
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            tma_load_5d(smem_a + s * A_STAGE_BYTES + uu * kTileBytes, &tmX, &a_full[s], cb * 64, w0 + kw - 1,
                        h0 + kh - 1, d0 + kd - 1, n);
          }
          ++a_it;
        }
      }
    }
  } else if (warp_id == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
    const int k16s = rows_pad / 16;
    uint32_t a_it = 0, b_it = 0;
    for (long long kb = kb0; kb < kb1; ++kb) {
      const int bs = b_it % B_STAGES;
      mbar_wait(&b_full[bs], (b_it / B_STAGES) & 1u);
      const uint32_t b_addr = smem_u32(smem_b + bs * B_STAGE_BYTES);
      for (int lp = 0; lp < npairs; ++lp) {
        const int as = a_it % A_STAGES;
        mbar_wait(&a_full[as], (a_it / A_STAGES) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + as * A_STAGE_BYTES);
          for (int k = 0; k < k16s; ++k) {  // 16 voxel rows (2048 B) per UMMA
            umma_bf16(tmem_base + (uint32_t)(lp * NT), make_smem_desc(a_addr + k * 2048, g.a_lbo, g.a_sbo),
                      make_smem_desc(b_addr + k * 2048, g.b_lbo, g.b_sbo), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&a_empty[as]);
          if (lp == npairs - 1) {
            umma_commit(&b_empty[bs]);
            if (kb == kb1 - 1) umma_commit(tmem_full_bar);
          }
        }
        __syncwarp();
        ++a_it;
      }
      ++b_it;
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    for (int lp = 0; lp < npairs; ++lp) {
      const int unit = 2 * (pair0 + lp) + (row >> 6);
      const int ci_in = row & 63;
      float* dst = partial + (((long long)blockIdx.y * g.units + unit) * 64 + ci_in) * g.Cout + ntile * NT;
#pragma unroll 1
      for (int j = 0; j < NT / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(lp * NT + j * 32), v);
        tmem_ld_wait();
        if (unit < g.units) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(dst + j * 32 + c * 4) = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 512);
}

// dw[co][ci][tap] = sum_s partial[s][tap*cin_blocks + ci/64][ci%64][co]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int units,
                                    int cin_blocks, int Cin, int Cout) {
  const long long total = (long long)units * 64 * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(long long)s * total + i];
    const int co = (int)(i % Cout);
    const int ci_in = (int)((i / Cout) % 64);
    const int unit = (int)(i / ((long long)Cout * 64));
    const int tap = unit / cin_blocks, cb = unit - tap * cin_blocks;
    dw[((long long)co * Cin + cb * 64 + ci_in) * 27 + tap] = acc;
  }
}

// fp32 [Cout][Cin][27] -> bf16 wf[tap][Cout][Cin], wd[26-tap][Cin][Cout]
__global__ void pack_conv3_weights_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wf,
                                          __nv_bfloat16* __restrict__ wd) {
  const long long total = (long long)Cout * Cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 27);
    const int ci = (int)((i / 27) % Cin);
    const int co = (int)(i / (27ll * Cin));
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[((long long)tap * Cout + co) * Cin + ci] = v;
    if (wd) wd[((long long)(26 - tap) * Cin + ci) * Cout + co] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Pick the box (wt, ht, dt) of one voxel tile: rows = wt*ht*dt <= 128 (and a multiple of `row_mult`),
// maximising the fraction of useful rows over the whole volume, preferring long contiguous W runs.
static void pick_tile(int W, int H, int D, int row_mult, bool cost_per_row, int& wt, int& ht, int& dt) {
  double best = -1.0;
  wt = ht = dt = 1;
  for (int a = 1; a <= 128 && a <= W; ++a)
    for (int b = 1; a * b <= 128 && b <= H; ++b)
      for (int c = 1; a * b * c <= 128 && c <= D; ++c) {
        const int rows = a * b * c;
        if (rows % row_mult) continue;
        const double tiles = (double)cdiv(W, a) * cdiv(H, b) * cdiv(D, c);
        // fprop: every tile costs a full 128-row MMA; wgrad: cost follows the rows actually reduced
        // (plus a small per-box overhead).  Tie-break on wider W (longer contiguous TMA runs).
        const double cost = cost_per_row ? tiles * (((rows + 15) & ~15) + 8.0) : tiles * 128.0;
        const double score = ((double)W * H * D) / cost + 1e-6 * a + 1e-9 * b;
        if (score > best) {
          best = score;
          wt = a;
          ht = b;
          dt = c;
        }
      }
}

static int make_act_tmap(CUtensorMap* tm, const void* base, int N, int D, int H, int W, int C, int wt, int ht, int dt) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  uint32_t box[5] = {64, (uint32_t)wt, (uint32_t)ht, (uint32_t)dt, 1};
  return make_tmap_bf16(tm, base, 5, dims, strides, box);
}

template <int BLOCK_N, int STAGES, int EPI>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ConvGeom& g,
                        long long tiles, int nblocks, const ToOneEpilogue& ep, cudaStream_t st) {
  constexpr int smem = STAGES * (kTileBytes + BLOCK_N * 128) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(conv3_igemm_kernel<BLOCK_N, STAGES, EPI>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                   "cudaFuncSetAttribute(conv3_igemm)"))
      return -1;
    attr_set = true;
  }
  dim3 grid((unsigned)tiles, (unsigned)nblocks);
  conv3_igemm_kernel<BLOCK_N, STAGES, EPI><<<grid, 192, smem, st>>>(tmA, tmB, tmC, g, ep);
  SIVAE_LAUNCH_OK("conv3_igemm_kernel");
  return 0;
}

int conv3_igemm(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                cudaStream_t st) {
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64, "conv3_igemm: Cin=%d must be a multiple of 64", Cin);
  SIVAE_CHECK(Cout % 64 == 0 && Cout >= 64, "conv3_igemm: Cout=%d must be a multiple of 64", Cout);
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "conv3_igemm: empty tensor");
  ConvGeom g;
  g.N = N; g.D = D; g.H = H; g.W = W;
  pick_tile(W, H, D, 1, false, g.wt, g.ht, g.dt);
  g.tiles_w = cdiv(W, g.wt); g.tiles_h = cdiv(H, g.ht); g.tiles_d = cdiv(D, g.dt);
  g.rows = g.wt * g.ht * g.dt;
  g.cin_blocks = Cin / 64;
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "conv3_igemm: too many tiles");
  const int block_n = (Cout % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB, tmC;
  if (make_act_tmap(&tmA, x, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  if (make_act_tmap(&tmC, y, N, D, H, W, Cout, g.wt, g.ht, g.dt)) return -1;
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 27};
    uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    uint32_t box[3] = {64, (uint32_t)block_n, 1};
    if (make_tmap_bf16(&tmB, wpack, 3, dims, strides, box)) return -1;
  }
  const ToOneEpilogue ep{};
  if (block_n == 128) return launch_igemm<128, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 128, ep, st);
  return launch_igemm<64, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 64, ep, st);
}

// fp32 [C][27] (one output channel) -> bf16 [27][16][C]; row 0 = the filter (tap-flipped when `flip`), rows 1..15 zero
__global__ void pack_to1_weights_kernel(const float* __restrict__ w, int C, int flip, __nv_bfloat16* __restrict__ wp) {
  const int total = 27 * 16 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, r = (i / C) % 16, tap = i / (16 * C);
    wp[i] = __float2bfloat16_rn(r == 0 ? w[c * 27 + (flip ? 26 - tap : tap)] : 0.f);
  }
}

size_t conv3_to1_workspace_bytes(int C) { return (size_t)27 * 16 * C * sizeof(__nv_bfloat16); }

// y[v] = act(bias + sum_{tap,c} w[c][tap'] * x[v + delta(tap)][c]) on tcgen05 (N = 16 tile, column 0 used)
int conv3_to1(const void* x, const float* w, const float* bias, float* y, int N, int D, int H, int W, int C, int flip,
              int act, const uint8_t* mask, float p, unsigned long long seed, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  SIVAE_CHECK(C % 64 == 0 && C >= 64, "conv3_to1: C=%d must be a multiple of 64", C);
  SIVAE_CHECK(p >= 0.f && p < 1.f, "conv3_to1: dropout p=%f out of range", p);
  SIVAE_CHECK(ws && ws_bytes >= conv3_to1_workspace_bytes(C), "conv3_to1: workspace too small");
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "conv3_to1: empty tensor");
  pack_to1_weights_kernel<<<cdiv(27 * 16 * C, 256), 256, 0, st>>>(w, C, flip, (__nv_bfloat16*)ws);
  SIVAE_LAUNCH_OK("pack_to1_weights_kernel");
  ConvGeom g;
  g.N = N; g.D = D; g.H = H; g.W = W;
  pick_tile(W, H, D, 1, false, g.wt, g.ht, g.dt);
  g.tiles_w = cdiv(W, g.wt); g.tiles_h = cdiv(H, g.ht); g.tiles_d = cdiv(D, g.dt);
  g.rows = g.wt * g.ht * g.dt;
  g.cin_blocks = C / 64;
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "conv3_to1: too many tiles");
  CUtensorMap tmA, tmB;
  if (make_act_tmap(&tmA, x, N, D, H, W, C, g.wt, g.ht, g.dt)) return -1;
  {
    uint64_t dims[3] = {(uint64_t)C, 16, 27};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)16 * C * 2};
    uint32_t box[3] = {64, 16, 1};
    if (make_tmap_bf16(&tmB, ws, 3, dims, strides, box)) return -1;
  }
  ToOneEpilogue ep;
  ep.y = y; ep.bias = bias; ep.mask = mask; ep.seed = make_seed_ref(seed); ep.p = p; ep.act = act;
  return launch_igemm<16, 4, 1>(tmA, tmB, tmA, g, tiles, 1, ep, st);
}

// ---- wgrad ----
static void wgrad_plan(int N, int D, int H, int W, int Cin, int Cout, WgradGeom& g, int& nt, int& ntiles, int& splits) {
  g.N = N; g.D = D; g.H = H; g.W = W;
  pick_tile(W, H, D, 1, true, g.wt, g.ht, g.dt);
  g.tiles_w = cdiv(W, g.wt); g.tiles_h = cdiv(H, g.ht); g.tiles_d = cdiv(D, g.dt);
  g.rows = g.wt * g.ht * g.dt;
  g.cin_blocks = Cin / 64;
  g.units = 27 * g.cin_blocks;
  g.pairs_total = (g.units + 1) / 2;
  nt = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  ntiles = Cout / nt;
  g.ppc = 512 / nt;
  g.groups = cdiv(g.pairs_total, g.ppc);
  g.Cout = Cout;
  g.total_kb = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  int want = 148 / (g.groups * ntiles);
  if (want < 1) want = 1;
  if ((long long)want > g.total_kb) want = (int)g.total_kb;
  g.kb_per_split = (g.total_kb + want - 1) / want;
  splits = (int)((g.total_kb + g.kb_per_split - 1) / g.kb_per_split);
  g.a_lbo = kTileBytes; g.a_sbo = 1024; g.b_lbo = kTileBytes; g.b_sbo = 1024;
}

size_t conv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  if (Cin % 64 || Cout % 64 || N <= 0) return 0;
  WgradGeom g; int nt, ntiles, splits;
  wgrad_plan(N, D, H, W, Cin, Cout, g, nt, ntiles, splits);
  return (size_t)splits * g.units * 64 * Cout * sizeof(float);
}

template <int NT, int A_STAGES>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradGeom& g, int ntiles, int splits,
                        float* partial, cudaStream_t st) {
  constexpr int smem = 2 * (NT / 64) * kTileBytes + A_STAGES * 2 * kTileBytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(conv3_wgrad_kernel<NT, A_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem),
                   "cudaFuncSetAttribute(conv3_wgrad)"))
      return -1;
    attr_set = true;
  }
  dim3 grid((unsigned)(g.groups * ntiles), (unsigned)splits);
  conv3_wgrad_kernel<NT, A_STAGES><<<grid, 192, smem, st>>>(tmX, tmDY, g, partial);
  SIVAE_LAUNCH_OK("conv3_wgrad_kernel");
  return 0;
}

int conv3_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, int N, int D, int H, int W,
                int Cin, int Cout, cudaStream_t st) {
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64 && Cout % 64 == 0 && Cout >= 64,
              "conv3_wgrad: Cin=%d, Cout=%d must be multiples of 64", Cin, Cout);
  WgradGeom g; int nt, ntiles, splits;
  wgrad_plan(N, D, H, W, Cin, Cout, g, nt, ntiles, splits);
  const size_t need = (size_t)splits * g.units * 64 * Cout * sizeof(float);
  SIVAE_CHECK(ws != nullptr && ws_bytes >= need, "conv3_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  // debugging knob: override the MN-major descriptor strides "a_lbo,a_sbo,b_lbo,b_sbo"
  if (const char* e = getenv("SIVAE_WGRAD_DESC")) {
    unsigned a, b, c, d;
    if (sscanf(e, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) { g.a_lbo = a; g.a_sbo = b; g.b_lbo = c; g.b_sbo = d; }
  }
  CUtensorMap tmX, tmDY;
  if (make_act_tmap(&tmX, x, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  if (make_act_tmap(&tmDY, dy, N, D, H, W, Cout, g.wt, g.ht, g.dt)) return -1;
  int rc;
  if (nt == 256) rc = launch_wgrad<256, 2>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  else if (nt == 128) rc = launch_wgrad<128, 3>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  else rc = launch_wgrad<64, 4>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  if (rc) return rc;
  const long long total = (long long)g.units * 64 * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>((const float*)ws, dw, splits, g.units, g.cin_blocks, Cin, Cout);
  SIVAE_LAUNCH_OK("wgrad_reduce_kernel");
  return 0;
}

int pack_conv3_weights(const float* w, int Cout, int Cin, void* wf, void* wd, cudaStream_t st) {
  SIVAE_CHECK(Cout > 0 && Cin > 0, "pack_conv3_weights: bad dims");
  const long long total = (long long)Cout * Cin * 27;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_conv3_weights_kernel<<<blocks, 256, 0, st>>>(w, Cout, Cin, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd);
  SIVAE_LAUNCH_OK("pack_conv3_weights_kernel");
  return 0;
}

}  // namespace sivae
