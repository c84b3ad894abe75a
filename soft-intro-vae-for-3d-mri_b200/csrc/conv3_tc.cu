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
#include <utility>

#include "sivae_common.cuh"

namespace sivae {

// SM count of the current device (persistent kernels launch one CTA per SM); B200: 148.
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      n = 148;
  }
  return n;
}

// Optional (SIVAE_CONV_PRIORITY=1): launch the tensor-core convolution kernels at the highest stream priority (per-launch
// attribute, kept by CUDA graph capture), so that when a BatchNorm grid of the pass on the other stream and a persistent
// convolution are pending together the block scheduler places the convolution CTAs first and the BatchNorm blocks fill
// what is left of each SM.  Measured neutral (4 A/B pairs: 59.44 vs 59.70 ms, inside the run-to-run spread): the step is
// power-capped, more overlap mostly buys lower clocks.  Default off.
static int conv_priority(bool* use) {
  static int prio = 0, enabled = -1;
  if (enabled < 0) {
    int least = 0, greatest = 0;
    const char* e = getenv("SIVAE_CONV_PRIORITY");
    enabled = (e != nullptr && atoi(e) != 0) && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess &&
              greatest < least;
    prio = greatest;
  }
  *use = enabled != 0;
  return prio;
}
template <typename... KArgs, typename... Args>
static void launch_conv(void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  bool use = false;
  attr[0].id = cudaLaunchAttributePriority;
  attr[0].val.priority = conv_priority(&use);
  cfg.attrs = attr;
  cfg.numAttrs = use ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through SIVAE_LAUNCH_OK
}

static constexpr int kTileRows = 128;            // UMMA M (fprop) / max voxel rows per box
static constexpr int kTileBytes = kTileRows * 128;  // one 128-row x 64-channel bf16 tile

struct ConvGeom {
  int N, D, H, W;
  int wt, ht, dt;  // box extents of one tile of output voxels
  int tiles_w, tiles_h, tiles_d;
  int rows;        // wt*ht*dt  (<= 128)
  int cin_blocks;  // (GEMM-K channels) / 64
  int mode;        // kTapsPlain / kTapsUpFprop / kTapsUpDgrad
};

// Tap schedules of the implicit GEMM.
//   kTapsPlain   : 27 taps, offsets (kd-1, kh-1, kw-1), weights wpack[tap].
//   kTapsUpFprop : nearest-Upsample(2) folded into the convolution (nn.Upsample models/models.py:58 followed by
//                  Conv3d :59).  Output voxel (2d+pd, 2h+ph, 2w+pw) only sees 2x2x2 distinct low-resolution inputs, so
//                  each of the 8 output parities (blockIdx.z) is an 8-tap convolution on the low-res grid with
//                  pre-summed weights wpack[parity*8 + abc] and offsets (a-1+pd, b-1+ph, c-1+pw); it stores to the
//                  parity sub-lattice of the high-res output (tensor map tmC.m[parity]).  27 -> 8 MACs per output.
//   kTapsUpDgrad : the transpose: dX_low = sum over 8 parities x 8 taps of dY on the parity sub-lattice
//                  (tensor map tmA.m[parity]) at offsets -(a-1+pd), ... with weights wpack[parity*8 + abc] (transposed).
static constexpr int kTapsPlain = 0, kTapsUpFprop = 1, kTapsUpDgrad = 2;

struct TmapPack {
  CUtensorMap m[8];
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
conv3_igemm_kernel(const __grid_constant__ TmapPack tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ TmapPack tmC, const ConvGeom g, const ToOneEpilogue ep) {
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
  const int parity = blockIdx.z;   // kTapsUpFprop only (gridDim.z == 8), else 0
  const int ntaps = g.mode == kTapsPlain ? 27 : g.mode == kTapsUpFprop ? 8 : 64;
  const int num_kb = ntaps * g.cin_blocks;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA.m[0]);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC.m[parity]);
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
    // Nested tap / K-block loops with incremental stage and phase counters: no per-iteration divisions in the single
    // producing thread (its bookkeeping has to stay well below the ~400 cycles a K block's MMAs take).
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)g.rows * 128u + (uint32_t)B_BYTES;
      int s = 0;
      uint32_t ph = 0;
      auto issue = [&](const CUtensorMap* amap, int cb, int ow, int oh, int od, int wtap) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], tx_bytes);
        uint8_t* a_dst = smem + s * STAGE_BYTES;
        tma_load_5d(a_dst, amap, &full_bar[s], cb * 64, w0 + ow, h0 + oh, d0 + od, n);
        tma_load_3d(a_dst + kTileBytes, &tmB, &full_bar[s], cb * 64, nb * BLOCK_N, wtap);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      };
      if (g.mode == kTapsPlain) {
        int tap = 0;
        for (int od = -1; od <= 1; ++od)
          for (int oh = -1; oh <= 1; ++oh)
            for (int ow = -1; ow <= 1; ++ow, ++tap)
              for (int cb = 0; cb < g.cin_blocks; ++cb) issue(&tmA.m[0], cb, ow, oh, od, tap);
      } else if (g.mode == kTapsUpFprop) {
        const int pd = parity >> 2, php = (parity >> 1) & 1, pw = parity & 1;
        for (int tap = 0; tap < 8; ++tap)
          for (int cb = 0; cb < g.cin_blocks; ++cb)
            issue(&tmA.m[0], cb, (tap & 1) - 1 + pw, ((tap >> 1) & 1) - 1 + php, (tap >> 2) - 1 + pd, parity * 8 + tap);
      } else {
        for (int pp = 0; pp < 8; ++pp)
          for (int abc = 0; abc < 8; ++abc)
            for (int cb = 0; cb < g.cin_blocks; ++cb)
              issue(&tmA.m[pp], cb, -((abc & 1) - 1 + (pp & 1)), -(((abc >> 1) & 1) - 1 + ((pp >> 1) & 1)),
                    -((abc >> 2) - 1 + (pp >> 2)), pp * 8 + abc);
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: one elected thread runs the whole loop (a per-step elect + __syncwarp costs more issue slots
    // than the four UMMAs of a K block; the descriptors are base + stage * constant) =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, 1024);
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t adesc = desc0 + (uint64_t)(s * (uint32_t)(STAGE_BYTES >> 4));
        const uint64_t bdesc = adesc + (uint64_t)(kTileBytes >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 64-channel K block = 4 x UMMA_K(16); +32 bytes = +2 in the (address >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
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
        float r = (__uint_as_float(v[0]) + __uint_as_float(v[1])) + (ep.bias ? ep.bias[0] : 0.f);
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
      const uint32_t tile = smem_u32(out_stage + (j >> 1) * kTileBytes + row * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
        pk.y = pack_bf16x2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
        pk.z = pack_bf16x2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
        pk.w = pack_bf16x2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
        const int chunk = (j & 1) * 4 + c;
        sts128(tile + ((uint32_t)(chunk ^ (row & 7)) << 4), pk);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp_id == 2 && lane == 0) {
#pragma unroll
      for (int jb = 0; jb < BLOCK_N / 64; ++jb)
        tma_store_5d(&tmC.m[parity], out_stage + jb * kTileBytes, nb * BLOCK_N + jb * 64, w0, h0, d0, n);
      tma_store_commit();
      tma_store_wait_read_all();
    }
    }  // EPI == 0
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================================================================================
// fprop / dgrad, Cin = 64: persistent "kw-slab" kernel -- input tiles are reused across taps.
//
// The tap-by-tap kernel above re-fetches a 16 KB input tile and an 8 KB weight slab for each of the 27 taps of every
// 128-voxel tile (648 KB of L2 -> SMEM traffic per tile: L2-bandwidth bound, profiles/r01a).  Here one work item is a
// PAIR of output tiles -- the same 8(w) x 16(h) patch at depths d0 and d0+1 -- and for each kw shift the producer
// loads ONE tall box {64 ch, 8 w, 18 h, 4 d} (72 KB).  Because a w-run of 8 voxels is exactly one 1024-byte swizzle
// atom, the A operand of tap (kd, kh) of tile j is the same buffer at byte offset ((kd + j) * 18 + kh) * 1024: nine taps
// x two tiles are served by descriptor arithmetic alone, and each weight slab feeds both tiles.
//   L2 -> SMEM bytes per 128 outputs: 27*16 KB + 27*8 KB = 648 KB  ->  (3*72 KB + 27*8 KB) / 2 = 216 KB.
// The CTA is persistent (one per SM, static round-robin over work items) with everything double-buffered so that no
// role ever drains: two A boxes (the box of the next kw / next item streams in while the current one is multiplied),
// a deep ring of weight slabs, two TMEM accumulator stages (epilogue of item i overlaps the MMAs of item i+1) and two
// output staging tiles for the TMA stores.
// 7 warps: A producer, MMA issuer, B producer, 4 epilogue warps.
// =================================================================================================
struct KwGeom {
  int N, D, H, W;
  int tiles_w, tiles_h, tiles_d;   // patches of 8 x 16, depth pairs: tiles_d = ceil(D/2)
  int nblk;                        // Cout / 64
  long long items;                 // tiles_w * tiles_h * tiles_d * N * nblk
};

static constexpr int kKwW = 8, kKwH = 16;
static constexpr int kKwBoxRows = kKwW * (kKwH + 2) * 4;     // 576 voxel rows per tall box
static constexpr int kKwABytes = kKwBoxRows * 128;           // 73,728
static constexpr int kKwBStages = 6;
static constexpr int kKwBBytes = 64 * 128;                   // one 64 x 64 weight slab
static constexpr int kKwSmem = 2 * kKwABytes + kKwBStages * kKwBBytes + 2 * kTileBytes + 1024 + 256;

__global__ void __launch_bounds__(224, 1)
conv3_kw64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const KwGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + 2 * kKwABytes;
  uint8_t* smem_o = smem_b + kKwBStages * kKwBBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_o + 2 * kTileBytes);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kKwBStages;
  uint64_t* acc_full = b_empty + kKwBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);   // one arrival per epilogue warp
    }
    for (int s = 0; s < kKwBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 256);   // 2 stages x 2 tiles x 64 fp32 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // work item -> (output-channel block, patch origin, depth pair, sample); nb fastest so both halves of a wide Cout
  // hit the same input box in L2 back to back
  auto decode = [&](long long item, int& nb, int& w0, int& h0, int& d0, int& n) {
    nb = (int)(item % g.nblk); item /= g.nblk;
    const int tw = (int)(item % g.tiles_w); item /= g.tiles_w;
    const int th = (int)(item % g.tiles_h); item /= g.tiles_h;
    const int td = (int)(item % g.tiles_d);
    n = (int)(item / g.tiles_d);
    w0 = tw * kKwW; h0 = th * kKwH; d0 = td * 2;
  };

  if (warp_id == 0) {
    // ===== A producer: one tall box per kw shift, two buffers, running ahead across work items =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x) {
        int nb, w0, h0, d0, n;
        decode(item, nb, w0, h0, d0, n);
        for (int kw = 0; kw < 3; ++kw, ++it) {
          const int s = it & 1;
          mbar_wait(&a_empty[s], ((it >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&a_full[s], kKwABytes);
          tma_load_5d(smem + s * kKwABytes, &tmA, &a_full[s], 0, w0 + kw - 1, h0 - 1, d0 - 1, n);
        }
      }
    }
  } else if (warp_id == 2) {
    // ===== B producer: weight slab of tap (kd, kh, kw) in MMA order =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x) {
        const int nb = (int)(item % g.nblk);
        for (int kw = 0; kw < 3; ++kw)
          for (int kdh = 0; kdh < 9; ++kdh, ++it) {
            const int s = it % kKwBStages;
            mbar_wait(&b_empty[s], ((it / kKwBStages) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], kKwBBytes);
            tma_load_3d(smem_b + s * kKwBBytes, &tmB, &b_full[s], 0, nb * 64, kdh * 3 + kw);
          }
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    uint32_t a_it = 0, b_it = 0, acc_it = 0;
    for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
      const uint32_t as = acc_it & 1;
      mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 128u;
      for (int kw = 0; kw < 3; ++kw, ++a_it) {
        const uint32_t s = a_it & 1;
        mbar_wait(&a_full[s], (a_it >> 1) & 1u);
        const uint32_t a_base = smem_u32(smem + s * kKwABytes);
        for (int kdh = 0; kdh < 9; ++kdh, ++b_it) {
          const uint32_t bs = b_it % kKwBStages;
          mbar_wait(&b_full[bs], (b_it / kKwBStages) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const int kd = kdh / 3, kh = kdh - kd * 3;
            const uint32_t b_addr = smem_u32(smem_b + bs * kKwBBytes);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint32_t a_addr = a_base + (uint32_t)((kd + j) * (kKwH + 2) + kh) * 1024u;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem + (uint32_t)(j * 64), make_smem_desc(a_addr + k * 32, 16, 1024),
                          make_smem_desc(b_addr + k * 32, 16, 1024), idesc, (kw | kdh | k) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[bs]);
            if (kdh == 8) {
              umma_commit(&a_empty[s]);
              if (kw == 2) umma_commit(&acc_full[as]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue (warps 3..6 <-> TMEM lane quadrants 3,0,1,2): TMEM -> bf16 -> swizzled smem -> TMA store =====
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const bool issuer = (warp_id == 3 && lane == 0);
    uint32_t acc_it = 0, st_it = 0;
    for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
      int nb, w0, h0, d0, n;
      decode(item, nb, w0, h0, d0, n);
      const uint32_t as = acc_it & 1;
      mbar_wait(&acc_full[as], (acc_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j, ++st_it) {
        uint8_t* tile = smem_o + (st_it & 1) * kTileBytes;
        // the store issued two tiles ago read this staging buffer: it must have drained before it is overwritten
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 128u + (uint32_t)(j * 64);
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32u, v1);
        tmem_ld_wait();
        if (j == 1) {   // both tiles of this accumulator stage are in registers: hand the stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        uint8_t* dst = tile + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v0[c * 8 + 0]), __uint_as_float(v0[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v0[c * 8 + 2]), __uint_as_float(v0[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v0[c * 8 + 4]), __uint_as_float(v0[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v0[c * 8 + 6]), __uint_as_float(v0[c * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = pk;
          pk.x = pack_bf16x2(__uint_as_float(v1[c * 8 + 0]), __uint_as_float(v1[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v1[c * 8 + 2]), __uint_as_float(v1[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v1[c * 8 + 4]), __uint_as_float(v1[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v1[c * 8 + 6]), __uint_as_float(v1[c * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + (((4 + c) ^ (row & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (d0 + j < g.D) tma_store_5d(&tmC, tile, nb * 64, w0, h0, d0 + j, n);
          tma_store_commit();   // a (possibly empty) group per tile keeps the wait_group arithmetic uniform
        }
      }
    }
    if (issuer) tma_store_wait_read_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 256);
}

// =================================================================================================
// fprop / dgrad, Cin = Cout = 64: persistent "kd-fused" kernel -- three depth taps share one A read.
//
// Measured (profiles/r01c): with N = 64 the tcgen05 pipe is bound by SHARED-MEMORY OPERAND READS, not by L2 or the
// MMA floor.  One UMMA 128x64x16 reads 4 KB of A + 2 KB of B = 192 B per floor-cycle against a 128 B/cycle tensor-core
// read port (l1tex__data_pipe_tc_wavefronts_mem_shared at 64 % of peak when the tensor pipe is 43 % active) -- which is
// why halving the L2 traffic (conv3_kw64_kernel) bought nothing.  The fix is more MACs per operand byte:
//
// an input tile of plane p, shifted by (kh, kw), feeds output plane p+1 through tap kd=0, plane p through kd=1 and
// plane p-1 through kd=2, always on the SAME (h, w) rows = the same TMEM lanes.  So a CTA keeps the accumulators of
// P = 4 consecutive output planes side by side in TMEM, in DESCENDING plane order (plane i of the chunk at column
// 64*(P-1-i)), and issues ONE UMMA 128 x 192 x 16 per input tile and K step against the stacked weight slab
// [kd=0 | kd=1 | kd=2] (192 rows): D columns [c, c+192) are exactly the accumulators of planes p+1, p, p-1.
// A is read once for three taps: 4 KB + 6 KB per 3 x 128x64x16 MACs instead of 3 x 6 KB (1.8x fewer operand bytes);
// chunk-edge tiles use the N = 128 / 64 sub-slabs.  The very first contribution to every accumulator (tap kd=0 of
// pass (kh,kw)=(0,0)) is issued separately with accumulate = 0.
//   work item  = 8(w) x 16(h) patch x 4 planes;   9 (kh,kw) passes x 6 input tiles of 16 KB, 9 x 24 KB of weights.
//   L2 -> SMEM = (54*16 + 9*24) KB / 4 tiles = 270 KB per 128 outputs (tap-by-tap kernel: 648 KB).
// Persistent, one CTA per SM; A ring 8 x 16 KB, B ring 2 x 24 KB, two TMEM stages of 256 columns (the epilogue of
// item i overlaps the MMAs of item i+1), two output staging tiles for the TMA stores.
// 7 warps: A producer, MMA issuer, B producer, 4 epilogue warps.
// =================================================================================================
struct KdGeom {
  int N, D, H, W;
  int tiles_w, tiles_h, tiles_d;   // 8 x 16 patches, chunks of 4 planes
  int cin_blocks;                  // Cin / 64 (K blocks per tap)
  long long items;
};
static constexpr int kKdP = 4;                       // output planes per work item
static constexpr int kKdAStages = 8;
static constexpr int kKdBStages = 2;
static constexpr int kKdBBytes = 3 * 64 * 128;       // [kd=0 | kd=1 | kd=2] slabs of one (kh, kw)
static constexpr int kKdSmem = kKdAStages * kTileBytes + kKdBStages * kKdBBytes + 2 * kTileBytes + 1024 + 256;

__global__ void __launch_bounds__(224, 1)
conv3_kd3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const KdGeom g, float* __restrict__ stats_partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + kKdAStages * kTileBytes;
  uint8_t* smem_o = smem_b + kKdBStages * kKdBBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_o + 2 * kTileBytes);
  uint64_t* a_empty = a_full + kKdAStages;
  uint64_t* b_full = a_empty + kKdAStages;
  uint64_t* b_empty = b_full + kKdBStages;
  uint64_t* acc_full = b_empty + kKdBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < kKdAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kKdBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 512);   // 2 stages x 4 planes x 64 fp32 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](long long item, int& w0, int& h0, int& d0, int& n) {
    const int tw = (int)(item % g.tiles_w); item /= g.tiles_w;
    const int th = (int)(item % g.tiles_h); item /= g.tiles_h;
    const int td = (int)(item % g.tiles_d);
    n = (int)(item / g.tiles_d);
    w0 = tw * kKwW; h0 = th * kKwH; d0 = td * kKdP;
  };

  if (warp_id == 0) {
    // ===== A producer: per (kh, kw) pass the P + 2 input planes of the chunk, one 16 KB tile each =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x) {
        int w0, h0, d0, n;
        decode(item, w0, h0, d0, n);
        for (int pass = 0; pass < 9; ++pass) {
          const int kh = pass / 3, kw = pass - kh * 3;
          for (int cb = 0; cb < g.cin_blocks; ++cb)
            for (int t = 0; t < kKdP + 2; ++t, ++it) {
              const int s = it % kKdAStages;
              mbar_wait(&a_empty[s], ((it / kKdAStages) & 1u) ^ 1u);
              mbar_expect_tx(&a_full[s], kTileBytes);
              tma_load_5d(smem + s * kTileBytes, &tmA, &a_full[s], cb * 64, w0 + kw - 1, h0 + kh - 1, d0 - 1 + t, n);
            }
        }
      }
    }
  } else if (warp_id == 2) {
    // ===== B producer: the three kd slabs of one (kh, kw), stacked =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x)
        for (int pass = 0; pass < 9; ++pass)
          for (int cb = 0; cb < g.cin_blocks; ++cb, ++it) {
            const int s = it % kKdBStages;
            mbar_wait(&b_empty[s], ((it / kKdBStages) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], kKdBBytes);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
              tma_load_3d(smem_b + s * kKdBBytes + kd * (64 * 128), &tmB, &b_full[s], cb * 64, 0, kd * 9 + pass);
          }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: one elected thread for the whole loop =====
    // The tile loop is fully unrolled so that column offsets, slab offsets and instruction descriptors are immediates:
    // a single thread issues every UMMA, and its bookkeeping must stay well below the ~600 cycles a tile's MMAs take.
    // Ring slots / phases are counters with a wrap, descriptors are base + slot * constant, and there is no per-step
    // elect / __syncwarp (see upconv3_fused_kernel for the measurement behind this).
    if (elect_one()) {
      uint32_t sa = 0, pa = 0, bs = 0, pb = 0, acc_it = 0;
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 16, 1024);
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
        const uint32_t as = acc_it & 1;
        mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256u;
        for (int pc = 0; pc < 9 * g.cin_blocks; ++pc) {   // (pass, K block) pairs, K block fastest
          mbar_wait(&b_full[bs], pb);
          const uint64_t bdesc = bdesc0 + (uint64_t)(bs * (uint32_t)(kKdBBytes >> 4));
          const bool pass0 = (pc == 0);                            // first tap, first K block: overwrite
          const bool last = (pc == 9 * g.cin_blocks - 1);
#pragma unroll
          for (int t = 0; t < kKdP + 2; ++t) {
            mbar_wait(&a_full[sa], pa);
            tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(sa * (uint32_t)(kTileBytes >> 4));
            // input plane t of the chunk feeds output plane i = t - kd through tap kd, 0 <= i < P  (all compile time)
            constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0);
            const int kd_lo0 = t - (kKdP - 1) > 0 ? t - (kKdP - 1) : 0;
            const int kd_hi = t < 2 ? t : 2;
            int kd_lo = kd_lo0;
            if (pass0 && kd_lo0 == 0) {
              // first contribution to output plane i = t: overwrite
              const uint32_t col = (uint32_t)(64 * (kKdP - 1 - t));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem + col, adesc + 2 * k, bdesc + 2 * k, idesc64, k != 0 ? 1u : 0u);
              kd_lo = 1;
            }
            if (kd_lo <= kd_hi) {
              const int nun = kd_hi - kd_lo + 1;                                  // stacked taps: 1..3
              const uint32_t col = (uint32_t)(64 * (kKdP - 1 - (t - kd_lo)));    // highest output plane first
              const uint32_t idesc = nun == 3 ? make_idesc_bf16(128, 192, 0, 0)
                                     : nun == 2 ? make_idesc_bf16(128, 128, 0, 0) : idesc64;
              const uint64_t bd = bdesc + (uint64_t)(kd_lo * (64 * 128 / 16));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem + col, adesc + 2 * k, bd + 2 * k, idesc, 1u);
            }
            umma_commit(&a_empty[sa]);
            if (t == kKdP + 1) {
              umma_commit(&b_empty[bs]);
              if (last) umma_commit(&acc_full[as]);
            }
            if (++sa == kKdAStages) { sa = 0; pa ^= 1u; }
          }
          if (++bs == kKdBStages) { bs = 0; pb ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue (warps 3..6 <-> TMEM lane quadrants 3,0,1,2): TMEM -> bf16 -> swizzled smem -> TMA store =====
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const bool issuer = (warp_id == 3 && lane == 0);
    // fused BatchNorm statistics (stats_partial != NULL): every epilogue warp sums, per channel pair of its lane, the
    // bf16-rounded values of the 32 rows it staged -- i.e. statistics of the tensor exactly as stored
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    uint32_t acc_it = 0, st_it = 0;
    for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
      int w0, h0, d0, n;
      decode(item, w0, h0, d0, n);
      const uint32_t as = acc_it & 1;
      mbar_wait(&acc_full[as], (acc_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < kKdP; ++i, ++st_it) {
        uint8_t* tile = smem_o + (st_it & 1) * kTileBytes;
        // the store issued two tiles ago read this staging buffer: it must have drained before it is overwritten
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256u + (uint32_t)(64 * (kKdP - 1 - i));
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32u, v1);
        tmem_ld_wait();
        if (i == kKdP - 1) {   // the whole accumulator stage is in registers / staged: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        const uint32_t dst = smem_u32(tile) + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v0[c * 8 + 0]), __uint_as_float(v0[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v0[c * 8 + 2]), __uint_as_float(v0[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v0[c * 8 + 4]), __uint_as_float(v0[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v0[c * 8 + 6]), __uint_as_float(v0[c * 8 + 7]));
          sts128(dst + ((uint32_t)(c ^ (row & 7)) << 4), pk);
          pk.x = pack_bf16x2(__uint_as_float(v1[c * 8 + 0]), __uint_as_float(v1[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v1[c * 8 + 2]), __uint_as_float(v1[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v1[c * 8 + 4]), __uint_as_float(v1[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v1[c * 8 + 6]), __uint_as_float(v1[c * 8 + 7]));
          sts128(dst + ((uint32_t)((4 + c) ^ (row & 7)) << 4), pk);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (d0 + i < g.D) tma_store_5d(&tmC, tile, 0, w0, h0, d0 + i, n);
          tma_store_commit();   // a (possibly empty) group per tile keeps the wait_group arithmetic uniform
        }
        if (stats_partial != nullptr && d0 + i < g.D)
          tile_channel_sums<kKwW>(smem_u32(tile), q, lane, g.W - w0, g.H - h0, s1a, s1b, s2a, s2b);
      }
    }
    if (stats_partial != nullptr) {
      // partial[(blk * 2 + {0: sum, 1: sum of squares}) * 64 + c], blk = CTA * 4 + epilogue warp (bn finalize layout)
      float* dst = stats_partial + (size_t)(blockIdx.x * 4 + (warp_id - 3)) * 2 * 64;
      *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(s1a, s1b);
      *reinterpret_cast<float2*>(dst + 64 + 2 * lane) = make_float2(s2a, s2b);
    }
    if (issuer) tma_store_wait_read_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 512);
}

// =================================================================================================
// Upsample(2) + Conv3d fprop, Cout = 64: persistent kernel with shared-tile tap stacking.
//
// In the folded form (see kTapsUpFprop) output parity (pd,ph,pw) at low-res voxel v is an 8-tap convolution over the
// low-res inputs v + (ad-1+pd, ah-1+ph, aw-1+pw).  The tap-by-tap kernel treats the 8 parities x 8 taps as 64 separate
// (A tile, weight slab) pairs.  But a low-res tile shifted by (oh, ow) in the plane is shared by every (ph,ah) / (pw,aw)
// combination with ah-1+ph = oh and aw-1+pw = ow -- 4 units for (0,0), 2 for the edge shifts, 1 for the corners -- and
// all of them land on the same voxel rows = the same TMEM lanes.  A work item is therefore
//     8(w) x 16(h) low-res patch  x  one low-res plane d  x  one depth parity pd
// with the four (ph,pw) parity accumulators side by side in TMEM (column block pw*2 + ph); per depth tap ad it walks the
// 9 in-plane shifts of input plane d + ad - 1 + pd and issues
//     (0,0): two UMMAs N = 128 (all four parities)     (0,-1) / (0,+1): ONE UMMA N = 128 (the two ph of one pw)
//     (-1,0) / (+1,0): two UMMAs N = 64 (column blocks not adjacent)       corners: one UMMA N = 64
// = 12 A reads instead of 16 per input plane, and 18 TMA tile loads per item instead of 32.  (The (0,0) tile is issued as
// two N = 128 steps: a 32 KB weight slot would leave too few ring slots to hide the L2 latency of the weight stream.)
// The (0,0) steps of ad = 0 come first and overwrite all four accumulators (accumulate = 0).  The weight slabs of a step
// are TMA-loaded back to back into one 16 KB ring slot so the stacked operand is contiguous.
// Persistent, one CTA per SM: A ring 5 x 16 KB, B ring 7 x 16 KB, two TMEM stages x 256 columns, two staging tiles;
// the four parity tiles are stored through the strided parity-sub-lattice tensor maps; BatchNorm statistics fused as in
// conv3_kd3_kernel.  7 warps: A producer, MMA issuer, B producer, 4 epilogue warps.
// =================================================================================================
struct UpGeom {
  int N, D, H, W;                  // low-res lattice
  int tiles_w, tiles_h;
  int cin_blocks;
  long long items;                 // tiles_w * tiles_h * D * 2 * N
};
static constexpr int kUpAStages = 5, kUpBStages = 7;   // (6 B stages would leave room for a co-resident BatchNorm block: A/B neutral)
static constexpr int kUpBSlot = 2 * 64 * 128;       // up to two stacked 64 x 64 slabs
static constexpr int kUpSmem = kUpAStages * kTileBytes + kUpBStages * kUpBSlot + 2 * kTileBytes + 1024 + 256;
static constexpr int kUpSteps = 10;

// schedule of one depth tap: in-plane shift (oh, ow), its 1..2 units (ph, pw) (accumulator column block = pw * 2 + ph),
// stacked (one UMMA N = 128) or separate (two UMMAs N = 64), and whether the step loads a new A tile / is the last
// user of the current one.  The (0,0) tile feeds all four parities: two N = 128 steps on the same A tile (a 32 KB weight
// slot would leave too few slots to hide the L2 latency of the weight stream).
struct UpStep { int8_t oh, ow, n, stacked, new_a, rel_a; int8_t ph[2], pw[2]; };
#define SIVAE_UP_STEPS                                                                                   \
  {0, 0, 2, 1, 1, 0, {0, 1}, {0, 0}}, {0, 0, 2, 1, 0, 1, {0, 1}, {1, 1}}, {0, -1, 2, 1, 1, 1, {0, 1}, {0, 0}}, \
  {0, 1, 2, 1, 1, 1, {0, 1}, {1, 1}}, {-1, 0, 2, 0, 1, 1, {0, 0}, {0, 1}}, {1, 0, 2, 0, 1, 1, {1, 1}, {0, 1}}, \
  {-1, -1, 1, 1, 1, 1, {0, 0}, {0, 0}}, {-1, 1, 1, 1, 1, 1, {0, 0}, {1, 0}}, {1, -1, 1, 1, 1, 1, {1, 0}, {0, 0}}, \
  {1, 1, 1, 1, 1, 1, {1, 0}, {1, 0}}
// compile-time copy: the MMA issuer's step loop is fully unrolled so every field folds into an immediate (with run-time
// table look-ups the single issuing thread, not the tensor pipe, set the pace: ~1000 cycles of bookkeeping per step)
__host__ __device__ constexpr UpStep up_step_c(int t) {
  constexpr UpStep tab[kUpSteps] = {SIVAE_UP_STEPS};
  return tab[t];
}
__device__ __constant__ UpStep kUpStepTab[kUpSteps] = {SIVAE_UP_STEPS};

__global__ void __launch_bounds__(224, 1)
upconv3_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ TmapPack tmC, const UpGeom g, float* __restrict__ stats_partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + kUpAStages * kTileBytes;
  uint8_t* smem_o = smem_b + kUpBStages * kUpBSlot;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_o + 2 * kTileBytes);
  uint64_t* a_empty = a_full + kUpAStages;
  uint64_t* b_full = a_empty + kUpAStages;
  uint64_t* b_empty = b_full + kUpBStages;
  uint64_t* acc_full = b_empty + kUpBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int p = 0; p < 8; ++p) prefetch_tmap(&tmC.m[p]);
    for (int s = 0; s < kUpAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kUpBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](long long item, int& w0, int& h0, int& d, int& pd, int& n) {
    pd = (int)(item & 1); item >>= 1;
    const int tw = (int)(item % g.tiles_w); item /= g.tiles_w;
    const int th = (int)(item % g.tiles_h); item /= g.tiles_h;
    d = (int)(item % g.D);
    n = (int)(item / g.D);
    w0 = tw * kKwW; h0 = th * kKwH;
  };

  if (warp_id == 0) {
    // ===== A producer: per depth tap the 9 shifted low-res tiles (x cin_blocks K blocks) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x) {
        int w0, h0, d, pd, n;
        decode(item, w0, h0, d, pd, n);
        for (int cb = 0; cb < g.cin_blocks; ++cb)      // K block outermost: every (tile, K block) is one A stage
          for (int ad = 0; ad < 2; ++ad) {
#pragma unroll
            for (int t = 0; t < kUpSteps; ++t) {         // unrolled: the schedule folds into immediates
              if (!up_step_c(t).new_a) continue;
              const int s = it % kUpAStages;
              mbar_wait(&a_empty[s], ((it / kUpAStages) & 1u) ^ 1u);
              mbar_expect_tx(&a_full[s], kTileBytes);
              tma_load_5d(smem + s * kTileBytes, &tmA, &a_full[s], cb * 64, w0 + up_step_c(t).ow, h0 + up_step_c(t).oh,
                          d + ad - 1 + pd, n);
              ++it;
            }
          }
      }
    }
  } else if (warp_id == 2) {
    // ===== B producer: the 1..4 weight slabs of a shifted tile, stacked in unit order =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x) {
        const int pd = (int)(item & 1);
        for (int cb = 0; cb < g.cin_blocks; ++cb)
          for (int ad = 0; ad < 2; ++ad) {
#pragma unroll
            for (int t = 0; t < kUpSteps; ++t, ++it) {   // unrolled: slab indices become base + immediate
              const int s = it % kUpBStages;
              mbar_wait(&b_empty[s], ((it / kUpBStages) & 1u) ^ 1u);
              mbar_expect_tx(&b_full[s], (uint32_t)up_step_c(t).n * (64u * 128u));
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (u >= up_step_c(t).n) continue;
                const int ph = up_step_c(t).ph[u], pw = up_step_c(t).pw[u];
                const int ah = up_step_c(t).oh + 1 - ph, aw = up_step_c(t).ow + 1 - pw;
                const int slab = (pd * 4 + ph * 2 + pw) * 8 + ad * 4 + ah * 2 + aw;
                tma_load_3d(smem_b + s * kUpBSlot + u * (64 * 128), &tmB, &b_full[s], cb * 64, 0, slab);
              }
            }
          }
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: ONE thread for the whole loop =====
    // The issue loop is a single dependent instruction stream: at ~85 SASS instructions per 4-UMMA step (a warp-wide
    // elect + reconvergence, ring indices by division, descriptors rebuilt from shared-memory addresses) it ran at
    // ~680 cycles per step against 256 cycles of tensor work (ncu: tensor pipe 37 % active, the issuing warp never
    // waiting on a barrier).  So: no per-step elect / __syncwarp, ring slot + phase kept as counters with a wrap,
    // descriptors = base + slot * constant.  elect.sync (not lane == 0): only then does the compiler treat the region
    // as uniform and feed UTCHMMA / UTCBAR from uniform registers; under a lane test it wraps EVERY tcgen05
    // instruction in its own elect loop (~9 extra instructions each).
    if (elect_one()) {
      uint32_t sa_next = 0, pa = 0, sa = 0;          // A ring: next slot to take, its phase, slot in use
      uint32_t sb = 0, pb = 0;                       // B ring
      uint32_t acc_it = 0;
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 16, 1024);
      for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
        const uint32_t as = acc_it & 1;
        mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256u;
        for (int cb = 0; cb < g.cin_blocks; ++cb)
          for (int ad = 0; ad < 2; ++ad) {
            const uint32_t first = (cb == 0 && ad == 0) ? 0u : 1u;   // steps 0, 1 of the first tap overwrite all 4 parities
            const bool last = (cb == g.cin_blocks - 1 && ad == 1);
#pragma unroll
            for (int t = 0; t < kUpSteps; ++t) {
              constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0), idesc128 = make_idesc_bf16(128, 128, 0, 0);
              if (up_step_c(t).new_a) {
                sa = sa_next;
                mbar_wait(&a_full[sa], pa);
                if (++sa_next == kUpAStages) { sa_next = 0; pa ^= 1u; }
              }
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint64_t adesc = adesc0 + (uint64_t)(sa * (uint32_t)(kTileBytes >> 4));
              const uint64_t bdesc = bdesc0 + (uint64_t)(sb * (uint32_t)(kUpBSlot >> 4));
              if (up_step_c(t).stacked) {
                const uint32_t col = (uint32_t)(up_step_c(t).pw[0] * 2 + up_step_c(t).ph[0]) * 64u;
#pragma unroll
                for (int k = 0; k < 4; ++k)   // +32 bytes per K step = +2 in the descriptor's (address >> 4) field
                  umma_bf16(d_tmem + col, adesc + 2 * k, bdesc + 2 * k, up_step_c(t).n == 2 ? idesc128 : idesc64,
                            (k == 0 && t < 2) ? first : 1u);
              } else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const uint32_t col = (uint32_t)(up_step_c(t).pw[u] * 2 + up_step_c(t).ph[u]) * 64u;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(d_tmem + col, adesc + 2 * k, bdesc + (uint64_t)(u * (64 * 128 / 16) + 2 * k), idesc64, 1u);
                }
              }
              if (up_step_c(t).rel_a) umma_commit(&a_empty[sa]);
              umma_commit(&b_empty[sb]);
              if (last && t == kUpSteps - 1) umma_commit(&acc_full[as]);
              if (++sb == kUpBStages) { sb = 0; pb ^= 1u; }
            }
          }
      }
    }
  } else {
    // ===== epilogue (warps 3..6 <-> TMEM lane quadrants 3,0,1,2): four parity tiles per item =====
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const bool issuer = (warp_id == 3 && lane == 0);
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    uint32_t acc_it = 0, st_it = 0;
    for (long long item = blockIdx.x; item < g.items; item += gridDim.x, ++acc_it) {
      int w0, h0, d, pd, n;
      decode(item, w0, h0, d, pd, n);
      const uint32_t as = acc_it & 1;
      mbar_wait(&acc_full[as], (acc_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j, ++st_it) {          // column block j = pw * 2 + ph
        uint8_t* tile = smem_o + (st_it & 1) * kTileBytes;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256u + (uint32_t)(j * 64);
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32u, v1);
        tmem_ld_wait();
        if (j == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        const uint32_t dst = smem_u32(tile) + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v0[c * 8 + 0]), __uint_as_float(v0[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v0[c * 8 + 2]), __uint_as_float(v0[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v0[c * 8 + 4]), __uint_as_float(v0[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v0[c * 8 + 6]), __uint_as_float(v0[c * 8 + 7]));
          sts128(dst + ((uint32_t)(c ^ (row & 7)) << 4), pk);
          pk.x = pack_bf16x2(__uint_as_float(v1[c * 8 + 0]), __uint_as_float(v1[c * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v1[c * 8 + 2]), __uint_as_float(v1[c * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v1[c * 8 + 4]), __uint_as_float(v1[c * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v1[c * 8 + 6]), __uint_as_float(v1[c * 8 + 7]));
          sts128(dst + ((uint32_t)((4 + c) ^ (row & 7)) << 4), pk);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          const int parity = pd * 4 + (j & 1) * 2 + (j >> 1);
          tma_store_5d(&tmC.m[parity], tile, 0, w0, h0, d, n);
          tma_store_commit();
        }
        if (stats_partial != nullptr)
          tile_channel_sums<kKwW>(smem_u32(tile), q, lane, g.W - w0, g.H - h0, s1a, s1b, s2a, s2b);
      }
    }
    if (stats_partial != nullptr) {
      float* dst = stats_partial + (size_t)(blockIdx.x * 4 + (warp_id - 3)) * 2 * 64;
      *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(s1a, s1b);
      *reinterpret_cast<float2*>(dst + 64 + 2 * lane) = make_float2(s2a, s2b);
    }
    if (issuer) tma_store_wait_read_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 512);
}

// =================================================================================================
// wgrad
// =================================================================================================
struct WgradGeom {
  int N, D, H, W;       // lattice of the GEMM-K voxels (low-res grid in the upsample-fused mode)
  int wt, ht, dt, tiles_w, tiles_h, tiles_d;
  int rows;             // voxel rows per box (<= 128); consumed in 16-row K steps, tail rows zeroed
  int cin_blocks;       // Cin/64
  int mode;             // kTapsPlain, or kTapsUpFprop (upsample folded in: 8 parity classes x 8 taps)
  int nclasses;         // 1 (plain) or 8 (output parities); all units of a CTA share the dy operand of one class
  int units_pc;         // units per class: taps(27|8) * cin_blocks   (one unit = one tap x one 64-wide Cin block)
  int pairs_pc;         // ceil(units_pc/2)
  int ppc;              // unit pairs accumulated per CTA (TMEM: ppc * NT columns)
  int groups_pc;        // ceil(pairs_pc / ppc)
  int units;            // nclasses * units_pc
  int Cout;
  long long total_kb;   // voxel boxes in the whole lattice
  long long kb_per_split;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // descriptor strides (bytes), see make_smem_desc
};

template <int NT, int A_STAGES>
__global__ void __launch_bounds__(192, 1)
conv3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ TmapPack tmDY,
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
  const int groups_all = g.groups_pc * g.nclasses;
  const int grp_all = blockIdx.x % groups_all;
  const int ntile = blockIdx.x / groups_all;
  const int cls = grp_all / g.groups_pc;
  const int pair0 = (grp_all % g.groups_pc) * g.ppc;
  const int npairs = min(g.ppc, g.pairs_pc - pair0);
  const long long kb0 = (long long)blockIdx.y * g.kb_per_split;
  const long long kb1 = min(g.total_kb, kb0 + g.kb_per_split);

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY.m[cls]);
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
    if (elect_one()) {
      uint32_t a_it = 0, b_it = 0;
      // voxel-box coordinates advance incrementally (one 64-bit decode per CTA, none per box)
      int w0, h0, d0, n;
      decode_tile(g, kb0, w0, h0, d0, n);
      for (long long kb = kb0; kb < kb1; ++kb) {
        if (kb != kb0) {
          w0 += g.wt;
          if (w0 >= g.tiles_w * g.wt) {
            w0 = 0; h0 += g.ht;
            if (h0 >= g.tiles_h * g.ht) {
              h0 = 0; d0 += g.dt;
              if (d0 >= g.tiles_d * g.dt) { d0 = 0; ++n; }
            }
          }
        }
        {
          const int s = b_it % B_STAGES;
          mbar_wait(&b_empty[s], ((b_it / B_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(&b_full[s], tile_tx * (NT / 64));
#pragma unroll
          for (int jc = 0; jc < NT / 64; ++jc)
            tma_load_5d(smem_b + s * B_STAGE_BYTES + jc * kTileBytes, &tmDY.m[cls], &b_full[s], ntile * NT + jc * 64, w0,
                        h0, d0, n);
          ++b_it;
        }
        for (int lp = 0; lp < npairs; ++lp) {
          const int s = a_it % A_STAGES;
          mbar_wait(&a_empty[s], ((a_it / A_STAGES) & 1u) ^ 1u);
          const int u0 = 2 * (pair0 + lp);
          const int nu = (u0 + 1 < g.units_pc) ? 2 : 1;
          mbar_expect_tx(&a_full[s], tile_tx * nu);
          for (int uu = 0; uu < nu; ++uu) {
            const int u = u0 + uu;
            const int tap = g.cin_blocks == 1 ? u : u / g.cin_blocks, cb = u - tap * g.cin_blocks;
            int od, oh, ow;
            if (g.mode == kTapsPlain) {
              od = tap / 9 - 1; oh = (tap / 3) % 3 - 1; ow = tap % 3 - 1;
            } else {
              od = (tap >> 2) - 1 + (cls >> 2); oh = ((tap >> 1) & 1) - 1 + ((cls >> 1) & 1); ow = (tap & 1) - 1 + (cls & 1);
            }
            tma_load_5d(smem_a + s * A_STAGE_BYTES + uu * kTileBytes, &tmX, &a_full[s], cb * 64, w0 + ow, h0 + oh,
                        d0 + od, n);
          }
          ++a_it;
        }
      }
    }
  } else if (warp_id == 1) {
    // one elected thread issues the whole loop (no per-step elect / __syncwarp; ring counters with a wrap; descriptors
    // = base + stage * constant)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
      const int k16s = rows_pad / 16;
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), g.a_lbo, g.a_sbo);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), g.b_lbo, g.b_sbo);
      uint32_t as = 0, pa = 0, bs = 0, pb = 0;
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&b_full[bs], pb);
        const uint64_t bdesc = bdesc0 + (uint64_t)(bs * (uint32_t)(B_STAGE_BYTES >> 4));
        for (int lp = 0; lp < npairs; ++lp) {
          mbar_wait(&a_full[as], pa);
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(as * (uint32_t)(A_STAGE_BYTES >> 4));
          for (int k = 0; k < k16s; ++k)  // 16 voxel rows (2048 B = +128 in the descriptor's address field) per UMMA
            umma_bf16(tmem_base + (uint32_t)(lp * NT), adesc + 128 * k, bdesc + 128 * k, idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[as]);
          if (lp == npairs - 1) {
            umma_commit(&b_empty[bs]);
            if (kb == kb1 - 1) umma_commit(tmem_full_bar);
          }
          if (++as == A_STAGES) { as = 0; pa ^= 1u; }
        }
        if (++bs == B_STAGES) { bs = 0; pb ^= 1u; }
      }
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    for (int lp = 0; lp < npairs; ++lp) {
      const int ul = 2 * (pair0 + lp) + (row >> 6);          // unit within the class
      const int unit = cls * g.units_pc + ul;
      const int ci_in = row & 63;
      float* dst = partial + (((long long)blockIdx.y * g.units + unit) * 64 + ci_in) * g.Cout + ntile * NT;
#pragma unroll 1
      for (int j = 0; j < NT / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(lp * NT + j * 32), v);
        tmem_ld_wait();
        if (ul < g.units_pc) {
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

// =================================================================================================
// wgrad, Cin = 64: persistent "kw-slab" split-K kernel (the wgrad twin of conv3_kw64_kernel).
//
// The generic kernel above fetches one tap-shifted 16 KB input tile per (tap, voxel box): 27 x 16 KB + dy per 128 voxels,
// and with one CTA per SM its 4-deep ring cannot keep enough bytes in flight (511 TFLOP/s on 64->64 @ 8x80x96x80).
// Here a CTA owns ONE kw shift (blockIdx.x % 3) and a contiguous range of work items; an item is the pair of 8 x 16
// voxel patches at depths d0, d0+1.  Per item it loads one tall input box {64 ch, 8 w, 18 h, 4 d} (72 KB) and the two
// dy tiles; the 9 (kd, kh) taps x 2 tiles are MN-major A operands at byte offset ((kd + j) * 18 + kh) * 1024 inside the
// box.  Units are paired into M = 128 MMAs through the descriptor's leading-dimension offset:
//     acc 0..2: (kd,0)+(kd,1)  LBO = 1 patch row;   acc 3: (0,2)+(1,2)  LBO = 18 patch rows;
//     acc 4   : (1,2)+(2,2)    (first half is a duplicate and is dropped by the epilogue).
// 5 accumulators x 64 fp32 columns stay in TMEM across the whole K range; fp32 partials go to the same
// [split][tap][ci][Cout] workspace layout as the generic kernel, so wgrad_reduce_kernel finishes the job.
//   L2 -> SMEM bytes per 128 voxels: 27*16 + 2*16 = 464 KB  ->  3 * (72 + 32) / 2 = 156 KB.
// =================================================================================================
struct WgKwGeom {
  int N, D, H, W;
  int tiles_w, tiles_h, tiles_d;
  int ntiles;                 // Cout / 64
  int Cout;
  long long items, items_per_split;
};
static constexpr int kWgKwBStages = 4;
static constexpr int kWgKwSmem = 2 * kKwABytes + kWgKwBStages * kTileBytes + 1024 + 256;

__global__ void __launch_bounds__(224, 1)   // 7 warps: X producer, MMA issuer, dy producer, 4 epilogue warps
conv3_wgrad_kw64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                        const WgKwGeom g, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + 2 * kKwABytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + kWgKwBStages * kTileBytes);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kWgKwBStages;
  uint64_t* tmem_full_bar = b_empty + kWgKwBStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kw = blockIdx.x % 3;
  const int ntile = blockIdx.x / 3;
  const long long it0 = (long long)blockIdx.y * g.items_per_split;
  const long long it1 = min(g.items, it0 + g.items_per_split);

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kWgKwBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](long long item, int& w0, int& h0, int& d0, int& n) {
    const int tw = (int)(item % g.tiles_w); item /= g.tiles_w;
    const int th = (int)(item % g.tiles_h); item /= g.tiles_h;
    const int td = (int)(item % g.tiles_d);
    n = (int)(item / g.tiles_d);
    w0 = tw * kKwW; h0 = th * kKwH; d0 = td * 2;
  };

  if (warp_id == 0) {
    // ===== X producer: the tall box of this CTA's kw shift =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = it0; item < it1; ++item, ++it) {
        int w0, h0, d0, n;
        decode(item, w0, h0, d0, n);
        const int s = it & 1;
        mbar_wait(&a_empty[s], ((it >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&a_full[s], kKwABytes);
        tma_load_5d(smem + s * kKwABytes, &tmX, &a_full[s], 0, w0 + kw - 1, h0 - 1, d0 - 1, n);
      }
    }
  } else if (warp_id == 2) {
    // ===== dy producer: the two output-gradient tiles of the item (depth d0 + 1 may be out of range: zero fill) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = it0; item < it1; ++item) {
        int w0, h0, d0, n;
        decode(item, w0, h0, d0, n);
        for (int j = 0; j < 2; ++j, ++it) {
          const int s = it % kWgKwBStages;
          mbar_wait(&b_empty[s], ((it / kWgKwBStages) & 1u) ^ 1u);
          mbar_expect_tx(&b_full[s], kTileBytes);
          tma_load_5d(smem_b + s * kTileBytes, &tmDY, &b_full[s], ntile * 64, w0, h0, d0 + j, n);
        }
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: one elected thread for the whole loop =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), kTileBytes, 1024);
      uint32_t a_it = 0, bs = 0, pb = 0;
      for (long long item = it0; item < it1; ++item, ++a_it) {
        const uint32_t s = a_it & 1;
        mbar_wait(&a_full[s], (a_it >> 1) & 1u);
        const uint32_t a_base = smem_u32(smem + s * kKwABytes);
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&b_full[bs], pb);
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + (uint64_t)(bs * (uint32_t)(kTileBytes >> 4));
          const bool first = (item == it0 && j == 0);
#pragma unroll
          for (int acc = 0; acc < 5; ++acc) {
            // base unit (kd, kh) and the pair's leading-dimension offset (in 1024-byte patch rows)
            const int kd = acc < 3 ? acc : acc - 3;
            const int kh = acc < 3 ? 0 : 2;
            const uint32_t lbo = acc < 3 ? 1024u : (uint32_t)(kKwH + 2) * 1024u;
            const uint64_t adesc = make_smem_desc(a_base + (uint32_t)((kd + j) * (kKwH + 2) + kh) * 1024u, lbo, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k)   // 16 voxel rows (2048 B = +128 in the descriptor's address field) per UMMA
              umma_bf16(tmem_base + (uint32_t)(acc * 64), adesc + 128 * k, bdesc + 128 * k, idesc,
                        (first && k == 0) ? 0u : 1u);
          }
          umma_commit(&b_empty[bs]);
          if (j == 1) {
            umma_commit(&a_empty[s]);
            if (item == it1 - 1) umma_commit(tmem_full_bar);
          }
          if (++bs == kWgKwBStages) { bs = 0; pb ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> fp32 partials [split][tap][ci][Cout] =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const int half = row >> 6, ci = row & 63;
#pragma unroll 1
    for (int acc = 0; acc < 5; ++acc) {
      int kd, kh;
      if (acc < 3) { kd = acc; kh = half; }
      else if (acc == 3) { kd = half; kh = 2; }
      else { kd = 1 + half; kh = 2; }
      const int tap = (kd * 3 + kh) * 3 + kw;
      float* dst = partial + (((long long)blockIdx.y * 27 + tap) * 64 + ci) * g.Cout + ntile * 64;
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + j * 32), v);
        tmem_ld_wait();
        if (!(acc == 4 && half == 0)) {
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

// =================================================================================================
// wgrad of Upsample(2) + Conv3d, Cin = 64: persistent tall-box split-K kernel (twin of conv3_wgrad_kw64_kernel).
//
// dWup[p*8 + abc][ci][co] = sum_v x_lo[v + (ad-1+pd, ah-1+ph, aw-1+pw)][ci] * dy_hi[2v + p][co]   (then
// wgrad_reduce_up_kernel folds the 64 slabs back onto the 27 taps).  A CTA owns one (ph, pw, aw) combination -- i.e. one
// in-plane w shift aw-1+pw and the two h shifts ah-1+ph -- and a contiguous range of work items (8 x 16 low-res patch x 2
// planes).  Per item: ONE tall low-res box {64 ch, 8 w, 18 h, 4 d} (72 KB) and the four dy parity tiles (pd x plane).
// The unit (pd, ad, ah) of plane j reads the box at byte offset ((j + ad + pd) * 18 + ah + ph) * 1024; ah = 0 / 1 are
// paired into one M = 128 UMMA through the leading-dimension offset (1 patch row), so four accumulators (pd, ad) x 64
// fp32 columns stay in TMEM over the whole K range.  Partials use the generic [split][p*8+abc][ci][Cout] layout.
//   L2 -> SMEM bytes per 128 low-res voxels (all 64 slabs): 64*16 + 8*16 = 1152 KB  ->  8 * (72 + 64) / 2 = 544 KB.
// =================================================================================================
struct WgUpGeom {
  int N, D, H, W;                  // low-res lattice
  int tiles_w, tiles_h, tiles_d;
  int ntiles;                      // Cout / 64
  int Cout;
  long long items, items_per_split;
};

__global__ void __launch_bounds__(224, 1)   // 7 warps: X producer, MMA issuer, dy producer, 4 epilogue warps
upconv3_wgrad_tall_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ TmapPack tmDY,
                          const WgUpGeom g, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + 2 * kKwABytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + kWgKwBStages * kTileBytes);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kWgKwBStages;
  uint64_t* tmem_full_bar = b_empty + kWgKwBStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int combo = blockIdx.x & 7;              // (ph, pw, aw)
  const int ntile = blockIdx.x >> 3;
  const int ph = combo >> 2, pw = (combo >> 1) & 1, aw = combo & 1;
  const long long it0 = (long long)blockIdx.y * g.items_per_split;
  const long long it1 = min(g.items, it0 + g.items_per_split);

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY.m[ph * 2 + pw]);
    prefetch_tmap(&tmDY.m[4 + ph * 2 + pw]);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kWgKwBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](long long item, int& w0, int& h0, int& d0, int& n) {
    const int tw = (int)(item % g.tiles_w); item /= g.tiles_w;
    const int th = (int)(item % g.tiles_h); item /= g.tiles_h;
    const int td = (int)(item % g.tiles_d);
    n = (int)(item / g.tiles_d);
    w0 = tw * kKwW; h0 = th * kKwH; d0 = td * 2;
  };

  if (warp_id == 0) {
    // ===== X producer: the tall low-res box of this CTA's w shift =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = it0; item < it1; ++item, ++it) {
        int w0, h0, d0, n;
        decode(item, w0, h0, d0, n);
        const int s = it & 1;
        mbar_wait(&a_empty[s], ((it >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&a_full[s], kKwABytes);
        tma_load_5d(smem + s * kKwABytes, &tmX, &a_full[s], 0, w0 + aw - 1 + pw, h0 - 1, d0 - 1, n);
      }
    }
  } else if (warp_id == 2) {
    // ===== dy producer: parity tiles (pd, ph, pw) of planes d0, d0 + 1 (plane d0 + 1 may be out of range: zero fill) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (long long item = it0; item < it1; ++item) {
        int w0, h0, d0, n;
        decode(item, w0, h0, d0, n);
        for (int j = 0; j < 2; ++j)
          for (int pd = 0; pd < 2; ++pd, ++it) {
            const int s = it % kWgKwBStages;
            mbar_wait(&b_empty[s], ((it / kWgKwBStages) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], kTileBytes);
            tma_load_5d(smem_b + s * kTileBytes, &tmDY.m[pd * 4 + ph * 2 + pw], &b_full[s], ntile * 64, w0, h0, d0 + j, n);
          }
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: one elected thread for the whole loop =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), kTileBytes, 1024);
      uint32_t a_it = 0, bs = 0, pb = 0;
      for (long long item = it0; item < it1; ++item, ++a_it) {
        const uint32_t s = a_it & 1;
        mbar_wait(&a_full[s], (a_it >> 1) & 1u);
        const uint32_t a_base = smem_u32(smem + s * kKwABytes);
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {                     // (plane j, depth parity pd), pd fastest
          const int j = jp >> 1, pd = jp & 1;
          mbar_wait(&b_full[bs], pb);
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + (uint64_t)(bs * (uint32_t)(kTileBytes >> 4));
          const bool first = (item == it0 && j == 0);      // first contribution to the accumulators of this pd
#pragma unroll
          for (int ad = 0; ad < 2; ++ad) {
            // units (pd, ad, ah = 0 | 1): box rows ((j + ad + pd) * 18 + ph + ah) * 8, paired through LBO = one patch row
            const uint64_t adesc = make_smem_desc(a_base + (uint32_t)((j + ad + pd) * (kKwH + 2) + ph) * 1024u, 1024, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k)   // 16 voxel rows (2048 B = +128 in the descriptor's address field) per UMMA
              umma_bf16(tmem_base + (uint32_t)((pd * 2 + ad) * 64), adesc + 128 * k, bdesc + 128 * k, idesc,
                        (first && k == 0) ? 0u : 1u);
          }
          umma_commit(&b_empty[bs]);
          if (jp == 3) {
            umma_commit(&a_empty[s]);
            if (item == it1 - 1) umma_commit(tmem_full_bar);
          }
          if (++bs == kWgKwBStages) { bs = 0; pb ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> fp32 partials [split][p*8 + abc][ci][Cout] =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const int ah = row >> 6, ci = row & 63;
#pragma unroll 1
    for (int acc = 0; acc < 4; ++acc) {
      const int pd = acc >> 1, ad = acc & 1;
      const int unit = (pd * 4 + ph * 2 + pw) * 8 + ad * 4 + ah * 2 + aw;
      float* dst = partial + (((long long)blockIdx.y * 64 + unit) * 64 + ci) * g.Cout + ntile * 64;
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + jj * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dst + jj * 32 + c * 4) = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 256);
}

// upsample-fused mode: dw[co][ci][k] = sum_s sum_parity partial[s][(parity*8 + abc(k,parity))*cin_blocks + ci/64][ci%64][co]
// (the transpose of the weight pre-summation: every 3x3x3 tap receives exactly one contribution per output parity)
__global__ void wgrad_reduce_up_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits,
                                       int cin_blocks, int Cin, int Cout) {
  const long long per_split = 64ll * cin_blocks * 64 * Cout;
  const long long total = 27ll * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int k = (int)(i / ((long long)Cout * Cin));
    const int kd = k / 9, kh = (k / 3) % 3, kw = k % 3;
    const int cb = ci >> 6, ci_in = ci & 63;
    float acc = 0.f;
    for (int p = 0; p < 8; ++p) {
      const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
      const int a = pd == 0 ? (kd == 0 ? 0 : 1) : (kd == 2 ? 1 : 0);
      const int b = ph == 0 ? (kh == 0 ? 0 : 1) : (kh == 2 ? 1 : 0);
      const int c = pw == 0 ? (kw == 0 ? 0 : 1) : (kw == 2 ? 1 : 0);
      const int unit = (p * 8 + (a << 2 | b << 1 | c)) * cin_blocks + cb;
      const float* src = partial + ((long long)unit * 64 + ci_in) * Cout + co;
      for (int s = 0; s < splits; ++s) acc += src[(long long)s * per_split];
    }
    dw[((long long)co * Cin + ci) * 27 + k] = acc;
  }
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
static int launch_igemm(const TmapPack& tmA, const CUtensorMap& tmB, const TmapPack& tmC, const ConvGeom& g,
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
  dim3 grid((unsigned)tiles, (unsigned)nblocks, g.mode == kTapsUpFprop ? 8u : 1u);
  launch_conv(conv3_igemm_kernel<BLOCK_N, STAGES, EPI>, grid, 192, smem, st, tmA, tmB, tmC, g, ep);
  SIVAE_LAUNCH_OK("conv3_igemm_kernel");
  return 0;
}

static void fill_geom(ConvGeom& g, int N, int D, int H, int W, int kch, int mode) {
  g.N = N; g.D = D; g.H = H; g.W = W;
  pick_tile(W, H, D, 1, false, g.wt, g.ht, g.dt);
  g.tiles_w = cdiv(W, g.wt); g.tiles_h = cdiv(H, g.ht); g.tiles_d = cdiv(D, g.dt);
  g.rows = g.wt * g.ht * g.dt;
  g.cin_blocks = kch / 64;
  g.mode = mode;
}

// tensor map of the parity-(pd,ph,pw) sub-lattice of a high-res NDHWC tensor [N][2D][2H][2W][C], seen as [N][D][H][W][C]
static int make_parity_tmap(CUtensorMap* tm, const void* base, int N, int D, int H, int W, int C, int parity, int wt,
                            int ht, int dt) {
  const int pd = parity >> 2, ph = (parity >> 1) & 1, pw = parity & 1;
  const uint64_t row = (uint64_t)C * 2;                 // bytes per voxel
  const uint64_t sw = 2 * row, sh = 2 * (2ull * W) * row, sd = 2 * (2ull * H) * (2ull * W) * row;
  const uint64_t sn = (2ull * D) * (2ull * H) * (2ull * W) * row;
  const uint8_t* b = (const uint8_t*)base + (((uint64_t)pd * (2ull * H) + ph) * (2ull * W) + pw) * row;
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  uint64_t strides[4] = {sw, sh, sd, sn};
  uint32_t box[5] = {64, (uint32_t)wt, (uint32_t)ht, (uint32_t)dt, 1};
  return make_tmap_bf16(tm, b, 5, dims, strides, box);
}

static int make_weight_tmap(CUtensorMap* tm, const void* wpack, int taps, int rows, int kch, int block_n) {
  uint64_t dims[3] = {(uint64_t)kch, (uint64_t)rows, (uint64_t)taps};
  uint64_t strides[2] = {(uint64_t)kch * 2, (uint64_t)rows * kch * 2};
  uint32_t box[3] = {64, (uint32_t)block_n, 1};
  return make_tmap_bf16(tm, wpack, 3, dims, strides, box);
}

// stats / stats_blocks: when both are non-NULL and the persistent kd-fused kernel takes the shape, the kernel also
// writes per-block BatchNorm partial sums of its output to `stats` and *stats_blocks = number of partial blocks;
// otherwise *stats_blocks = 0 and the caller runs the separate statistics pass.
static int conv3_igemm_impl(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                            float* stats, int* stats_blocks, cudaStream_t st) {
  if (stats_blocks) *stats_blocks = 0;
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64, "conv3_igemm: Cin=%d must be a multiple of 64", Cin);
  SIVAE_CHECK(Cout % 64 == 0 && Cout >= 64, "conv3_igemm: Cout=%d must be a multiple of 64", Cout);
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "conv3_igemm: empty tensor");
  // Cout = 64 (Cin = 64, 128, ...): persistent kd-fused kernel when 8 x 16 x 4 chunks tile the volume well
  // (SIVAE_CONV_KD=0 disables it, =force takes it for every Cout = 64 shape -- the parity tests use both).
  if (Cout == 64) {
    const char* kdenv = getenv("SIVAE_CONV_KD");
    const bool off = kdenv != nullptr && kdenv[0] == '0';
    const bool force = kdenv != nullptr && kdenv[0] == 'f';
    KdGeom kg;
    kg.N = N; kg.D = D; kg.H = H; kg.W = W;
    kg.tiles_w = cdiv(W, kKwW); kg.tiles_h = cdiv(H, kKwH); kg.tiles_d = cdiv(D, kKdP);
    kg.cin_blocks = Cin / 64;
    kg.items = (long long)kg.tiles_w * kg.tiles_h * kg.tiles_d * N;
    const double eff = ((double)W * H * D) / ((double)kg.tiles_w * kKwW * kg.tiles_h * kKwH * kg.tiles_d * kKdP);
    if (!off && (force || (eff >= 0.8 && kg.items >= 2 * num_sms()))) {
      CUtensorMap tA, tB, tC;
      if (make_act_tmap(&tA, x, N, D, H, W, Cin, kKwW, kKwH, 1)) return -1;
      if (make_act_tmap(&tC, y, N, D, H, W, 64, kKwW, kKwH, 1)) return -1;
      if (make_weight_tmap(&tB, wpack, 27, 64, Cin, 64)) return -1;
      static bool attr_set = false;
      if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(conv3_kd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKdSmem),
                       "cudaFuncSetAttribute(conv3_kd3)")) return -1;
        attr_set = true;
      }
      const unsigned ctas = (unsigned)(kg.items < (long long)num_sms() ? kg.items : (long long)num_sms());
      const bool fuse_stats = stats != nullptr && stats_blocks != nullptr;
      launch_conv(conv3_kd3_kernel, dim3(ctas), 224, kKdSmem, st, tA, tB, tC, kg, fuse_stats ? stats : (float*)nullptr);
      SIVAE_LAUNCH_OK("conv3_kd3_kernel");
      if (fuse_stats) *stats_blocks = (int)ctas * 4;
      return 0;
    }
  }
  // Cin = 64: persistent kw-slab kernel.  OPT-IN (SIVAE_CONV_KW=1 when 8 x 16 x 2 patches tile the volume well, =force for
  // every Cin = 64 shape): parity-tested, but it cut L2 traffic 3x without getting faster than the tap-by-tap kernel
  // (1.31 vs 1.26 ms on 64->64 @ 8x80x96x80) -- the evidence that these layers are bound by shared-memory operand
  // reads, which the kd-fused kernel above addresses.
  if (Cin == 64 && getenv("SIVAE_CONV_KW") != nullptr) {
    const char* kwenv = getenv("SIVAE_CONV_KW");
    const bool off = kwenv[0] == '0';
    const bool force = kwenv[0] == 'f';
    KwGeom kg;
    kg.N = N; kg.D = D; kg.H = H; kg.W = W;
    kg.tiles_w = cdiv(W, kKwW); kg.tiles_h = cdiv(H, kKwH); kg.tiles_d = cdiv(D, 2);
    kg.nblk = Cout / 64;
    kg.items = (long long)kg.tiles_w * kg.tiles_h * kg.tiles_d * N * kg.nblk;
    const double eff = ((double)W * H * D) / ((double)kg.tiles_w * kKwW * kg.tiles_h * kKwH * kg.tiles_d * 2);
    if (!off && (force || (eff >= 0.8 && kg.items >= 148))) {
      CUtensorMap tA, tB, tC;
      {
        uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        uint64_t strides[4] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128, (uint64_t)D * H * W * 128};
        uint32_t box[5] = {64, (uint32_t)kKwW, (uint32_t)(kKwH + 2), 4, 1};
        if (make_tmap_bf16(&tA, x, 5, dims, strides, box)) return -1;
      }
      if (make_act_tmap(&tC, y, N, D, H, W, Cout, kKwW, kKwH, 1)) return -1;
      if (make_weight_tmap(&tB, wpack, 27, Cout, 64, 64)) return -1;
      static bool attr_set = false;
      if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(conv3_kw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKwSmem),
                       "cudaFuncSetAttribute(conv3_kw64)")) return -1;
        attr_set = true;
      }
      const unsigned ctas = (unsigned)(kg.items < (long long)num_sms() ? kg.items : (long long)num_sms());
      launch_conv(conv3_kw64_kernel, dim3(ctas), 224, kKwSmem, st, tA, tB, tC, kg);
      SIVAE_LAUNCH_OK("conv3_kw64_kernel");
      return 0;
    }
  }
  ConvGeom g;
  fill_geom(g, N, D, H, W, Cin, kTapsPlain);
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "conv3_igemm: too many tiles");
  // Tile width for Cout % 256 == 0 at the latent resolution (80 voxel tiles at batch 8): N = 128 gives 160 CTAs on 148
  // SMs, N = 256 one wave of 80 CTAs that each read their A tile once.  With a 3-stage ring N = 256 measured SLOWER
  // (55.8 vs 51.0 us per launch: 68 idle SMs cost more than the saved A re-reads), with a 4-stage ring (4 x 48 KB)
  // FASTER (46.9 us, step 66.6 -> 65.7 ms): the k-step rate of these one-CTA-per-SM chains is TMA latency / ring depth.
  // Grids whose N = 128 form already fits one wave keep N = 128 with the 6-stage ring below (more CTAs, deeper ring).
  // SIVAE_N256: 0 = never, 1 = 3-stage, 4 = 4-stage whenever tiles <= SMs; unset = the rule above.
  int n256_stages = 0;
  if (Cout % 256 == 0 && tiles <= num_sms()) {
    const char* e = getenv("SIVAE_N256");
    if (e != nullptr) n256_stages = e[0] == '4' ? 4 : e[0] == '1' ? 3 : 0;
    else if (tiles * (Cout / 128) > (long long)num_sms()) n256_stages = 4;
  }
  const int block_n = n256_stages ? 256 : (Cout % 128 == 0) ? 128 : 64;
  TmapPack tmA, tmC;
  CUtensorMap tmB;
  if (make_act_tmap(&tmA.m[0], x, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  if (make_act_tmap(&tmC.m[0], y, N, D, H, W, Cout, g.wt, g.ht, g.dt)) return -1;
  for (int i = 1; i < 8; ++i) { tmA.m[i] = tmA.m[0]; tmC.m[i] = tmC.m[0]; }
  if (make_weight_tmap(&tmB, wpack, 27, Cout, Cin, block_n)) return -1;
  const ToOneEpilogue ep{};
  if (block_n == 256) {
    if (n256_stages == 4) return launch_igemm<256, 4, 0>(tmA, tmB, tmC, g, tiles, Cout / 256, ep, st);
    return launch_igemm<256, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 256, ep, st);
  }
  // Grids that leave every SM at most one CTA (latent-resolution layers of small batches) are bound by the TMA
  // round-trip of a 3-stage ring (one k-step per latency/3): a 6-stage ring keeps twice the bytes in flight.  With more
  // CTAs than SMs the 3-stage kernel wins because two of its CTAs share an SM (SIVAE_DEEP_RING=0 disables, =1 forces).
  if (block_n == 128) {
    const char* e = getenv("SIVAE_DEEP_RING");
    const bool deep = e ? e[0] == '1' : tiles * (Cout / 128) <= (long long)num_sms();
    if (deep) return launch_igemm<128, 6, 0>(tmA, tmB, tmC, g, tiles, Cout / 128, ep, st);
    return launch_igemm<128, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 128, ep, st);
  }
  return launch_igemm<64, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 64, ep, st);
}

int conv3_igemm(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                cudaStream_t st) {
  return conv3_igemm_impl(x, wpack, y, N, D, H, W, Cin, Cout, nullptr, nullptr, st);
}

// pointwise.cu
int bn_train_coeffs(const void* y, long long nvox, int C, const float* gamma, const float* beta, float* rm, float* rv,
                    long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                    void* ws, size_t ws_bytes, cudaStream_t st);
int bn_coeffs_from_partials(const float* partial, int nblocks, long long nvox, int C, const float* gamma,
                            const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                            float* mean, float* invstd, float* scale, float* shift, cudaStream_t st);
size_t bn_workspace_bytes(int C);

// Convolution + train-mode BatchNorm coefficients of its output in one call (models/models.py:17-18, :21-22: Conv3d
// followed by BatchNorm3d): the statistics come out of the convolution's epilogue when the persistent kernel runs,
// from the separate bn_stats pass otherwise.
int conv3_igemm_bn(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                   const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                   float eps, float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  SIVAE_CHECK(ws && ws_bytes >= bn_workspace_bytes(Cout), "conv3_igemm_bn: workspace too small");
  int blocks = 0;
  const bool allow = getenv("SIVAE_NO_FUSED_STATS") == nullptr;
  int rc = conv3_igemm_impl(x, wpack, y, N, D, H, W, Cin, Cout, allow ? (float*)ws : nullptr, allow ? &blocks : nullptr, st);
  if (rc) return rc;
  const long long nvox = (long long)N * D * H * W;
  if (blocks > 0)
    return bn_coeffs_from_partials((const float*)ws, blocks, nvox, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean,
                                   invstd, scale, shift, st);
  return bn_train_coeffs(y, nvox, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale, shift, ws, ws_bytes,
                         st);
}

// y_hi[n, 2d+pd, 2h+ph, 2w+pw, co] = conv3(upsample2(x_lo), w):  8 parity-specific 2x2x2 convolutions on the low-res grid.
// x_lo [N][D][H][W][Cin], wup bf16 [64 = parity*8+abc][Cout][Cin], y_hi [N][2D][2H][2W][Cout].
static int upconv3_fprop_impl(const void* x_lo, const void* wup, void* y_hi, int N, int D, int H, int W, int Cin,
                             int Cout, float* stats, int* stats_blocks, cudaStream_t st) {
  if (stats_blocks) *stats_blocks = 0;
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64 && Cout % 64 == 0 && Cout >= 64,
              "upconv3_fprop: Cin=%d, Cout=%d must be multiples of 64", Cin, Cout);
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "upconv3_fprop: empty tensor");
  // Cout = 64: persistent kernel with shared-tile tap stacking when 8 x 16 patches tile the low-res plane well
  // (SIVAE_UPCONV_FUSED=0 disables it, =force takes it for every Cout = 64 shape -- the parity tests use both)
  if (Cout == 64) {
    const char* e = getenv("SIVAE_UPCONV_FUSED");
    const bool off = e != nullptr && e[0] == '0';
    const bool force = e != nullptr && e[0] == 'f';
    UpGeom ug;
    ug.N = N; ug.D = D; ug.H = H; ug.W = W;
    ug.tiles_w = cdiv(W, kKwW); ug.tiles_h = cdiv(H, kKwH);
    ug.cin_blocks = Cin / 64;
    ug.items = (long long)ug.tiles_w * ug.tiles_h * D * 2 * N;
    const double eff = ((double)W * H) / ((double)ug.tiles_w * kKwW * ug.tiles_h * kKwH);
    if (!off && (force || (eff >= 0.8 && ug.items >= 2 * num_sms()))) {
      CUtensorMap tA, tB;
      TmapPack tC;
      if (make_act_tmap(&tA, x_lo, N, D, H, W, Cin, kKwW, kKwH, 1)) return -1;
      for (int p = 0; p < 8; ++p)
        if (make_parity_tmap(&tC.m[p], y_hi, N, D, H, W, Cout, p, kKwW, kKwH, 1)) return -1;
      if (make_weight_tmap(&tB, wup, 64, Cout, Cin, 64)) return -1;
      static bool attr_set = false;
      if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(upconv3_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem),
                       "cudaFuncSetAttribute(upconv3_fused)")) return -1;
        attr_set = true;
      }
      const unsigned ctas = (unsigned)(ug.items < (long long)num_sms() ? ug.items : (long long)num_sms());
      const bool fuse_stats = stats != nullptr && stats_blocks != nullptr;
      launch_conv(upconv3_fused_kernel, dim3(ctas), 224, kUpSmem, st, tA, tB, tC, ug, fuse_stats ? stats : (float*)nullptr);
      SIVAE_LAUNCH_OK("upconv3_fused_kernel");
      if (fuse_stats) *stats_blocks = (int)ctas * 4;
      return 0;
    }
  }
  ConvGeom g;
  fill_geom(g, N, D, H, W, Cin, kTapsUpFprop);
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "upconv3_fprop: too many tiles");
  const int block_n = (Cout % 128 == 0) ? 128 : 64;
  TmapPack tmA, tmC;
  CUtensorMap tmB;
  if (make_act_tmap(&tmA.m[0], x_lo, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  for (int i = 1; i < 8; ++i) tmA.m[i] = tmA.m[0];
  for (int p = 0; p < 8; ++p)
    if (make_parity_tmap(&tmC.m[p], y_hi, N, D, H, W, Cout, p, g.wt, g.ht, g.dt)) return -1;
  if (make_weight_tmap(&tmB, wup, 64, Cout, Cin, block_n)) return -1;
  const ToOneEpilogue ep{};
  if (block_n == 128) return launch_igemm<128, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 128, ep, st);
  return launch_igemm<64, 3, 0>(tmA, tmB, tmC, g, tiles, Cout / 64, ep, st);
}

int upconv3_fprop(const void* x_lo, const void* wup, void* y_hi, int N, int D, int H, int W, int Cin, int Cout,
                  cudaStream_t st) {
  return upconv3_fprop_impl(x_lo, wup, y_hi, N, D, H, W, Cin, Cout, nullptr, nullptr, st);
}

// Upsample(2) + Conv3d + train-mode BatchNorm3d coefficients of the (high-res) output in one call
// (models/models.py:58-60); statistics from the convolution epilogue when the persistent kernel runs.
int upconv3_fprop_bn(const void* x_lo, const void* wup, void* y_hi, int N, int D, int H, int W, int Cin, int Cout,
                     const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                     float eps, float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  SIVAE_CHECK(ws && ws_bytes >= bn_workspace_bytes(Cout), "upconv3_fprop_bn: workspace too small");
  int blocks = 0;
  const bool allow = getenv("SIVAE_NO_FUSED_STATS") == nullptr;
  int rc = upconv3_fprop_impl(x_lo, wup, y_hi, N, D, H, W, Cin, Cout, allow ? (float*)ws : nullptr,
                              allow ? &blocks : nullptr, st);
  if (rc) return rc;
  const long long nvox = (long long)N * D * H * W * 8;
  if (blocks > 0)
    return bn_coeffs_from_partials((const float*)ws, blocks, nvox, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean,
                                   invstd, scale, shift, st);
  return bn_train_coeffs(y_hi, nvox, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale, shift, ws,
                         ws_bytes, st);
}

// dx_lo = upsample2^T(conv3^T(dy_hi)):  one 64-tap (8 parities x 8 taps) implicit GEMM gathering dy on its parity
// sub-lattices.  dy_hi [N][2D][2H][2W][Cout], wupT bf16 [64][Cin][Cout], dx_lo [N][D][H][W][Cin]  (D,H,W low-res).
int upconv3_dgrad(const void* dy_hi, const void* wupT, void* dx_lo, int N, int D, int H, int W, int Cin, int Cout,
                  cudaStream_t st) {
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64 && Cout % 64 == 0 && Cout >= 64,
              "upconv3_dgrad: Cin=%d, Cout=%d must be multiples of 64", Cin, Cout);
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "upconv3_dgrad: empty tensor");
  ConvGeom g;
  fill_geom(g, N, D, H, W, Cout, kTapsUpDgrad);
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "upconv3_dgrad: too many tiles");
  const int block_n = (Cin % 128 == 0) ? 128 : 64;
  TmapPack tmA, tmC;
  CUtensorMap tmB;
  for (int p = 0; p < 8; ++p)
    if (make_parity_tmap(&tmA.m[p], dy_hi, N, D, H, W, Cout, p, g.wt, g.ht, g.dt)) return -1;
  if (make_act_tmap(&tmC.m[0], dx_lo, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  for (int i = 1; i < 8; ++i) tmC.m[i] = tmC.m[0];
  if (make_weight_tmap(&tmB, wupT, 64, Cin, Cout, block_n)) return -1;
  const ToOneEpilogue ep{};
  if (block_n == 128) return launch_igemm<128, 3, 0>(tmA, tmB, tmC, g, tiles, Cin / 128, ep, st);
  return launch_igemm<64, 3, 0>(tmA, tmB, tmC, g, tiles, Cin / 64, ep, st);
}

// fp32 [Cout][Cin][27] -> bf16 wup[parity*8+abc][Cout][Cin] and wupT[parity*8+abc][Cin][Cout]: per axis and output
// parity p the three taps collapse onto two low-res offsets, p=0: {k0}, {k1+k2};  p=1: {k0+k1}, {k2}.
__global__ void pack_upconv3_weights_kernel(const float* __restrict__ w, int Cout, int Cin,
                                            __nv_bfloat16* __restrict__ wup, __nv_bfloat16* __restrict__ wupT) {
  const long long total = 64ll * Cout * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int pa = (int)(i / ((long long)Cin * Cout));
    const int p = pa >> 3, abc = pa & 7;
    const float* src = w + ((long long)co * Cin + ci) * 27;
    float acc = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
      const int pd = p >> 2, a = abc >> 2;
      if ((pd == 0 ? (kd == 0 ? 0 : 1) : (kd == 2 ? 1 : 0)) != a) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ph = (p >> 1) & 1, b = (abc >> 1) & 1;
        if ((ph == 0 ? (kh == 0 ? 0 : 1) : (kh == 2 ? 1 : 0)) != b) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int pw = p & 1, c = abc & 1;
          if ((pw == 0 ? (kw == 0 ? 0 : 1) : (kw == 2 ? 1 : 0)) != c) continue;
          acc += src[(kd * 3 + kh) * 3 + kw];
        }
      }
    }
    const __nv_bfloat16 v = __float2bfloat16_rn(acc);
    if (wup) wup[((long long)pa * Cout + co) * Cin + ci] = v;
    if (wupT) wupT[((long long)pa * Cin + ci) * Cout + co] = v;
  }
}

int pack_upconv3_weights(const float* w, int Cout, int Cin, void* wup, void* wupT, cudaStream_t st) {
  SIVAE_CHECK(Cout > 0 && Cin > 0, "pack_upconv3_weights: bad dims");
  const long long total = 64ll * Cout * Cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_upconv3_weights_kernel<<<blocks, 256, 0, st>>>(w, Cout, Cin, (__nv_bfloat16*)wup, (__nv_bfloat16*)wupT);
  SIVAE_LAUNCH_OK("pack_upconv3_weights_kernel");
  return 0;
}

// =================================================================================================
// C = 64 -> 1 channel, 3x3x3: "all taps as GEMM columns", persistent and streaming along depth.
// P[v_in][tap] = sum_c x[v_in][c] * w[c][tap] is one tiny GEMM per INPUT voxel (N = 27 taps), so an input row is
// multiplied ONCE for all the outputs it feeds; the convolution is then out[v] = sum_tap P[v + delta(tap)][tap], a
// 27-term gather from shared memory.  The weights are split bf16(w) | bf16(w - bf16(w)) (~fp32 products: the output
// feeds the loss directly) and stacked into ONE N = 64 operand (columns 0..31 hi partials, 32..63 lo partials).
//
// A work item is a 16(w) x 8(h) patch over a CHUNK of output planes.  The CTA walks the chunk's input planes
// d_start-1 .. d_end: each plane's 18 x 10 halo (180 voxel rows, one 23 KB TMA) is loaded once, turned into P once
// (2 M-tiles x 4 UMMA 128x64x16) and its P rows serve the three output planes around it; the gather for output plane d
// runs as soon as P of planes d-1, d, d+1 sit in the 4-deep P ring.  Per 128 outputs: 23 KB of TMA, 8 UMMAs and
// 180 x 27 floats of P traffic -- the first version of this kernel (one 18 x 10 x 3 box per output plane) moved 3x that
// and was bound by shared-memory bandwidth (ncu: every pipe < 30 % busy, profiles/r01c).
// Persistent CTA, one per SM: resident weights, 4-deep input ring, ring of eight 64-column TMEM slots handed over per
// M-tile, 5-deep P ring.  18 warps: TMA producer, MMA issuer, 2 x 4 converter warps (TMEM -> hi + lo -> P) and 2 x 4
// gather warps (27-term gather -> bias / ReLU / dropout -> fp32 store), handing P planes over through mbarriers; the
// two groups of a kind take alternate planes (each is one warp per scheduler running a dependent chain).  With ONE
// warp group doing both halves the kernel ran at the length of that group's dependency chain (ncu source view of the
// round-2 build: 39 % of its samples in TMEM -> P, 44 % in gather / Philox / store, 17 % at the barrier between them,
// every pipe < 30 % busy, 300 us against 100 us of HBM time); the two halves now overlap plane by plane.
// =================================================================================================
static constexpr int kHW = 16, kHH = 8;                        // outputs per patch (w, h)
static constexpr int kHBW = kHW + 2, kHBH = kHH + 2, kHBD = 3; // input halo (kHBD only used by the c1 kernels' 3-plane halo)
static constexpr int kHRows = kHBW * kHBH * kHBD;              // 540 (3-plane halo of the c1 kernels)
static constexpr int kSPRows = kHBW * kHBH;                    // 180 voxel rows per input plane
static constexpr int kPStride = 27;                            // floats per P row (odd: conflict-free)
static constexpr int kSASlot = ((kSPRows * 128 + 1023) / 1024) * 1024;     // 23,552: stride of the input ring
static constexpr int kSASlots = 4;
static constexpr int kSARegion = (kSASlots - 1) * kSASlot + 2 * kTileBytes;  // the 2nd M-tile reads 76 rows past a plane
static constexpr int kHBBytes = 64 * 128;                                   // [hi 32 taps | lo 32 taps] x 64 channels
static constexpr int kSPSlots = 5;
static constexpr int kSPSlotFloats = kSPRows * kPStride;                    // 4,860
static constexpr int kSPBytes = kSPSlots * kSPSlotFloats * 4;               // 97,200
static constexpr int kHSlots = 8;
static constexpr int kTo1Groups = 2;                           // converter groups = gather groups (4 warps each)
static constexpr int kTo1Threads = 64 + 2 * kTo1Groups * 128;
static constexpr int kTo1Smem = kSARegion + kHBBytes + kSPBytes + 1024 + 512;

struct To1Geom {
  int N, D, H, W;
  int tiles_w, tiles_h;
  int chunks, dc;        // depth chunks per volume, planes per chunk
  int items;             // tiles_w * tiles_h * chunks * N
};

__global__ void __launch_bounds__(kTo1Threads, 1)
conv3_to1_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const To1Geom g,
                      const ToOneEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + kSARegion;
  float* P = reinterpret_cast<float*>(smem_b + kHBBytes);          // [kSPSlots][180][kPStride]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(P) + kSPBytes);
  uint64_t* a_empty = a_full + kSASlots;
  uint64_t* slot_full = a_empty + kSASlots;
  uint64_t* slot_empty = slot_full + kHSlots;
  uint64_t* p_full = slot_empty + kHSlots;
  uint64_t* p_empty = p_full + kSPSlots;
  uint64_t* b_full = p_empty + kSPSlots;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kSASlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kHSlots; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 4); }
    for (int s = 0; s < kSPSlots; ++s) { mbar_init(&p_full[s], 128); mbar_init(&p_empty[s], 3 * 128); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp_id == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // item -> patch origin, sample, output planes [d_lo, d_hi)
  auto decode = [&](int id, int& w0, int& h0, int& d_lo, int& d_hi, int& n) {
    const int tw = id % g.tiles_w; id /= g.tiles_w;
    const int th = id % g.tiles_h; id /= g.tiles_h;
    const int ch = id % g.chunks;
    n = id / g.chunks;
    w0 = tw * kHW; h0 = th * kHH;
    d_lo = ch * g.dc;
    d_hi = min(g.D, d_lo + g.dc);
  };

  if (warp_id == 0) {
    // ===== TMA producer: weights once, then one input plane at a time =====
    if (elect_one()) {
      mbar_expect_tx(b_full, kHBBytes);
      tma_load_3d(smem_b, &tmB, b_full, 0, 0, 0);
      tma_load_3d(smem_b + 32 * 128, &tmB, b_full, 0, 0, 1);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
        int w0, h0, d_lo, d_hi, n;
        decode(item, w0, h0, d_lo, d_hi, n);
        for (int z = d_lo - 1; z <= d_hi; ++z, ++it) {
          const int s = it % kSASlots;
          mbar_wait(&a_empty[s], ((it / kSASlots) & 1u) ^ 1u);
          mbar_expect_tx(&a_full[s], (uint32_t)kSPRows * 128u);
          tma_load_5d(smem + s * kSASlot, &tmA, &a_full[s], 0, w0 - 1, h0 - 1, z, n);
        }
      }
    }
  } else if (warp_id == 1) {
    // ===== MMA issuer: per input plane 2 M-tiles x 4 K steps of UMMA 128 x 64 x 16, one TMEM slot per M-tile =====
    // One elected thread for the whole loop: with 4 UMMAs of N = 64 (128 cycles of tensor work) per M-tile, a per-step
    // elect + __syncwarp and rebuilt descriptors (~85 instructions, ~680 cycles) made the issuer the pace setter.
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(b_full, 0);
      const uint64_t bdesc = make_smem_desc(smem_u32(smem_b), 16, 1024);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, 1024);
      uint32_t s = 0, pa = 0, slot = 0, ps = 1;               // ps: phase to wait for on slot_empty (starts "free")
      for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
        int w0, h0, d_lo, d_hi, n;
        decode(item, w0, h0, d_lo, d_hi, n);
        for (int z = d_lo - 1; z <= d_hi; ++z) {
          mbar_wait(&a_full[s], pa);
          const uint64_t adesc = adesc0 + (uint64_t)(s * (uint32_t)(kSASlot >> 4));
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mbar_wait(&slot_empty[slot], ps);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)   // +32 bytes per K step = +2 in the descriptor's (address >> 4) field
              umma_bf16(tmem_base + slot * 64u, adesc + (uint64_t)(m * (kTileBytes >> 4) + 2 * k), bdesc + 2 * k, idesc,
                        k != 0 ? 1u : 0u);
            umma_commit(&slot_full[slot]);
            if (m == 1) umma_commit(&a_empty[s]);
            if (++slot == kHSlots) { slot = 0; ps ^= 1u; }
          }
          if (++s == kSASlots) { s = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp_id < 2 + 4 * kTo1Groups) {
    // ===== converters (kTo1Groups x 4 warps <-> TMEM lane quadrants by warp_id % 4): P[plane][row][tap] = hi + lo =====
    // P plane `pc` (counted over the CTA's whole run) lives in ring slot pc % kSPSlots and is converted by group
    // pc % kTo1Groups; its slot is free again once the gather steps pc, pc+1, pc+2 have arrived on p_empty.
    const int q = warp_id & 3;
    const uint32_t grp = (uint32_t)(warp_id - 2) >> 2;
    const uint32_t p_addr = smem_u32(P);                  // shared-window address of the P ring: STS below
    uint32_t pc = 0;
    for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
      int w0, h0, d_lo, d_hi, n;
      decode(item, w0, h0, d_lo, d_hi, n);
      const int planes = d_hi - d_lo + 2;
      for (int zi = 0; zi < planes; ++zi, ++pc) {
        if (pc % kTo1Groups != grp) continue;
        const uint32_t pslot = pc % kSPSlots;
        mbar_wait(&p_empty[pslot], ((pc / kSPSlots) & 1u) ^ 1u);
        const uint32_t Pz = p_addr + pslot * (kSPSlotFloats * 4u);
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          const uint32_t sl = 2u * pc + (uint32_t)m;       // the MMA warp's M-tile counter
          const uint32_t slot = sl % kHSlots;
          const int prow = m * 128 + q * 32 + lane;
          mbar_wait(&slot_full[slot], (sl / kHSlots) & 1u);
          tc_fence_after();
          if (m * 128 + q * 32 < kSPRows) {           // warp-uniform: this quadrant holds rows of the plane
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 64u;
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32u, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[slot]);
            if (prow < kSPRows) {
              const uint32_t dst = Pz + (uint32_t)(prow * kPStride) * 4u;
#pragma unroll
              for (int t = 0; t < 27; ++t) sts_f32(dst + t * 4u, __uint_as_float(v0[t]) + __uint_as_float(v1[t]));
            }
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[slot]);
          }
        }
        mbar_arrive(&p_full[pslot]);                       // every converter thread: its rows of plane pc are in P
      }
    }
  } else {
    // ===== gather warps (kTo1Groups x 4): one output voxel of the patch per thread; step pc by group pc % kTo1Groups =====
    const uint32_t grp = (uint32_t)(warp_id - 2 - 4 * kTo1Groups) >> 2;
    const int t128 = ((warp_id - 2) & 3) * 32 + lane;
    const int ow = t128 % kHW, oh = t128 / kHW;
    const float bias = ep.bias ? ep.bias[0] : 0.f;
    const float inv_keep = 1.f / (1.f - ep.p);
    const uint32_t p_addr = smem_u32(P);                  // shared-window address of the P ring: LDS below
    uint32_t pc = 0;
    for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
      int w0, h0, d_lo, d_hi, n;
      decode(item, w0, h0, d_lo, d_hi, n);
      const int w = w0 + ow, h = h0 + oh;
      const bool inside = (w < g.W && h < g.H);
      const int planes = d_hi - d_lo + 2;
      for (int zi = 0; zi < planes; ++zi, ++pc) {
        if (pc % kTo1Groups != grp) continue;
        // slots of planes pc-2, pc-1, pc (the converters of the other group(s) filled some of them: wait for all three)
        uint32_t sk[3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const uint32_t pk = pc + (uint32_t)kd - 2u;
          sk[kd] = pk % kSPSlots;
          if (pc + (uint32_t)kd >= 2u) mbar_wait(&p_full[sk[kd]], (pk / kSPSlots) & 1u);
        }
        if (zi >= 2 && inside) {
          // output plane d = d_lo + zi - 2 gathers tap kd from input plane (zi - 2 + kd)
          float acc = 0.f;
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const uint32_t Pk = p_addr + sk[kd] * (kSPSlotFloats * 4u);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const int row = (oh + t9 / 3) * kHBW + ow + t9 % 3;
              acc += lds_f32(Pk + (uint32_t)(row * kPStride + kd * 9 + t9) * 4u);
            }
          }
          const int d = d_lo + zi - 2;
          const long long vox = (((long long)n * g.D + d) * g.H + h) * g.W + w;
          float r = acc + bias;
          if (ep.act == 1) {
            r = fmaxf(r, 0.f);
            if (ep.mask != nullptr) r = ep.mask[vox] ? r * inv_keep : 0.f;
            else if (ep.p > 0.f) r = philox_keep(resolve_seed(ep.seed), (unsigned long long)vox, ep.p) ? r * inv_keep : 0.f;
          }
          ep.y[vox] = r;
        }
        // every step (also the two that open an item and gather nothing) signs off the three planes of its window:
        // plane q has collected all its arrivals when steps q, q+1 and q+2 are done, whichever group ran them
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
          if (pc + (uint32_t)kd >= 2u) mbar_arrive(&p_empty[sk[kd]]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) tmem_dealloc(tmem_base, 512);
}

// fp32 [64][27] -> bf16 [2][32][64]: slab 0 = bf16(w) per tap row (rows 27..31 zero), slab 1 = bf16(w - bf16(w))
__global__ void pack_to1_halo_weights_kernel(const float* __restrict__ w, int flip, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * 32 * 64) return;
  const int c = i % 64, tap = (i / 64) % 32, part = i / (64 * 32);
  float v = 0.f;
  if (tap < 27) v = w[c * 27 + (flip ? 26 - tap : tap)];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  wp[i] = part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
}

// =================================================================================================
// 1 -> 64 channels, 3x3x3 (encoder stem Conv3d(1,64,3), models/models.py:92; input gradient of the decoder tail):
// y[v][c] = bias[c] + sum_tap w[c][tap] * x1[v + delta(tap)] as a tensor-core GEMM whose A operand (the im2col of
// the one-channel fp32 input) is BUILT IN SHARED MEMORY by the CTA's threads: M = 128 voxels (16 x 8 x 1), N = 64,
// K = 96 = [x_hi(32) | x_lo(32) | x_hi(32)] against B = [w_hi | w_hi | w_lo]  (bf16 split of both operands:
// x_hi*w_hi + x_lo*w_hi + x_hi*w_lo recovers ~16 mantissa bits of the fp32 product).
// HBM-bound on the 128 B/voxel output write instead of FMA-bound (1728 FMA/voxel on CUDA cores).
// Persistent CTA, one per SM, three concurrent roles over double-buffered A tiles / TMEM stages / output staging:
//   warps 0-3  builders : prefetch the next item's fp32 halo into registers, stage it, build one A row per thread
//   warp  4    MMA      : weights loaded once (TMA), 6 UMMA 128x64x16 per item
//   warps 5-8  epilogue : TMEM -> +bias -> bf16 -> swizzled smem -> TMA store
// =================================================================================================
static constexpr int kC1ABytes = 2 * kTileBytes;          // two 64-wide K blocks of the 128-row A tile
static constexpr int kC1BBytes = 2 * 64 * 128;            // [2 K blocks][64 channels][64]
static constexpr int kC1Halo = (kHRows + 3) & ~3;         // floats per staged halo
static constexpr int kC1Smem = 2 * kC1ABytes + kC1BBytes + kTileBytes + 2 * kC1Halo * 4 + 64 * 4 + 1024 + 256;

template <bool HAS_BIAS>
__global__ void __launch_bounds__(288, 2)
c1_to_c64_tc_kernel(const float* __restrict__ x1, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int N, int D, int H, int W,
                    int tiles_w, int tiles_h, int items, float* __restrict__ stats_partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + 2 * kC1ABytes;
  uint8_t* smem_o = smem_b + kC1BBytes;
  float* xs = reinterpret_cast<float*>(smem_o + kTileBytes);       // [2][kC1Halo]
  float* bias_s = xs + 2 * kC1Halo;                                // [64], 16-byte aligned (kC1Halo % 4 == 0)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bias_s + 64);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp_id == 4) {
    if (lane == 0) {
      prefetch_tmap(&tmB);
      prefetch_tmap(&tmC);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a_full[s], 128);     // every builder thread arrives after fencing its row
        mbar_init(&a_empty[s], 1);
        mbar_init(&acc_full[s], 1);
        mbar_init(&acc_empty[s], 4);
      }
      mbar_init(b_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 128);
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](int id, int& w0, int& h0, int& d0, long long& n) {   // 32-bit: items < 2^31 (host check)
    const int tw = id % tiles_w; id /= tiles_w;
    const int th = id % tiles_h; id /= tiles_h;
    d0 = id % D;
    n = id / D;
    w0 = tw * kHW; h0 = th * kHH;
  };

  if (warp_id == 4) {
    // ===== MMA issuer =====
    if (elect_one()) {
      mbar_expect_tx(b_full, kC1BBytes);
      tma_load_3d(smem_b, &tmB, b_full, 0, 0, 0);
      tma_load_3d(smem_b + 64 * 128, &tmB, b_full, 0, 0, 1);
    }
    __syncwarp();
    if (elect_one()) {                       // one elected thread issues the whole loop
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(b_full, 0);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 16, 1024);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, 1024);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const uint32_t s = it & 1, ph = (it >> 1) & 1u;
        mbar_wait(&acc_empty[s], ph ^ 1u);
        mbar_wait(&a_full[s], ph);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(s * (uint32_t)(kC1ABytes >> 4));
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) {   // K block 0: 4 steps [x_hi | x_lo], K block 1: 2 steps [x_hi]
          const int kb = kk >> 2, k = kk & 3;
          umma_bf16(tmem_base + s * 64u, adesc + (uint64_t)(kb * (kTileBytes >> 4) + 2 * k),
                    bdesc0 + (uint64_t)(kb * (64 * 128 >> 4) + 2 * k), idesc, kk != 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else if (warp_id < 4) {
    // ===== builders: one A row (output voxel) per thread =====
    const int row = threadIdx.x;                          // 0..127
    const int ow = row % kHW, oh = row / kHW;
    constexpr int kPer = (kHRows + 127) / 128;            // halo floats per thread (5)
    float pre[kPer];
    // this thread's halo cells never change: offsets from the halo origin and packed (wx, hy, dz) for the bounds tests
    int off[kPer], cell[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = row + j * 128;
      const int wx = i % kHBW, hy = (i / kHBW) % kHBH, dz = i / (kHBW * kHBH);
      off[j] = (dz * H + hy) * W + wx;
      cell[j] = i < kHRows ? (wx | (hy << 8) | (dz << 16)) : -1;
    }
    auto fetch = [&](int item) {
      int w0, h0, d0; long long n;
      decode(item, w0, h0, d0, n);
      const float* origin = x1 + (((n * D + (d0 - 1)) * H + (h0 - 1)) * W + (w0 - 1));
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int d = d0 - 1 + (cell[j] >> 16), h = h0 - 1 + ((cell[j] >> 8) & 255), w = w0 - 1 + (cell[j] & 255);
        float v = 0.f;
        if (cell[j] >= 0 && (unsigned)d < (unsigned)D && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
          v = __ldg(origin + off[j]);
        pre[j] = v;
      }
    };
    if ((int)blockIdx.x < items) fetch(blockIdx.x);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const uint32_t s = it & 1, ph = (it >> 1) & 1u;
      const uint32_t xh = smem_u32(xs + s * kC1Halo);       // shared-window address: LDS / STS below
#pragma unroll
      for (int j = 0; j < kPer; ++j)
        if (row + j * 128 < kHRows) sts_f32(xh + (uint32_t)(row + j * 128) * 4u, pre[j]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (item + (int)gridDim.x < items) fetch(item + gridDim.x);     // next item's halo: in flight while this row is built
      uint32_t hi[16], lo[16];                              // 32 bf16 each: taps 0..26, then zeros
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = 0.f, b = 0.f;
        if (2 * j < 27) a = lds_f32(xh + (uint32_t)((((2 * j) / 9) * kHBH + oh + ((2 * j) / 3) % 3) * kHBW + ow + (2 * j) % 3) * 4u);
        if (2 * j + 1 < 27)
          b = lds_f32(xh + (uint32_t)((((2 * j + 1) / 9) * kHBH + oh + ((2 * j + 1) / 3) % 3) * kHBW + ow + (2 * j + 1) % 3) * 4u);
        hi[j] = pack_bf16x2(a, b);                          // one packed conversion; bf16 -> fp32 is a 16-bit shift
        lo[j] = pack_bf16x2(a - __uint_as_float(hi[j] << 16), b - __uint_as_float(hi[j] & 0xffff0000u));
      }
      mbar_wait(&a_empty[s], ph ^ 1u);                      // the MMAs that read this A buffer two items ago are done
      const uint32_t r0 = smem_u32(smem + s * kC1ABytes + row * 128);   // K block 0: [x_hi | x_lo]
      const uint32_t r1 = r0 + kTileBytes;                              // K block 1: [x_hi | (never read)]
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 vh = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        const uint4 vl = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        sts128(r0 + ((uint32_t)(c ^ (row & 7)) << 4), vh);
        sts128(r0 + ((uint32_t)((4 + c) ^ (row & 7)) << 4), vl);
        sts128(r1 + ((uint32_t)(c ^ (row & 7)) << 4), vh);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[s]);
    }
  } else {
    // ===== epilogue (warps 5..8 <-> TMEM lane quadrants 1,2,3,0) =====
    const int q = warp_id & 3;
    const int row = q * 32 + lane;
    const bool issuer = (warp_id == 5 && lane == 0);
    const uint32_t bias_addr = smem_u32(bias_s);
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      int w0, h0, d0; long long n;
      decode(item, w0, h0, d0, n);
      const uint32_t s = it & 1;
      uint8_t* tile = smem_o;                               // single staging tile (two CTAs share an SM)
      mbar_wait(&acc_full[s], (it >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * 64u;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
      // the previous item's store must have read the staging tile before it is overwritten.  (Storing the rows straight
      // from registers instead -- no staging, fences or block barriers -- was measured for the bias-less launches:
      // bit-identical, but the step got 1.3 ms SLOWER; 16-byte pieces at a 128-byte stride write half sectors.)
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const uint32_t tile_row = smem_u32(tile) + (uint32_t)row * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float f[8], g8[8];
#pragma unroll
        // bias: 128-bit broadcast reads (4 per 8 channels instead of 16 scalar ones).  Folding it into a free K slot of
        // the GEMM was tried: faster still, but a bf16 hi + lo bias moves single roundings of the output (0.07 % of
        // the elements), and the fp32 add keeps this kernel bit-identical to the build the parity evidence was taken on.
        // Without a bias (input gradient of the decoder tail, 7 of the 12 launches of a step) the add is skipped: the
        // accumulator is never -0 (the zero K-padding slots contribute +0 products), so acc + 0.0f == acc bit for bit.
        if constexpr (HAS_BIAS) {
          const float4 b0 = lds128_f32(bias_addr + (uint32_t)(c * 8) * 4u), b1 = lds128_f32(bias_addr + (uint32_t)(c * 8 + 4) * 4u);
          const float4 b2 = lds128_f32(bias_addr + (uint32_t)(32 + c * 8) * 4u), b3 = lds128_f32(bias_addr + (uint32_t)(36 + c * 8) * 4u);
          const float bl[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const float bh[8] = {b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(v0[c * 8 + e]) + bl[e];
            g8[e] = __uint_as_float(v1[c * 8 + e]) + bh[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(v0[c * 8 + e]);
            g8[e] = __uint_as_float(v1[c * 8 + e]);
          }
        }
        uint4 pk;
        pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]);
        pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
        sts128(tile_row + ((uint32_t)(c ^ (row & 7)) << 4), pk);
        pk.x = pack_bf16x2(g8[0], g8[1]); pk.y = pack_bf16x2(g8[2], g8[3]);
        pk.z = pack_bf16x2(g8[4], g8[5]); pk.w = pack_bf16x2(g8[6], g8[7]);
        sts128(tile_row + ((uint32_t)((4 + c) ^ (row & 7)) << 4), pk);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (issuer) {
        tma_store_5d(&tmC, tile, 0, w0, h0, d0, (int)n);
        tma_store_commit();
      }
      if (stats_partial != nullptr)   // fused BatchNorm statistics of the stored (bf16-rounded) values
        tile_channel_sums<kHW>(smem_u32(tile), q, lane, W - w0, H - h0, s1a, s1b, s2a, s2b);
    }
    if (stats_partial != nullptr) {
      float* dst = stats_partial + (size_t)(blockIdx.x * 4 + (warp_id - 5)) * 2 * 64;
      *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(s1a, s1b);
      *reinterpret_cast<float2*>(dst + 64 + 2 * lane) = make_float2(s2a, s2b);
    }
    if (issuer) tma_store_wait_read_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 4) tmem_dealloc(tmem_base, 128);
}

// fp32 [64][27] -> bf16 [2][64][64]: K block 0 = [w_hi(32) | w_hi(32)], K block 1 = [w_lo(32) | 0]  (tap-flipped if `flip`)
__global__ void pack_c1_to_c64_weights_kernel(const float* __restrict__ w, int flip, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * 64 * 64) return;
  const int k = i % 64, c = (i / 64) % 64, kb = i / (64 * 64);
  const int tap = k % 32;
  float v = 0.f;
  if (tap < 27) v = w[c * 27 + (flip ? 26 - tap : tap)];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16 o = __float2bfloat16_rn(0.f);
  if (kb == 0) o = hi;
  else if (k < 32) o = lo;
  wp[i] = o;
}

size_t c1_to_c64_workspace_bytes() { return (size_t)2 * 64 * 64 * sizeof(__nv_bfloat16); }

static int c1_to_c64_tc_impl(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W,
                            int flip, void* ws, size_t ws_bytes, float* stats, int* stats_blocks, cudaStream_t st) {
  SIVAE_CHECK(ws && ws_bytes >= c1_to_c64_workspace_bytes(), "c1_to_c64: workspace too small");
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "c1_to_c64: empty tensor");
  pack_c1_to_c64_weights_kernel<<<32, 256, 0, st>>>(w, flip, (__nv_bfloat16*)ws);
  SIVAE_LAUNCH_OK("pack_c1_to_c64_weights_kernel");
  CUtensorMap tmB, tmC;
  {
    uint64_t dims[3] = {64, 64, 2};
    uint64_t strides[2] = {128, 64 * 128};
    uint32_t box[3] = {64, 64, 1};
    if (make_tmap_bf16(&tmB, ws, 3, dims, strides, box)) return -1;
  }
  if (make_act_tmap(&tmC, y, N, D, H, W, 64, kHW, kHH, 1)) return -1;
  const int tiles_w = cdiv(W, kHW), tiles_h = cdiv(H, kHH);
  const long long items = (long long)tiles_w * tiles_h * D * N;
  SIVAE_CHECK(items < (1ll << 31), "c1_to_c64: too many tiles");
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(c1_to_c64_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1Smem),
                   "cudaFuncSetAttribute(c1_to_c64_tc)") ||
        check_cuda(cudaFuncSetAttribute(c1_to_c64_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1Smem),
                   "cudaFuncSetAttribute(c1_to_c64_tc)"))
      return -1;
    attr_set = true;
  }
  // two persistent CTAs per SM (102 KB of shared memory, 288 threads x <= 112 registers, 128 TMEM columns each): every
  // role is a single warp per scheduler, so a second CTA is what hides the LDS / ALU latencies of the builders
  const long long cap = 2ll * num_sms();
  const unsigned ctas = (unsigned)(items < cap ? items : cap);
  if (bias != nullptr)
    c1_to_c64_tc_kernel<true><<<ctas, 288, kC1Smem, st>>>(x1, tmB, tmC, bias, N, D, H, W, tiles_w, tiles_h, (int)items, stats);
  else
    c1_to_c64_tc_kernel<false><<<ctas, 288, kC1Smem, st>>>(x1, tmB, tmC, bias, N, D, H, W, tiles_w, tiles_h, (int)items, stats);
  SIVAE_LAUNCH_OK("c1_to_c64_tc_kernel");
  if (stats != nullptr && stats_blocks != nullptr) *stats_blocks = (int)ctas * 4;
  return 0;
}

int c1_to_c64_tc(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int flip,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  return c1_to_c64_tc_impl(x1, w, bias, y, N, D, H, W, flip, ws, ws_bytes, nullptr, nullptr, st);
}

int bn_coeffs_from_partials(const float* partial, int nblocks, long long nvox, int C, const float* gamma,
                            const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                            float* mean, float* invstd, float* scale, float* shift, cudaStream_t st);
size_t bn_workspace_bytes(int C);

// Encoder stem Conv3d(1,64,3,bias) + the train-mode statistics of the BatchNorm3d behind it (models/models.py:92-93):
// the channel sums come out of the convolution's epilogue.
int c1_to_c64_bn(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int flip,
                 const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                 float* mean, float* invstd, float* scale, float* shift, void* ws_pack, size_t ws_pack_bytes,
                 void* ws_bn, size_t ws_bn_bytes, cudaStream_t st) {
  SIVAE_CHECK(ws_bn && ws_bn_bytes >= bn_workspace_bytes(64), "c1_to_c64_bn: BN workspace too small");
  int blocks = 0;
  int rc = c1_to_c64_tc_impl(x1, w, bias, y, N, D, H, W, flip, ws_pack, ws_pack_bytes, (float*)ws_bn, &blocks, st);
  if (rc) return rc;
  return bn_coeffs_from_partials((const float*)ws_bn, blocks, (long long)N * D * H * W, 64, gamma, beta, rm, rv, nbt,
                                 momentum, eps, mean, invstd, scale, shift, st);
}

// =================================================================================================
// Weight gradient of the thin 3x3x3 convolutions with C = 64 (encoder stem, decoder tail) on tensor cores:
//   dw[c][tap] = sum_v xc[v][c] * x1[v + delta(tap)],   sum_c[c] = sum_v xc[v][c],   sum_1 = sum_v x1[v].
// GEMM D[m][c] += A[v][m] * B[v][c] over voxel rows v (K), both operands MN-major: B = the NDHWC tile of xc straight
// from TMA, A = [x_hi(tap 0..26), 1.0, 0.. | x_lo(tap 0..26), 0..] built in shared memory from the fp32 halo of x1
// (rows 0..26: hi products, row 27: sum_c, rows 32..58: lo products; rows 64..127 of the M=128 MMA are don't-care).
// Persistent CTAs stride over 16x8x1 voxel tiles; fp32 partials per CTA, deterministic finalize.
// =================================================================================================
static constexpr int kWg1Stages = 2;

__global__ void __launch_bounds__(192)
wgrad_c1_tc_kernel(const float* __restrict__ x1, const __grid_constant__ CUtensorMap tmXC, int N, int D, int H, int W,
                   int tiles_w, int tiles_h, long long total_tiles, float* __restrict__ partial) {
  constexpr int A_STAGE = 2 * kTileBytes;   // chunk 0 built, chunk 1 (M rows 64..127) never read back
  constexpr int B_STAGE = kTileBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kWg1Stages * A_STAGE;
  float* xs = reinterpret_cast<float*>(smem_b + kWg1Stages * B_STAGE);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(xs + ((kHRows + 3) & ~3));
  uint64_t* b_full = a_full + kWg1Stages;
  uint64_t* empty = b_full + kWg1Stages;       // stage (A and B) released by the MMA commit
  uint64_t* tmem_full_bar = empty + kWg1Stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* red = reinterpret_cast<float*>(tmem_ptr_smem + 2);   // 4 floats: per-warp sum_1

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 0-3 builders/epilogue, 4 TMA, 5 MMA
  const long long first = blockIdx.x, step = gridDim.x;
  const long long my_tiles = first < total_tiles ? (total_tiles - first + step - 1) / step : 0;

  if (warp_id == 4 && lane == 0) {
    prefetch_tmap(&tmXC);
    for (int s = 0; s < kWg1Stages; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&b_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_id == 5) tmem_alloc(tmem_ptr_smem, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto decode = [&](long long id, int& w0, int& h0, int& d0, long long& n) {
    const int tw = (int)(id % tiles_w); id /= tiles_w;
    const int th = (int)(id % tiles_h); id /= tiles_h;
    d0 = (int)(id % D);
    n = id / D;
    w0 = tw * kHW; h0 = th * kHH;
  };

  if (warp_id == 4) {
    if (elect_one()) {
      for (long long it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % kWg1Stages);
        mbar_wait(&empty[s], (uint32_t)((it / kWg1Stages) & 1) ^ 1u);
        int w0, h0, d0; long long n;
        decode(first + it * step, w0, h0, d0, n);
        mbar_expect_tx(&b_full[s], B_STAGE);
        tma_load_5d(smem_b + s * B_STAGE, &tmXC, &b_full[s], 0, w0, h0, d0, (int)n);
      }
    }
  } else if (warp_id == 5) {
    if (elect_one()) {                       // one elected thread issues the whole loop
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), kTileBytes, 1024);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), kTileBytes, 1024);
      uint32_t s = 0, ph = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        mbar_wait(&a_full[s], ph);
        mbar_wait(&b_full[s], ph);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(s * (uint32_t)(A_STAGE >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)(s * (uint32_t)(B_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 voxel rows (2048 B = +128 in the descriptor's address field) per UMMA
          umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (it == my_tiles - 1) umma_commit(tmem_full_bar);
        if (++s == kWg1Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int row = threadIdx.x;                          // voxel row of the tile
    const int ow = row % kHW, oh = row / kHW;
    float s1 = 0.f;
    const uint32_t xs_addr = smem_u32(xs);                // shared-window address of the halo: LDS / STS below
    constexpr int kPer = (kHRows + 127) / 128;            // halo floats per thread (5)
    float pre[kPer];
    // the fp32 halo of the NEXT tile is fetched into registers while the current A tile is built, so the global-load
    // latency is off the per-tile critical path
    // this thread's halo cells never change: offsets from the halo origin and packed (wx, hy, dz) for the bounds tests
    int off[kPer], cell[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = row + j * 128;
      const int wx = i % kHBW, hy = (i / kHBW) % kHBH, dz = i / (kHBW * kHBH);
      off[j] = (dz * H + hy) * W + wx;
      cell[j] = i < kHRows ? (wx | (hy << 8) | (dz << 16)) : -1;
    }
    auto fetch = [&](long long id) {
      int w0, h0, d0; long long n;
      decode(id, w0, h0, d0, n);
      const float* origin = x1 + (((n * D + (d0 - 1)) * H + (h0 - 1)) * W + (w0 - 1));
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int d = d0 - 1 + (cell[j] >> 16), h = h0 - 1 + ((cell[j] >> 8) & 255), w = w0 - 1 + (cell[j] & 255);
        float v = 0.f;
        if (cell[j] >= 0 && (unsigned)d < (unsigned)D && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
          v = __ldg(origin + off[j]);
        pre[j] = v;
      }
    };
    if (my_tiles > 0) fetch(first);
    for (long long it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % kWg1Stages);
      asm volatile("bar.sync 1, 128;" ::: "memory");      // previous iteration's halo reads are done
#pragma unroll
      for (int j = 0; j < kPer; ++j)
        if (row + j * 128 < kHRows) sts_f32(xs_addr + (uint32_t)(row + j * 128) * 4u, pre[j]);
      if (it + 1 < my_tiles) fetch(first + (it + 1) * step);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = 0.f, b = 0.f;
        if (2 * j < 27) a = lds_f32(xs_addr + (uint32_t)((((2 * j) / 9) * kHBH + oh + ((2 * j) / 3) % 3) * kHBW + ow + (2 * j) % 3) * 4u);
        if (2 * j + 1 < 27)
          b = lds_f32(xs_addr + (uint32_t)((((2 * j + 1) / 9) * kHBH + oh + ((2 * j + 1) / 3) % 3) * kHBW + ow + (2 * j + 1) % 3) * 4u);
        if (2 * j + 1 == 27) b = 1.0f;                     // row 27 of D accumulates sum_v xc[v][c]
        hi[j] = pack_bf16x2(a, b);
        lo[j] = pack_bf16x2(a - __uint_as_float(hi[j] << 16), b - __uint_as_float(hi[j] & 0xffff0000u));
      }
      s1 += lds_f32(xs_addr + (uint32_t)((1 * kHBH + oh + 1) * kHBW + ow + 1) * 4u);   // centre tap; zero outside the volume
      mbar_wait(&empty[s], (uint32_t)((it / kWg1Stages) & 1) ^ 1u);
      const uint32_t r0 = smem_u32(smem_a + s * A_STAGE + row * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts128(r0 + ((uint32_t)(c ^ (row & 7)) << 4), make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]));
        sts128(r0 + ((uint32_t)((4 + c) ^ (row & 7)) << 4), make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]));
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[s]);
    }
    // ---- epilogue: rows 0..63 of D -> partial[cta][m][c]; per-CTA sum_1 ----
    s1 = warp_sum(s1);
    if (lane == 0) red[warp_id] = s1;
    float* dst = partial + (long long)blockIdx.x * (64 * 64 + 4);   // 16-byte aligned stride
    if (my_tiles > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      uint32_t v[32];
      if (my_tiles > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp_id * 32) << 16) + (uint32_t)(j * 32), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
      if (warp_id < 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dst + row * 64 + j * 32 + c * 4) = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) dst[64 * 64] = red[0] + red[1] + red[2] + red[3];
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 5) tmem_dealloc(tmem_base, 64);
}

// one warp per output element: lanes stride over the per-CTA partials (fp64, fixed order -> deterministic)
__global__ void wgrad_c1_tc_finalize_kernel(const float* __restrict__ partial, int nctas, int flip, float* dw,
                                            float* sum_c, float* sum_1) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // 0 .. 27*64 (dw), then 64 (sum_c), then 1
  const int lane = threadIdx.x & 31;
  const int per = 64 * 64 + 4;
  if (i > 27 * 64 + 64) return;
  double a = 0.0;
  if (i < 27 * 64) {
    const int t = i / 64, c = i % 64;
    const int m = flip ? 26 - t : t;
    for (int b = lane; b < nctas; b += 32)
      a += (double)partial[(long long)b * per + m * 64 + c] + (double)partial[(long long)b * per + (32 + m) * 64 + c];
  } else if (i < 27 * 64 + 64) {
    for (int b = lane; b < nctas; b += 32) a += (double)partial[(long long)b * per + 27 * 64 + (i - 27 * 64)];
  } else {
    for (int b = lane; b < nctas; b += 32) a += (double)partial[(long long)b * per + 64 * 64];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane != 0) return;
  if (i < 27 * 64) dw[(i % 64) * 27 + i / 64] = (float)a;
  else if (i < 27 * 64 + 64) { if (sum_c) sum_c[i - 27 * 64] = (float)a; }
  else if (sum_1) sum_1[0] = (float)a;
}

static constexpr int kWg1Ctas = 148 * 2;
size_t wgrad_c1_tc_workspace_bytes() { return (size_t)kWg1Ctas * (64 * 64 + 4) * sizeof(float); }

int wgrad_c1_tc(const void* xc, const float* x1, float* dw, float* sum_c, float* sum_1, int N, int D, int H, int W,
                int flip, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(ws && ws_bytes >= wgrad_c1_tc_workspace_bytes(), "wgrad_c1_tc: workspace too small");
  SIVAE_CHECK(N > 0 && D > 0 && H > 0 && W > 0, "wgrad_c1_tc: empty tensor");
  CUtensorMap tmXC;
  if (make_act_tmap(&tmXC, xc, N, D, H, W, 64, kHW, kHH, 1)) return -1;
  const int tiles_w = cdiv(W, kHW), tiles_h = cdiv(H, kHH);
  const long long total = (long long)tiles_w * tiles_h * D * N;
  const int ctas = (int)(total < kWg1Ctas ? total : kWg1Ctas);
  constexpr int smem = kWg1Stages * (2 * kTileBytes + kTileBytes) + ((kHRows + 3) & ~3) * 4 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(wgrad_c1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                   "cudaFuncSetAttribute(wgrad_c1_tc)"))
      return -1;
    attr_set = true;
  }
  wgrad_c1_tc_kernel<<<ctas, 192, smem, st>>>(x1, tmXC, N, D, H, W, tiles_w, tiles_h, total, (float*)ws);
  SIVAE_LAUNCH_OK("wgrad_c1_tc_kernel");
  wgrad_c1_tc_finalize_kernel<<<cdiv(27 * 64 + 64 + 1, 8), 256, 0, st>>>((const float*)ws, ctas, flip, dw, sum_c, sum_1);
  SIVAE_LAUNCH_OK("wgrad_c1_tc_finalize_kernel");
  return 0;
}

// fp32 [C][27] (one output channel) -> bf16 [27][16][C]; the filter (tap-flipped when `flip`) is split into
// row 0 = bf16(w) and row 1 = bf16(w - bf16(w)) so the two accumulator columns add up to (almost) fp32 weights
// at no extra MMA cost (the N = 16 tile is the minimum anyway); rows 2..15 are zero.
__global__ void pack_to1_weights_kernel(const float* __restrict__ w, int C, int flip, __nv_bfloat16* __restrict__ wp) {
  const int total = 27 * 16 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, r = (i / C) % 16, tap = i / (16 * C);
    const float v = w[c * 27 + (flip ? 26 - tap : tap)];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    wp[i] = r == 0 ? hi : r == 1 ? __float2bfloat16_rn(v - __bfloat162float(hi)) : __float2bfloat16_rn(0.f);
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
  if (C == 64 && getenv("SIVAE_TO1_TAPWISE") == nullptr) {
    // halo variant: the input tile is fetched once, all 27 taps are GEMM columns
    pack_to1_halo_weights_kernel<<<16, 256, 0, st>>>(w, flip, (__nv_bfloat16*)ws);
    SIVAE_LAUNCH_OK("pack_to1_halo_weights_kernel");
    CUtensorMap tmA, tmB;
    if (make_act_tmap(&tmA, x, N, D, H, W, 64, kHBW, kHBH, 1)) return -1;
    {
      uint64_t dims[3] = {64, 32, 2};
      uint64_t strides[2] = {128, 32 * 128};
      uint32_t box[3] = {64, 32, 1};
      if (make_tmap_bf16(&tmB, ws, 3, dims, strides, box)) return -1;
    }
    ToOneEpilogue ep;
    ep.y = y; ep.bias = bias; ep.mask = mask; ep.seed = make_seed_ref(seed); ep.p = p; ep.act = act;
    To1Geom tg;
    tg.N = N; tg.D = D; tg.H = H; tg.W = W;
    tg.tiles_w = cdiv(W, kHW); tg.tiles_h = cdiv(H, kHH);
    // depth chunks: ~20 planes each (2 extra input planes per chunk = 10 % overhead), more chunks when the patches
    // alone cannot fill the SMs
    const long long patches = (long long)tg.tiles_w * tg.tiles_h * N;
    int chunks = cdiv(D, 20);
    while (patches * chunks < 2ll * num_sms() && chunks < D) ++chunks;
    tg.dc = cdiv(D, chunks);
    tg.chunks = cdiv(D, tg.dc);
    SIVAE_CHECK(patches * tg.chunks < (1ll << 31), "conv3_to1: too many tiles");
    tg.items = (int)(patches * tg.chunks);
    static bool attr_set = false;
    if (!attr_set) {
      if (check_cuda(cudaFuncSetAttribute(conv3_to1_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTo1Smem),
                     "cudaFuncSetAttribute(conv3_to1_halo)"))
        return -1;
      attr_set = true;
    }
    const unsigned ctas = (unsigned)(tg.items < num_sms() ? tg.items : num_sms());
    conv3_to1_halo_kernel<<<ctas, kTo1Threads, kTo1Smem, st>>>(tmA, tmB, tg, ep);
    SIVAE_LAUNCH_OK("conv3_to1_halo_kernel");
    return 0;
  }
  pack_to1_weights_kernel<<<cdiv(27 * 16 * C, 256), 256, 0, st>>>(w, C, flip, (__nv_bfloat16*)ws);
  SIVAE_LAUNCH_OK("pack_to1_weights_kernel");
  ConvGeom g;
  fill_geom(g, N, D, H, W, C, kTapsPlain);
  const long long tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  SIVAE_CHECK(tiles < (1ll << 31), "conv3_to1: too many tiles");
  TmapPack tmA;
  CUtensorMap tmB;
  if (make_act_tmap(&tmA.m[0], x, N, D, H, W, C, g.wt, g.ht, g.dt)) return -1;
  for (int i = 1; i < 8; ++i) tmA.m[i] = tmA.m[0];
  if (make_weight_tmap(&tmB, ws, 27, 16, C, 16)) return -1;
  ToOneEpilogue ep;
  ep.y = y; ep.bias = bias; ep.mask = mask; ep.seed = make_seed_ref(seed); ep.p = p; ep.act = act;
  return launch_igemm<16, 4, 1>(tmA, tmB, tmA, g, tiles, 1, ep, st);
}

// ---- wgrad ----
static void wgrad_plan(int N, int D, int H, int W, int Cin, int Cout, int mode, WgradGeom& g, int& nt, int& ntiles,
                       int& splits) {
  g.N = N; g.D = D; g.H = H; g.W = W;
  pick_tile(W, H, D, 1, true, g.wt, g.ht, g.dt);
  g.tiles_w = cdiv(W, g.wt); g.tiles_h = cdiv(H, g.ht); g.tiles_d = cdiv(D, g.dt);
  g.rows = g.wt * g.ht * g.dt;
  g.cin_blocks = Cin / 64;
  g.mode = mode;
  g.nclasses = mode == kTapsPlain ? 1 : 8;
  g.units_pc = (mode == kTapsPlain ? 27 : 8) * g.cin_blocks;
  g.pairs_pc = (g.units_pc + 1) / 2;
  g.units = g.nclasses * g.units_pc;
  nt = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  ntiles = Cout / nt;
  g.ppc = 512 / nt;
  g.groups_pc = cdiv(g.pairs_pc, g.ppc);
  g.Cout = Cout;
  g.total_kb = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  int want = 148 / (g.groups_pc * g.nclasses * ntiles);
  if (want < 1) want = 1;
  if ((long long)want > g.total_kb) want = (int)g.total_kb;
  g.kb_per_split = (g.total_kb + want - 1) / want;
  splits = (int)((g.total_kb + g.kb_per_split - 1) / g.kb_per_split);
  g.a_lbo = kTileBytes; g.a_sbo = 1024; g.b_lbo = kTileBytes; g.b_sbo = 1024;
}

// Plan of the persistent kw-slab wgrad kernel; returns false when the generic kernel should run instead
// (SIVAE_WGRAD_KW=0 disables it, =force takes it for every Cin = 64 plain-mode shape).
static bool wgrad_kw_plan(int N, int D, int H, int W, int Cin, int Cout, int mode, WgKwGeom& g, int& splits) {
  if (mode != kTapsPlain || Cin != 64 || Cout % 64 != 0) return false;
  const char* e = getenv("SIVAE_WGRAD_KW");
  if (e != nullptr && e[0] == '0') return false;
  const bool force = e != nullptr && e[0] == 'f';
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.tiles_w = cdiv(W, kKwW); g.tiles_h = cdiv(H, kKwH); g.tiles_d = cdiv(D, 2);
  g.ntiles = Cout / 64;
  g.Cout = Cout;
  g.items = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  const double eff = ((double)W * H * D) / ((double)g.tiles_w * kKwW * g.tiles_h * kKwH * g.tiles_d * 2);
  int want = num_sms() / (3 * g.ntiles);
  if (want < 1) want = 1;
  if ((long long)want > g.items) want = (int)g.items;
  g.items_per_split = (g.items + want - 1) / want;
  splits = (int)((g.items + g.items_per_split - 1) / g.items_per_split);
  return force || (eff >= 0.8 && g.items_per_split >= 8);
}

// Plan of the persistent tall-box upconv wgrad kernel (SIVAE_UPWGRAD_TALL=0 disables it, =force takes every Cin = 64 shape)
static bool wgrad_up_plan(int N, int D, int H, int W, int Cin, int Cout, int mode, WgUpGeom& g, int& splits) {
  if (mode != kTapsUpFprop || Cin != 64 || Cout % 64 != 0) return false;
  const char* e = getenv("SIVAE_UPWGRAD_TALL");
  if (e != nullptr && e[0] == '0') return false;
  const bool force = e != nullptr && e[0] == 'f';
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.tiles_w = cdiv(W, kKwW); g.tiles_h = cdiv(H, kKwH); g.tiles_d = cdiv(D, 2);
  g.ntiles = Cout / 64;
  g.Cout = Cout;
  g.items = (long long)g.tiles_w * g.tiles_h * g.tiles_d * N;
  const double eff = ((double)W * H * D) / ((double)g.tiles_w * kKwW * g.tiles_h * kKwH * g.tiles_d * 2);
  int want = num_sms() / (8 * g.ntiles);
  if (want < 1) want = 1;
  if ((long long)want > g.items) want = (int)g.items;
  g.items_per_split = (g.items + want - 1) / want;
  splits = (int)((g.items + g.items_per_split - 1) / g.items_per_split);
  return force || (eff >= 0.8 && g.items_per_split >= 8);
}

static size_t wgrad_ws_bytes(int N, int D, int H, int W, int Cin, int Cout, int mode) {
  if (Cin % 64 || Cout % 64 || N <= 0) return 0;
  WgradGeom g; int nt, ntiles, splits;
  wgrad_plan(N, D, H, W, Cin, Cout, mode, g, nt, ntiles, splits);
  size_t need = (size_t)splits * g.units * 64 * Cout * sizeof(float);
  WgKwGeom kg; int ksplits;
  if (wgrad_kw_plan(N, D, H, W, Cin, Cout, mode, kg, ksplits)) {
    const size_t kneed = (size_t)ksplits * 27 * 64 * Cout * sizeof(float);
    if (kneed > need) need = kneed;
  }
  WgUpGeom ug; int usplits;
  if (wgrad_up_plan(N, D, H, W, Cin, Cout, mode, ug, usplits)) {
    const size_t uneed = (size_t)usplits * 64 * 64 * Cout * sizeof(float);
    if (uneed > need) need = uneed;
  }
  return need;
}
size_t conv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  return wgrad_ws_bytes(N, D, H, W, Cin, Cout, kTapsPlain);
}
size_t upconv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  return wgrad_ws_bytes(N, D, H, W, Cin, Cout, kTapsUpFprop);
}

template <int NT, int A_STAGES>
static int launch_wgrad(const CUtensorMap& tmX, const TmapPack& tmDY, const WgradGeom& g, int ntiles, int splits,
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
  dim3 grid((unsigned)(g.groups_pc * g.nclasses * ntiles), (unsigned)splits);
  launch_conv(conv3_wgrad_kernel<NT, A_STAGES>, grid, 192, smem, st, tmX, tmDY, g, partial);
  SIVAE_LAUNCH_OK("conv3_wgrad_kernel");
  return 0;
}

// mode kTapsPlain : x, dy on the same [N][D][H][W] lattice.
// mode kTapsUpFprop: x is the LOW-res input [N][D][H][W][Cin], dy the HIGH-res output gradient [N][2D][2H][2W][Cout].
static int wgrad_impl(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, int N, int D, int H, int W,
                      int Cin, int Cout, int mode, cudaStream_t st) {
  SIVAE_CHECK(Cin % 64 == 0 && Cin >= 64 && Cout % 64 == 0 && Cout >= 64,
              "conv3_wgrad: Cin=%d, Cout=%d must be multiples of 64", Cin, Cout);
  {
    WgKwGeom kg; int ksplits;
    if (wgrad_kw_plan(N, D, H, W, Cin, Cout, mode, kg, ksplits)) {
      const size_t kneed = (size_t)ksplits * 27 * 64 * Cout * sizeof(float);
      SIVAE_CHECK(ws != nullptr && ws_bytes >= kneed, "conv3_wgrad: workspace too small (%zu < %zu)", ws_bytes, kneed);
      CUtensorMap tX, tDY;
      {
        uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        uint64_t strides[4] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128, (uint64_t)D * H * W * 128};
        uint32_t box[5] = {64, (uint32_t)kKwW, (uint32_t)(kKwH + 2), 4, 1};
        if (make_tmap_bf16(&tX, x, 5, dims, strides, box)) return -1;
      }
      if (make_act_tmap(&tDY, dy, N, D, H, W, Cout, kKwW, kKwH, 1)) return -1;
      static bool attr_set = false;
      if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(conv3_wgrad_kw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kWgKwSmem), "cudaFuncSetAttribute(conv3_wgrad_kw64)")) return -1;
        attr_set = true;
      }
      launch_conv(conv3_wgrad_kw64_kernel, dim3(3u * (unsigned)kg.ntiles, (unsigned)ksplits), 224, kWgKwSmem, st, tX, tDY, kg,
                  (float*)ws);
      SIVAE_LAUNCH_OK("conv3_wgrad_kw64_kernel");
      const long long total = 27ll * 64 * Cout;
      int blocks = (int)((total + 255) / 256);
      wgrad_reduce_kernel<<<blocks, 256, 0, st>>>((const float*)ws, dw, ksplits, 27, 1, Cin, Cout);
      SIVAE_LAUNCH_OK("wgrad_reduce_kernel");
      return 0;
    }
  }
  {
    WgUpGeom ug; int usplits;
    if (wgrad_up_plan(N, D, H, W, Cin, Cout, mode, ug, usplits)) {
      const size_t uneed = (size_t)usplits * 64 * 64 * Cout * sizeof(float);
      SIVAE_CHECK(ws != nullptr && ws_bytes >= uneed, "upconv3_wgrad: workspace too small (%zu < %zu)", ws_bytes, uneed);
      CUtensorMap tX;
      TmapPack tDY;
      {
        uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        uint64_t strides[4] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128, (uint64_t)D * H * W * 128};
        uint32_t box[5] = {64, (uint32_t)kKwW, (uint32_t)(kKwH + 2), 4, 1};
        if (make_tmap_bf16(&tX, x, 5, dims, strides, box)) return -1;
      }
      for (int p = 0; p < 8; ++p)
        if (make_parity_tmap(&tDY.m[p], dy, N, D, H, W, Cout, p, kKwW, kKwH, 1)) return -1;
      static bool attr_set = false;
      if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(upconv3_wgrad_tall_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kWgKwSmem), "cudaFuncSetAttribute(upconv3_wgrad_tall)")) return -1;
        attr_set = true;
      }
      launch_conv(upconv3_wgrad_tall_kernel, dim3(8u * (unsigned)ug.ntiles, (unsigned)usplits), 224, kWgKwSmem, st, tX, tDY, ug,
                  (float*)ws);
      SIVAE_LAUNCH_OK("upconv3_wgrad_tall_kernel");
      const long long total = 27ll * Cin * Cout;
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      wgrad_reduce_up_kernel<<<blocks, 256, 0, st>>>((const float*)ws, dw, usplits, 1, Cin, Cout);
      SIVAE_LAUNCH_OK("wgrad_reduce_up_kernel");
      return 0;
    }
  }
  WgradGeom g; int nt, ntiles, splits;
  wgrad_plan(N, D, H, W, Cin, Cout, mode, g, nt, ntiles, splits);
  const size_t need = (size_t)splits * g.units * 64 * Cout * sizeof(float);
  SIVAE_CHECK(ws != nullptr && ws_bytes >= need, "conv3_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  // debugging knob: override the MN-major descriptor strides "a_lbo,a_sbo,b_lbo,b_sbo"
  if (const char* e = getenv("SIVAE_WGRAD_DESC")) {
    unsigned a, b, c, d;
    if (sscanf(e, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) { g.a_lbo = a; g.a_sbo = b; g.b_lbo = c; g.b_sbo = d; }
  }
  CUtensorMap tmX;
  TmapPack tmDY;
  if (make_act_tmap(&tmX, x, N, D, H, W, Cin, g.wt, g.ht, g.dt)) return -1;
  if (mode == kTapsPlain) {
    if (make_act_tmap(&tmDY.m[0], dy, N, D, H, W, Cout, g.wt, g.ht, g.dt)) return -1;
    for (int i = 1; i < 8; ++i) tmDY.m[i] = tmDY.m[0];
  } else {
    for (int p = 0; p < 8; ++p)
      if (make_parity_tmap(&tmDY.m[p], dy, N, D, H, W, Cout, p, g.wt, g.ht, g.dt)) return -1;
  }
  int rc;
  if (nt == 256) rc = launch_wgrad<256, 2>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  else if (nt == 128) rc = launch_wgrad<128, 3>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  else rc = launch_wgrad<64, 4>(tmX, tmDY, g, ntiles, splits, (float*)ws, st);
  if (rc) return rc;
  if (mode == kTapsPlain) {
    const long long total = (long long)g.units * 64 * Cout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>((const float*)ws, dw, splits, g.units, g.cin_blocks, Cin, Cout);
    SIVAE_LAUNCH_OK("wgrad_reduce_kernel");
  } else {
    const long long total = 27ll * Cin * Cout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_up_kernel<<<blocks, 256, 0, st>>>((const float*)ws, dw, splits, g.cin_blocks, Cin, Cout);
    SIVAE_LAUNCH_OK("wgrad_reduce_up_kernel");
  }
  return 0;
}

int conv3_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, int N, int D, int H, int W,
                int Cin, int Cout, cudaStream_t st) {
  return wgrad_impl(x, dy, dw, ws, ws_bytes, N, D, H, W, Cin, Cout, kTapsPlain, st);
}
int upconv3_wgrad(const void* x_lo, const void* dy_hi, float* dw, void* ws, size_t ws_bytes, int N, int D, int H, int W,
                  int Cin, int Cout, cudaStream_t st) {
  return wgrad_impl(x_lo, dy_hi, dw, ws, ws_bytes, N, D, H, W, Cin, Cout, kTapsUpFprop, st);
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
