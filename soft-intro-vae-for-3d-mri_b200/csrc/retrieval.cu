// retrieval.cu -- top-k similarity search over latent vectors (NEXT-4).
//
// The reference's stated end goal is content-based image retrieval on the encoder's latents (README.md:4-11); the
// repository itself only extracts latents (logistic1.ipynb cell 7) and analyses them with sklearn on the host.  This
// file provides the missing device-side search: for every query vector the k most similar database vectors under
// cosine similarity or (negative squared) L2 distance.
//   scores[q][d] = <Q_q, D_d>                          tiled fp32 SGEMM (64 x 64 tiles, exact fp32 FMA accumulation)
//   cosine: scores / (|Q_q| |D_d|)     l2: -(|Q_q|^2 + |D_d|^2 - 2 <Q_q, D_d>)
//   top-k per query row: one warp per row, per-lane sorted candidate lists, k rounds of warp arg-max
// Sizes in scope: 10^4 x 10^4 vectors of dimension 1200 (240 GFLOP, 400 MB of scores): not a hot path, kept simple.
#include "sivae_common.cuh"

namespace sivae {

static constexpr int kSimTile = 64, kSimK = 16, kTopKMax = 32;

__global__ void row_sqnorm_kernel(const float* __restrict__ x, int n, int dim, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float s = 0.f;
  for (int j = threadIdx.x & 31; j < dim; j += 32) {
    const float v = x[(size_t)row * dim + j];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[row] = s;
}

// scores[q][d] for a 64 x 64 tile per block; metric 0 = cosine, 1 = negative squared L2
__global__ void __launch_bounds__(256)
sim_scores_kernel(const float* __restrict__ Q, const float* __restrict__ Dm, int nq, int nd, int dim,
                  const float* __restrict__ qn, const float* __restrict__ dn, int metric, float* __restrict__ S) {
  __shared__ float sq[kSimK][kSimTile + 1], sd[kSimK][kSimTile + 1];
  const int q0 = blockIdx.y * kSimTile, d0 = blockIdx.x * kSimTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < dim; k0 += kSimK) {
    for (int i = threadIdx.x; i < kSimTile * kSimK; i += 256) {
      const int r = i / kSimK, c = i % kSimK;                 // consecutive threads read consecutive k: coalesced
      const int k = k0 + c;
      sq[c][r] = (q0 + r < nq && k < dim) ? Q[(size_t)(q0 + r) * dim + k] : 0.f;
      sd[c][r] = (d0 + r < nd && k < dim) ? Dm[(size_t)(d0 + r) * dim + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kSimK; ++c) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sq[c][ty * 4 + i]; b[i] = sd[c][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + tx * 4 + j;
      if (d >= nd) continue;
      float v = acc[i][j];
      if (metric == 0) v = v * rsqrtf(fmaxf(qn[q] * dn[d], 1e-30f));
      else v = -(qn[q] + dn[d] - 2.f * v);
      S[(size_t)q * nd + d] = v;
    }
  }
}

// one warp per query row: every lane keeps the best k of its strided share (descending, ties -> lower index first),
// then k rounds of warp arg-max pop the global winners.  Deterministic.
__global__ void topk_rows_kernel(const float* __restrict__ S, int nq, int nd, int k, float* __restrict__ out_s,
                                 int* __restrict__ out_i) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  float bs[kTopKMax];
  int bi[kTopKMax];
#pragma unroll
  for (int t = 0; t < kTopKMax; ++t) { bs[t] = -INFINITY; bi[t] = 0x7fffffff; }
  for (int d = lane; d < nd; d += 32) {
    float v = S[(size_t)row * nd + d];
    int id = d;
    if (v > bs[k - 1] || (v == bs[k - 1] && id < bi[k - 1])) {
      // insertion into the sorted list (k <= 32, fully unrolled compare-exchange chain)
#pragma unroll
      for (int t = 0; t < kTopKMax; ++t) {
        if (t < k && (v > bs[t] || (v == bs[t] && id < bi[t]))) {
          const float tv = bs[t]; const int ti = bi[t];
          bs[t] = v; bi[t] = id; v = tv; id = ti;
        }
      }
    }
  }
  int head = 0;
  for (int r = 0; r < k; ++r) {
    float v = -INFINITY; int id = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < kTopKMax; ++t) if (t == head) { v = bs[t]; id = bi[t]; }
    float wv = v; int wi = id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
    }
    if (wi == id && wv == v && id != 0x7fffffff) ++head;       // this lane owned the winner: pop it
    if (lane == 0) { out_s[(size_t)row * k + r] = wv; out_i[(size_t)row * k + r] = wi == 0x7fffffff ? -1 : wi; }
  }
}

size_t similarity_workspace_bytes(int nq, int nd) {
  if (nq <= 0 || nd <= 0) return 0;
  return ((size_t)nq * nd + (size_t)nq + (size_t)nd) * sizeof(float);
}

int similarity_topk(const float* q, const float* db, int nq, int nd, int dim, int metric, int k, float* out_scores,
                    int* out_index, void* ws, size_t ws_bytes, cudaStream_t st) {
  SIVAE_CHECK(nq > 0 && nd > 0 && dim > 0, "similarity_topk: empty input");
  SIVAE_CHECK(metric == 0 || metric == 1, "similarity_topk: metric must be 0 (cosine) or 1 (l2)");
  SIVAE_CHECK(k >= 1 && k <= kTopKMax && k <= nd, "similarity_topk: k=%d out of range (1..min(%d, nd))", k, kTopKMax);
  SIVAE_CHECK(ws != nullptr && ws_bytes >= similarity_workspace_bytes(nq, nd), "similarity_topk: workspace too small");
  float* S = (float*)ws;
  float* qn = S + (size_t)nq * nd;
  float* dn = qn + nq;
  row_sqnorm_kernel<<<cdiv(nq, 8), 256, 0, st>>>(q, nq, dim, qn);
  SIVAE_LAUNCH_OK("row_sqnorm_kernel");
  row_sqnorm_kernel<<<cdiv(nd, 8), 256, 0, st>>>(db, nd, dim, dn);
  SIVAE_LAUNCH_OK("row_sqnorm_kernel");
  sim_scores_kernel<<<dim3((unsigned)cdiv(nd, kSimTile), (unsigned)cdiv(nq, kSimTile)), 256, 0, st>>>(q, db, nq, nd, dim,
                                                                                                     qn, dn, metric, S);
  SIVAE_LAUNCH_OK("sim_scores_kernel");
  topk_rows_kernel<<<cdiv(nq, 4), 128, 0, st>>>(S, nq, nd, k, out_scores, out_index);
  SIVAE_LAUNCH_OK("topk_rows_kernel");
  return 0;
}

}  // namespace sivae
